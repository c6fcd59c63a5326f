"""Minimal C-header reader for the cuDecomp public ABI: enumerators with their values, struct members in order with
their types and array bounds, #define'd magics/versions and function prototypes reduced to (return type, parameter
types). The same reader is applied to the reference's include/cudecomp.h (by make_golden.py, in the build container)
and to this repo's include/cudecomp.h (by tests/test_abi_golden.py); parameter NAMES are dropped, everything else
must match token for token."""
import re


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    return src


def norm(s):
    s = re.sub(r"\s+", " ", s.strip())
    s = re.sub(r"\s*\*\s*", "* ", s).strip()
    return s


def split_decl(decl):
    """'const int32_t input_halo_extents[]' -> ('const int32_t[]'), 'cudecompHandle_t* handle' -> 'cudecompHandle_t*'"""
    decl = norm(decl)
    m = re.match(r"^(.*?)([A-Za-z_]\w*)?((?:\s*\[[^\]]*\])*)$", decl)
    base, name, arr = m.group(1), m.group(2), m.group(3) or ""
    if name and not base.strip():  # a lone type name such as 'void'
        base, name = name, None
    return norm(base) + re.sub(r"\s+", "", arr)


def parse_header(src):
    src = strip_comments(src)
    out = dict(enums={}, structs={}, defines={}, functions={}, inline_wrappers={})
    for body, name in re.findall(r"typedef\s+enum\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        vals, cur = [], -1
        for item in [i.strip() for i in body.split(",") if i.strip()]:
            if "=" in item:
                k, v = [x.strip() for x in item.split("=")]
                cur = int(v, 0)
            else:
                k, cur = item, cur + 1
            vals.append([k, cur])
        out["enums"][name] = vals
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        members = []
        for decl in [d.strip() for d in body.split(";") if d.strip()]:
            m = re.match(r"^(.*?)(\w+)((?:\s*\[[^\]]*\])*)$", norm(decl))
            members.append([norm(m.group(1)), m.group(2), re.sub(r"\s+", "", m.group(3) or "")])
        out["structs"][name] = members
    for name, val in re.findall(r"#define\s+(CUDECOMP_\w+)\s+(?:INT32_C\()?\s*(0x[0-9a-fA-F]+|\d+)\)?\s*$", src, flags=re.M):
        out["defines"][name] = int(val, 0)
    # prototypes: 'type name(args);' at file scope -- inline wrappers have a body instead of ';'
    for ret, name, args, tail in re.findall(r"(?:^|\n)\s*((?:static\s+inline\s+)?(?:const\s+)?\w+\s*\**)\s*(cudecomp\w+)\s*\(([^)]*)\)\s*(;|\{)",
                                            src):
        params = [split_decl(a) for a in args.split(",")] if args.strip() and args.strip() != "void" else []
        entry = dict(ret=norm(ret.replace("static inline", "")), params=params)
        (out["functions"] if tail == ";" else out["inline_wrappers"])[name] = entry
    return out


def fortran_bindings(src):
    """names the Fortran module binds with bind(C, name="...")"""
    return sorted(set(n for n in re.findall(r'bind\s*\(\s*c\s*,\s*name\s*=\s*"(\w+)"\s*\)', src, flags=re.I) if n.startswith("cudecomp")))
