"""The GPU-only Python paths (bench.py's native arm, the benchmark callers, the GPU worker, smoke()) cannot run in the
CPU suite, so a typo in them would first show up on the GPU box. This test walks their ASTs and reports every name that
is read somewhere but bound nowhere (no assignment, argument, import, def, comprehension target, global or builtin)."""
import ast
import builtins
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["bench.py", "__graft_entry__.py", "bench/fft_benchmark.py", "bench/halo_benchmark.py", "bench/nccl_restated.py",
         "tests/_worker.py", "tests/_launcher.py", "cudecomp_b200/capi.py", "scripts/autotune_bench.py",
         "scripts/r2_autotune_check.py", "scripts/r2_summarize.py", "scripts/explain_plan.py"]


class Scope:
    def __init__(self, parent=None):
        self.parent = parent
        self.bound = set()

    def has(self, name):
        s = self
        while s is not None:
            if name in s.bound:
                return True
            s = s.parent
        return False


def bind_targets(node, scope):
    for n in ast.walk(node):
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            scope.bound.add(n.id)


def collect_bindings(body, scope):
    """Everything a block binds in ITS scope (not descending into nested function / class bodies)."""
    stack = list(body)
    while stack:
        n = stack.pop()
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            scope.bound.add(n.name)
            continue
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                scope.bound.add((a.asname or a.name).split(".")[0])
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            scope.bound.update(n.names)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            scope.bound.add(n.name)
        elif isinstance(n, ast.Lambda):
            continue
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            scope.bound.add(n.id)
        for child in ast.iter_child_nodes(n):
            if isinstance(child, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
                continue  # own scope, handled in check()
            stack.append(child)


def check(node, scope, problems, filename):
    if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
        inner = Scope(scope)
        args = node.args
        for a in args.posonlyargs + args.args + args.kwonlyargs + ([args.vararg] if args.vararg else []) + \
                ([args.kwarg] if args.kwarg else []):
            inner.bound.add(a.arg)
        for d in args.defaults + [d for d in args.kw_defaults if d is not None]:
            check(d, scope, problems, filename)
        body = node.body if isinstance(node.body, list) else [node.body]
        collect_bindings(body, inner)
        for b in body:
            check(b, inner, problems, filename)
        return
    if isinstance(node, ast.ClassDef):
        inner = Scope(scope)
        collect_bindings(node.body, inner)
        for b in node.body:
            check(b, inner, problems, filename)
        return
    if isinstance(node, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
        inner = Scope(scope)
        for gen in node.generators:
            bind_targets(gen.target, inner)
        for gen in node.generators:
            check(gen.iter, inner, problems, filename)
            for cond in gen.ifs:
                check(cond, inner, problems, filename)
        for part in ([node.key, node.value] if isinstance(node, ast.DictComp) else [node.elt]):
            check(part, inner, problems, filename)
        return
    if isinstance(node, ast.Name) and isinstance(node.ctx, ast.Load):
        if not scope.has(node.id) and not hasattr(builtins, node.id):
            problems.append("%s:%d: name '%s' is never bound" % (filename, node.lineno, node.id))
        return
    for child in ast.iter_child_nodes(node):
        check(child, scope, problems, filename)


@pytest.mark.parametrize("rel", FILES)
def test_every_name_that_is_read_is_bound_somewhere(rel):
    path = os.path.join(ROOT, rel)
    tree = ast.parse(open(path).read(), filename=rel)
    module = Scope()
    module.bound.update({"__file__", "__name__", "__doc__"})
    collect_bindings(tree.body, module)
    problems = []
    for n in tree.body:
        check(n, module, problems, rel)
    assert not problems, "\n".join(problems)


def test_the_checker_sees_an_unbound_name():
    tree = ast.parse("def f(a):\n    b = a + 1\n    return b + c\n")
    module = Scope()
    collect_bindings(tree.body, module)
    problems = []
    for n in tree.body:
        check(n, module, problems, "x.py")
    assert problems == ["x.py:3: name 'c' is never bound"]


@pytest.mark.parametrize("rel", ["bench.py", "bench/halo_benchmark.py", "bench/fft_benchmark.py", "bench/nccl_restated.py"])
def test_every_args_attribute_has_an_option(rel):
    """`args.<name>` must be an option the script's argument parser defines."""
    import re
    src = open(os.path.join(ROOT, rel)).read()
    used = set(re.findall(r"\bargs\.([a-zA-Z_]\w*)", src))
    defined = set()
    for m in re.finditer(r'add_argument\(\s*"(--?[\w-]+)"(?:,\s*"(--[\w-]+)")?([^)]*)\)', src):
        dest = re.search(r'dest="(\w+)"', m.group(3))
        defined.add(dest.group(1) if dest else (m.group(2) or m.group(1)).lstrip("-").replace("-", "_"))
    assert used <= defined, sorted(used - defined)
