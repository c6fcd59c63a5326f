"""The reference's example programs (examples/cc/basic_usage/basic_usage.cu, basic_usage_autotune.cu and the
Taylor-Green solver examples/cc/taylor_green/tg.cu) are compiled UNMODIFIED against this repo's cudecomp.h and linked
with libcudecomp.so by oracle/ref_tests.mk (outputs in oracle/_ref/, built where /root/reference is mounted). Host-only
check of the boundary for real applications: the binaries exist, every cudecomp* / MPI_* symbol they import is exported
by the library, and the dynamic loader resolves the library through the binary's own run path."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
LIB = os.path.join(ROOT, "cudecomp_b200", "lib", "libcudecomp.so")
EXAMPLES = ["example_basic_usage", "example_basic_usage_autotune", "example_tg"]


def dynamic_symbols(path, defined):
    out = subprocess.run(["nm", "-D", "--defined-only" if defined else "--undefined-only", path], capture_output=True,
                         text=True, check=True).stdout
    return {line.split()[-1].split("@")[0] for line in out.splitlines() if line.strip()}


@pytest.mark.parametrize("name", EXAMPLES)
def test_reference_example_links_against_the_library(name):
    exe = os.path.join(REF, name)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref binaries not built (make -f oracle/ref_tests.mk needs /root/reference)")
    exported = dynamic_symbols(LIB, True)
    wanted = {s for s in dynamic_symbols(exe, False) if s.startswith(("cudecomp", "MPI_"))}
    assert "cudecompInit" in wanted and "cudecompGridDescCreateVersioned" in wanted and "MPI_Init" in wanted
    assert any(s.startswith("cudecompTranspose") for s in wanted)
    assert wanted <= exported, sorted(wanted - exported)
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    line = [l for l in ldd.splitlines() if "libcudecomp.so" in l]
    assert line and os.path.realpath(line[0].split("=>")[1].split("(")[0].strip()) == os.path.realpath(LIB), ldd
