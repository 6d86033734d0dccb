"""The C-ABI library loads and exports every symbol include/cenet_b200.h declares (no compute calls: CPU-only)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "cenet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cenet_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from cenet_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run python -m cenet_b200.build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_table_matches_header():
    from cenet_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()
    lib = _lib.load()
    assert lib.cenet_abi_version() == 1
    assert lib.cenet_ccu_nchunk(3136) == 25
    assert lib.cenet_loss_nblocks(4097) == 2


def test_gemm_args_struct_layout():
    """ctypes mirror and the C struct must agree; the C side is compiled from the header, so check against gcc."""
    import subprocess
    import tempfile
    from cenet_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "cenet_b200.h"
int main(){ printf("%zu %zu %zu %zu %zu\n", sizeof(cenet_gemm_args), offsetof(cenet_gemm_args, Wt),
  offsetof(cenet_gemm_args, alpha), offsetof(cenet_gemm_args, res2), offsetof(cenet_gemm_args, impl)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(v) for v in subprocess.check_output([exe]).split()]
    G = _lib.GemmArgs
    assert got == [ctypes.sizeof(G), G.Wt.offset, G.alpha.offset, G.res2.offset, G.impl.offset]

