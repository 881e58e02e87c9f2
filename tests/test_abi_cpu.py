"""CPU checks of the drop-in boundary (no compute calls): the built library exports every symbol include/vist3a_sm100.h
declares, the ctypes structs have the C layout, the product path has no CPU fallback and never touches oracle/."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vist3a_sm100.h")


@pytest.fixture(scope="module")
def lib():
    from vist3a_b200 import _lib, build

    build.build()
    return _lib


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vist3a_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 26
    h = lib.load()
    missing = [n for n in names if not hasattr(h, n)]
    assert not missing, missing
    assert sorted(lib.EXPORTS) == names            # the Python binding lists exactly the header's entry points
    assert h.vist3a_abi_version() == 5
    assert h.vist3a_launch_count() == 0            # nothing was launched by loading / symbol lookup


def test_ctypes_structs_match_the_c_layout(lib, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vist3a_sm100.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(vist3a_gemm_args), '
                   'sizeof(vist3a_fmha_args), sizeof(vist3a_rowmap), sizeof(vist3a_conv), offsetof(vist3a_gemm_args, conv), offsetof(vist3a_fmha_args, q_row_scale));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(lib.GemmArgs), C.sizeof(lib.FmhaArgs), C.sizeof(lib.RowMap), C.sizeof(lib.Conv), lib.GemmArgs.conv.offset,
            lib.FmhaArgs.q_row_scale.offset]
    assert got == want


def test_errors_are_reported_not_swallowed(lib):
    h = lib.load()
    rc = h.vist3a_gemm(None, None)                  # null args: validated before any CUDA call
    assert rc == lib.ERR_INVALID and b"null" in h.vist3a_last_error()
    with pytest.raises(lib.Vist3aError):
        lib.check(rc)


def test_no_cpu_fallback_in_the_product_path():
    from vist3a_b200 import ops

    a = torch.randn(8, 8).bfloat16()
    for call in (lambda: ops.gemm(a, a), lambda: ops.layernorm(a), lambda: ops.fmha(a.view(1, 8, 1, 8), a.view(1, 8, 1, 8), a.view(1, 8, 1, 8)),
                 lambda: ops.im2col_stitch(torch.randn(1, 16, 2, 4, 4))):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vist3a_b200")
    bad = []
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                bad.append(fn)
    assert not bad, bad
    code = "import sys; sys.path.insert(0, %r); import vist3a_b200.wan_dit, vist3a_b200.stitched_decoder, vist3a_b200.t23d, vist3a_b200.pipeline; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)


def test_weight_manifests_match_the_oracles():
    from oracle import decoder_ref as D
    from oracle import wan_dit_ref as R
    from vist3a_b200 import stitched_decoder as SD
    from vist3a_b200 import wan_dit as WD

    assert SD.param_shapes(SD.DecoderConfig()) == D.param_shapes(D.FULL)
    assert WD.param_shapes(WD.WAN_1_3B_CONFIG) == R.param_shapes(R.WAN_1_3B)
    assert WD.param_shapes(WD.WAN_14B_CONFIG) == R.param_shapes(R.WAN_14B)
