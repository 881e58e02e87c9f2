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
    assert h.vist3a_abi_version() == 8
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


def test_argument_validation_precedes_any_device_work(lib):
    """Every entry point validates sizes / alignment / geometry before it touches CUDA or dereferences a pointer, so the checks can be
    exercised on a GPU-less host with placeholder addresses; a call that passes validation then fails on the missing device."""
    import ctypes as C

    h = lib.load()
    no_gpu = not torch.cuda.is_available()   # with a device present the placeholder addresses must never reach a launch
    P = 0x1000        # placeholder, 256-byte aligned, never dereferenced
    cam = (C.c_float * 16)(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1)
    K = (C.c_float * 9)(100, 0, 32, 0, 100, 32, 0, 0, 1)

    def err():
        return h.vist3a_last_error().decode()

    # workspace sizes are pure host arithmetic
    n = 1000
    vb = h.vist3a_voxel_fusion_workspace_bytes(n)
    assert vb >= n * (8 + 8 + 4 + 4 + 4 + 4) and h.vist3a_voxel_fusion_workspace_bytes(0) == 0
    pb = h.vist3a_gs_project_workspace_bytes(n)
    assert pb >= n * (40 + 8 + 4 + 4)
    assert h.vist3a_gs_rasterize_workspace_bytes(5000, 64, 64) > 5000 * 24

    # voxel fusion
    args = dict(pts=P, feats=P, ld=83, C_=83, conf=P, cs=1, N=n, vs=0.002, vp=P, vf=P, inv=None, cnt=None, nv=P, ws=P, wsb=vb, st=None)
    call = lambda a: h.vist3a_voxel_fusion(a["pts"], a["feats"], a["ld"], a["C_"], a["conf"], a["cs"], a["N"], a["vs"], a["vp"], a["vf"],
                                           a["inv"], a["cnt"], a["nv"], a["ws"], a["wsb"], a["st"])
    assert call({**args, "pts": None}) == lib.ERR_INVALID and "null" in err()
    assert call({**args, "N": 0}) == lib.ERR_INVALID and "N must be" in err()
    assert call({**args, "C_": 126, "ld": 126}) == lib.ERR_INVALID and "feature dim" in err()
    assert call({**args, "vs": 0.0}) == lib.ERR_INVALID and "voxel_size" in err()
    assert call({**args, "ws": P + 16}) == lib.ERR_INVALID and "aligned" in err()
    assert call({**args, "wsb": vb - 1}) == lib.ERR_INVALID and "workspace of" in err()
    assert not no_gpu or call(args) in (lib.ERR_ARCH, lib.ERR_CUDA)        # valid arguments: only the device is missing

    # rasteriser, phase 1
    pargs = dict(m=P, c=P, o=P, h=P, dsh=25, deg=4, N=n, vm=cam, K=K, W=64, H=64, near=1e-10, far=1e10, rc=0.1, eps=0.3, ws=P, wsb=pb, ni=P, st=None)
    pcall = lambda a: h.vist3a_gs_project(a["m"], a["c"], a["o"], a["h"], a["dsh"], a["deg"], a["N"], a["vm"], a["K"], a["W"], a["H"], a["near"],
                                          a["far"], a["rc"], a["eps"], a["ws"], a["wsb"], a["ni"], a["st"])
    assert pcall({**pargs, "deg": 5}) == lib.ERR_INVALID and "sh_degree" in err()
    assert pcall({**pargs, "dsh": 16}) == lib.ERR_INVALID and "sh_degree" in err()
    assert pcall({**pargs, "near": -1.0}) == lib.ERR_INVALID and "near_plane" in err()
    assert pcall({**pargs, "far": 0.0}) == lib.ERR_INVALID and "near_plane" in err()
    assert pcall({**pargs, "W": 0}) == lib.ERR_INVALID and "image size" in err()
    assert pcall({**pargs, "wsb": pb - 1}) == lib.ERR_INVALID and "workspace of" in err()
    assert not no_gpu or pcall(pargs) in (lib.ERR_ARCH, lib.ERR_CUDA)

    # rasteriser, phase 2
    bg = (C.c_float * 3)(0, 0, 0)
    rb = h.vist3a_gs_rasterize_workspace_bytes(5000, 64, 64)
    assert h.vist3a_gs_rasterize(P, n, 5000, 64, 64, bg, P, rb - 1, P, P, P, None) == lib.ERR_INVALID and "workspace of" in err()
    assert h.vist3a_gs_rasterize(P, n, 1 << 32, 64, 64, bg, P, rb, P, P, P, None) == lib.ERR_INVALID and "2^32" in err()
    assert not no_gpu or h.vist3a_gs_rasterize(P, n, 5000, 64, 64, bg, P, rb, P, P, P, None) in (lib.ERR_ARCH, lib.ERR_CUDA)
    assert not no_gpu or h.vist3a_launch_count() == 0


def test_gemm_and_attention_argument_validation(lib):
    """the two tensor-core entry points reject bad geometry with a message, before any tensor map is encoded or kernel launched"""
    h = lib.load()
    P = 0x1000

    def err():
        return h.vist3a_last_error().decode()

    def gemm(**kw):
        a = lib.GemmArgs()
        a.A = a.W = a.C = P
        a.M, a.N, a.K = 256, 128, 64
        a.lda = a.ldw = 64
        a.ldc = 128
        a.in_dtype = a.out_dtype = 0       # bf16
        for k, v in kw.items():
            if k.startswith("conv_"):
                setattr(a.conv, k[5:], v)
            else:
                setattr(a, k, v)
        return h.vist3a_gemm(C.byref(a), None)

    assert gemm(A=None) == lib.ERR_INVALID and "null" in err()
    assert gemm(K=0) == lib.ERR_INVALID and "positive" in err()
    assert gemm(ldw=60) == lib.ERR_INVALID and "ldw" in err()
    assert gemm(lda=63) == lib.ERR_INVALID and "lda" in err()
    assert gemm(ldc=120) == lib.ERR_INVALID and "ldc" in err()
    assert gemm(A=P + 8) == lib.ERR_INVALID and "aligned" in err()
    assert gemm(in_dtype=7) == lib.ERR_INVALID and "in_dtype" in err()
    assert gemm(post_act=1) == lib.ERR_UNSUPPORTED and "post_act" in err()
    # implicit-GEMM convolution: K and M must agree with the geometry, channels with the 64-element K block
    conv = dict(conv_enabled=1, conv_kh=3, conv_kw=3, conv_pad_y=1, conv_pad_x=1, conv_n_img=1, conv_h=16, conv_w=16, conv_c_in=64, K=9 * 64, ldw=9 * 64)
    assert gemm(**{**conv, "conv_c_in": 32, "K": 9 * 32, "ldw": 9 * 32}) == lib.ERR_UNSUPPORTED and "c_in" in err()
    assert gemm(**{**conv, "K": 8 * 64, "ldw": 9 * 64}) == lib.ERR_INVALID and "kh*kw*c_in" in err()
    assert gemm(**{**conv, "M": 255}) == lib.ERR_INVALID and "n_img*h_out*w_out" in err()

    def fmha(**kw):
        a = lib.FmhaArgs()
        a.Q = a.K = a.V = a.O = P
        a.batch, a.heads, a.len_q, a.len_kv, a.head_dim = 1, 2, 128, 128, 128
        for f in ("q", "k", "v", "o"):
            setattr(a, f + "_bs", 128 * 256)
            setattr(a, f + "_rs", 256)
            setattr(a, f + "_hs", 128)
        a.scale = 0.088
        for k, v in kw.items():
            setattr(a, k, v)
        return h.vist3a_fmha_fwd(C.byref(a), None)

    assert fmha(Q=None) == lib.ERR_INVALID and "null" in err()
    assert fmha(head_dim=96) == lib.ERR_UNSUPPORTED and "head_dim" in err()
    assert fmha(len_kv=0) == lib.ERR_INVALID
    assert fmha(k_rs=250) == lib.ERR_INVALID and "strides" in err()
    assert fmha(V=P + 4) == lib.ERR_INVALID
    if not torch.cuda.is_available():
        assert gemm() in (lib.ERR_ARCH, lib.ERR_CUDA) and fmha() in (lib.ERR_ARCH, lib.ERR_CUDA)


def test_every_compute_entry_point_rejects_zeroed_arguments(lib):
    """robustness of the boundary: NULL pointers / zero sizes come back as VIST3A_ERR_INVALID with a message from every compute entry point
    (no crash, no launch) -- the reference's FFI-less Python would raise here; a C caller gets a status code"""
    h = lib.load()
    skip = {"vist3a_last_error", "vist3a_abi_version", "vist3a_launch_count", "vist3a_set_pdl", "vist3a_voxel_fusion_workspace_bytes",
            "vist3a_gs_project_workspace_bytes", "vist3a_gs_rasterize_workspace_bytes", "vist3a_quantile_workspace_bytes",
            "vist3a_compact_rows_workspace_bytes", "vist3a_fmha_workspace_bytes"}
    n0 = h.vist3a_launch_count()
    checked = 0
    for name in lib.EXPORTS:
        if name in skip:
            continue
        fn = getattr(h, name)
        assert fn.argtypes is not None, name
        args = []
        for t in fn.argtypes:
            if t in (C.c_void_p, C.c_char_p) or (hasattr(t, "_type_") and not isinstance(t._type_, str)):
                args.append(None)
            elif t in (C.c_float, C.c_double):
                args.append(0.0)
            else:
                args.append(0)
        assert fn(*args) == lib.ERR_INVALID, name
        msg = h.vist3a_last_error().decode()
        assert ":" in msg, (name, msg)            # "<entry point>: <what is wrong>"
        checked += 1
    assert checked == len(lib.EXPORTS) - len(skip) and h.vist3a_launch_count() == n0


def test_fmha_pair_key_split_plan_is_consistent(lib):
    """The work decomposition of the CTA-pair attention kernel (persistent clusters; the tail units laid end to end and cut into one equal
    range per cluster; a merge kernel) is index arithmetic shared by three places: the kernel's per-cluster item list, the host's list of cut
    units and the merge kernel's walk over a unit's pieces.  v3a_debug_fmha_pair_plan_check rebuilds all three with the functions the kernels
    use (no CUDA call) and checks coverage (every key step of every unit exactly once), the item-list bound, slot uniqueness and that the
    merge visits exactly the slots that were written.  Swept over unit counts around multiples of the cluster count, short and long key
    sequences, odd cluster counts, with the split / the persistence switched off."""
    h = lib.load()
    fn = h.v3a_debug_fmha_pair_plan_check
    fn.restype = C.c_int
    fn.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    out = (C.c_int * 5)()
    split_seen = whole_seen = 0
    for slots in (1, 2, 7, 66, 74, 128, 200):
        for units in list(range(1, 40)) + [73, 74, 75, 96, 147, 148, 149, 192, 222, 640, 960, 4439, 4441, 5000]:
            for n_kv in (1, 3, 4, 7, 8, 9, 16, 31, 32, 48, 105, 257):
                for allow_split, persistent in ((1, 1), (0, 1), (1, 0)):
                    rc = fn(units, n_kv, slots, allow_split, persistent, out)
                    assert rc == 0, (rc, units, n_kv, slots, allow_split, persistent)
                    clusters, n_full, tail, n_cut, partial = list(out)
                    assert n_full + tail == units and clusters >= 1
                    if not allow_split or not persistent or n_kv < 8:
                        assert tail == 0 and partial == 0
                    if tail:
                        split_seen += 1
                        assert clusters <= min(slots, 128) and partial <= 2 * clusters and n_cut <= tail
                    else:
                        whole_seen += 1
    assert split_seen > 500 and whole_seen > 500
    # the BASELINE shapes on a B200 (74 cluster slots): B2 H12 L4096 -> 192 units of 32 steps: two whole waves + 44 tail units in 74 ranges
    assert fn(192, 32, 74, 1, 1, out) == 0 and list(out)[:3] == [74, 148, 44]
    assert fn(192, 4, 74, 1, 1, out) == 0 and list(out)[:3] == [74, 192, 0]          # cross-attention (512 keys): too short to cut
    assert fn(24, 32, 74, 1, 1, out) == 0 and list(out)[:3] == [74, 0, 24]           # fewer units than clusters: everything is "tail"
