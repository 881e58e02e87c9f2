"""ctypes binding of libvist3a_sm100.so (the C ABI declared in include/vist3a_sm100.h).

There is exactly one compute path in this package: the sm_100a kernels behind this library.  If the
library is missing or the device is not a B200, every op raises -- there is no eager/CPU fallback
(the CPU code under oracle/ is test infrastructure and is never imported from here).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libvist3a_sm100.so"

OK = 0
ERR_INVALID, ERR_ARCH, ERR_CUDA, ERR_UNSUPPORTED = -1, -2, -3, -4
DTYPE_BF16, DTYPE_F32 = 0, 1
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_SILU, ACT_RELU = 0, 1, 2, 3, 4
GEMM_FLAG_2CTA, GEMM_FLAG_1CTA, GEMM_FLAG_BN176, GEMM_FLAG_MULTICAST, GEMM_FLAG_STAGED = 1, 2, 4, 8, 16

# every symbol include/vist3a_sm100.h declares (tests check the built library exports all of them)
EXPORTS = (
    "vist3a_last_error",
    "vist3a_abi_version",
    "vist3a_launch_count",
    "vist3a_set_pdl",
    "vist3a_gemm",
    "vist3a_fmha_fwd",
    "vist3a_fmha_workspace_bytes",
    "vist3a_layernorm",
    "vist3a_rmsnorm_rope",
    "vist3a_row_rinv",
    "vist3a_modulation",
    "vist3a_skinny_linear",
    "vist3a_timestep_features",
    "vist3a_patchify",
    "vist3a_unpatchify",
    "vist3a_cfg_combine",
    "vist3a_axpby_n",
    "vist3a_im2col_stitch",
    "vist3a_im2col_nhwc",
    "vist3a_qknorm_rope2d",
    "vist3a_bilinear_nhwc",
    "vist3a_depth_to_space",
    "vist3a_attention_small",
    "vist3a_fma_rows",
    "vist3a_bias_act_t",
    "vist3a_rgb_to_nhwc4pad",
    "vist3a_rgb01_views_to_nhwc4pad",
    "vist3a_patch_embed_im2col",
    "vist3a_pose_to_cameras",
    "vist3a_gaussian_epilogue",
    "vist3a_gaussian_adapter",
    "vist3a_voxel_fusion",
    "vist3a_gs_project",
    "vist3a_gs_rasterize",
    "vist3a_vae_rmsnorm",
    "vist3a_softmax_rows",
    "vist3a_time_interleave",
    "vist3a_transpose_bf16",
    "vist3a_depth_to_space2_bf16",
    "vist3a_latent_to_ndhwc",
    "vist3a_vae_frames_out",
    "vist3a_resize_planes",
    "vist3a_depth_conf",
    "vist3a_quantile_f32",
    "vist3a_compact_rows",
    "vist3a_voxel_fusion_workspace_bytes",
    "vist3a_gs_project_workspace_bytes",
    "vist3a_gs_rasterize_workspace_bytes",
    "vist3a_quantile_workspace_bytes",
    "vist3a_compact_rows_workspace_bytes",
)


class RowMap(C.Structure):
    _fields_ = [("rpg", C.c_int64), ("gstride", C.c_int64), ("goff", C.c_int64)]


class Conv(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("pad_y", C.c_int32), ("pad_x", C.c_int32),
                ("n_img", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c_in", C.c_int32), ("kt", C.c_int32),
                ("pix_stride", C.c_int64), ("row_stride", C.c_int64), ("img_stride", C.c_int64)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p),
        ("W", C.c_void_p),
        ("C", C.c_void_p),
        ("bias", C.c_void_p),
        ("gate", C.c_void_p),
        ("residual", C.c_void_p),
        ("residual2", C.c_void_p),
        ("M", C.c_int64),
        ("N", C.c_int64),
        ("K", C.c_int64),
        ("lda", C.c_int64),
        ("ldw", C.c_int64),
        ("ldc", C.c_int64),
        ("ldr", C.c_int64),
        ("rows_per_batch", C.c_int64),
        ("gate_bstride", C.c_int64),
        ("cmap", RowMap),
        ("rmap", RowMap),
        ("conv", Conv),
        ("in_dtype", C.c_int32),
        ("out_dtype", C.c_int32),
        ("act", C.c_int32),
        ("post_act", C.c_int32),
        ("round_linear", C.c_int32),
        ("round_gate", C.c_int32),
        ("flags", C.c_uint32),
    ]


class FmhaArgs(C.Structure):
    _fields_ = [
        ("Q", C.c_void_p),
        ("K", C.c_void_p),
        ("V", C.c_void_p),
        ("O", C.c_void_p),
        ("batch", C.c_int64),
        ("heads", C.c_int64),
        ("len_q", C.c_int64),
        ("len_kv", C.c_int64),
        ("head_dim", C.c_int64),
        ("q_bs", C.c_int64),
        ("q_rs", C.c_int64),
        ("q_hs", C.c_int64),
        ("k_bs", C.c_int64),
        ("k_rs", C.c_int64),
        ("k_hs", C.c_int64),
        ("v_bs", C.c_int64),
        ("v_rs", C.c_int64),
        ("v_hs", C.c_int64),
        ("o_bs", C.c_int64),
        ("o_rs", C.c_int64),
        ("o_hs", C.c_int64),
        ("scale", C.c_float),
        ("flags", C.c_uint32),
        ("q_row_scale", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
    ]


class Vist3aError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libvist3a_sm100 error {code}: {msg}")
        self.code = code


_lib = None


def load(build_if_missing: bool = False) -> C.CDLL:
    """Load the kernel library.  Raises if it has not been built (python -m vist3a_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if build_if_missing or os.environ.get("VIST3A_BUILD_ON_IMPORT") == "1":
            from . import build as _b

            _b.build()
        else:
            raise FileNotFoundError(
                f"{LIB_PATH} not found: build the sm_100a kernels first (python -m vist3a_b200.build). "
                "vist3a_b200 has no CPU or eager fallback."
            )
    lib = C.CDLL(str(LIB_PATH))
    lib.vist3a_last_error.restype = C.c_char_p
    lib.vist3a_abi_version.restype = C.c_int
    lib.vist3a_launch_count.restype = C.c_int64
    lib.vist3a_set_pdl.restype = C.c_int
    lib.vist3a_set_pdl.argtypes = [C.c_int32]
    for name in EXPORTS[4:-5]:
        getattr(lib, name).restype = C.c_int
    lib.vist3a_voxel_fusion_workspace_bytes.restype = C.c_int64
    lib.vist3a_voxel_fusion_workspace_bytes.argtypes = [C.c_int64]
    lib.vist3a_gs_project_workspace_bytes.restype = C.c_int64
    lib.vist3a_gs_project_workspace_bytes.argtypes = [C.c_int64]
    lib.vist3a_gs_rasterize_workspace_bytes.restype = C.c_int64
    lib.vist3a_gs_rasterize_workspace_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int64]
    lib.vist3a_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    lib.vist3a_fmha_fwd.argtypes = [C.POINTER(FmhaArgs), C.c_void_p]
    lib.vist3a_fmha_workspace_bytes.argtypes = [C.POINTER(FmhaArgs)]
    lib.vist3a_fmha_workspace_bytes.restype = C.c_int64
    i64, i32, f32, vp, u32 = C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_uint32
    lib.vist3a_layernorm.argtypes = [vp, i32, i64, vp, i32, i64, i64, i64, i64, vp, i64, vp, i64, f32, i32, C.POINTER(RowMap), C.POINTER(RowMap), vp]
    lib.vist3a_rmsnorm_rope.argtypes = [vp, i64, i64, i64, i64, vp, f32, vp, vp, i64, i64, i64, vp]
    lib.vist3a_row_rinv.argtypes = [vp, i64, i64, i64, f32, vp, vp]
    lib.vist3a_modulation.argtypes = [vp, vp, i32, i32, vp, i64, i64, i64, u32, vp]
    lib.vist3a_skinny_linear.argtypes = [vp, i32, i64, vp, i32, i64, vp, vp, i32, i64, i64, i64, i64, i32, i32, vp, vp, i64, vp]
    lib.vist3a_timestep_features.argtypes = [vp, vp, i32, i64, i64, vp]
    lib.vist3a_patchify.argtypes = [vp, i32, vp, i64, i64, i64, i64, i64, vp]
    lib.vist3a_unpatchify.argtypes = [vp, i32, i64, vp, i32, i64, i64, i64, i64, i64, vp]
    lib.vist3a_cfg_combine.argtypes = [vp, vp, i32, f32, vp, i64, vp]
    lib.vist3a_axpby_n.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(f32), i64, vp]
    lib.vist3a_im2col_stitch.argtypes = [vp, i32, vp, i64, i64, i64, i64, i64, vp]
    lib.vist3a_im2col_nhwc.argtypes = [vp, vp, i64, i64, i64, i64, i64, i32, i32, i32, i32, vp]
    lib.vist3a_qknorm_rope2d.argtypes = [vp, i64, i64, i64, vp, vp, vp, vp, f32, vp, vp, i64, i64, i64, i64, vp]
    lib.vist3a_bilinear_nhwc.argtypes = [vp, vp, i64, i64, i64, i64, i64, i64, vp, vp, vp, vp]
    lib.vist3a_depth_to_space.argtypes = [vp, vp, i64, i64, i64, i64, i32, vp]
    lib.vist3a_attention_small.argtypes = [vp, vp, i64, i64, i64, i64, f32, vp]
    lib.vist3a_fma_rows.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, i64, i64, vp]
    lib.vist3a_rgb_to_nhwc4pad.argtypes = [vp, i32, vp, i64, i64, i64, i64, vp]
    lib.vist3a_rgb01_views_to_nhwc4pad.argtypes = [vp, i32, vp, i64, i64, i64, i64, vp]
    lib.vist3a_patch_embed_im2col.argtypes = [vp, i32, vp, i64, i64, i64, i64, i32, C.POINTER(f32), C.POINTER(f32), vp]
    lib.vist3a_bias_act_t.argtypes = [vp, i64, vp, i32, vp, vp, i64, vp, i64, i64, i64, i32, vp]
    lib.vist3a_pose_to_cameras.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, vp]
    lib.vist3a_gaussian_epilogue.argtypes = [vp, i64, i64, vp, f32, vp, i64, vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp, vp, vp,
                                             vp, vp, vp]
    lib.vist3a_gaussian_adapter.argtypes = [vp, vp, i64, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.vist3a_voxel_fusion.argtypes = [vp, vp, i64, i64, vp, i64, i64, f32, vp, vp, vp, vp, vp, vp, i64, vp]
    fp = C.POINTER(f32)
    lib.vist3a_gs_project.argtypes = [vp, vp, vp, vp, i64, i32, i64, fp, fp, i64, i64, f32, f32, f32, f32, vp, i64, vp, vp]
    lib.vist3a_gs_rasterize.argtypes = [vp, i64, i64, i64, i64, fp, vp, i64, vp, vp, vp, vp]
    lib.vist3a_vae_rmsnorm.argtypes = [vp, i64, vp, vp, i64, i64, i64, i32, vp]
    lib.vist3a_softmax_rows.argtypes = [vp, vp, i64, i64, i64, i64, f32, vp]
    lib.vist3a_time_interleave.argtypes = [vp, i64, vp, i64, i64, i64, i64, vp]
    lib.vist3a_transpose_bf16.argtypes = [vp, i64, vp, i64, i64, i64, vp]
    lib.vist3a_depth_to_space2_bf16.argtypes = [vp, vp, i64, i64, i64, i64, i64, vp]
    lib.vist3a_latent_to_ndhwc.argtypes = [vp, i32, vp, i64, i64, i64, vp]
    lib.vist3a_vae_frames_out.argtypes = [vp, i64, vp, i64, vp]
    lib.vist3a_resize_planes.argtypes = [vp, vp, i64, i64, i64, i64, i64, vp]
    lib.vist3a_depth_conf.argtypes = [vp, i64, i64, vp, f32, vp, i64, vp]
    lib.vist3a_quantile_workspace_bytes.restype = C.c_int64
    lib.vist3a_quantile_workspace_bytes.argtypes = [i64]
    lib.vist3a_quantile_f32.argtypes = [vp, i64, f32, vp, vp, i64, vp]
    lib.vist3a_compact_rows_workspace_bytes.restype = C.c_int64
    lib.vist3a_compact_rows_workspace_bytes.argtypes = [i64]
    lib.vist3a_compact_rows.argtypes = [vp, vp, i32, i64, vp, i64, i64, vp, vp, vp, vp, vp, vp, i64, vp]
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != OK:
        raise Vist3aError(rc, load().vist3a_last_error().decode())


def set_pdl(enable: bool) -> bool:
    """Turn programmatic dependent launch of the hot kernels on/off (default on); returns the previous setting."""
    return bool(load().vist3a_set_pdl(1 if enable else 0))


def launch_count() -> int:
    return int(load().vist3a_launch_count())
