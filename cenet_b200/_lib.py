"""ctypes binding of libcenet_b200.so (the C ABI declared in include/cenet_b200.h).

There is deliberately no fallback: if the library is missing or a kernel reports an error, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcenet_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU, ACT_LEAKY, ACT_SILU, ACT_SIGMOID, ACT_GELU_GRAD = range(7)
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05, GEMM_MMA = -1, 0, 1, 2

vp, ll, i32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_float


class AttnTcArgs(C.Structure):
    """Mirror of `cenet_attn_tc_args` (include/cenet_b200.h)."""
    _fields_ = [("q", vp), ("k", vp), ("v", vp), ("o", vp), ("lse", vp),
                ("ldq", ll), ("ldk", ll), ("ldv", ll), ("ldo", ll), ("bq", ll), ("bk", ll), ("bv", ll), ("bo", ll),
                ("B", i32), ("heads", i32), ("Nq", i32), ("Nk", i32), ("D", i32), ("scale", f32), ("lse_base2", i32)]


class WgradJob(C.Structure):
    """Mirror of `cenet_wgrad_job` (include/cenet_b200.h)."""
    _fields_ = [("src", vp), ("dst", vp), ("stride", ll), ("S", i32), ("N", i32), ("K", i32), ("T", i32), ("blk0", i32),
                ("src_ld", i32)]


class GemmArgs(C.Structure):
    """Mirror of `cenet_gemm_args` (include/cenet_b200.h)."""
    _fields_ = [
        ("M", i32), ("N", i32), ("K", i32),
        ("batch", i32), ("batch_inner", i32),
        ("A", vp), ("a_dtype", i32), ("lda", ll), ("a_bs_outer", ll), ("a_bs_inner", ll),
        ("conv", i32), ("Bimg", i32), ("H", i32), ("W", i32), ("Cin", i32), ("KH", i32), ("KW", i32),
        ("stride", i32), ("pad", i32), ("Ho", i32), ("Wo", i32),
        ("Wt", vp), ("w_dtype", i32), ("ldw", ll), ("w_bs_outer", ll), ("w_bs_inner", ll), ("w_nmajor", i32),
        ("C", vp), ("c_dtype", i32), ("ldc", ll), ("c_bs_outer", ll), ("c_bs_inner", ll),
        ("alpha", f32), ("bias", vp), ("bias_per_row", i32), ("row_scale", vp),
        ("act", i32), ("slope", f32), ("act_after_res", i32),
        ("res1", vp), ("res1_dtype", i32), ("ldr1", ll), ("res1_cscale", vp), ("res1_scale", f32),
        ("res2", vp), ("res2_dtype", i32), ("ldr2", ll),
        ("mul", vp), ("mul_dtype", i32), ("ldmul", ll), ("mul_act", i32),
        ("impl", i32),
        # training extensions (zero == inference behaviour)
        ("a_mmajor", i32), ("rs_div", i32), ("post_row_scale", vp), ("post_rs_div", i32),
        ("k_scale", vp), ("k_scale_div", i32), ("k_scale_bs", ll),
        ("split_ws", vp), ("split_ws_elems", ll),
    ]


# name -> argument ctypes (all return int unless noted)
_SIGS = {
    "cenet_gemm": [C.POINTER(GemmArgs), vp],
    "cenet_layernorm": [vp, i32, vp, i32, vp, vp, ll, i32, f32, vp],
    "cenet_softmax_rows": [vp, i32, ll, i32, ll, vp],
    "cenet_row_stats": [vp, i32, ll, i32, ll, i32, vp, vp],
    "cenet_mixffn_tail": [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "cenet_dwconv3x3": [vp, i32, ll, vp, i32, ll, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, vp],
    "cenet_nhwc_to_nchw": [vp, i32, ll, vp, i32, i32, i32, i32, i32, i32, vp],
    "cenet_nchw_to_nhwc": [vp, i32, vp, i32, ll, i32, i32, i32, vp],
    "cenet_im2col": [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "cenet_upsample2x_ac": [vp, i32, vp, i32, i32, i32, i32, i32, vp],
    "cenet_maxpool2_scale": [vp, i32, vp, i32, ll, i32, vp, i32, i32, i32, i32, vp],
    "cenet_affine_gate": [vp, i32, vp, i32, vp, vp, vp, i32, i32, i32, vp],
    "cenet_fea_combine": [vp, vp, vp, i32, vp, i32, i32, i32, i32, C.POINTER(f32), i32, vp],
    "cenet_dog_combine": [vp, vp, vp, i32, vp, i32, i32, i32, i32, C.POINTER(f32), i32, i32, vp],
    "cenet_diff_combine": [vp, i32, ll, ll, f32, vp],
    "cenet_rmsnorm_seg": [vp, i32, vp, i32, ll, i32, i32, f32, f32, vp],
    "cenet_diffattn_flash": [vp, vp, i32, i32, i32, i32, f32, f32, f32, vp, vp],
    "cenet_diffattn_flash_padded": [vp, vp, i32, i32, i32, i32, i32, i32, f32, f32, f32, vp, vp],
    "cenet_sr_attention": [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, f32, vp],
    "cenet_nonlocal_flash": [vp, vp, i32, i32, i32, f32, vp],
    "cenet_ccu_gate": [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "cenet_srm_gate": [vp, vp, vp, vp, f32, f32, i32, i32, i32, vp],
    "cenet_pool_branch": [vp, i32, ll, i32, vp, i32, ll, i32, vp, vp, vp, f32, vp, i32, i32, i32, i32, vp],
    "cenet_stem5x5": [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp],
    "cenet_head_upsample_argmax": [vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "cenet_dice_ce": [vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, f32, vp],
    "cenet_seg_loss": [vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, f32, f32, f32, vp],
    # ---- training (see include/cenet_b200.h) ----
    "cenet_dwconv3x3_train": [vp, i32, ll, vp, i32, ll, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, vp],
    "cenet_gemm_wgrad": [vp, i32, ll, vp, i32, ll, ll, i32, i32, i32, vp, i32, vp, vp, i32, vp, ll, vp],
    "cenet_gemm_wgrad_partial": [vp, i32, ll, vp, i32, ll, ll, i32, i32, i32, vp, i32, i32, vp, vp, i32, vp, ll,
                                 C.POINTER(i32), C.POINTER(i32), vp],
    "cenet_wgrad_reduce_batch": [vp, i32, i32, vp],
    "cenet_colsum": [vp, i32, ll, ll, i32, vp, i32, vp, vp, ll, vp],
    "cenet_row_scale": [vp, i32, vp, vp, ll, i32, vp],
    "cenet_droppath_mask": [vp, vp, i32, i32, C.c_ulonglong, vp, vp],
    "cenet_smallk_dgrad": [vp, i32, vp, ll, vp, i32, ll, ll, i32, i32, vp],
    "cenet_conv_wgrad": [vp, i32, ll, vp, i32, ll, i32, i32, i32, i32, i32, i32, vp, vp, ll, vp],
    "cenet_layernorm_bwd": [vp, vp, i32, vp, f32, ll, i32, vp, i32, vp, vp, vp, ll, C.POINTER(i32), vp],
    "cenet_bn_stats": [vp, i32, ll, ll, i32, vp, vp, vp, vp, vp, f32, f32, i32, vp, vp, vp, vp, vp, ll, vp],
    "cenet_affine_act": [vp, i32, ll, vp, vp, vp, i32, ll, vp, vp, vp, i32, ll, ll, i32, i32, f32, vp],
    "cenet_bn_bwd": [vp, i32, vp, i32, ll, vp, i32, ll, vp, vp, vp, ll, i32, i32, f32, vp, i32, i32, vp, vp, vp, i32, ll,
                     i32, i32, vp, ll, vp],
    "cenet_dwconv3x3_wgrad": [vp, i32, ll, vp, i32, ll, i32, i32, i32, i32, i32, i32, vp, vp, vp, ll, vp],
    "cenet_sumpool2": [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp],
    "cenet_col2im": [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "cenet_flash_fwd": [vp, ll, vp, ll, vp, ll, vp, ll, vp, i32, i32, i32, i32, i32, i32, i32, f32, vp],
    "cenet_diffattn_fwd_train": [vp, vp, vp, i32, i32, i32, i32, vp, vp],
    "cenet_flash_bwd": [vp, ll, vp, ll, vp, ll, vp, vp, ll, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, vp, ll, vp],
    "cenet_softmax_bwd_rows": [vp, vp, i32, ll, i32, vp],
    "cenet_lambda_fwd": [vp, vp, vp, vp, i32, f32, vp, vp],
    "cenet_lambda_bwd": [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp],
    "cenet_diff_rmsnorm_fwd": [vp, i32, vp, vp, ll, i32, i32, f32, f32, vp],
    "cenet_diff_rmsnorm_bwd": [vp, vp, i32, vp, vp, vp, ll, i32, i32, f32, f32, vp, ll, vp],
    "cenet_fea_bwd": [vp, vp, vp, i32, vp, vp, i32, vp, vp, i32, i32, i32, i32, vp, vp, i32, i32, i32, vp, ll, vp],
    "cenet_nchw_to_nhwc_slice": [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp],
    "cenet_add": [vp, vp, i32, ll, i32, vp],
    "cenet_ccu_stats": [vp, i32, vp, vp, i32, i32, i32, vp, ll, vp],
    "cenet_ccu_mlp_fwd": [vp, vp, vp, vp, vp, vp, vp, vp, f32, f32, vp, vp, i32, i32, i32, vp],
    "cenet_ccu_mlp_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "cenet_ccu_dgate": [vp, vp, i32, vp, i32, i32, i32, vp, ll, vp],
    "cenet_ccu_apply_bwd": [vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "cenet_row_stats_arg": [vp, i32, vp, vp, ll, i32, vp],
    "cenet_srm_fwd": [vp, vp, vp, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, i32, i32, i32, i32, vp, ll, vp],
    "cenet_row_dot": [vp, vp, i32, vp, ll, i32, vp],
    "cenet_srm_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, ll, vp],
    "cenet_srm_apply_bwd": [vp, vp, vp, i32, vp, vp, vp, vp, vp, ll, i32, vp],
    "cenet_silu_mul_fwd": [vp, vp, vp, i32, ll, vp],
    "cenet_silu_mul_bwd": [vp, vp, vp, vp, vp, i32, ll, vp],
    "cenet_ls_combine_fwd": [vp, vp, vp, i32, vp, vp, vp, vp, vp, ll, i32, vp],
    "cenet_ls_combine_bwd": [vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, ll, i32, vp, ll, vp],
    "cenet_resample": [vp, i32, ll, vp, i32, ll, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp],
    "cenet_maxpool2_scale_bwd": [vp, i32, ll, vp, i32, vp, vp, vp, i32, i32, i32, i32, vp, ll, vp],
    "cenet_head_upsample_bwd": [vp, vp, i32, i32, i32, i32, vp],
    "cenet_adamw": [vp, vp, vp, vp, ll, vp, vp],
    "cenet_attn_tc": [vp, vp],
    "cenet_volume_labels_counts": [vp, i32, i32, vp, vp, vp, i32, vp, vp, i32, i32, i32, i32, vp],
    "cenet_gather_cast": [vp, vp, vp, i32, ll, vp],
}
_PLAIN = {  # no stream, different return types
    "cenet_last_error": ([], C.c_char_p),
    "cenet_abi_version": ([], i32),
    "cenet_mixffn_tail_supported": ([i32, i32, i32, i32], i32),
    "cenet_launch_count": ([], ll),
    "cenet_ccu_nchunk": ([i32], i32),
    "cenet_loss_nblocks": ([ll], i32),
    "cenet_wgrad_reduce_blocks": ([C.POINTER(WgradJob)], i32),
    "cenet_wgrad_plan_query": ([ll, i32, i32, i32, i32, i32, ll, C.POINTER(i32)], i32),
}
EXPORTS = sorted(list(_SIGS) + list(_PLAIN))

_lib = None


def load():
    """Load the shared library once; raises with a build hint when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: run `python -m cenet_b200.build` (needs nvcc, sm_100a). "
                           "cenet_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i32
    for name, (args, res) in _PLAIN.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().cenet_last_error()
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else 'unknown error'}")


def call(name: str, *args):
    check(getattr(load(), name)(*args), name)
