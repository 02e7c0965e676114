"""ctypes binding of include/attndm_b200.h (the C-ABI of libattndm_b200.so).

There is no fallback: if the library is missing or the tensors are not CUDA
tensors, the call raises.  Only raw device pointers cross the boundary.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# ATTNDM_LIB selects an A/B build variant of the same library (attentiondm_b200.build.VARIANTS); never a fallback
LIB_PATH = os.environ.get("ATTNDM_LIB") or os.path.join(_HERE, "libattndm_b200.so")

ROWS_PLAIN, ROWS_HALO = 0, 1
PRE_NONE, PRE_SILU, PRE_GN_SILU = 0, 1, 2
CONV_SIMT, CONV_TCGEN05 = 0, 1

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float


class AttnQuant(C.Structure):
    _fields_ = [("scale", C.c_float), ("zero_point", C.c_float), ("bits", C.c_int)]


# name -> argtypes, in the order of include/attndm_b200.h
SIGNATURES = {
    "attndm_act_quant": [vp, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp, vp, f32, vp, vp, i32, vp, vp],
    "attndm_conv_f32_tc_fits": [i64, i32, i32],
    "attndm_split_tf32": [vp, i64, vp, vp, vp],
    "attndm_gemm_tf32x3": [vp, vp, i64, i32, vp, vp, i32, vp, vp, vp],
    "attndm_gn_stats": [vp, i32, i32, i32, i32, vp, vp],
    "attndm_gn_stats_cat": [vp, i32, i32, i32, vp, i32, i32, i32, i32, vp, vp],
    "attndm_act_quant_cat_fits": [i32, i32, i32, i32],
    "attndm_act_quant_cat": [vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp, vp, f32, vp, vp, i32, vp],
    "attndm_act_quant_cat2": [vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, vp, f32, vp, vp, i32,
                              vp, vp, vp, vp, i32, vp],
    "attndm_gn_act_quant_fits": [i32, i32, i32],
    "attndm_gn_act_quant": [vp, i32, i32, i32, i32, vp, vp, f32, vp, vp, i32, vp, vp, i32, vp, vp],
    "attndm_gn_silu": [vp, i32, i32, i32, i32, vp, vp, vp, f32, vp, vp],
    "attndm_minmax_workspace_blocks": [],
    "attndm_minmax_c": [vp, i64, i32, vp, vp, vp, vp],
    "attndm_group_ranges": [vp, vp, i32, i32, f32, f32, vp, vp, vp, vp],
    "attndm_calib_mix": [vp, i64, i32, i32, vp, vp, i32, vp, vp, f32, vp],
    "attndm_kth_value": [vp, i64, i64, vp, vp, vp],
    "attndm_weight_clamp_pack": [vp, i32, i32, i32, i32, vp, vp, vp, vp],
    "attndm_weight_to_i8": [vp, i32, i32, i32, vp, vp, i32, vp, i32, vp, vp, vp, vp],
    "attndm_qconv_i8": [vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "attndm_conv_f32": [vp, i32, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp],
    "attndm_attention": [vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, f32, AttnQuant, AttnQuant, vp],
    "attndm_scale_add": [vp, vp, vp, vp, i64, vp],
    "attndm_maxpool2": [vp, i32, i32, i32, i32, vp, vp],
    "attndm_upsample_concat": [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp],
    "attndm_timestep_embedding": [vp, i32, i32, vp, vp],
    "attndm_ddim_step": [vp, vp, vp, vp, vp, vp, i64, vp],
    "attndm_ddim_step_hist": [vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, i32, vp],
    "attndm_stage_tables": [vp, i64, i32, vp, i32, vp, vp],
    "attndm_ddpm_step": [vp, vp, vp, vp, vp, vp, i64, vp],
    "attndm_noise_mix": [vp, vp, vp, vp, i64, vp],
    "attndm_sq_err": [vp, vp, i32, i64, vp, vp],
    "attndm_alpha_entropy_grad": [vp, i32, i32, f32, vp, vp, vp],
    "attndm_rowprog": [vp, vp, i32, i32, i32, i32, i32, i32, vp, i64, vp, i32, vp],
    "attndm_bcast_rows": [vp, vp, i32, i32, i32, vp, vp],
    "attndm_rowprog_smem_bytes": [i32, i32, i32, i32],
    "attndm_rowprog_packed_weight_bytes": [i32, i32],
    "attndm_rowprog_pack_weights": [vp, i32, i32, vp, vp],
    "attndm_version": [],
    "attndm_device_supported": [],
}

_lib = None


def lib():
    """Load the shared library (built in-tree by attentiondm_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"attentiondm_b200: {LIB_PATH} is missing -- run `python -m attentiondm_b200.build` "
                "(there is no CPU / PyTorch fallback for this path)")
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        L.attndm_rowprog_packed_weight_bytes.restype = C.c_longlong
        L.attndm_last_error.restype = C.c_char_p
        L.attndm_last_error.argtypes = []
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().attndm_last_error().decode(errors="replace")
        raise RuntimeError(f"attentiondm_b200.{what} failed ({rc}): {msg}")


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("attentiondm_b200: expected a CUDA tensor; this path has no CPU fallback")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


# launch counter: bench.py reports how many of OUR kernels' entry points ran
launches = 0


def call(name: str, *args):
    global launches
    launches += 1
    check(getattr(lib(), name)(*args), name)
