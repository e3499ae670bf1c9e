"""ctypes binding of the C-ABI library (`include/hdiff_b200.h`).  No torch types cross the boundary:
device pointers travel as integers, sizes as C ints, the CUDA stream as an opaque handle."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HDIFF_LIB_PATH") or os.path.join(HERE, "libhdiff_b200.so")      # (HDIFF_LIB_PATH: the lab build, for timing experiments)

P, I, L, F, U64, D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_double

# name -> argument types (all return int status).  Kept in the order of include/hdiff_b200.h.
PROTOTYPES = {
    "hd_conv_simt": [I, P, I, P, I, I, I, P, P, P, L, P, P, I, I, I, I, I, I, I, P],
    "hd_wgrad_simt": [I, P, I, P, I, I, I, P, I, I, I, P, I, I, I, I, P],
    "hd_attn_fwd_simt": [I, P, P, P, I, I, I, P],
    "hd_attn_bwd_simt": [I, P, P, P, P, P, P, I, I, I, P],
    "hd_linear_fwd": [P, I, I, L, P, P, P, I, L, I, I, P],
    "hd_linear_bwd_x": [P, I, I, L, P, I, P, L, P, L, I, P],
    "hd_linear_bwd_w": [P, I, I, L, P, I, L, I, P, P, P],
    "hd_embedding_fwd": [P, I, I, P, I, P, P],
    "hd_embedding_bwd": [P, I, P, I, P, L, P],
    "hd_gather_pack": [I, P, P, P, L, P, P],
    "hd_scatter_unpack": [P, P, L, P, P],
    "hd_gn_stats": [I, P, I, P, I, I, L, I, P, P],
    "hd_gn_v2": [I, I, I, I, L, I],
    "hd_gn_apply": [I, P, I, P, I, I, L, I, P, P, P, F, I, F, U64, P, P],
    "hd_gn_bwd_reduce": [I, P, I, P, I, I, L, I, P, P, P, F, I, F, U64, P, P, P, P, P, P],
    "hd_gn_bwd_apply": [I, P, I, P, I, I, L, I, P, P, P, F, I, F, U64, P, P, P, P, P, P, P, P, P, L, I, I, P],
    "hd_gn_bwd_fused": [I, P, I, P, I, I, L, I, P, P, P, F, I, F, U64, P, P, P, P, P, P, P, P, P, P, P],
    "hd_colsum": [I, P, I, I, L, I, P, L, P, P],
    "hd_q_sample": [P, P, P, P, P, P, I, L, I, P],
    "hd_mse_fwd": [P, P, P, L, P],
    "hd_mse_bwd": [P, P, P, P, L, P],
    "hd_sampler_step": [P, P, P, P, F, F, P, P, I, P, L, P],
    "hd_add_int": [P, I, P],
    "hd_sqnorm": [P, L, P, I, P],
    "hd_adamw_flat": [P, P, P, P, L, P, F, D, D, D, D, D, I, P],
    "hd_conv_tc": [P, I, P, I, I, P, P, P, L, P, P, I, I, I, I, I, I, I, P, P],
    "hd_gn_group_sums": [P, I, P, I, I, I, P, P],
    "hd_pad_nchw": [P, I, P, I, L, P],
    "hd_conv_tc_stats_staged": [I, I, I, I, I, I, I, I],
    "hd_conv_tc_supported": [I, I, I, I, I, I, I, I],
    "hd_wgrad_tc": [P, I, P, I, I, P, I, I, P, P, L, I, I, I, I, P],
    "hd_wgrad_tc_supported": [I, I, I, I, I, I, I, I],
    "hd_wgrad_tc_workspace": [I, I, I, I, I, I, I, I, I],
    "hd_attn_fwd_tc": [P, P, P, I, I, I, P],
    "hd_attn_bwd_tc": [P, P, P, P, P, P, I, I, I, P],
    "hd_attn_fwd_tc_scaled": [P, P, P, I, I, F, P],
    "hd_attn_bwd_tc_scaled": [P, P, P, P, P, P, I, I, F, P],
    "hd_attn_tc_supported": [I, I],
    "hd_attn_wide_tc_supported": [I, I],
    "hd_attn_fwd_wide_tc": [P, P, P, I, I, I, P],
    "hd_attn_bwd_wide_tc": [P, P, P, P, P, P, I, I, I, P],
    "hd_attn_bwd_tc_supported": [I, I],
    "hd_upsample_nearest": [I, P, P, I, I, I, I, I, I, P],
    "hd_upsample_nearest_bwd": [I, P, P, I, I, I, I, I, I, P],
    "hd_image_affine": [P, I, P, F, F, L, P],
    "hd_resize_bilinear_u8": [P, I, I, I, I, P, I, I, I, P],
    "hd_sq_err_u8": [P, P, I, L, P, P],
    "hd_uiqm_workspace": [I],
    "hd_uiqm_u8": [P, I, I, I, P, L, P, P],
    "hd_ssim_u8": [P, P, I, I, I, I, I, P, P],
    "hd_rgb2lab_u8": [P, L, P, P],
    "hd_lab_tables_host": [P, P],
    "hd_uciqe_workspace": [I],
    "hd_uciqe_u8": [P, I, I, I, P, L, P, P],
    "hd_mha_supported": [I, I],
    "hd_mha_pack_heads": [P, P, I, I, I, I, I, F, P],
    "hd_mha_unpack_heads": [P, P, I, I, I, I, I, F, P],
    "hd_mha_fwd": [I, P, P, P, I, I, I, I, P],
    "hd_mha_bwd": [I, P, P, P, P, P, P, I, I, I, I, P],
}
NON_STATUS = {"hd_uiqm_workspace", "hd_uciqe_workspace", "hd_gn_v2", "hd_mha_supported", "hd_conv_tc_supported", "hd_conv_tc_stats_staged", "hd_wgrad_tc_supported", "hd_attn_tc_supported", "hd_attn_bwd_tc_supported", "hd_attn_wide_tc_supported",
              "hd_wgrad_tc_workspace"}

# lab library only (include/hdiff_b200_lab.h): hardware probes and timing experiments, not part of the product ABI
LAB_PROTOTYPES = {
    "hd_probe_shift": [P, P, P, I, I, P],
    "hd_conv_dbg_read": [P],
    "hd_probe_queue": [P, P, I, I, I, I, I, P, P],
    "hd_probe_pair": [P, P, P, I, I, I, I, P, I, I, I, I, I, P],
}
LAB_LIB_PATH = os.path.join(HERE, "libhdiff_b200_lab.so")

_lib = None


class HdiffError(RuntimeError):
    pass


def load():
    """Load the CUDA library.  There is no fallback: a missing library is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HdiffError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(nvcc, sm_100a). There is no CPU or library fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    lib.hd_last_error.restype = C.c_char_p
    lib.hd_last_error.argtypes = []
    lib.hd_abi_version.restype = I
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here means the .so is stale
        fn.argtypes = args
        fn.restype = L if name in ("hd_wgrad_tc_workspace", "hd_uiqm_workspace", "hd_uciqe_workspace") else I
    _lib = lib
    return lib


def load_lab():
    """The lab build (product kernels compiled with -DHDIFF_LAB + the probes): scripts/probe_*.py and scripts/conv_clock.py only."""
    if not os.path.exists(LAB_LIB_PATH):
        raise HdiffError(f"{LAB_LIB_PATH} is missing: run `python -m hdiff_b200.build --lab`")
    lib = C.CDLL(LAB_LIB_PATH)
    lib.hd_last_error.restype = C.c_char_p
    for table in (PROTOTYPES, LAB_PROTOTYPES):
        for name, args in table.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = L if name == "hd_wgrad_tc_workspace" else I
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().hd_last_error().decode(errors="replace")
        raise HdiffError(f"{what} failed with status {rc}: {msg}")
