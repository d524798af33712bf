"""ctypes binding of libgptst_b200.so (the C ABI of include/gptst_b200.h).  Fails loudly when absent."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgptst_b200.so")
_lock = threading.Lock()
_lib = None

_f = C.c_void_p  # device pointers are passed as integers
_i, _l = C.c_int, C.c_long

# name -> (restype, argtypes); must list every symbol declared in include/gptst_b200.h
SIGNATURES = {
    "gptst_gproj_fwd": (_i, [_f, _f, _f, _f, _f, _i, _i, _l, _l, _i, _i, _i, _f]),
    "gptst_gproj_splits": (_i, [_i, _i, _i]),
    "gptst_gproj_bwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _l, _l, _i, _i, _i, _i, _f]),
    "gptst_tmix": (_i, [_f, _f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "gptst_tmix_dM_splits": (_i, [_i, _i]),
    "gptst_tmix_dM": (_i, [_f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "gptst_tmix_bwd_splits": (_i, [_i, _i]),
    "gptst_tmix_bwd": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_route_fwd": (_i, [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_hop_fwd": (_i, [_f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_recon": (_i, [_f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_hop_e1": (_i, [_f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_recon_hop": (_i, [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_recon_hop_fused": (_i, [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_dv_dcr": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_hop_bwd": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_hop_bwd_parts": (_i, [_i]),
    "gptst_cap_hop_bwd2": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_dv_dcr_hoprows": (_i, [_f] * 9 + [_i] * 6 + [_f]),
    "gptst_cap_hop_bwd_cols": (_i, [_f] * 7 + [_i] * 5 + [_f]),
    "gptst_cap_route_bwd_parts": (_i, [_i, _i, _i, _i, _i]),
    "gptst_cap_route_bwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_route2_supported": (_i, [_i, _i, _i]),
    "gptst_cap_route_bwd_dz": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_route_fwd_z": (_i, [_f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _f]),
    "gptst_cap_route_bwd_dz_z": (_i, [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "gptst_linear_bwd_acc_splits": (_i, [_l, _i]),
    "gptst_linear_bwd_acc": (_i, [_f, _f, _f, _f, _f, _f, _l, _i, _i, _i, _f]),
    "gptst_mask_labels": (_i, [_f, _f, _f, _l, _i, _f]),
    "gptst_mask_select": (_i, [_f, _f, _f, _f, _f, _f, _f, _l, _i, _i, _i, _i, _f]),
    "gptst_mask_ws_ints": (_i, []),
    "gptst_mask_adaptive": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _f, _l, _i, _i, _i, _f]),
    "gptst_mask_random": (_i, [_f, _f, _f, _f, _l, _f]),
    "gptst_time_mlp_chunks": (_i, [_i]),
    "gptst_time_mlp_grad_floats": (_i, [_i, _i]),
    "gptst_time_mlp_fwd": (_i, [_f] * 16 + [_i, _i, _i, _l, _f]),
    "gptst_time_mlp_bwd": (_i, [_f] * 10 + [_i, _i, _i, _l, _f]),
    "gptst_table_fwd": (_i, [_f, _f, _f, _i, _i, _i, _f]),
    "gptst_table_bwd": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _f]),
    "gptst_table_bwd2_chunks": (_i, [_i, _i, _f, _f]),
    "gptst_table_bwd2": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _f]),
    "gptst_mn_fwd": (_i, [_f, _f, _i, _i, _i, _f]),
    "gptst_mn_bwd": (_i, [_f, _f, _f, _i, _i, _i, _f]),
    "gptst_score_head_fwd": (_i, [_f, _f, _f, _f, _l, _i, _i, _f]),
    "gptst_score_head_bwd_parts": (_i, [_l]),
    "gptst_score_head_bwd": (_i, [_f, _f, _f, _f, _f, _f, _l, _i, _i, _f]),
    "gptst_sum_partials": (_i, [_f, _f, _f, _f, _i, _f]),
    "gptst_affine1_fwd": (_i, [_f, _f, _f, _f, _l, _i, _f]),
    "gptst_masked_affine1_fwd": (_i, [_f, _l, _f, C.c_float, _f, _f, _f, _f, _l, _i, _f]),
    "gptst_affine1_bwd_parts": (_i, [_l]),
    "gptst_affine1_bwd": (_i, [_f, _f, _f, _l, _i, _i, _f]),
    "gptst_gproj3_bwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _l, _l, _i, _i, _i, _i, _i, _f]),
    "gptst_hypertem_wfrag_bytes": (_l, [_i]),
    "gptst_hypertem_mask_pad_rows": (_i, []),
    "gptst_hypertem_pack_w": (_i, [_f, _f, _f, _i, _f]),
    "gptst_hypertem_fwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _f]),
    "gptst_hypertem_bwd": (_i, [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _f]),
    "gptst_hypertem_dw": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "gptst_tmix_dM2": (_i, [_f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "gptst_proj_out_fwd": (_i, [_f, _f, _f, _f, _l, _i, _i, _f]),
    "gptst_proj_out_bwd_parts": (_i, [_l]),
    "gptst_proj_out_bwd": (_i, [_f, _f, _f, _f, _f, _l, _i, _i, _i, _f]),
    "gptst_loss_parts": (_i, []),
    "gptst_pretrain_loss": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _f, _l, _i, _i, _i, _i, C.c_float, C.c_float, C.c_float,
                                C.c_float, _f]),
    "gptst_opt_chunk": (_i, []),
    "gptst_adam_clip": (_i, [_f, _f, _i, _f, _f, _f, _f, _f]),
    "gptst_cap_hop_ev": (_i, [_f] * 4 + [_i] * 5 + [_f]),
    "gptst_cap_recon_proj": (_i, [_f] * 7 + [_i] * 5 + [_f]),
    "gptst_glu_tconv_fwd": (_i, [_f] * 8 + [_i] * 6 + [_f]),
    "gptst_tconv_fwd": (_i, [_f] * 4 + [_i] * 6 + [_f]),
    "gptst_glu_gate_bwd": (_i, [_f] * 4 + [_i] * 4 + [_f]),
    "gptst_tconv_dw_splits": (_i, [_i] * 4),
    "gptst_tconv_dw": (_i, [_f] * 4 + [_i] * 7 + [_f]),
    "gptst_gate_fwd": (_i, [_f] * 7 + [_l, _i, _i, _f]),
    "gptst_gate_blend": (_i, [_f] * 6 + [_l, _f]),
    "gptst_gate_bwd": (_i, [_f] * 7 + [_l, _f]),
    "gptst_version": (C.c_char_p, []),
}


class GptstLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library.  No fallback: raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise GptstLibraryError(
                    f"{LIB_PATH} not found: build it with `python gpt-st_b200/build.py` "
                    "(or __graft_entry__.build()). There is no CPU / PyTorch fallback for this path.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)  # AttributeError if the symbol is missing
                fn.restype, fn.argtypes = res, args
            _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc == -1:
        raise GptstLibraryError(f"{what}: NULL / empty argument")
    if rc == -2:
        raise GptstLibraryError(f"{what}: unsupported shape (D in {{64,128}}, T == 12, H <= 16, prec in {{1,3}})")
    # cudaGetLastError() after the launch: a launch-configuration error of THIS call, or a sticky error left by an earlier
    # asynchronous kernel of the process (run with CUDA_LAUNCH_BLOCKING=1 to attribute it)
    raise GptstLibraryError(f"{what}: CUDA error {rc} (cudaGetLastError after the launch; may stem from an earlier asynchronous kernel)")
