"""ctypes binding of libmop_b200.so (the C ABI in include/mop_b200.h).

There is NO CPU fallback: if the shared library is missing or a tensor is not a
CUDA tensor the call raises.  Build with ``python -m multioptpy_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libmop_b200.so")

MOP_OK = 0


class MopError(RuntimeError):
    pass


_p = C.c_void_p
_i = C.c_int
_d = C.c_double
_sz = C.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/mop_b200.h
SIGNATURES = {
    "mop_version": (_i, []),
    "mop_last_error": (C.c_char_p, []),
    "mop_hessian_update_workspace_bytes": (_sz, [_i, _i]),
    "mop_hessian_update": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mop_project_trrot": (_i, [_i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "mop_eigh_workspace_bytes": (_sz, [_i, _i, _i]),
    "mop_eigh": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "mop_rsirfo_workspace_bytes": (_sz, [_i, _i, _i]),
    "mop_rsirfo_step": (_i, [_i, _i, _i, _i, _i, _i, _d, _d, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                             _p, _p, _p, _p, _sz, _p]),
    "mop_rsirfo_step_packed": (_i, [_i, _i, _i, _i, _i, _d, _d, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                    _p, _sz, _p]),
    "mop_rsirfo_step_packed_begin": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mop_rsirfo_step_packed_finish": (_i, [_i, _i, _i, _i, _d, _d, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mop_constraint_project": (_i, [_i, _i, _i, _d, _p, _p, _p, _p, _p, _p, _p, _p]),
    "mop_crsirfo_finalize": (_i, [_i, _i, _d, _p, _p, _p, _p, _p, _p, _p, _p]),
    "mop_add_inplace": (_i, [_sz, _p, _p, _p]),
    "mop_tridiag_stage_count": (_i, [_i]),
    "mop_hessian_sr_workspace_bytes": (_sz, [_i, _i]),
    "mop_hessian_sr_correction": (_i, [_i, _i, _p, _p, _i, _p, _i, _d, _d, _d, _p, _p, _p, _sz, _p]),
    "mop_rsirfo_step_mixed": (_i, [_i, _i, _p, _i, _i, _d, _d, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                   _p, _sz, _p]),
    "mop_pack_lower": (_i, [_i, _i, _p, _p, _p]),
    "mop_unpack_lower": (_i, [_i, _i, _p, _p, _p]),
    "mop_rsprfo_workspace_bytes": (_sz, [_i, _i, _i]),
    "mop_rsprfo_step": (_i, [_i, _i, _i, _i, _i, _d, _d] + [_p] * 16 + [_sz, _p]),
    "mop_rsirfo_spectral_workspace_bytes": (_sz, [_i, _i]),
    "mop_rsirfo_spectral_step": (_i, [_i, _i, _i, _i, _d, _d, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mop_connectivity": (_i, [_i, _i, _p, _p, _i, _d, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "mop_fischer_workspace_bytes": (_sz, [_i, _i]),
    "mop_fischer_hessian": (_i, [_i, _i, _p, _p, _i, _p, _p, _p, _p, _sz, _p]),
    "mop_fischer_d3old_hessian": (_i, [_i, _i, _p, _p, _i, _d, _d, _d, _d, _p, _p, _p, _p, _sz, _p]),
    "mop_fix_atoms_gather": (_i, [_i, _i, _i, _p, _p, _p, _p]),
    "mop_fix_atoms_schur": (_i, [_i, _i, _i, _p, _p, _p, _p, _p]),
    "mop_hessian_ts_modify": (_i, [_i, _i, _p, _p, _p, _p, _p, _p]),
    "mop_hessian_clip_eigvals": (_i, [_i, _i, _p, _p, _p, _p]),
    "mop_fischer_d3_hessian": (_i, [_i, _i, _p, _p, _i, _d, _d, _d, _d, _p, _p, _p, _p, _sz, _p]),
    "mop_bias_term_bytes": (_sz, []),
    "mop_bias_terms": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "mop_kabsch": (_i, [_i, _i, _p, _p, _p, _p, _p, _p]),
    "mop_check_convergence": (_i, [_i, _i, _p, _p, _d, _d, _d, _d, _p, _p, _p]),
    "mop_ric_bmatrix": (_i, [_i, _i, _p, _p, _p]),
    "mop_ric_partial_rows": (_i, [_i, _i, _p, _i, _p, _p, _p]),
    "mop_ric_grad_to_cart": (_i, [_i, _i, _p, _p, _p, _p]),
    "mop_ric_hess_workspace_bytes": (_sz, [_i, _i, _i]),
    "mop_ric_hess_to_cart": (_i, [_i, _i, _p, _p, _i, _p, _p, _p, _sz, _p]),
    "mop_ric_kmatrix": (_i, [_i, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _i, _p, _p]),
    "mop_ric_pb_workspace_bytes": (_sz, [_i, _i]),
    "mop_ric_pb_int_grad": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "mop_ric_pb_cart_grad": (_i, [_i, _i, _i, _p, _p, _p, _p]),
    "mop_swart_workspace_bytes": (_sz, [_i, _i]),
    "mop_swart_hessian": (_i, [_i, _i, _p, _p, _i, _p, _p, _p, _p, _sz, _p]),
    "mop_lindh_workspace_bytes": (_sz, [_i, _i]),
    "mop_lindh_hessian": (_i, [_i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "mop_afir": (_i, [_i, _i, _p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p]),
    "mop_bneb_force": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "mop_neb_ayala": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "mop_neb_limit_tr": (_i, [_i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "mop_neb_redistribute": (_i, [_i, _i, _i, _i, _p, _p, _p, _p]),
    "mop_neb_fire_blend": (_i, [_i, _i, _d, _p, _p, _p, _p, _p, _p]),
    "mop_neb_fire_advance": (_i, [_i, _i, _d, _i, _p, _p, _p, _p, _p, _p]),
    "mop_outer_trust_radius": (_i, [_i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _d, _d, _p]),
    "mop_clamp_and_move": (_i, [_i, _i, _p, _p, _p, _p, _p]),
}

# Private symbols (csrc/mop_private.h): measurement probes and tuning / diagnostic hooks used by bench.py and tools/.
PRIVATE_SIGNATURES = {
    "mop_priv_bench_dfma": (_i, [_i, _i, _p, _p]),
    "mop_priv_bench_fill": (_i, [_p, _sz, _d, _p]),
    "mop_priv_fast_rcp": (_i, [_p, _p, _sz, _p]),
    "mop_priv_latency": (_i, [_p, _p]),
    "mop_priv_barrier_latency": (_i, [_i, _p, _p]),
    "mop_priv_spectrum_timing": (_i, [_p]),
    "mop_priv_large_timing": (_i, [_p]),
    "mop_priv_tridiag_blk_timing": (_i, [_p]),
    "mop_priv_tridiag_cluster_timing": (_i, [_p]),
    "mop_priv_large_cluster": (_i, [_i]),
    "mop_priv_tridiag_cluster_sym": (_i, [_i]),
    "mop_priv_tridiag_cluster_ablate": (_i, [_i]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MopError(
            f"{LIB_PATH} not found: the CUDA library is required (no CPU fallback). "
            "Build it with `python -m multioptpy_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(SIGNATURES.items()) + list(PRIVATE_SIGNATURES.items()):
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != MOP_OK:
        msg = load().mop_last_error().decode("utf-8", "replace")
        raise MopError(f"{what or 'mop call'} failed (code {rc}): {msg}")
