"""ctypes binding of libhb_b200.so (C ABI in include/hb_b200.h).

The library is the product: if it is missing, or there is no CUDA device, every compute call
raises — there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HB_B200_LIB") or os.path.join(HERE, "libhb_b200.so")  # env: dev experiments

HB_OK = 0
HB_ERR_ARG = -1
HB_ERR_CUDA = -2
HB_ERR_UNDERFLOW = -3
HB_ERR_NOMEM = -4
HB_ERR_NODEVICE = -5

WANT_VITERBI = 1
WANT_GAMMA = 2
WANT_LL = 4


class HBError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libhb_b200 error {code}: {msg}")
        self.code = code


class UnderflowError(HBError, FloatingPointError):
    """HiddenMarkov::Viterbi's stop("Problems With Underflow")."""


_p, _i32, _i64, _d = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); every symbol include/hb_b200.h declares
SIGNATURES = {
    "hb_version": (_i32, []),
    "hb_last_error": (C.c_char_p, []),
    "hb_device_count": (_i32, [C.POINTER(C.c_int)]),
    "hb_set_device": (_i32, [C.c_int]),
    "hb_release_workspace": (_i32, []),
    "hb_status_dev": (_i32, [_p]),
    "hb_last_host_io": (_i32, [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "hb_last_kernels": (_i32, [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "hb_set_binom_lut_max": (_i32, [_i32]),
    "hb_set_host_threads": (_i32, [_i32]),
    "hb_set_norm3_mode": (_i32, [_i32]),
    "hb_set_binom2_mode": (_i32, [_i32]),
    "hb_binom2_stats_dev": (_i32, [_p, C.POINTER(C.c_int32)]),
    "hb_norm3_stats_dev": (_i32, [_p, C.POINTER(C.c_int64)]),
    "hb_norm3_flagged_dev": (_i32, [_p, _p, _i64, C.POINTER(C.c_int64)]),
    "hb_hmm_norm_host": (_i32, [_p, _i64, _i64, _p, _i32, _p, _p, _i32, _p, _p, _i32, _p, _p, _p]),
    "hb_hmm_norm_dev": (_i32, [_p, _i64, _i64, _i64, _p, _i32, _p, _p, _i32, _p, _p, _i32, _p, _i64,
                               _p, _i64, _p, _p]),
    "hb_hmm_norm_tiled_dev": (_i32, [_p, _i64, _i64, _p, _i32, _p, _p, _i32, _p, _p, _i32, _p, _p, _p, _p]),
    "hb_hmm_binom_tiled_dev": (_i32, [_p, _p, _i64, _i64, _p, _i32, _p, _d, _p, _p, _i32, _p, _p, _p, _p]),
    "hb_hmm_binom_packed_tiled_dev": (_i32, [_p, _i64, _i64, _p, _i32, _p, _d, _p, _p, _i32, _p, _p, _p, _p]),
    "hb_pack_counts_dev": (_i32, [_p, _p, _i64, _p, _p]),
    "hb_fbr_stats_dev": (_i32, [_p, C.POINTER(C.c_int32)]),
    "hb_forwardback_norm_cm_dev": (_i32, [_p, _i64, _i64, _p, _i32, _p, _p, _i32, _p, _p, _p, _p]),
    "hb_tile_dev": (_i32, [_p, _i64, _i64, _i64, _i32, _p, _p]),
    "hb_untile_dev": (_i32, [_p, _i64, _i64, _i32, _p, _i64, _p]),
    "hb_hmm_binom_host": (_i32, [_p, _p, _i64, _i64, _p, _i32, _p, _d, _p, _p, _i32, _p, _p, _p]),
    "hb_hmm_binom_dev": (_i32, [_p, _p, _i64, _i64, _i64, _p, _i32, _p, _d, _p, _p, _i32, _p, _i64,
                                _p, _i64, _p, _p]),
    "hb_transpose_f64_dev": (_i32, [_p, _i64, _i64, _p, _i64, _p]),
    "hb_transpose_f64_to_i32_dev": (_i32, [_p, _i64, _i64, _p, _i64, _p]),
    "hb_transpose_u8_back_dev": (_i32, [_p, _i64, _i64, _i64, _p, _p]),
    "hb_transpose_f64_back_dev": (_i32, [_p, _i64, _i64, _i64, _p, _p]),
    "hb_chain_sd_dev": (_i32, [_p, _i64, _i64, _i64, _p, _i32, _i32, _p, _p]),
    "hb_pool_mean_dev": (_i32, [_p, _i64, _i64, _i64, _p, _i32, _p, _p]),
    "hb_pool_mean_relay_dev": (_i32, [_p, _i64, _i64, _i64, _p, _i32, _p, _i32, _p, _p, _p, _p]),
    "hb_pooled_sd_dev": (_i32, [_p, _i64, _i32, _p, _p]),
    "hb_pool_count_pos_dev": (_i32, [_p, _i64, _i64, _i64, _p, _i32, _p, _p]),
    "hb_runmean_dev": (_i32, [_p, _i64, _i64, _i64, _i32, _p, _i64, _p]),
    "hb_runmean_ratio_dev": (_i32, [_p, _p, _i64, _i64, _i64, _i32, _p, _i64, _p]),
    "hb_dist_sq_dev": (_i32, [_p, _i64, _i64, _i64, _p, _p]),
    "hb_dist_sq_rect_dev": (_i32, [_p, _i64, _i64, _p, _i64, _i64, _i64, _p, _i64, _p]),
    "hb_ward_linkage_dev": (_i32, [_p, _i64, _p, _p, _p, _p]),
    "hb_cutree_host": (_i32, [_p, _p, _p, _i64, _p, _i32, _p]),
    "hb_group_gexp_host": (_i32, [_p, _i64, _i64, _i32, _p, _i32, _p]),
    "hb_group_allele_host": (_i32, [_p, _p, _i64, _i64, _i32, _p, _i32, _p]),
    "hb_set_gexp_mats_dev": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i32, _d, _p, _p, _i32, _p, _i64,
                                    _p, _p, _p, _p]),
    "hb_set_gexp_mats_host": (_i32, [_p, _i64, _i64, _p, _i64, _i32, _d, _p, _p, _i32, _p, _p, _p, _p]),
    "hb_row_means_dev": (_i32, [_p, _i64, _i64, _i64, _p, _p]),
    "hb_col_scale_stats_dev": (_i32, [_p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "hb_gexp_apply_dev": (_i32, [_p, _i64, _i64, _p, _i64, _p, _p, _p, _p, _i64, _p]),
    "hb_mean_nonzero_dev": (_i32, [_p, _i64, _i64, _i64, _i32, _p, _p, _p]),
    "hb_set_allele_mats_dev": (_i32, [_p, _p, _i64, _i64, _i64, _p, _p, _i32, _d, _i32, _p, _p, _i64, _p, _p,
                                      _p, _p, _p, _p, _p]),
    "hb_set_allele_mats_host": (_i32, [_p, _p, _i64, _i64, _p, _p, _i32, _d, _i32, _p, _p, _p, _p, _p, _p,
                                       _p, _p]),
    "hb_retest_gexp_host": (_i32, [_p, _i64, _i64, _d, _p, _p, _p, _p]),
    "hb_retest_gexp_dev": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _d, _p, _p, _p, _p, _p, _p, _i32, _p]),
    "hb_retest_allele_model_dev": (_i32, [_p, _p, _i64, _i64, _i64, _p, _i64, _p, _p, _p, _d, _d, _i32, _p, _p]),
    "hb_retest_comb_dev": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _d, _p, _p, _p, _p, _p, _p, _p, _i32, _p]),
    "hb_retest_allele_model_host": (_i32, [_p, _p, _i64, _i64, _p, _p, _p, _d, _d, _p]),
    "hb_retest_allele_lr_dev": (_i32, [_p, _p, _i64, _i64, _i64, _p, _i64, _d, _d, _p, _p]),
    "hb_viterbi_long_logb_dev": (_i32, [_p, _i32, _i64, _i32, _p, _p, _i32, _p, _p]),
    "hb_forwardback_long_logb_dev": (_i32, [_p, _i32, _i64, _i32, _p, _p, _i32, _p, _p, _p]),
    "hb_longchain_stats_dev": (_i32, [_p, _p]),
    "hb_set_longchain_mode": (_i32, [_i32]),
    "hb_set_ward_mode": (_i32, [_i32]),
    "hb_set_dist_mode": (_i32, [_i32]),
    "hb_gexp_pooled_viterbi_dev": (_i32, [_p, _i64, _i64, _i64, _p, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "hb_allele_pooled_viterbi_dev": (_i32, [_p, _p, _i64, _i64, _i64, _p, _i32, _p, _p, _p, _p, _p,
                                            _p, _p]),
    "hb_gexp_pooled_viterbi_host": (_i32, [_p, _i64, _i64, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "hb_allele_pooled_viterbi_host": (_i32, [_p, _p, _i64, _i64, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "hb_upload_gexp_host": (_i32, [_p, _i64, _i64, C.POINTER(_p)]),
    "hb_upload_counts_host": (_i32, [_p, _p, _i64, _i64, C.POINTER(_p)]),
    "hb_free_handle": (_i32, [_p]),
    "hb_handle_info": (_i32, [_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                              C.POINTER(C.c_int64)]),
    "hb_gexp_round_handle": (_i32, [_p, _p, _i64, _p, _i64, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "hb_allele_round_handle": (_i32, [_p, _p, _i64, _p, _i64, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "hb_group_handle": (_i32, [_p, _p, _i64, _p, _i64, _i32, _p, _i32, _p]),
    "hb_retest_gexp_handle": (_i32, [_p, _p, _i64, _p, _i64, _d, _p, _p, _p, _p]),
    "hb_retest_allele_handle": (_i32, [_p, _p, _i64, _p, _i64, _p, _p, _p, _d, _d, _p]),
}

_lib = None


def load():
    """Load libhb_b200.so (built in-tree by honeybadger_b200.build); raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m honeybadger_b200.build` "
                "(nvcc, sm_100a).  honeybadger_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    if rc == HB_OK:
        return
    msg = load().hb_last_error().decode("utf-8", "replace")
    if rc == HB_ERR_UNDERFLOW:
        raise UnderflowError(rc, msg)
    raise HBError(rc, msg)


def device_count() -> int:
    n = C.c_int(0)
    rc = load().hb_device_count(C.byref(n))
    return n.value if rc == HB_OK else 0


def require_device():
    if device_count() < 1:
        raise HBError(HB_ERR_NODEVICE, "no CUDA device visible; honeybadger_b200 has no CPU fallback")
