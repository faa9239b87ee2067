"""
ctypes binding of the C ABI declared in include/mfk.h (libmfk_b200.so, sm_100a).

There is no CPU or alternative-backend fallback: `lib()` raises if the shared library is
missing or cannot be loaded, and every wrapper raises RuntimeError on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (MFK_LIB_PATH: another build of the same library, for A/B experiments of kernel variants)
LIB_PATH = os.environ.get("MFK_LIB_PATH") or os.path.join(_HERE, "csrc", "libmfk_b200.so")

KERNEL_IDS = {"linear": 0, "sigmoid": 1, "rbf": 2}
ERR_NAMES = {1: "MFK_ERR_ARG", 2: "MFK_ERR_CUDA", 3: "MFK_ERR_UNSUPPORTED", 4: "MFK_ERR_NO_DEVICE"}


class MfkError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{ERR_NAMES.get(status, status)}: {message}")
        self.status = status


class PlanOpts(C.Structure):
    _fields_ = [("n_workers", C.c_int32), ("warps_per_cta", C.c_int32), ("n_factors", C.c_int32),
                ("hot_min_degree", C.c_uint32), ("stripe_slack", C.c_uint32), ("schedule", C.c_uint32), ("no_hot_users", C.c_uint32)]


class PlanInfo(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_users", C.c_int32), ("n_items", C.c_int32), ("n_workers", C.c_int32),
                ("n_ctas", C.c_int32), ("warps_per_cta", C.c_int32), ("max_items_per_worker", C.c_int32),
                ("max_worker_ratings", C.c_int64), ("max_item_degree", C.c_int64), ("max_user_degree", C.c_int64),
                ("n_hot_items", C.c_int32), ("n_steps", C.c_int32), ("n_hot_ratings", C.c_int64),
                ("n_hot_users", C.c_int32), ("flat", C.c_int32), ("n_hot_user_ratings", C.c_int64),
                ("n_hot_workers", C.c_int32), ("n_hot_user_workers", C.c_int32), ("hot_max_slots", C.c_int32),
                ("hot_parallel", C.c_int32)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_p = C.c_void_p
_i32, _i64, _f32, _int = C.c_int32, C.c_int64, C.c_float, C.c_int

# name -> (restype, argtypes); mirrors include/mfk.h one-for-one
SIGNATURES = {
    "mfk_last_error": (C.c_char_p, []),
    "mfk_abi_version": (_int, []),
    "mfk_device_query": (_int, [_int, C.POINTER(_int), C.POINTER(_int), C.POINTER(C.c_size_t)]),
    "mfk_plan_create": (_int, [C.POINTER(_p), _p, _p, _p, _i64, _i32, _i32, C.POINTER(PlanOpts), _p]),
    "mfk_plan_destroy": (_int, [_p]),
    "mfk_plan_get_info": (_int, [_p, C.POINTER(PlanInfo)]),
    "mfk_plan_order": (_int, [_p, _p, _p]),
    "mfk_plan_assignment": (_int, [_p, _p, _p, _p]),
    "mfk_plan_stats": (_int, [_p, _p, _p]),
    "mfk_plan_set_phases": (_int, [_p, C.c_uint32]),
    "mfk_kmf_sgd_epoch": (_int, [_p, _int, _p, _p, _p, _p, _i32, _i32, _f32, _f32, _f32, _f32, _f32, _f32, _int,
                                 _int, _p]),
    "mfk_sse_workspace_bytes": (C.c_size_t, []),
    "mfk_kmf_sse": (_int, [_int, _p, _p, _p, _i64, _p, _p, _p, _p, _i32, _i32, _f32, _f32, _f32, _f32, _p, _p, _p]),
    "mfk_kmf_sse_plan": (_int, [_p, _int, _p, _p, _p, _p, _i32, _i32, _f32, _f32, _f32, _f32, _p, _p, _p]),
    "mfk_kmf_predict": (_int, [_int, _p, _p, _i64, _p, _p, _p, _p, _i32, _i32, _f32, _f32, _f32, _f32, _int, _p, _p,
                               _p]),
    "mfk_bias_sgd_epoch": (_int, [_p, _p, _p, _f32, _f32, _f32, _int, _int, _p]),
    "mfk_csr_create": (_int, [C.POINTER(_p), _p, _p, _p, _i64, _i32, _i32, _p]),
    "mfk_csr_destroy": (_int, [_p]),
    "mfk_csr_export": (_int, [_p, _p, _p, _p, _p, _p, _p, _p]),
    "mfk_bias_als_epoch": (_int, [_p, _p, _p, _f32, _f32, _p]),
    "mfk_bias_sse": (_int, [_p, _p, _p, _i64, _p, _p, _f32, _p, _p, _p]),
    "mfk_bias_predict": (_int, [_p, _p, _i64, _p, _p, _f32, _f32, _f32, _int, _p, _p, _p]),
    "mfk_score_workspace_bytes": (C.c_size_t, [_i64, _i32, _i32, _i32]),
    "mfk_score_topk": (_int, [_int, _p, _i64, _p, _p, _p, _p, _i32, _i32, _i32, _f32, _f32, _f32, _f32, _p, _p, _i32,
                              _int, _p, _p, _p, _p]),
    "mfk_topk_merge": (_int, [_p, _p, _i64, _i32, _i32, _int, _f32, _f32, _p, _p, _p]),
    "mfk_first_appearance": (_int, [_p, _p, _i64, _p, _p, _p, _p]),
    "mfk_has_duplicate_pairs": (_int, [_p, _p, _i64, _i64, _p, _p]),
    "mfk_kmf_sgd_host": (_int, [_int, _p, _p, _p, _i64, _i32, _i32, _p, _p, _p, _p, _i32, _i32, _f32, _i32, _f32,
                                _f32, _f32, _f32, _f32, _int, _int, C.POINTER(PlanOpts), _p, _p]),
    "mfk_bias_sgd_host": (_int, [_p, _p, _p, _i64, _i32, _i32, _p, _p, _f32, _i32, _f32, _f32, _int, _int,
                                 C.POINTER(PlanOpts), _p, _p]),
    "mfk_bias_als_host": (_int, [_p, _p, _p, _i64, _i32, _i32, _p, _p, _f32, _i32, _f32, _p]),
}

_LIB = None


def lib():
    """Load libmfk_b200.so (once).  Fails loudly: there is no fallback implementation."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m matrix_factorization_b200.build` "
                "(or __graft_entry__.build()).  matrix_factorization_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here means the .so does not match include/mfk.h
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(status: int):
    if status != 0:
        raise MfkError(status, lib().mfk_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array / None as c_void_p."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def stream_ptr(stream=None):
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)
