"""
oracle/oracle.py -- Python face of the CPU oracle.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

* the numeric loops are the fp64 C restatement in mf_oracle.c (loaded with ctypes);
* the host-side logic (id mapping, update filtering, recommend) is restated here in
  numpy/pandas, each function citing the reference lines it follows.

Parity status: pinned -- tests/test_oracle_golden.py checks everything here against
fixtures generated from the reference itself by oracle/gen_golden.py.
"""
import ctypes
import os
import subprocess

import numpy as np
import pandas as pd

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KERNELS = {"linear": 0, "sigmoid": 1, "rbf": 2}

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_c = ctypes


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (seconds).  ORACLE_FAST=1 (bench.py's CPU arm) builds liboracle_fast.so with
    -O3 -march=native instead; floating-point contraction and fast-math stay off, so the arithmetic is the same."""
    fast = os.environ.get("ORACLE_FAST") == "1"
    name = "liboracle.so"
    if fast:  # -march=native code must not travel to another CPU: the file name carries the host CPU's signature
        import hashlib
        try:
            sig = "".join(l for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags")))[:20000]
        except OSError:
            sig = "unknown"
        name = "liboracle_fast_%s.so" % hashlib.sha1(sig.encode()).hexdigest()[:10]
    so = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "mf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        opt = ["-O3", "-march=native", "-ffp-contract=off"] if fast else ["-O2"]
        subprocess.check_call(["gcc"] + opt + ["-fPIC", "-std=c11", "-fno-fast-math", "-shared", "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_kmf_replay.argtypes = [
            _c.c_int, _i32p, _i32p, _f64p, _c.c_void_p, _c.c_int64, _c.c_double, _f64p, _f64p,
            _f64p, _f64p, _c.c_int, _c.c_double, _c.c_double, _c.c_double, _c.c_double,
            _c.c_double, _c.c_int, _c.c_int,
        ]
        L.orc_kmf_replay.restype = None
        L.orc_kmf_rmse.argtypes = [
            _c.c_int, _i32p, _i32p, _f64p, _c.c_int64, _c.c_double, _f64p, _f64p, _f64p, _f64p,
            _c.c_int, _c.c_double, _c.c_double, _c.c_double,
        ]
        L.orc_kmf_rmse.restype = _c.c_double
        L.orc_kmf_sgd.argtypes = [
            _c.c_int, _i32p, _i32p, _f64p, _c.c_int64, _c.c_double, _f64p, _f64p, _f64p, _f64p,
            _c.c_int, _c.c_int, _c.c_double, _c.c_double, _c.c_double, _c.c_double, _c.c_double,
            _c.c_int, _c.c_int, _c.c_uint64, _f64p,
        ]
        L.orc_kmf_sgd.restype = None
        L.orc_kmf_predict.argtypes = [
            _c.c_int, _i32p, _i32p, _c.c_int64, _c.c_double, _f64p, _f64p, _f64p, _f64p,
            _c.c_int, _c.c_double, _c.c_double, _c.c_double, _c.c_int, _f64p, _u8p,
        ]
        L.orc_kmf_predict.restype = None
        L.orc_bias_rmse.argtypes = [_i32p, _i32p, _f64p, _c.c_int64, _c.c_double, _f64p, _f64p]
        L.orc_bias_rmse.restype = _c.c_double
        L.orc_bias_replay.argtypes = [
            _i32p, _i32p, _f64p, _c.c_void_p, _c.c_int64, _c.c_double, _f64p, _f64p,
            _c.c_double, _c.c_double, _c.c_int, _c.c_int,
        ]
        L.orc_bias_replay.restype = None
        L.orc_bias_sgd.argtypes = [
            _i32p, _i32p, _f64p, _c.c_int64, _c.c_double, _f64p, _f64p, _c.c_int, _c.c_double,
            _c.c_double, _c.c_int, _c.c_int, _c.c_uint64, _f64p,
        ]
        L.orc_bias_sgd.restype = None
        L.orc_bias_als.argtypes = [
            _i32p, _i32p, _f64p, _c.c_int64, _c.c_double, _f64p, _f64p, _c.c_int, _c.c_int,
            _c.c_int, _c.c_double, _f64p,
        ]
        L.orc_bias_als.restype = None
        L.orc_bias_predict.argtypes = [
            _i32p, _i32p, _c.c_int64, _c.c_double, _f64p, _f64p, _c.c_double, _c.c_double,
            _c.c_int, _f64p, _u8p,
        ]
        L.orc_bias_predict.restype = None
        _LIB = L
    return _LIB


def _ids(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _order_ptr(order):
    if order is None:
        return None, None
    o = np.ascontiguousarray(order, dtype=np.int64)
    return o, o.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------ KernelMF loops
def kmf_replay(kernel, u, i, r, order, mu, bu, bi, P, Q, lr, reg, gamma=0.01, min_rating=0.0,
               max_rating=5.0, update_user_params=True, update_item_params=True):
    """Apply the reference update rule in an explicit order; returns new (P, Q, bu, bi) fp64 copies.
    Restates kernel_matrix_factorization.py:374-425 + kernels.py:108-327."""
    u, i, r = _ids(u), _ids(i), _f64(r)
    P, Q, bu, bi = (np.array(x, dtype=np.float64, order="C", copy=True) for x in (P, Q, bu, bi))
    keep, optr = _order_ptr(order)
    n = len(u) if order is None else len(keep)
    F = P.shape[1] if P.ndim == 2 else 0
    lib().orc_kmf_replay(KERNELS[kernel], u, i, r, optr, n, float(mu), bu, bi, P.reshape(-1),
                         Q.reshape(-1), F, float(lr), float(reg), float(gamma), float(min_rating),
                         float(max_rating), int(update_user_params), int(update_item_params))
    return P, Q, bu, bi


def kmf_rmse(kernel, u, i, r, mu, bu, bi, P, Q, gamma=0.01, min_rating=0.0, max_rating=5.0):
    """kernel_matrix_factorization.py:240-317."""
    P, Q = _f64(P), _f64(Q)
    return float(lib().orc_kmf_rmse(KERNELS[kernel], _ids(u), _ids(i), _f64(r), len(u), float(mu),
                                    _f64(bu), _f64(bi), P.reshape(-1), Q.reshape(-1), P.shape[1],
                                    float(gamma), float(min_rating), float(max_rating)))


def kmf_sgd(kernel, u, i, r, mu, bu, bi, P, Q, n_epochs, lr, reg, gamma=0.01, min_rating=0.0,
            max_rating=5.0, update_user_params=True, update_item_params=True, seed=1):
    """Full port of _sgd (kernel_matrix_factorization.py:320-445) with its own shuffle stream.
    Returns (P, Q, bu, bi, train_rmse)."""
    u, i, r = (np.array(x, copy=True) for x in (_ids(u), _ids(i), _f64(r)))
    P, Q, bu, bi = (np.array(x, dtype=np.float64, order="C", copy=True) for x in (P, Q, bu, bi))
    rm = np.zeros(n_epochs, dtype=np.float64)
    lib().orc_kmf_sgd(KERNELS[kernel], u, i, r, len(u), float(mu), bu, bi, P.reshape(-1),
                      Q.reshape(-1), P.shape[1], int(n_epochs), float(lr), float(reg),
                      float(gamma), float(min_rating), float(max_rating), int(update_user_params),
                      int(update_item_params), int(seed), rm)
    return P, Q, bu, bi, rm.tolist()


def kmf_predict(kernel, u, i, mu, bu, bi, P, Q, gamma=0.01, min_rating=0.0, max_rating=5.0,
                bound_ratings=True):
    """kernel_matrix_factorization.py:448-541; ids of -1 are unknown."""
    u, i = _ids(u), _ids(i)
    P, Q = _f64(P), _f64(Q)
    pred = np.zeros(len(u), dtype=np.float64)
    poss = np.zeros(len(u), dtype=np.uint8)
    lib().orc_kmf_predict(KERNELS[kernel], u, i, len(u), float(mu), _f64(bu), _f64(bi),
                          P.reshape(-1), Q.reshape(-1), P.shape[1], float(gamma),
                          float(min_rating), float(max_rating), int(bound_ratings), pred, poss)
    return pred, poss.astype(bool)


# ------------------------------------------------------------------ BaselineModel loops
def bias_rmse(u, i, r, mu, bu, bi):
    """baseline_model.py:183-212."""
    return float(lib().orc_bias_rmse(_ids(u), _ids(i), _f64(r), len(u), float(mu), _f64(bu), _f64(bi)))


def bias_replay(u, i, r, order, mu, bu, bi, lr, reg, update_user_params=True,
                update_item_params=True):
    """baseline_model.py:255-266 in an explicit order."""
    bu, bi = (np.array(x, dtype=np.float64, copy=True) for x in (bu, bi))
    keep, optr = _order_ptr(order)
    n = len(u) if order is None else len(keep)
    lib().orc_bias_replay(_ids(u), _ids(i), _f64(r), optr, n, float(mu), bu, bi, float(lr),
                          float(reg), int(update_user_params), int(update_item_params))
    return bu, bi


def bias_sgd(u, i, r, mu, bu, bi, n_epochs, lr, reg, update_user_params=True,
             update_item_params=True, seed=1):
    """baseline_model.py:215-280."""
    u, i, r = (np.array(x, copy=True) for x in (_ids(u), _ids(i), _f64(r)))
    bu, bi = (np.array(x, dtype=np.float64, copy=True) for x in (bu, bi))
    rm = np.zeros(n_epochs, dtype=np.float64)
    lib().orc_bias_sgd(u, i, r, len(u), float(mu), bu, bi, int(n_epochs), float(lr), float(reg),
                       int(update_user_params), int(update_item_params), int(seed), rm)
    return bu, bi, rm.tolist()


def bias_als(u, i, r, mu, n_users, n_items, n_epochs, reg, bu=None, bi=None):
    """baseline_model.py:283-362."""
    bu = np.zeros(n_users) if bu is None else np.array(bu, dtype=np.float64, copy=True)
    bi = np.zeros(n_items) if bi is None else np.array(bi, dtype=np.float64, copy=True)
    rm = np.zeros(n_epochs, dtype=np.float64)
    lib().orc_bias_als(_ids(u), _ids(i), _f64(r), len(u), float(mu), bu, bi, int(n_users),
                       int(n_items), int(n_epochs), float(reg), rm)
    return bu, bi, rm.tolist()


def bias_predict(u, i, mu, bu, bi, min_rating=0.0, max_rating=5.0, bound_ratings=True):
    """baseline_model.py:365-417."""
    u, i = _ids(u), _ids(i)
    pred = np.zeros(len(u), dtype=np.float64)
    poss = np.zeros(len(u), dtype=np.uint8)
    lib().orc_bias_predict(u, i, len(u), float(mu), _f64(bu), _f64(bi), float(min_rating),
                           float(max_rating), int(bound_ratings), pred, poss)
    return pred, poss.astype(bool)


# ------------------------------------------------------------------ host logic restatements
def preprocess_fit(user_ids, item_ids, ratings):
    """recommender_base.py:120-164 for type='fit' on positional arrays: duplicate check,
    row shuffle drawn from numpy's GLOBAL RNG (DataFrame.sample(frac=1) ==
    np.random.choice(n, n, replace=False) == permutation), first-appearance id maps on the
    SHUFFLED rows.  Returns (u_int, i_int, r_shuffled, user_id_map, item_id_map, perm)."""
    user_ids, item_ids = np.asarray(user_ids), np.asarray(item_ids)
    df = pd.DataFrame({"user_id": user_ids, "item_id": item_ids})
    if df.duplicated(subset=["user_id", "item_id"]).sum() != 0:  # :127-128
        raise ValueError("Duplicate user-item ratings in matrix")
    n = len(user_ids)
    perm = np.random.choice(n, size=n, replace=False)  # :131 (what DataFrame.sample draws)
    us, it = user_ids[perm], item_ids[perm]
    ucodes, uuniq = pd.factorize(us)  # first-appearance order, :135-140
    icodes, iuniq = pd.factorize(it)
    umap = {k: j for j, k in enumerate(uuniq.tolist())}
    imap = {k: j for j, k in enumerate(iuniq.tolist())}
    return ucodes.astype(np.int64), icodes.astype(np.int64), np.asarray(ratings)[perm], umap, imap, perm


def map_predict(user_ids, item_ids, user_id_map, item_id_map):
    """recommender_base.py:163-168 for type='predict': unknown ids -> -1."""
    u = np.array([user_id_map.get(x, -1) for x in np.asarray(user_ids).tolist()], dtype=np.int64)
    i = np.array([item_id_map.get(x, -1) for x in np.asarray(item_ids).tolist()], dtype=np.int64)
    return u, i


def recommend_candidates(item_id_map, items_known):
    """recommender_base.py:245-250: all known items in internal-id order minus items_known."""
    items = list(item_id_map.keys())
    if items_known is not None:
        known = set(list(items_known))
        items = [x for x in items if x not in known]
    return items


def topk_desc(scores, k):
    """Reference ordering of recommender_base.py:259-260 (sort desc, head) as a stable rule:
    highest score first, ties by lower candidate position."""
    scores = np.asarray(scores, dtype=np.float64)
    idx = np.argsort(-scores, kind="stable")[:k]
    return idx, scores[idx]
