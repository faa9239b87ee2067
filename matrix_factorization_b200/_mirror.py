"""
Device mirrors of the estimators' host arrays.

The public attributes (user_features, item_features, user_biases, item_biases) stay float64
numpy arrays like the reference's.  Uploading them on every predict()/recommend() call would
dominate small requests, so the fp32 device copy made by the last fit/update/predict is kept
here, keyed by the identity of the numpy array (weak reference: the mirror dies with the
array, and is never pickled).

The reference reads the host arrays on every call, so in-place edits by user code
(`model.item_features[j] = ...`, re-initialising P/Q before another `_sgd`) must not be served
from a stale mirror: every lookup re-checks a content fingerprint of the host array -- the
exact sum / sum of squares for arrays up to 2^18 elements, a 65 536-element strided sample
plus both ends beyond that -- and re-uploads on a mismatch.  `invalidate(array)` forces the
re-upload (needed only for edits of a large array that miss every sampled element).
"""
from __future__ import annotations

import weakref

import numpy as np

from . import engine

_ROWS: dict[int, tuple] = {}
_VECS: dict[int, tuple] = {}


_FULL_CHECK_MAX = 1 << 18
_SAMPLE = 1 << 16


def fingerprint(arr: np.ndarray) -> tuple:
    """Cheap content fingerprint of a host array (see the module docstring)."""
    a = np.asarray(arr)
    n = a.size
    if n == 0:
        return (a.shape, 0.0, 0.0, 0.0, 0.0)
    flat = a.reshape(-1) if a.flags.c_contiguous else a.ravel()
    s = flat if n <= _FULL_CHECK_MAX else flat[:: max(1, n // _SAMPLE)]
    s = s.astype(np.float64, copy=False)
    return (a.shape, float(s.sum()), float(np.dot(s, s)), float(flat[0]), float(flat[-1]))


def _register(table, arr: np.ndarray, tensor):
    key = id(arr)

    def _drop(_ref, key=key, table=table):
        table.pop(key, None)

    table[key] = (weakref.ref(arr, _drop), tensor, fingerprint(arr))


def _lookup(table, arr: np.ndarray):
    ent = table.get(id(arr))
    if ent is not None and ent[0]() is arr and ent[2] == fingerprint(arr):
        return ent[1]
    return None


def rows(arr: np.ndarray):
    """fp32 [n, ld] device mirror of a float [n, F] host array (uploaded on first use)."""
    t = _lookup(_ROWS, arr)
    if t is None or t.shape[0] != arr.shape[0] or t.shape[1] != engine.round_up4(arr.shape[1]):
        t = engine.upload_rows(arr)
        _register(_ROWS, arr, t)
    return t


def vec(arr: np.ndarray):
    t = _lookup(_VECS, arr)
    if t is None or t.shape[0] != arr.shape[0]:
        t = engine.upload_vec(arr)
        _register(_VECS, arr, t)
    return t


def set_rows(arr: np.ndarray, tensor):
    _register(_ROWS, arr, tensor)


def set_vec(arr: np.ndarray, tensor):
    _register(_VECS, arr, tensor)


def invalidate(arr: np.ndarray):
    _ROWS.pop(id(arr), None)
    _VECS.pop(id(arr), None)
