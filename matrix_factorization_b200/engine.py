"""
Thin device layer between the estimators and the C ABI (include/mfk.h).

torch is plumbing only: it owns device memory (tensors) and the CUDA stream; every numeric
operation is a call into libmfk_b200.so.  No operation here has a CPU or torch fallback --
without a CUDA device (or without the built library) the calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KERNEL_IDS, PlanInfo, PlanOpts, check, lib, ptr, stream_ptr


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("matrix_factorization_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def device():
    torch = _torch()
    return torch.device("cuda", torch.cuda.current_device())


def round_up4(f: int) -> int:
    return (int(f) + 3) & ~3


def upload_rows(a: np.ndarray, ld: int | None = None):
    """[n, F] host array -> [n, ld] fp32 device tensor, zero-padded columns (ld % 4 == 0)."""
    torch = _torch()
    a = np.asarray(a)
    n, f = a.shape
    ld = round_up4(f) if ld is None else ld
    t = torch.zeros((n, ld), dtype=torch.float32, device=device())
    if n and f:
        src = torch.from_numpy(np.ascontiguousarray(a))
        t[:, :f].copy_(src.to(device(), non_blocking=False).to(torch.float32))
    return t


def upload_vec(a: np.ndarray, dtype=None):
    torch = _torch()
    dtype = dtype or torch.float32
    a = np.ascontiguousarray(np.asarray(a))
    if a.size == 0:
        return torch.zeros((0,), dtype=dtype, device=device())
    return torch.from_numpy(a).to(device()).to(dtype)


def download(t, dtype=np.float64, cols: int | None = None) -> np.ndarray:
    a = t.detach().cpu().numpy()
    if cols is not None:
        a = a[:, :cols]
    return np.array(a, dtype=dtype, order="C", copy=True)


class Plan:
    """Owner of an mfk_plan handle (stratified conflict-free SGD schedule)."""

    NO_HOT_SPLIT = 0xFFFFFFFF

    def __init__(self, u, i, r, n_users: int, n_items: int, n_factors: int = 0, n_workers: int = 0,
                 warps_per_cta: int = 0, hot_min_degree: int = NO_HOT_SPLIT, stripe_slack: int = 0,
                 schedule: int = 0, hot_users: bool = True):
        torch = _torch()
        assert u.dtype == torch.int32 and i.dtype == torch.int32 and r.dtype == torch.float32
        self._h = C.c_void_p()
        self.n = int(u.numel())
        opts = PlanOpts(int(n_workers), int(warps_per_cta), int(n_factors), int(hot_min_degree), int(stripe_slack), int(schedule),
                        0 if hot_users else 1)
        check(lib().mfk_plan_create(C.byref(self._h), ptr(u), ptr(i), ptr(r), self.n, int(n_users), int(n_items),
                                    C.byref(opts), stream_ptr()))

    @property
    def handle(self):
        return self._h

    def info(self) -> dict:
        inf = PlanInfo()
        check(lib().mfk_plan_get_info(self._h, C.byref(inf)))
        return inf.as_dict()

    def order(self):
        """int64 device tensor: a sequential order equivalent to one parallel epoch of this plan."""
        torch = _torch()
        out = torch.empty((self.n,), dtype=torch.int64, device=device())
        check(lib().mfk_plan_order(self._h, ptr(out), stream_ptr()))
        return out

    def assignment(self):
        torch = _torch()
        w = torch.empty((self.n,), dtype=torch.int32, device=device())
        s = torch.empty((self.n,), dtype=torch.int32, device=device())
        check(lib().mfk_plan_assignment(self._h, ptr(w), ptr(s), stream_ptr()))
        return w, s

    def set_phases(self, mask: int = 7):
        """Diagnostics: run only some phases of a split plan (bit 0 hot items, bit 1 hot users, bit 2 the rest)."""
        check(lib().mfk_plan_set_phases(self._h, int(mask)))

    def stats(self):
        """[n_workers, 4] int64 host array: cycles, blocked cycles, 4-chains, singles of the last epoch."""
        torch = _torch()
        info = self.info()
        W, H, HU = info["n_workers"], info["n_hot_workers"], info["n_hot_user_workers"]
        out = torch.zeros((12 * (W + H + HU),), dtype=torch.int64, device=device())
        check(lib().mfk_plan_stats(self._h, ptr(out), stream_ptr()))
        out = out.cpu().numpy()
        self.last_profile = out[4 * W:12 * W].reshape(W, 8)  # phase counters (MFK_RING_PROFILE builds only)
        hot = out[12 * W:12 * (W + H)]
        self.hot_stats = hot[:4 * H].reshape(H, 4)    # cycles, blocked cycles, batches, ratings per hot item
        self.hot_profile = hot[4 * H:].reshape(H, 8)  # phase cycles of the hot kernel
        hotu = out[12 * (W + H):]
        self.hot_user_stats = hotu[:4 * HU].reshape(HU, 4)
        self.hot_user_profile = hotu[4 * HU:].reshape(HU, 8)
        return out[:4 * W].reshape(W, 4)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().mfk_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Csr:
    """Owner of an mfk_csr handle (CSR by user + CSC by item)."""

    def __init__(self, u, i, r, n_users: int, n_items: int):
        _torch()
        self._h = C.c_void_p()
        self.n, self.n_users, self.n_items = int(u.numel()), int(n_users), int(n_items)
        check(lib().mfk_csr_create(C.byref(self._h), ptr(u), ptr(i), ptr(r), self.n, self.n_users, self.n_items,
                                   stream_ptr()))

    @property
    def handle(self):
        return self._h

    def export(self):
        """(row_ptr, col, val, col_ptr, row, cval) as device tensors (copies)."""
        torch = _torch()
        d = device()
        row_ptr = torch.empty((self.n_users + 1,), dtype=torch.int64, device=d)
        col_ptr = torch.empty((self.n_items + 1,), dtype=torch.int64, device=d)
        col = torch.empty((self.n,), dtype=torch.int32, device=d)
        row = torch.empty((self.n,), dtype=torch.int32, device=d)
        val = torch.empty((self.n,), dtype=torch.float32, device=d)
        cval = torch.empty((self.n,), dtype=torch.float32, device=d)
        check(lib().mfk_csr_export(self._h, ptr(row_ptr), ptr(col), ptr(val), ptr(col_ptr), ptr(row), ptr(cval),
                                   stream_ptr()))
        return row_ptr, col, val, col_ptr, row, cval

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().mfk_csr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_SSE_WS = {}


def _sse_ws():
    torch = _torch()
    dev = torch.cuda.current_device()
    if dev not in _SSE_WS:
        _SSE_WS[dev] = torch.empty((lib().mfk_sse_workspace_bytes(),), dtype=torch.uint8, device=device())
    return _SSE_WS[dev]


def kmf_sgd_epoch(plan: Plan, kernel: str, P, Q, bu, bi, n_factors, mu, lr, reg, gamma, lo, hi, upd_user=True,
                  upd_item=True):
    check(lib().mfk_kmf_sgd_epoch(plan.handle, KERNEL_IDS[kernel], ptr(P), ptr(Q), ptr(bu), ptr(bi), int(n_factors),
                                  int(P.shape[1]), float(mu), float(lr), float(reg), float(gamma), float(lo),
                                  float(hi), int(bool(upd_user)), int(bool(upd_item)), stream_ptr()))


def kmf_sse(kernel: str, u, i, r, P, Q, bu, bi, n_factors, mu, gamma, lo, hi, out):
    """out: 1-element float64 device tensor receiving sum((r - pred)^2)."""
    check(lib().mfk_kmf_sse(KERNEL_IDS[kernel], ptr(u), ptr(i), ptr(r), int(u.numel()), ptr(P), ptr(Q), ptr(bu),
                            ptr(bi), int(n_factors), int(P.shape[1]), float(mu), float(gamma), float(lo), float(hi),
                            ptr(_sse_ws()), ptr(out), stream_ptr()))


def kmf_sse_plan(plan: Plan, kernel: str, P, Q, bu, bi, n_factors, mu, gamma, lo, hi, out):
    check(lib().mfk_kmf_sse_plan(plan.handle, KERNEL_IDS[kernel], ptr(P), ptr(Q), ptr(bu), ptr(bi), int(n_factors),
                                 int(P.shape[1]), float(mu), float(gamma), float(lo), float(hi), ptr(_sse_ws()),
                                 ptr(out), stream_ptr()))


def kmf_predict(kernel: str, u, i, P, Q, bu, bi, n_factors, mu, gamma, lo, hi, bound):
    torch = _torch()
    n = int(u.numel())
    pred = torch.empty((n,), dtype=torch.float32, device=device())
    poss = torch.empty((n,), dtype=torch.uint8, device=device())
    check(lib().mfk_kmf_predict(KERNEL_IDS[kernel], ptr(u), ptr(i), n, ptr(P), ptr(Q), ptr(bu), ptr(bi),
                                int(n_factors), int(P.shape[1]), float(mu), float(gamma), float(lo), float(hi),
                                int(bool(bound)), ptr(pred), ptr(poss), stream_ptr()))
    return pred, poss


def bias_sgd_epoch(plan: Plan, bu, bi, mu, lr, reg, upd_user=True, upd_item=True):
    check(lib().mfk_bias_sgd_epoch(plan.handle, ptr(bu), ptr(bi), float(mu), float(lr), float(reg),
                                   int(bool(upd_user)), int(bool(upd_item)), stream_ptr()))


def bias_als_epoch(csr: Csr, bu, bi, mu, reg):
    check(lib().mfk_bias_als_epoch(csr.handle, ptr(bu), ptr(bi), float(mu), float(reg), stream_ptr()))


def bias_sse(u, i, r, bu, bi, mu, out):
    check(lib().mfk_bias_sse(ptr(u), ptr(i), ptr(r), int(u.numel()), ptr(bu), ptr(bi), float(mu), ptr(_sse_ws()),
                             ptr(out), stream_ptr()))


def bias_predict(u, i, bu, bi, mu, lo, hi, bound):
    torch = _torch()
    n = int(u.numel())
    pred = torch.empty((n,), dtype=torch.float32, device=device())
    poss = torch.empty((n,), dtype=torch.uint8, device=device())
    check(lib().mfk_bias_predict(ptr(u), ptr(i), n, ptr(bu), ptr(bi), float(mu), float(lo), float(hi),
                                 int(bool(bound)), ptr(pred), ptr(poss), stream_ptr()))
    return pred, poss


def score_topk(kernel: str, users, P, Q, bu, bi, n_items, n_factors, mu, gamma, lo, hi, k, bound, mask_ptr=None,
               mask_items=None):
    """users int32 [m]; optional CSR-style mask (int64 [m+1], int32 [nnz]).  Returns (scores [m,k] fp32,
    items [m,k] int32) device tensors; rows are padded with (-inf, -1)."""
    torch = _torch()
    m = int(users.numel())
    scores = torch.empty((m, k), dtype=torch.float32, device=device())
    items = torch.empty((m, k), dtype=torch.int32, device=device())
    ws = torch.empty((lib().mfk_score_workspace_bytes(m, int(n_items), int(n_factors), int(k)),), dtype=torch.uint8,
                     device=device())
    check(lib().mfk_score_topk(KERNEL_IDS[kernel], ptr(users), m, ptr(P), ptr(Q), ptr(bu), ptr(bi), int(n_items),
                               int(n_factors), int(P.shape[1]), float(mu), float(gamma), float(lo), float(hi),
                               ptr(mask_ptr), ptr(mask_items), int(k), int(bool(bound)), ptr(scores), ptr(items),
                               ptr(ws), stream_ptr()))
    return scores, items


def topk_merge(scores_in, items_in, k, bound, lo, hi):
    """[m, c] candidate lists (unbounded scores, global item ids, -1 padding) -> best k per row."""
    torch = _torch()
    m, c = scores_in.shape
    scores = torch.empty((m, k), dtype=torch.float32, device=device())
    items = torch.empty((m, k), dtype=torch.int32, device=device())
    check(lib().mfk_topk_merge(ptr(scores_in), ptr(items_in), int(m), int(c), int(k), int(bool(bound)), float(lo),
                               float(hi), ptr(scores), ptr(items), stream_ptr()))
    return scores, items
