"""
Small-batch prediction path for serving-style callers.

The reference's API handler predicts one user against a few hundred candidate items per request
(project_template/app/api.py:43-52: `model.predict(frame, bound_ratings=False)`); on the CPU that is ~0.3 ms.  The
general `predict()` of this package uploads the ids, re-checks the parameter mirrors and synchronises twice per call,
which a 500-row request does not amortise.  `Predictor` removes all of that from the request path:

* the fp32 device parameters are captured once (`refresh()` re-captures them after a refit / update_users);
* ids go through a pinned host buffer of fixed capacity, padded with -1 (= unknown id, which the predict kernel
  handles like the reference's `-1` rows, kernel_matrix_factorization.py:487-512);
* host->device copy, predict kernel and device->host copies are ONE captured CUDA graph, replayed per request on a
  private stream, followed by a single stream synchronisation.

Predictions equal `model.predict()` on the same rows (same kernel).
"""
from __future__ import annotations

import numpy as np

from . import _mirror, engine
from ._lib import check, lib, ptr


class Predictor:
    def __init__(self, model, capacity: int = 1024, bound_ratings: bool = True):
        torch = engine._torch()
        if capacity < 1:
            raise ValueError("capacity must be >= 1")
        self.model = model
        self._torch = torch
        self.capacity = int(capacity)
        self.bound = bool(bound_ratings)
        self._stream = torch.cuda.Stream()
        self._h_ids = torch.full((2, self.capacity), -1, dtype=torch.int32).pin_memory()
        self._ids_np = self._h_ids.numpy()
        self._d_ids = torch.full((2, self.capacity), -1, dtype=torch.int32, device=engine.device())
        self._d_pred = torch.zeros((self.capacity,), dtype=torch.float32, device=engine.device())
        self._d_poss = torch.zeros((self.capacity,), dtype=torch.uint8, device=engine.device())
        self._h_pred = torch.zeros((self.capacity,), dtype=torch.float32).pin_memory()
        self._h_poss = torch.zeros((self.capacity,), dtype=torch.uint8).pin_memory()
        self._pred_np, self._poss_np = self._h_pred.numpy(), self._h_poss.numpy()
        self._graph = None
        self.refresh()

    # ------------------------------------------------------------------ graph
    def _launch(self):
        m = self.model
        is_kmf = hasattr(m, "user_features")
        self._d_ids.copy_(self._h_ids, non_blocking=True)
        if is_kmf:
            check(lib().mfk_kmf_predict(engine.KERNEL_IDS[m.kernel], ptr(self._d_ids[0]), ptr(self._d_ids[1]), self.capacity,
                                        ptr(self._P), ptr(self._Q), ptr(self._bu), ptr(self._bi), int(m.n_factors),
                                        int(self._P.shape[1]), float(m.global_mean), float(m.gamma), float(m.min_rating),
                                        float(m.max_rating), int(self.bound), ptr(self._d_pred), ptr(self._d_poss),
                                        engine.stream_ptr()))
        else:
            check(lib().mfk_bias_predict(ptr(self._d_ids[0]), ptr(self._d_ids[1]), self.capacity, ptr(self._bu), ptr(self._bi),
                                         float(m.global_mean), float(m.min_rating), float(m.max_rating), int(self.bound),
                                         ptr(self._d_pred), ptr(self._d_poss), engine.stream_ptr()))
        self._h_pred.copy_(self._d_pred, non_blocking=True)
        self._h_poss.copy_(self._d_poss, non_blocking=True)

    def refresh(self):
        """(Re-)capture the model's current parameters and the request graph.  Call after fit() / update_users()."""
        torch = engine._torch()
        m = self.model
        if hasattr(m, "user_features"):
            self._P, self._Q = _mirror.rows(m.user_features), _mirror.rows(m.item_features)
        self._bu, self._bi = _mirror.vec(m.user_biases), _mirror.vec(m.item_biases)
        self._umap, self._imap = m.user_id_map, m.item_id_map
        torch.cuda.synchronize()
        with torch.cuda.stream(self._stream):
            self._launch()  # warm-up outside the capture (lazy module / context work)
        self._stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._stream):
            self._launch()
        self._graph = g

    # ------------------------------------------------------------------ requests
    def _run(self, n: int):
        torch = self._torch
        with torch.cuda.stream(self._stream):  # (a graph replays on the CURRENT stream)
            self._graph.replay()
        self._stream.synchronize()
        return self._pred_np[:n].astype(np.float64), self._poss_np[:n].astype(bool)

    def predict_pairs(self, user, items):
        """Predictions of ONE user for a list of item ids (raw ids).  Returns (predictions float64 [n], possible bool [n])."""
        n = len(items)
        if n > self.capacity:
            raise ValueError(f"{n} items exceed the predictor's capacity {self.capacity}")
        ids = self._ids_np
        ids[0, :n] = self._umap.get(user, -1)
        imap = self._imap
        ids[1, :n] = [imap.get(i, -1) for i in items]
        ids[:, n:] = -1
        return self._run(n)

    def predict(self, X) -> list:
        """Drop-in for `model.predict(X, bound_ratings=...)` on frames of up to `capacity` rows (larger frames are served
        in chunks).  Sets `model.predictions_possible` like the estimator does."""
        users, items = X["user_id"].tolist(), X["item_id"].tolist()
        out, poss = [], []
        umap, imap, ids = self._umap, self._imap, self._ids_np
        for s in range(0, len(users), self.capacity):
            n = min(self.capacity, len(users) - s)
            ids[0, :n] = [umap.get(u, -1) for u in users[s:s + n]]
            ids[1, :n] = [imap.get(i, -1) for i in items[s:s + n]]
            ids[:, n:] = -1
            p, q = self._run(n)
            out += p.tolist()
            poss += q.tolist()
        self.model.predictions_possible = poss
        return out
