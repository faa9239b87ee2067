"""
RecommenderBase -- host-side mirror of the reference's estimator base class
(matrix_factorization/recommender_base.py:14-271): sklearn BaseEstimator/RegressorMixin,
id maps, `_preprocess_data` (fit / update / predict modes) and `recommend`.

The SEMANTICS are the reference's, bit-exact for the id maps (first-appearance order on the
frame shuffled with numpy's global RNG, recommender_base.py:131-140); the IMPLEMENTATION works
on numpy arrays (factorize + one permutation) instead of per-row pandas maps so that it scales
to 10^8 rows, and `recommend` runs the batched scoring kernel instead of predict-all + sort.
"""
from __future__ import annotations

from abc import ABCMeta, abstractmethod
from typing import Any, Tuple, Union

import numpy as np
import pandas as pd
from sklearn.base import BaseEstimator, RegressorMixin


def _first_appearance_order(codes: np.ndarray, n_unique: int) -> np.ndarray:
    """rank[c] = position of code c in first-appearance order of `codes` (O(N), no hashing)."""
    n = len(codes)
    first_pos = np.full(n_unique, n, dtype=np.int64)
    # repeated indices: the LAST assignment wins, so assigning in reverse keeps the first position
    first_pos[codes[::-1]] = np.arange(n - 1, -1, -1, dtype=np.int64)
    order = np.argsort(first_pos, kind="stable")  # codes sorted by first appearance (absent ones last)
    rank = np.empty(n_unique, dtype=np.int64)
    rank[order] = np.arange(n_unique, dtype=np.int64)
    return rank, order, first_pos


def _has_duplicate_pairs(ucodes: np.ndarray, icodes: np.ndarray, n_items: int) -> bool:
    """Duplicate (user, item) check of recommender_base.py:127-128 on factorized codes."""
    n = len(ucodes)
    if n < 2:
        return False
    key = ucodes.astype(np.int64) * np.int64(max(n_items, 1)) + icodes.astype(np.int64)
    if n >= 2_000_000:
        try:  # large inputs: sort the keys on the device (plumbing, not the hot path)
            import torch

            if torch.cuda.is_available():
                k = torch.from_numpy(key).cuda()
                k, _ = torch.sort(k)
                return bool((k[1:] == k[:-1]).any().item())
        except Exception:
            pass
    key.sort()
    return bool((key[1:] == key[:-1]).any())


GPU_PREPROCESS_MIN_ROWS = 2_000_000


def _gpu_preprocess_wanted(users: np.ndarray, items: np.ndarray, n: int) -> bool:
    """The GPU id-mapping path serves integer raw ids from GPU_PREPROCESS_MIN_ROWS rows on (below that the host path is
    faster than the transfers); MFB_GPU_PREPROCESS=1 / 0 forces / forbids it."""
    import os

    env = os.environ.get("MFB_GPU_PREPROCESS")
    if env == "0" or n == 0:
        return False
    if users.dtype.kind not in "iu" or items.dtype.kind not in "iu" or users.dtype.itemsize > 8:
        return False
    if users.dtype == np.uint64 or items.dtype == np.uint64:
        return False
    if env == "1":
        return True
    if n < GPU_PREPROCESS_MIN_ROWS:
        return False
    from . import engine

    try:
        return bool(engine._torch().cuda.is_available())
    except Exception:
        return False


class RecommenderBase(BaseEstimator, RegressorMixin, metaclass=ABCMeta):
    """
    Abstract base of the recommenders (reference: recommender_base.py:14-95).

    Arguments:
        min_rating, max_rating -- rating bounds used when clipping predictions (defaults 0 / 5)
        verbose -- 1 prints one line per epoch while fitting (default 0 here, 1 in the subclasses)

    Attributes set by fit: n_users, n_items, global_mean, user_id_map, item_id_map.
    """

    @abstractmethod
    def __init__(self, min_rating: float = 0, max_rating: float = 5, verbose: int = 0):
        self.min_rating = min_rating
        self.max_rating = max_rating
        self.verbose = verbose
        return

    # recommender_base.py:53-95
    @property
    def known_users(self):
        return set(self.user_id_map.keys())

    @property
    def known_items(self):
        return set(self.item_id_map.keys())

    def contains_user(self, user_id: Any) -> bool:
        return user_id in self.known_users

    def contains_item(self, item_id: Any) -> bool:
        return item_id in self.known_items

    # ------------------------------------------------------------------ preprocessing
    def _preprocess_arrays(self, X: pd.DataFrame, y: pd.Series = None, type: str = "fit"):
        """
        Array form of `_preprocess_data` (recommender_base.py:97-173).  Returns a dict with
        u, i (int64 internal ids, -1 for unknown in predict mode), r (ratings or None) in the
        SHUFFLED row order the reference would hand to `_sgd`, plus known_users / new_users
        for type='update'.
        """
        X = X.loc[:, ["user_id", "item_id"]]
        if type != "predict":
            X["rating"] = y  # index-aligned like the reference (:123)
        users = X["user_id"].to_numpy()
        items = X["item_id"].to_numpy()
        ratings = X["rating"].to_numpy() if type != "predict" else None
        n = len(users)

        if type == "fit" and _gpu_preprocess_wanted(users, items, n):
            return self._preprocess_fit_gpu(users, items, ratings)

        if type in ("fit", "update"):
            ucodes0, uuniq0 = pd.factorize(users)
            icodes0, iuniq0 = pd.factorize(items)
            if _has_duplicate_pairs(ucodes0, icodes0, len(iuniq0)):  # :127-128
                raise ValueError("Duplicate user-item ratings in matrix")
            # :131  X.sample(frac=1) draws np.random.choice(n, n, replace=False) from the GLOBAL RNG
            perm = np.random.choice(n, size=n, replace=False) if n > 0 else np.zeros(0, dtype=np.int64)
            ucodes0, icodes0, ratings = ucodes0[perm], icodes0[perm], ratings[perm]

        if type == "fit":
            # :135-140  first-appearance ids on the shuffled rows
            urank, uorder, _ = _first_appearance_order(ucodes0, len(uuniq0))
            irank, iorder, _ = _first_appearance_order(icodes0, len(iuniq0))
            user_ids = uuniq0[uorder]
            item_ids = iuniq0[iorder]
            self.user_id_map = {k: j for j, k in enumerate(user_ids.tolist())}
            self.item_id_map = {k: j for j, k in enumerate(item_ids.tolist())}
            self.n_users = len(user_ids)
            self.n_items = len(item_ids)
            return {"u": urank[ucodes0], "i": irank[icodes0], "r": ratings}

        if type == "update":
            # :144-145 keep only ratings of known items
            imap = self.item_id_map
            item_internal = np.array([imap.get(k, -1) for k in iuniq0.tolist()], dtype=np.int64)
            keep = item_internal[icodes0] >= 0
            ucodes0, icodes0, ratings = ucodes0[keep], icodes0[keep], ratings[keep]
            # :148-160 users in first-appearance order of the shuffled, filtered rows
            urank, uorder, first_pos = _first_appearance_order(ucodes0, len(uuniq0))
            present = uorder[: int((first_pos < len(ucodes0)).sum())]
            new_users, known_users = [], []
            new_user_id = max(self.user_id_map.values()) + 1
            user_internal = np.full(len(uuniq0), -1, dtype=np.int64)
            for c in present.tolist():
                user = uuniq0[c]
                user = user.item() if hasattr(user, "item") else user
                if user in self.user_id_map:
                    known_users.append(user)
                    user_internal[c] = self.user_id_map[user]
                    continue
                new_users.append(user)
                self.user_id_map[user] = new_user_id
                user_internal[c] = new_user_id
                new_user_id += 1
            return {"u": user_internal[ucodes0], "i": item_internal[icodes0], "r": ratings,
                    "known_users": known_users, "new_users": new_users}

        # predict: :163-168 unknown ids -> -1
        ucodes, uuniq = pd.factorize(users)
        icodes, iuniq = pd.factorize(items)
        umap, imap = self.user_id_map, self.item_id_map
        uint = np.array([umap.get(k, -1) for k in uuniq.tolist()] + [-1], dtype=np.int64)
        iint = np.array([imap.get(k, -1) for k in iuniq.tolist()] + [-1], dtype=np.int64)
        return {"u": uint[ucodes], "i": iint[icodes], "r": None}  # code -1 (NaN id) hits the sentinel

    def _preprocess_fit_gpu(self, users: np.ndarray, items: np.ndarray, ratings: np.ndarray) -> dict:
        """type='fit' for integer raw ids on the GPU: the shuffle permutation is drawn on the host
        from numpy's global RNG exactly like the reference (:131), the first-appearance id assignment (:133-140) and the
        duplicate check (:125-128) run in mfk_first_appearance / mfk_has_duplicate_pairs.  Bit-exact with the host path."""
        import ctypes as C

        from . import engine
        from ._lib import check, lib, ptr, stream_ptr

        torch = engine._torch()
        n = len(users)
        dev = engine.device()
        d_users = torch.from_numpy(np.array(users, dtype=np.int64, copy=True)).to(dev)
        d_items = torch.from_numpy(np.array(items, dtype=np.int64, copy=True)).to(dev)
        # duplicates do not depend on the order: check before the RNG is touched, like the reference (:127 precedes :131)
        ui = torch.empty((n,), dtype=torch.int32, device=dev)
        ii = torch.empty((n,), dtype=torch.int32, device=dev)
        uq = torch.empty((n,), dtype=torch.int64, device=dev)
        iq = torch.empty((n,), dtype=torch.int64, device=dev)
        nu, ni, dup = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        check(lib().mfk_first_appearance(ptr(d_users), None, n, ptr(ui), ptr(uq), C.byref(nu), stream_ptr()))
        check(lib().mfk_first_appearance(ptr(d_items), None, n, ptr(ii), ptr(iq), C.byref(ni), stream_ptr()))
        check(lib().mfk_has_duplicate_pairs(ptr(ui), ptr(ii), n, max(1, ni.value), C.byref(dup), stream_ptr()))
        if dup.value:
            raise ValueError("Duplicate user-item ratings in matrix")
        perm = np.random.choice(n, size=n, replace=False)
        d_perm = torch.from_numpy(perm.astype(np.int64, copy=False)).to(dev)
        check(lib().mfk_first_appearance(ptr(d_users), ptr(d_perm), n, ptr(ui), ptr(uq), C.byref(nu), stream_ptr()))
        check(lib().mfk_first_appearance(ptr(d_items), ptr(d_perm), n, ptr(ii), ptr(iq), C.byref(ni), stream_ptr()))
        user_ids = uq[: nu.value].cpu().numpy()
        item_ids = iq[: ni.value].cpu().numpy()
        self.user_id_map = {k: j for j, k in enumerate(user_ids.tolist())}
        self.item_id_map = {k: j for j, k in enumerate(item_ids.tolist())}
        self.n_users, self.n_items = len(user_ids), len(item_ids)
        return {"u": ui.cpu().numpy().astype(np.int64), "i": ii.cpu().numpy().astype(np.int64), "r": ratings[perm]}

    def _preprocess_data(
        self, X: pd.DataFrame, y: pd.Series = None, type: str = "fit"
    ) -> Union[pd.DataFrame, Tuple[pd.DataFrame, list, list]]:
        """
        DataFrame form with the reference's signature and return values
        (recommender_base.py:97-173): X with columns user_id, item_id (internal ids) and rating;
        for type='update' also (known_users, new_users).
        """
        out = self._preprocess_arrays(X, y, type)
        cols = {"user_id": out["u"], "item_id": out["i"]}
        if out["r"] is not None:
            cols["rating"] = out["r"]
        frame = pd.DataFrame(cols)
        if type == "update":
            return frame, out["known_users"], out["new_users"]
        return frame

    @abstractmethod
    def fit(self, X: pd.DataFrame, y: pd.Series):
        return self

    @abstractmethod
    def predict(self, X: pd.DataFrame, bound_ratings: bool = True) -> list:
        return []

    # ------------------------------------------------------------------ recommend
    def _score_topk(self, user_internal: np.ndarray, k: int, mask_ptr, mask_items, bound_ratings: bool):
        """(scores [m,k] float64, items [m,k] int32 internal ids) from the scoring kernel."""
        raise NotImplementedError

    def _masked_internal(self, items_known) -> np.ndarray:
        imap = self.item_id_map
        out = [imap[x] for x in set(list(items_known)) if x in imap]  # unknown ids are ignored (:248-250)
        return np.array(sorted(out), dtype=np.int32)

    def recommend(
        self,
        user: Any,
        amount: int = 10,
        items_known: list = None,
        include_user: bool = True,
        bound_ratings: bool = True,
    ) -> pd.DataFrame:
        """
        Top `amount` items for `user`, best first (recommender_base.py:214-271): candidates = all
        known items minus `items_known`, ranked on the UNBOUNDED prediction, clipped after
        selection.  Columns: user_id (optional), item_id, rating_pred; the index is the item's
        position in the filtered candidate list, as in the reference.
        """
        masked = self._masked_internal(items_known) if items_known is not None else np.zeros(0, dtype=np.int32)
        n_cand = self.n_items_total() - len(masked)
        k = int(max(0, min(amount, n_cand)))
        item_ids = self._internal_to_raw_items()
        if k == 0:
            out = pd.DataFrame({"user_id": pd.Series([], dtype=object), "item_id": item_ids[:0],
                                "rating_pred": np.zeros(0)})
        else:
            uint = self.user_id_map.get(user, -1)
            if uint >= 0:
                mask_ptr = np.array([0, len(masked)], dtype=np.int64)
                scores, items = self._score_topk(np.array([uint], dtype=np.int32), k, mask_ptr, masked, bound_ratings)
                scores, items = scores[0], items[0].astype(np.int64)
            else:
                # unknown user: bias-only / zero-vector predictions through the predict kernel
                cand = np.setdiff1d(np.arange(self.n_items_total(), dtype=np.int64), masked, assume_unique=True)
                pred = np.asarray(self._predict_internal(np.full(len(cand), -1, dtype=np.int64), cand, False)[0])
                top = np.argsort(-pred, kind="stable")[:k]
                items, scores = cand[top], pred[top]
                if bound_ratings:
                    scores = np.clip(scores, self.min_rating, self.max_rating)
            position = items - np.searchsorted(masked, items)  # index in the filtered candidate list
            out = pd.DataFrame({"user_id": user, "item_id": item_ids[items], "rating_pred": scores}, index=position)
        if not include_user:
            out.drop(["user_id"], axis="columns", inplace=True)
        return out

    def predictor(self, capacity: int = 1024, bound_ratings: bool = True):
        """Low-latency predictor for small requests (one user x a few hundred items, project_template/app/api.py:43-52):
        parameters captured on the device, one CUDA-graph replay per request.  See serving.Predictor."""
        from .serving import Predictor

        return Predictor(self, capacity=capacity, bound_ratings=bound_ratings)

    def n_items_total(self) -> int:
        return len(self.item_id_map)

    def _internal_to_raw_items(self) -> np.ndarray:
        # The reference rebuilds list(item_id_map.keys()) on every call (recommender_base.py:245).  The cached
        # array is tied to the dict OBJECT it was built from: every fit() assigns a new item_id_map (same keys, a
        # different first-appearance order), which must never be served from the previous fit's cache.
        cache = getattr(self, "_raw_items_cache", None)
        if cache is None or cache[0] is not self.item_id_map or len(cache[1]) != len(self.item_id_map):
            keys = list(self.item_id_map.keys())
            arr = np.array(keys) if len(keys) else np.zeros(0, dtype=np.int64)
            if arr.ndim != 1:  # tuple-like ids
                arr = np.empty(len(keys), dtype=object)
                arr[:] = keys
            cache = (self.item_id_map, arr)
            self._raw_items_cache = cache
        return cache[1]

    def _internal_to_raw_users(self) -> np.ndarray:
        keys = list(self.user_id_map.keys())
        arr = np.array(keys) if len(keys) else np.zeros(0, dtype=np.int64)
        if arr.ndim != 1:
            arr = np.empty(len(keys), dtype=object)
            arr[:] = keys
        return arr

    def recommend_all(self, users=None, amount: int = 10, items_known: pd.DataFrame = None,
                      bound_ratings: bool = True) -> pd.DataFrame:
        """
        Batched recommend: the top `amount` items for every user in `users`
        (default: all known users) in one scoring pass.  `items_known` is a DataFrame with columns
        user_id, item_id (typically the training ratings); each user's known items are excluded.
        Returns a long DataFrame user_id, item_id, rating_pred, rank (0 = best).
        """
        raw_users = self._internal_to_raw_users() if users is None else np.asarray(list(users))
        umap = self.user_id_map
        uint = np.array([umap.get(u.item() if hasattr(u, "item") else u, -1) for u in raw_users], dtype=np.int64)
        if (uint < 0).any():
            raise ValueError("recommend_all: unknown user ids (use recommend() for cold-start users)")
        m = len(uint)
        mask_ptr = mask_items = None
        if items_known is not None and len(items_known):
            ku = pd.Series(items_known["user_id"].to_numpy()).map(umap).to_numpy(dtype=np.float64, na_value=-1.0)
            ki = pd.Series(items_known["item_id"].to_numpy()).map(self.item_id_map).to_numpy(dtype=np.float64, na_value=-1.0)
            ok = (ku >= 0) & (ki >= 0)
            ku, ki = ku[ok].astype(np.int64), ki[ok].astype(np.int64)
            # rows of the requested users, grouped in request order
            slot = np.full(max(self.user_id_map.values()) + 1, -1, dtype=np.int64)
            slot[uint] = np.arange(m)
            s = slot[ku]
            sel = s >= 0
            s, ki = s[sel], ki[sel]
            order = np.lexsort((ki, s))  # by user slot, item ids ascending inside a row (scoring kernel contract)
            mask_items = ki[order].astype(np.int32)
            mask_ptr = np.zeros(m + 1, dtype=np.int64)
            np.cumsum(np.bincount(s, minlength=m), out=mask_ptr[1:])
        k = int(max(1, min(amount, self.n_items_total())))
        scores, items = self._score_topk(uint.astype(np.int32), k, mask_ptr, mask_items, bound_ratings)
        valid = items.reshape(-1) >= 0
        raw_items = self._internal_to_raw_items()
        out = pd.DataFrame({
            "user_id": np.repeat(raw_users, k)[valid],
            "item_id": raw_items[items.reshape(-1)[valid]],
            "rating_pred": scores.reshape(-1)[valid],
            "rank": np.tile(np.arange(k), m)[valid],
        })
        return out
