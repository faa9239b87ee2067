"""
BaselineModel -- drop-in for matrix_factorization/baseline_model.py of the reference:
r_ui ~ mu + b_u + b_i, fitted by bias SGD (stratified conflict-free schedule) or by ALS
(two segmented reductions per epoch over a CSR/CSC layout built once on the GPU).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import pandas as pd

from . import _mirror, engine
from .kernel_matrix_factorization import _split_X, _upload_ratings
from .recommender_base import RecommenderBase


class BaselineModel(RecommenderBase):
    """
    Global mean + user bias + item bias (reference docstring: baseline_model.py:10-39).

    Arguments:
        method {str} -- 'sgd' or 'als' (default: 'sgd')
        n_epochs {int} -- epochs (default: 100)
        reg {float} -- L2 regularisation (default: 1)
        lr {float} -- learning rate, sgd only (default: 0.01)
        min_rating, max_rating -- rating range (default: 0, 5)
        verbose -- 1 prints one line per epoch (default: 1)

    Attributes: n_users, n_items, global_mean, user_biases, item_biases, user_id_map,
        item_id_map, train_rmse, predictions_possible (after predict).
    """

    def __init__(
        self,
        method: str = "sgd",
        n_epochs: int = 100,
        reg: float = 1,
        lr: float = 0.01,
        min_rating: int = 0,
        max_rating: int = 5,
        verbose=1,
    ):
        if method not in ("sgd", "als"):
            raise ValueError('Method param must be either "sgd" or "als"')

        super().__init__(min_rating=min_rating, max_rating=max_rating, verbose=verbose)

        self.method = method
        self.n_epochs = n_epochs
        self.reg = reg
        self.lr = lr
        return

    def fit(self, X: pd.DataFrame, y: pd.Series):
        """Fit the mean + bias model (baseline_model.py:63-102)."""
        data = self._preprocess_arrays(X=X, y=y, type="fit")
        self.global_mean = pd.Series(data["r"]).mean()
        self.user_biases = np.zeros(self.n_users)
        self.item_biases = np.zeros(self.n_items)
        Xa = (data["u"], data["i"], data["r"])
        if self.method == "sgd":
            self.user_biases, self.item_biases, self.train_rmse = _sgd(
                X=Xa,
                global_mean=self.global_mean,
                user_biases=self.user_biases,
                item_biases=self.item_biases,
                n_epochs=self.n_epochs,
                lr=self.lr,
                reg=self.reg,
                verbose=self.verbose,
            )
        elif self.method == "als":
            self.user_biases, self.item_biases, self.train_rmse = _als(
                X=Xa,
                global_mean=self.global_mean,
                user_biases=self.user_biases,
                item_biases=self.item_biases,
                n_epochs=self.n_epochs,
                reg=self.reg,
                verbose=self.verbose,
            )
        return self

    def _predict_internal(self, u, i, bound_ratings):
        return _predict(
            X=(u, i),
            global_mean=self.global_mean,
            min_rating=self.min_rating,
            max_rating=self.max_rating,
            user_biases=self.user_biases,
            item_biases=self.item_biases,
            bound_ratings=bound_ratings,
        )

    def predict(self, X: pd.DataFrame, bound_ratings: bool = True) -> list:
        """Predicted ratings in the order of X (baseline_model.py:104-134)."""
        if X.shape[0] == 0:
            return []
        data = self._preprocess_arrays(X=X, type="predict")
        predictions, predictions_possible = self._predict_internal(data["u"], data["i"], bound_ratings)
        self.predictions_possible = predictions_possible
        return predictions

    def update_users(
        self,
        X: pd.DataFrame,
        y: pd.Series,
        lr: float = 0.01,
        n_epochs: int = 20,
        verbose: int = 0,
    ):
        """User biases of new / re-passed users by SGD with item biases frozen (baseline_model.py:136-180)."""
        data = self._preprocess_arrays(X=X, y=y, type="update")
        for user in data["known_users"]:
            self.user_biases[self.user_id_map[user]] = 0
        if data["known_users"]:
            _mirror.invalidate(self.user_biases)
        self.user_biases = np.append(self.user_biases, np.zeros(len(data["new_users"])))
        self.user_biases, _, self.train_rmse = _sgd(
            X=(data["u"], data["i"], data["r"]),
            global_mean=self.global_mean,
            user_biases=self.user_biases,
            item_biases=self.item_biases,
            n_epochs=n_epochs,
            lr=lr,
            reg=self.reg,
            verbose=verbose,
            update_item_params=False,
        )
        return

    def _score_topk(self, user_internal, k, mask_ptr, mask_items, bound_ratings):
        # linear kernel with zero factors: rank key = b_i (the most popular items, as the reference notes)
        torch = engine._torch()
        dev = engine.device()
        P = torch.zeros((len(self.user_biases), 4), dtype=torch.float32, device=dev)
        Q = torch.zeros((len(self.item_biases), 4), dtype=torch.float32, device=dev)
        bu, bi = _mirror.vec(self.user_biases), _mirror.vec(self.item_biases)
        users = engine.upload_vec(np.asarray(user_internal, dtype=np.int32), torch.int32)
        mp = engine.upload_vec(np.asarray(mask_ptr, dtype=np.int64), torch.int64) if mask_ptr is not None else None
        mi = None
        if mask_ptr is not None:
            mi = engine.upload_vec(np.asarray(mask_items, dtype=np.int32), torch.int32)
            if mi.numel() == 0:
                mi = torch.zeros((1,), dtype=torch.int32, device=dev)
        scores, items = engine.score_topk("linear", users, P, Q, bu, bi, len(self.item_biases), 4, self.global_mean,
                                          0.0, self.min_rating, self.max_rating, k, bound_ratings, mp, mi)
        return scores.cpu().numpy().astype(np.float64), items.cpu().numpy()

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_raw_items_cache", None)
        return state


# ---------------------------------------------------------------------------------------
# Operator boundary (same names / arguments / results as the reference's njit functions)
# ---------------------------------------------------------------------------------------
def _calculate_rmse(X, global_mean: float, user_biases: np.ndarray, item_biases: np.ndarray):
    """baseline_model.py:183-212."""
    torch = engine._torch()
    u, i, r = _split_X(X)
    n = len(u)
    if n == 0:
        return float("nan")
    du, di, dr = _upload_ratings(u, i, r)
    out = torch.zeros((1,), dtype=torch.float64, device=engine.device())
    engine.bias_sse(du, di, dr, _mirror.vec(user_biases), _mirror.vec(item_biases), global_mean, out)
    return math.sqrt(float(out.item()) / n)


def _sgd(
    X,
    global_mean: float,
    user_biases: np.ndarray,
    item_biases: np.ndarray,
    n_epochs: int,
    lr: float,
    reg: float,
    verbose: int,
    update_user_params: bool = True,
    update_item_params: bool = True,
    plan_options: dict = None,
    return_order: bool = False,
) -> Tuple[np.ndarray, np.ndarray, list]:
    """Bias SGD (baseline_model.py:215-280); biases updated in place and returned with train_rmse."""
    torch = engine._torch()
    u, i, r = _split_X(X)
    n = len(u)
    du, di, dr = _upload_ratings(u, i, r)
    bu, bi = _mirror.vec(user_biases), _mirror.vec(item_biases)
    plan = engine.Plan(du, di, dr, len(user_biases), len(item_biases), n_factors=0, **(plan_options or {}))
    order = plan.order().cpu().numpy() if return_order else None
    sse = torch.zeros((max(n_epochs, 1),), dtype=torch.float64, device=engine.device())
    for epoch in range(n_epochs):
        engine.bias_sgd_epoch(plan, bu, bi, global_mean, lr, reg, update_user_params, update_item_params)
        engine.bias_sse(du, di, dr, bu, bi, global_mean, sse[epoch:epoch + 1])
        if verbose == 1:
            rmse = math.sqrt(float(sse[epoch].item()) / n) if n else float("nan")
            print("Epoch ", epoch + 1, "/", n_epochs, " -  train_rmse:", rmse)
    train_rmse = []
    if n_epochs > 0:
        train_rmse = [math.sqrt(v / n) if n else float("nan") for v in sse[:n_epochs].cpu().numpy().tolist()]
    plan.close()
    if update_user_params:
        user_biases[...] = engine.download(bu)
        _mirror.set_vec(user_biases, bu)  # (re-registers the mirror with the new content fingerprint)
    if update_item_params:
        item_biases[...] = engine.download(bi)
        _mirror.set_vec(item_biases, bi)
    out = (user_biases, item_biases, train_rmse)
    return out + (order,) if return_order else out


def _als(
    X,
    global_mean: float,
    user_biases: np.ndarray,
    item_biases: np.ndarray,
    n_epochs: int,
    reg: float,
    verbose: int,
) -> Tuple[np.ndarray, np.ndarray, list]:
    """Alternating least squares on the biases (baseline_model.py:283-362)."""
    torch = engine._torch()
    u, i, r = _split_X(X)
    n = len(u)
    du, di, dr = _upload_ratings(u, i, r)
    bu, bi = _mirror.vec(user_biases), _mirror.vec(item_biases)
    csr = engine.Csr(du, di, dr, len(user_biases), len(item_biases))
    sse = torch.zeros((max(n_epochs, 1),), dtype=torch.float64, device=engine.device())
    for epoch in range(n_epochs):
        engine.bias_als_epoch(csr, bu, bi, global_mean, reg)
        engine.bias_sse(du, di, dr, bu, bi, global_mean, sse[epoch:epoch + 1])
        if verbose == 1:
            rmse = math.sqrt(float(sse[epoch].item()) / n) if n else float("nan")
            print("Epoch ", epoch + 1, "/", n_epochs, " -  train_rmse:", rmse)
    train_rmse = []
    if n_epochs > 0:
        train_rmse = [math.sqrt(v / n) if n else float("nan") for v in sse[:n_epochs].cpu().numpy().tolist()]
    csr.close()
    # the reference returns FRESH arrays from _als (:329, :340); callers rebind the attributes
    _mirror.invalidate(user_biases)
    _mirror.invalidate(item_biases)
    user_biases = engine.download(bu)
    item_biases = engine.download(bi)
    _mirror.set_vec(user_biases, bu)
    _mirror.set_vec(item_biases, bi)
    return user_biases, item_biases, train_rmse


def _predict(
    X,
    global_mean: float,
    min_rating: int,
    max_rating: int,
    user_biases: np.ndarray,
    item_biases: np.ndarray,
    bound_ratings: bool,
) -> Tuple[list, list]:
    """baseline_model.py:365-417; id -1 = unknown."""
    u, i, _ = _split_X(X, with_rating=False)
    if len(u) == 0:
        return [], []
    du, di, _ = _upload_ratings(u, i)
    pred, poss = engine.bias_predict(du, di, _mirror.vec(user_biases), _mirror.vec(item_biases), global_mean,
                                     min_rating, max_rating, bound_ratings)
    return pred.cpu().numpy().astype(np.float64).tolist(), poss.cpu().numpy().astype(bool).tolist()
