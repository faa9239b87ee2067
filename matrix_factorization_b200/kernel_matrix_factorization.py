"""
KernelMF -- drop-in for matrix_factorization/kernel_matrix_factorization.py of the reference.

Same constructor, defaults, quirks and fitted attributes; the numeric
loops (`_sgd`, `_calculate_rmse`, `_predict`) keep the reference's signatures but run the
sm_100a kernels of libmfk_b200.so through the C ABI.  The per-epoch `np.random.shuffle` of the
reference (:371, numba's private RNG -- not reproducible) is replaced by the plan's stratified
conflict-free order, which the parity tests replay through the fp64 oracle.
"""
from __future__ import annotations

import math
from typing import Tuple, Union

import numpy as np
import pandas as pd

from . import _mirror, engine
from .recommender_base import RecommenderBase

KERNELS = ("linear", "sigmoid", "rbf")


def _split_X(X, with_rating=True):
    """Accept the reference's N x 3 (or N x 2) float matrix or a tuple of 1-D arrays."""
    if isinstance(X, (tuple, list)):
        u, i = np.asarray(X[0]), np.asarray(X[1])
        r = np.asarray(X[2]) if with_rating else None
    else:
        X = np.asarray(X)
        u, i = X[:, 0], X[:, 1]
        r = X[:, 2] if with_rating else None
    return u, i, r


def _upload_ratings(u, i, r=None):
    torch = engine._torch()
    du = engine.upload_vec(np.asarray(u).astype(np.int32, copy=False), torch.int32)
    di = engine.upload_vec(np.asarray(i).astype(np.int32, copy=False), torch.int32)
    dr = engine.upload_vec(np.asarray(r).astype(np.float32, copy=False), torch.float32) if r is not None else None
    return du, di, dr


class KernelMF(RecommenderBase):
    """
    Kernel Matrix Factorization: thin matrices P (users) and Q (items) plus bias vectors, fitted by
    SGD on the squared error of r_ui ~ K(p_u, q_i) (reference docstring:
    kernel_matrix_factorization.py:19-50).

    Arguments:
        n_factors {int} -- number of latent factors (default: 100)
        n_epochs {int} -- epochs (default: 100)
        kernel {str} -- 'linear', 'sigmoid' or 'rbf' (default: 'linear')
        gamma {str or float} -- rbf coefficient; 'auto' = 1/n_factors, resolved at construction
        reg {float} -- L2 regularisation (default: 1 -- the reference's actual default)
        lr {float} -- learning rate (default: 0.01)
        init_mean, init_sd {float} -- normal initialisation of P and Q (default: 0, 0.1)
        min_rating, max_rating -- rating range (default: 0, 5)
        verbose {int} -- 1 prints one line per epoch (default: 1)

    Attributes: n_users, n_items, global_mean, user_biases, item_biases, user_features,
        item_features (float64 numpy, like the reference), user_id_map, item_id_map, train_rmse,
        predictions_possible (after predict).
    """

    def __init__(
        self,
        n_factors: int = 100,
        n_epochs: int = 100,
        kernel: str = "linear",
        gamma: Union[str, float] = "auto",
        reg: float = 1,
        lr: float = 0.01,
        init_mean: float = 0,
        init_sd: float = 0.1,
        min_rating: int = 0,
        max_rating: int = 5,
        verbose: int = 1,
    ):
        if kernel not in KERNELS:
            raise ValueError("Kernel must be one of linear, sigmoid, or rbf")

        super().__init__(min_rating=min_rating, max_rating=max_rating, verbose=verbose)

        self.n_factors = n_factors
        self.n_epochs = n_epochs
        self.kernel = kernel
        self.gamma = 1 / n_factors if gamma == "auto" else gamma  # resolved here, as in the reference (:74)
        self.reg = reg
        self.lr = lr
        self.init_mean = init_mean
        self.init_sd = init_sd
        return

    def fit(self, X: pd.DataFrame, y: pd.Series):
        """Decompose the rating matrix (kernel_matrix_factorization.py:81-128)."""
        data = self._preprocess_arrays(X=X, y=y, type="fit")
        self.global_mean = pd.Series(data["r"]).mean()  # X["rating"].mean() on the shuffled rows (:90)

        self.user_biases = np.zeros(self.n_users)
        self.item_biases = np.zeros(self.n_items)
        # same draws, same order as the reference (:97-102): users first, then items, global RNG
        self.user_features = np.random.normal(self.init_mean, self.init_sd, (self.n_users, self.n_factors))
        self.item_features = np.random.normal(self.init_mean, self.init_sd, (self.n_items, self.n_factors))

        (
            self.user_features,
            self.item_features,
            self.user_biases,
            self.item_biases,
            self.train_rmse,
        ) = _sgd(
            X=(data["u"], data["i"], data["r"]),
            global_mean=self.global_mean,
            user_biases=self.user_biases,
            item_biases=self.item_biases,
            user_features=self.user_features,
            item_features=self.item_features,
            n_epochs=self.n_epochs,
            kernel=self.kernel,
            gamma=self.gamma,
            lr=self.lr,
            reg=self.reg,
            min_rating=self.min_rating,
            max_rating=self.max_rating,
            verbose=self.verbose,
        )
        return self

    def _predict_internal(self, u, i, bound_ratings):
        return _predict(
            X=(u, i),
            global_mean=self.global_mean,
            user_biases=self.user_biases,
            item_biases=self.item_biases,
            user_features=self.user_features,
            item_features=self.item_features,
            min_rating=self.min_rating,
            max_rating=self.max_rating,
            kernel=self.kernel,
            gamma=self.gamma,
            bound_ratings=bound_ratings,
        )

    def predict(self, X: pd.DataFrame, bound_ratings: bool = True) -> list:
        """Predicted ratings in the order of X (kernel_matrix_factorization.py:130-163)."""
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
        """
        Fit the user parameters of new / re-passed users with the item side frozen
        (kernel_matrix_factorization.py:165-237).  Quirks kept: re-passed users are
        re-initialised, `lr`/`n_epochs` are the ARGUMENTS, reg is self.reg, n_users is not bumped.
        """
        data = self._preprocess_arrays(X=X, y=y, type="update")
        known_users, new_users = data["known_users"], data["new_users"]
        n_new_users = len(new_users)

        for user in known_users:  # :190-199 one RNG call per known user, in list order
            user_index = self.user_id_map[user]
            self.user_biases[user_index] = 0
            self.user_features[user_index, :] = np.random.normal(self.init_mean, self.init_sd, (1, self.n_factors))
        if known_users:
            _mirror.invalidate(self.user_features)
            _mirror.invalidate(self.user_biases)

        self.user_biases = np.append(self.user_biases, np.zeros(n_new_users))  # :202
        new_user_features = np.random.normal(self.init_mean, self.init_sd, (n_new_users, self.n_factors))
        self.user_features = np.concatenate((self.user_features, new_user_features), axis=0)  # :205-210

        (
            self.user_features,
            self.item_features,
            self.user_biases,
            self.item_biases,
            self.train_rmse,
        ) = _sgd(
            X=(data["u"], data["i"], data["r"]),
            global_mean=self.global_mean,
            user_biases=self.user_biases,
            item_biases=self.item_biases,
            user_features=self.user_features,
            item_features=self.item_features,
            n_epochs=n_epochs,
            kernel=self.kernel,
            gamma=self.gamma,
            lr=lr,
            reg=self.reg,
            min_rating=self.min_rating,
            max_rating=self.max_rating,
            verbose=verbose,
            update_item_params=False,
        )
        return

    def _score_topk(self, user_internal, k, mask_ptr, mask_items, bound_ratings):
        torch = engine._torch()
        P, Q = _mirror.rows(self.user_features), _mirror.rows(self.item_features)
        bu, bi = _mirror.vec(self.user_biases), _mirror.vec(self.item_biases)
        users = engine.upload_vec(np.asarray(user_internal, dtype=np.int32), torch.int32)
        mp = engine.upload_vec(np.asarray(mask_ptr, dtype=np.int64), torch.int64) if mask_ptr is not None else None
        mi = None
        if mask_ptr is not None:
            mi = engine.upload_vec(np.asarray(mask_items, dtype=np.int32), torch.int32)
            if mi.numel() == 0:
                mi = torch.zeros((1,), dtype=torch.int32, device=engine.device())
        scores, items = engine.score_topk(self.kernel, users, P, Q, bu, bi, self.item_features.shape[0],
                                          self.n_factors, self.global_mean, self.gamma, self.min_rating,
                                          self.max_rating, k, bound_ratings, mp, mi)
        return scores.cpu().numpy().astype(np.float64), items.cpu().numpy()

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_raw_items_cache", None)  # device mirrors live outside the instance
        return state


# ---------------------------------------------------------------------------------------
# Operator boundary: same names / arguments / results as the reference's njit functions.
# ---------------------------------------------------------------------------------------
def _calculate_rmse(
    X,
    global_mean: float,
    user_biases: np.ndarray,
    item_biases: np.ndarray,
    user_features: np.ndarray,
    item_features: np.ndarray,
    min_rating: float,
    max_rating: float,
    kernel: str,
    gamma: float,
):
    """RMSE of the unclipped predictions over X (kernel_matrix_factorization.py:240-317)."""
    torch = engine._torch()
    u, i, r = _split_X(X)
    n = len(u)
    if n == 0:
        return float("nan")
    du, di, dr = _upload_ratings(u, i, r)
    P, Q = _mirror.rows(user_features), _mirror.rows(item_features)
    bu, bi = _mirror.vec(user_biases), _mirror.vec(item_biases)
    out = torch.zeros((1,), dtype=torch.float64, device=engine.device())
    engine.kmf_sse(kernel, du, di, dr, P, Q, bu, bi, user_features.shape[1], global_mean, gamma, min_rating,
                   max_rating, out)
    return math.sqrt(float(out.item()) / n)


def _sgd(
    X,
    global_mean: float,
    user_biases: np.ndarray,
    item_biases: np.ndarray,
    user_features: np.ndarray,
    item_features: np.ndarray,
    n_epochs: int,
    kernel: str,
    gamma: float,
    lr: float,
    reg: float,
    min_rating: float,
    max_rating: float,
    verbose: int,
    update_user_params: bool = True,
    update_item_params: bool = True,
    plan_options: dict = None,
    return_order: bool = False,
) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, list]:
    """
    SGD over the ratings X (kernel_matrix_factorization.py:320-445).  Parameters are updated IN
    PLACE and returned, with the list of per-epoch training RMSEs.  X is the reference's N x 3
    float matrix (user, item, rating) or a tuple of three 1-D arrays.

    Extras (not in the reference): `plan_options` (n_workers / warps_per_cta overrides) and
    `return_order=True`, which appends the plan's sequential-replay order to the result.
    """
    torch = engine._torch()
    if kernel not in KERNELS:
        raise ValueError("Kernel must be one of linear, sigmoid, or rbf")
    u, i, r = _split_X(X)
    n = len(u)
    F = user_features.shape[1]
    du, di, dr = _upload_ratings(u, i, r)
    P, Q = _mirror.rows(user_features), _mirror.rows(item_features)
    bu, bi = _mirror.vec(user_biases), _mirror.vec(item_biases)
    opts = dict(plan_options or {})
    if "hot_min_degree" not in opts:
        # the exact mini-batch path for the most-rated items exists for the linear kernel with item updates
        opts["hot_min_degree"] = 0 if (kernel == "linear" and update_item_params and "n_workers" not in opts) \
            else engine.Plan.NO_HOT_SPLIT
    plan = engine.Plan(du, di, dr, user_features.shape[0], item_features.shape[0], n_factors=F, **opts)
    order = plan.order().cpu().numpy() if return_order else None
    sse = torch.zeros((max(n_epochs, 1),), dtype=torch.float64, device=engine.device())
    train_rmse = []
    for epoch in range(n_epochs):
        engine.kmf_sgd_epoch(plan, kernel, P, Q, bu, bi, F, global_mean, lr, reg, gamma, min_rating, max_rating,
                             update_user_params, update_item_params)
        engine.kmf_sse_plan(plan, kernel, P, Q, bu, bi, F, global_mean, gamma, min_rating, max_rating,
                            sse[epoch:epoch + 1])
        if verbose == 1:
            rmse = math.sqrt(float(sse[epoch].item()) / n) if n else float("nan")
            print("Epoch ", epoch + 1, "/", n_epochs, " -  train_rmse:", rmse)
    if n_epochs > 0:
        host = sse[:n_epochs].cpu().numpy()
        train_rmse = [math.sqrt(v / n) if n else float("nan") for v in host.tolist()]
    plan.close()
    # in-place write-back (the reference mutates its inputs and returns them, :445)
    if update_user_params:
        user_features[...] = engine.download(P, cols=F)
        user_biases[...] = engine.download(bu)
        _mirror.set_rows(user_features, P)  # (re-registers the mirror with the new content fingerprint)
        _mirror.set_vec(user_biases, bu)
    if update_item_params:
        item_features[...] = engine.download(Q, cols=F)
        item_biases[...] = engine.download(bi)
        _mirror.set_rows(item_features, Q)
        _mirror.set_vec(item_biases, bi)
    out = (user_features, item_features, user_biases, item_biases, train_rmse)
    return out + (order,) if return_order else out


def _predict(
    X,
    global_mean: float,
    user_biases: np.ndarray,
    item_biases: np.ndarray,
    user_features: np.ndarray,
    item_features: np.ndarray,
    min_rating: int,
    max_rating: int,
    kernel: str,
    gamma: float,
    bound_ratings: bool,
) -> Tuple[list, list]:
    """
    Predicted rating per (user, item) row; id -1 = unknown (kernel_matrix_factorization.py:448-541).
    Returns (list[float], list[bool]) like the reference.
    """
    u, i, _ = _split_X(X, with_rating=False)
    if len(u) == 0:
        return [], []
    du, di, _ = _upload_ratings(u, i)
    P, Q = _mirror.rows(user_features), _mirror.rows(item_features)
    bu, bi = _mirror.vec(user_biases), _mirror.vec(item_biases)
    pred, poss = engine.kmf_predict(kernel, du, di, P, Q, bu, bi, user_features.shape[1], global_mean, gamma,
                                    min_rating, max_rating, bound_ratings)
    return pred.cpu().numpy().astype(np.float64).tolist(), poss.cpu().numpy().astype(bool).tolist()
