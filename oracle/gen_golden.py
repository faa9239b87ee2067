"""
oracle/gen_golden.py -- generates tests/golden/*.npz / *.json by running the REFERENCE ITSELF
(imported from /root/reference; never copied) in the build container.

    python oracle/gen_golden.py            # rewrites tests/golden/

The reference cannot travel to the GPU box, so the vectors are committed as small fixtures
together with this script.  They pin (a) the C/numpy oracle (tests/test_oracle_golden.py,
CPU) and (b) the CUDA path (tests/test_*_gpu.py).

Known reference breakages under pandas 3 and how each fixture avoids them:
  * predict() with unknown ids raises TypeError      -> `_predict` njit is called with -1 ids
  * update_users() raises numba TypingError          -> `_preprocess_data('update')` + the array
    growth lines + `_sgd(np.array(X, order="F"), update_item_params=False)` are run directly
"""
import json
import os
import sys

import numba as nb
import numpy as np
import pandas as pd

REF = os.environ.get("MF_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from matrix_factorization import BaselineModel, KernelMF  # noqa: E402  (the reference)
from matrix_factorization import baseline_model as ref_bm  # noqa: E402
from matrix_factorization import kernel_matrix_factorization as ref_kmf  # noqa: E402
from matrix_factorization import kernels as ref_k  # noqa: E402

from matrix_factorization_b200.data import synth_ratings, split_rows  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


@nb.njit()
def _seed_numba(s):
    np.random.seed(s)


@nb.njit()
def _replay(kernel, u, i, r, order, mu, bu, bi, P, Q, lr, reg, gamma, a, c, uu, ui):
    for k in range(order.shape[0]):
        j = order[k]
        if kernel == 0:
            ref_k.kernel_linear_sgd_update(u[j], i[j], r[j], mu, bu, bi, P, Q, lr, reg, uu, ui)
        elif kernel == 1:
            ref_k.kernel_sigmoid_sgd_update(u[j], i[j], r[j], mu, bu, bi, P, Q, lr, reg, a, c, uu, ui)
        else:
            ref_k.kernel_rbf_sgd_update(u[j], i[j], r[j], P, Q, lr, reg, gamma, a, c, uu, ui)


def kat_updates():
    """SURVEY.md 9.2 known-answer vectors, recomputed from the reference functions."""
    out = {}
    base = dict(p=[0.1, 0.2], q=[0.3, -0.1], bu=0.1, bi=-0.2, mu=3.0, r=4.0, lr=0.01, reg=0.005,
                a=0.0, c=5.0, gamma=0.5)
    for name, kid, uu, ui in [("linear_tt", 0, True, True), ("linear_tf", 0, True, False),
                              ("linear_ft", 0, False, True), ("sigmoid_tt", 1, True, True),
                              ("sigmoid_tf", 1, True, False), ("rbf_tt", 2, True, True),
                              ("rbf_tf", 2, True, False)]:
        P = np.array([base["p"]], dtype=np.float64)
        Q = np.array([base["q"]], dtype=np.float64)
        bu = np.array([base["bu"]])
        bi = np.array([base["bi"]])
        _replay(kid, np.array([0]), np.array([0]), np.array([base["r"]]), np.array([0]), base["mu"],
                bu, bi, P, Q, base["lr"], base["reg"], base["gamma"], base["a"], base["c"], uu, ui)
        out[name] = dict(kernel=kid, upd_user=uu, upd_item=ui, bu=bu[0], bi=bi[0], p=P[0].tolist(),
                         q=Q[0].tolist())
    p, q = np.array(base["p"]), np.array(base["q"])
    out["pred_linear"] = ref_k.kernel_linear(base["mu"], base["bu"], base["bi"], p, q)
    out["pred_sigmoid"] = ref_k.kernel_sigmoid(base["mu"], base["bu"], base["bi"], p, q, 0.0, 5.0)
    out["pred_rbf"] = ref_k.kernel_rbf(p, q, 0.5, 0.0, 5.0)
    out["inputs"] = base
    # ALS KAT
    X = np.array([[0, 0, 5], [0, 1, 3], [1, 0, 4], [1, 1, 1], [2, 1, 2]], dtype=np.float64)
    bu, bi, rm = ref_bm._als(X, 3.0, np.zeros(3), np.zeros(2), 2, 0.5, 0)
    out["als"] = dict(X=X.tolist(), mu=3.0, reg=0.5, n_epochs=2, bu=bu.tolist(), bi=bi.tolist(),
                      train_rmse=list(rm))
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(out, f, indent=1)


def small_problem(seed, U=37, I=23, N=400, F=12):
    rng = np.random.default_rng(seed)
    keys = rng.choice(U * I, N, replace=False)
    u, i = (keys // I).astype(np.int64), (keys % I).astype(np.int64)
    r = rng.integers(1, 6, N).astype(np.float64)
    P = rng.normal(0, 0.1, (U, F))
    Q = rng.normal(0, 0.1, (I, F))
    bu = rng.normal(0, 0.05, U)
    bi = rng.normal(0, 0.05, I)
    return u, i, r, P, Q, bu, bi


def replay_vectors():
    """One pass of the reference update functions over a random small problem in a fixed order,
    plus the reference RMSE and the reference _predict (with -1 ids) on the result."""
    for kid, kname, lr, gamma in [(0, "linear", 0.01, 0.01), (1, "sigmoid", 0.05, 0.01), (2, "rbf", 0.3, 0.05)]:
        for uu, ui in [(True, True), (True, False)]:
            u, i, r, P, Q, bu, bi = small_problem(100 + kid)
            rng = np.random.default_rng(5 + kid)
            order = rng.permutation(len(u)).astype(np.int64)
            P0, Q0, bu0, bi0 = P.copy(), Q.copy(), bu.copy(), bi.copy()
            mu, reg, a, c = float(r.mean()), 0.02, 0.0, 5.0
            _replay(kid, u, i, r, order, mu, bu, bi, P, Q, lr, reg, gamma, a, c, uu, ui)
            X = np.stack([u, i, r], axis=1).astype(np.float64)
            rmse = ref_kmf._calculate_rmse(X, mu, bu, bi, P, Q, a, a + c, kname, gamma)
            pu = np.concatenate([u[:50], [-1, -1, 3]]).astype(np.float64)
            pi = np.concatenate([i[:50], [-1, 2, -1]]).astype(np.float64)
            Xp = np.stack([pu, pi], axis=1)
            pb, possb = ref_kmf._predict(Xp, mu, bu, bi, P, Q, 0.0, 5.0, kname, gamma, True)
            pn, _ = ref_kmf._predict(Xp, mu, bu, bi, P, Q, 0.0, 5.0, kname, gamma, False)
            np.savez(os.path.join(OUT, f"replay_{kname}_{int(uu)}{int(ui)}.npz"), u=u, i=i, r=r,
                     order=order, P0=P0, Q0=Q0, bu0=bu0, bi0=bi0, P=P, Q=Q, bu=bu, bi=bi, mu=mu,
                     lr=lr, reg=reg, gamma=gamma, rmse=rmse, pred_u=pu, pred_i=pi,
                     pred_bound=np.array(list(pb)), pred_unbound=np.array(list(pn)),
                     possible=np.array(list(possb)))


def sgd_njit_vectors():
    """The reference's own njit _sgd for 3 single epochs; the 4th column of X carries the row id so
    the order its private shuffle produced can be replayed by the oracle."""
    for kid, kname, lr, gamma in [(0, "linear", 0.01, 0.01), (1, "sigmoid", 0.05, 0.01), (2, "rbf", 0.3, 0.05)]:
        u, i, r, P, Q, bu, bi = small_problem(200 + kid)
        bu[:] = 0
        bi[:] = 0
        P0, Q0 = P.copy(), Q.copy()
        mu, reg = float(r.mean()), 0.02
        X = np.stack([u, i, r, np.arange(len(u))], axis=1).astype(np.float64)
        _seed_numba(11 + kid)
        orders, rmses = [], []
        for _ in range(3):
            P, Q, bu, bi, rm = ref_kmf._sgd(X, mu, bu, bi, P, Q, 1, kname, gamma, lr, reg, 0.0, 5.0, 0)
            orders.append(X[:, 3].astype(np.int64).copy())
            rmses.append(rm[0])
        np.savez(os.path.join(OUT, f"sgd_{kname}.npz"), u=u, i=i, r=r, P0=P0, Q0=Q0, mu=mu, lr=lr,
                 reg=reg, gamma=gamma, orders=np.stack(orders), rmse=np.array(rmses), P=P, Q=Q,
                 bu=bu, bi=bi)


def baseline_vectors():
    u, i, r, _, _, _, _ = small_problem(300, U=60, I=40, N=900)
    mu = float(r.mean())
    X = np.stack([u, i, r, np.arange(len(u))], axis=1).astype(np.float64)
    bu, bi = np.zeros(60), np.zeros(40)
    _seed_numba(21)
    orders, rmses = [], []
    for _ in range(3):
        bu, bi, rm = ref_bm._sgd(X, mu, bu, bi, 1, 0.01, 0.02, 0)
        orders.append(X[:, 3].astype(np.int64).copy())
        rmses.append(rm[0])
    # update_users flavour: item biases frozen
    bu2, bi2 = bu.copy(), bi.copy()
    X2 = np.stack([u, i, r, np.arange(len(u))], axis=1).astype(np.float64)
    bu2, bi2, rm2 = ref_bm._sgd(X2, mu, bu2, bi2, 1, 0.01, 0.02, 0, True, False)
    X3 = np.stack([u, i, r], axis=1).astype(np.float64)
    abu, abi, arm = ref_bm._als(X3, mu, np.zeros(60), np.zeros(40), 4, 0.5, 0)
    pu = np.concatenate([u[:30], [-1, -1, 3]]).astype(np.float64)
    pi = np.concatenate([i[:30], [-1, 2, -1]]).astype(np.float64)
    pb, poss = ref_bm._predict(np.stack([pu, pi], axis=1), mu, 0.0, 5.0, abu * 8, abi * 8, True)
    pn, _ = ref_bm._predict(np.stack([pu, pi], axis=1), mu, 0.0, 5.0, abu * 8, abi * 8, False)
    np.savez(os.path.join(OUT, "baseline.npz"), u=u, i=i, r=r, mu=mu, lr=0.01, reg=0.02,
             orders=np.stack(orders), rmse=np.array(rmses), bu=bu, bi=bi,
             order_frozen=X2[:, 3].astype(np.int64), bu_frozen=bu2, bi_frozen=bi2, rmse_frozen=rm2[0],
             als_bu=abu, als_bi=abi, als_rmse=np.array(list(arm)), als_reg=0.5,
             pred_u=pu, pred_i=pi, pred_bound=np.array(list(pb)), pred_unbound=np.array(list(pn)),
             possible=np.array(list(poss)))


def preprocess_vectors():
    """RecommenderBase._preprocess_data run directly: fit / update / predict modes + recommend."""
    df = synth_ratings(50, 40, 600, seed=77, min_per_user=3)
    train = df.iloc[:450].reset_index(drop=True)
    upd = df.iloc[450:].reset_index(drop=True).copy()
    # make some update rows refer to unknown items / brand-new users
    upd.loc[upd.index[:15], "item_id"] = upd["item_id"].iloc[:15] + 1000
    upd.loc[upd.index[15:60], "user_id"] = upd["user_id"].iloc[15:60] + 5000
    upd = upd.drop_duplicates(subset=["user_id", "item_id"]).reset_index(drop=True)

    m = BaselineModel(method="als", n_epochs=1, verbose=0)
    np.random.seed(1234)
    Xf = m._preprocess_data(train[["user_id", "item_id"]], train["rating"], type="fit")
    fit_u, fit_i, fit_r = Xf["user_id"].to_numpy(), Xf["item_id"].to_numpy(), Xf["rating"].to_numpy()
    umap_keys = np.array(list(m.user_id_map.keys()))
    imap_keys = np.array(list(m.item_id_map.keys()))
    np.random.seed(4321)
    Xu, known, new = m._preprocess_data(upd[["user_id", "item_id"]], upd["rating"], type="update")
    # predict mode: cast ids to float so NaN fits (pandas 3 breakage)
    pq = pd.DataFrame({"user_id": [train.user_id[0], 999999, train.user_id[5], 999998],
                       "item_id": [train.item_id[3], train.item_id[4], 888888, 888887]}).astype(np.float64)
    Xp = m._preprocess_data(pq, type="predict")
    np.savez(os.path.join(OUT, "preprocess.npz"),
             train_user=train.user_id.to_numpy(), train_item=train.item_id.to_numpy(),
             train_rating=train.rating.to_numpy(), fit_seed=1234, fit_u=fit_u.astype(np.int64),
             fit_i=fit_i.astype(np.int64), fit_r=fit_r, umap_keys=umap_keys, imap_keys=imap_keys,
             upd_user=upd.user_id.to_numpy(), upd_item=upd.item_id.to_numpy(),
             upd_rating=upd.rating.to_numpy(), upd_seed=4321, upd_u=Xu["user_id"].to_numpy().astype(np.int64),
             upd_i=Xu["item_id"].to_numpy().astype(np.int64), upd_r=Xu["rating"].to_numpy(),
             upd_known=np.array(known), upd_new=np.array(new),
             umap_keys_after=np.array(list(m.user_id_map.keys())),
             pq_user=pq.user_id.to_numpy(), pq_item=pq.item_id.to_numpy(),
             pq_u=Xp["user_id"].to_numpy().astype(np.int64), pq_i=Xp["item_id"].to_numpy().astype(np.int64))


def fit_vectors():
    """Full reference fits on planted data: final train RMSE, test RMSE, recommend lists.
    Statistical goldens (the per-epoch shuffle differs), tolerance 1e-3 by north_star."""
    res = {}
    U, I, N, seed = 943, 1682, 100_000, 1001
    df = synth_ratings(U, I, N, seed=seed, min_per_user=20)
    train, test = split_rows(df, 0.1, seed=3)
    cfgs = {
        "linear": dict(kernel="linear", n_factors=32, n_epochs=20, lr=0.005, reg=0.02),
        "sigmoid": dict(kernel="sigmoid", n_factors=32, n_epochs=20, lr=0.01, reg=0.005),
        "rbf": dict(kernel="rbf", n_factors=32, n_epochs=20, lr=0.5, reg=0.005, gamma=0.01),
    }
    # np.random.seed is FIXED (same row shuffle, id maps and init as a seeded fit of the new
    # implementation); only numba's private per-epoch shuffle stream varies -> the spread below is
    # the reference's own order-to-order noise.
    for name, kw in cfgs.items():
        tr, te = [], []
        for s in range(4):
            np.random.seed(50)
            _seed_numba(60 + s)
            m = KernelMF(verbose=0, **kw).fit(train[["user_id", "item_id"]], train["rating"])
            pred = m.predict(test[["user_id", "item_id"]].astype(np.float64))
            tr.append(m.train_rmse[-1])
            te.append(float(np.sqrt(np.mean((np.array(pred) - test.rating.to_numpy()) ** 2))))
        res[name] = dict(params=kw, train_rmse=tr, test_rmse=te)
    for method, kw in {"sgd": dict(method="sgd", n_epochs=20, reg=0.005, lr=0.01),
                       "als": dict(method="als", n_epochs=20, reg=0.5)}.items():
        tr, te = [], []
        for s in range(4):
            np.random.seed(50)
            _seed_numba(60 + s)
            m = BaselineModel(verbose=0, **kw).fit(train[["user_id", "item_id"]], train["rating"])
            pred = m.predict(test[["user_id", "item_id"]].astype(np.float64))
            tr.append(m.train_rmse[-1])
            te.append(float(np.sqrt(np.mean((np.array(pred) - test.rating.to_numpy()) ** 2))))
        res["baseline_" + method] = dict(params=kw, train_rmse=tr, test_rmse=te)
    res["data"] = dict(n_users=U, n_items=I, n_ratings=N, seed=seed, min_per_user=20,
                       split_seed=3, test_frac=0.1, np_seed=50)
    with open(os.path.join(OUT, "fit_rmse.json"), "w") as f:
        json.dump(res, f, indent=1)

    # recommend(): deterministic given fitted parameters -> store params + reference lists
    np.random.seed(99)
    _seed_numba(99)
    m = KernelMF(n_factors=8, n_epochs=5, lr=0.01, reg=0.02, verbose=0).fit(
        train[["user_id", "item_id"]], train["rating"])
    users = list(m.user_id_map.keys())[:6]
    recs = {}
    for uid in users:
        known = train.loc[train.user_id == uid, "item_id"].tolist() + [987654321]
        rec = m.recommend(user=uid, amount=10, items_known=known, include_user=True, bound_ratings=True)
        recs[str(uid)] = dict(items=rec["item_id"].tolist(), scores=rec["rating_pred"].tolist(),
                              index=rec.index.tolist(), known=known)
    np.savez(os.path.join(OUT, "recommend.npz"), P=m.user_features, Q=m.item_features,
             bu=m.user_biases, bi=m.item_biases, mu=m.global_mean,
             umap_keys=np.array(list(m.user_id_map.keys())), imap_keys=np.array(list(m.item_id_map.keys())),
             recs=json.dumps(recs))


def split_vectors():
    """train_update_test_split (utils.py:8-72) under a fixed numpy seed: the row indices of the six frames."""
    from matrix_factorization import train_update_test_split as ref_split

    df = synth_ratings(60, 40, 1500, seed=5, min_per_user=6)
    np.random.seed(17)
    Xi, yi, Xu, yu, Xt, yt = ref_split(df, frac_new_users=0.25)
    nxt = np.random.random()  # the RNG state after the call is part of the contract (what fit() draws next)
    np.savez(os.path.join(OUT, "split.npz"), seed=17, frac=0.25, idx_initial=Xi.index.to_numpy(), idx_update=Xu.index.to_numpy(),
             idx_test=Xt.index.to_numpy(), y_initial=yi.to_numpy(), y_update=yu.to_numpy(), y_test=yt.to_numpy(), next_draw=nxt)


if __name__ == "__main__":
    kat_updates()
    replay_vectors()
    sgd_njit_vectors()
    baseline_vectors()
    preprocess_vectors()
    fit_vectors()
    split_vectors()
    print("golden fixtures written to", OUT)
