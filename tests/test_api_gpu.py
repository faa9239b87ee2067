"""GPU tests of the callers either side of the hot path: batched recommend_all + the evaluate loop (f2),
GridSearchCV with worker processes (f1)."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def _fit(n_users=1400, n_items=300, n=60_000, kernel="linear", seed=21):
    import matrix_factorization_b200 as mfb
    from matrix_factorization_b200.data import synth_ratings

    df = synth_ratings(n_users, n_items, n, seed=seed, min_per_user=8)
    np.random.seed(seed)
    kw = dict(lr=0.01) if kernel != "rbf" else dict(lr=0.3, gamma=0.05)
    m = mfb.KernelMF(n_factors=32, n_epochs=8, kernel=kernel, reg=0.01, verbose=0, **kw)
    m.fit(df[["user_id", "item_id"]], df["rating"])
    return m, df


@pytest.mark.parametrize("kernel", ["linear", "rbf"])
def test_recommend_all_equals_per_user_recommend_on_1000_users(kernel):
    """f2: one batched scoring pass with the training items masked gives, for every user, the list that the reference's
    per-user recommend(user, items_known=...) call gives (recommender_base.py:214-271); ties within 1e-5 may swap."""
    m, df = _fit(kernel=kernel)
    users = df["user_id"].unique()[:1100].tolist()
    k = 12
    known = df[["user_id", "item_id"]]
    ra = m.recommend_all(users=users, amount=k, items_known=known)
    known_by_user = known.groupby("user_id")["item_id"].apply(list).to_dict()
    sizes = ra.groupby("user_id").size()
    assert ra["user_id"].nunique() == len(users)
    assert all(sizes[u] == min(k, m.n_items - len(known_by_user[u])) for u in users)  # (a user may have rated almost everything)
    by_user = {u: g.sort_values("rank") for u, g in ra.groupby("user_id")}
    rng = np.random.default_rng(0)
    for u in rng.choice(users, 120, replace=False).tolist() + users[:5]:
        one = m.recommend(user=u, amount=k, items_known=known_by_user[u])
        a, b = by_user[u], one
        assert not set(a["item_id"]) & set(known_by_user[u])
        sa, sb = a["rating_pred"].to_numpy(), b["rating_pred"].to_numpy()
        np.testing.assert_allclose(sa, sb, atol=2e-5)
        ia, ib = a["item_id"].to_numpy(), b["item_id"].to_numpy()
        diff = ia != ib
        if diff.any():  # only ties (within 1e-5) may be ordered differently
            assert np.all(np.abs(sa[diff] - sb[diff]) < 1e-5) and set(ia[diff]) == set(ib[diff])


def test_evaluate_topk_on_recommend_all_matches_per_user_loop():
    """f2: precision / recall / NDCG@k from the batched path equal the reference-style loop over recommend()
    (project_template/pipeline/evaluate.py:61-111)."""
    from matrix_factorization_b200 import evaluation as ev

    m, df = _fit(n_users=500, n_items=200, n=20_000, seed=5)
    k, n_test, thr, seed = 10, 3, 4.0, 7
    res = ev.evaluate_topk(df, m, k, thr, n_test, seed)
    train, test = ev.holdout_split(df, n_test, thr, seed)
    rel = test.groupby("user_id")["item_id"].apply(set).to_dict()
    kn = train.groupby("user_id")["item_id"].apply(list).to_dict()
    ps, rs, ns = [], [], []
    for u in list(kn)[:150]:
        rec = m.recommend(user=u, amount=k, items_known=kn[u], include_user=False)
        hit = np.array([1 if i in rel[u] else 0 for i in rec["item_id"].tolist()], dtype=np.float64)
        ps.append(hit.mean())
        rs.append(hit.sum() / max(1, len(rel[u])))
        ideal = np.sort(hit)[::-1]
        disc = 1.0 / np.log2(np.arange(2, hit.size + 2))
        ns.append((hit * disc).sum() / (ideal * disc).sum() if hit.sum() > 0 else 0.0)
    sub = ev.topk_metrics(m.recommend_all(users=list(kn)[:150], amount=k, items_known=train[["user_id", "item_id"]]),
                          test[test["user_id"].isin(list(kn)[:150])], k)
    assert res.n_users == len(kn) and 0.0 <= res.precision <= 1.0
    assert abs(sub.precision - np.mean(ps)) < 1e-9 and abs(sub.recall - np.mean(rs)) < 1e-9 and abs(sub.ndcg - np.mean(ns)) < 1e-9


def test_grid_search_cv_with_worker_processes():
    """f1: GridSearchCV(n_jobs=2) clones and pickles the estimator into worker processes, each of which creates its own
    CUDA context lazily (examples/recommender-system.ipynb:3366-3376)."""
    from sklearn.model_selection import GridSearchCV

    import matrix_factorization_b200 as mfb
    from matrix_factorization_b200.data import synth_ratings

    df = synth_ratings(300, 120, 9000, seed=3, min_per_user=6)
    X, y = df[["user_id", "item_id"]], df["rating"]
    grid = GridSearchCV(mfb.KernelMF(n_epochs=6, lr=0.01, verbose=0), {"n_factors": [8, 16], "reg": [0.005, 0.05]},
                        scoring="neg_root_mean_squared_error", cv=3, n_jobs=2, refit=True)
    grid.fit(X, y)
    assert set(grid.best_params_) == {"n_factors", "reg"}
    assert np.all(np.isfinite(grid.cv_results_["mean_test_score"])) and grid.best_score_ < 0
    pred = grid.best_estimator_.predict(X.head(50))
    assert len(pred) == 50 and np.all(np.isfinite(pred))


def test_small_batch_predictor_matches_predict_and_is_fast():
    """f4: one user x 500 candidate items through the captured-graph predictor equals model.predict(bound_ratings=False)
    (project_template/app/api.py:43-52) and takes well under the reference's ~0.3 ms CPU path."""
    import time

    import matrix_factorization_b200 as mfb
    from matrix_factorization_b200.data import synth_ratings

    df = synth_ratings(2000, 900, 80_000, seed=8, min_per_user=8)
    np.random.seed(8)
    for model in (mfb.KernelMF(n_factors=64, n_epochs=4, lr=0.01, reg=0.01, verbose=0), mfb.BaselineModel(method="als", n_epochs=4, verbose=0)):
        model.fit(df[["user_id", "item_id"]], df["rating"])
        user = df["user_id"].iloc[0]
        items = df["item_id"].unique()[:500].tolist() + [10 ** 9]  # one unknown item
        frame = pd.DataFrame({"user_id": [user] * len(items), "item_id": items})
        want = model.predict(frame, bound_ratings=False)
        want_possible = list(model.predictions_possible)
        pr = model.predictor(capacity=512, bound_ratings=False)
        got, possible = pr.predict_pairs(user, items)
        np.testing.assert_allclose(got, want, atol=1e-6)
        assert possible.tolist() == want_possible and not possible[-1]
        assert pr.predict(frame) == pytest.approx(want, abs=1e-6) and model.predictions_possible == want_possible
        # unknown user: bias-only prediction, flagged
        g2, p2 = pr.predict_pairs(-12345, items[:10])
        assert not p2.any()
        ts = []
        for _ in range(300):
            t0 = time.perf_counter()
            pr.predict_pairs(user, items)
            ts.append(time.perf_counter() - t0)
        med = float(np.median(ts[50:]))
        assert med < 0.3e-3, f"median request latency {med * 1e3:.3f} ms"
    # a refit is picked up after refresh()
    model.fit(df[["user_id", "item_id"]], df["rating"])
    pr.refresh()
    np.testing.assert_allclose(pr.predict_pairs(user, items)[0], model.predict(frame, bound_ratings=False), atol=1e-6)


def test_gpu_preprocess_is_bit_exact_with_the_host_path(monkeypatch, golden_dir):
    """f3: first-appearance id assignment + duplicate check on the GPU (mfk_first_appearance / mfk_has_duplicate_pairs)
    against the host path and the reference-generated fixture (recommender_base.py:125-164), same RNG consumption."""
    import os

    import matrix_factorization_b200 as mfb

    rng = np.random.default_rng(4)
    n, U, I = 300_000, 5000, 1200
    keys = rng.choice(U * I, n, replace=False)
    raw_u = rng.permutation(10 * U)[:U][keys // I].astype(np.int64) - 7   # arbitrary (also negative) integer ids
    raw_i = rng.permutation(10 * I)[:I][keys % I].astype(np.int64)
    X = pd.DataFrame({"user_id": raw_u, "item_id": raw_i})
    y = pd.Series(rng.integers(1, 6, n).astype(np.float64))
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MFB_GPU_PREPROCESS", mode)
        m = mfb.BaselineModel(verbose=0)
        np.random.seed(99)
        out = m._preprocess_arrays(X.copy(), y, "fit")
        res[mode] = (out, list(m.user_id_map.keys()), list(m.item_id_map.keys()), np.random.random(), m.n_users, m.n_items)
    a, b = res["0"], res["1"]
    assert np.array_equal(a[0]["u"], b[0]["u"]) and np.array_equal(a[0]["i"], b[0]["i"]) and np.array_equal(a[0]["r"], b[0]["r"])
    assert a[1] == b[1] and a[2] == b[2] and a[3] == b[3] and a[4:] == b[4:]
    # the reference's own output (fixture)
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    monkeypatch.setenv("MFB_GPU_PREPROCESS", "1")
    m = mfb.BaselineModel(method="als", verbose=0)
    train = pd.DataFrame({"user_id": g["train_user"], "item_id": g["train_item"]})
    np.random.seed(int(g["fit_seed"]))
    Xf = m._preprocess_data(train, pd.Series(g["train_rating"]), type="fit")
    assert np.array_equal(Xf["user_id"].to_numpy(), g["fit_u"]) and np.array_equal(Xf["item_id"].to_numpy(), g["fit_i"])
    assert np.array_equal(Xf["rating"].to_numpy(), g["fit_r"])
    assert list(m.user_id_map.keys()) == g["umap_keys"].tolist() and list(m.item_id_map.keys()) == g["imap_keys"].tolist()
    # duplicates raise before the RNG is consumed
    Xd = pd.concat([X.head(1000), X.head(1)], ignore_index=True)
    np.random.seed(5)
    with pytest.raises(ValueError, match="Duplicate user-item ratings in matrix"):
        m._preprocess_arrays(Xd, pd.Series(np.ones(1001)), "fit")
    nxt = np.random.random()
    np.random.seed(5)
    assert nxt == np.random.random()
    # and a whole fit goes through it
    monkeypatch.setenv("MFB_GPU_PREPROCESS", "1")
    np.random.seed(1)
    k1 = mfb.KernelMF(n_factors=8, n_epochs=2, verbose=0).fit(X.head(50_000), y.head(50_000))
    monkeypatch.setenv("MFB_GPU_PREPROCESS", "0")
    np.random.seed(1)
    k0 = mfb.KernelMF(n_factors=8, n_epochs=2, verbose=0).fit(X.head(50_000), y.head(50_000))
    assert k0.train_rmse == k1.train_rmse and np.array_equal(k0.user_features, k1.user_features)
