"""GPU parity tests of the BaselineModel path (bias SGD, ALS, CSR/CSC build)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_bias_sgd_matches_reference_replay(golden_dir):
    from matrix_factorization_b200 import baseline_model as bm
    from oracle import oracle as orc

    g = np.load(os.path.join(golden_dir, "baseline.npz"))
    mu, lr, reg = float(g["mu"]), float(g["lr"]), float(g["reg"])
    X = np.stack([g["u"], g["i"], g["r"]], axis=1).astype(np.float64)
    for opts, uu, ui in [(None, True, True), (dict(n_workers=6, warps_per_cta=3), True, True), (None, True, False)]:
        bu, bi = np.full(60, 0.01), np.full(40, -0.02)
        bu0, bi0 = bu.copy(), bi.copy()
        b1, b2, rm, order = bm._sgd(X, mu, bu, bi, 2, lr, reg, 0, uu, ui, plan_options=opts, return_order=True)
        assert b1 is bu and b2 is bi
        buo, bio = bu0, bi0
        for e in range(2):
            buo, bio = orc.bias_replay(g["u"], g["i"], g["r"], order, mu, buo, bio, lr, reg, uu, ui)
            assert abs(rm[e] - orc.bias_rmse(g["u"], g["i"], g["r"], mu, buo, bio)) < 1e-5
        assert np.max(np.abs(bu - buo)) < 1e-5 and np.max(np.abs(bi - bio)) < 1e-5
        if not ui:
            assert np.array_equal(bi, bi0)


def test_als_matches_reference(golden_dir):
    from matrix_factorization_b200 import baseline_model as bm

    g = np.load(os.path.join(golden_dir, "baseline.npz"))
    X = np.stack([g["u"], g["i"], g["r"]], axis=1).astype(np.float64)
    bu, bi, rm = bm._als(X, float(g["mu"]), np.zeros(60), np.zeros(40), 4, float(g["als_reg"]), 0)
    np.testing.assert_allclose(bu, g["als_bu"], atol=2e-6)
    np.testing.assert_allclose(bi, g["als_bi"], atol=2e-6)
    np.testing.assert_allclose(rm, g["als_rmse"], atol=2e-6)
    kat = json.load(open(os.path.join(golden_dir, "kat.json")))["als"]
    Xk = np.array(kat["X"])
    bu, bi, rm = bm._als(Xk, kat["mu"], np.zeros(3), np.zeros(2), kat["n_epochs"], kat["reg"], 0)
    np.testing.assert_allclose(bu, kat["bu"], atol=1e-6)
    np.testing.assert_allclose(bi, kat["bi"], atol=1e-6)
    np.testing.assert_allclose(rm, kat["train_rmse"], atol=1e-6)


def test_csr_csc_bit_exact_vs_scipy():
    import scipy.sparse as sp
    import torch
    from matrix_factorization_b200 import engine

    rng = np.random.default_rng(4)
    U, I, N = 257, 131, 6000
    keys = rng.choice(U * I, N, replace=False)
    u, i = (keys // I).astype(np.int32), (keys % I).astype(np.int32)
    r = rng.integers(1, 11, N).astype(np.float32) / 2
    csr = engine.Csr(torch.tensor(u).cuda(), torch.tensor(i).cuda(), torch.tensor(r).cuda(), U, I)
    row_ptr, col, val, col_ptr, row, cval = (t.cpu().numpy() for t in csr.export())
    ref = sp.coo_matrix((r, (u, i)), shape=(U, I)).tocsr()
    ref.sort_indices()
    assert np.array_equal(row_ptr, ref.indptr) and np.array_equal(col, ref.indices) and np.array_equal(val, ref.data)
    refc = ref.tocsc()
    refc.sort_indices()
    assert np.array_equal(col_ptr, refc.indptr) and np.array_equal(row, refc.indices) and np.array_equal(cval, refc.data)


def test_bias_predict_and_rmse(golden_dir):
    from matrix_factorization_b200 import baseline_model as bm

    g = np.load(os.path.join(golden_dir, "baseline.npz"))
    Xp = np.stack([g["pred_u"], g["pred_i"]], axis=1)
    for bound, key in [(True, "pred_bound"), (False, "pred_unbound")]:
        pred, poss = bm._predict(Xp, float(g["mu"]), 0, 5, g["als_bu"] * 8, g["als_bi"] * 8, bound)
        np.testing.assert_allclose(pred, g[key], atol=5e-6)
        assert poss == g["possible"].tolist()
    X = np.stack([g["u"], g["i"], g["r"]], axis=1).astype(np.float64)
    assert abs(bm._calculate_rmse(X, float(g["mu"]), g["bu"], g["bi"]) - float(g["rmse"][-1])) < 1e-5


@pytest.mark.parametrize("method", ["sgd", "als"])
def test_baseline_fit_rmse_within_reference_band(golden_dir, method):
    import matrix_factorization_b200 as mfb
    from matrix_factorization_b200.data import synth_ratings, split_rows

    res = json.load(open(os.path.join(golden_dir, "fit_rmse.json")))
    d = res["data"]
    df = synth_ratings(d["n_users"], d["n_items"], d["n_ratings"], seed=d["seed"], min_per_user=d["min_per_user"])
    train, test = split_rows(df, d["test_frac"], seed=d["split_seed"])
    np.random.seed(d["np_seed"])
    key = "baseline_" + method
    m = mfb.BaselineModel(verbose=0, **res[key]["params"]).fit(train[["user_id", "item_id"]], train["rating"])
    pred = np.array(m.predict(test[["user_id", "item_id"]]))
    test_rmse = float(np.sqrt(np.mean((pred - test.rating.to_numpy()) ** 2)))
    ref_tr, ref_te = np.array(res[key]["train_rmse"]), np.array(res[key]["test_rmse"])
    tol = 1e-3 if method == "sgd" else 2e-5
    assert abs(m.train_rmse[-1] - ref_tr.mean()) < tol + 2 * (ref_tr.max() - ref_tr.min())
    assert abs(test_rmse - ref_te.mean()) < tol + 2 * (ref_te.max() - ref_te.min())
    rec = m.recommend(user=train.user_id.iloc[0], amount=5)
    top_items = sorted(m.item_id_map, key=lambda k: -m.item_biases[m.item_id_map[k]])[:5]
    assert rec.item_id.tolist() == top_items  # "most popular" behaviour of the baseline
