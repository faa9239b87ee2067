"""CPU: the C/numpy oracle against fixtures produced by the reference itself
(oracle/gen_golden.py).  This is what pins the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

KN = ["linear", "sigmoid", "rbf"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_kat_updates(golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "kat.json")))
    b = kat["inputs"]
    for name, g in kat.items():
        if not isinstance(g, dict) or "kernel" not in g:
            continue
        P, Q, bu, bi = orc.kmf_replay(KN[g["kernel"]], [0], [0], [b["r"]], None, b["mu"], [b["bu"]],
                                      [b["bi"]], [b["p"]], [b["q"]], b["lr"], b["reg"], b["gamma"],
                                      b["a"], b["a"] + b["c"], g["upd_user"], g["upd_item"])
        np.testing.assert_allclose(P[0], g["p"], rtol=0, atol=1e-15, err_msg=name)
        np.testing.assert_allclose(Q[0], g["q"], rtol=0, atol=1e-15, err_msg=name)
        assert abs(bu[0] - g["bu"]) < 1e-15 and abs(bi[0] - g["bi"]) < 1e-15, name
    # SURVEY 9.2 literal values (typed in from the survey, independent of gen_golden)
    assert kat["linear_tt"]["bu"] == pytest.approx(0.11089500000000001, abs=1e-16)
    assert kat["rbf_tt"]["q"][1] == pytest.approx(-0.10192162190327958, abs=1e-16)
    assert kat["sigmoid_tt"]["p"][0] == pytest.approx(0.09988598777295116, abs=1e-16)
    for kname in KN:
        pred, _ = orc.kmf_predict(kname, [0], [0], b["mu"], [b["bu"]], [b["bi"]], [b["p"]], [b["q"]],
                                  b["gamma"], b["a"], b["a"] + b["c"], bound_ratings=False)
        assert pred[0] == pytest.approx(kat["pred_" + kname], abs=1e-14)
    assert kat["pred_linear"] == pytest.approx(2.9099999999999997, abs=1e-15)
    assert kat["pred_sigmoid"] == pytest.approx(4.741692822669576, abs=1e-15)
    assert kat["pred_rbf"] == pytest.approx(4.685337316887017, abs=1e-15)


def test_kat_als(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "kat.json")))["als"]
    X = np.array(g["X"])
    bu, bi, rm = orc.bias_als(X[:, 0], X[:, 1], X[:, 2], g["mu"], 3, 2, g["n_epochs"], g["reg"])
    np.testing.assert_allclose(bu, g["bu"], atol=1e-15)
    np.testing.assert_allclose(bi, g["bi"], atol=1e-15)
    np.testing.assert_allclose(rm, g["train_rmse"], atol=1e-15)
    np.testing.assert_allclose(bu, [0.6963809523809525, -0.5036190476190475, -0.14603174603174596], atol=1e-15)
    np.testing.assert_allclose(rm, [0.4531588299592849, 0.34639121151372754], atol=1e-15)


@pytest.mark.parametrize("kname", KN)
@pytest.mark.parametrize("flags", ["11", "10"])
def test_replay_rmse_predict(golden_dir, kname, flags):
    g = _load(golden_dir, f"replay_{kname}_{flags}.npz")
    P, Q, bu, bi = orc.kmf_replay(kname, g["u"], g["i"], g["r"], g["order"], float(g["mu"]), g["bu0"],
                                  g["bi0"], g["P0"], g["Q0"], float(g["lr"]), float(g["reg"]),
                                  float(g["gamma"]), 0.0, 5.0, flags[0] == "1", flags[1] == "1")
    for a, b in [(P, g["P"]), (Q, g["Q"]), (bu, g["bu"]), (bi, g["bi"])]:
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-12)
    if flags[1] == "0":  # update_users flavour: item side bit-unchanged
        assert np.array_equal(Q, g["Q0"]) and np.array_equal(bi, g["bi0"])
    rmse = orc.kmf_rmse(kname, g["u"], g["i"], g["r"], float(g["mu"]), bu, bi, P, Q, float(g["gamma"]))
    assert rmse == pytest.approx(float(g["rmse"]), abs=1e-12)
    for bound, key in [(True, "pred_bound"), (False, "pred_unbound")]:
        pred, poss = orc.kmf_predict(kname, g["pred_u"], g["pred_i"], float(g["mu"]), bu, bi, P, Q,
                                     float(g["gamma"]), 0.0, 5.0, bound)
        np.testing.assert_allclose(pred, g[key], rtol=0, atol=1e-12)
        assert np.array_equal(poss, g["possible"])


@pytest.mark.parametrize("kname", KN)
def test_reference_njit_sgd_replayed(golden_dir, kname):
    """Three epochs of the reference's own njit _sgd, replayed by the oracle in the recorded order."""
    g = _load(golden_dir, f"sgd_{kname}.npz")
    P, Q = g["P0"], g["Q0"]
    bu, bi = np.zeros(P.shape[0]), np.zeros(Q.shape[0])
    for e in range(g["orders"].shape[0]):
        P, Q, bu, bi = orc.kmf_replay(kname, g["u"], g["i"], g["r"], g["orders"][e], float(g["mu"]), bu,
                                      bi, P, Q, float(g["lr"]), float(g["reg"]), float(g["gamma"]))
        rmse = orc.kmf_rmse(kname, g["u"], g["i"], g["r"], float(g["mu"]), bu, bi, P, Q, float(g["gamma"]))
        assert rmse == pytest.approx(float(g["rmse"][e]), abs=1e-12)
    for a, b in [(P, g["P"]), (Q, g["Q"]), (bu, g["bu"]), (bi, g["bi"])]:
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-12)


def test_baseline_loops(golden_dir):
    g = _load(golden_dir, "baseline.npz")
    mu = float(g["mu"])
    bu, bi = np.zeros(60), np.zeros(40)
    for e in range(3):
        bu, bi = orc.bias_replay(g["u"], g["i"], g["r"], g["orders"][e], mu, bu, bi, float(g["lr"]), float(g["reg"]))
        assert orc.bias_rmse(g["u"], g["i"], g["r"], mu, bu, bi) == pytest.approx(float(g["rmse"][e]), abs=1e-13)
    np.testing.assert_allclose(bu, g["bu"], atol=1e-13)
    np.testing.assert_allclose(bi, g["bi"], atol=1e-13)
    bu2, bi2 = orc.bias_replay(g["u"], g["i"], g["r"], g["order_frozen"], mu, bu, bi, float(g["lr"]),
                               float(g["reg"]), True, False)
    np.testing.assert_allclose(bu2, g["bu_frozen"], atol=1e-13)
    assert np.array_equal(bi2, g["bi_frozen"])
    abu, abi, arm = orc.bias_als(g["u"], g["i"], g["r"], mu, 60, 40, 4, float(g["als_reg"]))
    np.testing.assert_allclose(abu, g["als_bu"], atol=1e-13)
    np.testing.assert_allclose(abi, g["als_bi"], atol=1e-13)
    np.testing.assert_allclose(arm, g["als_rmse"], atol=1e-13)
    for bound, key in [(True, "pred_bound"), (False, "pred_unbound")]:
        pred, poss = orc.bias_predict(g["pred_u"], g["pred_i"], mu, g["als_bu"] * 8, g["als_bi"] * 8, 0.0, 5.0, bound)
        np.testing.assert_allclose(pred, g[key], atol=1e-13)
        assert np.array_equal(poss, g["possible"])


def test_port_sgd_converges_like_reference(golden_dir):
    """The port's _sgd (own shuffle stream) lands in the reference's RMSE band on the fit fixture."""
    from matrix_factorization_b200.data import synth_ratings, split_rows

    res = json.load(open(os.path.join(golden_dir, "fit_rmse.json")))
    d = res["data"]
    df = synth_ratings(d["n_users"], d["n_items"], d["n_ratings"], seed=d["seed"], min_per_user=d["min_per_user"])
    train, _ = split_rows(df, d["test_frac"], seed=d["split_seed"])
    np.random.seed(d["np_seed"])
    u, i, r, umap, imap, _ = orc.preprocess_fit(train.user_id.to_numpy(), train.item_id.to_numpy(), train.rating.to_numpy())
    kw = res["linear"]["params"]
    mu = float(r.mean())
    P0 = np.random.normal(0, 0.1, (len(umap), kw["n_factors"]))
    Q0 = np.random.normal(0, 0.1, (len(imap), kw["n_factors"]))
    *_, rm = orc.kmf_sgd("linear", u, i, r, mu, np.zeros(len(umap)), np.zeros(len(imap)), P0, Q0,
                         kw["n_epochs"], kw["lr"], kw["reg"], seed=5)
    ref = np.array(res["linear"]["train_rmse"])
    assert abs(rm[-1] - ref.mean()) < 1e-3


def test_preprocess_restatement(golden_dir):
    g = _load(golden_dir, "preprocess.npz")
    np.random.seed(int(g["fit_seed"]))
    u, i, r, umap, imap, _ = orc.preprocess_fit(g["train_user"], g["train_item"], g["train_rating"])
    assert np.array_equal(u, g["fit_u"]) and np.array_equal(i, g["fit_i"]) and np.array_equal(r, g["fit_r"])
    assert list(umap.keys()) == g["umap_keys"].tolist() and list(imap.keys()) == g["imap_keys"].tolist()
    pu, pi = orc.map_predict(g["pq_user"], g["pq_item"], umap, imap)
    assert np.array_equal(pu, g["pq_u"]) and np.array_equal(pi, g["pq_i"])
    with pytest.raises(ValueError):
        orc.preprocess_fit([1, 1], [2, 2], [3.0, 4.0])
