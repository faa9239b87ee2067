"""GPU parity tests of the KernelMF path, through the C ABI (ctypes) -- compared with the fp64
oracle on the same seeded inputs and with the fixtures generated from the reference."""
import json
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

KN = ["linear", "sigmoid", "rbf"]
HP = {"linear": (0.01, 0.01), "sigmoid": (0.05, 0.01), "rbf": (0.3, 0.05)}  # lr, gamma


def _mods():
    from matrix_factorization_b200 import kernel_matrix_factorization as kmf
    from oracle import oracle as orc

    return kmf, orc


def _problem(seed, U, I, N, F, hot=0.0):
    rng = np.random.default_rng(seed)
    keys = rng.choice(U * I, N, replace=False)
    u, i = (keys // I).astype(np.int64), (keys % I).astype(np.int64)
    if hot > 0:  # concentrate a share of the ratings on item 0 / user 0 (hot chains)
        m = rng.random(N) < hot
        i[m] = 0
        keep = np.unique(u * I + i, return_index=True)[1]
        u, i = u[keep], i[keep]
    r = rng.integers(1, 6, len(u)).astype(np.float64)
    P = rng.normal(0, 0.1, (U, F))
    Q = rng.normal(0, 0.1, (I, F))
    return u, i, r, P, Q, rng.normal(0, 0.05, U), rng.normal(0, 0.05, I)


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-12, np.max(np.abs(b))))


def test_plan_is_conflict_free_and_order_is_permutation():
    import torch
    from matrix_factorization_b200 import engine

    u, i, r, *_ = _problem(1, 300, 200, 9000, 4, hot=0.05)
    du, di, dr = (torch.tensor(u, dtype=torch.int32).cuda(), torch.tensor(i, dtype=torch.int32).cuda(),
                  torch.tensor(r, dtype=torch.float32).cuda())
    for opts in [dict(), dict(n_workers=16, warps_per_cta=4), dict(n_workers=7, warps_per_cta=7), dict(n_workers=1, warps_per_cta=1),
                 dict(stripe_slack=1), dict(stripe_slack=3), dict(n_workers=16, warps_per_cta=4, stripe_slack=4),
                 dict(schedule=2), dict(schedule=1)]:
        plan = engine.Plan(du, di, dr, 300, 200, n_factors=4, **opts)
        info = plan.info()
        W, R = info["n_workers"], info["n_steps"]
        assert W == info["n_ctas"] * info["warps_per_cta"]
        # (stripes per worker as asked, unless there are too few users to fill them)
        assert R % W == 0 and (opts.get("stripe_slack", 0) in (0, R // W) or R // W == max(1, 300 // W))
        w, s = (t.cpu().numpy().astype(np.int64) for t in plan.assignment())
        order = plan.order().cpu().numpy()
        assert np.array_equal(np.sort(order), np.arange(len(u)))          # permutation
        assert np.all(np.diff(s[order]) >= 0)                              # step-major
        assert s.min() >= 0 and s.max() < R and w.max() < W
        # a user's stripe (c * worker + step) mod R is the same wherever the user is rated: the stripe travels
        # around the ring, worker w + 1 hands it to worker w with a lag of c = R / W steps
        stripe = ((R // W) * w + s) % R
        first = np.full(300, -1, np.int64)
        first[u[::-1]] = stripe[::-1]
        assert np.array_equal(stripe, first[u])
        # inside one step (wave) a user / an item is touched by exactly one worker
        for ids in (u, i):
            key = s * (ids.max() + 1) + ids
            srt = np.argsort(key, kind="stable")
            same = key[srt][1:] == key[srt][:-1]
            assert np.all(w[srt][1:][same] == w[srt][:-1][same])
        assert info["max_item_degree"] == np.bincount(i).max() and info["max_user_degree"] == np.bincount(u).max()
        plan.close()


@pytest.mark.parametrize("kname", KN)
@pytest.mark.parametrize("flags", ["11", "10"])
def test_one_epoch_matches_reference_replay(golden_dir, kname, flags):
    """Golden inputs; the GPU epoch equals the reference update rule replayed in the emitted order."""
    kmf, orc = _mods()
    g = np.load(os.path.join(golden_dir, f"replay_{kname}_{flags}.npz"))
    uu, ui = flags[0] == "1", flags[1] == "1"
    for opts in [None, dict(n_workers=6, warps_per_cta=3), dict(n_workers=24, warps_per_cta=4),
                 dict(n_workers=6, warps_per_cta=3, stripe_slack=2), dict(stripe_slack=4),
                 dict(schedule=2), dict(n_workers=6, warps_per_cta=3, stripe_slack=2, schedule=2)]:
        P, Q, bu, bi = (g[k].copy() for k in ("P0", "Q0", "bu0", "bi0"))
        X = np.stack([g["u"], g["i"], g["r"]], axis=1).astype(np.float64)
        P2, Q2, bu2, bi2, rm, order = kmf._sgd(X, float(g["mu"]), bu, bi, P, Q, 1, kname, float(g["gamma"]),
                                              float(g["lr"]), float(g["reg"]), 0.0, 5.0, 0, uu, ui,
                                              plan_options=opts, return_order=True)
        assert P2 is P and Q2 is Q  # in place, like the reference
        Po, Qo, buo, bio = orc.kmf_replay(kname, g["u"], g["i"], g["r"], order, float(g["mu"]), g["bu0"], g["bi0"],
                                          g["P0"], g["Q0"], float(g["lr"]), float(g["reg"]), float(g["gamma"]),
                                          0.0, 5.0, uu, ui)
        assert _rel(P, Po) < 1e-4 and _rel(Q, Qo) < 1e-4
        assert np.max(np.abs(bu - buo)) < 1e-5 and np.max(np.abs(bi - bio)) < 1e-5
        if not ui:  # update_users flavour: the item side is bit-unchanged
            assert np.array_equal(Q, g["Q0"]) and np.array_equal(bi, g["bi0"])
        ref_rmse = orc.kmf_rmse(kname, g["u"], g["i"], g["r"], float(g["mu"]), buo, bio, Po, Qo, float(g["gamma"]))
        assert abs(rm[0] - ref_rmse) < 1e-5


def test_single_update_kat(golden_dir):
    kmf, _ = _mods()
    kat = json.load(open(os.path.join(golden_dir, "kat.json")))
    b = kat["inputs"]
    for name, g in kat.items():
        if not isinstance(g, dict) or "kernel" not in g:
            continue
        P, Q = np.array([b["p"]]), np.array([b["q"]])
        bu, bi = np.array([b["bu"]]), np.array([b["bi"]])
        kmf._sgd(np.array([[0, 0, b["r"]]]), b["mu"], bu, bi, P, Q, 1, KN[g["kernel"]], b["gamma"], b["lr"],
                 b["reg"], b["a"], b["a"] + b["c"], 0, g["upd_user"], g["upd_item"])
        np.testing.assert_allclose(P[0], g["p"], rtol=2e-6, atol=1e-7, err_msg=name)
        np.testing.assert_allclose(Q[0], g["q"], rtol=2e-6, atol=1e-7, err_msg=name)
        assert abs(bu[0] - g["bu"]) < 1e-6 and abs(bi[0] - g["bi"]) < 1e-6, name


@pytest.mark.parametrize("kname", KN)
def test_rmse_and_predict_match_reference(golden_dir, kname):
    kmf, _ = _mods()
    g = np.load(os.path.join(golden_dir, f"replay_{kname}_11.npz"))
    X = np.stack([g["u"], g["i"], g["r"]], axis=1).astype(np.float64)
    rmse = kmf._calculate_rmse(X, float(g["mu"]), g["bu"], g["bi"], g["P"], g["Q"], 0.0, 5.0, kname, float(g["gamma"]))
    assert abs(rmse - float(g["rmse"])) < 1e-5
    Xp = np.stack([g["pred_u"], g["pred_i"]], axis=1)
    for bound, key in [(True, "pred_bound"), (False, "pred_unbound")]:
        pred, poss = kmf._predict(Xp, float(g["mu"]), g["bu"], g["bi"], g["P"], g["Q"], 0, 5, kname, float(g["gamma"]), bound)
        assert isinstance(pred, list) and isinstance(poss, list) and isinstance(poss[0], bool)
        np.testing.assert_allclose(pred, g[key], rtol=0, atol=5e-6)
        assert poss == g["possible"].tolist()


@pytest.mark.parametrize("kname,F,U,I,N,hot", [
    ("linear", 100, 943, 1682, 100_000, 0.0),   # config 1 shape
    ("linear", 128, 600, 400, 30_000, 0.1),     # full 512-byte rows, hot item chain
    ("sigmoid", 100, 500, 300, 20_000, 0.05),
    ("rbf", 100, 500, 300, 20_000, 0.05),
    ("linear", 256, 300, 200, 8_000, 0.1),      # NV = 2 (Netflix config row width)
    ("linear", 50, 300, 200, 8_000, 0.0),       # n_factors % 4 != 0 -> padded row stride
    ("linear", 300, 200, 150, 4_000, 0.0),      # NV = 4
    ("rbf", 600, 120, 100, 2_000, 0.0),         # NV = 8
])
def test_one_epoch_parity_default_plan(kname, F, U, I, N, hot):
    """Seeded medium problems with the auto-sized plan: factors within 1e-4 relative of the replay."""
    kmf, orc = _mods()
    lr, gamma = HP[kname]
    u, i, r, P, Q, bu, bi = _problem(F + N, U, I, N, F, hot=hot)
    if kname == "rbf":
        bu[:] = 0
        bi[:] = 0
    P0, Q0, bu0, bi0 = P.copy(), Q.copy(), bu.copy(), bi.copy()
    mu = float(r.mean())
    *_, rm, order = kmf._sgd((u, i, r), mu, bu, bi, P, Q, 1, kname, gamma, lr, 0.02, 0.0, 5.0, 0, return_order=True)
    Po, Qo, buo, bio = orc.kmf_replay(kname, u, i, r, order, mu, bu0, bi0, P0, Q0, lr, 0.02, gamma)
    assert _rel(P, Po) < 1e-4 and _rel(Q, Qo) < 1e-4, (_rel(P, Po), _rel(Q, Qo))
    assert np.max(np.abs(bu - buo)) < 2e-5 and np.max(np.abs(bi - bio)) < 2e-5
    assert abs(rm[0] - orc.kmf_rmse(kname, u, i, r, mu, buo, bio, Po, Qo, gamma)) < 2e-5


def test_multi_epoch_replay_and_determinism():
    kmf, orc = _mods()
    u, i, r, P, Q, bu, bi = _problem(77, 400, 250, 15_000, 64, hot=0.05)
    mu = float(r.mean())
    runs = []
    for _ in range(2):
        Pa, Qa, bua, bia = P.copy(), Q.copy(), bu.copy(), bi.copy()
        *_, rm, order = kmf._sgd((u, i, r), mu, bua, bia, Pa, Qa, 3, "linear", 0.01, 0.01, 0.02, 0.0, 5.0, 0, return_order=True)
        runs.append((Pa, Qa, bua, bia, rm))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])  # bit-deterministic
    Po, Qo, buo, bio = P, Q, bu, bi
    for e in range(3):
        Po, Qo, buo, bio = orc.kmf_replay("linear", u, i, r, order, mu, buo, bio, Po, Qo, 0.01, 0.02)
        assert abs(runs[0][4][e] - orc.kmf_rmse("linear", u, i, r, mu, buo, bio, Po, Qo)) < 2e-5
    assert _rel(runs[0][0], Po) < 1e-4 and _rel(runs[0][1], Qo) < 1e-4


def test_edge_cases():
    kmf, _ = _mods()
    P, Q, bu, bi = np.zeros((3, 8)) + 0.1, np.zeros((2, 8)) + 0.1, np.zeros(3), np.zeros(2)
    # empty rating set: nothing changes, no crash
    out = kmf._sgd(np.zeros((0, 3)), 3.0, bu, bi, P, Q, 2, "linear", 0.1, 0.01, 0.02, 0, 5, 0)
    assert len(out[4]) == 2 and np.all(P == np.float64(np.float32(0.1)))  # device storage is fp32
    assert kmf._predict(np.zeros((0, 2)), 3.0, bu, bi, P, Q, 0, 5, "linear", 0.1, True) == ([], [])
    # a single user with every item (maximally ragged): one worker chain
    X = np.array([[0, 0, 4.0], [0, 1, 2.0]])
    kmf._sgd(X, 3.0, bu, bi, P, Q, 1, "linear", 0.1, 0.01, 0.02, 0, 5, 0)
    assert np.all(P[1:] == np.float64(np.float32(0.1))) and not np.any(P[0] == np.float64(np.float32(0.1)))
    with pytest.raises(ValueError):
        kmf._sgd(X, 3.0, bu, bi, P, Q, 1, "poly", 0.1, 0.01, 0.02, 0, 5, 0)
    # ids out of range are rejected by the C ABI, not silently clipped
    from matrix_factorization_b200._lib import MfkError
    with pytest.raises(MfkError):
        kmf._sgd(np.array([[7, 0, 1.0]]), 3.0, bu, bi, P, Q, 1, "linear", 0.1, 0.01, 0.02, 0, 5, 0)


@pytest.mark.parametrize("kname", KN)
def test_full_fit_rmse_within_reference_band(golden_dir, kname):
    """KernelMF.fit end to end (same np seed => same shuffle, id maps and init as the reference run):
    final train / test RMSE within 1e-3 of the reference (plus its own order-to-order spread)."""
    import matrix_factorization_b200 as mfb
    from matrix_factorization_b200.data import synth_ratings, split_rows

    res = json.load(open(os.path.join(golden_dir, "fit_rmse.json")))
    d = res["data"]
    df = synth_ratings(d["n_users"], d["n_items"], d["n_ratings"], seed=d["seed"], min_per_user=d["min_per_user"])
    train, test = split_rows(df, d["test_frac"], seed=d["split_seed"])
    np.random.seed(d["np_seed"])
    m = mfb.KernelMF(verbose=0, **res[kname]["params"]).fit(train[["user_id", "item_id"]], train["rating"])
    pred = np.array(m.predict(test[["user_id", "item_id"]]))
    test_rmse = float(np.sqrt(np.mean((pred - test.rating.to_numpy()) ** 2)))
    ref_tr, ref_te = np.array(res[kname]["train_rmse"]), np.array(res[kname]["test_rmse"])
    tol_tr = 1e-3 + 2 * (ref_tr.max() - ref_tr.min())
    tol_te = 1e-3 + 2 * (ref_te.max() - ref_te.min())
    assert len(m.train_rmse) == res[kname]["params"]["n_epochs"]
    assert abs(m.train_rmse[-1] - ref_tr.mean()) < tol_tr, (m.train_rmse[-1], ref_tr)
    assert abs(test_rmse - ref_te.mean()) < tol_te, (test_rmse, ref_te)
    assert m.user_features.dtype == np.float64 and m.user_features.shape == (m.n_users, m.n_factors)
    assert isinstance(m.train_rmse, list) and isinstance(m.global_mean, float)


def test_estimator_api_update_users_recommend_pickle(golden_dir):
    import pickle
    import matrix_factorization_b200 as mfb
    from matrix_factorization_b200.data import synth_ratings

    df = synth_ratings(120, 90, 4000, seed=21, min_per_user=8)
    np.random.seed(4)
    Xi, yi, Xu, yu, Xt, yt = mfb.train_update_test_split(df, frac_new_users=0.2)
    m = mfb.KernelMF(n_factors=12, n_epochs=8, lr=0.01, reg=0.02, verbose=0).fit(Xi, yi)
    n_users0, Q0, bi0 = m.n_users, m.item_features.copy(), m.item_biases.copy()
    P_old = m.user_features.copy()
    m.update_users(Xu, yu, lr=0.01, n_epochs=5, verbose=0)
    new_users = set(Xu.user_id) - set(Xi.user_id)
    assert m.user_features.shape[0] == n_users0 + len(new_users) and m.n_users == n_users0  # quirk
    assert np.array_equal(m.item_features, Q0) and np.array_equal(m.item_biases, bi0)        # items frozen
    assert np.array_equal(m.user_features[:n_users0], P_old)                                  # old users untouched
    assert len(m.train_rmse) == 5
    pred = m.predict(Xt)
    assert len(pred) == len(Xt) and all(0 <= p <= 5 for p in pred)
    assert m.predictions_possible.count(True) + m.predictions_possible.count(False) == len(Xt)
    # cold start: unknown user and item -> global mean (linear kernel)
    cold = m.predict(pd.DataFrame({"user_id": [10**9], "item_id": [10**9]}), bound_ratings=False)
    assert abs(cold[0] - m.global_mean) < 1e-6 and m.predictions_possible == [False]
    # recommend: agrees with predict-all + sort, known items excluded
    user = Xi.user_id.iloc[0]
    known = Xi.loc[Xi.user_id == user, "item_id"].tolist()
    rec = m.recommend(user=user, amount=7, items_known=known + [123456789], include_user=False)
    assert list(rec.columns) == ["item_id", "rating_pred"] and len(rec) == 7
    assert not set(rec.item_id) & set(known)
    cand = [it for it in m.item_id_map if it not in set(known)]
    allp = np.array(m.predict(pd.DataFrame({"user_id": user, "item_id": cand}), bound_ratings=False))
    top = np.argsort(-allp, kind="stable")[:7]
    assert rec.index.tolist() == top.tolist()
    np.testing.assert_allclose(rec.rating_pred.to_numpy(), np.clip(allp[top], 0, 5), atol=1e-5)
    # pickle round trip (device mirrors are not part of the state)
    m2 = pickle.loads(pickle.dumps(m))
    assert np.array_equal(m2.user_features, m.user_features)
    assert m2.predict(Xt) == pred
    # batched recommend agrees with the per-user call
    ra = m.recommend_all(users=[user], amount=7, items_known=Xi.assign(rating=yi))
    assert ra.item_id.tolist() == rec.item_id.tolist()


def test_recommend_matches_reference_lists(golden_dir):
    import matrix_factorization_b200 as mfb

    g = np.load(os.path.join(golden_dir, "recommend.npz"))
    recs = json.loads(str(g["recs"]))
    m = mfb.KernelMF(n_factors=8, verbose=0)
    m.user_features, m.item_features = g["P"].copy(), g["Q"].copy()
    m.user_biases, m.item_biases, m.global_mean = g["bu"].copy(), g["bi"].copy(), float(g["mu"])
    m.user_id_map = {k: j for j, k in enumerate(g["umap_keys"].tolist())}
    m.item_id_map = {k: j for j, k in enumerate(g["imap_keys"].tolist())}
    m.n_users, m.n_items = len(m.user_id_map), len(m.item_id_map)
    for uid, ref in recs.items():
        rec = m.recommend(user=int(uid), amount=10, items_known=ref["known"])
        got_items, ref_items = rec.item_id.tolist(), ref["items"]
        got_s, ref_s = rec.rating_pred.to_numpy(), np.array(ref["scores"])
        np.testing.assert_allclose(got_s, ref_s, atol=1e-5)
        for a, b, sa in zip(got_items, ref_items, range(10)):
            if a != b:  # only allowed inside a tie (within 1e-5)
                assert np.sum(np.abs(ref_s - ref_s[sa]) < 1e-5) > 1, (uid, sa, a, b)
        same = [a == b for a, b in zip(got_items, ref_items)]
        assert rec.index.tolist() == ref["index"] or not all(same)
        assert rec.user_id.tolist() == [int(uid)] * 10


@pytest.mark.parametrize("F,U,I,N,hot,min_deg", [
    (128, 900, 300, 40_000, 0.5, 200),    # several hot items, full 512-byte rows, batches of 64 + ragged tails
    (100, 600, 200, 20_000, 0.4, 100),    # n_factors % 128 != 0
    (256, 400, 150, 12_000, 0.4, 150),    # NV = 2
    (32, 300, 100, 6_000, 0.3, 64),
])
def test_hot_item_minibatch_path_matches_replay(F, U, I, N, hot, min_deg):
    """Hot/cold split plan: the most-rated items are resolved by the CTA-cooperative exact mini-batch kernel
    (Gram matrix + forward substitution); the emitted order (hot phase, then cold phase) replays to the same
    factors through the sequential fp64 oracle."""
    import torch
    from matrix_factorization_b200 import engine
    kmf, orc = _mods()
    rng = np.random.default_rng(F + N)
    # a handful of very popular items on top of a uniform background
    n_hot_items = 5
    keys = rng.choice(U * I, N, replace=False)
    u, i = (keys // I).astype(np.int64), (keys % I).astype(np.int64)
    m = rng.random(len(u)) < hot
    i[m] = rng.integers(0, n_hot_items, m.sum())
    keep = np.unique(u * I + i, return_index=True)[1]
    u, i = u[keep], i[keep]
    r = rng.integers(1, 6, len(u)).astype(np.float64)
    P, Q = rng.normal(0, 0.1, (U, F)), rng.normal(0, 0.1, (I, F))
    bu, bi = rng.normal(0, 0.05, U), rng.normal(0, 0.05, I)
    P0, Q0, bu0, bi0 = P.copy(), Q.copy(), bu.copy(), bi.copy()
    mu, lr, reg = float(r.mean()), 0.01, 0.02
    # the plan really is split, and every wave (hot steps first, then cold steps) is conflict-free
    plan = engine.Plan(torch.tensor(u, dtype=torch.int32).cuda(), torch.tensor(i, dtype=torch.int32).cuda(),
                       torch.tensor(r, dtype=torch.float32).cuda(), U, I, n_factors=F, hot_min_degree=min_deg)
    info = plan.info()
    assert info["n_hot_items"] >= n_hot_items - 1 and info["n_hot_ratings"] > 0 and info["n"] == len(u)
    w, s = (t.cpu().numpy().astype(np.int64) for t in plan.assignment())
    order = plan.order().cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(len(u))) and np.all(np.diff(s[order]) >= 0)
    for ids in (u, i):
        key = s * (ids.max() + 1) + ids
        srt = np.argsort(key, kind="stable")
        same = key[srt][1:] == key[srt][:-1]
        assert np.all(w[srt][1:][same] == w[srt][:-1][same])
    plan.close()
    for uu, ui in [(True, True), (True, False)]:
        Pa, Qa, bua, bia = P0.copy(), Q0.copy(), bu0.copy(), bi0.copy()
        *_, rm, order = kmf._sgd((u, i, r), mu, bua, bia, Pa, Qa, 2, "linear", 0.01, lr, reg, 0.0, 5.0, 0, uu, ui,
                                 plan_options=dict(hot_min_degree=min_deg), return_order=True)
        Po, Qo, buo, bio = P0, Q0, bu0, bi0
        for e in range(2):
            Po, Qo, buo, bio = orc.kmf_replay("linear", u, i, r, order, mu, buo, bio, Po, Qo, lr, reg, 0.01, 0.0, 5.0, uu, ui)
            assert abs(rm[e] - orc.kmf_rmse("linear", u, i, r, mu, buo, bio, Po, Qo)) < 2e-5
        assert _rel(Pa, Po) < 1e-4 and _rel(Qa, Qo) < 1e-4, (_rel(Pa, Po), _rel(Qa, Qo))
        assert np.max(np.abs(bua - buo)) < 2e-5 and np.max(np.abs(bia - bio)) < 2e-5
        if not ui:
            assert np.array_equal(Qa, Q0) and np.array_equal(bia, bi0)


@pytest.mark.parametrize("kname", KN)
@pytest.mark.parametrize("W,slack", [(1, 1), (1, 2), (2, 2), (4, 1), (8, 2), (16, 3), (64, 1)])
def test_flat_schedule_every_worker_geometry(kname, W, slack):
    """The flat (dataflow) schedule with few workers: cells of many chunks, row hand-over inside and across chunks,
    item chains that continue from one cell into the next -- every geometry replays exactly."""
    kmf, orc = _mods()
    lr, gamma = HP[kname]
    u, i, r, P, Q, bu, bi = _problem(W * 10 + slack, 300, 200, 20_000, 32, hot=0.05)
    if kname == "rbf":
        bu[:] = 0
        bi[:] = 0
    P0, Q0, bu0, bi0 = P.copy(), Q.copy(), bu.copy(), bi.copy()
    mu = float(r.mean())
    opts = dict(schedule=3, n_workers=W, stripe_slack=slack, hot_min_degree=0xFFFFFFFF)
    *_, rm, order = kmf._sgd((u, i, r), mu, bu, bi, P, Q, 1, kname, gamma, lr, 0.02, 0.0, 5.0, 0, plan_options=opts,
                             return_order=True)
    Po, Qo, buo, bio = orc.kmf_replay(kname, u, i, r, order, mu, bu0, bi0, P0, Q0, lr, 0.02, gamma)
    assert _rel(P, Po) < 1e-4 and _rel(Q, Qo) < 1e-4, (_rel(P, Po), _rel(Q, Qo))
    assert np.max(np.abs(bu - buo)) < 2e-5 and np.max(np.abs(bi - bio)) < 2e-5
