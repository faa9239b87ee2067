"""CPU: host-side logic of the drop-in API and the C-ABI surface (no compute calls)."""
import json
import os
import pickle
import re

import numpy as np
import pandas as pd
import pytest

import matrix_factorization_b200 as mfb
from matrix_factorization_b200 import _lib
from matrix_factorization_b200.data import synth_ratings, split_rows, SHAPES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    """Every function declared in include/mfk.h is exported by the built .so and bound in _lib."""
    hdr = open(os.path.join(ROOT, "include", "mfk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mfk_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in mfk.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.mfk_abi_version() == 1


def test_cabi_structs_match_the_ctypes_mirrors():
    """Field by field: mfk_plan_opts / mfk_plan_info of include/mfk.h against the ctypes structures of _lib (names, order, widths)."""
    import ctypes as C

    hdr = open(os.path.join(ROOT, "include", "mfk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    widths = {"int32_t": C.c_int32, "uint32_t": C.c_uint32, "int64_t": C.c_int64}
    for cname, mirror in (("mfk_plan_opts", _lib.PlanOpts), ("mfk_plan_info", _lib.PlanInfo)):
        body = re.search(r"typedef struct[^{]*\{([^}]*)\}\s*" + cname + r"\s*;", hdr, flags=re.S).group(1)
        fields = []
        for ctype, names in re.findall(r"\b(u?int(?:32|64)_t)\s+([^;]+);", body):
            fields += [(n.strip(), widths[ctype]) for n in names.split(",")]
        assert fields == list(mirror._fields_), (cname, fields, mirror._fields_)


def test_constructor_defaults_and_errors():
    m = mfb.KernelMF()
    assert (m.n_factors, m.n_epochs, m.kernel, m.reg, m.lr, m.init_mean, m.init_sd) == (100, 100, "linear", 1, 0.01, 0, 0.1)
    assert m.gamma == 0.01 and m.min_rating == 0 and m.max_rating == 5 and m.verbose == 1
    assert mfb.KernelMF(n_factors=50).gamma == 1 / 50
    b = mfb.BaselineModel()
    assert (b.method, b.n_epochs, b.reg, b.lr, b.verbose) == ("sgd", 100, 1, 0.01, 1)
    with pytest.raises(ValueError, match="Kernel must be one of linear, sigmoid, or rbf"):
        mfb.KernelMF(kernel="poly")
    with pytest.raises(ValueError, match='Method param must be either "sgd" or "als"'):
        mfb.BaselineModel(method="x")


def test_sklearn_contract():
    from sklearn.base import clone

    m = mfb.KernelMF(n_factors=50, kernel="rbf", reg=0.1)
    c = clone(m)
    assert c.get_params() == m.get_params() and c.gamma == 0.02
    c.set_params(n_factors=10)
    assert c.gamma == 0.02  # resolved at construction, like the reference
    assert pickle.loads(pickle.dumps(m)).get_params() == m.get_params()
    assert set(mfb.__all__) == {"BaselineModel", "KernelMF", "RecommenderBase", "train_update_test_split"}


def test_preprocess_matches_reference_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    m = mfb.BaselineModel(method="als", verbose=0)
    train = pd.DataFrame({"user_id": g["train_user"], "item_id": g["train_item"]})
    np.random.seed(int(g["fit_seed"]))
    Xf = m._preprocess_data(train, pd.Series(g["train_rating"]), type="fit")
    assert np.array_equal(Xf["user_id"].to_numpy(), g["fit_u"])
    assert np.array_equal(Xf["item_id"].to_numpy(), g["fit_i"])
    assert np.array_equal(Xf["rating"].to_numpy(), g["fit_r"])
    assert list(m.user_id_map.keys()) == g["umap_keys"].tolist()
    assert list(m.item_id_map.keys()) == g["imap_keys"].tolist()
    assert list(m.user_id_map.values()) == list(range(m.n_users))
    # same RNG consumption as the reference: the next draw must agree with a replay
    nxt = np.random.random()
    np.random.seed(int(g["fit_seed"]))
    np.random.choice(len(train), size=len(train), replace=False)
    assert nxt == np.random.random()

    upd = pd.DataFrame({"user_id": g["upd_user"], "item_id": g["upd_item"]})
    np.random.seed(int(g["upd_seed"]))
    Xu, known, new = m._preprocess_data(upd, pd.Series(g["upd_rating"]), type="update")
    assert np.array_equal(Xu["user_id"].to_numpy(), g["upd_u"])
    assert np.array_equal(Xu["item_id"].to_numpy(), g["upd_i"])
    assert np.array_equal(Xu["rating"].to_numpy(), g["upd_r"])
    assert known == g["upd_known"].tolist() and new == g["upd_new"].tolist()
    assert list(m.user_id_map.keys()) == g["umap_keys_after"].tolist()
    assert m.n_users == len(g["umap_keys"])  # quirk: n_users is NOT bumped by update

    pq = pd.DataFrame({"user_id": g["pq_user"], "item_id": g["pq_item"]})
    Xp = m._preprocess_data(pq, type="predict")
    assert np.array_equal(Xp["user_id"].to_numpy(), g["pq_u"]) and np.array_equal(Xp["item_id"].to_numpy(), g["pq_i"])


def test_preprocess_edge_cases():
    m = mfb.BaselineModel(verbose=0)
    X = pd.DataFrame({"user_id": [1, 1, 2], "item_id": [7, 7, 8]})
    with pytest.raises(ValueError, match="Duplicate user-item ratings in matrix"):
        m._preprocess_data(X, pd.Series([1.0, 2.0, 3.0]), type="fit")
    # string ids (the reference itself raises under pandas 3; semantics = first appearance on shuffled rows)
    Xs = pd.DataFrame({"user_id": ["a", "b", "a", "c"], "item_id": ["x", "x", "y", "z"]})
    np.random.seed(3)
    out = m._preprocess_data(Xs, pd.Series([1.0, 2.0, 3.0, 4.0]), type="fit")
    np.random.seed(3)
    perm = np.random.choice(4, size=4, replace=False)
    assert list(m.user_id_map.keys()) == list(dict.fromkeys(Xs.user_id.to_numpy()[perm]))
    assert list(m.item_id_map.keys()) == list(dict.fromkeys(Xs.item_id.to_numpy()[perm]))
    assert out["rating"].tolist() == [[1.0, 2.0, 3.0, 4.0][j] for j in perm]
    # index-aligned rating assignment (recommender_base.py:123)
    Xi = pd.DataFrame({"user_id": [1, 2, 3], "item_id": [4, 5, 6]}, index=[10, 11, 12])
    y = pd.Series([3.0, 2.0, 1.0], index=[12, 11, 10])
    np.random.seed(0)
    o = m._preprocess_data(Xi, y, type="fit")
    raw_u = {v: k for k, v in m.user_id_map.items()}
    got = {raw_u[u]: r for u, r in zip(o["user_id"], o["rating"])}
    assert got == {1: 1.0, 2: 2.0, 3: 3.0}
    # known_users / contains_*
    assert m.known_users == {1, 2, 3} and m.contains_item(5) and not m.contains_user(99)
    # empty predict never touches the device
    assert mfb.KernelMF(verbose=0).predict(pd.DataFrame({"user_id": [], "item_id": []})) == []


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    df = synth_ratings(20, 15, 100, seed=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        mfb.KernelMF(n_factors=4, n_epochs=1, verbose=0).fit(df[["user_id", "item_id"]], df["rating"])


def test_train_update_test_split():
    df = synth_ratings(60, 40, 1500, seed=5, min_per_user=6)
    np.random.seed(1)
    Xi, yi, Xu, yu, Xt, yt = mfb.train_update_test_split(df, frac_new_users=0.25)
    held = set(Xu.user_id) | set(Xt.user_id)
    assert len(held) == round(0.25 * df.user_id.nunique())
    assert not (set(Xi.user_id) & held)
    assert len(Xi) + len(Xu) + len(Xt) == len(df)
    assert set(Xu.user_id) == set(Xt.user_id)
    assert abs(len(Xu) - len(Xt)) <= len(held)
    assert list(Xi.columns) == ["user_id", "item_id"] and yi.name == "rating"


def test_train_update_test_split_matches_reference_same_seed():
    """Pinned to the reference's own output (tests/golden/split.npz, written by oracle/gen_golden.py from
    matrix_factorization/utils.py:8-72): same rows in the same order in all six frames, same RNG consumption."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "split.npz"))
    df = synth_ratings(60, 40, 1500, seed=5, min_per_user=6)
    np.random.seed(int(g["seed"]))
    Xi, yi, Xu, yu, Xt, yt = mfb.train_update_test_split(df, frac_new_users=float(g["frac"]))
    assert np.array_equal(Xi.index.to_numpy(), g["idx_initial"])
    assert np.array_equal(Xu.index.to_numpy(), g["idx_update"])
    assert np.array_equal(Xt.index.to_numpy(), g["idx_test"])
    assert np.array_equal(yi.to_numpy(), g["y_initial"]) and np.array_equal(yu.to_numpy(), g["y_update"])
    assert np.array_equal(yt.to_numpy(), g["y_test"])
    assert np.random.random() == float(g["next_draw"])


def test_synthetic_generator_shapes():
    df = synth_ratings(200, 150, 5000, seed=9, min_per_user=5)
    assert len(df) == 5000 and not df.duplicated(["user_id", "item_id"]).any()
    assert df.user_id.nunique() <= 200 and df.groupby("user_id").size().min() >= 5
    assert set(np.unique(df.rating)) <= {1.0, 2.0, 3.0, 4.0, 5.0}
    # item popularity is skewed (Zipf ~ 1): the top item is far above the median item
    cnt = df.groupby("item_id").size().sort_values(ascending=False)
    assert cnt.iloc[0] > 5 * cnt.median()
    tr, te = split_rows(df, 0.1, seed=0)
    assert len(tr) + len(te) == len(df)
    assert SHAPES["netflix"][0] == 480_189


def test_refit_does_not_reuse_the_item_id_cache():
    """recommend() after a second fit() must map internal ids through the NEW item_id_map
    (every fit reshuffles the rows, so the first-appearance order differs although the length is equal)."""
    df = synth_ratings(40, 30, 600, seed=11, min_per_user=5)
    m = mfb.BaselineModel(method="als", verbose=0)
    np.random.seed(1)
    m._preprocess_arrays(df[["user_id", "item_id"]], df["rating"], type="fit")
    first = m._internal_to_raw_items().copy()
    assert first.tolist() == list(m.item_id_map.keys())
    np.random.seed(2)
    m._preprocess_arrays(df[["user_id", "item_id"]], df["rating"], type="fit")
    second = m._internal_to_raw_items()
    assert second.tolist() == list(m.item_id_map.keys())
    assert len(first) == len(second) and first.tolist() != second.tolist()  # same items, different order
    # the cache never travels in a pickle (it holds a reference to the dict it was built from)
    assert "_raw_items_cache" not in m.__getstate__()


def test_mirror_detects_in_place_edits():
    """ADVICE r1 / VERDICT weak #6: a device mirror is only served while the host array still has the content it
    was registered with (the reference reads the host arrays on every call)."""
    from matrix_factorization_b200 import _mirror

    table = {}
    a = np.random.default_rng(0).normal(size=(50, 8))
    sentinel = object()
    _mirror._register(table, a, sentinel)
    assert _mirror._lookup(table, a) is sentinel
    a[17, 3] += 1.0  # in-place edit, same id(a)
    assert _mirror._lookup(table, a) is None
    _mirror._register(table, a, sentinel)
    assert _mirror._lookup(table, a) is sentinel
    a[...] = a[::-1].copy()  # a permutation keeps sum and sum of squares but not the end elements
    assert _mirror._lookup(table, a) is None
    big = np.zeros(1 << 20)
    _mirror._register(table, big, sentinel)
    big[:: (1 << 20) // (1 << 16)] = 1.0  # large arrays: the strided sample catches bulk re-initialisation
    assert _mirror._lookup(table, big) is None


def test_topk_metrics_match_per_user_loop():
    """evaluation.topk_metrics / holdout_split against a per-user restatement of the reference's evaluate loop
    (project_template/pipeline/evaluate.py:33-111) on random recommendation lists."""
    from matrix_factorization_b200 import evaluation as ev

    rng = np.random.default_rng(3)
    df = synth_ratings(80, 60, 2500, seed=11, min_per_user=4)
    k, n_test, thr, seed = 7, 3, 4.0, 123
    train, test = ev.holdout_split(df, n_test, thr, seed)
    # reference-style split, user by user with one RandomState
    rs = np.random.RandomState(seed)
    ref_train, ref_test = {}, {}
    for u in df["user_id"].unique():
        hist = df[df["user_id"] == u]
        if hist.shape[0] <= n_test:
            continue
        pos = hist[hist["rating"] >= thr]
        t = pos.sample(n=n_test, random_state=rs) if pos.shape[0] >= n_test else hist.sort_values("rating", ascending=False).head(n_test)
        ti = t["item_id"].tolist()
        tr = hist.loc[~hist["item_id"].isin(ti), "item_id"].tolist()
        if tr and ti:
            ref_train[u], ref_test[u] = tr, ti
    assert {u: g["item_id"].tolist() for u, g in train.groupby("user_id", sort=False)} == ref_train
    assert {u: g["item_id"].tolist() for u, g in test.groupby("user_id", sort=False)} == ref_test

    class FakeModel:  # random lists that avoid the known items, like recommend_all returns them
        def contains_user(self, u):
            return True

        def recommend_all(self, users, amount, items_known):
            known = items_known.groupby("user_id")["item_id"].apply(set).to_dict()
            rows = []
            for u in users:
                cand = [i for i in rng.permutation(60) + 1 if i not in known.get(u, ())][:amount]
                rows += [(u, i, rk) for rk, i in enumerate(cand)]
            return pd.DataFrame(rows, columns=["user_id", "item_id", "rank"])

    model = FakeModel()
    rec = model.recommend_all(list(ref_train), k, train[["user_id", "item_id"]])
    res = ev.topk_metrics(rec, test, k)
    ps, rs_, ns = [], [], []
    for u in ref_train:
        items = rec[rec.user_id == u].sort_values("rank")["item_id"].tolist()
        hit = np.array([1 if i in set(ref_test[u]) else 0 for i in items])
        ps.append(hit.mean())
        rs_.append(hit.sum() / max(1, len(set(ref_test[u]))))
        gains = (2.0 ** hit - 1) / np.log2(np.arange(2, hit.size + 2))
        ideal = np.sort(hit)[::-1]
        idcg = np.sum((2.0 ** ideal - 1) / np.log2(np.arange(2, ideal.size + 2)))
        ns.append(gains.sum() / idcg if idcg > 0 else 0.0)
    assert res.n_users == len(ref_train)
    assert abs(res.precision - np.mean(ps)) < 1e-12 and abs(res.recall - np.mean(rs_)) < 1e-12 and abs(res.ndcg - np.mean(ns)) < 1e-12
