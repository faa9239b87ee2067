"""GPU tests of the scoring / top-k path (recommend): kernel vs numpy fp64, all three kernels, masks,
item-sharded top-k + merge (single-GPU emulation of the all-gather)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref_scores(kernel, P, Q, bu, bi, mu, gamma, lo, hi):
    if kernel == "linear":
        return mu + bu[:, None] + bi[None, :] + P @ Q.T
    if kernel == "sigmoid":
        return lo + (hi - lo) / (1.0 + np.exp(-(mu + bu[:, None] + bi[None, :] + P @ Q.T)))
    d2 = ((P[:, None, :] - Q[None, :, :]) ** 2).sum(-1)
    return lo + (hi - lo) * np.exp(-gamma * d2)


def _check_lists(scores, items, ref, k, masked_sets, tol=2e-5):
    for row in range(ref.shape[0]):
        s = ref[row].copy()
        s[list(masked_sets[row])] = -np.inf
        order = np.argsort(-s, kind="stable")[:k]
        n_valid = int(np.isfinite(s).sum())
        kk = min(k, n_valid)
        np.testing.assert_allclose(scores[row, :kk], s[order][:kk], atol=tol, rtol=0)
        got, exp = items[row, :kk], order[:kk]
        for a, b, pos in zip(got, exp, range(kk)):
            if a != b:  # only inside a tie within tol
                assert abs(s[a] - s[b]) <= 2 * tol, (row, pos, a, b, s[a], s[b])
        assert not (set(got.tolist()) & masked_sets[row])
        assert np.all(items[row, kk:] == -1)


@pytest.mark.parametrize("kernel", ["linear", "sigmoid", "rbf"])
@pytest.mark.parametrize("F,U,I,k", [(16, 70, 333, 10), (100, 40, 1000, 50), (128, 130, 257, 7)])
def test_score_topk_matches_numpy(kernel, F, U, I, k):
    import torch
    from matrix_factorization_b200 import engine

    rng = np.random.default_rng(F + I)
    P, Q = rng.normal(0, 0.3, (U, F)), rng.normal(0, 0.3, (I, F))
    bu, bi = rng.normal(0, 0.2, U), rng.normal(0, 0.2, I)
    mu, gamma, lo, hi = 3.2, 0.05, 0.0, 5.0
    users = rng.permutation(U)[: U - 3].astype(np.int32)
    masked = [set(rng.choice(I, rng.integers(0, 40), replace=False).tolist()) for _ in users]
    masked[0] = set(range(I - 3))  # fewer than k candidates left
    mask_ptr = np.zeros(len(users) + 1, dtype=np.int64)
    mask_ptr[1:] = np.cumsum([len(m) for m in masked])
    mask_items = np.concatenate([np.array(sorted(m), dtype=np.int32) for m in masked])
    dP, dQ = engine.upload_rows(P), engine.upload_rows(Q)
    dbu, dbi = engine.upload_vec(bu), engine.upload_vec(bi)
    sc, it = engine.score_topk(kernel, torch.tensor(users).cuda(), dP, dQ, dbu, dbi, I, F, mu, gamma, lo, hi, k, False,
                               torch.tensor(mask_ptr).cuda(), torch.tensor(mask_items).cuda())
    ref = _ref_scores(kernel, P[users], Q, bu[users], bi, mu, gamma, lo, hi)
    _check_lists(sc.cpu().numpy().astype(np.float64), it.cpu().numpy(), ref, k, masked)
    # bounded variant clips after selection
    sc2, it2 = engine.score_topk(kernel, torch.tensor(users).cuda(), dP, dQ, dbu, dbi, I, F, mu, gamma, 1.0, 4.0, k, True,
                                 torch.tensor(mask_ptr).cuda(), torch.tensor(mask_items).cuda())
    s2 = sc2.cpu().numpy()
    assert np.all((s2[np.isfinite(s2)] >= 1.0) & (s2[np.isfinite(s2)] <= 4.0))


@pytest.mark.parametrize("G", [2, 4])
def test_item_sharded_topk_merge_equals_unsharded(G):
    """What the G-rank recommend does (local top-k per item stripe, all-gather, merge), emulated on one GPU."""
    import torch
    from matrix_factorization_b200 import engine
    from matrix_factorization_b200.dist import deal_balanced

    rng = np.random.default_rng(G)
    U, I, F, k = 50, 400, 32, 20
    P, Q = rng.normal(0, 0.3, (U, F)), rng.normal(0, 0.3, (I, F))
    bu, bi = rng.normal(0, 0.2, U), rng.normal(0, 0.2, I)
    users = torch.arange(U, dtype=torch.int32).cuda()
    masked = [set(rng.choice(I, 15, replace=False).tolist()) for _ in range(U)]
    mask_ptr = torch.tensor(np.arange(U + 1) * 15, dtype=torch.int64).cuda()
    mask_items = torch.tensor(np.concatenate([sorted(m) for m in masked]), dtype=torch.int32).cuda()
    dP, dbu = engine.upload_rows(P), engine.upload_vec(bu)
    stripe, _ = deal_balanced(rng.integers(1, 50, I), G)
    cand_s, cand_i = [], []
    for g in range(G):
        gl = np.nonzero(stripe == g)[0]
        g2l = np.full(I, -1, dtype=np.int64)
        g2l[gl] = np.arange(len(gl))
        loc = g2l[mask_items.cpu().numpy()]
        keep = loc >= 0
        rows = np.repeat(np.arange(U), 15)
        mp = np.zeros(U + 1, dtype=np.int64)
        mp[1:] = np.cumsum(np.bincount(rows[keep], minlength=U))
        sc, it = engine.score_topk("linear", users, dP, engine.upload_rows(Q[gl]), dbu, engine.upload_vec(bi[gl]), len(gl), F,
                                   3.0, 0.1, 0.0, 5.0, k, False, torch.tensor(mp).cuda(),
                                   torch.tensor(loc[keep].astype(np.int32)).cuda())
        it = it.cpu().numpy()
        cand_s.append(sc)
        cand_i.append(torch.tensor(np.where(it >= 0, gl[np.clip(it, 0, None)], -1).astype(np.int32)).cuda())
    ms, mi = engine.topk_merge(torch.cat(cand_s, 1).contiguous(), torch.cat(cand_i, 1).contiguous(), k, True, 0.0, 5.0)
    full_s, full_i = engine.score_topk("linear", users, dP, engine.upload_rows(Q), dbu, engine.upload_vec(bi), I, F, 3.0,
                                       0.1, 0.0, 5.0, k, True, mask_ptr, mask_items)
    assert torch.equal(mi, full_i)
    assert torch.allclose(ms, full_s, atol=1e-6)


def test_tensor_path_agrees_with_simt_path_on_a_large_shape():
    """tcgen05 split-TF32 path vs the fp32 SIMT path (forced with MFK_SCORE_SIMT=1 in a subprocess): several user
    tiles, several k-blocks (F=256), ragged item tile, k=50, per-user masks."""
    import os
    import subprocess
    import sys
    import tempfile

    code = """
import sys, numpy as np, torch
sys.path.insert(0, %r)
from matrix_factorization_b200 import engine
rng = np.random.default_rng(11)
U, I, F, k = 700, 3000, 256, 50
P, Q = rng.normal(0, 0.2, (U, F)), rng.normal(0, 0.2, (I, F))
bu, bi = rng.normal(0, 0.2, U), rng.normal(0, 0.2, I)
users = torch.tensor(rng.permutation(U)[:650].astype(np.int32)).cuda()
lens = rng.integers(0, 120, 650)
mp = np.zeros(651, dtype=np.int64); mp[1:] = np.cumsum(lens)
mi = np.concatenate([np.sort(rng.choice(I, n, replace=False)) for n in lens]).astype(np.int32)
out = {}
for kern in ("linear", "rbf"):
    sc, it = engine.score_topk(kern, users, engine.upload_rows(P), engine.upload_rows(Q), engine.upload_vec(bu),
                               engine.upload_vec(bi), I, F, 3.0, 0.02, 0.0, 5.0, k, False,
                               torch.tensor(mp).cuda(), torch.tensor(mi).cuda())
    out[kern + "_s"] = sc.cpu().numpy(); out[kern + "_i"] = it.cpu().numpy()
np.savez(sys.argv[1], **out)
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for tag, env in (("tc", {}), ("simt", {"MFK_SCORE_SIMT": "1"})):
            f = os.path.join(d, tag + ".npz")
            subprocess.run([sys.executable, "-c", code, f], check=True, env={**os.environ, **env}, timeout=300)
            res[tag] = dict(np.load(f))
    for kern in ("linear", "rbf"):
        a_s, b_s = res["tc"][kern + "_s"], res["simt"][kern + "_s"]
        a_i, b_i = res["tc"][kern + "_i"], res["simt"][kern + "_i"]
        np.testing.assert_allclose(a_s, b_s, atol=2e-5, rtol=0)
        diff = a_i != b_i
        # item ids may differ only inside ties within 1e-5 (fp32 summation order differs between the two paths)
        assert diff.mean() < 0.01
        rows, cols = np.nonzero(diff)
        for r, c in zip(rows, cols):
            if c == a_i.shape[1] - 1:
                continue  # the k-th place may be tied with the (k+1)-th item, which is outside both lists
            assert np.sum(np.abs(b_s[r] - b_s[r, c]) < 2e-5) > 1, (kern, r, c)
