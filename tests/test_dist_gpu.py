"""GPU parity tests of the multi-GPU pieces.

* `DsgdTrainer` (the class bench.py --gpus N times): G trainers emulated in ONE process on one GPU -- every rank's
  sub-epoch runs through the real trainer (block plans, in-place stripe buffers), the ring shift is a buffer copy -- and
  the result is compared with the fp64 oracle replaying the concatenated block orders.
* the same over NCCL with one process per GPU, and the item-sharded recommend with its all-gather merge, when the box
  has >= 2 GPUs (skipped otherwise).
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-12, np.max(np.abs(b))))


def _problem(seed=11, U=900, I=500, N=60_000, F=64, hot=0.35):
    rng = np.random.default_rng(seed)
    keys = rng.choice(U * I, N, replace=False)
    u, i = (keys // I).astype(np.int64), (keys % I).astype(np.int64)
    m = rng.random(N) < hot  # a few heavily rated items and very active users: every phase of the block plans runs
    i[m] = rng.integers(0, 6, m.sum())
    m2 = (~m) & (rng.random(N) < 0.15)
    u[m2] = rng.integers(0, 5, m2.sum())
    keep = np.unique(u * I + i, return_index=True)[1]
    rng.shuffle(keep)
    u, i = u[keep], i[keep]
    r = rng.integers(1, 6, len(u)).astype(np.float64)
    P0, Q0 = rng.normal(0, 0.1, (U, F)), rng.normal(0, 0.1, (I, F))
    return u, i, r, P0, Q0, U, I, F


def _make_trainer(rank, G, prob, part, dev, hot_min_degree):
    import torch
    from matrix_factorization_b200.dist import DsgdTrainer

    u, i, r, P0, Q0, U, I, F = prob
    mine = np.nonzero(part["block_u"] == rank)[0]
    users = np.nonzero(part["user_stripe"] == rank)[0]
    users = users[np.argsort(part["user_local"][users])]
    items = np.nonzero(part["item_stripe"] == rank)[0]
    items = items[np.argsort(part["item_local"][items])]
    t = lambda a, dt: torch.tensor(a, dtype=dt, device=dev)
    tr = DsgdTrainer(rank, G, t(part["user_local"][u[mine]], torch.int32), t(part["item_stripe"][i[mine]], torch.int64),
                     t(part["item_local"][i[mine]], torch.int32), t(r[mine], torch.float32), len(users),
                     part["items_per_stripe"].tolist(), F, t(P0[users], torch.float32).contiguous(),
                     t(Q0[items], torch.float32).contiguous(), torch.zeros(len(users), device=dev),
                     torch.zeros(len(items), device=dev), dev, hot_min_degree=hot_min_degree)
    return tr, mine, users, items


def _replay(prob, part, G, orders, epochs, mu, lr, reg):
    from oracle import oracle as orc

    u, i, r, P0, Q0, U, I, F = prob
    P, Q, bu, bi = P0.astype(np.float32).astype(np.float64), Q0.astype(np.float32).astype(np.float64), np.zeros(U), np.zeros(I)
    for _ in range(epochs):
        for s in range(G):
            for g in range(G):
                P, Q, bu, bi = orc.kmf_replay("linear", u, i, r, orders[(g, (g + s) % G)], mu, bu, bi, P, Q, lr, reg)
    return P, Q, bu, bi


@pytest.mark.parametrize("G", [2, 3])
def test_dsgd_trainer_emulated_ranks_match_oracle_replay(G):
    import torch
    from matrix_factorization_b200.dist import partition

    dev = torch.device("cuda", 0)
    prob = _problem()
    u, i, r, P0, Q0, U, I, F = prob
    part = partition(u, i, U, I, G)
    trs = [_make_trainer(g, G, prob, part, dev, 150) for g in range(G)]
    assert any(tr.plans[j].info()["n_hot_items"] > 0 for tr, *_ in trs for j in range(G))
    mu, lr, reg, epochs = float(r.mean()), 0.01, 0.02, 2
    orders = {}
    for g, (tr, mine, _, _) in enumerate(trs):
        for j in range(G):
            orders[(g, j)] = mine[tr.block_order(j).cpu().numpy()]
    assert np.array_equal(np.sort(np.concatenate(list(orders.values()))), np.arange(len(u)))
    for _ in range(epochs):
        for s in range(G):
            for tr, *_ in trs:
                assert tr.sub_epoch(s, "linear", mu, lr, reg, 0.01, 0.0, 5.0) == (tr.rank + s) % G
            # ring shift by hand: rank g's stripe goes to rank g - 1
            held = [tr.qbuf[tr.cur].clone() for tr, *_ in trs]
            for g, (tr, *_) in enumerate(trs):
                tr.qbuf[tr.cur].copy_(held[(g + 1) % G])
    Po, Qo, buo, bio = _replay(prob, part, G, orders, epochs, mu, lr, reg)
    sse = 0.0
    for g, (tr, mine, users, items) in enumerate(trs):
        q, b = tr.home_stripe()
        assert _rel(tr.P.cpu().numpy()[:, :F].astype(np.float64), Po[users]) < 1e-4
        assert _rel(q.cpu().numpy()[:, :F].astype(np.float64), Qo[items]) < 1e-4
        assert np.max(np.abs(tr.bu.cpu().numpy() - buo[users])) < 5e-5 and np.max(np.abs(b.cpu().numpy() - bio[items])) < 5e-5
    torch.cuda.synchronize()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from matrix_factorization_b200 import engine
    from matrix_factorization_b200.dist import partition, sharded_topk

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    prob = _problem()
    u, i, r, P0, Q0, U, I, F = prob
    part = partition(u, i, U, I, world)
    tr, mine, users, items = _make_trainer(rank, world, prob, part, dev, 150)
    mu, lr, reg, epochs = float(r.mean()), 0.01, 0.02, 2
    orders = {(rank, j): mine[tr.block_order(j).cpu().numpy()] for j in range(world)}
    sse = None
    for _ in range(epochs):
        tr.epoch("linear", mu, lr, reg, 0.01, 0.0, 5.0)
        sse = tr.sse_epoch("linear", mu, 0.01, 0.0, 5.0)
    q, b = tr.home_stripe()
    res = {"rank": rank, "users": users, "items": items, "orders": orders, "P": tr.P.cpu().numpy()[:, :F], "bu": tr.bu.cpu().numpy(),
           "Q": q.cpu().numpy()[:, :F], "bi": b.cpu().numpy(), "sse": float(sse.item())}
    # ---- item-sharded recommend: every rank scores all users against its item stripe, all-gather + merge
    k = 10
    Pfull = torch.tensor(P0, dtype=torch.float32, device=dev)
    Qfull = torch.tensor(Q0, dtype=torch.float32, device=dev)
    bu_f, bi_f = torch.zeros(U, device=dev), torch.linspace(-0.2, 0.2, I, device=dev)
    l2g = torch.tensor(items, dtype=torch.int32, device=dev)
    g2l = torch.full((I,), -1, dtype=torch.int32, device=dev)
    g2l[l2g.long()] = torch.arange(len(items), dtype=torch.int32, device=dev)
    req = torch.arange(0, U, 3, dtype=torch.int32, device=dev)
    # known-item mask: the user's rated items (CSR over the requested users, global ids ascending)
    rows = [np.sort(i[u == int(x)]) for x in req.cpu().numpy()]
    mp_ = torch.tensor(np.concatenate([[0], np.cumsum([len(x) for x in rows])]), dtype=torch.int64, device=dev)
    mi_ = torch.tensor(np.concatenate(rows) if len(rows) else np.zeros(0), dtype=torch.int32, device=dev)
    sc, it = sharded_topk("linear", req, Pfull, bu_f, Qfull[l2g.long()].contiguous(), bi_f[l2g.long()].contiguous(), l2g, F, mu, 0.01,
                          0.0, 5.0, k, True, mask_ptr=mp_, mask_items_global=mi_, global_to_local=g2l)
    sc1, it1 = engine.score_topk("linear", req, Pfull, Qfull, bu_f, bi_f, I, F, mu, 0.01, 0.0, 5.0, k, True, mp_, mi_)
    res["topk_items_equal"] = bool(torch.equal(it.cpu(), it1.cpu()))
    res["topk_score_err"] = float((sc - sc1).abs().max().item())
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        torch.save(gathered, out)
    dist.destroy_process_group()


def test_dsgd_trainer_and_sharded_recommend_over_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp
    from matrix_factorization_b200.dist import partition
    from oracle import oracle as orc

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world, out = 2, str(tmp_path / "res.pt")
    mp.spawn(_nccl_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    gathered = torch.load(out, weights_only=False)
    prob = _problem()
    u, i, r, P0, Q0, U, I, F = prob
    part = partition(u, i, U, I, world)
    orders = {}
    for res in gathered:
        orders.update(res["orders"])
    mu = float(r.mean())
    Po, Qo, buo, bio = _replay(prob, part, world, orders, 2, mu, 0.01, 0.02)
    sse_o = orc.kmf_rmse("linear", u, i, r, mu, buo, bio, Po, Qo) ** 2 * len(u)
    for res in gathered:
        assert _rel(res["P"].astype(np.float64), Po[res["users"]]) < 1e-4 and _rel(res["Q"].astype(np.float64), Qo[res["items"]]) < 1e-4
        assert np.max(np.abs(res["bu"] - buo[res["users"]])) < 5e-5 and np.max(np.abs(res["bi"] - bio[res["items"]])) < 5e-5
        assert abs(res["sse"] - sse_o) / sse_o < 1e-4
        assert res["topk_items_equal"] and res["topk_score_err"] < 1e-5
