"""CPU (gloo, world_size 2): the host-side DSGD logic -- stripe partition, sub-epoch schedule and the
ring shift of item stripes -- reproduces a sequential replay of the same ratings in block order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from matrix_factorization_b200.dist import deal_balanced, partition, subepoch_schedule


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(seed=3, U=40, I=30, N=500, F=6):
    rng = np.random.default_rng(seed)
    keys = rng.choice(U * I, N, replace=False)
    u, i = keys // I, keys % I
    r = rng.integers(1, 6, N).astype(np.float64)
    return u, i, r, rng.normal(0, 0.1, (U, F)), rng.normal(0, 0.1, (I, F)), U, I, F


def _worker(rank, world, port, out):
    from oracle import oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    u, i, r, P0, Q0, U, I, F = _problem()
    part = partition(u, i, U, I, world)
    my_users = np.nonzero(part["user_stripe"] == rank)[0]
    my_items = np.nonzero(part["item_stripe"] == rank)[0]
    mu, lr, reg = float(r.mean()), 0.02, 0.01
    P = P0.copy()                       # only rows of my_users are meaningful on this rank
    bu = np.zeros(U)
    # stripe buffer: full-size Q / bi arrays, only the rows of the stripe currently held are valid
    Q, bi = Q0.copy(), np.zeros(I)
    held = rank
    for s, blocks in subepoch_schedule(world):
        j = dict(blocks)[rank]
        assert j == held
        m = (part["block_u"] == rank) & (part["block_i"] == j)
        idx = np.nonzero(m)[0]
        P, Q, bu, bi = orc.kmf_replay("linear", u, i, r, idx, mu, bu, bi, P, Q, lr, reg)
        # ring shift: send the held stripe's rows to rank-1, receive the next stripe from rank+1
        items_held = np.nonzero(part["item_stripe"] == held)[0]
        nxt = (held + 1) % world
        items_next = np.nonzero(part["item_stripe"] == nxt)[0]
        send = torch.from_numpy(np.concatenate([Q[items_held], bi[items_held, None]], axis=1).copy())
        recv = torch.zeros((len(items_next), F + 1), dtype=torch.float64)
        if rank % 2 == 0:
            dist.send(send, (rank - 1) % world)
            dist.recv(recv, (rank + 1) % world)
        else:
            dist.recv(recv, (rank + 1) % world)
            dist.send(send, (rank - 1) % world)
        Q[items_next] = recv[:, :F].numpy()
        bi[items_next] = recv[:, F].numpy()
        held = nxt
    assert held == rank  # after G shifts every stripe is home again
    res = {"rank": rank, "users": my_users, "items": my_items, "P": P[my_users], "bu": bu[my_users],
           "Q": Q[my_items], "bi": bi[my_items]}
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        torch.save(gathered, out)
    dist.destroy_process_group()


def test_dsgd_two_ranks_equals_sequential_replay(tmp_path):
    from oracle import oracle as orc

    world, out = 2, str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    gathered = torch.load(out, weights_only=False)
    u, i, r, P0, Q0, U, I, F = _problem()
    part = partition(u, i, U, I, world)
    order = []
    for s, blocks in subepoch_schedule(world):
        for g, j in blocks:
            order.append(np.nonzero((part["block_u"] == g) & (part["block_i"] == j))[0])
        # blocks of one sub-epoch never share a user or an item
        us = [set(u[o]) for o in order[-world:]]
        its = [set(i[o]) for o in order[-world:]]
        assert not (us[0] & us[1]) and not (its[0] & its[1])
    order = np.concatenate(order)
    assert np.array_equal(np.sort(order), np.arange(len(u)))
    P, Q, bu, bi = orc.kmf_replay("linear", u, i, r, order, float(r.mean()), np.zeros(U), np.zeros(I), P0, Q0, 0.02, 0.01)
    for res in gathered:
        np.testing.assert_allclose(res["P"], P[res["users"]], atol=1e-13)
        np.testing.assert_allclose(res["bu"], bu[res["users"]], atol=1e-13)
        np.testing.assert_allclose(res["Q"], Q[res["items"]], atol=1e-13)
        np.testing.assert_allclose(res["bi"], bi[res["items"]], atol=1e-13)


def test_partition_is_balanced_and_consistent():
    from matrix_factorization_b200.data import synth_pairs

    u, i = synth_pairs(500, 300, 20000, seed=5)
    for G in (2, 4, 8):
        part = partition(u, i, 500, 300, G)
        cnt_u = np.bincount(part["block_u"], minlength=G)
        cnt_i = np.bincount(part["block_i"], minlength=G)
        assert cnt_u.max() / cnt_u.mean() < 1.1 and cnt_i.max() / cnt_i.mean() < 1.25
        # local ids are a bijection inside every stripe
        for s in range(G):
            loc = part["user_local"][part["user_stripe"] == s]
            assert np.array_equal(np.sort(loc), np.arange(len(loc)))
        sched = subepoch_schedule(G)
        seen = {(g, j) for _, blocks in sched for g, j in blocks}
        assert len(seen) == G * G
        for _, blocks in sched:
            assert len({j for _, j in blocks}) == G  # every item stripe is held by exactly one rank
    b, l = deal_balanced(np.array([5, 1, 9, 3]), 2)
    assert b.tolist() == [1, 0, 0, 1] and l.tolist() == [0, 1, 0, 1]
