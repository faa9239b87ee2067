"""
Multi-GPU DSGD for KernelMF: one process per GPU, torch.distributed for the plumbing.

Users are dealt into G nnz-balanced stripes (rank g owns stripe g of P, the user biases and the
ratings of its users, pre-split into G item blocks); items are dealt into G stripes.  At sub-epoch
s rank g holds item stripe (g + s) mod G and runs the single-GPU stratified kernel on block
(g, (g + s) mod G); then every rank sends its item stripe (factors + biases) to rank g-1 and
receives the next one from rank g+1 (ring shift over NVLink / NVSwitch).  G sub-epochs = 1 epoch.
Blocks in flight never share a user or an item, so the step-major order of the G block plans,
concatenated sub-epoch by sub-epoch, is a valid sequential replay order.

The host-side partitioning logic (`partition`) is pure numpy and is what the gloo CPU tests
exercise; `DsgdTrainer` needs CUDA + NCCL.
"""
from __future__ import annotations

import math

import numpy as np


def deal_balanced(deg: np.ndarray, n_bins: int):
    """Snake-deal ids by descending degree into n_bins (nnz balance).  Returns (bin_of, local_index)."""
    order = np.argsort(-deg, kind="stable")
    rnd, pos = np.divmod(np.arange(len(deg)), n_bins)
    b = np.where(rnd % 2 == 1, n_bins - 1 - pos, pos)
    bin_of = np.empty(len(deg), dtype=np.int64)
    local = np.empty(len(deg), dtype=np.int64)
    bin_of[order] = b
    local[order] = rnd
    return bin_of, local


def partition(u: np.ndarray, i: np.ndarray, n_users: int, n_items: int, G: int):
    """Stripe assignment for G ranks.  Returns dict with per-user / per-item (stripe, local id), stripe
    sizes and, per rating, its block (user stripe, item stripe)."""
    du = np.bincount(u, minlength=n_users)
    di = np.bincount(i, minlength=n_items)
    us, ul = deal_balanced(du, G)
    is_, il = deal_balanced(di, G)
    return {
        "user_stripe": us, "user_local": ul, "item_stripe": is_, "item_local": il,
        "users_per_stripe": np.bincount(us, minlength=G), "items_per_stripe": np.bincount(is_, minlength=G),
        "block_u": us[u], "block_i": is_[i],
    }


def subepoch_schedule(G: int):
    """[(sub-epoch s, [(rank g, item stripe j), ...])]: the G blocks processed concurrently in each sub-epoch."""
    return [(s, [(g, (g + s) % G) for g in range(G)]) for s in range(G)]


class DsgdTrainer:
    """Rank-local state of a G-GPU DSGD fit.  Construct on every rank with the ratings of the rank's
    user stripe (local user ids, GLOBAL item stripe / local item ids).

    The item stripe a rank currently holds lives in ONE flat buffer `[max_items * ld factors | max_items biases]`, so
    that the SGD kernel works on it in place (Q and bi are views) and one NCCL send / recv moves it to the ring
    neighbour; a second buffer receives the incoming stripe."""

    def __init__(self, rank: int, world: int, u_local, item_stripe, item_local, r, n_users_local: int,
                 items_per_stripe, n_factors: int, P, Q_stripe, bu, bi_stripe, device, hot_min_degree: int = 0):
        import torch
        from . import engine

        self.rank, self.G, self.F = rank, world, n_factors
        self.device = device
        self.P, self.bu = P, bu
        self.n = int(u_local.numel())
        ld = P.shape[1]
        self.ld = ld
        self.max_items = int(max(items_per_stripe))
        M = self.max_items
        self.buf_len = (M * (ld + 1) + 3) & ~3  # (a multiple of 16 bytes: gathered stripes stay aligned)
        self.qbuf = [torch.zeros((self.buf_len,), dtype=torch.float32, device=device) for _ in range(2)]
        self.cur = 0
        self.items_per_stripe = [int(x) for x in items_per_stripe]
        q, b = self.stripe_views(0)
        q[: Q_stripe.shape[0]].copy_(Q_stripe)
        b[: Q_stripe.shape[0]].copy_(bi_stripe)
        # one plan per item stripe (block (rank, j)); hot_min_degree = 0: the block's longest item / user chains go to the
        # exact mini-batch phases (threshold chosen per block by the plan builder)
        self.plans, self.block_n, self.block_data = [], [], []
        for j in range(world):
            m = item_stripe == j
            bu_, bi_, br_ = u_local[m].contiguous(), item_local[m].contiguous(), r[m].contiguous()
            self.block_data.append((bu_, bi_, br_, torch.nonzero(m).flatten()))
            self.block_n.append(int(bu_.numel()))
            self.plans.append(engine.Plan(bu_, bi_, br_, n_users_local, max(1, self.items_per_stripe[j]), n_factors=n_factors,
                                          hot_min_degree=hot_min_degree))
        self.sse = torch.zeros((1,), dtype=torch.float64, device=device)

    def stripe_views(self, which: int):
        """(Q [max_items, ld], bi [max_items]) views of stripe buffer `which`."""
        M, ld = self.max_items, self.ld
        buf = self.qbuf[which]
        return buf[: M * ld].view(M, ld), buf[M * ld: M * ld + M]

    def block_order(self, j: int):
        """Sequential order of block (rank, j) as positions into the arrays this trainer was built from."""
        if self.block_n[j] == 0:
            return self.block_data[j][3]
        return self.block_data[j][3][self.plans[j].order()]

    def sub_epoch(self, s: int, kernel, mu, lr, reg, gamma, lo, hi):
        """Local stratified SGD on the block this rank holds in sub-epoch s (item stripe (rank + s) mod G), in place."""
        from . import engine

        j = (self.rank + s) % self.G
        if self.block_n[j] > 0:
            q, b = self.stripe_views(self.cur)
            engine.kmf_sgd_epoch(self.plans[j], kernel, self.P, q, self.bu, b, self.F, mu, lr, reg, gamma, lo, hi)
        return j

    def shift(self):
        """Ring shift: the held stripe goes to rank - 1, the next one comes from rank + 1 (one send + one recv)."""
        import torch.distributed as dist

        G, g = self.G, self.rank
        if G > 1:
            ops = [dist.P2POp(dist.isend, self.qbuf[self.cur], (g - 1) % G),
                   dist.P2POp(dist.irecv, self.qbuf[1 - self.cur], (g + 1) % G)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            self.cur = 1 - self.cur

    def epoch(self, kernel, mu, lr, reg, gamma, lo, hi):
        """One DSGD epoch: G sub-epochs of (local stratified SGD on one block, ring shift of the item stripe).
        After G shifts every stripe is back on its home rank."""
        for s in range(self.G):
            self.sub_epoch(s, kernel, mu, lr, reg, gamma, lo, hi)
            self.shift()

    def sse_epoch(self, kernel, mu, gamma, lo, hi):
        """Sum of squared errors of the rank's ratings: all-gather the item stripes, one SSE pass per block."""
        import torch
        import torch.distributed as dist
        from . import engine

        G, ld, M = self.G, self.ld, self.max_items
        mine = self.qbuf[self.cur]
        if G > 1:
            allq = torch.empty((G, self.buf_len), dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(allq.view(-1), mine)
        else:
            allq = mine.view(1, -1)
        total = torch.zeros((1,), dtype=torch.float64, device=self.device)
        for j in range(G):
            if self.block_n[j] == 0:
                continue
            bu_, bi_, br_, _ = self.block_data[j]
            engine.kmf_sse(kernel, bu_, bi_, br_, self.P, allq[j][: M * ld].view(M, ld), self.bu, allq[j][M * ld: M * ld + M], self.F, mu,
                           gamma, lo, hi, self.sse)
            total += self.sse
        if G > 1:
            dist.all_reduce(total)
        return total

    def home_stripe(self):
        """(Q stripe [n_items_of_stripe, ld], bi stripe) of this rank's home item stripe."""
        q, b = self.stripe_views(self.cur)
        nj = self.items_per_stripe[self.rank]
        return q[:nj], b[:nj]


def sharded_topk(kernel, users, P, bu, Q_local, bi_local, local_to_global, n_factors, mu, gamma, lo, hi, k, bound,
                 mask_ptr=None, mask_items_global=None, global_to_local=None, group=None, mask_local=None):
    """
    Item-sharded recommend: every rank scores ALL requested users against ITS item
    stripe (local top-k with the known-item mask restricted to the stripe), the per-rank lists are
    all-gathered and merged (score desc, ties by lower global item id), then clipped.
    users int32 [m] (global user ids, P / bu replicated); Q_local [n_local, ld]; local_to_global int32
    [n_local]; mask in CSR form over GLOBAL item ids with global_to_local int32 [n_items] (-1 = not
    on this rank); or mask_local = (ptr int64 [m + 1], local item ids int32, ascending inside a row) when the caller
    keeps the stripe's part of the known-item lists (a rank of the DSGD grid does).  Returns (scores [m, k], items
    [m, k] global ids) on every rank.
    """
    import torch
    import torch.distributed as dist
    from . import engine

    m = int(users.numel())
    n_local = int(Q_local.shape[0])
    k_loc = max(1, min(k, n_local))
    mp = mi = None
    if mask_local is not None:
        mp, mi = mask_local
    elif mask_ptr is not None:
        loc = global_to_local[mask_items_global.long()]
        keep = loc >= 0
        # rows keep their CSR structure: count kept entries per row
        row = torch.repeat_interleave(torch.arange(m, device=users.device), (mask_ptr[1:] - mask_ptr[:-1]))
        # the scoring kernel wants item ids ascending inside each row: sort by (row, local id)
        rk, lk = row[keep], loc[keep]
        order = torch.argsort(rk * (n_local + 1) + lk)
        mi = lk[order].int().contiguous()
        if mi.numel() == 0:
            mi = torch.zeros((1,), dtype=torch.int32, device=users.device)
        mp = torch.zeros(m + 1, dtype=torch.int64, device=users.device)
        mp[1:] = torch.cumsum(torch.bincount(rk, minlength=m), 0)
    sc, it = engine.score_topk(kernel, users, P, Q_local, bu, bi_local, n_local, n_factors, mu, gamma, lo, hi, k_loc,
                               False, mp, mi)
    valid = it >= 0
    git = torch.where(valid, local_to_global[it.clamp(min=0).long()].int(), torch.full_like(it, -1))
    if k_loc < k:  # pad to a common width
        pad = k - k_loc
        sc = torch.cat([sc, torch.full((m, pad), float("-inf"), device=sc.device)], 1)
        git = torch.cat([git, torch.full((m, pad), -1, dtype=torch.int32, device=sc.device)], 1)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        all_sc = [torch.empty_like(sc) for _ in range(world)]
        all_it = [torch.empty_like(git) for _ in range(world)]
        dist.all_gather(all_sc, sc.contiguous(), group=group)
        dist.all_gather(all_it, git.contiguous(), group=group)
        sc, git = torch.cat(all_sc, 1).contiguous(), torch.cat(all_it, 1).contiguous()
    return engine.topk_merge(sc, git, k, bound, lo, hi)
