"""bench.py's N > 1 leg: DSGD over the GPUs of one node (one process per GPU, NCCL)."""
from __future__ import annotations

import json
import math
import os
import time

import numpy as np


def _partition_torch(u, i, U, I, G):
    """GPU version of dist.partition (same snake dealing by descending degree)."""
    import torch

    def deal(ids, n):
        deg = torch.bincount(ids.long(), minlength=n)
        order = torch.argsort(deg, descending=True, stable=True)
        k = torch.arange(n, device=ids.device)
        rnd, pos = k // G, k % G
        b = torch.where(rnd % 2 == 1, G - 1 - pos, pos)
        bin_of = torch.empty(n, dtype=torch.int64, device=ids.device)
        local = torch.empty(n, dtype=torch.int64, device=ids.device)
        bin_of[order] = b
        local[order] = rnd
        return bin_of, local

    us, ul = deal(u, U)
    is_, il = deal(i, I)
    return us, ul, is_, il


def run(args):
    import torch
    import torch.distributed as dist

    import bench
    from . import engine
    from .dist import DsgdTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    G = world

    wl = bench.gen_workload(args.workload, dev, uniform=args.uniform)  # same seed => same data on every rank
    F, U, I, N = wl["F"], wl["U"], wl["I"], wl["N"]
    us, ul, is_, il = _partition_torch(wl["u"], wl["i"], U, I, G)
    mine = us[wl["u"].long()] == rank
    u_loc = ul[wl["u"][mine].long()].int()
    it = wl["i"][mine].long()
    i_stripe, i_loc = is_[it], il[it].int()
    r_loc = wl["r"][mine]
    n_users_local = int((us == rank).sum().item())
    items_per_stripe = torch.bincount(is_, minlength=G).cpu().tolist()
    mu = float(wl["r"].double().mean().item())
    n_local = int(u_loc.numel())
    # item-sharded recommend: the known-item lists of ALL users restricted to this rank's item stripe (local item ids,
    # CSR built on the GPU), and every user's row in the gathered user matrix (stripe-major, stripes padded alike)
    in_stripe = is_[wl["i"].long()] == rank
    csr = engine.Csr(wl["u"][in_stripe], il[wl["i"][in_stripe].long()].int(), wl["r"][in_stripe], U, items_per_stripe[rank])
    rec_ptr, rec_col, *_ = csr.export()
    rec_ptr, rec_col = rec_ptr.clone(), rec_col.clone()
    csr.close()
    max_users = int(torch.bincount(us, minlength=G).max().item())
    rec_rows = (us * max_users + ul).int().contiguous()
    stripe_items = torch.nonzero(is_ == rank).flatten()
    local_to_global = torch.empty(items_per_stripe[rank], dtype=torch.int32, device=dev)
    local_to_global[il[stripe_items]] = stripe_items.int()
    del wl["u"], wl["i"], wl["r"], mine, it, in_stripe
    torch.cuda.empty_cache()

    ld = engine.round_up4(F)

    def fresh_params():
        g = torch.Generator(device=dev).manual_seed(5 + rank)
        P = torch.zeros(n_users_local, ld, device=dev)
        P[:, :F] = torch.randn(n_users_local, F, device=dev, generator=g) * 0.1
        Q = torch.zeros(items_per_stripe[rank], ld, device=dev)
        Q[:, :F] = torch.randn(items_per_stripe[rank], F, device=dev, generator=g) * 0.1
        return P, Q, torch.zeros(n_users_local, device=dev), torch.zeros(items_per_stripe[rank], device=dev)

    P, Q, bu, bi = fresh_params()
    t0 = time.perf_counter()
    tr = DsgdTrainer(rank, G, u_loc, i_stripe, i_loc, r_loc, n_users_local, items_per_stripe, F, P, Q, bu, bi, dev)
    torch.cuda.synchronize()
    plan_ms = 1e3 * (time.perf_counter() - t0)

    lr, reg, gamma = wl["lr"], wl["reg"], 1.0 / F
    sse_hist = []

    def step():
        tr.epoch("linear", mu, lr, reg, gamma, 0.0, 5.0)
        sse_hist.append(tr.sse_epoch("linear", mu, gamma, 0.0, 5.0))

    for _ in range(args.warmup):
        step()
    clocks = bench.ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.synchronize()
    dist.barrier()
    if clocks:
        clocks.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clk = clocks.stop() if clocks else None

    # kernel-only time of the local SGD launches (for the roofline line), one more epoch
    ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = 0.0
    for j in range(G):
        if tr.block_n[j] == 0:
            continue
        nj = tr.items_per_stripe[j]
        ka.record()
        qv, bv = tr.stripe_views(tr.cur)
        engine.kmf_sgd_epoch(tr.plans[j], "linear", tr.P, qv, tr.bu, bv, F, mu, lr, reg, gamma, 0.0, 5.0)
        kb.record()
        torch.cuda.synchronize()
        kms += ka.elapsed_time(kb)

    # end-to-end: pinned host shards -> device -> plans -> n_epochs epochs (+RMSE) -> parameters back on the host
    hu, hs, hi, hr = (t.cpu().pin_memory() for t in (u_loc, i_stripe.int(), i_loc, r_loc))
    P2, Q2, bu2, bi2 = fresh_params()
    hP, hQ, hbu, hbi = (t.cpu().pin_memory() for t in (P2, Q2, bu2, bi2))
    del tr
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    du, ds, di_, dr = (t.to(dev, non_blocking=True) for t in (hu, hs, hi, hr))
    dP, dQ, dbu, dbi = (t.to(dev, non_blocking=True) for t in (hP, hQ, hbu, hbi))
    tr2 = DsgdTrainer(rank, G, du, ds.long(), di_, dr, n_users_local, items_per_stripe, F, dP, dQ, dbu, dbi, dev)
    e2e_sse = None
    for _ in range(wl["n_epochs"]):
        tr2.epoch("linear", mu, lr, reg, gamma, 0.0, 5.0)
        e2e_sse = tr2.sse_epoch("linear", mu, gamma, 0.0, 5.0)
    qs, bis = tr2.home_stripe()
    hP.copy_(dP, non_blocking=True)
    hbu.copy_(dbu, non_blocking=True)
    hQ.copy_(qs, non_blocking=True)
    hbi.copy_(bis, non_blocking=True)
    e2e_rmse = math.sqrt(float(e2e_sse.item()) / N)
    torch.cuda.synchronize()
    dist.barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())

    # ---- recommend, item-sharded: all-gather of the user stripes, every rank scores all users against its item
    #      stripe (known items of the stripe masked), all-gather of the per-rank lists, merge
    from .dist import sharded_topk

    k_rec = 50
    Ps = torch.zeros(max_users, ld, device=dev)  # this rank's users, padded to the longest stripe
    Ps[:n_users_local] = dP
    bs = torch.zeros(max_users, device=dev)
    bs[:n_users_local] = dbu
    Pall = torch.empty(G * max_users, ld, device=dev)
    ball = torch.empty(G * max_users, device=dev)
    rec_ms = []
    rec_out = None
    for rep_ in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        ra, rb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ra.record()
        dist.all_gather_into_tensor(Pall.view(-1), Ps.view(-1))
        dist.all_gather_into_tensor(ball, bs)
        rec_out = sharded_topk("linear", rec_rows, Pall, ball, qs, bis, local_to_global, F, mu, gamma, 0.0, 5.0, k_rec, True,
                               mask_local=(rec_ptr, rec_col))
        rb.record()
        torch.cuda.synchronize()
        t = torch.tensor([ra.elapsed_time(rb)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep_ > 0:
            rec_ms.append(float(t.item()))
    rec = {"metric": "recommend users/s", "value": U / (min(rec_ms) * 1e-3), "unit": "users/s", "users": U, "k": k_rec,
           "ms": min(rec_ms), "mask": "each user's training items",
           "path": f"item-sharded over {G} GPUs: all-gather of the user stripes, local tcgen05 top-{k_rec} per item stripe, "
                   "all-gather of the lists, mfk_topk_merge"}

    if rank == 0:
        peak, peak_src = bench.load_peaks()
        bpu = 16 * F + 28
        achieved = bpu * n_local / (kms * 1e-3) / 1e9
        rmse = [math.sqrt(float(x.item()) / N) for x in sse_hist]
        line = {
            "metric": "KernelMF SGD rating-updates/s", "value": N * args.steps / (total_ms * 1e-3),
            "unit": "rating-updates/s", "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": bench.workload_name(args), "n_factors": F, "kernel": "linear", "lr": lr, "reg": reg,
                       "parallelism": f"dsgd{G}: {G}x{G} user/item block grid, item stripes ring-shifted with NCCL send/recv",
                       "step": "1 epoch = G sub-epochs (local stratified SGD + ring shift) + all-gather RMSE pass",
                       "plan_build_ms": plan_ms, "l2": "per-rank working set exceeds L2, no explicit flush",
                       "train_rmse_first_last": [rmse[0], rmse[-1]]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "rank 0: the SGD launches (k_sgd_batch hot items / hot users, k_sgd_flat) of its G blocks of one epoch",
                         "kernel_ms": kms, "bytes_per_update": bpu, "peak_source": peak_src},
            "cpu_baseline": None,
            "e2e": {"value": N * wl["n_epochs"] / e2e_s, "unit": "rating-updates/s",
                    "h2d_bytes_per_step": int(n_local * 16 + (n_users_local + items_per_stripe[0]) * (ld + 1) * 4),
                    "d2h_bytes_per_step": int((n_users_local + items_per_stripe[0]) * (ld + 1) * 4),
                    "call": f"per rank: pinned host shard -> H2D -> {G} block plans -> {wl['n_epochs']} DSGD epochs + RMSE -> D2H",
                    "seconds_per_call": e2e_s, "train_rmse_last": e2e_rmse},
            "recommend": rec,
            "gpu_launches": int(args.steps * (G + 2 * G)),
            "clocks": clk,
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
