"""Diagnostics: the local work of ONE rank of a G-GPU DSGD epoch, run on one GPU (no NCCL): build the G x G partition,
take rank `--rank`'s G blocks, time each block's SGD launch (and its phases).
    python tools/dsgd_blocks.py --workload ml-20m --G 8 [--hot-min-degree D] [--uniform]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from matrix_factorization_b200 import engine
    from matrix_factorization_b200.dist_bench import _partition_torch

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ml-20m")
    ap.add_argument("--uniform", action="store_true")
    ap.add_argument("--G", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--hot-min-degree", type=int, default=0)
    ap.add_argument("--phases", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = bench.gen_workload(args.workload, dev, uniform=args.uniform)
    F, U, I, N = wl["F"], wl["U"], wl["I"], wl["N"]
    G, rank = args.G, args.rank
    us, ul, is_, il = _partition_torch(wl["u"], wl["i"], U, I, G)
    mine = us[wl["u"].long()] == rank
    u_loc = ul[wl["u"][mine].long()].int()
    it = wl["i"][mine].long()
    i_stripe, i_loc = is_[it], il[it].int()
    r_loc = wl["r"][mine]
    n_users_local = int((us == rank).sum().item())
    items_per_stripe = torch.bincount(is_, minlength=G).cpu().tolist()
    mu = float(wl["r"].double().mean().item())
    ld = engine.round_up4(F)
    g = torch.Generator(device=dev).manual_seed(5)
    P = torch.randn(n_users_local, ld, device=dev, generator=g) * 0.1
    bu = torch.zeros(n_users_local, device=dev)
    tot = 0.0
    for j in range(G):
        m = i_stripe == j
        bu_, bi_, br_ = u_loc[m].contiguous(), i_loc[m].contiguous(), r_loc[m].contiguous()
        nj = max(1, items_per_stripe[j])
        Q = torch.randn(nj, ld, device=dev, generator=g) * 0.1
        bi = torch.zeros(nj, device=dev)
        plan = engine.Plan(bu_, bi_, br_, n_users_local, nj, n_factors=F, hot_min_degree=args.hot_min_degree)
        info = plan.info()
        res = {}
        for ph in ([7, 1, 2, 4] if args.phases else [7]):
            plan.set_phases(ph)
            ms = []
            for _ in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                engine.kmf_sgd_epoch(plan, "linear", P, Q, bu, bi, F, mu, wl["lr"], wl["reg"], 1.0 / F, 0.0, 5.0)
                b.record()
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            res[ph] = float(np.median(ms[1:]))
        tot += res[7]
        print(f"block ({rank},{j}): n={int(bu_.numel())} hot_items={info['n_hot_items']} hot_ratings={info['n_hot_ratings']} "
              f"hot_users={info['n_hot_users']}/{info['n_hot_user_ratings']} max_item_deg={info['max_item_degree']} "
              f"max_user_deg={info['max_user_degree']} flat={info.get('flat')} W={info['n_workers']} ms={res}")
        plan.close()
    print(f"G={G} rank {rank}: sum of block launches {tot:.3f} ms per epoch ({N} ratings in the job)")


if __name__ == "__main__":
    main()
