"""Diagnostics: hand-off latency of the ring (k_sgd_ring) -- an epoch over very few ratings, so that the epoch time is
the time progress needs to travel around the ring.
    python tools/ring_latency.py [--n 200000] [--workers 2220 --warps 15] [--slack 1]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from matrix_factorization_b200 import engine

    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200000)
    ap.add_argument("--users", type=int, default=138493)
    ap.add_argument("--items", type=int, default=26744)
    ap.add_argument("--factors", type=int, default=128)
    ap.add_argument("--workers", type=int, default=2220)
    ap.add_argument("--warps", type=int, default=15)
    ap.add_argument("--slack", type=int, default=1)
    ap.add_argument("--epochs", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    U, I, F, N = args.users, args.items, args.factors, args.n
    key = torch.randperm(U * I, device=dev, generator=g)[:N] if U * I < 2 ** 31 else torch.unique(
        torch.randint(0, U * I, (int(N * 1.1),), device=dev, generator=g))[:N]
    u, i = (key // I).int(), (key % I).int()
    r = torch.randint(1, 6, (len(u),), device=dev, generator=g).float()
    P = torch.randn(U, F, device=dev, generator=g) * 0.1
    Q = torch.randn(I, F, device=dev, generator=g) * 0.1
    bu, bi = torch.zeros(U, device=dev), torch.zeros(I, device=dev)
    plan = engine.Plan(u, i, r, U, I, n_factors=F, n_workers=args.workers, warps_per_cta=args.warps,
                       hot_min_degree=engine.Plan.NO_HOT_SPLIT, stripe_slack=args.slack)
    info = plan.info()
    ms = []
    for e in range(args.epochs):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        engine.kmf_sgd_epoch(plan, "linear", P, Q, bu, bi, F, 3.0, 0.01, 0.01, 1.0 / F, 0.0, 5.0)
        b.record()
        torch.cuda.synchronize()
        ms.append(round(a.elapsed_time(b), 3))
    st = plan.stats()
    W = info["n_workers"]
    print("n", len(u), "workers", W, "steps", info["n_steps"], "epoch ms", ms,
          "us per worker hop", round(1e3 * ms[-1] / W, 3), "max total Mcyc", st[:W, 0].max() / 1e6,
          "median blocked Mcyc", float(np.median(st[:W, 1])) / 1e6)


if __name__ == "__main__":
    main()
