"""Profiling driver: ONE hot item rated by every user -- a single chain on a single CTA of the batch engine, so that ncu's
per-kernel counters and stall samples describe the chain itself (nothing waits on a ring neighbour).
    python tools/prof_chain.py [--users 138000] [--factors 128] [--epochs 3]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from matrix_factorization_b200 import engine

    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=138000)
    ap.add_argument("--items", type=int, default=1)
    ap.add_argument("--factors", type=int, default=128)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--lr", type=float, default=0.001)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    U, I, F = args.users, args.items, args.factors
    g = torch.Generator(device=dev).manual_seed(5)
    u = torch.arange(U, device=dev, dtype=torch.int32).repeat(I)
    i = torch.arange(I, device=dev, dtype=torch.int32).repeat_interleave(U)
    perm = torch.randperm(U * I, device=dev, generator=g)
    u, i = u[perm].contiguous(), i[perm].contiguous()
    r = torch.randint(1, 6, (U * I,), device=dev, generator=g).float()
    P = torch.randn(U, F, device=dev, generator=g) * 0.1
    Q = torch.randn(I, F, device=dev, generator=g) * 0.1
    bu, bi = torch.zeros(U, device=dev), torch.zeros(I, device=dev)
    plan = engine.Plan(u, i, r, U, I, n_factors=F, hot_min_degree=0)
    print(plan.info())
    for _ in range(args.epochs):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        engine.kmf_sgd_epoch(plan, "linear", P, Q, bu, bi, F, 3.0, args.lr, 0.005, 1.0 / F, 0.0, 5.0)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print("epoch ms", ms, "ns per chained rating", 1e6 * ms / U)


if __name__ == "__main__":
    main()
