"""Profiling driver: the phases of one epoch on a bench workload, nothing else (for ncu).
    python tools/prof_hot.py [--workload ml-20m] [--phases 1] [--epochs 2]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from matrix_factorization_b200 import engine

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ml-20m")
    ap.add_argument("--phases", type=int, default=1)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--hot-min-degree", type=int, default=0)
    ap.add_argument("--slack", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = bench.gen_workload(args.workload, dev)
    F, U, I = wl["F"], wl["U"], wl["I"]
    g = torch.Generator(device=dev).manual_seed(5)
    P = torch.randn(U, F, device=dev, generator=g) * 0.1
    Q = torch.randn(I, F, device=dev, generator=g) * 0.1
    bu, bi = torch.zeros(U, device=dev), torch.zeros(I, device=dev)
    mu = float(wl["r"].double().mean().item())
    plan = engine.Plan(wl["u"], wl["i"], wl["r"], U, I, n_factors=F, hot_min_degree=args.hot_min_degree, stripe_slack=args.slack)
    plan.set_phases(args.phases)
    for _ in range(args.epochs):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        engine.kmf_sgd_epoch(plan, "linear", P, Q, bu, bi, F, mu, wl["lr"], wl["reg"], 1.0 / F, 0.0, 5.0)
        b.record()
        torch.cuda.synchronize()
        print("epoch ms", a.elapsed_time(b), plan.info())


if __name__ == "__main__":
    main()
