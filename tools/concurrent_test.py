"""Experiment: do two cooperative SGD launches on two streams overlap?  Two half-plans over disjoint user halves of the
bench workload (own parameter arrays): hot phase of half A next to the cold phases of half B."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from matrix_factorization_b200 import engine

    dev = torch.device("cuda", 0)
    wl = bench.gen_workload("ml-20m", dev)
    F, U, I = wl["F"], wl["U"], wl["I"]
    half = (wl["u"] % 2) == 0
    plans, params = [], []
    for h in (half, ~half):
        u, i, r = wl["u"][h].contiguous(), wl["i"][h].contiguous(), wl["r"][h].contiguous()
        plans.append(engine.Plan(u, i, r, U, I, n_factors=F, hot_min_degree=0))
        g = torch.Generator(device=dev).manual_seed(5)
        params.append((torch.randn(U, F, device=dev, generator=g) * 0.1, torch.randn(I, F, device=dev, generator=g) * 0.1,
                       torch.zeros(U, device=dev), torch.zeros(I, device=dev)))
        print(plans[-1].info())
    plans[0].set_phases(1)
    plans[1].set_phases(6)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(k):
        P, Q, bu, bi = params[k]
        engine.kmf_sgd_epoch(plans[k], "linear", P, Q, bu, bi, F, 3.5, wl["lr"], wl["reg"], 1.0 / F, 0.0, 5.0)

    def timed(fn):
        ms = []
        for _ in range(4):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms[1:]))

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            run(0)
        with torch.cuda.stream(s2):
            run(1)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    print("hot(A) alone ms", timed(lambda: run(0)))
    print("rest(B) alone ms", timed(lambda: run(1)))
    print("both, two streams ms", timed(both))
    for w in (64, 96):
        os.environ["MFK_HOT_WORKERS"] = str(w)
        pa = engine.Plan(wl["u"][half].contiguous(), wl["i"][half].contiguous(), wl["r"][half].contiguous(), U, I, n_factors=F,
                         hot_min_degree=0)
        pa.set_phases(1)
        pb = engine.Plan(wl["u"][~half].contiguous(), wl["i"][~half].contiguous(), wl["r"][~half].contiguous(), U, I, n_factors=F,
                         hot_min_degree=0, n_workers=148 - w, schedule=3)
        pb.set_phases(6)
        plans[0], plans[1] = pa, pb
        print("hot workers", w, "flat workers", 148 - w, pa.info()["n_hot_workers"], pb.info()["n_workers"], pb.info()["n_hot_user_workers"],
              "hot alone", timed(lambda: run(0)), "rest alone", timed(lambda: run(1)), "both", timed(both))


if __name__ == "__main__":
    main()
