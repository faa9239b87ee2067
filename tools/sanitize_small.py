"""Small runs of every SGD engine and the scoring kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from matrix_factorization_b200 import engine

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
for F, lr in ((96, 0.001), (200, 0.001), (64, 0.01)):
    U, I, N = 600, 60, 8000
    keys = rng.choice(U * I, N, replace=False)
    u, i = (keys // I).astype(np.int32), (keys % I).astype(np.int32)
    m = rng.random(N) < 0.4
    i[m] = rng.integers(0, 3, m.sum())
    keep = np.unique(u.astype(np.int64) * I + i, return_index=True)[1]
    u, i = u[keep], i[keep]
    r = rng.integers(1, 6, len(u)).astype(np.float32)
    tu, ti, tr = (torch.tensor(x).to(dev) for x in (u, i, r))
    ld = engine.round_up4(F)
    P = torch.randn(U, ld, device=dev) * 0.1
    Q = torch.randn(I, ld, device=dev) * 0.1
    bu, bi = torch.zeros(U, device=dev), torch.zeros(I, device=dev)
    plan = engine.Plan(tu, ti, tr, U, I, n_factors=F, hot_min_degree=16)
    for kern in ("linear", "sigmoid"):
        engine.kmf_sgd_epoch(plan, kern, P, Q, bu, bi, F, 3.0, lr, 0.02, 0.01, 0.0, 5.0)
    torch.cuda.synchronize()
    print("F", F, "lr", lr, plan.info()["n_hot_items"], "hot items ok", float(P.abs().max()))
    users = torch.arange(U, dtype=torch.int32, device=dev)
    sc, it = engine.score_topk("linear", users, P, Q, bu, bi, I, F, 3.0, 0.01, 0.0, 5.0, 10, True, None, None)
    torch.cuda.synchronize()
    print("score ok", int(it.min()))
