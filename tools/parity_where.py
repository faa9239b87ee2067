"""Where does the one-epoch fp32-vs-oracle error sit?  Rows with the largest error, their degrees and phases."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from matrix_factorization_b200 import engine
from matrix_factorization_b200.data import synth_ratings_torch
from oracle import oracle as orc

dev = torch.device("cuda", 0)
U, I, n, F, seed = 40_000, 8_000, 2_000_000, 128, 1234
u, i, r = synth_ratings_torch(U, I, n, seed, dev, grid_step=0.5)
g = torch.Generator(device=dev).manual_seed(seed)
P = torch.randn(U, F, device=dev, generator=g) * 0.1
Q = torch.randn(I, F, device=dev, generator=g) * 0.1
bu, bi = torch.zeros(U, device=dev), torch.zeros(I, device=dev)
P0, Q0 = P.double().cpu().numpy(), Q.double().cpu().numpy()
mu = float(r.double().mean().item())
for hot in (0, engine.Plan.NO_HOT_SPLIT):
    Pd, Qd, bud, bid = P.clone(), Q.clone(), bu.clone(), bi.clone()
    plan = engine.Plan(u, i, r, U, I, n_factors=F, hot_min_degree=hot)
    engine.kmf_sgd_epoch(plan, "linear", Pd, Qd, bud, bid, F, mu, 0.001, 0.005, 1.0 / F, 0.0, 5.0)
    order = plan.order().cpu().numpy()
    Po, Qo, buo, bio = orc.kmf_replay("linear", u.cpu().numpy(), i.cpu().numpy(), r.double().cpu().numpy(), order, mu,
                                      np.zeros(U), np.zeros(I), P0, Q0, 0.001, 0.005)
    eQ = np.abs(Qd.double().cpu().numpy() - Qo).max(1)
    eP = np.abs(Pd.double().cpu().numpy() - Po).max(1)
    di = np.bincount(i.cpu().numpy(), minlength=I)
    du = np.bincount(u.cpu().numpy(), minlength=U)
    print("hot" if hot == 0 else "no-hot-split", plan.info()["n_hot_items"], "max|Q|", np.abs(Qo).max(), "max|P|", np.abs(Po).max())
    for k in np.argsort(-eQ)[:6]:
        print("  item", k, "deg", di[k], "err", eQ[k], "|q|max", np.abs(Qo[k]).max())
    for k in np.argsort(-eP)[:4]:
        print("  user", k, "deg", du[k], "err", eP[k], "|p|max", np.abs(Po[k]).max())
    # fp32 rounding floor: the oracle's own result rounded to fp32 at every step is not available; report eps * sqrt(deg) * |q|
    k = int(np.argmax(eQ))
    print("  eps*sqrt(deg)*|q| for the worst item:", 6e-8 * np.sqrt(di[k]) * np.abs(Qo[k]).max())
