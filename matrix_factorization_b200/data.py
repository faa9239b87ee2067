"""
Synthetic MovieLens/Netflix-shaped ratings.  No dataset ships with the
reference (its .gitignore drops /data) and there is no network, so every measurement and
parity run uses these generators.

* unique (user, item) pairs with Zipf-like marginals: user activity ~ rank^-0.8, item
  popularity ~ rank^-1.0, sampled then de-duplicated until N pairs exist;
* ratings from a planted model  r = clip(grid(3.5 + b_u + b_i + p_u.q_i + eps));
* raw ids are a random permutation of 1..U / 1..I (int64) so the id-map path is exercised.

`synth_ratings` is the numpy generator (tests, goldens, small configs).  `synth_ratings_torch`
draws from the same distribution with torch ops on a device for the 20M/100M-row configs.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

# name -> (n_users, n_items, n_ratings, rating grid step, min ratings per user)
SHAPES = {
    "ml-100k": (943, 1682, 100_000, 1.0, 20),
    "ml-1m": (6040, 3706, 1_000_000, 1.0, 20),
    "ml-20m": (138_493, 26_744, 20_000_000, 0.5, 20),
    "netflix": (480_189, 17_770, 100_480_507, 1.0, 1),
}

PLANT_RANK = 16


def _zipf_cdf(n: int, expo: float) -> np.ndarray:
    w = np.arange(1, n + 1, dtype=np.float64) ** (-expo)
    c = np.cumsum(w)
    return c / c[-1]


def _planted(rng, n_users, n_items):
    bu = rng.normal(0.0, 0.3, n_users)
    bi = rng.normal(0.0, 0.3, n_items)
    pu = rng.normal(0.0, 0.35, (n_users, PLANT_RANK))
    qi = rng.normal(0.0, 0.35, (n_items, PLANT_RANK))
    return bu, bi, pu, qi


def _grid(x: np.ndarray, step: float) -> np.ndarray:
    lo = 1.0 if step == 1.0 else step
    return np.clip(np.round(x / step) * step, lo, 5.0)


def synth_pairs(n_users, n_items, n_ratings, seed, uniform=False, min_per_user=0,
                zipf_user=0.8, zipf_item=1.0):
    """N unique (u, i) pairs (0-based ranks: low rank = popular), int64 arrays in random order."""
    if n_ratings > n_users * n_items:
        raise ValueError("more ratings than cells")
    rng = np.random.default_rng(seed)
    cu = None if uniform else _zipf_cdf(n_users, zipf_user)
    ci = None if uniform else _zipf_cdf(n_items, zipf_item)

    def draw_items(m):
        return rng.integers(0, n_items, m) if uniform else np.searchsorted(ci, rng.random(m))

    def draw_users(m):
        return rng.integers(0, n_users, m) if uniform else np.searchsorted(cu, rng.random(m))

    keys = np.empty(0, dtype=np.int64)
    if min_per_user > 0:
        m = int(min(min_per_user, n_items))
        for _ in range(6):  # top up users that are still short after de-duplication
            cnt = np.bincount(keys // n_items, minlength=n_users) if len(keys) else np.zeros(n_users, np.int64)
            short = np.nonzero(cnt < m)[0]
            if len(short) == 0:
                break
            uu = np.repeat(short, np.ceil((m - cnt[short]) * 1.5).astype(np.int64) + 1)
            keys = np.unique(np.concatenate([keys, uu * n_items + draw_items(len(uu))]))
    base = keys
    while len(keys) < n_ratings:
        need = n_ratings - len(keys)
        m = int(need * 1.3) + 1024
        keys = np.unique(np.concatenate([keys, draw_users(m) * n_items + draw_items(m)]))
    if len(keys) > n_ratings:  # trim extras, never the min-per-user base pairs
        extra = np.setdiff1d(keys, base, assume_unique=True)
        extra = rng.permutation(extra)[: n_ratings - len(base)] if len(base) < n_ratings else extra[:0]
        keys = np.concatenate([base[:n_ratings], extra])
    keys = rng.permutation(keys)
    return keys // n_items, keys % n_items


def synth_ratings(n_users, n_items, n_ratings, seed, grid_step=1.0, uniform=False,
                  min_per_user=0, raw_ids=True) -> pd.DataFrame:
    """DataFrame[user_id, item_id, rating] of a planted low-rank model on Zipf-shaped pairs."""
    u, i = synth_pairs(n_users, n_items, n_ratings, seed, uniform=uniform, min_per_user=min_per_user)
    rng = np.random.default_rng(seed + 7919)
    if uniform:
        r = rng.integers(1, 6, n_ratings).astype(np.float64)
    else:
        bu, bi, pu, qi = _planted(rng, n_users, n_items)
        x = 3.5 + bu[u] + bi[i] + np.einsum("nk,nk->n", pu[u], qi[i]) + rng.normal(0.0, 0.7, n_ratings)
        r = _grid(x, grid_step)
    if raw_ids:
        uperm = rng.permutation(n_users).astype(np.int64) + 1
        iperm = rng.permutation(n_items).astype(np.int64) + 1
        u, i = uperm[u], iperm[i]
    return pd.DataFrame({"user_id": u.astype(np.int64), "item_id": i.astype(np.int64), "rating": r})


def synth_config(name: str, seed: int | None = None, uniform=False, scale: float = 1.0) -> pd.DataFrame:
    """One of SHAPES, optionally scaled down (users, items by sqrt(scale); ratings by scale)."""
    U, I, N, step, mpu = SHAPES[name]
    if scale != 1.0:
        U, I, N = max(8, int(U * scale ** 0.5)), max(8, int(I * scale ** 0.5)), max(64, int(N * scale))
    if seed is None:
        seed = 1000 + list(SHAPES).index(name)
    return synth_ratings(U, I, N, seed, grid_step=step, uniform=uniform, min_per_user=mpu)


def split_rows(df: pd.DataFrame, test_frac=0.1, seed=0):
    """90/10 train/test split by rows."""
    rng = np.random.default_rng(seed)
    m = rng.random(len(df)) < test_frac
    return df[~m].reset_index(drop=True), df[m].reset_index(drop=True)


def synth_ratings_torch(n_users, n_items, n_ratings, seed, device, grid_step=1.0, uniform=False,
                        zipf_user=0.8, zipf_item=1.0):
    """Same distribution as synth_ratings, generated with torch on `device` for the big configs.
    Returns (u int32, i int32, r float32) tensors of internal 0-based ids in random order.
    (Data generation only -- never part of a timed region or of the product path.)"""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)

    def cdf(n, expo):
        w = torch.arange(1, n + 1, device=device, dtype=torch.float64) ** (-expo)
        c = torch.cumsum(w, 0)
        return c / c[-1]

    cu, ci = (None, None) if uniform else (cdf(n_users, zipf_user), cdf(n_items, zipf_item))
    keys = torch.empty(0, dtype=torch.int64, device=device)
    while keys.numel() < n_ratings:
        m = int((n_ratings - keys.numel()) * 1.3) + 4096
        if uniform:
            uu = torch.randint(0, n_users, (m,), device=device, generator=g)
            ii = torch.randint(0, n_items, (m,), device=device, generator=g)
        else:
            uu = torch.searchsorted(cu, torch.rand(m, device=device, dtype=torch.float64, generator=g))
            ii = torch.searchsorted(ci, torch.rand(m, device=device, dtype=torch.float64, generator=g))
            uu.clamp_(max=n_users - 1)
            ii.clamp_(max=n_items - 1)
        keys = torch.unique(torch.cat([keys, uu * n_items + ii]))
    perm = torch.randperm(keys.numel(), device=device, generator=g)
    keys = keys[perm[:n_ratings]]
    u, i = keys // n_items, keys % n_items
    if uniform:
        r = torch.randint(1, 6, (n_ratings,), device=device, generator=g).float()
    else:
        bu = torch.randn(n_users, device=device, generator=g) * 0.3
        bi = torch.randn(n_items, device=device, generator=g) * 0.3
        pu = torch.randn(n_users, PLANT_RANK, device=device, generator=g) * 0.35
        qi = torch.randn(n_items, PLANT_RANK, device=device, generator=g) * 0.35
        x = 3.5 + bu[u] + bi[i] + (pu[u] * qi[i]).sum(1) + torch.randn(n_ratings, device=device, generator=g) * 0.7
        lo = 1.0 if grid_step == 1.0 else grid_step
        r = (torch.round(x / grid_step) * grid_step).clamp_(lo, 5.0)
    # ids are shuffled ranks so popular users/items are not the low ids
    up = torch.randperm(n_users, device=device, generator=g)
    ip = torch.randperm(n_items, device=device, generator=g)
    return up[u].int(), ip[i].int(), r.float()
