"""
Top-k evaluation on the batched recommend path.

The reference's callers loop `model.recommend(user, amount=k, items_known=train_items)` over the users in Python
(project_template/pipeline/evaluate.py:61-111) and average precision / recall / NDCG@k.  `evaluate_topk` keeps that
signature and those per-user definitions, but asks the model for every user's list at once (`recommend_all`: one
scoring pass with each user's training items masked inside the kernel).

Per user with hit vector h (1 where the recommended item is relevant) over the `m` recommended items:
    precision = mean(h)                       recall = sum(h) / max(1, |relevant|)
    ndcg = sum((2^h - 1) / log2(rank + 1)) / the same sum for h sorted descending      (0 when there is no hit)
Users whose history is too short to split, or that end up with no training or no test item, are skipped -- like
the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np
import pandas as pd


@dataclass
class TopKResult:
    precision: float
    recall: float
    ndcg: float
    n_users: int = 0


def holdout_split(ratings: pd.DataFrame, n_test: int, positive_threshold: float, seed: int) -> Tuple[pd.DataFrame, pd.DataFrame]:
    """
    Per-user hold-out with the reference's rule (evaluate.py:33-57): users in first-appearance order, one RandomState
    for all of them; test = `n_test` items sampled from the user's ratings >= positive_threshold (the n_test highest
    rated ones if there are not enough), train = the user's other items.  Users with <= n_test ratings are dropped.
    Returns (train, test) frames with the columns of `ratings`.
    """
    rng = np.random.RandomState(seed)
    train_parts, test_parts = [], []
    for _, hist in ratings.groupby("user_id", sort=False):
        if hist.shape[0] <= n_test:
            continue
        pos = hist[hist["rating"] >= positive_threshold]
        if pos.shape[0] >= n_test:
            test = pos.sample(n=n_test, random_state=rng)
        else:
            test = hist.sort_values("rating", ascending=False).head(n_test)
        train = hist.loc[~hist["item_id"].isin(test["item_id"].tolist())]
        if len(train) == 0 or len(test) == 0:
            continue
        train_parts.append(train)
        test_parts.append(test)
    if not train_parts:
        return ratings.iloc[:0], ratings.iloc[:0]
    return pd.concat(train_parts), pd.concat(test_parts)


def topk_metrics(rec: pd.DataFrame, test: pd.DataFrame, k: int) -> TopKResult:
    """Metrics of recommendation lists `rec` (columns user_id, item_id, rank as returned by recommend_all) against the
    relevant items in `test` (columns user_id, item_id)."""
    if len(rec) == 0:
        return TopKResult(0.0, 0.0, 0.0, 0)
    rel = test[["user_id", "item_id"]].drop_duplicates().assign(_hit=1)
    m = rec.merge(rel, on=["user_id", "item_id"], how="left")
    hit = m["_hit"].fillna(0).to_numpy(dtype=np.float64)
    users, inv = np.unique(m["user_id"].to_numpy(), return_inverse=True)
    rank = m["rank"].to_numpy()
    n_rec = np.bincount(inv, minlength=len(users)).astype(np.float64)
    hits = np.bincount(inv, weights=hit, minlength=len(users))
    n_rel = rel.groupby("user_id").size().reindex(users).fillna(0).to_numpy(dtype=np.float64)
    precision = hits / np.maximum(n_rec, 1.0)
    recall = hits / np.maximum(n_rel, 1.0)
    dcg = np.bincount(inv, weights=hit / np.log2(rank + 2.0), minlength=len(users))
    # ideal: the user's hits moved to the front of the list
    disc = np.concatenate([[0.0], np.cumsum(1.0 / np.log2(np.arange(2, k + 2)))])
    idcg = disc[np.minimum(hits.astype(np.int64), k)]
    ndcg = np.where(idcg > 0, dcg / np.where(idcg > 0, idcg, 1.0), 0.0)
    return TopKResult(float(precision.mean()), float(recall.mean()), float(ndcg.mean()), int(len(users)))


def evaluate_topk(ratings: pd.DataFrame, model, k: int, positive_threshold: float, n_test: int, seed: int) -> TopKResult:
    """
    Drop-in for the reference's `evaluate_topk` (evaluate.py:61-111): split every user's history into train / test,
    recommend `k` items per user with the training items excluded, average precision / recall / NDCG@k over the users.
    `model` is a fitted estimator of this package; users or items it does not know are skipped / never recommended.
    """
    for col in ("user_id", "item_id", "rating"):
        if col not in ratings.columns:
            raise ValueError(f"ratings is missing column {col!r}")
    train, test = holdout_split(ratings, n_test, positive_threshold, seed)
    users = [u for u in train["user_id"].unique().tolist() if model.contains_user(u)]
    if not users:
        return TopKResult(0.0, 0.0, 0.0, 0)
    rec = model.recommend_all(users=users, amount=k, items_known=train[["user_id", "item_id"]])
    return topk_metrics(rec, test[test["user_id"].isin(users)], k)
