"""Host-only helpers (reference: matrix_factorization/utils.py)."""
from typing import Tuple

import numpy as np
import pandas as pd
from sklearn.model_selection import train_test_split


def train_update_test_split(
    X: pd.DataFrame, frac_new_users: float
) -> Tuple[pd.DataFrame, pd.Series, pd.DataFrame, pd.Series, pd.DataFrame, pd.Series]:
    """
    Three-way split for testing `update_users` (reference: utils.py:8-72): a random
    `round(frac_new_users * n_users)` users are held out; all other users' ratings form
    train_initial (shuffled); every held-out user's ratings are split 50/50, stratified by user,
    into train_update and test_update.

    Usage: fit on train_initial, `update_users` with train_update, evaluate on test_update.

    Returns X_train_initial, y_train_initial, X_train_update, y_train_update, X_test_update,
    y_test_update  (X_* have columns user_id, item_id; y_* is the rating column).
    """
    all_users = X["user_id"].unique()
    n_held = round(frac_new_users * len(all_users))
    held_out = np.random.choice(all_users, size=n_held, replace=False)

    is_held = X["user_id"].isin(held_out)
    initial = X[~is_held].sample(frac=1, replace=False)
    held = X[is_held]
    update, test = train_test_split(held, stratify=held["user_id"], test_size=0.5)

    cols = ["user_id", "item_id"]
    return (initial[cols], initial["rating"], update[cols], update["rating"], test[cols], test["rating"])
