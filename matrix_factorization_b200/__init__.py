"""
matrix_factorization_b200 -- B200-native (sm_100a) drop-in for the training and scoring path of
the `matrix_factorization` package: KernelMF / BaselineModel fit, predict, update_users,
recommend.  The numeric work runs in hand-written CUDA kernels behind the C ABI of
include/mfk.h (libmfk_b200.so); there is no CPU fallback.

The reference's three classical recommenders (UserUserCF, ItemItemCF, ContentBasedRecommender)
are outside this path and are deliberately not provided.
"""
from .baseline_model import BaselineModel
from .kernel_matrix_factorization import KernelMF
from .recommender_base import RecommenderBase
from .utils import train_update_test_split

__all__ = ["BaselineModel", "KernelMF", "RecommenderBase", "train_update_test_split"]
