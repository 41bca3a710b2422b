"""Evaluator hot path (reference: rec_retrieval/evaluator/__init__.py)."""
from .enums import MetricType
from .evaluator import Evaluator
from .metrics import NDCG, BaseMetric, Recall
from .sharded import (MR_SCORE_BF16, MR_SCORE_TF32X1, MR_SCORE_TF32X3, ShardedItemTable, replicate_from_host,
                      shard_bounds)

__all__ = ["Evaluator", "MetricType", "Recall", "NDCG", "BaseMetric", "ShardedItemTable", "shard_bounds", "replicate_from_host",
           "MR_SCORE_TF32X3", "MR_SCORE_TF32X1", "MR_SCORE_BF16"]
