"""Linear merge (A10); reference: rec_retrieval/merger/algorithms/linear.py."""
from __future__ import annotations

from typing import List

from ... import _lib
from ..types import FlattenedModel
from ._common import as_rows, merge_axpy, weights_tensor


def merge_linear(models: List[FlattenedModel], weights: List[float], **__) -> FlattenedModel:
    """merged = sum_k weights[k] * models[k] from a zero accumulator (linear.py:21-27); base_model ignored."""
    assert len(models) == len(weights), "Number of models and weights should match."
    rows = as_rows(models)
    w = weights_tensor(weights, rows[0].device)
    return merge_axpy(None, rows, w, _lib.MR_ORDER_LINEAR, src_is_model=False)
