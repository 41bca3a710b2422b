"""Task arithmetic (A1, A2 of SURVEY.md section 8(a)); reference: rec_retrieval/merger/algorithms/task_vector.py."""
from __future__ import annotations

from typing import List

import torch

from ... import _lib
from ..layout import alloc_rows
from ..types import FlattenedModel, FlattenedModel2D
from ._common import as_rows, merge_axpy, weights_tensor


def get_task_vectors(base_model: FlattenedModel, models: List[FlattenedModel], **__) -> FlattenedModel2D:
    """T[k] = models[k] - base_model, stacked to (K, d)   (task_vector.py:8-10).
    One kernel reads the base once and writes all K rows; rows are 256-byte aligned."""
    rows = as_rows(models)
    K, d = len(rows), base_model.numel()
    T = alloc_rows(K, d, base_model.device)
    rc = _lib.load().mr_task_vectors(_lib.dptr(base_model, torch.float32), _lib.ptr_array(rows), K, d,
                                     _lib.dptr(T), T.stride(0), _lib.stream_handle())
    _lib.check(rc, "mr_task_vectors")
    return T


def merge_task_vector(base_model: FlattenedModel, models: List[FlattenedModel], weights: List[float], **__
                      ) -> FlattenedModel:
    """merged = base + sum_k weights[k] * (models[k] - base), accumulated base-first in k order with unfused
    fp32 ops exactly like the reference loop (task_vector.py:28-34) -> bit-identical results."""
    assert len(models) == len(weights), "Number of models and weights should match."
    rows = as_rows(models)
    w = weights_tensor(weights, base_model.device)
    return merge_axpy(base_model, rows, w, _lib.MR_ORDER_BASE_FIRST, src_is_model=True)
