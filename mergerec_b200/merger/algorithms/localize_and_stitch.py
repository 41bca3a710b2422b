"""Dataless Localize-and-Stitch vectors (reference: rec_retrieval/merger/algorithms/localize_and_stitch.py).

Localization keeps, per model, the ``int(density * d)`` largest ``|tau_k|`` over the whole flat vector -- the same
exact order statistic as the TIES trim (`mr_ties_select`, lowest flat index first among equal magnitudes, where
the reference inherits ``torch.topk``'s unspecified order); stitching scales every kept entry by one over the
number of models that kept that position.  One `mr_ties_build` pass in ``MR_TIES_LNS`` mode.
"""
from __future__ import annotations

from typing import List

import torch

from ... import _lib
from ..layout import alloc_rows
from ..types import FlattenedModel, FlattenedModel2D
from ._common import as_rows, merge_axpy
from .ties import _build, ties_select


def get_localize_and_stitch_vectors(base_model: FlattenedModel, models: List[FlattenedModel], density: float = 0.05,
                                    **__) -> FlattenedModel2D:
    """(K, d): ``(mask_k / max(sum_j mask_j, 1)) * (models[k] - base)`` (localize_and_stitch.py:26-49).
    ``int(density * d) <= 0`` returns zeros like the reference (:31-33)."""
    assert len(models) > 0, "models must be non-empty."
    rows = as_rows(models)
    K, d = len(rows), base_model.numel()
    out = alloc_rows(K, d, base_model.device)
    if int(density * d) <= 0:
        return out.zero_()
    cut = ties_select(base_model, rows, density)
    _build(base_model, rows, cut, _lib.MR_TIES_LNS, out=out, ldo=max(out.stride(0), d))
    return out


def merge_localize_and_stitch(base_model: FlattenedModel, models: List[FlattenedModel], weights: List[float],
                              density: float = 0.05, **__) -> FlattenedModel:
    """``base + sum_dim0_k(w_k * vectors[k])`` (localize_and_stitch.py:52-82): the weights scale the stitched
    vectors, the sum runs in ``torch.sum(dim=0)`` order."""
    assert len(models) == len(weights), "Number of models and weights should match."
    vectors = get_localize_and_stitch_vectors(base_model, models, density=density)
    w = torch.tensor([float(x) for x in weights], dtype=torch.float32, device=base_model.device).reshape(1, -1)
    return merge_axpy(base_model, list(vectors.unbind(0)), w, _lib.MR_ORDER_SUM_FIRST, False)
