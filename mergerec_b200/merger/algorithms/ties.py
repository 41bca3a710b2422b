"""TIES (A6-A9); reference: rec_retrieval/merger/algorithms/ties.py.  Filled in by csrc/ties.cu."""
from __future__ import annotations

from typing import List

from ..types import FlattenedModel, FlattenedModel2D


def get_ties_vectors(base_model: FlattenedModel, models: List[FlattenedModel], density: float, **__) -> FlattenedModel2D:
    raise NotImplementedError("TIES kernels are not built yet")


def merge_ties(base_model: FlattenedModel, models: List[FlattenedModel], weights: List[float], density: float, **__
               ) -> FlattenedModel:
    raise NotImplementedError("TIES kernels are not built yet")
