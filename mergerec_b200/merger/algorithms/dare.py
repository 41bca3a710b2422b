"""DARE merge (drop and rescale); reference: rec_retrieval/merger/algorithms/dare.py:9-31.

``merged = base; for i: merged += dropout(weights[i] * (m_i - base), p=density)`` -- note that the reference hands
``density`` to ``dropout`` as the DROP probability.  torch's dropout is ``input * (bernoulli(1 - p) / (1 - p))``; given
the same keep masks the CUDA kernel (``mr_merge_dare``) reproduces the reference bit for bit.  The masks themselves are
random: without ``masks=`` they are drawn on the GPU from ``generator`` (torch's CUDA Philox stream, which is not the
CPU generator's stream -- same distribution, different bits), so runs are reproducible for a seeded generator but do
not replay a CPU run of the reference; pass the reference's masks to do that (the tests do)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from ... import _lib
from ..types import FlattenedModel
from ._common import as_rows, weights_tensor

__all__ = ["merge_dare"]


def merge_dare(base_model: FlattenedModel, models: List[FlattenedModel], weights: List[float], density: float,
               masks: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None, **__) -> FlattenedModel:
    assert len(models) == len(weights), "Number of models and weights should match."
    p = float(density)
    if p < 0.0 or p > 1.0:
        raise ValueError(f"dropout probability has to be between 0 and 1, but got {p}")
    rows = as_rows(models)
    K, d = len(rows), base_model.numel()
    dev = base_model.device
    if masks is None:
        masks = torch.empty((K, d), dtype=torch.float32, device=dev).bernoulli_(1.0 - p, generator=generator) != 0
    masks = masks.to(device=dev)
    if masks.shape != (K, d):
        raise ValueError(f"masks must have shape {(K, d)}, got {tuple(masks.shape)}")
    keep = masks.to(torch.uint8).contiguous()
    if p == 1.0:                 # torch: `input * zeros`
        keep, scale = torch.zeros_like(keep), 0.0
    elif p == 0.0:               # torch returns the input unchanged
        keep, scale = torch.ones_like(keep), 1.0
    else:
        scale = float(np.float32(1.0) / np.float32(1.0 - p))      # `noise.div_(1 - p)` on an fp32 tensor
    out = torch.empty_like(base_model)
    w = weights_tensor(weights, dev)
    rc = _lib.load().mr_merge_dare(_lib.dptr(base_model, torch.float32), _lib.ptr_array(rows), K, d, _lib.dptr(w),
                                   _lib.dptr(keep), keep.stride(0) if d else d, scale, _lib.dptr(out), _lib.stream_handle())
    _lib.check(rc, "mr_merge_dare")
    return out
