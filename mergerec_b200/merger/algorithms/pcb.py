"""PCB merging; reference: rec_retrieval/merger/algorithms/pcb.py:9-73.

``get_pcb_vectors`` = three exact order statistics per model plus one build pass, all on the GPU:

* the 1 % / 99 % clamp of ``|tau_k|`` (``_clamp(torch.abs(task_vectors), 0.01, 0.01)``, pcb.py:17-27,42) are the
  ``d - int(d*0.01)``-th and ``d - int(d*0.99 - 1)``-th LARGEST magnitudes -- found with the TIES selection kernels;
* the lower clamp of the balancing weights (``_clamp(task_pcb, 1 - density, 0)``, pcb.py:53) is the
  ``int(d*(1-density))``-th smallest of values that only exist on the fly -- ``mr_pcb_vectors`` finds it with three
  histogram passes and never materialises a (K, d) temporary.

Floating-point contract: torch's CPU ``exp`` / ``tanh`` and CUDA's differ in the last bit, so the vectors agree with the
reference to ~1e-6 of the row scale except in columns holding an element whose balancing weight lies within a few ulp
of the clamp (there the reference's normalisation is discontinuous: ``scale / max(sum(scale), 1e-12)``)."""
from __future__ import annotations

from typing import List

import torch

from ... import _lib
from ..layout import alloc_rows
from ..types import FlattenedModel
from ._common import as_rows, merge_axpy, weights_tensor
from .ties import select_kth_largest

__all__ = ["get_pcb_vectors", "merge_pcb"]


def _magnitude_of(cut: torch.Tensor) -> torch.Tensor:
    """fp32 magnitude held in the high word of the selection keys."""
    return (cut >> 32).to(torch.int32).view(torch.float32)


def get_pcb_vectors(base_model: FlattenedModel, models: List[FlattenedModel], density: float = 0.2,
                    return_diagnostics: bool = False, force_dense: bool = False, **__):
    lib = _lib.load()
    rows = as_rows(models)
    K, d = len(rows), base_model.numel()
    dev = base_model.device
    i_lo, i_hi = int(d * 0.01), int(d * (1 - 0.01) - 1)          # pcb.py:20-21 with min_ratio = max_ratio = 0.01
    i_hi = i_hi if i_hi >= 0 else d + i_hi                       # sorted_x[-1] for tiny d, like the reference's indexing
    if i_lo == 0:
        # d < 100: sorted_x[0] is the row minimum of |tau| (pcb.py:20); the order-statistic select answers "keep
        # everything" for k = d with cut 0 rather than the d-th largest magnitude, so take the minimum directly
        lo = torch.stack([(r - base_model).abs().min() for r in rows]).to(torch.float32).contiguous()
    else:
        lo = _magnitude_of(select_kth_largest(base_model, rows, d - i_lo)).contiguous()
    hi = _magnitude_of(select_kth_largest(base_model, rows, d - i_hi)).contiguous()
    q_index = int(d * (1 - density))                             # pcb.py:53: min_ratio = 1 - density, max_ratio = 0
    out = alloc_rows(K, d, dev)
    ldo = max(out.stride(0), d)
    task = alloc_rows(K, d, dev) if return_diagnostics else None
    thr = torch.empty((K, 2), dtype=torch.float32, device=dev) if return_diagnostics else None
    ws_bytes = int(lib.mr_pcb_workspace_bytes(K))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    status = torch.zeros(K, dtype=torch.int32, device=dev)
    for dense in ((1,) if force_dense else (0, 1)):
        # fast path: dense radix passes on a 1/32 sample, then two windowed passes over everything; a window that misses
        # the wanted rank is reported in `status` and the three dense passes over the whole vector run instead
        rc = lib.mr_pcb_vectors(_lib.dptr(base_model, torch.float32), _lib.ptr_array(rows), K, d, _lib.dptr(lo), _lib.dptr(hi),
                                q_index, dense, _lib.dptr(status), _lib.dptr(out), ldo, _lib.dptr(task), _lib.dptr(thr),
                                _lib.dptr(ws), ws_bytes, _lib.stream_handle())
        _lib.check(rc, "mr_pcb_vectors")
        if bool((status.cpu() == 1).all()):
            break
    else:
        raise _lib.MergeRecLibraryError("PCB quantile search failed")
    if return_diagnostics:
        return out, task, thr, lo, hi
    return out


def merge_pcb(base_model: FlattenedModel, models: List[FlattenedModel], weights: List[float], density: float = 0.2, **__
              ) -> FlattenedModel:
    """``merged = base.clone(); merged += weights[i] * pcb_vector_i`` (pcb.py:61-73): base-first accumulation."""
    assert len(models) == len(weights), "Number of models and weights should match."
    vectors = get_pcb_vectors(base_model, models, density=density)
    return merge_axpy(base_model, list(vectors.unbind(0)), weights_tensor(weights, base_model.device),
                      _lib.MR_ORDER_BASE_FIRST, False)
