"""Shared argument plumbing for the merge entry points."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from ... import _lib


def as_rows(models) -> List[torch.Tensor]:
    """A list of contiguous 1-D fp32 CUDA rows from a list of flat tensors or a (K, d) tensor."""
    rows = list(models.unbind(0)) if isinstance(models, torch.Tensor) else list(models)
    out = []
    for r in rows:
        if not r.is_cuda:
            raise _lib.MergeRecLibraryError("flat models must live on the GPU (use flatten_model / ModelMerger)")
        if r.dtype != torch.float32:
            r = r.float()
        if r.dim() != 1 or r.stride(0) != 1:
            r = r.reshape(-1).contiguous()
        out.append(r)
    return out


def weights_tensor(weights: Sequence[float], device) -> torch.Tensor:
    """Python floats -> fp32 lambdas (the reference's scalar-mul kernels round the double to fp32 first)."""
    return torch.tensor([float(w) for w in weights], dtype=torch.float32, device=device).reshape(1, -1)


def merge_axpy(base: Optional[torch.Tensor], rows: Sequence[torch.Tensor], w: torch.Tensor, order: int,
               src_is_model: bool, seg_end: Optional[torch.Tensor] = None, seg_group: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    K, d = len(rows), rows[0].numel()
    assert w.dtype == torch.float32 and w.is_contiguous() and w.shape[-1] == K
    G = w.numel() // K
    P = 1 if seg_end is None else seg_end.numel()
    if out is None:
        out = torch.empty(d, dtype=torch.float32, device=rows[0].device)
    rc = lib.mr_merge_axpy(_lib.dptr(base), _lib.ptr_array(rows), K, d, _lib.dptr(w, torch.float32), G,
                           _lib.dptr(seg_end), _lib.dptr(seg_group), P, order, int(src_is_model),
                           _lib.dptr(out, torch.float32), _lib.stream_handle())
    _lib.check(rc, "mr_merge_axpy")
    return out
