"""TIES (A6-A9 of SURVEY.md section 8(a)); reference: rec_retrieval/merger/algorithms/ties.py.

The magnitude trim is global over the flat vector like the reference (``numel = base_model.numel()``,
ties.py:14-15).  Where the reference leaves the choice among equal magnitudes at the threshold to
``torch.topk``'s unspecified order, this implementation keeps the LOWEST flat indices (canonical tie rule);
away from exact threshold ties the results are bit-identical to the reference.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from ... import _lib
from ..layout import alloc_rows
from ..types import FlattenedModel
from ._common import as_rows


def ties_topk_count(density: float, numel: int) -> int:
    """``int(density * numel)`` in Python double arithmetic (ties.py:14-15)."""
    return int(density * numel)


def select_kth_largest(base_model: FlattenedModel, rows: Sequence[torch.Tensor], k_cnt: int,
                       w: Optional[torch.Tensor] = None, defer_status: bool = False):
    """Per-model 64-bit keys (int64 tensor of K, bit pattern of the uint64) of the ``k_cnt``-th largest
    ``|w_k (m_k - base)|`` under the order (magnitude, lower flat index first): model k keeps element j iff
    ``(bits(|u_kj|) << 32 | (0xFFFFFFFF - j)) >= cut[k]``.  Runs the stream-ordered sampled-bracket select and,
    if a model's bracket missed or overflowed (adversarial inputs), the exact multi-pass select.

    ``defer_status=True`` keeps the call fully asynchronous (no host synchronisation, CUDA-graph capturable): it
    returns ``(cut, status)`` and the caller must pass ``status`` to :func:`verify_select_status` before trusting a
    result built from ``cut`` -- a non-1 entry means "rerun with ``defer_status=False``"."""
    lib = _lib.load()
    K, d = len(rows), base_model.numel()
    dev = base_model.device
    ws_bytes = int(lib.mr_ties_workspace_bytes(d, K))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    cut = torch.zeros(K, dtype=torch.int64, device=dev)      # defined even when a deferred select fails (status says so)
    status = torch.zeros(K, dtype=torch.int32, device=dev)
    args = (_lib.dptr(base_model, torch.float32), _lib.ptr_array(rows), K, d, _lib.dptr(w), int(k_cnt), _lib.dptr(cut),
            _lib.dptr(status), _lib.dptr(ws), ws_bytes, _lib.stream_handle())
    _lib.check(lib.mr_ties_select(*args), "mr_ties_select")
    if defer_status:
        return cut, status
    st = status.cpu()
    if bool((st != 1).any()):
        _lib.check(lib.mr_ties_select_exact(*args), "mr_ties_select_exact")
        st = status.cpu()
        if bool((st != 1).any()):
            raise _lib.MergeRecLibraryError(f"TIES select failed with status {st.tolist()}")
    return cut


def verify_select_status(status: torch.Tensor) -> None:
    """Host check of a deferred select: raises unless every model's sampled bracket held its cut."""
    st = status.cpu()
    if bool((st != 1).any()):
        raise _lib.MergeRecLibraryError(
            f"deferred TIES select needs the exact path (status {st.tolist()}): rerun with defer_status=False")


def ties_select(base_model: FlattenedModel, rows: Sequence[torch.Tensor], density: float,
                w: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Cut keys of the TIES trim: the ``int(density * d)`` largest magnitudes survive (ties.py:14-23)."""
    return select_kth_largest(base_model, rows, ties_topk_count(density, base_model.numel()), w)


def _build(base_model, rows, cut, mode, w=None, G=1, seg_end=None, seg_group=None, out=None, ldo=0,
           trim_mask=None, elect_mask=None):
    K, d = len(rows), base_model.numel()
    P = 1 if seg_end is None else seg_end.numel()
    rc = _lib.load().mr_ties_build(_lib.dptr(base_model, torch.float32), _lib.ptr_array(rows), K, d, _lib.dptr(cut),
                                   mode, _lib.dptr(w), G, _lib.dptr(seg_end), _lib.dptr(seg_group), P, _lib.dptr(out),
                                   ldo, _lib.dptr(trim_mask), _lib.dptr(elect_mask), _lib.stream_handle())
    _lib.check(rc, "mr_ties_build")


def select_build(base_model: FlattenedModel, rows: Sequence[torch.Tensor], k_cnt: int, mode: int, out: torch.Tensor,
                 ldo: int = 0, w: Optional[torch.Tensor] = None, G: int = 1, seg_end: Optional[torch.Tensor] = None,
                 seg_group: Optional[torch.Tensor] = None, defer_status: bool = False):
    """Selection and build in ONE pass over the data (`mr_ties_select_build`: build with a provisional cut while the
    keys around it are collected, finish the exact cut, rebuild the few mis-decided columns) -- bit-identical to
    `select_kth_largest` + `_build` for `MR_TIES_VECTORS` / `MR_TIES_FUSED_MERGE`.  Returns the cut keys; if a model's
    sampled bracket missed (adversarial inputs) the exact select and the plain build run instead.
    `defer_status=True`: no host synchronisation (CUDA-graph capturable); returns `(cut, status)` and the caller must
    pass `status` to :func:`verify_select_status` before trusting `out`."""
    lib = _lib.load()
    K, d = len(rows), base_model.numel()
    dev = base_model.device
    P = 1 if seg_end is None else seg_end.numel()
    ws_bytes = int(lib.mr_ties_workspace_bytes(d, K))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    cut = torch.empty(K, dtype=torch.int64, device=dev)
    status = torch.zeros(K, dtype=torch.int32, device=dev)
    parr = _lib.ptr_array(rows)
    _lib.check(lib.mr_ties_select_build(_lib.dptr(base_model, torch.float32), parr, K, d, int(k_cnt), mode, _lib.dptr(w), G,
                                        _lib.dptr(seg_end), _lib.dptr(seg_group), P, _lib.dptr(out), ldo, _lib.dptr(cut),
                                        _lib.dptr(status), _lib.dptr(ws), ws_bytes, _lib.stream_handle()),
               "mr_ties_select_build")
    if defer_status:
        return cut, status
    if bool((status.cpu() != 1).any()):
        args = (_lib.dptr(base_model, torch.float32), parr, K, d, None, int(k_cnt), _lib.dptr(cut), _lib.dptr(status),
                _lib.dptr(ws), ws_bytes, _lib.stream_handle())
        _lib.check(lib.mr_ties_select_exact(*args), "mr_ties_select_exact")
        st = status.cpu()
        if bool((st != 1).any()):
            raise _lib.MergeRecLibraryError(f"TIES select failed with status {st.tolist()}")
        _build(base_model, rows, cut, mode, w=w, G=G, seg_end=seg_end, seg_group=seg_group, out=out, ldo=ldo)
    return cut


def get_ties_vectors(base_model: FlattenedModel, models: List[FlattenedModel], density: float,
                     return_masks: bool = False, **__):
    """(K, d) TIES vectors: per model the trimmed update where it agrees with the elected sign, divided by the
    per-column count of such survivors, so that ``sum_k lambda_k * That[k]`` is the lambda-weighted TIES delta
    (ties.py:55-72).  ``return_masks=True`` (extension) also returns the (K, d) bool trim and elect masks and
    the cut keys."""
    rows = as_rows(models)
    K, d = len(rows), base_model.numel()
    That = alloc_rows(K, d, base_model.device)
    if not return_masks:
        # one pass over the data: speculative build + exact cut + fix-up of the mis-decided columns
        select_build(base_model, rows, ties_topk_count(density, d), _lib.MR_TIES_VECTORS, That, ldo=max(That.stride(0), d))
        return That
    cut = ties_select(base_model, rows, density)
    trim = elect = None
    if return_masks:
        trim = torch.empty((K, d), dtype=torch.uint8, device=base_model.device)
        elect = torch.empty((K, d), dtype=torch.uint8, device=base_model.device)
    _build(base_model, rows, cut, _lib.MR_TIES_VECTORS, out=That, ldo=max(That.stride(0), d),
           trim_mask=trim, elect_mask=elect)
    if return_masks:
        return That, trim.bool(), elect.bool(), cut
    return That


def merge_ties(base_model: FlattenedModel, models: List[FlattenedModel], weights: List[float], density: float, **__
               ) -> FlattenedModel:
    """``base + sum_k trim_k(weights[k] * (models[k] - base))`` -- the weight is applied BEFORE the trim and there
    is no sign election or mean, exactly like the reference function of this name (ties.py:75-83)."""
    assert len(models) == len(weights), "Number of models and weights should match."
    rows = as_rows(models)
    dev = base_model.device
    # `update *= weights[i]` only runs for a non-empty list (ties.py:20); K >= 1 here so it always does
    w = torch.tensor([float(x) for x in weights], dtype=torch.float32, device=dev)
    cut = ties_select(base_model, rows, density, w)
    out = torch.empty_like(base_model)
    _build(base_model, rows, cut, _lib.MR_TIES_TRIMSUM, w=w, out=out)
    return out


def merge_ties_lambda(base_model: FlattenedModel, models: List[FlattenedModel], density: float, w: torch.Tensor,
                      seg_end: Optional[torch.Tensor] = None, seg_group: Optional[torch.Tensor] = None,
                      cut: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                      one_pass: bool = True) -> FlattenedModel:
    """Extension: ``get_ties_vectors`` followed by the (task- or layer-wise) lambda merge in ONE pass that never
    materialises the (K, d) TIES vectors.  Bit-identical to
    ``base + (w[g][:, None] * get_ties_vectors(...)).sum(0)`` evaluated block by block like
    weight_learning/module/layer_wise.py:76-82.  ``w`` is (G, K) fp32 on the device."""
    rows = as_rows(models)
    K = len(rows)
    if out is None:
        out = torch.empty_like(base_model)
    assert w.dtype == torch.float32 and w.is_contiguous() and w.shape[-1] == K
    if cut is None and one_pass:
        # one pass over the data (select folded into the fused build).  Measured on K = 8 BLaIR-base: 2.3 ms against
        # 2.46 ms for select + fused build in the same run (`one_pass=False`); this mode's pass is instruction-bound
        # (1.1 G warp-instructions), not HBM-bound
        select_build(base_model, rows, ties_topk_count(density, base_model.numel()), _lib.MR_TIES_FUSED_MERGE, out, w=w,
                     G=w.numel() // K, seg_end=seg_end, seg_group=seg_group)
        return out
    if cut is None:
        cut = ties_select(base_model, rows, density)
    _build(base_model, rows, cut, _lib.MR_TIES_FUSED_MERGE, w=w, G=w.numel() // K, seg_end=seg_end,
           seg_group=seg_group, out=out)
    return out
