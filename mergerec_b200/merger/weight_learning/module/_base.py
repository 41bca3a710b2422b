"""Collaborative-merging module (A3-A5, A11 of SURVEY.md section 8(a)).

Reference: rec_retrieval/merger/weight_learning/module/_base.py (+ task_wise.py, layer_wise.py).  Same
attributes and methods; the merge and its lambda-gradient run as CUDA kernels:

* forward  : ``merged = base + sum_k w[g,k] * T[k]`` in ONE streaming kernel (no (K,d) temporary), bit-identical
  to the reference's ``torch.sum(dim=0)`` order, written into one flat buffer whose slices become the
  encoder's parameters;
* backward : the P per-tensor gradients are read in place through a pointer table and reduced against T in one
  pass (``mr_lambda_grad``) -- no P x d SliceBackward temporaries, no (K,d) product.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Dict, List, Optional

import torch
from torch import nn

from .... import _lib
from ...algorithms._common import merge_axpy
from ...layout import FlatLayout, alloc_rows
from ..utils import load_weights


def _lambda_grad(grads: List[Optional[torch.Tensor]], layout: FlatLayout, T: torch.Tensor, seg_group: Optional[torch.Tensor],
                 G: int) -> torch.Tensor:
    """(G, K) = sum over tensors of <grad_p, T[k, off_p:off_p+n_p]> via mr_lambda_grad."""
    lib = _lib.load()
    dev = T.device
    K, d = T.shape
    keep = []  # keep contiguous copies alive until the kernel is queued
    ptrs = []
    for g in grads:
        if g is None:
            ptrs.append(0)
            continue
        if g.dtype != torch.float32 or not g.is_contiguous():
            g = g.to(torch.float32).contiguous()
        keep.append(g)
        ptrs.append(g.data_ptr())
    P = len(ptrs)
    # Pointer table and workspace are cached on the layout: a gradient set that repeats (steady-state allocator
    # addresses; always under CUDA-graph replay) uploads nothing, a new one goes up with ONE asynchronous copy from a
    # pinned host table that stays alive in the cache -- no blocking transfer, no allocation, so the whole backward
    # (and the collaborative step around it) can be captured in a CUDA graph.
    cache = layout._dev_cache.setdefault(("lambda_grad", str(dev)), {"tables": {}, "ws": None})
    key = tuple(ptrs)
    entry = cache["tables"].get(key)
    if entry is None:
        host = torch.tensor(ptrs, dtype=torch.int64).pin_memory()
        entry = (host, host.to(dev, non_blocking=True))
        if len(cache["tables"]) >= 16:
            cache["tables"].pop(next(iter(cache["tables"])))
        cache["tables"][key] = entry
    ptr_t = entry[1]
    seg_off, seg_len = layout.device_segments(dev)
    ws_bytes = int(lib.mr_lambda_grad_workspace_bytes(d, P, K))
    if cache["ws"] is None or cache["ws"].numel() < ws_bytes:
        cache["ws"] = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ws = cache["ws"]
    out = torch.empty((G, K), dtype=torch.float32, device=dev)
    rc = lib.mr_lambda_grad(_lib.dptr(ptr_t), _lib.dptr(seg_off), _lib.dptr(seg_len), _lib.dptr(seg_group), P, d,
                            _lib.dptr(T), T.stride(0), K, G, _lib.dptr(out), _lib.dptr(ws), ws_bytes,
                            _lib.stream_handle())
    _lib.check(rc, "mr_lambda_grad")
    del keep
    return out


class _MergedViews(torch.autograd.Function):
    """w (G,K) -> the P parameter views of one merged flat buffer.  Backward = the lambda-gradient reduction."""

    @staticmethod
    def forward(ctx, w: torch.Tensor, module: "TaskVectorMergingModuleBase"):
        merged = module._merge_flat(w.detach())
        ctx.module = module
        ctx.set_materialize_grads(False)
        views = module._layout.views(merged)
        return tuple(views.values())

    @staticmethod
    def backward(ctx, *grads):
        m = ctx.module
        _, seg_group, keys = m._blocks()
        gw = _lambda_grad(list(grads), m._layout, m._task_rows(), seg_group if len(keys) > 1 else None, len(keys))
        return gw, None


class _MergedFlat(torch.autograd.Function):
    """w (G,K) -> merged flat (d) vector (what ``_merge_task_vectors`` returns in the reference)."""

    @staticmethod
    def forward(ctx, w: torch.Tensor, module: "TaskVectorMergingModuleBase"):
        ctx.module = module
        return module._merge_flat(w.detach())

    @staticmethod
    def backward(ctx, grad):
        m = ctx.module
        _, seg_group, keys = m._blocks()
        grad = grad.to(torch.float32).contiguous()
        layout = m._layout
        grads = [grad[o:o + n] for o, n in zip(layout.offsets, layout.sizes)]
        gw = _lambda_grad(grads, layout, m._task_rows(), seg_group if len(keys) > 1 else None, len(keys))
        return gw, None


class TaskVectorMergingModuleBase(nn.Module, ABC):
    LAYER_WISE = False

    def __init__(self, base_model_tensor: torch.Tensor, task_vectors_tensor: torch.Tensor, model_without_params,
                 shape_dict: Dict[str, torch.Size], disable_softmax: bool = False):
        super().__init__()
        self.model = model_without_params
        self.shape_dict = shape_dict
        self.disable_softmax = disable_softmax
        # frozen base / task vectors (requires_grad=False), as in the reference (_base.py:26-27)
        self.base_model_tensor = nn.Parameter(base_model_tensor, requires_grad=False)
        self.task_vectors_tensor = nn.Parameter(task_vectors_tensor, requires_grad=False)
        self.global_weights = nn.ParameterDict()
        self.global_biases = nn.ParameterDict()
        self.per_weights = nn.ParameterDict()
        self._layout = FlatLayout.from_shape_dict(shape_dict)
        assert self._layout.d == base_model_tensor.numel(), "shape_dict does not match the flat base model."
        self._rows_cache = None

    # ---- reference API -------------------------------------------------------------------------------
    def trainable_parameters(self, freeze_global_weight: bool = False, freeze_global_bias: bool = False,
                             freeze_per_weight: bool = False):
        params = []
        if not freeze_global_weight:
            params.extend(self.global_weights.parameters())
        if not freeze_global_bias:
            params.extend(self.global_biases.parameters())
        if not freeze_per_weight:
            params.extend(self.per_weights.parameters())
        return params

    def serialize_weights(self):
        return {
            "global_weights": {k: v.tolist() for k, v in self.global_weights.items()},
            "global_biases": {k: v.tolist() for k, v in self.global_biases.items()},
            "per_weights": {k: v.tolist() for k, v in self.per_weights.items()},
        }

    @torch.no_grad()
    def load_weights_from_dict(self, weights):
        """Same schema and checks as the reference (_base.py:53-76); values land on the parameter's device."""
        for field, store, truncate in (("global_weights", self.global_weights, False),
                                       ("global_biases", self.global_biases, False),
                                       ("per_weights", self.per_weights, True)):
            for k, v in weights[field].items():
                assert k in store, f"Key '{k}' not found in {field}."
                v = torch.tensor(v)
                if truncate:
                    v = v[: store[k].numel()]  # the reference silently truncates per_weights (_base.py:72)
                assert v.shape == store[k].shape, f"Shape mismatch for key '{k}', ({v.shape} != {store[k].shape})"
                store[k].data = v.to(device=store[k].device, dtype=store[k].dtype)

    def forward(self, batch):
        self.load_weights()
        return self.model(batch)

    def load_weights(self):
        """Merge and plant the P parameter views into the wrapped model (task_wise.py:50-55)."""
        views = _MergedViews.apply(self._effective_weights(), self)
        load_weights(self.model, list(views), self.shape_dict)
        return self.model

    def get_state_dict(self):
        views = _MergedViews.apply(self._effective_weights(), self)
        return dict(zip(self.shape_dict.keys(), views))

    def _merge_task_vectors(self) -> torch.Tensor:
        """The merged flat (d) vector, differentiable w.r.t. the lambdas."""
        return _MergedFlat.apply(self._effective_weights(), self)

    # ---- internals -----------------------------------------------------------------------------------
    @abstractmethod
    def _group_keys(self) -> List[str]:
        raise NotImplementedError

    def _effective_weights(self) -> torch.Tensor:
        """(G, K): w_g = gw_g * (softmax(pw_g) unless disable_softmax) + gb_g, same op order as the reference
        (task_wise.py:37-42, layer_wise.py:67-74); stays in torch so autograd chains to gw / gb / pw."""
        rows = []
        for key in self._group_keys():
            pw = self.per_weights[key]
            if not self.disable_softmax:
                pw = torch.softmax(pw, dim=0)
            rows.append(self.global_weights[key] * pw + self.global_biases[key])
        return torch.stack(rows, dim=0)

    def _blocks(self):
        return self._layout.device_blocks(self.LAYER_WISE, self.base_model_tensor.device)

    def _task_rows(self) -> torch.Tensor:
        """(K, d) task vectors with 16-byte aligned rows (repacked once if the given tensor's rows are not)."""
        T = self.task_vectors_tensor.data
        if not T.is_cuda:
            raise _lib.MergeRecLibraryError("the merging module must live on a CUDA device (no CPU fallback)")
        key = (T.data_ptr(), T.stride(0))
        if self._rows_cache is None or self._rows_cache[0] != key:
            rows = T
            if T.dtype != torch.float32 or T.stride(1) != 1 or (T.stride(0) * 4) % 16 or T.data_ptr() % 16:
                rows = alloc_rows(T.shape[0], T.shape[1], T.device)
                rows.copy_(T)
            self._rows_cache = (key, rows)
        return self._rows_cache[1]

    def _merge_flat(self, w: torch.Tensor) -> torch.Tensor:
        base = self.base_model_tensor.data
        rows = self._task_rows()
        w = w.to(torch.float32).contiguous()
        seg_end, seg_group, keys = self._blocks()
        if len(keys) == 1 and not self.LAYER_WISE:
            return merge_axpy(base, list(rows.unbind(0)), w, _lib.MR_ORDER_SUM_FIRST, src_is_model=False)
        return merge_axpy(base, list(rows.unbind(0)), w, _lib.MR_ORDER_SUM_FIRST, src_is_model=False, seg_end=seg_end,
                          seg_group=seg_group)
