"""Collaborative-merging distillation step (reference: rec_retrieval/module/distiller/sequence/module.py:16-108,
teacher logits: merge_train.py:116-126).

``DistillSequenceModule`` keeps the reference's constructor, ``forward`` dispatch, ``_forward_distill`` and
``configure_optimizers``; Lightning's trainer hooks are out of scope, so it is a plain ``nn.Module``.  The per-sample
python loop of the reference (one GEMV, one host->device teacher-row copy and ~10 small kernels per sample) is replaced
by three CUDA launches per step (csrc/distill.cu): catalogue logits for all samples with every domain's item table
read once, the loss + d loss/d logits of every sample, and -- in backward -- the gradient w.r.t. the representations
from a second pass over the tables.  Teacher logits either come from device-resident score matrices
(``make_score_embeddings``: what merge_train.py builds on the CPU) or are recomputed per step from the normalised
teacher embeddings (``TeacherScores(on_the_fly=True)``), which never materialises the (num_sequences, num_items)
matrices."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, List, Literal, Optional, Sequence

import torch
from torch import nn

from .... import _lib
from ...recommender.loss_fn import DistillLossBase, LossSpec, launch_distill_loss

__all__ = ["DistillSequenceModule", "BatchDistillationSequence", "TeacherScores", "make_score_embeddings",
           "normalize_rows", "distill_logits", "fused_distill_losses"]


@dataclass
class BatchDistillationSequence:
    """Fields `_forward_distill` reads (ref: rec_retrieval/types/model_batch.py, sequence/module.py:59-63)."""
    sequence: Any
    dataset_indexes: List[int]
    sequence_ids: List[int]


def normalize_rows(x: torch.Tensor) -> torch.Tensor:
    """``x / x.norm(dim=-1, keepdim=True)`` (merge_train.py:122-123) via ``mr_normalize_rows``."""
    dev = _lib.require_cuda()
    x = x.to(device=dev, dtype=torch.float32).contiguous()
    out = torch.empty_like(x)
    _lib.check(_lib.load().mr_normalize_rows(_lib.dptr(x), x.shape[0], x.shape[1], _lib.dptr(out), _lib.stream_handle()),
               "mr_normalize_rows")
    return out


def make_score_embeddings(item_embedding: torch.Tensor, sequence_embedding: torch.Tensor) -> torch.Tensor:
    """Teacher logits of one domain, device resident: normalise both tables, then every (sequence, item) dot product
    in plain fp32 (``mr_scores_fp32``).  ref: merge_train.py:116-126."""
    items = normalize_rows(item_embedding)
    seqs = normalize_rows(sequence_embedding)
    Q, E = seqs.shape
    N = items.shape[0]
    out = torch.empty((Q, N), dtype=torch.float32, device=items.device)
    _lib.check(_lib.load().mr_scores_fp32(_lib.dptr(seqs), Q, _lib.dptr(items), N, E, _lib.dptr(out), N, _lib.stream_handle()),
               "mr_scores_fp32")
    return out


def _tables(item_embeddings: Sequence[torch.Tensor]):
    nD = len(item_embeddings)
    ptrs = (C.c_void_p * nD)()
    rows = (C.c_int64 * nD)()
    E = None
    for d, t in enumerate(item_embeddings):
        if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or not t.is_contiguous():
            raise _lib.MergeRecLibraryError("item tables must be contiguous fp32 CUDA matrices (no CPU path)")
        if E is None:
            E = t.shape[1]
        elif t.shape[1] != E:
            raise ValueError("item tables disagree on the embedding dimension")
        ptrs[d] = t.data_ptr()
        rows[d] = t.shape[0]
    return ptrs, rows, nD, E


def _ld(item_embeddings, dataset_indexes) -> int:
    for b, d in enumerate(dataset_indexes):
        if not 0 <= int(d) < len(item_embeddings):
            raise ValueError(f"sample {b} has dataset index {d} outside [0, {len(item_embeddings)})")
    n = max(item_embeddings[d].shape[0] for d in dataset_indexes)
    return (n + 3) // 4 * 4


def distill_logits(rep: torch.Tensor, item_embeddings: Sequence[torch.Tensor], dataset_indexes: Sequence[int],
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """logits[b, :N_d] = rep[b] @ item_embeddings[d].T with d = dataset_indexes[b]  (sequence/module.py:64-66).
    Row b is defined on its first N_d columns only."""
    lib = _lib.load()
    B, E = rep.shape
    ptrs, rows, nD, Et = _tables(item_embeddings)
    if Et != E:
        raise ValueError(f"representation dim {E} != item table dim {Et}")
    ld = _ld(item_embeddings, dataset_indexes)
    if out is None:
        out = torch.empty((B, ld), dtype=torch.float32, device=rep.device)
    for b0 in range(0, B, _lib.MR_DISTILL_MAX_B // 2):   # <= 64 samples per call keeps the group count legal
        nb = min(_lib.MR_DISTILL_MAX_B // 2, B - b0)
        dom = (C.c_int32 * nb)(*[int(d) for d in dataset_indexes[b0:b0 + nb]])
        rc = lib.mr_distill_logits(C.c_void_p(rep[b0].data_ptr()), nb, E, ptrs, rows, nD, dom, C.c_void_p(out[b0].data_ptr()),
                                   out.stride(0), _lib.stream_handle())
        _lib.check(rc, "mr_distill_logits")
    return out


class TeacherScores:
    """Single-domain ("teacher") logits for the distillation step.

    * materialised (default): ``scores[d]`` is the device-resident (num_sequences_d, num_items_d) matrix the reference
      builds on the CPU (merge_train.py:126) -- rows are handed to the loss kernel as pointers, no copies;
    * ``on_the_fly=True``: keep only the normalised embeddings and recompute the B needed rows each step with the same
      table-streaming kernel that produces the merged model's logits."""

    def __init__(self, item_embeddings: Sequence[torch.Tensor], sequence_embeddings: Sequence[torch.Tensor],
                 on_the_fly: bool = False):
        self.on_the_fly = on_the_fly
        if on_the_fly:
            self.items = [normalize_rows(t) for t in item_embeddings]
            self.sequences = [normalize_rows(t) for t in sequence_embeddings]
            self.scores = None
        else:
            self.scores = [make_score_embeddings(i, s) for i, s in zip(item_embeddings, sequence_embeddings)]

    # teacher matrices above this many bytes (all domains together) stay in pinned host memory, like the reference keeps
    # them on the CPU (merge_train.py:126), and only the B rows of a step are copied to the device
    MAX_DEVICE_BYTES = 32 << 30

    @classmethod
    def from_scores(cls, score_embeddings: Sequence[torch.Tensor], max_device_bytes: Optional[int] = None) -> "TeacherScores":
        """Wrap precomputed (num_sequences_d, num_items_d) teacher matrices.  Up to `max_device_bytes` in total they are
        moved to HBM (rows are then handed to the loss kernel as pointers, no copies); larger sets stay on the host in
        pinned memory and `rows` stages the B needed rows per step."""
        self = cls.__new__(cls)
        self.on_the_fly = False
        dev = _lib.require_cuda()
        limit = cls.MAX_DEVICE_BYTES if max_device_bytes is None else int(max_device_bytes)
        total = sum(int(s.numel()) * 4 for s in score_embeddings)
        if total <= limit:
            self.scores = [s.to(device=dev, dtype=torch.float32).contiguous() for s in score_embeddings]
        else:
            self.scores = []
            for s in score_embeddings:
                h = s.detach().to(device="cpu", dtype=torch.float32).contiguous()
                self.scores.append(h if h.is_pinned() else h.pin_memory())
        return self

    def width(self, d: int) -> int:
        """Number of item columns of domain d's teacher rows."""
        return int(self.items[d].shape[0]) if self.on_the_fly else int(self.scores[d].shape[1])

    def rows(self, dataset_indexes: Sequence[int], sequence_ids: Sequence[int],
             item_embeddings: Optional[Sequence[torch.Tensor]] = None):
        """(keep-alive tensor(s), device addresses of the B teacher rows).  With `item_embeddings` given, every teacher
        row is checked to be exactly as wide as the item table the loss kernel will read it against (a narrower teacher
        matrix would be read out of bounds): ValueError otherwise."""
        if item_embeddings is not None:
            for d in set(int(x) for x in dataset_indexes):
                if not 0 <= d < len(item_embeddings):
                    raise ValueError(f"dataset index {d} outside [0, {len(item_embeddings)})")
                if self.width(d) != int(item_embeddings[d].shape[0]):
                    raise ValueError(f"teacher scores of dataset {d} have {self.width(d)} columns but its item table has "
                                     f"{int(item_embeddings[d].shape[0])} rows")
        if not self.on_the_fly:
            for d, s in zip(dataset_indexes, sequence_ids):
                if not 0 <= s < self.scores[d].shape[0]:
                    raise IndexError(f"sequence id {s} outside the teacher matrix of dataset {d}")
            if self.scores and not self.scores[0].is_cuda:
                # host-resident matrices: stage the B rows (one pinned -> device copy per row, stream-ordered)
                dev = _lib.require_cuda()
                ld = (max(self.scores[d].shape[1] for d in dataset_indexes) + 3) // 4 * 4
                t = torch.empty((len(dataset_indexes), ld), dtype=torch.float32, device=dev)
                for b, (d, s) in enumerate(zip(dataset_indexes, sequence_ids)):
                    row = self.scores[d][s]
                    t[b, :row.numel()].copy_(row, non_blocking=True)
                return t, [t.data_ptr() + b * t.stride(0) * 4 for b in range(t.shape[0])]
            ptrs = []
            for d, s in zip(dataset_indexes, sequence_ids):
                m = self.scores[d]
                ptrs.append(m.data_ptr() + int(s) * m.stride(0) * 4)
            return None, ptrs
        rep = torch.stack([self.sequences[d][s] for d, s in zip(dataset_indexes, sequence_ids)])
        t = distill_logits(rep, self.items, dataset_indexes)
        return t, [t.data_ptr() + b * t.stride(0) * 4 for b in range(t.shape[0])]


class _FusedDistill(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rep, item_embeddings, dataset_indexes, teacher_ptrs, spec):
        rep = rep.contiguous()
        logits = distill_logits(rep, item_embeddings, dataset_indexes)
        n = [item_embeddings[d].shape[0] for d in dataset_indexes]
        loss, gz = launch_distill_loss(logits, teacher_ptrs, n, spec, want_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(gz)
        ctx.tables, ctx.dataset_indexes, ctx.E = item_embeddings, list(dataset_indexes), rep.shape[1]
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (gz,) = ctx.saved_tensors
        lib = _lib.load()
        B, E = gz.shape[0], ctx.E
        ptrs, rows, nD, _ = _tables(ctx.tables)
        grad_out = grad_out.to(torch.float32).contiguous()
        grad_rep = torch.empty((B, E), dtype=torch.float32, device=gz.device)
        ws_bytes = int(lib.mr_distill_grad_workspace_bytes(E))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=gz.device)
        step = _lib.MR_DISTILL_MAX_B // 2
        for b0 in range(0, B, step):
            nb = min(step, B - b0)
            dom = (C.c_int32 * nb)(*[int(d) for d in ctx.dataset_indexes[b0:b0 + nb]])
            rc = lib.mr_distill_grad(C.c_void_p(gz[b0].data_ptr()), gz.stride(0), C.c_void_p(grad_out[b0:].data_ptr()), nb, E,
                                     ptrs, rows, nD, dom, C.c_void_p(grad_rep[b0].data_ptr()), _lib.dptr(ws), ws_bytes,
                                     _lib.stream_handle())
            _lib.check(rc, "mr_distill_grad")
        return grad_rep, None, None, None, None


def fused_distill_losses(rep: torch.Tensor, item_embeddings: Sequence[torch.Tensor], dataset_indexes: Sequence[int],
                         teacher_ptrs: Optional[Sequence[int]], spec: LossSpec) -> torch.Tensor:
    """Per-sample distillation losses (B,) -- the body of the reference loop, sequence/module.py:63-71."""
    _lib.require_cuda()
    if rep.dim() != 2 or rep.dtype != torch.float32 or not rep.is_cuda:
        raise _lib.MergeRecLibraryError("representations must be a (B, E) fp32 CUDA tensor (no CPU path)")
    if len(dataset_indexes) != rep.shape[0]:
        raise ValueError("one dataset index per sample expected")
    return _FusedDistill.apply(rep, list(item_embeddings), list(dataset_indexes), teacher_ptrs, spec)


class DistillSequenceModule(nn.Module):
    """ref: sequence/module.py:16-108 (same constructor arguments and attributes)."""

    def __init__(self, merged_model: nn.Module, score_embeddings, loss_fn: DistillLossBase,
                 similarity: Literal["dot", "cosine"], learning_rate: float = 5e-5, trainable_args_kwargs: dict | None = None):
        super().__init__()
        self.merged_model = merged_model
        self.loss_fn = loss_fn
        self.similarity = similarity
        self.learning_rate = learning_rate
        self.trainable_args_kwargs = trainable_args_kwargs or {}
        self.score_embeddings = score_embeddings if isinstance(score_embeddings, TeacherScores) \
            else TeacherScores.from_scores(score_embeddings)
        self.item_embeddings = None   # list of (num_items_d, E) device tensors, injected like callbacks.py:85-110 does
        self._valid_metrics = []

    def _maybe_normalize(self, matrix: torch.Tensor):
        if self.similarity == "cosine":
            return nn.functional.normalize(matrix, p=2, dim=-1)
        return matrix

    def forward(self, batch):
        if isinstance(batch, BatchDistillationSequence) or hasattr(batch, "dataset_indexes"):
            return self._forward_distill(batch)
        if hasattr(batch, "sequence"):
            return self._forward_sequence_encoding(batch.sequence)
        if hasattr(batch, "items"):
            return self._forward_sequence_encoding(batch.items)
        raise ValueError(f"Invalid batch type {type(batch)}")

    def _forward_sequence_encoding(self, sequence_batch):
        return self._maybe_normalize(self.merged_model.forward(sequence_batch))

    def _forward_distill(self, batch):
        if self.item_embeddings is None:
            raise RuntimeError("item_embeddings have not been injected")
        rep = self._forward_sequence_encoding(batch.sequence).to(torch.float32)
        spec = self.loss_fn.spec
        keep, ptrs = (None, None)
        if spec.needs_teacher:
            keep, ptrs = self.score_embeddings.rows(batch.dataset_indexes, batch.sequence_ids, self.item_embeddings)
        losses = fused_distill_losses(rep, self.item_embeddings, batch.dataset_indexes, ptrs, spec)
        del keep
        return losses.mean()

    def training_step(self, batch, batch_idx: int = 0):
        return self._forward_distill(batch)

    def validation_step(self, batch, batch_idx: int = 0, dataloader_idx: int = 0):
        loss = self._forward_distill(batch)
        self._valid_metrics.append(loss.item())
        return loss

    def configure_optimizers(self):
        return torch.optim.Adam(self.merged_model.trainable_parameters(**self.trainable_args_kwargs),
                                lr=self.learning_rate, weight_decay=0.0)
