"""Distillation losses (reference: rec_retrieval/module/recommender/loss_fn.py:20-280; same class names, constructor
arguments and ``forward(merged_model_logits, single_model_logits)`` contract).  Every class is a thin description
(`spec`) of one mix handled by ``mr_distill_loss`` (csrc/distill.cu): the loss of each logits row and its gradient
come out of one CUDA kernel launch for the whole batch of rows; ``forward`` returns the mean over rows, which is what
every reference loss reduces to for equally long rows (CE / entropy / hinge / ListNet ``.mean()``, ``kl_div(...,
"batchmean")``, ``mse_loss(..., "mean")``).  No CPU path: the tensors must live on the GPU."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch
from torch import nn

from ... import _lib
from ...merger.enums import LossType

__all__ = [
    "DistillLossBase", "DistillCELoss", "DistillKDLoss", "DistillMSELoss", "DistillPairwiseLoss", "DistillListNetLoss",
    "DistillAdaMergingLoss", "DistillAdaMergingKDLoss", "MergedPseudoLabelLoss", "MergedPseudoLabelKDLoss",
    "SinglePseudoLabelLoss", "SinglePseudoLabelKDLoss", "distill_loss_factory", "LossSpec", "distill_loss_rows",
]

(MR_LOSS_CE, MR_LOSS_KD, MR_LOSS_MSE, MR_LOSS_ADAMERGING, MR_LOSS_ADAMERGING_KD, MR_LOSS_MERGED_PSEUDO_LABEL,
 MR_LOSS_MERGED_PSEUDO_LABEL_KD, MR_LOSS_SINGLE_PSEUDO_LABEL, MR_LOSS_SINGLE_PSEUDO_LABEL_KD, MR_LOSS_PAIRWISE,
 MR_LOSS_LISTNET) = range(11)
_NO_TEACHER = (MR_LOSS_ADAMERGING, MR_LOSS_MERGED_PSEUDO_LABEL)


@dataclass(frozen=True)
class LossSpec:
    """What ``mr_distill_loss`` needs to know about a loss object."""
    loss_type: int
    temperature: float = 1.0
    coefficient: float = 0.0
    margin: float = 0.0

    @property
    def needs_teacher(self) -> bool:
        return self.loss_type not in _NO_TEACHER


def launch_distill_loss(logits: torch.Tensor, teacher_ptrs: Optional[Sequence[int]], n_per_sample: Sequence[int],
                        spec: LossSpec, want_grad: bool):
    """One ``mr_distill_loss`` launch per <= 128 rows.  logits (B, ld) fp32 CUDA, row b valid on [:n_b];
    teacher_ptrs = device addresses of the B teacher rows.  Returns (loss (B,), grad_logits (B, ld) or None)."""
    lib = _lib.load()
    B, ld = logits.shape
    loss = torch.empty(B, dtype=torch.float32, device=logits.device)
    gz = torch.empty_like(logits) if want_grad else None
    for b0 in range(0, B, _lib.MR_DISTILL_MAX_B):
        nb = min(_lib.MR_DISTILL_MAX_B, B - b0)
        tarr = None
        if teacher_ptrs is not None:
            tarr = (C.c_void_p * nb)(*[int(p) for p in teacher_ptrs[b0:b0 + nb]])
        narr = (C.c_int64 * nb)(*[int(n) for n in n_per_sample[b0:b0 + nb]])
        rc = lib.mr_distill_loss(C.c_void_p(logits[b0].data_ptr()), logits.stride(0), tarr, narr, nb, spec.loss_type,
                                 float(spec.temperature), float(spec.coefficient), float(spec.margin),
                                 C.c_void_p(loss[b0:].data_ptr()),
                                 C.c_void_p(gz[b0].data_ptr()) if want_grad else None, gz.stride(0) if want_grad else 0,
                                 _lib.stream_handle())
        _lib.check(rc, "mr_distill_loss")
    return loss, gz


class _LossRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, merged: torch.Tensor, single: Optional[torch.Tensor], spec: LossSpec):
        R, N = merged.shape
        ptrs = None
        if single is not None:
            ptrs = [single.data_ptr() + r * single.stride(0) * 4 for r in range(R)]
        loss, gz = launch_distill_loss(merged, ptrs, [N] * R, spec, want_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(gz)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (gz,) = ctx.saved_tensors
        return gz * grad_out.unsqueeze(1), None, None


def distill_loss_rows(merged_model_logits: torch.Tensor, single_model_logits: Optional[torch.Tensor], spec: LossSpec) -> torch.Tensor:
    """Per-row losses (R,) of (R, N) merged-model logits against (R, N) teacher logits."""
    _lib.require_cuda()
    if merged_model_logits.dim() != 2:
        raise ValueError("expected (rows, num_items) logits")
    merged = merged_model_logits.to(torch.float32).contiguous()
    single = None
    if spec.needs_teacher:
        if single_model_logits is None or single_model_logits.shape != merged_model_logits.shape:
            raise ValueError("single_model_logits must have the shape of merged_model_logits")
        single = single_model_logits.detach().to(device=merged.device, dtype=torch.float32).contiguous()
    return _LossRows.apply(merged, single, spec)


class DistillLossBase(nn.Module):
    """ref: loss_fn.py:20-33.  Subclasses set ``spec``."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    @property
    def spec(self) -> LossSpec:
        raise NotImplementedError("Subclasses should implement this method.")

    def forward(self, merged_model_logits: torch.Tensor, single_model_logits: torch.Tensor):
        return distill_loss_rows(merged_model_logits, single_model_logits, self.spec).mean()


class DistillCELoss(DistillLossBase):
    """CE against the teacher's argmax (ref: loss_fn.py:36-43)."""
    spec = LossSpec(MR_LOSS_CE)


class DistillKDLoss(DistillLossBase):
    """T^2 * KL(softmax(t/T) || softmax(z/T)) (ref: loss_fn.py:46-59)."""

    def __init__(self, temperature: float):
        super().__init__()
        self.temperature = temperature

    @property
    def spec(self):
        return LossSpec(MR_LOSS_KD, temperature=self.temperature)


class DistillAdaMergingLoss(DistillLossBase):
    """Entropy of softmax(z) with the +1e-8 inside the log (ref: loss_fn.py:62-68)."""
    spec = LossSpec(MR_LOSS_ADAMERGING)


class DistillAdaMergingKDLoss(DistillLossBase):
    """entropy + coefficient * KD (ref: loss_fn.py:71-87)."""

    def __init__(self, temperature: float, coefficient: float):
        super().__init__()
        self.temperature = temperature
        self.coefficient = coefficient

    @property
    def spec(self):
        return LossSpec(MR_LOSS_ADAMERGING_KD, temperature=self.temperature, coefficient=self.coefficient)


class MergedPseudoLabelLoss(DistillLossBase):
    """CE against the merged model's own argmax (ref: loss_fn.py:90-105)."""
    spec = LossSpec(MR_LOSS_MERGED_PSEUDO_LABEL)


class MergedPseudoLabelKDLoss(DistillKDLoss):
    """ref: loss_fn.py:108-129."""

    def __init__(self, temperature: float, coefficient: float):
        super().__init__(temperature)
        self.coefficient = coefficient

    @property
    def spec(self):
        return LossSpec(MR_LOSS_MERGED_PSEUDO_LABEL_KD, temperature=self.temperature, coefficient=self.coefficient)


class SinglePseudoLabelLoss(DistillLossBase):
    """CE against the teacher's argmax (ref: loss_fn.py:132-153)."""
    spec = LossSpec(MR_LOSS_SINGLE_PSEUDO_LABEL)


class SinglePseudoLabelKDLoss(DistillKDLoss):
    """ref: loss_fn.py:156-177."""

    def __init__(self, temperature: float, coefficient: float):
        super().__init__(temperature)
        self.coefficient = coefficient

    @property
    def spec(self):
        return LossSpec(MR_LOSS_SINGLE_PSEUDO_LABEL_KD, temperature=self.temperature, coefficient=self.coefficient)


class DistillMSELoss(DistillLossBase):
    """ref: loss_fn.py:180-187."""
    spec = LossSpec(MR_LOSS_MSE)


class DistillPairwiseLoss(DistillLossBase):
    """Hinge on the teacher's best vs second-best item (ref: loss_fn.py:190-211)."""

    def __init__(self, margin: float):
        super().__init__()
        self.margin = margin

    @property
    def spec(self):
        return LossSpec(MR_LOSS_PAIRWISE, margin=self.margin)


class DistillListNetLoss(DistillLossBase):
    """ref: loss_fn.py:214-229 (``eps`` is accepted and unused there too)."""

    def __init__(self, temperature: float, eps: float = 1e-8):
        super().__init__()
        self.temperature = temperature
        self.eps = eps

    @property
    def spec(self):
        return LossSpec(MR_LOSS_LISTNET, temperature=self.temperature)


def distill_loss_factory(loss_type: LossType, temperature: float | None = None, **kwargs) -> DistillLossBase:
    """ref: loss_fn.py:233-280 -- same cases, same error messages."""
    name = loss_type.name if isinstance(loss_type, LossType) else None
    if name == "CE":
        return DistillCELoss(**kwargs)
    if name == "KD":
        if temperature is None:
            raise ValueError("Temperature must be provided for KDLoss.")
        return DistillKDLoss(temperature)
    if name == "MSE":
        return DistillMSELoss(**kwargs)
    if name == "ADAMERGING":
        return DistillAdaMergingLoss(**kwargs)
    if name == "ADAMERGING_KD":
        if temperature is None:
            raise ValueError("Temperature must be provided for AdaMergingKDLoss.")
        if "coefficient" not in kwargs:
            raise ValueError("Coefficient must be provided for AdaMergingKDLoss.")
        return DistillAdaMergingKDLoss(temperature, kwargs["coefficient"])
    if name == "MERGED_PSEUDO_LABEL":
        return MergedPseudoLabelLoss()
    if name == "MERGED_PSEUDO_LABEL_KD":
        if temperature is None:
            raise ValueError("Temperature must be provided for MergedPseudoLabelKDLoss.")
        if "coefficient" not in kwargs:
            raise ValueError("Coefficient must be provided for MergedPseudoLabelKDLoss.")
        return MergedPseudoLabelKDLoss(temperature, kwargs["coefficient"])
    if name == "SINGLE_PSEUDO_LABEL":
        return SinglePseudoLabelLoss()
    if name == "SINGLE_PSEUDO_LABEL_KD":
        if temperature is None:
            raise ValueError("Temperature must be provided for SinglePseudoLabelKDLoss.")
        if "coefficient" not in kwargs:
            raise ValueError("Coefficient must be provided for SinglePseudoLabelKDLoss.")
        return SinglePseudoLabelKDLoss(temperature, kwargs["coefficient"])
    raise ValueError(f"Unknown loss type: {loss_type}")
