"""state_dict <-> flat vector plumbing (A0 of SURVEY.md section 8(a)).

Same functions as the reference's rec_retrieval/merger/utils/model_operations.py, except that
``flatten_model`` lands the flat vector directly in HBM (one device allocation, one copy per tensor, no host
``torch.cat``), because every consumer is a CUDA kernel.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple, Union

import torch

from ... import _lib
from ..layout import FlatLayout
from ..types import FlattenedModel, ShapeDict, StateDict

__all__ = ["check_model_shape", "flatten_model", "unflatten_model", "align_dict_key_order"]


def check_model_shape(models: Sequence[StateDict], base_model: Optional[StateDict] = None) -> None:
    """AssertionError unless all models (and the base) share key set and per-key shapes.
    reference: model_operations.py:15-44 (same messages)."""
    first = models[0]
    want = set(first.keys())
    for m in models:
        assert set(m.keys()) == want, "Models have different architectures."
    for name in want:
        for m in models[1:]:
            assert m[name].shape == first[name].shape, "Models have different shapes."
    if base_model is not None:
        assert set(base_model.keys()) == want, "Base model has different architecture from the others."
        for name in want:
            assert base_model[name].shape == first[name].shape, "Base model has different shapes."


def flatten_model(model: StateDict, device: Optional[torch.device] = None) -> Tuple[FlattenedModel, ShapeDict]:
    """Concatenate every tensor (dict order, reshape(-1)) into one fp32 CUDA vector.
    Integer buffers are promoted to fp32 exactly as ``torch.cat`` promotes them in the reference
    (model_operations.py:47-63)."""
    device = device or _lib.require_cuda()
    shape_dict = {k: v.shape for k, v in model.items()}
    layout = FlatLayout.from_shape_dict(shape_dict)
    if any(not t.is_cuda for t in model.values()):
        # host tensors: the streamed loader (pinned runs coalesced, pageable tensors packed through a pinned staging ring
        # on a copy stream); the current stream waits for its event, the host does not
        from .loader import default_loader
        flat, ready = default_loader(device).load(model, layout)
        torch.cuda.current_stream(device).wait_event(ready)
        return flat, shape_dict
    flat = torch.empty(layout.d, dtype=torch.float32, device=device)
    for t, off, n in zip(model.values(), layout.offsets, layout.sizes):
        if n:
            flat[off:off + n].copy_(t.detach().reshape(-1), non_blocking=True)
    return flat, shape_dict


def unflatten_model(model: FlattenedModel, shape_dict: ShapeDict) -> StateDict:
    """Views of the flat vector with the original shapes (model_operations.py:66-90)."""
    return FlatLayout.from_shape_dict(shape_dict).views(model)


def align_dict_key_order(*models: Union[StateDict, None], key_order: Optional[List[str]] = None
                         ) -> Iterable[Union[StateDict, None]]:
    """Re-key every dict into one common order: ``key_order`` if given, else the sorted keys of the first
    non-None dict (model_operations.py:93-136; ``None`` entries pass through)."""
    if key_order is None:
        first = next((m for m in models if isinstance(m, dict)), None)
        if first is None:
            raise ValueError("At least one model must be provided to infer the key order.")
        key_order = sorted(first.keys())
    wanted = set(key_order)
    out = []
    for m in models:
        if m is None:
            out.append(None)
            continue
        assert isinstance(m, dict), f"Model must be a dictionary, got {type(m)}."
        missing_keys, extra_keys = wanted - set(m.keys()), set(m.keys()) - wanted
        assert not missing_keys and not extra_keys, (
            f"All models must have the same set of keys. Missing keys: {missing_keys}, Extra keys: {extra_keys}")
        out.append({k: m[k] for k in key_order})
    return out
