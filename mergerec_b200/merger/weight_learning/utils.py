"""Functional-parameter plumbing of the collaborative-merging module
(reference: rec_retrieval/merger/weight_learning/utils.py)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from ..layout import FlatLayout
from ..types import ShapeDict, StateDict


def del_attr(obj, names: Sequence[str]) -> None:
    for n in names[:-1]:
        obj = getattr(obj, n)
    delattr(obj, names[-1])


def set_attr(obj, names: Sequence[str], val: torch.Tensor, shape=None) -> None:
    """Plant ``val`` (reshaped to ``shape`` when given) at the dotted attribute path (utils.py:11-15)."""
    for n in names[:-1]:
        obj = getattr(obj, n)
    setattr(obj, names[-1], val if shape is None else val.reshape(shape))


def make_functional(mod: torch.nn.Module) -> Tuple[Tuple[torch.Tensor, ...], List[str]]:
    """Delete every nn.Parameter of ``mod`` (they are re-planted as views of the merged flat vector before each
    forward) and return their former values and names (utils.py:18-26)."""
    orig, names = [], []
    for name, p in list(mod.named_parameters()):
        orig.append(p.data.clone())
        names.append(name)
        del_attr(mod, name.split("."))
    return tuple(orig), names


def get_state_dict(params: torch.Tensor, shape_dict: ShapeDict) -> StateDict:
    """Slice + reshape views of a flat vector (utils.py:29-40); AssertionError when sizes disagree."""
    layout = FlatLayout.from_shape_dict(shape_dict)
    assert layout.d == len(params), "Not all parameters are loaded."
    return layout.views(params)


def load_weights(mod: torch.nn.Module, params, shape_dict: ShapeDict) -> None:
    """Plant one tensor per state_dict entry into ``mod`` (utils.py:43-51).  ``params`` is a flat (d) vector or
    an already split dict/sequence of per-tensor views."""
    if isinstance(params, torch.Tensor):
        params = get_state_dict(params, shape_dict)
    if isinstance(params, dict):
        params = [params[k] for k in shape_dict.keys()]
    assert len(params) == len(shape_dict), "Not all parameters are loaded."
    for (name, _), p in zip(shape_dict.items(), params):
        set_attr(mod, name.split("."), p)
