"""``ModelMerger`` -- the training-free merge front door (reference: rec_retrieval/merger/merger.py:10-107).

Same constructor, attributes and ``merge`` contract as the reference; the flat vectors live in HBM and every
merge is one CUDA kernel launch through the C ABI.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch

from .algorithms import merge_dare, merge_linear, merge_pcb, merge_task_vector, merge_ties
from .types import FlattenedModel, ShapeDict, StateDict
from .utils.model_operations import align_dict_key_order, check_model_shape, flatten_model, unflatten_model

_NEEDS_BASE = {"task_vector": "Task vector", "ties": "TIES", "dare": "DARE", "pcb": "PCB"}


class ModelMerger:
    def __init__(self, models: Sequence[StateDict], base_model: Optional[StateDict] = None,
                 align_key_order: bool = True):
        """Validate K state_dicts (+ optional base), fix a common key order and flatten each into one fp32
        vector on the GPU.  ``align_key_order=False`` only checks that the orders already agree
        (merger.py:25-30).  Without a base model the first model supplies ``shape_dict`` (merger.py:33-38)."""
        check_model_shape(models, base_model)
        if align_key_order:
            *models, base_model = align_dict_key_order(*models, base_model)
        else:
            assert self._keys_are_aligned(models, base_model), "Model keys are not aligned."

        self.models: List[FlattenedModel] = []
        self.base_model: Optional[FlattenedModel] = None
        self.shape_dict: ShapeDict
        if base_model is None:
            first, self.shape_dict = flatten_model(models[0])
            self.models.append(first)
            models = models[1:]
        else:
            self.base_model, self.shape_dict = flatten_model(base_model)
        self.models.extend(flatten_model(m)[0] for m in models)

    @torch.no_grad()
    def merge(self, merge_type: str, weights: Union[Sequence[float], float], **kwargs) -> StateDict:
        """Run one named merge and return the merged state_dict (views of one flat CUDA vector).

        ``weights`` is a float (applied to every model) or a list of floats -- anything else raises
        ``ValueError`` exactly like the reference (merger.py:60-64).  Unknown merge types raise ``ValueError``
        (merger.py:87-88); ``"dare"`` draws its keep masks from torch's CUDA generator unless ``masks=`` is passed."""
        if isinstance(weights, float):
            weights = [weights] * len(self.models)
        elif not (isinstance(weights, list) and all(isinstance(w, float) for w in weights)):
            raise ValueError("Weights should be a float or a list of floats.")

        if merge_type in _NEEDS_BASE and self.base_model is None:
            raise ValueError(f"{_NEEDS_BASE[merge_type]} merge requires a base model.")
        call = dict(models=self.models, base_model=self.base_model, weights=weights, **kwargs)
        if merge_type == "linear":
            flat = merge_linear(**call)
        elif merge_type == "task_vector":
            flat = merge_task_vector(**call)
        elif merge_type == "ties":
            flat = merge_ties(**call)
        elif merge_type == "pcb":
            flat = merge_pcb(**call)
        elif merge_type == "dare":
            flat = merge_dare(**call)
        else:
            raise ValueError(f"Merge type '{merge_type}' is not supported.")
        return unflatten_model(flat, self.shape_dict)

    @staticmethod
    def _keys_are_aligned(models: Sequence[StateDict], base_model: Optional[StateDict]) -> bool:
        order = list(models[0].keys())
        others = list(models[1:]) + ([base_model] if base_model is not None else [])
        return all(list(m.keys()) == order for m in others)
