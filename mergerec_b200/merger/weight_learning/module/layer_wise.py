"""One lambda per domain model and encoder layer (reference: rec_retrieval/merger/weight_learning/module/layer_wise.py)."""
from __future__ import annotations

from typing import Dict, List

import torch
from torch import nn

from ...layout import FlatLayout
from ._base import TaskVectorMergingModuleBase


def group_parameters_by_layer(shape_dict):
    """{group key: [(tensor name, start, end), ...]}: tensors whose name contains ``encoder.layer.`` belong to the
    group named by the 4th dotted component (the layer index); everything else to ``"others"``
    (layer_wise.py:13-33)."""
    return FlatLayout.from_shape_dict(shape_dict).layer_groups()


class TaskVectorMergingModuleLayerWise(TaskVectorMergingModuleBase):
    LAYER_WISE = True

    def __init__(self, base_model_tensor: torch.Tensor, task_vectors_tensor: torch.Tensor, model_without_params,
                 shape_dict: Dict[str, torch.Size], initial_global_weight: float = 1.0, initial_global_bias: float = 0.0,
                 initial_per_weight: float = 0.2, disable_softmax: bool = False):
        super().__init__(base_model_tensor, task_vectors_tensor, model_without_params, shape_dict, disable_softmax)
        self._num_model_parameters = base_model_tensor.size(0)
        self.layer_groups = group_parameters_by_layer(shape_dict)
        K = task_vectors_tensor.size(0)
        for layer_key in self.layer_groups:
            self.global_weights[layer_key] = nn.Parameter(torch.full((1,), initial_global_weight))
            self.global_biases[layer_key] = nn.Parameter(torch.full((1,), initial_global_bias))
            self.per_weights[layer_key] = nn.Parameter(torch.full((K,), initial_per_weight))

    def _group_keys(self) -> List[str]:
        return list(self.layer_groups.keys())
