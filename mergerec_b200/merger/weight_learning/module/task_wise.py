"""One lambda per domain model (reference: rec_retrieval/merger/weight_learning/module/task_wise.py)."""
from __future__ import annotations

from typing import Dict, List

import torch
from torch import nn

from ._base import TaskVectorMergingModuleBase


class TaskVectorMergingModuleTaskWise(TaskVectorMergingModuleBase):
    LAYER_WISE = False

    def __init__(self, base_model_tensor: torch.Tensor, task_vectors_tensor: torch.Tensor, model_without_params,
                 shape_dict: Dict[str, torch.Size], initial_global_weight: float = 1.0, initial_global_bias: float = 0.0,
                 initial_per_weight: float = 0.2, disable_softmax: bool = True):
        super().__init__(base_model_tensor, task_vectors_tensor, model_without_params, shape_dict, disable_softmax)
        K = task_vectors_tensor.size(0)
        self.global_weights["all"] = nn.Parameter(torch.full((1,), initial_global_weight))
        self.global_biases["all"] = nn.Parameter(torch.full((1,), initial_global_bias))
        self.per_weights["all"] = nn.Parameter(torch.full((K,), initial_per_weight))

    def _group_keys(self) -> List[str]:
        return ["all"]
