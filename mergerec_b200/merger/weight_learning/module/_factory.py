"""``load_merging_module`` (reference: rec_retrieval/merger/weight_learning/module/_factory.py:27-127)."""
from __future__ import annotations

from typing import List, Optional, Set

import torch

from ...algorithms.localize_and_stitch import get_localize_and_stitch_vectors
from ...algorithms.pcb import get_pcb_vectors
from ...algorithms.task_vector import get_task_vectors
from ...algorithms.ties import get_ties_vectors
from ...enums import LearnType, MergeType
from ...merger import ModelMerger
from ...types import StateDict
from ..utils import make_functional
from ._base import TaskVectorMergingModuleBase
from .layer_wise import TaskVectorMergingModuleLayerWise
from .task_wise import TaskVectorMergingModuleTaskWise


def _check_isinstance_state_dict(t) -> None:
    if not isinstance(t, dict):
        raise ValueError(f"Expected a state dict, got {type(t)}")
    for k, v in t.items():
        if not isinstance(k, str):
            raise ValueError(f"Expected a string key, got {type(k)}")
        if not isinstance(v, torch.Tensor):
            raise ValueError(f"Expected a tensor value, got {type(v)}")


def load_merging_module(merge_type: MergeType, learn_type: LearnType, model: torch.nn.Module,
                        pretrain_state_dict: StateDict, finetune_state_dicts: List[StateDict], ignore_keys: Set[str],
                        ties_density: Optional[float] = None, initial_global_weight: float = 1.0,
                        initial_global_bias: float = 0.0, initial_per_weight: float = 0.2,
                        disable_softmax: bool = False) -> TaskVectorMergingModuleBase:
    """Build the collaborative-merging module.  Key order = the pre-trained dict's order restricted to
    ``pretrain.keys() & finetune[0].keys() - ignore_keys`` (_factory.py:55-66); the wrapped ``model`` loses its
    parameters (``make_functional``) and gets them back as views of the merged vector on every forward.
    The flat base and the (K, d) vectors are built directly in HBM."""
    assert isinstance(merge_type, MergeType), f"Invalid merge type: {merge_type}"
    assert isinstance(learn_type, LearnType), f"Invalid learn type: {learn_type}"
    _check_isinstance_state_dict(pretrain_state_dict)
    for ckpt in finetune_state_dicts:
        _check_isinstance_state_dict(ckpt)

    keys_to_keep = set(pretrain_state_dict.keys() & finetune_state_dicts[0].keys()) - set(ignore_keys)
    pretrain_state_dict = {k: v for k, v in pretrain_state_dict.items() if k in keys_to_keep}
    order = list(pretrain_state_dict.keys())
    finetune_state_dicts = [{k: ckpt[k] for k in order if k in ckpt} for ckpt in finetune_state_dicts]

    make_functional(model)
    merger = ModelMerger(models=finetune_state_dicts, base_model=pretrain_state_dict, align_key_order=False)
    if merge_type is MergeType.TASK_VECTOR:
        vectors = get_task_vectors(base_model=merger.base_model, models=merger.models)
    elif merge_type is MergeType.TIES:
        assert ties_density is not None, "Density should be provided for ties merging."
        vectors = get_ties_vectors(base_model=merger.base_model, models=merger.models, density=ties_density)
    elif merge_type is MergeType.LOCALIZE_AND_STITCH:
        assert ties_density is not None, "Density should be provided for localize-and-stitch merging."
        vectors = get_localize_and_stitch_vectors(base_model=merger.base_model, models=merger.models,
                                                  density=ties_density)
    elif merge_type is MergeType.PCB:
        vectors = get_pcb_vectors(base_model=merger.base_model, models=merger.models, density=ties_density)
    else:
        raise ValueError(f"Invalid merge type: {merge_type}")

    if learn_type is LearnType.TASK_WISE:
        cls = TaskVectorMergingModuleTaskWise
    elif learn_type is LearnType.LAYER_WISE:
        cls = TaskVectorMergingModuleLayerWise
    else:
        raise ValueError(f"Invalid learn type: {learn_type}")
    module = cls(merger.base_model.detach().requires_grad_(False), vectors.detach().requires_grad_(False), model,
                 shape_dict=merger.shape_dict, initial_global_weight=initial_global_weight,
                 initial_global_bias=initial_global_bias, initial_per_weight=initial_per_weight,
                 disable_softmax=disable_softmax)
    return module.to(merger.base_model.device)
