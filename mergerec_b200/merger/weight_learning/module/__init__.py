from ._base import TaskVectorMergingModuleBase
from ._factory import load_merging_module
from .layer_wise import TaskVectorMergingModuleLayerWise, group_parameters_by_layer
from .task_wise import TaskVectorMergingModuleTaskWise

__all__ = ["TaskVectorMergingModuleBase", "load_merging_module", "TaskVectorMergingModuleLayerWise",
           "TaskVectorMergingModuleTaskWise", "group_parameters_by_layer"]
