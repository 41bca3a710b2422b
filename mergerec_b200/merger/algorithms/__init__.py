from .linear import merge_linear
from .task_vector import get_task_vectors, merge_task_vector
from .ties import get_ties_vectors, merge_ties

__all__ = ["merge_linear", "merge_task_vector", "merge_ties", "get_task_vectors", "get_ties_vectors"]
