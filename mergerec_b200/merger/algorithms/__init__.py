from .dare import merge_dare
from .linear import merge_linear
from .localize_and_stitch import get_localize_and_stitch_vectors, merge_localize_and_stitch
from .pcb import get_pcb_vectors, merge_pcb
from .task_vector import get_task_vectors, merge_task_vector
from .ties import get_ties_vectors, merge_ties

__all__ = ["merge_linear", "merge_task_vector", "merge_ties", "get_task_vectors", "get_ties_vectors",
           "get_localize_and_stitch_vectors", "merge_localize_and_stitch", "get_pcb_vectors", "merge_pcb", "merge_dare"]
