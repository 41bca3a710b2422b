from .module import BatchDistillationItem, DistillModule  # noqa: F401
