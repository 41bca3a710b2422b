from .merger import ModelMerger

__all__ = ["ModelMerger"]
