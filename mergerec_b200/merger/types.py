"""Type aliases of the merger API (names as in the reference's rec_retrieval/merger/types.py:15-19)."""
from pathlib import Path
from typing import Dict, Union

import torch

PathStr = Union[str, Path]
FlattenedModel = torch.Tensor    # (d,) fp32: every state_dict tensor reshaped to 1-D and concatenated
FlattenedModel2D = torch.Tensor  # (K, d) fp32: one flat row per domain model / task vector
StateDict = Dict[str, torch.Tensor]
ShapeDict = Dict[str, torch.Size]

__all__ = ["PathStr", "FlattenedModel", "FlattenedModel2D", "StateDict", "ShapeDict"]
