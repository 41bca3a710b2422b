"""Enums of the merger API (member names/values as in the reference's rec_retrieval/merger/enums.py:11-40)."""
from enum import Enum

__all__ = ["MergeType", "LearnType", "LossType"]

MergeType = Enum("MergeType", {n: n for n in ("TASK_VECTOR", "TIES", "PCB", "LOCALIZE_AND_STITCH")})
MergeType.__doc__ = "How the per-domain vectors are built before lambda-weighting."

LearnType = Enum("LearnType", {n: n for n in ("TASK_WISE", "LAYER_WISE")})
LearnType.__doc__ = "Granularity of the learnable lambdas: one per domain, or one per domain and layer group."

LossType = Enum("LossType", {n: n for n in (
    "CE", "KD", "MSE", "ADAMERGING", "ADAMERGING_KD", "MERGED_PSEUDO_LABEL", "SINGLE_PSEUDO_LABEL",
    "MERGED_PSEUDO_LABEL_KD", "SINGLE_PSEUDO_LABEL_KD")})
LossType.__doc__ = "Distillation losses of the collaborative-merging caller (kept for config parity only)."
