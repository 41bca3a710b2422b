"""Item-side distillation (reference: rec_retrieval/module/distiller/item/module.py:18-160).

Same step as the sequence distiller with the roles of the batch fields changed: the merged model encodes ITEM texts
(`batch.items`), the logits are taken against the dataset's item table and the teacher row is
`score_embeddings[dataset_index][item_id]` (:87-101).  It therefore shares the three kernels of
`DistillSequenceModule._forward_distill` (csrc/distill.cu); only the batch plumbing differs."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List

import torch

from ..sequence.module import DistillSequenceModule, fused_distill_losses

__all__ = ["DistillModule", "BatchDistillationItem"]


@dataclass
class BatchDistillationItem:
    """Fields `_forward_distill` reads (ref: rec_retrieval/types/model_batch.py, item/module.py:87-93)."""
    items: Any
    dataset_indexes: List[int]
    item_ids: List[int]


class DistillModule(DistillSequenceModule):
    """ref: item/module.py:18-160 (same constructor as the sequence module)."""

    def forward(self, batch):
        if hasattr(batch, "item_ids"):
            return self._forward_distill(batch)
        if hasattr(batch, "items"):
            return self._forward_item_encoding(batch.items)
        raise ValueError(f"Invalid batch type {type(batch)}")

    def _forward_item_encoding(self, batch):
        return self._forward_sequence_encoding(batch)

    def _forward_distill(self, batch):
        if self.item_embeddings is None:
            raise RuntimeError("item_embeddings have not been injected")
        rep = self._forward_item_encoding(batch.items).to(torch.float32)
        spec = self.loss_fn.spec
        keep, ptrs = (None, None)
        if spec.needs_teacher:
            keep, ptrs = self.score_embeddings.rows(batch.dataset_indexes, batch.item_ids, self.item_embeddings)
        losses = fused_distill_losses(rep, list(self.item_embeddings), batch.dataset_indexes, ptrs, spec)
        del keep
        return losses.mean()
