"""Item-table production for the distillation step and the evaluator (first clause).

Reference: rec_retrieval/module/callbacks.py:18-50 (`ItemEncoderMixin.encode_items` / `inject_item_embeddings`) and
:85-110 (`MultiDatasetItemEncodingCallback`).  Same names and hook signatures; Lightning's `Callback` base and `tqdm`
are out of scope, so the hooks are plain methods a training loop calls.  The encoder forward stays PyTorch; what this
produces -- one contiguous fp32 `(num_items, E)` table per dataset on the GPU -- is exactly the layout
`mr_distill_logits`, `mr_distill_grad` and `ShardedItemTable` stream."""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
from torch import nn

__all__ = ["ItemEncoderMixin", "ItemEncodingCallback", "MultiDatasetItemEncodingCallback", "WeightCheckpointCallback"]


def _to_device(batch, device):
    return batch.to(device) if hasattr(batch, "to") else batch


def _device_of(module: nn.Module):
    for p in module.parameters():
        return p.device
    for b in module.buffers():
        return b.device
    return torch.device("cpu")


class ItemEncoderMixin:
    @staticmethod
    @torch.no_grad()
    def encode_items(item_dataloader: Iterable, pl_module: nn.Module) -> torch.Tensor:
        """Run the module over every item batch in eval mode and concatenate the outputs (callbacks.py:19-38)."""
        train_status = pl_module.training
        pl_module.eval()
        device = getattr(pl_module, "device", None) or _device_of(pl_module)
        chunks = [pl_module.forward(_to_device(batch, device)) for batch in item_dataloader]
        item_embeddings = torch.cat(chunks, dim=0).to(torch.float32).contiguous()
        pl_module.train(train_status)
        return item_embeddings

    def inject_item_embeddings(self, item_dataloader: Iterable, pl_module: nn.Module, requires_grad: bool = False):
        """callbacks.py:40-50."""
        pl_module.item_embeddings = nn.Parameter(self.encode_items(item_dataloader, pl_module), requires_grad=requires_grad)


class ItemEncodingCallback(ItemEncoderMixin):
    """callbacks.py:53-66."""

    def __init__(self, item_dataloader: Iterable | None = None):
        self.item_dataloader = item_dataloader

    def on_train_epoch_start(self, trainer, pl_module):
        self.inject_item_embeddings(self.item_dataloader, pl_module)

    def on_test_epoch_start(self, trainer, pl_module):
        if pl_module.item_embeddings is None:
            self.inject_item_embeddings(self.item_dataloader, pl_module)


class MultiDatasetItemEncodingCallback(ItemEncoderMixin):
    """One table per dataset, kept in a `ParameterList` and encoded once (callbacks.py:85-118)."""

    def __init__(self, item_dataloaders: Sequence[Iterable]):
        self.item_dataloaders = item_dataloaders

    def inject_item_embeddings(self, item_dataloaders: Sequence[Iterable], pl_module: nn.Module, requires_grad: bool = False):
        if pl_module.item_embeddings is not None:
            return
        tables: List[nn.Parameter] = [nn.Parameter(self.encode_items(dl, pl_module), requires_grad=requires_grad)
                                      for dl in item_dataloaders]
        pl_module.item_embeddings = nn.ParameterList(tables)

    def on_train_epoch_start(self, trainer, pl_module):
        self.inject_item_embeddings(self.item_dataloaders, pl_module)

    def on_test_epoch_start(self, trainer, pl_module):
        if pl_module.item_embeddings is None:
            self.inject_item_embeddings(self.item_dataloaders, pl_module)


class WeightCheckpointCallback:
    """Keep the best lambda set seen during validation and restore it afterwards (callbacks.py:177-205).

    `monitor` is a regex matched with `re.fullmatch` against the names in `trainer.callback_metrics`; the score of an
    epoch is the mean of the matching metrics, lower is better.  Built on `serialize_weights` /
    `load_weights_from_dict` of the merging module (weight_learning/module/_base.py:67-89)."""

    def __init__(self, monitor: str = "val/loss"):
        self.monitor = monitor
        self.best_score = float("inf")
        self.best_weights = None

    def on_validation_epoch_end(self, trainer, pl_module):
        import re
        scores = []
        for name, value in trainer.callback_metrics.items():
            if re.fullmatch(self.monitor, name):
                scores.append(value.item() if hasattr(value, "item") else float(value))
        if len(scores) == 0:
            raise RuntimeError(f"No metrics found matching the monitor pattern: {self.monitor}")
        current_score = sum(scores) / len(scores)
        if current_score < self.best_score:
            self.best_score = current_score
            self.best_weights = pl_module.merged_model.serialize_weights()

    def load_weights(self, pl_module):
        if self.best_weights is not None:
            pl_module.merged_model.load_weights_from_dict(self.best_weights)
