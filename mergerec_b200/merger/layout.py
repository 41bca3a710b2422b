"""Flat-layout bookkeeping shared by every merger kernel.

The reference concatenates a state_dict's tensors in dict order into one fp32 vector
(rec_retrieval/merger/utils/model_operations.py:47-63) and later slices it back by running offsets
(:66-90, weight_learning/utils.py:30-51).  ``FlatLayout`` is that offset table, plus the block / lambda-group
tables the kernels consume:

* task-wise  (weight_learning/module/task_wise.py:36-48): one block [0, d), one group ``"all"``;
* layer-wise (weight_learning/module/layer_wise.py:13-33): one block per tensor; group key is
  ``name.split(".")[3]`` when ``"encoder.layer."`` occurs in the name, else ``"others"``; group ids follow
  first appearance.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Mapping, Sequence, Tuple

import torch

ROW_ALIGN = 64  # floats; rows of (K, d) buffers start 256-byte aligned so 128-bit loads stay legal for any d


def _numel(shape: Sequence[int]) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


def padded_ld(d: int) -> int:
    return (d + ROW_ALIGN - 1) // ROW_ALIGN * ROW_ALIGN


def alloc_rows(K: int, d: int, device, zero: bool = False) -> torch.Tensor:
    """A (K, d) fp32 view whose rows are 256-byte aligned (row stride = padded_ld(d))."""
    ld = padded_ld(d)
    buf = (torch.zeros if zero else torch.empty)((K, ld), dtype=torch.float32, device=device)
    return buf[:, :d]


@dataclass
class FlatLayout:
    names: List[str]
    shapes: List[torch.Size]
    offsets: List[int]          # start of each tensor in the flat vector
    sizes: List[int]
    d: int
    _dev_cache: Dict[Tuple[str, str], tuple] = field(default_factory=dict, repr=False)

    @classmethod
    def from_shape_dict(cls, shape_dict: Mapping[str, Sequence[int]]) -> "FlatLayout":
        names, shapes, offsets, sizes = [], [], [], []
        off = 0
        for name, shape in shape_dict.items():
            n = _numel(shape)
            names.append(name)
            shapes.append(torch.Size(shape))
            offsets.append(off)
            sizes.append(n)
            off += n
        return cls(names, shapes, offsets, sizes, off)

    @property
    def shape_dict(self) -> Dict[str, torch.Size]:
        return dict(zip(self.names, self.shapes))

    # -- lambda groups -------------------------------------------------------------------------
    def layer_groups(self) -> "Dict[str, List[Tuple[str, int, int]]]":
        """{group key: [(tensor name, start, end), ...]} in first-appearance order (layer_wise.py:13-33)."""
        groups: Dict[str, List[Tuple[str, int, int]]] = {}
        for name, off, n in zip(self.names, self.offsets, self.sizes):
            key = name.split(".")[3] if "encoder.layer." in name else "others"
            groups.setdefault(key, []).append((name, off, off + n))
        return groups

    def block_table(self, layer_wise: bool):
        """(seg_end int64[P], seg_group int32[P], group_keys) on the host."""
        if not layer_wise:
            return [self.d], [0], ["all"]
        keys = list(self.layer_groups().keys())
        index = {k: i for i, k in enumerate(keys)}
        ends, grp = [], []
        for name, off, n in zip(self.names, self.offsets, self.sizes):
            key = name.split(".")[3] if "encoder.layer." in name else "others"
            ends.append(off + n)
            grp.append(index[key])
        return ends, grp, keys

    def device_blocks(self, layer_wise: bool, device) -> Tuple[torch.Tensor, torch.Tensor, List[str]]:
        """Device copies of block_table (cached per device)."""
        ck = ("lw" if layer_wise else "tw", str(device))
        if ck not in self._dev_cache:
            ends, grp, keys = self.block_table(layer_wise)
            self._dev_cache[ck] = (torch.tensor(ends, dtype=torch.int64, device=device),
                                   torch.tensor(grp, dtype=torch.int32, device=device), keys)
        return self._dev_cache[ck]

    def device_segments(self, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """(seg_off int64[P], seg_len int64[P]) on the device (cached)."""
        ck = ("seg", str(device))
        if ck not in self._dev_cache:
            self._dev_cache[ck] = (torch.tensor(self.offsets, dtype=torch.int64, device=device),
                                   torch.tensor(self.sizes, dtype=torch.int64, device=device))
        return self._dev_cache[ck]

    def views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Slice + reshape views of a flat vector, in layout order."""
        if flat.numel() != self.d:
            raise AssertionError("Flattened tensor size does not match the expected size.")
        return {n: flat[o:o + s].reshape(sh) for n, sh, o, s in zip(self.names, self.shapes, self.offsets, self.sizes)}
