"""Streamed host -> HBM loader of state_dicts straight into the flat layout.

The reference flattens on the CPU (``torch.cat`` of every tensor, rec_retrieval/merger/utils/model_operations.py:47-63,
0.7 s for four RoBERTa-base dicts) and leaves the device copy to the caller (merge_test.py:20-25 loads the checkpoints
with ``torch.load(map_location="cpu")``).  Here every tensor lands at its offset of ONE device vector:

* tensors that already sit in pinned host memory are copied in place, adjacent ones (views of one pinned buffer, an
  mmap-ed checkpoint) coalesced into a single ``cudaMemcpyAsync`` per run;
* pageable tensors (what ``torch.load`` returns) are packed -- dtype promotion to fp32 included, exactly what
  ``torch.cat`` does to the int64 ``position_ids`` buffer -- into a ring of pinned staging buffers by a few host threads
  while the previous buffer is on its way over PCIe, so the copy engine never waits for a pageable bounce;
* everything is issued on a dedicated copy stream; consumers synchronise with an event per model (``ready``), so the
  kernels over model k can be queued while model k + 1 is still in flight and nothing blocks the host.

No CUDA kernel of this package is involved -- this is the data format either side of the merger path.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence, Tuple

import torch

from ... import _lib
from ..layout import FlatLayout
from ..types import FlattenedModel, ShapeDict, StateDict

__all__ = ["StreamedFlatLoader", "flatten_models_streamed", "default_loader"]


class StreamedFlatLoader:
    def __init__(self, device: Optional[torch.device] = None, chunk_bytes: int = 64 << 20, n_buffers: int = 3,
                 host_threads: int = 4):
        self.device = device or _lib.require_cuda()
        self.stream = torch.cuda.Stream(device=self.device)
        self.chunk = max(int(chunk_bytes) // 4, 1024)          # fp32 elements per staging buffer
        self._bufs: List[torch.Tensor] = []
        self._free: List[Optional[torch.cuda.Event]] = []
        self._n_buffers = n_buffers
        self._next = 0
        self._pool = ThreadPoolExecutor(max_workers=max(1, host_threads)) if host_threads > 1 else None
        self.h2d_calls = 0                                     # diagnostics: cudaMemcpyAsync calls issued

    # ---------------------------------------------------------------------------------------------- staging ring
    def _staging(self) -> Tuple[int, torch.Tensor]:
        """Next staging buffer, once the copy that last read it has completed (host wait on that one event only)."""
        if len(self._bufs) < self._n_buffers:
            self._bufs.append(torch.empty(self.chunk, dtype=torch.float32, pin_memory=True))
            self._free.append(None)
            return len(self._bufs) - 1, self._bufs[-1]
        i = self._next
        self._next = (self._next + 1) % self._n_buffers
        if self._free[i] is not None:
            self._free[i].synchronize()
        return i, self._bufs[i]

    def _fill(self, buf: torch.Tensor, pieces: List[Tuple[int, torch.Tensor]]) -> None:
        """pieces: (position in the staging buffer, 1-D source slice).  Host copies with dtype promotion to fp32."""
        def one(p):
            pos, src = p
            buf[pos:pos + src.numel()].copy_(src)
        if self._pool is not None and len(pieces) > 1:
            list(self._pool.map(one, pieces))
        else:
            for p in pieces:
                one(p)

    @staticmethod
    def _split(src: torch.Tensor, parts: int) -> List[Tuple[int, torch.Tensor]]:
        n = src.numel()
        step = max((n + parts - 1) // parts, 1 << 18)
        return [(a, src[a:min(n, a + step)]) for a in range(0, n, step)]

    # ---------------------------------------------------------------------------------------------- one model
    @torch.no_grad()
    def load(self, model: StateDict, layout: Optional[FlatLayout] = None, out: Optional[torch.Tensor] = None
             ) -> Tuple[FlattenedModel, torch.cuda.Event]:
        """Queue the copies of one state_dict into a flat fp32 device vector.  Returns (flat, ready): `ready` is recorded
        on the copy stream after the last copy -- `torch.cuda.current_stream().wait_event(ready)` before using `flat`."""
        layout = layout or FlatLayout.from_shape_dict({k: v.shape for k, v in model.items()})
        flat = out if out is not None else torch.empty(layout.d, dtype=torch.float32, device=self.device)
        if flat.numel() != layout.d:
            raise AssertionError("Flattened tensor size does not match the expected size.")
        # the flat vector may have been allocated on the current stream a moment ago
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        run_ptr = run_len = run_off = None            # current run of adjacent pinned fp32 host tensors
        run_src: Optional[torch.Tensor] = None
        stage_idx, stage, stage_pos, stage_off = -1, None, 0, 0
        pieces: List[Tuple[int, torch.Tensor]] = []

        def flush_run():
            nonlocal run_ptr, run_src
            if run_ptr is None:
                return
            src = torch.as_strided(run_src, (run_len,), (1,), run_src.storage_offset())
            with torch.cuda.stream(self.stream):
                flat[run_off:run_off + run_len].copy_(src, non_blocking=True)
            self.h2d_calls += 1
            run_ptr = run_src = None

        def flush_stage():
            nonlocal stage, stage_pos, pieces
            if stage is None or stage_pos == 0:
                return
            self._fill(stage, pieces)
            with torch.cuda.stream(self.stream):
                flat[stage_off:stage_off + stage_pos].copy_(stage[:stage_pos], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self._free[stage_idx] = ev
            self.h2d_calls += 1
            stage, stage_pos, pieces = None, 0, []

        for t, off, n in zip(model.values(), layout.offsets, layout.sizes):
            if n == 0:
                continue
            src = t.detach().reshape(-1)
            if src.is_cuda:
                flush_run(); flush_stage()
                with torch.cuda.stream(self.stream):
                    flat[off:off + n].copy_(src, non_blocking=True)
                continue
            if src.dtype == torch.float32 and src.is_contiguous() and src.is_pinned():
                flush_stage()
                if run_ptr is not None and src.data_ptr() == run_ptr + 4 * run_len and run_off + run_len == off \
                        and src.untyped_storage().data_ptr() == run_src.untyped_storage().data_ptr():
                    run_len += n                      # adjacent in host memory and in the flat layout: extend the run
                else:
                    flush_run()
                    run_ptr, run_len, run_off, run_src = src.data_ptr(), n, off, src
                continue
            # pageable (or non-fp32) host tensor: pack it into the staging ring, splitting what does not fit
            flush_run()
            done = 0
            while done < n:
                if stage is None:
                    stage_idx, stage = self._staging()
                    stage_pos, stage_off = 0, off + done
                take = min(n - done, self.chunk - stage_pos)
                part = src[done:done + take]
                if take >= (1 << 20) and self._pool is not None:
                    pieces.extend((stage_pos + a, p) for a, p in self._split(part, self._pool._max_workers))
                else:
                    pieces.append((stage_pos, part))
                stage_pos += take
                done += take
                if stage_pos == self.chunk:
                    flush_stage()
        flush_run()
        flush_stage()
        ready = torch.cuda.Event()
        ready.record(self.stream)
        return flat, ready

    def close(self) -> None:
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None


_DEFAULT: dict = {}


def default_loader(device: Optional[torch.device] = None) -> StreamedFlatLoader:
    """One persistent loader per device (its pinned staging ring is allocated once and protected by events, so no call
    ever has to drain the copy stream)."""
    device = device or _lib.require_cuda()
    key = (device.type, device.index)
    if key not in _DEFAULT:
        _DEFAULT[key] = StreamedFlatLoader(device)
    return _DEFAULT[key]


def flatten_models_streamed(models: Sequence[StateDict], device: Optional[torch.device] = None,
                            loader: Optional[StreamedFlatLoader] = None) -> Tuple[List[FlattenedModel], ShapeDict]:
    """All dicts (same keys, same order) -> flat device vectors; the current stream waits for the last copy (no host sync)."""
    loader = loader or default_loader(device)
    shape_dict = {k: v.shape for k, v in models[0].items()}
    layout = FlatLayout.from_shape_dict(shape_dict)
    flats, ready = [], None
    for m in models:
        f, ready = loader.load(m, layout)
        flats.append(f)
    if ready is not None:
        torch.cuda.current_stream(loader.device).wait_event(ready)
    return flats, shape_dict
