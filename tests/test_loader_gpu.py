"""GPU: the streamed host -> HBM loader lands every tensor of a state_dict at its flat offset (reference layout:
rec_retrieval/merger/utils/model_operations.py:47-63), for pageable, pinned, mixed-dtype and device-resident tensors."""
import numpy as np
import pytest
import torch

from mergerec_b200 import synth
from mergerec_b200.merger import ModelMerger
from mergerec_b200.merger.layout import FlatLayout
from mergerec_b200.merger.utils.loader import StreamedFlatLoader, flatten_models_streamed
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _dicts(recformer, K=2, seed=3):
    shapes = synth.tiny_shapes(layers=3, hidden=40, ffn=72, vocab=3001, max_pos=18, recformer=recformer)
    base, models = synth.make_state_dicts(shapes, K, seed=seed, sigma=1e-2)
    return shapes, base, models


@pytest.mark.parametrize("recformer", [False, True])
@pytest.mark.parametrize("chunk_bytes", [1 << 12, 1 << 16, 64 << 20])
def test_pageable_state_dict_lands_in_the_flat_layout(recformer, chunk_bytes):
    """Pageable tensors (incl. the int64 position_ids buffer, promoted to fp32 like torch.cat does) through staging buffers
    much smaller / larger than the tensors: chunk boundaries inside tensors, many tensors per chunk."""
    shapes, base, models = _dicts(recformer)
    want, _ = orc.flatten_model(base)
    loader = StreamedFlatLoader(chunk_bytes=chunk_bytes, n_buffers=2, host_threads=3)
    flat, ready = loader.load({k: torch.from_numpy(v) for k, v in base.items()})
    torch.cuda.current_stream().wait_event(ready)
    assert np.array_equal(flat.cpu().numpy().view(np.uint32), want.view(np.uint32))
    loader.close()


def test_pinned_adjacent_tensors_are_coalesced_and_mixed_sources_work():
    shapes, base, models = _dicts(False)
    layout = FlatLayout.from_shape_dict(shapes)
    want, _ = orc.flatten_model(models[0])
    pinned_flat = torch.from_numpy(want.copy()).pin_memory()
    views = layout.views(pinned_flat)                       # one pinned buffer, every tensor a view: adjacent in memory
    loader = StreamedFlatLoader(host_threads=1)
    flat, ready = loader.load(views)
    torch.cuda.current_stream().wait_event(ready)
    assert torch.equal(flat.cpu(), torch.from_numpy(want))
    assert loader.h2d_calls == 1, "adjacent pinned tensors must go over in one copy"
    # mixed: some tensors pageable, some pinned, some already on the device
    mixed = {}
    for i, (k, v) in enumerate(models[1].items()):
        t = torch.from_numpy(v)
        mixed[k] = t.pin_memory() if i % 3 == 0 else (t.cuda() if i % 3 == 1 else t)
    flat2, ready2 = loader.load(mixed)
    torch.cuda.current_stream().wait_event(ready2)
    want2, _ = orc.flatten_model(models[1])
    assert np.array_equal(flat2.cpu().numpy().view(np.uint32), want2.view(np.uint32))
    loader.close()


def test_model_merger_from_host_state_dicts_uses_the_loader():
    """The public constructor with CPU state_dicts == the oracle's merge, bit for bit (flatten through the loader)."""
    shapes, base, models = _dicts(True, K=3, seed=8)
    tb = {k: torch.from_numpy(v) for k, v in base.items()}
    tm = [{k: torch.from_numpy(v) for k, v in m.items()} for m in models]
    merged = ModelMerger(tm, tb).merge("task_vector", 0.3)
    keys = sorted(base.keys())
    fb = np.concatenate([base[k].reshape(-1).astype(np.float32) for k in keys])
    fm = [np.concatenate([m[k].reshape(-1).astype(np.float32) for k in keys]) for m in models]
    want = orc.merge_task_vector(fb, fm, [0.3] * 3)
    got = torch.cat([merged[k].reshape(-1) for k in keys]).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    flats, shape_dict = flatten_models_streamed([tb] + tm)
    assert len(flats) == 4 and list(shape_dict) == list(tb)
