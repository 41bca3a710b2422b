"""Shared test helpers (input regeneration from golden_cases seeds, bit comparisons)."""
import os
from collections import OrderedDict

import numpy as np

from mergerec_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 8: np.uint64, 2: np.uint16, 1: np.uint8}[a.dtype.itemsize])


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert a.dtype == b.dtype, f"{what}: dtype {a.dtype} vs {b.dtype}"
    bad = bits(a) != bits(b)
    if bad.any():
        idx = np.argwhere(bad)[:5]
        raise AssertionError(f"{what}: {int(bad.sum())} of {a.size} elements differ bitwise; first at {idx.tolist()}: "
                             f"{a[tuple(idx[0])]!r} vs {b[tuple(idx[0])]!r}")


def state_dict_case(case):
    shapes = synth.tiny_shapes(recformer=case["recformer"])
    base, models = synth.make_state_dicts(shapes, case["K"], seed=case["seed"], sigma=1e-2)
    return shapes, base, models


def flatten_np(sd, keys=None):
    keys = list(sd.keys()) if keys is None else keys
    return np.concatenate([np.asarray(sd[k]).reshape(-1).astype(np.float32) for k in keys])


def shape_dict_of(sd, keys=None):
    keys = list(sd.keys()) if keys is None else keys
    return OrderedDict((k, tuple(sd[k].shape)) for k in keys)
