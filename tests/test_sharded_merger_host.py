"""CPU, gloo, world size 2: the sharded merger's host logic (mergerec_b200/merger/sharded.py) with oracle stand-ins for
the kernels -- global TIES trim from all-reduced radix histograms, ties straddling the cut resolved towards the lowest
global index by the rank-order scan, per-rank cut keys, slice all-gather.  Everything must be bit-identical to the
unsharded oracle."""
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist

from mergerec_b200 import synth
from mergerec_b200.merger.sharded import (flat_shard_bounds, gather_flat, get_ties_vectors_sharded, merge_linear_sharded,
                                          merge_task_vector_sharded, merge_ties_sharded, sharded_select)
from oracle import oracle as orc
from sharded_helpers import OracleKernels

CASES = [
    dict(K=3, d=4165, seed=41, tie_free=True, density=0.2, weights=[0.3, 0.5, 0.7]),
    dict(K=6, d=4165, seed=45, tie_free=False, quantize=2.5e-4, density=0.2, weights=[0.5] * 6),      # ties straddle the cut
    dict(K=5, d=70, seed=46, tie_free=False, quantize=1e-3, density=0.5, weights=[1.0, 0.5, 0.25, 2.0, 1.5]),
    dict(K=2, d=20, seed=47, tie_free=True, density=0.3, weights=[0.6, 0.4]),                           # rank 1 owns nothing
    # flattened tiny Recformer state_dicts (d = 14,346; an all-zero update block): dense enough that a refined window
    # is wider than the bin it refines -- the excess bins must be ignored
    dict(K=5, d=14346, seed=22, state_dict=True, density=0.2, weights=[0.4] * 5),
]


def test_flat_shard_bounds_cover_and_align():
    for d in (0, 1, 31, 32, 33, 4165, 124645632, 433610754):
        for world in (1, 2, 3, 8):
            b = [flat_shard_bounds(d, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == d
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all(lo % 32 == 0 for lo, _ in b if lo < d)
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 32 + 31
    with pytest.raises(ValueError):
        flat_shard_bounds(10, 2, 2)


def _case_inputs(case):
    if case.get("state_dict"):
        shapes = synth.tiny_shapes(recformer=True)
        base, models = synth.make_state_dicts(shapes, case["K"], seed=case["seed"], sigma=1e-2)
        flat = lambda sd: np.concatenate([np.asarray(sd[k]).reshape(-1).astype(np.float32) for k in sorted(base)])  # noqa: E731
        out = flat(base), [flat(m) for m in models]
        assert out[0].size == case["d"]
        return out
    return synth.make_flat(case["d"], case["K"], seed=case["seed"], tie_free=case.get("tie_free", False),
                           quantize=case.get("quantize"))


def _check_case(case, group, world, rank):
    base, models = _case_inputs(case)
    d = case["d"]
    lo, hi = flat_shard_bounds(d, world, rank)
    bl = torch.from_numpy(base[lo:hi].copy())
    ml = [torch.from_numpy(m[lo:hi].copy()) for m in models]
    That_l = get_ties_vectors_sharded(bl, ml, case["density"], d, group, kernels=OracleKernels)
    want = orc.ties_vectors(base, models, case["density"])
    ok = np.array_equal(That_l.numpy().view(np.uint32), want[:, lo:hi].view(np.uint32))
    merged_l = merge_ties_sharded(bl, ml, case["weights"], case["density"], d, group, kernels=OracleKernels)
    full = gather_flat(merged_l, d, group)
    ok = ok and np.array_equal(full.numpy().view(np.uint32), orc.merge_ties(base, models, case["weights"], case["density"]).view(np.uint32))
    tv = gather_flat(merge_task_vector_sharded(bl, ml, case["weights"], kernels=OracleKernels), d, group)
    ok = ok and np.array_equal(tv.numpy().view(np.uint32), orc.merge_task_vector(base, models, case["weights"]).view(np.uint32))
    ln = gather_flat(merge_linear_sharded(ml, case["weights"], kernels=OracleKernels), d, group)
    ok = ok and np.array_equal(ln.numpy().view(np.uint32), orc.merge_linear(models, case["weights"]).view(np.uint32))
    return bool(ok)


def test_world1_matches_oracle():
    for case in CASES:
        assert _check_case(case, None, 1, 0), case


def test_select_extremes():
    base, models = _case_inputs(CASES[0])
    bl, ml = torch.from_numpy(base), [torch.from_numpy(m) for m in models]
    assert sharded_select(bl, ml, 0, base.size, kernels=OracleKernels).tolist() == [-1, -1, -1]
    assert sharded_select(bl, ml, base.size, base.size, kernels=OracleKernels).tolist() == [0, 0, 0]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = all(_check_case(case, dist.group.WORLD, world, rank) for case in CASES)
        q.put((rank, bool(ok)))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merges_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]


def test_window_miss_falls_back_to_the_full_range():
    """A wrong first window (here: estimates of 0, or absurdly high) must be detected and replaced by lo = 0, shift = 20."""
    class BadLow(OracleKernels):
        @staticmethod
        def kth_largest_bits(base, rows, k, w):
            return torch.zeros(len(rows), dtype=torch.int64), None

    class BadHigh(OracleKernels):
        @staticmethod
        def kth_largest_bits(base, rows, k, w):
            return torch.full((len(rows),), 0x7F000000, dtype=torch.int64), None

    for case in CASES[:2]:
        base, models = _case_inputs(case)
        bl, ml = torch.from_numpy(base), [torch.from_numpy(m) for m in models]
        want = orc.ties_vectors(base, models, case["density"])
        for kern in (BadLow, BadHigh):
            got = get_ties_vectors_sharded(bl, ml, case["density"], case["d"], None, kernels=kern)
            assert np.array_equal(got.numpy().view(np.uint32), want.view(np.uint32))


def test_truncated_tie_list_rescans(monkeypatch):
    """Equal magnitudes straddling the cut are normally resolved from the (bin, index) list recorded by the last
    histogram level; when that list is too short for them the slice is rescanned -- same answer either way."""
    import mergerec_b200.merger.sharded as sh
    calls = []
    orig = sh._tie_index
    monkeypatch.setattr(sh, "_tie_index", lambda *a: (calls.append(1), orig(*a))[1])
    case = CASES[2]                     # a handful of equal magnitudes at the cut: resolved from the recorded list
    base, models = _case_inputs(case)
    bl, ml = torch.from_numpy(base), [torch.from_numpy(m) for m in models]
    want = orc.ties_vectors(base, models, case["density"])
    got = get_ties_vectors_sharded(bl, ml, case["density"], case["d"], None, kernels=OracleKernels)
    assert np.array_equal(got.numpy().view(np.uint32), want.view(np.uint32)) and not calls
    monkeypatch.setattr(sh, "CAND_CAP", 2)
    got = get_ties_vectors_sharded(bl, ml, case["density"], case["d"], None, kernels=OracleKernels)
    assert np.array_equal(got.numpy().view(np.uint32), want.view(np.uint32)) and calls
