"""CPU: the static unit schedule of the fused scoring kernel (csrc/score_topk.cu: make_plan + get_unit, exported host-only
as mr_score_topk_schedule).  Every (query block, item split) pair must be scheduled exactly once, the tile ranges of the
splits must tile [0, T), and the units that run at the same time (one wave of CTA pairs) must walk exactly S item streams
whose member counts are what the pacing counters expect -- for the BASELINE shapes, their per-GPU shards, and odd shapes."""
import ctypes

import numpy as np
import pytest

from mergerec_b200 import _lib

SHAPES = [(65536, 1_000_000, 100), (65536, 500_000, 100), (65536, 250_000, 100), (65536, 125_000, 100),   # config 5, 1-8 GPUs
          (32768, 200_000, 50), (32768, 25_000, 50), (256, 20_000, 10), (1, 1, 1), (255, 257, 7), (513, 100_003, 128),
          (100_000, 3_000, 20), (7_000, 70_000, 64), (40_000, 1_000, 1)]


@pytest.mark.parametrize("Q,N,K", SHAPES)
@pytest.mark.parametrize("mode", [0, 2])
def test_schedule_covers_every_unit_once(Q, N, K, mode):
    lib = _lib.load()
    plan = np.zeros(8, np.int32)
    U = int(lib.mr_score_topk_schedule(Q, N, K, mode, plan.ctypes.data_as(ctypes.c_void_p), None, 0))
    QB, T, S, wave, cg, streams, windows, win_tiles = (int(x) for x in plan)
    assert U == QB * S and QB == -(-Q // (128 * cg)) and T == -(-N // 256) and 1 <= S <= T and 1 <= wave <= 148 // cg
    units = np.zeros((U, 6), np.int32)
    assert int(lib.mr_score_topk_schedule(Q, N, K, mode, plan.ctypes.data_as(ctypes.c_void_p),
                                          units.ctypes.data_as(ctypes.c_void_p), U)) == U
    qb, split, t0, t1, stream, members = units.T
    # every (query block, split) exactly once
    assert (qb >= 0).all() and (qb < QB).all() and (split >= 0).all() and (split < S).all()
    assert len(set(zip(qb.tolist(), split.tolist()))) == U
    # the splits tile [0, T) and a unit's range is its split's
    bounds = [(s * T) // S for s in range(S + 1)]
    assert bounds[0] == 0 and bounds[-1] == T and all(b1 > b0 for b0, b1 in zip(bounds, bounds[1:]))
    assert np.array_equal(t0, np.array(bounds)[split]) and np.array_equal(t1, np.array(bounds)[split + 1])
    # wave w = units [w * wave, (w + 1) * wave): at most S streams per wave, one per split, never shared between waves;
    # `members` = number of units of the wave on that stream (what the pacing counters wait for)
    assert (stream >= 0).all() and (stream < streams).all()
    w = np.arange(U) // wave
    for ww in range(int(w.max()) + 1):
        sel = w == ww
        per = {}
        for st_, sp_, m_ in zip(stream[sel].tolist(), split[sel].tolist(), members[sel].tolist()):
            per.setdefault(st_, []).append((sp_, m_))
        assert len(per) <= S
        for st_, lst in per.items():
            assert len({sp for sp, _ in lst}) == 1 and all(m == len(lst) for _, m in lst)
        assert len({lst[0][0] for lst in per.values()}) == len(per)            # distinct splits
    assert len(set(stream[w == 0].tolist()) & set(stream[w == int(w.max())].tolist())) == (0 if w.max() > 0 else len(set(stream.tolist())))
    # the pacing counters cover the longest unit
    if win_tiles:
        assert windows >= -(-int((t1 - t0).max()) // win_tiles)


def test_schedule_config5_is_the_measured_layout():
    """BASELINE config 5 on one GPU: 256 query blocks x 2 splits, waves of 74 pairs = 37 + 37 units on two streams."""
    lib = _lib.load()
    plan = np.zeros(8, np.int32)
    U = int(lib.mr_score_topk_schedule(65536, 1_000_000, 100, 0, plan.ctypes.data_as(ctypes.c_void_p), None, 0))
    assert (U, int(plan[2]), int(plan[3])) == (512, 2, 74)
    units = np.zeros((U, 6), np.int32)
    lib.mr_score_topk_schedule(65536, 1_000_000, 100, 0, plan.ctypes.data_as(ctypes.c_void_p), units.ctypes.data_as(ctypes.c_void_p), U)
    first = units[:74]
    assert (first[:37, 1] == 0).all() and (first[37:, 1] == 1).all() and (first[:, 5] == 37).all()
    assert first[:37, 0].tolist() == list(range(37)) and first[37:, 0].tolist() == list(range(37))
