"""CPU: the numpy restatement of get_pcb_vectors / merge_pcb (oracle/oracle.py) against tests/golden/pcb.npz, the
outputs of the unmodified reference (rec_retrieval/merger/algorithms/pcb.py) on the same seeded inputs."""
import numpy as np
import pytest

import golden_cases as gc
from helpers import golden
from mergerec_b200 import synth
from oracle import oracle as orc


@pytest.mark.parametrize("case", gc.PCB_CASES, ids=lambda c: c["name"])
def test_pcb_oracle_matches_reference(case):
    g = golden("pcb")
    base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"])
    vec, task, q, mx, lo, hi = orc.pcb_vectors(base, models, case["density"], return_task=True)
    ref = g[f"{case['name']}/vectors"]
    assert (np.abs(vec - ref) / np.abs(ref).max(axis=1, keepdims=True)).max() < 2e-6
    # the support is the top `density` fraction of the balancing weights, row by row
    d = case["d"]
    assert np.array_equal((ref != 0).sum(axis=1) > 0, np.ones(case["K"], bool))
    assert abs(int((vec[0] != 0).sum()) - (d - int(d * (1 - case["density"])))) <= 1
    merged = orc.merge_pcb(base, models, case["weights"], case["density"])
    tau_max = max(np.abs(m - base).max() for m in models)
    assert np.abs(merged - g[f"{case['name']}/merged"]).max() <= 1e-6 * tau_max


def test_sum_dim0_helper_is_torch_order():
    """K = 8, d = 37: the 5 trailing columns use ATen's 4-way interleaved order."""
    rng = np.random.default_rng(0)
    X = (rng.standard_normal((8, 37)) * 1e3).astype(np.float32)
    got = orc._sum_dim0(X)
    seq = np.zeros(37, np.float32)
    for k in range(8):
        seq = (seq + X[k]).astype(np.float32)
    assert np.array_equal(got[:32], seq[:32])
    a = [np.float32(0)] * 4
    for j in range(32, 37):
        a = [np.float32(0)] * 4
        for k in range(8):
            a[k % 4] = np.float32(a[k % 4] + X[k, j])
        assert got[j] == np.float32(np.float32(np.float32(a[0] + a[1]) + a[2]) + a[3])
