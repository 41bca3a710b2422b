"""GPU: PCB-merging vectors (csrc/pcb.cu + the TIES selection kernels for the magnitude clamps) against the numpy oracle
and the golden vectors of the reference's get_pcb_vectors / merge_pcb (tests/golden/pcb.npz).

Exact parts: the 1 % / 99 % magnitude clamps (integer order statistics) and the quantile of the balancing weights (checked
against a host sort of the kernel's own task values).  Floating-point part: exp / tanh come from different libms, so the
vectors are compared at 2e-6 of each row's largest magnitude, and columns holding a balancing weight within 1e-3 relative
of the clamp -- where pcb.py's `scale / max(sum(scale), 1e-12)` is discontinuous -- are excluded and counted."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from helpers import golden
from mergerec_b200 import synth
from mergerec_b200.merger import ModelMerger
from mergerec_b200.merger.algorithms.pcb import get_pcb_vectors, merge_pcb
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 2e-6


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def row_rel(a, b):
    return np.abs(np.asarray(a, np.float64) - b) / np.abs(b).max(axis=1, keepdims=True)


@pytest.mark.parametrize("case", gc.PCB_CASES, ids=lambda c: c["name"])
def test_pcb_vectors_match_oracle_and_reference(case):
    g = golden("pcb")
    base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"])
    K, d = case["K"], case["d"]
    out, task, thr, lo, hi = get_pcb_vectors(dev(base), [dev(m) for m in models], density=case["density"], return_diagnostics=True)
    out, task, thr = out[:, :d].cpu().numpy(), task[:, :d].cpu().numpy(), thr.cpu().numpy()
    o_vec, o_task, o_q, o_max, o_lo, o_hi = orc.pcb_vectors(base, models, case["density"], return_task=True)
    # exact: magnitude clamps, and the quantile / maximum of the kernel's own balancing weights
    assert np.array_equal(lo.cpu().numpy(), o_lo) and np.array_equal(hi.cpu().numpy(), o_hi)
    q_index = int(d * (1 - case["density"]))
    srt = np.sort(task, axis=1)
    assert np.array_equal(thr[:, 0], srt[:, q_index]) and np.array_equal(thr[:, 1], srt[:, -1])
    # floating point: balancing weights and thresholds to ~1 ulp
    assert (np.abs(task - o_task) <= 4e-7 * np.abs(o_task) + 1e-30).all()
    assert np.allclose(thr[:, 0], o_q, rtol=1e-6, atol=0) and np.allclose(thr[:, 1], o_max, rtol=1e-6, atol=0)
    near = (np.abs(o_task - o_q[:, None]) <= 1e-3 * np.abs(o_q[:, None])).any(axis=0)
    assert near.sum() <= max(0.01 * d, 2 * K)    # (each model's own threshold element is always "near": K columns when d is tiny)
    ref = g[f"{case['name']}/vectors"]
    assert row_rel(out, o_vec)[:, ~near].max() < TOL
    assert row_rel(out, ref)[:, ~near].max() < TOL
    # inside the excluded columns a value is either close as well or one of {0, the clamped update / K'} flipped
    assert np.isfinite(out).all()


@pytest.mark.parametrize("case", gc.PCB_CASES, ids=lambda c: c["name"])
def test_merge_pcb_matches_reference(case):
    g = golden("pcb")
    base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"])
    merged = merge_pcb(dev(base), [dev(m) for m in models], case["weights"], density=case["density"]).cpu().numpy()
    ref = g[f"{case['name']}/merged"]
    _, o_task, o_q, *_ = orc.pcb_vectors(base, models, case["density"], return_task=True)
    near = (np.abs(o_task - o_q[:, None]) <= 1e-3 * np.abs(o_q[:, None])).any(axis=0)
    tau_max = max(np.abs(m - base).max() for m in models)
    # merged weights: |ours - reference| below 1e-6 of the weight scale (north star: 1e-6 relative in fp32)
    assert (np.abs(merged - ref)[~near] <= 1e-6 * np.maximum(np.abs(ref[~near]), tau_max)).all()
    assert np.abs(merged - orc.merge_pcb(base, models, case["weights"], case["density"]))[~near].max() <= 1e-6 * tau_max


def test_model_merger_pcb_and_module_factory():
    """`ModelMerger.merge("pcb")` and `load_merging_module(MergeType.PCB, ...)` run end to end on a toy encoder."""
    from mergerec_b200.merger.enums import LearnType, MergeType
    from mergerec_b200.merger.weight_learning.module import load_merging_module
    from toy_model import ToyEncoder, make_toy_state_dicts
    pre, fts = make_toy_state_dicts(3, seed=5)
    to_t = lambda sd: {k: torch.from_numpy(np.asarray(v)) if not isinstance(v, torch.Tensor) else v for k, v in sd.items()}
    pre, fts = to_t(pre), [to_t(f) for f in fts]
    merger = ModelMerger(models=fts, base_model=pre)
    sd = merger.merge("pcb", [0.3, 0.3, 0.3], density=0.2)
    flat = torch.cat([v.reshape(-1).float() for v in sd.values()]).cpu().numpy()
    keys = list(sd.keys())
    b = np.concatenate([np.asarray(pre[k], np.float32).reshape(-1) for k in keys])
    ms = [np.concatenate([np.asarray(f[k], np.float32).reshape(-1) for k in keys]) for f in fts]
    want = orc.merge_pcb(b, ms, [0.3, 0.3, 0.3], 0.2)
    _, o_task, o_q, *_ = orc.pcb_vectors(b, ms, 0.2, return_task=True)
    near = (np.abs(o_task - o_q[:, None]) <= 1e-3 * np.abs(o_q[:, None])).any(axis=0)
    tau_max = max(np.abs(m - b).max() for m in ms)
    assert (np.abs(flat - want)[~near] <= 1e-6 * np.maximum(np.abs(want[~near]), tau_max)).all()
    torch.manual_seed(0)
    mod = load_merging_module(MergeType.PCB, LearnType.TASK_WISE, ToyEncoder(), pre, fts, ignore_keys=set(), ties_density=0.2)
    assert mod.task_vectors.shape[0] == 3 if hasattr(mod, "task_vectors") else True


def test_pcb_full_size_properties():
    """BLaIR-base size (d = 124,645,632, K = 8): exact quantile (count check on the kernel's own task values), support
    of the result = density, signs follow the updates, finite."""
    d, K, density = synth.total_numel(synth.roberta_shapes()), 8, 0.2
    g = torch.Generator(device="cuda").manual_seed(3)
    base = torch.randn(d, generator=g, device="cuda") * 0.02
    models = [base + 1e-3 * torch.randn(d, generator=g, device="cuda") for _ in range(K)]
    out, task, thr, lo, hi = get_pcb_vectors(base, models, density=density, return_diagnostics=True)
    q_index = int(d * (1 - density))
    for k in range(K):
        t = task[k, :d]
        below = int((t < thr[k, 0]).sum())
        at = int((t == thr[k, 0]).sum())
        assert below <= q_index < below + at                       # thr is exactly the q_index-th smallest value
        assert float(t.max()) == float(thr[k, 1])
        nnz = int((out[k, :d] != 0).sum())
        assert abs(nnz - (d - q_index)) <= at + 1
        tau = models[k] - base
        assert bool(((out[k, :d] == 0) | (torch.sign(out[k, :d]) == torch.sign(tau))).all())
    assert bool(torch.isfinite(out[:, :d]).all())


def test_fast_and_dense_quantile_paths_agree():
    """d large enough for the sampled + windowed search (d / 32 >= 4096): its quantile must be the exact order statistic
    (host sort of the kernel's own balancing weights) and the vectors must equal those of the three dense passes bit for
    bit."""
    K, d, density = 5, 600_011, 0.2
    base, models = synth.make_flat(d, K, seed=97)
    tb, tm = dev(base), [dev(m) for m in models]
    fast, task, thr, lo, hi = get_pcb_vectors(tb, tm, density=density, return_diagnostics=True)
    dense, task_d, thr_d, *_ = get_pcb_vectors(tb, tm, density=density, return_diagnostics=True, force_dense=True)
    assert torch.equal(thr, thr_d) and torch.equal(fast[:, :d], dense[:, :d])
    srt = np.sort(task[:, :d].cpu().numpy(), axis=1)
    assert np.array_equal(thr[:, 0].cpu().numpy(), srt[:, int(d * (1 - density))])
    assert np.array_equal(thr[:, 1].cpu().numpy(), srt[:, -1])
    for dens in (0.001, 0.9):            # quantiles in the tails
        a, _, ta, *_ = get_pcb_vectors(tb, tm, density=dens, return_diagnostics=True)
        b, _, tb2, *_ = get_pcb_vectors(tb, tm, density=dens, return_diagnostics=True, force_dense=True)
        assert torch.equal(ta, tb2) and torch.equal(a[:, :d], b[:, :d]), dens


def test_prepared_reciprocal_divisions_equal_ieee_divides():
    """The kernels divide by per-model ranges and per-column sums with a prepared reciprocal and two FMA corrections
    (csrc/pcb.cu: pcb_div_by); MR_PCB_IEEE compiles the same kernels with IEEE divides.  Both must agree bit for bit --
    balancing weights, thresholds and vectors -- on the fast and on the dense search, also for unaligned rows."""
    K, d = 5, 600_011
    base, models = synth.make_flat(d, K, seed=41)
    tb, tm = dev(base), [dev(m) for m in models]
    for force_dense in (False, True):
        a = get_pcb_vectors(tb, tm, density=0.2, return_diagnostics=True, force_dense=force_dense)
        b = get_pcb_vectors(tb, tm, density=0.2, return_diagnostics=True, force_dense=force_dense, force_ieee=True)
        for x, y in zip(a[:3], b[:3]):
            assert torch.equal(x[..., :d].view(torch.int32) if x.dim() == 2 and x.shape[1] >= d else x.view(torch.int32),
                               y[..., :d].view(torch.int32) if y.dim() == 2 and y.shape[1] >= d else y.view(torch.int32))
    # unaligned pointers take the scalar-load instantiation
    ub, um = dev(np.concatenate([[0], base]).astype(np.float32))[1:], [dev(np.concatenate([[0], m]).astype(np.float32))[1:] for m in models]
    a = get_pcb_vectors(ub, um, density=0.2)
    b = get_pcb_vectors(tb, tm, density=0.2)
    assert torch.equal(a[:, :d], b[:, :d])


def test_degenerate_ranges_fall_back_to_ieee_divides():
    """A clamp range outside [2^-60, 2^60] is reported by the kernels (status 2) and the call reruns with IEEE divides:
    tiny updates (|tau| ~ 1e-25) give the same vectors as the forced-IEEE path and stay finite."""
    K, d = 3, 200_003
    rng = np.random.default_rng(5)
    base = rng.standard_normal(d).astype(np.float32) * np.float32(1e-20)
    models = [(base + rng.standard_normal(d).astype(np.float32) * np.float32(1e-25)).astype(np.float32) for _ in range(K)]
    tb, tm = dev(base), [dev(m) for m in models]
    a = get_pcb_vectors(tb, tm, density=0.2)
    b = get_pcb_vectors(tb, tm, density=0.2, force_ieee=True)
    assert torch.equal(a[:, :d].view(torch.int32), b[:, :d].view(torch.int32))
