"""Randomised parity (hypothesis) of the merger kernels against the CPU oracle: random K in [1, 16], flat lengths that
are not multiples of 4 / 32, pointers off the 16-byte grid, ragged block tables with tiny blocks, special values
(+-0, denormals, huge magnitudes), injected magnitude ties and edge densities ("property")."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, example, given, settings, strategies as st

from helpers import assert_bit_equal
from mergerec_b200 import _lib
from mergerec_b200.merger.algorithms import (get_localize_and_stitch_vectors, get_ties_vectors, merge_linear,
                                             merge_task_vector, merge_ties)
from mergerec_b200.merger.algorithms._common import merge_axpy
from mergerec_b200.merger.algorithms.ties import merge_ties_lambda
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
import os

# MR_HYPOTHESIS_EXAMPLES=1000 turns the suite into a soak test (the default keeps `pytest -m gpu` under a minute)
# Without the variable the examples are derived from the test body (derandomize): `pytest -m gpu` is reproducible; the
# soak run draws fresh random examples.
SETTINGS = dict(deadline=None, max_examples=int(os.environ.get("MR_HYPOTHESIS_EXAMPLES", "40")),
                derandomize="MR_HYPOTHESIS_EXAMPLES" not in os.environ,
                suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def make_inputs(rng, K, d, specials, quant):
    base = (rng.standard_normal(d) * 0.02).astype(np.float32)
    models = []
    for _ in range(K):
        tv = (rng.standard_normal(d) * 1e-3).astype(np.float32)
        if quant:                      # coarse grid -> many exactly equal magnitudes
            tv = (np.round(tv / np.float32(5e-4)) * np.float32(5e-4)).astype(np.float32)
        models.append((base + tv).astype(np.float32))
    if specials and d >= 8:
        idx = rng.choice(d, size=min(8, d), replace=False)
        vals = np.array([0.0, -0.0, 1e-42, -1e-42, 3e30, -3e30, 1.0, -1.0], np.float32)[: len(idx)]
        for m in models[: max(1, K // 2)]:
            m[idx] = vals
        base[idx[:2]] = np.array([-0.0, 0.0], np.float32)[: len(idx[:2])]
    return base, models


def shifted(arr, shift, d):
    """A device view of length d starting `shift` floats off a 16-byte boundary."""
    buf = torch.zeros(d + 8, dtype=torch.float32, device="cuda")
    buf[shift:shift + d] = dev(arr)
    return buf[shift:shift + d]


def random_blocks(rng, d):
    """Ragged block table: tiny, odd-sized and large blocks in random order."""
    ends, pos = [], 0
    while pos < d:
        pos = min(d, pos + int(rng.choice([1, 2, 3, 5, 31, 32, 33, 64, 100, 257, 1000])))
        ends.append(pos)
    G = int(rng.integers(1, 6))
    groups = rng.integers(0, G, size=len(ends)).astype(np.int32)
    groups[0] = G - 1                  # every group id below G may be unused except the largest: still valid
    return np.asarray(ends, np.int64), groups, G


@settings(**SETTINGS)
@given(K=st.integers(1, 16), d=st.integers(1, 3000), shift=st.integers(0, 3), seed=st.integers(0, 2 ** 31 - 1),
       specials=st.booleans())
def test_task_vector_and_linear_merge(K, d, shift, seed, specials):
    rng = np.random.Generator(np.random.PCG64(seed))
    base, models = make_inputs(rng, K, d, specials, False)
    w = [float(x) for x in rng.uniform(-1.0, 1.0, size=K)]
    tb, tm = shifted(base, shift, d), [shifted(m, shift, d) for m in models]
    assert_bit_equal(host(merge_task_vector(tb, tm, w)), orc.merge_task_vector(base, models, w), "merge_task_vector")
    assert_bit_equal(host(merge_linear(models=tm, weights=w)), orc.merge_linear(models, w), "merge_linear")


@settings(**SETTINGS)
@given(K=st.integers(1, 16), d=st.integers(8, 3000), seed=st.integers(0, 2 ** 31 - 1), layer_wise=st.booleans())
def test_lambda_merge_ragged_blocks(K, d, seed, layer_wise):
    """Blocks shorter than 8 take another ATen reduction path in the reference: only blocks the
    oracle models (n >= 8 or K <= 4) are generated for K >= 5 by merging tiny blocks into their neighbours."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base, models = make_inputs(rng, K, d, False, False)
    T = orc.task_vectors(base, models)
    if layer_wise:
        ends, groups, G = random_blocks(rng, d)
        if K >= 5:                     # drop block boundaries that would create blocks shorter than 8
            keep, last = [], 0
            for e in ends:
                if e - last >= 8 or e == d:
                    keep.append(e)
                    last = e
            if len(keep) >= 2 and keep[-1] - keep[-2] < 8:
                keep.pop(-2)
            ends = np.asarray(keep, np.int64)
            groups = groups[: len(ends)].copy()
            groups[0] = G - 1
        begins = np.concatenate([[0], ends[:-1]]).astype(np.int64)
    else:
        ends, groups, G, begins = np.asarray([d], np.int64), np.zeros(1, np.int32), 1, np.zeros(1, np.int64)
    w = rng.uniform(-0.5, 0.5, size=(G, K)).astype(np.float32)
    got = merge_axpy(dev(base), [dev(t) for t in T], dev(w), _lib.MR_ORDER_SUM_FIRST, False,
                     dev(ends) if layer_wise else None, dev(groups) if layer_wise else None)
    assert_bit_equal(host(got), orc.lambda_merge(base, T, w, begins, ends, groups), "lambda merge")


@settings(**SETTINGS)
@given(K=st.integers(1, 16), d=st.integers(40, 4000), shift=st.integers(0, 3), seed=st.integers(0, 2 ** 31 - 1),
       density=st.sampled_from([0.0, 0.001, 0.05, 0.2, 0.5, 0.999, 1.0]), quant=st.booleans(), specials=st.booleans())
@example(K=6, d=40, shift=0, seed=0, density=0.999, quant=False, specials=True)   # kept -0.0 updates under a minus election
def test_ties_family(K, d, shift, seed, density, quant, specials):
    rng = np.random.Generator(np.random.PCG64(seed))
    base, models = make_inputs(rng, K, d, specials, quant)
    tb, tm = shifted(base, shift, d), [shifted(m, shift, d) for m in models]
    That, trim, elect, cut = get_ties_vectors(tb, tm, density, return_masks=True)
    oT, otrim, oelect, ocut = orc.ties_vectors(base, models, density, return_masks=True)
    assert np.array_equal(host(cut).view(np.uint64), ocut), "cut keys"
    assert np.array_equal(host(trim), otrim) and np.array_equal(host(elect), oelect), "masks"
    assert_bit_equal(host(That), oT, "TIES vectors")
    assert_bit_equal(host(get_ties_vectors(tb, tm, density)), oT, "TIES vectors, one-pass select + build")
    w = [float(x) for x in rng.uniform(0.1, 1.0, size=K)]
    assert_bit_equal(host(merge_ties(tb, tm, w, density)), orc.merge_ties(base, models, w, density), "merge_ties")
    lam = rng.uniform(0.1, 0.5, size=(1, K)).astype(np.float32)
    fused = merge_ties_lambda(tb, tm, density, dev(lam))
    assert_bit_equal(host(fused), orc.lambda_merge(base, oT, lam), "fused TIES + lambda merge")
    assert_bit_equal(host(merge_ties_lambda(tb, tm, density, dev(lam), one_pass=False)), orc.lambda_merge(base, oT, lam),
                     "fused TIES + lambda merge, two-pass form")
    assert_bit_equal(host(get_localize_and_stitch_vectors(tb, tm, density)), orc.lns_vectors(base, models, density),
                     "localize-and-stitch vectors")
    # the sharded merger's select (per-slice estimate + windowed radix levels + tie scan) on a single rank
    from mergerec_b200.merger.sharded import get_ties_vectors_sharded, merge_ties_sharded
    assert_bit_equal(host(get_ties_vectors_sharded(tb, tm, density, d, None))[:, :d], oT, "sharded-path TIES vectors")
    assert_bit_equal(host(merge_ties_sharded(tb, tm, w, density, d, None)), orc.merge_ties(base, models, w, density),
                     "sharded-path merge_ties")


@settings(**dict(SETTINGS, max_examples=max(30, SETTINGS["max_examples"] // 4)))
@given(Q=st.integers(1, 600), N=st.integers(1, 3000), e4=st.integers(1, 33), k=st.integers(1, 128),
       id_base=st.integers(0, 2_000_000_000), seed=st.integers(0, 2 ** 31 - 1), cg=st.sampled_from([1, 2]),
       mode=st.sampled_from([0, 1, 2]), splits=st.sampled_from([0, 1, 2, 5]))
def test_fused_scoring_topk(Q, N, e4, k, id_base, seed, cg, mode, splits):
    """Fused tensor-core scoring + top-K on exact-grid embeddings (every summation order gives the same fp32 score, ties
    are plentiful): ids and scores bit-exact vs the oracle for ragged Q / N / E, every mode, both CTA-group variants,
    forced split counts and arbitrary shard offsets."""
    import os
    from mergerec_b200 import synth
    from mergerec_b200.evaluator import ShardedItemTable
    from mergerec_b200.evaluator.evaluator import score_topk
    from mergerec_b200.evaluator.sharded import split_tf32, to_bf16
    E = 4 * e4 if mode != 2 else 8 * ((e4 + 1) // 2)
    k = min(k, N)
    id_base = min(id_base, 2 ** 31 - 1 - N - 512)
    users, items, _ = synth.make_catalog(Q, N, E, kind="grid", seed=seed % 100000)
    os.environ["MR_SCORE_CTA_GROUP"], os.environ["MR_SCORE_SPLITS"] = str(cg), str(splits)
    try:
        table = ShardedItemTable(dev(items), id_base=id_base, bf16=(mode == 2))
        u_hi, u_lo = (to_bf16(dev(users)), None) if mode == 2 else split_tf32(dev(users))
        v, i = score_topk(u_hi, u_lo, table, k, mode)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("MR_SCORE_CTA_GROUP", None)
        os.environ.pop("MR_SCORE_SPLITS", None)
    scores = orc.scores_bf16(users, items) if mode == 2 else orc.scores_f32(users, items)
    ov, oi = orc.topk_rows(scores, k, id_base=id_base)
    assert np.array_equal(host(i), oi), "ids"
    assert_bit_equal(host(v), ov, "scores")


_DISTILL_LOSSES = [("CE", {}), ("KD", dict(temperature=2.0)), ("KD", dict(temperature=0.1)), ("MSE", {}), ("ADAMERGING", {}),
                   ("ADAMERGING_KD", dict(temperature=0.5, coefficient=0.3)), ("SINGLE_PSEUDO_LABEL_KD", dict(temperature=1.0, coefficient=0.5)),
                   ("LISTNET", dict(temperature=0.3))]


@settings(**dict(SETTINGS, max_examples=max(25, SETTINGS["max_examples"] // 8)))
@given(B=st.integers(1, 20), e4=st.integers(1, 256), D=st.integers(1, 5), seed=st.integers(0, 2 ** 31 - 1),
       loss=st.sampled_from(_DISTILL_LOSSES), big=st.booleans())
@example(B=1, e4=222, D=1, seed=2_147_483_646, loss=("KD", {"temperature": 0.1}), big=False)   # student ~ teacher: gradient cancels
def test_distill_step(B, e4, D, seed, loss, big):
    """Distillation step (logits -> loss -> representation gradient) on ragged tables, every embedding width that is a
    multiple of 4 up to 1024, 1..20 samples spread over 1..5 domains (so groups of 1..4 samples and repeated passes over
    a table): 2e-5 of the largest magnitude against the fp64 oracle.  (The losses whose label is an argmax of the merged
    model's own logits, and the hinge, are discontinuous in the logits and stay with the golden-vector tests.)"""
    from mergerec_b200.module.distiller.sequence.module import fused_distill_losses
    from test_distill_gpu import loss_object
    rng = np.random.Generator(np.random.PCG64(seed))
    E = 4 * e4
    rows = [int(rng.integers(1, 2000 if big else 200)) for _ in range(D)]
    tables = [(rng.standard_normal((n, E)) / np.sqrt(E)).astype(np.float32) for n in rows]
    dom = [int(x) for x in rng.integers(0, D, size=B)]
    rep = rng.standard_normal((B, E)).astype(np.float32)
    teacher = [(tables[d] @ rep[b] + 0.5 * rng.standard_normal(rows[d])).astype(np.float32) for b, d in enumerate(dom)]
    lname, kw = loss
    spec = loss_object(lname, kw).spec
    t_dev = [dev(t) for t in teacher]
    r = dev(rep).requires_grad_(True)
    losses = fused_distill_losses(r, [dev(t) for t in tables], dom, [t.data_ptr() for t in t_dev], spec)
    losses.mean().backward()
    o_losses, _, o_grad = orc.distill_step(rep, tables, dom, teacher, lname, **kw)
    scale_l = max(np.abs(o_losses).max(), 1e-3)
    assert np.abs(host(losses) - o_losses).max() <= 2e-5 * scale_l, (lname, rows, dom)
    # the gradient is sum_n gz[n] * items[n] with |gz| up to ~max(1, T, 1/T): when student and teacher nearly agree it
    # cancels to ~0, so the tolerance is anchored to the size of the terms as well as to the result
    T = kw.get("temperature", 1.0)
    term = max(1.0, T, 1.0 / T) * max(np.abs(t).max() for t in tables) / B
    assert np.abs(host(r.grad) - o_grad).max() <= 2e-5 * np.abs(o_grad).max() + 2e-7 * term, (lname, rows, dom)


@settings(**dict(SETTINGS, max_examples=max(20, SETTINGS["max_examples"] // 10)))
@given(K=st.integers(2, 16), d=st.integers(500, 6000), seed=st.integers(0, 2 ** 31 - 1),
       density=st.sampled_from([0.05, 0.2, 0.5, 0.9]), shift=st.integers(0, 3))
def test_pcb_vectors(K, d, seed, density, shift):
    """PCB vectors on random K / d / density and unaligned pointers: exact clamps and quantile, vectors within 2e-6 of the
    row scale outside the columns whose balancing weight sits within 1e-3 relative of the clamp (see tests/test_pcb_gpu.py)."""
    from hypothesis import assume
    from mergerec_b200.merger.algorithms.pcb import get_pcb_vectors
    rng = np.random.Generator(np.random.PCG64(seed))
    base, models = make_inputs(rng, K, d, False, False)
    o_vec, o_task, o_q, o_max, o_lo, o_hi = orc.pcb_vectors(base, models, density, return_task=True)
    assume(np.isfinite(o_vec).all() and (o_hi > o_lo).all() and (o_max > o_q).all())
    tb, tm = shifted(base, shift, d), [shifted(m, shift, d) for m in models]
    out, task, thr, lo, hi = get_pcb_vectors(tb, tm, density=density, return_diagnostics=True)
    out, task, thr = host(out)[:, :d], host(task)[:, :d], host(thr)
    assert np.array_equal(host(lo), o_lo) and np.array_equal(host(hi), o_hi)
    srt = np.sort(task, axis=1)
    assert np.array_equal(thr[:, 0], srt[:, int(d * (1 - density))]) and np.array_equal(thr[:, 1], srt[:, -1])
    near = (np.abs(o_task - o_q[:, None]) <= 1e-3 * np.abs(o_q[:, None])).any(axis=0)
    err = np.abs(out - o_vec) / np.abs(o_vec).max(axis=1, keepdims=True)
    assert err[:, ~near].max() < 2e-6
