"""GPU parity: the CUDA merge kernels (through the C ABI) against the golden vectors of the unmodified
reference and against the CPU oracle.  Everything here is bit-exact (the kernels use unfused fp32
arithmetic in the reference's order)."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from helpers import assert_bit_equal, flatten_np, golden, shape_dict_of, state_dict_case
from mergerec_b200 import _lib, synth
from mergerec_b200.merger import ModelMerger
from mergerec_b200.merger.algorithms import get_task_vectors, merge_linear, merge_task_vector
from mergerec_b200.merger.algorithms._common import merge_axpy
from mergerec_b200.merger.layout import FlatLayout, alloc_rows
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("case", gc.MERGE_FLAT_CASES, ids=lambda c: c["name"])
def test_merge_flat_vs_golden(case):
    g = golden("merge_flat")
    base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"])
    tb, tm = dev(base), [dev(m) for m in models]
    assert_bit_equal(host(merge_task_vector(tb, tm, case["weights"])), g[f"{case['name']}/task_vector"], "A1")
    assert_bit_equal(host(merge_linear(models=tm, weights=case["weights"])), g[f"{case['name']}/linear"], "A10")
    assert_bit_equal(host(get_task_vectors(tb, tm)), g[f"{case['name']}/task_vectors"], "A2")


@pytest.mark.parametrize("shift", [1, 2, 3])
def test_merge_unaligned_pointers(shift):
    """Flat vectors that start off a 16-byte boundary take the scalar path; same bits."""
    K, d = 5, 5003
    base, models = synth.make_flat(d + 8, K, seed=77)
    w = [0.11 * (k + 1) for k in range(K)]
    tb = dev(base)[shift:shift + d]
    tm = [dev(m)[shift:shift + d] for m in models]
    want = orc.merge_task_vector(base[shift:shift + d], [m[shift:shift + d] for m in models], w)
    assert_bit_equal(host(merge_task_vector(tb, tm, w)), want, "A1 unaligned")
    assert_bit_equal(host(merge_linear(models=tm, weights=w)), orc.merge_linear([m[shift:shift + d] for m in models], w),
                     "A10 unaligned")


@pytest.mark.parametrize("case", gc.MODEL_MERGER_CASES, ids=lambda c: c["name"])
def test_model_merger_vs_golden(case):
    g = golden("model_merger")
    _, base, models = state_dict_case(case)
    tb = {k: torch.from_numpy(v) for k, v in base.items()}
    tm = [{k: torch.from_numpy(v) for k, v in m.items()} for m in models]
    merger = ModelMerger(tm, tb)
    keys = list(g[f"{case['name']}/keys"])
    assert list(merger.shape_dict.keys()) == keys
    assert_bit_equal(host(merger.base_model), g[f"{case['name']}/base_flat"], "A0 flatten")
    for mt, w, kw in case["merges"]:
        sd = merger.merge(mt, w, **kw)
        assert list(sd.keys()) == keys
        for k in keys:
            assert tuple(sd[k].shape) == tuple(base[k].shape)
        last = [m for m in case["merges"] if m[0] == mt][-1]
        if (mt, w, kw) == last:
            flat = torch.cat([v.reshape(-1) for v in sd.values()])
            assert_bit_equal(host(flat), g[f"{case['name']}/{mt}"], mt)


def test_model_merger_errors():
    shapes = synth.tiny_shapes()
    base, models = synth.make_state_dicts(shapes, 2, seed=3)
    tb = {k: torch.from_numpy(v) for k, v in base.items()}
    tm = [{k: torch.from_numpy(v) for k, v in m.items()} for m in models]
    merger = ModelMerger(tm, tb)
    with pytest.raises(ValueError):
        merger.merge("task_vector", 1)  # ints are rejected like the reference (merger.py:60-64)
    with pytest.raises(ValueError):
        merger.merge("nope", 0.3)
    with pytest.raises(AssertionError):
        merger.merge("task_vector", [0.3])  # length mismatch (task_vector.py:28)
    with pytest.raises(ValueError):
        ModelMerger(tm).merge("task_vector", 0.3)  # no base model (merger.py:72-74)
    bad = dict(tm[1])
    bad.pop(next(iter(bad)))
    with pytest.raises(AssertionError):
        ModelMerger([tm[0], bad], tb)


@pytest.mark.parametrize("case", gc.LAMBDA_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("learn", ["task", "layer"])
@pytest.mark.parametrize("softmax", [0, 1])
def test_lambda_merge_vs_golden(case, learn, softmax):
    """A3 / A4 with the reference's own effective lambdas; ragged blocks, Recformer's mod-4 == 2 offsets and
    the interleaved torch.sum tail (K >= 5) are all in these cases."""
    g = golden("lambda_merge")
    _, base, models = state_dict_case(case)
    layout = FlatLayout.from_shape_dict(shape_dict_of(base))
    tb = dev(flatten_np(base))
    tm = [dev(flatten_np(m)) for m in models]
    T = get_task_vectors(tb, tm)
    tag = f"{case['name']}/{learn}/softmax{softmax}"
    seg_end, seg_group, keys = layout.device_blocks(learn == "layer", tb.device)
    assert keys == list(g[f"{tag}/keys"])
    w = dev(g[f"{tag}/w"])
    if learn == "task":
        out = merge_axpy(tb, list(T.unbind(0)), w, _lib.MR_ORDER_SUM_FIRST, src_is_model=False)
    else:
        out = merge_axpy(tb, list(T.unbind(0)), w, _lib.MR_ORDER_SUM_FIRST, src_is_model=False, seg_end=seg_end,
                         seg_group=seg_group)
    assert_bit_equal(host(out), g[f"{tag}/merged"], "A3/A4 merged")


@pytest.mark.parametrize("K", [1, 2, 4, 5, 7, 12, 16])
def test_lambda_merge_all_k_vs_oracle(K):
    d = 32 * 61 + 19  # 19-column interleaved tail
    base, models = synth.make_flat(d, K, seed=100 + K)
    T = orc.task_vectors(base, models)
    rng = np.random.Generator(np.random.PCG64(K))
    w = rng.uniform(-0.5, 0.9, size=(1, K)).astype(np.float32)
    rows = alloc_rows(K, d, "cuda")
    rows.copy_(dev(T))
    out = merge_axpy(dev(base), list(rows.unbind(0)), dev(w), _lib.MR_ORDER_SUM_FIRST, src_is_model=False)
    assert_bit_equal(host(out), orc.lambda_merge(base, T, w), f"A3 K={K}")


def test_merge_empty_and_bad_args():
    e = torch.empty(0, dtype=torch.float32, device="cuda")
    out = merge_task_vector(e, [e, e], [0.5, 0.5])
    assert out.numel() == 0
    x = torch.zeros(8, dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError):
        merge_task_vector(x, [x] * 17, [0.1] * 17)  # K > MR_MAX_K


@pytest.mark.parametrize("K,order", [(3, "base_first"), (8, "sum_first")])
def test_full_size_blair_base(K, order):
    """BASELINE config sizes (d = 124,645,632): bit-exact against the oracle at full size."""
    d = synth.total_numel(synth.roberta_shapes())
    assert d == 124_645_632
    g = torch.Generator(device="cuda").manual_seed(5)
    base = torch.randn(d, generator=g, device="cuda") * 0.02
    models = [base + 1e-3 * torch.randn(d, generator=g, device="cuda") for _ in range(K)]
    hb, hm = host(base), [host(m) for m in models]
    if order == "base_first":
        w = [0.3] * K
        got = host(merge_task_vector(base, models, w))
        want = orc.merge_task_vector(hb, hm, w)
    else:
        layout = FlatLayout.from_shape_dict(synth.roberta_shapes())
        seg_end, seg_group, keys = layout.device_blocks(True, base.device)
        assert len(keys) == 13
        rng = np.random.Generator(np.random.PCG64(5))
        w = rng.uniform(0.1, 0.5, size=(13, K)).astype(np.float32)
        T = get_task_vectors(base, models)
        got = host(merge_axpy(base, list(T.unbind(0)), dev(w), _lib.MR_ORDER_SUM_FIRST, False, seg_end, seg_group))
        sb, se, sg, _ = orc.segment_table(synth.roberta_shapes(), layer_wise=True)
        want = orc.lambda_merge(hb, orc.task_vectors(hb, hm), w, sb, se, sg)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_lambda_merge_many_blocks_needs_large_shared_memory():
    """A layer-wise table of 6,000 tensors (72 KB of block table, above the 48 KB default) must launch -- the kernel opts
    in to the larger dynamic shared memory -- and stay bit-exact; a table that cannot fit 224 KB is refused with a clear
    argument error instead of a launch failure."""
    P, K, G = 6000, 3, 7
    rng = np.random.Generator(np.random.PCG64(21))
    sizes = rng.integers(1, 40, size=P)
    seg_end = np.cumsum(sizes).astype(np.int64)
    seg_begin = seg_end - sizes
    seg_group = rng.integers(0, G, size=P).astype(np.int32)
    d = int(seg_end[-1])
    base, models = synth.make_flat(d, K, seed=5)
    T = orc.task_vectors(base, models)
    w = rng.uniform(0.1, 0.5, size=(G, K)).astype(np.float32)
    rows = alloc_rows(K, d, "cuda")
    rows.copy_(dev(T))
    out = merge_axpy(dev(base), list(rows.unbind(0)), dev(w), _lib.MR_ORDER_SUM_FIRST, False, dev(seg_end), dev(seg_group))
    assert_bit_equal(host(out), orc.lambda_merge(base, T, w, seg_begin, seg_end, seg_group), "P = 6000 blocks")
    P2 = 20000
    with pytest.raises(ValueError, match="does not fit"):
        merge_axpy(dev(base), list(rows.unbind(0)), dev(w), _lib.MR_ORDER_SUM_FIRST, False,
                   dev(np.arange(1, P2 + 1, dtype=np.int64)), dev(np.zeros(P2, np.int32)))
