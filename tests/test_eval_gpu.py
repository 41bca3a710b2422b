"""GPU parity for the evaluator (B1-B4): per-row top-K, shard merge, label rank, the TF32 split, plain-fp32 scoring
and the fused tensor-core scoring + top-K kernel, all through the C ABI, against the CPU oracle and the golden
vectors produced by the unmodified reference (tests/golden/evaluator.npz).

Bit-exact bar: ids and metric floats.  On exact-grid embeddings every summation order yields the same fp32 score,
so ids must equal the canonical (score desc, id asc) ranking of the reference's own CPU scores.  On Gaussian
embeddings ids may differ from an fp64 ranking only where fp64 itself shows a near-tie (bound in the test)."""
import os

import numpy as np
import pytest
import torch

import golden_cases as gc
from helpers import assert_bit_equal, golden
from mergerec_b200 import _lib, synth
from mergerec_b200.evaluator import Evaluator, NDCG, Recall, ShardedItemTable, shard_bounds
from mergerec_b200.evaluator.evaluator import score_topk, topk_rows
from mergerec_b200.evaluator.metrics import label_rank
from mergerec_b200.evaluator.sharded import MR_SCORE_BF16, MR_SCORE_TF32X1, MR_SCORE_TF32X3, split_tf32, topk_merge
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------ B2: topk_rows
@pytest.mark.parametrize("case", gc.EVAL_CASES, ids=lambda c: c["name"])
def test_topk_rows_and_metrics_vs_golden(case):
    g = golden("evaluator")
    users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
    kmax = max(case["ks"])
    scores = orc.scores_f32(users, items)
    vals, ids = topk_rows(dev(scores), kmax)
    ov, oi = orc.topk_rows(scores, kmax)
    assert np.array_equal(host(ids), oi), "ids vs oracle on the same score matrix"
    assert_bit_equal(host(vals), ov, "values vs oracle")
    ev = Evaluator(case["metrics"], case["ks"])
    res = ev(dev(scores), dev(labels), metric_prefix=case["prefix"])
    if case["kind"] == "grid":
        assert np.array_equal(host(ids), g[f"{case['name']}/canon_topk"]), "ids vs reference scores + stable sort"
        assert list(res.keys()) == list(g[f"{case['name']}/canon_keys"])
        assert np.array_equal(np.asarray(list(res.values()), np.float64), g[f"{case['name']}/canon_values"])
    else:
        want = orc.evaluate_ids(oi, labels, case["metrics"], case["ks"], case["prefix"])
        assert res == want
        assert list(res.keys()) == list(g[f"{case['name']}/raw_keys"])


def test_topk_rows_cpu_input_strided_and_edge_values():
    rng = np.random.Generator(np.random.PCG64(3))
    s = rng.standard_normal((37, 1003)).astype(np.float32)
    s[0, :50] = 9.0                       # a long run of ties: lowest ids first
    s[1, 7] = np.nan                      # NaN sorts first (torch.topk)
    s[2, 5], s[2, 900] = 0.0, -0.0        # -0 == +0: the lower id wins
    s[2, :5] = -1.0
    s[2, 6:900] = -1.0
    s[2, 901:] = -1.0
    s[3, :] = -np.inf
    s[4, 11] = np.inf
    wide = np.zeros((37, 1100), np.float32)
    wide[:, :1003] = s
    t = dev(wide)[:, :1003]               # row stride 1100 (not 16-byte aligned rows for odd q)
    for k in (1, 10, 100, 1003):
        v, i = topk_rows(t, k)
        ov, oi = orc.topk_rows(s, k)
        assert np.array_equal(host(i), oi), f"k={k}"
        assert np.array_equal(host(v).view(np.uint32), np.take_along_axis(s, oi.astype(np.int64), 1).view(np.uint32))
    v, i = topk_rows(torch.from_numpy(s), 10)   # host tensor in -> copied in chunks
    assert np.array_equal(host(i), orc.topk_rows(s, 10)[1])
    assert host(i)[0].tolist() == list(range(10)) and host(i)[1, 0] == 7 and host(i)[2, 0] == 5
    with pytest.raises(RuntimeError):
        topk_rows(t, 1004)
    e_v, e_i = topk_rows(torch.empty((0, 50), device="cuda"), 5)
    assert e_v.shape == (0, 5) and e_i.shape == (0, 5)


def test_topk_rows_matches_torch_topk_values():
    torch.manual_seed(0)
    s = torch.randn(300, 20000, device="cuda")
    v, i = topk_rows(s, 50)
    tv, ti = torch.topk(s, 50, dim=1)
    assert torch.equal(v, tv)
    assert torch.equal(i.long(), ti)      # continuous scores: no ties


# ------------------------------------------------------------------------------------------ merge / rank / split
def test_topk_merge_equals_unsharded():
    users, items, labels = synth.make_catalog(70, 5000, 24, kind="grid", seed=9)
    scores = orc.scores_f32(users, items)
    for k in (1, 20, 100):
        v_all, i_all = orc.topk_rows(scores, k)
        for world in (2, 3, 8):
            pv, pi = [], []
            for r in range(world):
                a, b = shard_bounds(5000, world, r)
                v, i = topk_rows(dev(np.ascontiguousarray(scores[:, a:b])), k, id_base=a)
                pv.append(v)
                pi.append(i)
            mv, mi = topk_merge(torch.stack(pv), torch.stack(pi), k)
            assert np.array_equal(host(mi), i_all) and np.array_equal(host(mv), v_all)
            ov, oi = orc.topk_merge(host(torch.stack(pv)), host(torch.stack(pi)))
            assert np.array_equal(host(mi), oi)


def test_topk_merge_empty_slots():
    vals = torch.tensor([[[3.0, 1.0, 0.0]], [[2.0, 2.0, -1.0]]], device="cuda")
    ids = torch.tensor([[[5, 9, -1]], [[7, 6, -1]]], dtype=torch.int32, device="cuda")
    v, i = topk_merge(vals, ids, 5)
    assert host(i)[0].tolist() == [5, 6, 7, 9, -1]
    assert host(v)[0, :4].tolist() == [3.0, 2.0, 2.0, 1.0] and np.isneginf(host(v)[0, 4])


def test_label_rank_and_metric_objects():
    g = golden("evaluator")
    case = gc.EVAL_CASES[4]   # grid_top100
    users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
    canon = g[f"{case['name']}/canon_topk"]
    rk = host(label_rank(dev(canon), dev(labels)))
    assert np.array_equal(rk, orc.label_rank(canon, labels))
    keys, values = list(g[f"{case['name']}/canon_keys"]), g[f"{case['name']}/canon_values"]
    for m in case["metrics"]:
        for k in case["ks"]:
            obj = {"RECALL": Recall, "NDCG": NDCG}[m](k)
            got = obj(y_true=torch.from_numpy(labels), y_pred=torch.from_numpy(canon.astype(np.int64)))
            assert got == values[keys.index(case["prefix"] + obj.name)], obj.name
    # labels outside the catalog / absent from the list
    ids = dev(np.array([[4, 2, 9], [1, 1, 1]], np.int32))
    assert host(label_rank(ids, dev(np.array([9, 77], np.int64)))).tolist() == [2, -1]


def test_split_tf32_bits():
    rng = np.random.Generator(np.random.PCG64(4))
    x = np.concatenate([rng.standard_normal(5000).astype(np.float32), np.array([0.0, -0.0, 1.0, 3.0e38, -7.5], np.float32)])
    hi, lo = split_tf32(dev(x))
    hi, lo = host(hi), host(lo)
    assert ((hi.view(np.uint32) & 0x1FFF) == 0).all() and ((lo.view(np.uint32) & 0x1FFF) == 0).all(), "tf32 grid"
    # rna = round to nearest, ties away from zero, on the 13 dropped mantissa bits
    u = x.view(np.uint32).astype(np.uint64)
    want_hi = (((u + 0x1000) & 0xFFFFE000) & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
    fin = np.isfinite(want_hi)
    assert np.array_equal(hi[fin].view(np.uint32), want_hi[fin].view(np.uint32))
    rel = np.abs((hi.astype(np.float64) + lo.astype(np.float64)) - x.astype(np.float64))
    big = np.abs(x) > 1e-30
    assert (rel[big & fin] <= np.abs(x[big & fin]).astype(np.float64) * 2.0 ** -21).all()


def scores_fp32(users, items, ldo=None):
    lib = _lib.load()
    du, di = dev(users), dev(items)
    Q, E = users.shape
    N = items.shape[0]
    ldo = N if ldo is None else ldo
    out = torch.zeros((Q, ldo), device="cuda")
    _lib.check(lib.mr_scores_fp32(_lib.dptr(du), Q, _lib.dptr(di), N, E, _lib.dptr(out), ldo, _lib.stream_handle()),
               "mr_scores_fp32")
    torch.cuda.synchronize()
    return host(out[:, :N])


def test_scores_fp32_kernel():
    users, items, _ = synth.make_catalog(130, 777, 48, kind="grid", seed=12)
    assert_bit_equal(scores_fp32(users, items, ldo=800), orc.scores_f32(users, items), "grid scores are exact in any order")
    ug, ig, _ = synth.make_catalog(64, 300, 100, kind="gauss", seed=13)
    assert np.abs(scores_fp32(ug, ig).astype(np.float64) - orc.scores_f64(ug, ig)).max() < 2e-6   # fp32 accumulation


# ------------------------------------------------------------------------------------------ B1+B2 fused (tcgen05)
def fused(users, items, k, mode=MR_SCORE_TF32X3, id_base=0):
    table = ShardedItemTable(dev(items), id_base=id_base)
    u_hi, u_lo = split_tf32(dev(users))
    v, i = score_topk(u_hi, u_lo, table, k, mode)
    torch.cuda.synchronize()
    return host(v), host(i)


FUSED_SHAPES = [
    # Q, N, E, k: ragged everything (partial query block, partial item tile, E not a multiple of 32)
    (64, 500, 32, 10), (200, 3000, 64, 100), (257, 1025, 36, 50), (1, 256, 8, 1), (300, 70000, 128, 128),
]


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("shape", FUSED_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_fused_grid_bit_exact(shape, cg, monkeypatch):
    monkeypatch.setenv("MR_SCORE_CTA_GROUP", str(cg))
    Q, N, E, k = shape
    users, items, _ = synth.make_catalog(Q, N, E, kind="grid", seed=21)
    scores = orc.scores_f32(users, items)
    ov, oi = orc.topk_rows(scores, min(k, N), id_base=1000)
    for mode in (MR_SCORE_TF32X3, MR_SCORE_TF32X1):      # grid values are exact in TF32: both modes are exact
        v, i = fused(users, items, min(k, N), mode, id_base=1000)
        assert np.array_equal(i, oi), f"ids, mode {mode}"
        assert_bit_equal(v, ov, f"scores, mode {mode}")


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("splits", [1, 3, 0])
def test_fused_golden_cfg1(cg, splits, monkeypatch):
    """BASELINE config 1 evaluator: Q=256, N=20,000, E=768, Recall@10 / NDCG@10 on grid (bit-exact vs the reference)."""
    monkeypatch.setenv("MR_SCORE_CTA_GROUP", str(cg))
    monkeypatch.setenv("MR_SCORE_SPLITS", str(splits))
    g = golden("evaluator")
    case = next(c for c in gc.EVAL_CASES if c["name"] == "grid_cfg1")
    users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
    ev = Evaluator(case["metrics"], case["ks"])
    _, ids = ev.topk_embeddings(dev(users), dev(items))
    assert np.array_equal(host(ids), g["grid_cfg1/canon_topk"])
    res = ev.evaluate_embeddings(dev(users), dev(items), dev(labels), metric_prefix=case["prefix"])
    assert list(res.keys()) == list(g["grid_cfg1/canon_keys"])
    assert np.array_equal(np.asarray(list(res.values()), np.float64), g["grid_cfg1/canon_values"])


def near_tie_rows(ids, users, items, k):
    """Rows whose ids differ from the fp64 ranking; a difference is explained when the two items swapped at that
    position have fp64 scores closer than 4 E 2^-24 |u||i|."""
    s64 = orc.scores_f64(users, items)
    order = np.argsort(-s64, axis=1, kind="stable")[:, :k]
    bad = np.unique(np.argwhere(order != ids)[:, 0])
    bound = 4 * users.shape[1] * 2.0 ** -24 * np.linalg.norm(users, axis=1).max() * np.linalg.norm(items, axis=1).max()
    unexplained = []
    for r in bad:
        pos = np.argwhere(order[r] != ids[r])[:, 0]
        if not all(abs(s64[r, ids[r, p]] - s64[r, order[r, p]]) < bound for p in pos):
            unexplained.append(int(r))
    return bad, unexplained


@pytest.mark.parametrize("cg", [1, 2])
def test_fused_gauss_fp32_faithful(cg, monkeypatch):
    monkeypatch.setenv("MR_SCORE_CTA_GROUP", str(cg))
    g = golden("evaluator")
    case = next(c for c in gc.EVAL_CASES if c["name"] == "gauss_cfg1")
    users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
    v, i = fused(users, items, 10)
    s64 = orc.scores_f64(users, items)
    err = np.abs(v.astype(np.float64) - np.take_along_axis(s64, i.astype(np.int64), 1)).max()
    # measured on B200: 7.2e-6 at scores ~0.9 (fp32 sgemm: 1.4e-7).  The operands are split exactly; the residual is
    # the tensor core's truncating fp32 accumulation, ~0.5 ulp(score) per tcgen05.mma (288 per dot product at E=768).
    print(f"3xTF32 max abs score error vs fp64: {err:.3e}")
    assert err < 2e-5, f"3xTF32 score error {err:.3e}"
    bad, unexplained = near_tie_rows(i, users, items, 10)
    assert not unexplained, f"rows {unexplained} reorder without an fp64 near-tie"
    print(f"rows whose ids differ from the fp64 ranking: {len(bad)} of {len(i)} (all at fp64 near-ties)")
    assert len(bad) <= 8
    # the reference's own ids (fp32 CPU sgemm + topk) on this tie-free case, and its metric floats
    ref_ids = g["gauss_cfg1/raw_topk"]
    assert (ref_ids != i).any(axis=1).sum() <= 8
    res = Evaluator(case["metrics"], case["ks"]).evaluate_embeddings(dev(users), dev(items), dev(labels))
    assert list(res.keys()) == list(g["gauss_cfg1/raw_keys"])
    if not (ref_ids != i).any():
        assert np.array_equal(np.asarray(list(res.values()), np.float64), g["gauss_cfg1/raw_values"])
    # plain 1xTF32 is visibly worse -- that is why it is not the default
    v1, _ = fused(users, items, 10, MR_SCORE_TF32X1)
    assert np.abs(v1.astype(np.float64) - np.sort(s64, axis=1)[:, ::-1][:, :10]).max() > 3 * err


def test_fused_equals_fp32_kernel_plus_topk_rows_at_scale():
    """Larger than the oracle likes: compare the tensor-core path with mr_scores_fp32 + mr_topk_rows on the GPU
    (grid inputs: both are exact), many query blocks and splits."""
    Q, N, E, k = 1500, 150000, 256, 100
    users, items, _ = synth.make_catalog(Q, N, E, kind="grid", seed=31)
    lib = _lib.load()
    du, di = dev(users), dev(items)
    scores = torch.empty((Q, N), device="cuda")
    _lib.check(lib.mr_scores_fp32(_lib.dptr(du), Q, _lib.dptr(di), N, E, _lib.dptr(scores), N, _lib.stream_handle()), "scores")
    rv, ri = topk_rows(scores, k)
    v, i = Evaluator(["RECALL"], [k]).topk_embeddings(du, di)
    assert torch.equal(i, ri) and torch.equal(v, rv)


def test_fused_sharded_equals_single(monkeypatch):
    """Shards scored separately (as ranks would) + mr_topk_merge == one table, bit for bit."""
    Q, N, E, k = 300, 20011, 64, 50
    users, items, _ = synth.make_catalog(Q, N, E, kind="gauss", seed=41)
    v1, i1 = fused(users, items, k)
    for world in (2, 4):
        pv, pi = [], []
        for r in range(world):
            a, b = shard_bounds(N, world, r)
            v, i = fused(users, items[a:b], k, id_base=a)
            pv.append(dev(v))
            pi.append(dev(i))
        mv, mi = topk_merge(torch.stack(pv), torch.stack(pi), k)
        assert np.array_equal(host(mi), i1) and np.array_equal(host(mv).view(np.uint32), v1.view(np.uint32))


def test_fused_argument_errors():
    lib = _lib.load()
    t = torch.zeros(64, device="cuda")
    assert lib.mr_score_topk(_lib.dptr(t), _lib.dptr(t), 4, _lib.dptr(t), _lib.dptr(t), 4, 6, 2, 0, 0, _lib.dptr(t),
                             _lib.dptr(t), None, 0, None) == -1      # E % 4 != 0
    assert lib.mr_score_topk(_lib.dptr(t), _lib.dptr(t), 4, _lib.dptr(t), _lib.dptr(t), 4, 8, 200, 0, 0, _lib.dptr(t),
                             _lib.dptr(t), None, 0, None) == -1      # K too large
    assert lib.mr_score_topk(_lib.dptr(t), _lib.dptr(t), 4, _lib.dptr(t), _lib.dptr(t), 4, 8, 2, 0, 0, _lib.dptr(t),
                             _lib.dptr(t), None, 0, None) == -3      # no workspace
    with pytest.raises(ValueError):
        Evaluator(["RECALL"], [200]).topk_embeddings(torch.zeros(4, 8), torch.zeros(300, 8))
    with pytest.raises(RuntimeError):
        Evaluator(["RECALL"], [20]).topk_embeddings(torch.zeros(4, 8), torch.zeros(10, 8))


# ------------------------------------------------------------------------------------------ bf16-compat mode (8f row 4)
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("case", [c for c in gc.EVAL_CASES if c["kind"] == "grid"], ids=lambda c: c["name"])
def test_fused_bf16_compat_grid(case, cg, monkeypatch):
    """bf16 operands, fp32 accumulation, bf16-rounded scores: bit-exact vs the reference's bf16 matmul + stable sort."""
    from mergerec_b200.evaluator import MR_SCORE_BF16
    from mergerec_b200.evaluator.sharded import to_bf16
    monkeypatch.setenv("MR_SCORE_CTA_GROUP", str(cg))
    g = golden("evaluator_bf16")
    users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
    assert np.array_equal(host(to_bf16(dev(users)).float()), orc.to_bf16(users))
    ev = Evaluator(case["metrics"], case["ks"])
    vals, ids = ev.topk_embeddings(dev(users), dev(items), mode=MR_SCORE_BF16)
    assert np.array_equal(host(ids), g[f"{case['name']}/canon_topk"])
    assert_bit_equal(host(vals), g[f"{case['name']}/canon_vals"], "bf16 scores")
    res = ev.evaluate_embeddings(dev(users), dev(items), dev(labels), metric_prefix=case["prefix"], mode=MR_SCORE_BF16)
    assert list(res.keys()) == list(g[f"{case['name']}/canon_keys"])
    assert np.array_equal(np.asarray(list(res.values()), np.float64), g[f"{case['name']}/canon_values"])


def test_fused_bf16_compat_gauss_and_shards():
    from mergerec_b200.evaluator import MR_SCORE_BF16
    Q, N, E, k = 300, 20011, 64, 50
    users, items, _ = synth.make_catalog(Q, N, E, kind="gauss", seed=43)
    ev = Evaluator(["RECALL"], [k])
    v, i = ev.topk_embeddings(dev(users), dev(items), mode=MR_SCORE_BF16)
    v, i = host(v), host(i)
    ref = orc.to_bf16(users).astype(np.float64) @ orc.to_bf16(items).astype(np.float64).T
    got = np.take_along_axis(ref, i.astype(np.int64), 1)
    assert np.abs(v - got).max() <= 2.0 ** -8 * np.abs(got).max(), "each returned score is its pair's bf16-rounded dot product"
    assert (np.diff(v, axis=1) <= 0).all(), "descending scores"
    same = np.diff(v, axis=1) == 0
    assert (np.diff(i, axis=1)[same] > 0).all(), "equal scores in ascending id order"
    kth = np.sort(orc.to_bf16(ref.astype(np.float32)), axis=1)[:, ::-1][:, k - 1]
    assert (np.abs(v[:, -1] - kth) <= 2.0 ** -7 * np.abs(kth)).all(), "k-th score within one bf16 step of the oracle's"
    # shards + merge == single table, bit for bit
    pv, pi = [], []
    for r in range(3):
        a, b = shard_bounds(N, 3, r)
        tv, ti = ev.topk_embeddings(dev(users), ShardedItemTable(dev(items[a:b]), id_base=a, n_total=N, bf16=True), mode=MR_SCORE_BF16)
        pv.append(tv)
        pi.append(ti)
    mv, mi = topk_merge(torch.stack(pv), torch.stack(pi), k)
    assert np.array_equal(host(mi), i) and np.array_equal(host(mv).view(np.uint32), v.view(np.uint32))
    with pytest.raises(ValueError):
        ev.topk_embeddings(dev(users), ShardedItemTable(dev(items)), mode=MR_SCORE_BF16)   # fp32-split table, bf16 mode


# ------------------------------------------------------------------------------------------ round 2 additions
def test_topk_merge_packed_equals_separate_lists():
    """The exchange buffer form (L, 2, Q, K) merges to the same lists as separate (L, Q, K) vals / ids."""
    from mergerec_b200.evaluator.sharded import topk_merge_packed
    rng = np.random.Generator(np.random.PCG64(11))
    for L, Q, K, k_out in ((2, 37, 10, 10), (8, 129, 100, 100), (3, 5, 7, 4)):
        vals = rng.standard_normal((L, Q, K)).astype(np.float32)
        vals[0, 0, :3] = vals[1, 0, :3]                           # equal scores across lists: lower id first
        ids = rng.permutation(L * Q * K).reshape(L, Q, K).astype(np.int32)
        ids[L - 1, Q - 1, K - 1] = -1                             # an empty slot
        packed = np.stack([vals.view(np.int32), ids], axis=1)     # (L, 2, Q, K)
        pv, pi = topk_merge_packed(dev(packed), k_out)
        mv, mi = topk_merge(dev(vals), dev(ids), k_out)
        assert np.array_equal(host(pi), host(mi))
        assert_bit_equal(host(pv), host(mv), "packed merge values")


@pytest.mark.parametrize("case", gc.EVAL_CFG4_CASES, ids=lambda c: c["name"])
def test_fused_cfg4_shape_vs_golden(case):
    """BASELINE config 4's evaluator shape -- N = 200,000 items, E = 1024, top-50 -- through the fused tensor-core path:
    ids, scores and metric floats bit-identical to the golden run of the unmodified reference (grid inputs)."""
    g = golden("evaluator_cfg4")
    users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
    ev = Evaluator(case["metrics"], case["ks"])
    vals, ids = ev.topk_embeddings(dev(users), dev(items))
    assert np.array_equal(host(ids), g[f"{case['name']}/canon_topk"])
    assert_bit_equal(host(vals), g[f"{case['name']}/canon_vals"], "fused scores at E=1024")
    res = ev.evaluate_embeddings(dev(users), dev(items), dev(labels), metric_prefix=case["prefix"])
    assert list(res.keys()) == list(g[f"{case['name']}/canon_keys"])
    assert np.array_equal(np.asarray(list(res.values()), np.float64), g[f"{case['name']}/canon_values"])


def test_fused_cfg4_shape_vs_fp32_kernel_more_queries():
    """E = 1024, K = 50, N = 200,000 with enough queries for several query blocks (Q = 1,100: 5 CTA pairs, a ragged
    last block): the fused path equals `mr_scores_fp32` + `mr_topk_rows` bit for bit on grid inputs."""
    from mergerec_b200 import _lib
    users, items, _ = synth.make_catalog(1100, 200_000, 1024, kind="grid", seed=55)
    tu, ti = dev(users), dev(items)
    ev = Evaluator(["RECALL"], [50])
    vals, ids = ev.topk_embeddings(tu, ti)
    sc = torch.empty((1100, 200_000), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().mr_scores_fp32(_lib.dptr(tu), 1100, _lib.dptr(ti), 200_000, 1024, _lib.dptr(sc), 200_000,
                                          _lib.stream_handle()), "mr_scores_fp32")
    rv, ri = topk_rows(sc, 50)
    assert torch.equal(ids, ri)
    assert_bit_equal(host(vals), host(rv), "fused vs fp32 kernel scores")


def test_prepared_queries_and_table_scratch_reuse():
    """`prepare_queries` + a reused ShardedItemTable give the same lists as raw tensors, call after call."""
    from mergerec_b200.evaluator import ShardedItemTable
    users, items, labels = synth.make_catalog(300, 5000, 64, kind="grid", seed=19)
    ev = Evaluator(["NDCG", "RECALL"], [10, 50])
    tu, ti, tl = dev(users), dev(items), dev(labels)
    v0, i0 = ev.topk_embeddings(tu, ti)
    table = ShardedItemTable(ti)
    q = ev.prepare_queries(tu)
    for _ in range(3):
        v, i = ev.topk_embeddings(q, table)
        assert torch.equal(i, i0) and torch.equal(v, v0)
    ws = table._ws
    assert ws is not None and table.workspace(16) is ws
    assert ev.evaluate_embeddings(q, table, tl) == ev.evaluate_embeddings(tu, ti, tl)
    with pytest.raises(ValueError):
        ev.topk_embeddings(ev.prepare_queries(tu, mode=MR_SCORE_BF16), table)


def test_topk_merge_sorted_fast_path_unsorted_fallback_and_duplicates():
    """`mr_topk_merge` merges sorted lists by rank counting and falls back to a sort when a list is not sorted; both
    must give the (score desc, id asc) order, with short rows padded by empty slots and duplicate keys kept."""
    rng = np.random.Generator(np.random.PCG64(23))
    L, Q, K = 5, 61, 37
    vals = rng.standard_normal((L, Q, K)).astype(np.float32)
    vals[:, :, ::5] = np.round(vals[:, :, ::5], 1)                          # plenty of equal scores across lists
    ids = rng.permutation(L * Q * K).reshape(L, Q, K).astype(np.int32)
    ids[1, 3, :] = -1                                                        # an entirely empty list for one row
    ids[:, 7, 2:] = -1                                                       # a row with only 2 candidates per list: 10 < k_out
    order = np.lexsort((ids, -vals), axis=-1)                                # sort every list: score desc, id asc, empties last
    empty = np.take_along_axis(ids, order, -1) < 0
    order = np.take_along_axis(order, np.argsort(empty, axis=-1, kind="stable"), -1)
    svals, sids = np.take_along_axis(vals, order, -1), np.take_along_axis(ids, order, -1)
    # expectation in numpy (the oracle's merge has no notion of empty slots): per row, all non-empty candidates ordered
    # by (score desc, id asc), padded with (-inf, -1)
    want_v = np.full((Q, L * K), -np.inf, np.float32)
    want_i = np.full((Q, L * K), -1, np.int32)
    for q in range(Q):
        v, i = vals[:, q, :].reshape(-1), ids[:, q, :].reshape(-1)
        keep = i >= 0
        v, i = v[keep], i[keep]
        o = np.lexsort((i, -v))
        want_v[q, :o.size], want_i[q, :o.size] = v[o], i[o]
    for k_out in (K, 12, 50):
        mv, mi = topk_merge(dev(svals), dev(sids), k_out)                    # sorted lists: rank-counting path
        uv, ui = topk_merge(dev(vals), dev(ids), k_out)                      # unsorted lists: sort fallback
        kk = min(k_out, want_i.shape[1])
        for got_v, got_i in ((mv, mi), (uv, ui)):
            assert np.array_equal(host(got_i)[:, :kk], want_i[:, :kk])
            valid = want_i[:, :kk] >= 0
            assert np.array_equal(host(got_v)[:, :kk][valid].view(np.uint32), want_v[:, :kk][valid].view(np.uint32))
            assert (host(got_i)[7, 10:] == -1).all() and np.isneginf(host(got_v)[7, 10:]).all()
    # the same (score, id) in two lists: both copies survive, next to each other
    dv = np.array([[[3.0, 2.0, 1.0]], [[3.0, 1.5, 0.5]]], np.float32)
    di = np.array([[[5, 6, 7]], [[5, 8, 9]]], np.int32)
    v, i = topk_merge(dev(dv), dev(di), 6)
    assert host(i)[0].tolist() == [5, 5, 6, 8, 7, 9] and host(v)[0].tolist() == [3.0, 3.0, 2.0, 1.5, 1.0, 0.5]


@pytest.mark.parametrize("kind,Q,N,E,k,first", [("grid", 300, 20011, 64, 50, 1000), ("gauss", 130, 5000, 768, 100, 64),
                                                 ("grid", 70, 90, 32, 50, 16), ("grid", 40, 3000, 128, 10, 4096)])
def test_streamed_host_table_equals_one_table(kind, Q, N, E, k, first):
    """`topk_embeddings_streamed` (host-resident item table copied in growing chunks behind the scoring) gives exactly the
    one-table result: ids, score bits, metric floats -- including chunks shorter than k and a single-chunk table."""
    users, items, labels = synth.make_catalog(Q, N, E, kind=kind, seed=29)
    ev = Evaluator(["NDCG", "RECALL"], [1, 10, k])
    tu, tl = dev(users), dev(labels)
    want_v, want_i = ev.topk_embeddings(tu, dev(items), k)
    for pinned in (False, True):
        h = torch.from_numpy(items)
        h = h.pin_memory() if pinned else h
        v, i = ev.topk_embeddings_streamed(tu, h, k, first_rows=first, max_chunks=7)
        assert torch.equal(i, want_i) and torch.equal(v.view(torch.int32), want_v.view(torch.int32))
    assert ev.evaluate_embeddings_streamed(tu, torch.from_numpy(items), tl, first_rows=first) == ev.evaluate_embeddings(tu, dev(items), tl)
    # a shard of a larger catalog: global ids = id_base + local row
    v, i = ev.topk_embeddings_streamed(tu, torch.from_numpy(items[N // 2:]), min(k, N - N // 2), id_base=N // 2, n_total=N, first_rows=first)
    wv, wi = ev.topk_embeddings(tu, ShardedItemTable(dev(items[N // 2:]), id_base=N // 2, n_total=N), min(k, N - N // 2))
    assert torch.equal(i, wi) and torch.equal(v.view(torch.int32), wv.view(torch.int32))
