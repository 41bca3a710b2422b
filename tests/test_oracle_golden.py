"""CPU: the oracle (oracle/merge_oracle.c + oracle/oracle.py) against the golden vectors produced by
the unmodified reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import golden_cases as gc
from helpers import assert_bit_equal, flatten_np, golden, shape_dict_of, state_dict_case
from mergerec_b200 import synth
from oracle import oracle as orc


@pytest.mark.parametrize("case", gc.MERGE_FLAT_CASES, ids=lambda c: c["name"])
def test_merge_flat(case):
    g = golden("merge_flat")
    base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"])
    assert_bit_equal(orc.merge_task_vector(base, models, case["weights"]), g[f"{case['name']}/task_vector"], "A1")
    assert_bit_equal(orc.merge_linear(models, case["weights"]), g[f"{case['name']}/linear"], "A10")
    assert_bit_equal(orc.task_vectors(base, models), g[f"{case['name']}/task_vectors"], "A2")


@pytest.mark.parametrize("case", gc.MODEL_MERGER_CASES, ids=lambda c: c["name"])
def test_model_merger_layout(case):
    g = golden("model_merger")
    _, base, models = state_dict_case(case)
    keys = sorted(base.keys())  # ModelMerger(align_key_order=True): merger/utils/model_operations.py:111-115
    assert list(g[f"{case['name']}/keys"]) == keys
    fb = flatten_np(base, keys)
    fm = [flatten_np(m, keys) for m in models]
    assert_bit_equal(fb, g[f"{case['name']}/base_flat"], "A0 flatten")
    seen = set()
    for mt, w, _ in case["merges"]:
        if mt in seen:
            continue  # golden stores the last merge of each type
        ws = [m for m in case["merges"] if m[0] == mt][-1][1]
        wl = [ws] * len(fm) if isinstance(ws, float) else ws
        out = orc.merge_task_vector(fb, fm, wl) if mt == "task_vector" else orc.merge_linear(fm, wl)
        assert_bit_equal(out, g[f"{case['name']}/{mt}"], mt)
        seen.add(mt)


@pytest.mark.parametrize("case", gc.LAMBDA_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("learn", ["task", "layer"])
@pytest.mark.parametrize("softmax", [0, 1])
def test_lambda_merge_and_grad(case, learn, softmax):
    g = golden("lambda_merge")
    _, base, models = state_dict_case(case)
    fb = flatten_np(base)
    fm = [flatten_np(m) for m in models]
    T = orc.task_vectors(fb, fm)
    sb, se, sg, keys = orc.segment_table(shape_dict_of(base), layer_wise=(learn == "layer"))
    tag = f"{case['name']}/{learn}/softmax{softmax}"
    assert list(g[f"{tag}/keys"]) == keys
    w = g[f"{tag}/w"]
    assert_bit_equal(orc.lambda_merge(fb, T, w, sb, se, sg), g[f"{tag}/merged"], "A3/A4 merged")
    # A5: reference autograd (fp32) vs fp64 oracle, chained through w = gw * pw' + gb
    lg = orc.lambda_grad(g[f"{tag}/grad_out"], T, len(keys), sb, se, sg)
    gw, pw = g[f"{tag}/global_weights"].astype(np.float64), g[f"{tag}/per_weights"].astype(np.float64)
    np.testing.assert_allclose(g[f"{tag}/grad_global_biases"][:, 0], lg.sum(axis=1), rtol=2e-4, atol=1e-6)
    if not softmax:
        np.testing.assert_allclose(g[f"{tag}/grad_per_weights"], gw * lg, rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(g[f"{tag}/grad_global_weights"][:, 0], (pw * lg).sum(axis=1), rtol=2e-4, atol=1e-6)


@pytest.mark.parametrize("case", gc.TIES_CASES, ids=lambda c: c["name"])
def test_ties(case):
    g = golden("ties")
    base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"], tie_free=case["tie_free"],
                                   quantize=case.get("quantize", 0.0))
    ref_T = g[f"{case['name']}/ties_vectors"]
    ref_M = g[f"{case['name']}/merge_ties"]
    That, trim, elect, cut = orc.ties_vectors(base, models, case["density"], return_masks=True)
    k_cnt = orc.ties_topk_count(case["density"], case["d"])
    assert (trim.sum(axis=1) == k_cnt).all(), "every model keeps exactly int(density*d) entries (ties.py:15,23)"
    merged = orc.merge_ties(base, models, case["weights"], case["density"])
    if case["tie_free"]:
        # no magnitude ties -> torch.topk's tie order cannot matter -> bit-exact vs the raw reference
        assert_bit_equal(That, ref_T, "A6-A8 get_ties_vectors")
        assert_bit_equal(merged, ref_M, "A9 merge_ties")
        return
    # inputs with ties: the canonical (lowest-index) rule may differ from torch.topk only at elements
    # whose magnitude equals the per-model threshold
    thr = (cut >> np.uint64(32)).astype(np.uint32)
    diff_cols = np.zeros(case["d"], bool)
    for k, m in enumerate(models):
        mag = np.abs(m - base).view(np.uint32)
        at_thr = mag == thr[k]
        differs = That[k].view(np.uint32) != ref_T[k].view(np.uint32)
        diff_cols |= differs
    # a column may differ only if some model has a threshold-magnitude entry in it
    col_has_thr = np.zeros(case["d"], bool)
    for k, m in enumerate(models):
        col_has_thr |= np.abs(m - base).view(np.uint32) == thr[k]
    assert not (diff_cols & ~col_has_thr).any(), "difference vs raw torch.topk outside threshold ties"


def test_ties_edge_densities():
    base, models = synth.make_flat(1000, 3, seed=5)
    for density, expect in [(0.0, 0), (1.0, 1000), (0.0004, 0), (0.9999, 999)]:
        _, trim, _, _ = orc.ties_vectors(base, models, density, return_masks=True)
        assert (trim.sum(axis=1) == expect).all()
    # model identical to base: every |u| == 0 -> all ties, the lowest indices survive
    _, trim, _, _ = orc.ties_vectors(base, [base.copy(), models[0]], 0.25, return_masks=True)
    assert trim[0, :250].all() and not trim[0, 250:].any()


@pytest.mark.parametrize("case", gc.EVAL_CASES, ids=lambda c: c["name"])
def test_evaluator(case):
    g = golden("evaluator")
    users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
    kmax = max(case["ks"])
    canon = g[f"{case['name']}/canon_topk"]
    if case["kind"] == "grid":
        scores = orc.scores_f32(users, items)
        # exact-grid inputs: any summation order gives the same fp32 bits as the reference's sgemm
        assert np.array_equal(scores, orc.scores_f64(users, items).astype(np.float32))
        vals, ids = orc.topk_rows(scores, kmax)
        assert np.array_equal(ids, canon), "canonical (score desc, id asc) ids"
        res = orc.evaluate_ids(ids, labels, case["metrics"], case["ks"], case["prefix"])
        assert list(res.keys()) == list(g[f"{case['name']}/canon_keys"])
        assert np.array_equal(np.asarray(list(res.values()), np.float64), g[f"{case['name']}/canon_values"])
        # raw torch.topk picks the same score multiset per row
        raw = g[f"{case['name']}/raw_topk"]
        assert np.array_equal(np.take_along_axis(scores, raw, 1), vals)
    else:
        # gaussian inputs: BLAS bits differ between hosts, so rank with fp64 scores and require agreement with
        # the stored reference ids wherever the fp64 gap is not a near-tie
        s64 = orc.scores_f64(users, items)
        _, ids = orc.topk_rows(s64.astype(np.float32), kmax)
        mism = ids != canon
        if mism.any():
            rows = np.unique(np.argwhere(mism)[:, 0])
            for r in rows:
                top = np.sort(s64[r])[::-1][: kmax + 1]
                assert np.min(np.abs(np.diff(top))) < 1e-6, f"row {r}: ids differ without a near-tie"
        res = orc.evaluate_ids(canon, labels, case["metrics"], case["ks"], case["prefix"])
        assert list(res.keys()) == list(g[f"{case['name']}/raw_keys"])
        # tie-free: canonical ids == raw torch.topk ids, so the full Evaluator result must match bit-for-bit
        assert np.array_equal(g[f"{case['name']}/raw_topk"], canon)
        assert np.array_equal(np.asarray(list(res.values()), np.float64), g[f"{case['name']}/raw_values"])


@pytest.mark.parametrize("case", gc.EVAL_CFG4_CASES, ids=lambda c: c["name"])
def test_evaluator_cfg4_shape(case):
    """BASELINE config 4's evaluator shape (N = 200,000, E = 1024, top-50): oracle ids / values / metric floats equal the
    golden run of the unmodified reference."""
    g = golden("evaluator_cfg4")
    users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
    kmax = max(case["ks"])
    scores = orc.scores_f32(users, items)
    vals, ids = orc.topk_rows(scores, kmax)
    assert np.array_equal(ids, g[f"{case['name']}/canon_topk"])
    assert np.array_equal(vals, g[f"{case['name']}/canon_vals"])
    res = orc.evaluate_ids(ids, labels, case["metrics"], case["ks"], case["prefix"])
    assert list(res.keys()) == list(g[f"{case['name']}/canon_keys"])
    assert np.array_equal(np.asarray(list(res.values()), np.float64), g[f"{case['name']}/canon_values"])


def test_ndcg_gain_table():
    g = golden("evaluator")
    table = np.asarray([1 / orc._log2_f32(r + 2) for r in range(1024)], np.float64)
    assert np.array_equal(table, g["ndcg_gain_table"]), "fp32 log2 table (metrics.py:84)"


def test_topk_merge_equals_unsharded():
    users, items, _ = synth.make_catalog(32, 1000, 16, kind="grid", seed=9)
    scores = orc.scores_f32(users, items)
    v_all, i_all = orc.topk_rows(scores, 20)
    parts = [(0, 300), (300, 650), (650, 1000)]
    pv, pi = zip(*[orc.topk_rows(np.ascontiguousarray(scores[:, a:b]), 20, id_base=a) for a, b in parts])
    mv, mi = orc.topk_merge(np.stack(pv), np.stack(pi))
    assert np.array_equal(mi, i_all) and np.array_equal(mv, v_all)
    assert np.array_equal(orc.label_rank(mi, mi[:, 3].astype(np.int64)), np.full(32, 3, np.int32))


@pytest.mark.parametrize("case", gc.LNS_CASES, ids=lambda c: c["name"])
def test_localize_and_stitch(case):
    """Oracle vs the reference's get_localize_and_stitch_vectors / merge_localize_and_stitch (lns.npz)."""
    g = golden("lns")
    base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"], tie_free=case["tie_free"])
    vec = orc.lns_vectors(base, models, case["density"])
    merged = orc.merge_localize_and_stitch(base, models, case["weights"], case["density"])
    ref_v, ref_m = g[f"{case['name']}/vectors"], g[f"{case['name']}/merged"]
    if case["tie_free"]:
        assert_bit_equal(vec, ref_v, "L&S vectors vs raw reference")
        assert_bit_equal(merged, ref_m, "L&S merge vs raw reference")
    else:
        # torch.topk's choice among equal magnitudes at the threshold is unspecified: differences from the raw
        # reference must be confined to columns holding a threshold-magnitude entry
        cut = orc.ties_select(base, models, case["density"])
        thr = (cut >> np.uint64(32)).astype(np.uint32)
        mags = np.stack([(m - base).astype(np.float32) for m in models]).view(np.uint32) & np.uint32(0x7FFFFFFF)
        at_thr = (mags == thr[:, None]).any(axis=0)
        diff = (vec.view(np.uint32) != ref_v.view(np.uint32)).any(axis=0)
        assert not (diff & ~at_thr).any()
        assert not ((merged.view(np.uint32) != ref_m.view(np.uint32)) & ~at_thr).any()


@pytest.mark.parametrize("case", [c for c in gc.EVAL_CASES if c["kind"] == "grid"], ids=lambda c: c["name"])
def test_evaluator_bf16_compat(case):
    """bf16-compat scoring (the reference's default bf16-mixed runs): oracle vs torch's own bf16 matmul + stable sort."""
    g = golden("evaluator_bf16")
    users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
    kmax = max(case["ks"])
    vals, ids = orc.topk_rows(orc.scores_bf16(users, items), kmax)
    assert np.array_equal(ids, g[f"{case['name']}/canon_topk"])
    assert_bit_equal(vals, g[f"{case['name']}/canon_vals"], "bf16 scores")
    res = orc.evaluate_ids(ids, labels, case["metrics"], case["ks"], case["prefix"])
    assert list(res.keys()) == list(g[f"{case['name']}/canon_keys"])
    assert np.array_equal(np.asarray(list(res.values()), np.float64), g[f"{case['name']}/canon_values"])
