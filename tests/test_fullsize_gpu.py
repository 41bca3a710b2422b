"""Full-size (BASELINE.json shapes) checks on the GPU, where the CPU oracle would take minutes: size-independent
properties plus comparisons with an independent implementation of the same IEEE arithmetic (torch's own CUDA
elementwise kernels, which never contract mul+add across ops) on identical inputs.

BLaIR-base flat length d = 124,645,632, K = 8 domain models (configs 2 and 3)."""
import numpy as np
import pytest
import torch

from mergerec_b200 import _lib, synth
from mergerec_b200.merger.algorithms import get_task_vectors, get_ties_vectors, merge_task_vector, merge_ties
from mergerec_b200.merger.algorithms._common import merge_axpy
from mergerec_b200.merger.algorithms.ties import merge_ties_lambda, ties_topk_count
from mergerec_b200.merger.layout import FlatLayout
from mergerec_b200.merger.weight_learning.module._base import _lambda_grad

pytestmark = pytest.mark.gpu

K = 8


@pytest.fixture(scope="module")
def blair():
    shapes = synth.roberta_shapes()
    d = synth.total_numel(shapes)
    assert d == 124_645_632
    g = torch.Generator(device="cuda").manual_seed(2024)
    base = torch.randn(d, generator=g, device="cuda") * 0.02
    models = [base + 1e-3 * torch.randn(d, generator=g, device="cuda") for _ in range(K)]
    yield shapes, d, base, models
    del base, models
    torch.cuda.empty_cache()


def test_task_vector_merge_equals_torch_elementwise(blair):
    _, d, base, models = blair
    w = [0.3, -0.25, 0.5, 0.125, 0.7, 0.05, 1.0, 0.33]
    got = merge_task_vector(base, models, w)
    acc = base.clone()
    for m, wk in zip(models, w):                 # task_vector.py:30-32, op by op (unfused on the GPU as on the CPU)
        acc = acc + torch.tensor(wk, dtype=torch.float32, device="cuda") * (m - base)
    assert torch.equal(got, acc)
    assert torch.equal(merge_task_vector(base, models, [0.0] * K), base + 0.0 * (models[0] - base))   # w = 0 -> base


def test_layerwise_lambda_merge_and_grad(blair):
    shapes, d, base, models = blair
    layout = FlatLayout.from_shape_dict(shapes)
    seg_end, seg_group, keys = layout.device_blocks(True, base.device)
    G = len(keys)
    assert G == 13
    T = get_task_vectors(base, models)
    rng = np.random.Generator(np.random.PCG64(5))
    w = torch.from_numpy(rng.uniform(0.1, 0.5, size=(G, K)).astype(np.float32)).cuda()
    merged = merge_axpy(base, list(T.unbind(0)), w, _lib.MR_ORDER_SUM_FIRST, False, seg_end, seg_group)
    # reference order per block (layer_wise.py:76-82): s = 0; s += w[g,k] * T[k]; merged = base + s  (all blocks of
    # BLaIR are multiples of 32 long: no interleaved tail)
    ends = seg_end.cpu().tolist()
    groups = seg_group.cpu().tolist()
    starts = [0] + ends[:-1]
    wrow = w[torch.tensor(groups, device="cuda")]                       # (P, K)
    lens = torch.tensor([e - s for s, e in zip(starts, ends)], device="cuda")
    wfull = torch.repeat_interleave(wrow, lens, dim=0)                  # (d, K) -- 4 GB, fine on a B200
    s = torch.zeros(d, device="cuda")
    for k in range(K):
        s = s + wfull[:, k] * T[k]
    assert torch.equal(merged, base + s)
    del wfull, s
    # lambda-gradient against an fp64 reduction
    g = torch.randn(d, device="cuda")
    seg_off, seg_len = layout.device_segments(base.device)
    grads = [g[o:o + n] for o, n in zip(seg_off.tolist(), seg_len.tolist())]
    got = _lambda_grad(grads, layout, T, seg_group, G).double()
    want = torch.zeros((G, K), dtype=torch.float64, device="cuda")
    for p, (o, n) in enumerate(zip(seg_off.tolist(), seg_len.tolist())):
        want[groups[p]] += T[:, o:o + n].double() @ g[o:o + n].double()
    rel = (got - want).abs().max() / want.abs().max()
    assert rel < 1e-5, f"lambda-gradient relative error {rel:.2e}"
    # deterministic: two runs give identical bits
    assert torch.equal(_lambda_grad(grads, layout, T, seg_group, G), _lambda_grad(grads, layout, T, seg_group, G))


def test_ties_full_size_properties(blair):
    shapes, d, base, models = blair
    density = 0.2
    k_cnt = ties_topk_count(density, d)
    assert k_cnt == 24_929_126            # int(0.2 * 124,645,632), SURVEY.md appendix A
    That, trim, elect, cut = get_ties_vectors(base, models, density, return_masks=True)
    cut_np = cut.cpu().numpy().view(np.uint64)
    for k in range(K):
        u = (models[k] - base).abs()
        kept = trim[k].bool()
        assert int(kept.sum()) == k_cnt, "the trim keeps exactly int(density * d) entries per model"
        thr_bits = int(cut_np[k] >> np.uint64(32))
        mags = u.view(torch.int32)
        assert int(mags[kept].min()) == thr_bits, "smallest kept magnitude is the cut magnitude"
        assert int(mags[~kept].max()) <= thr_bits, "nothing dropped is larger than the cut"
        # the same order statistic from an independent implementation
        assert float(torch.kthvalue(u, d - k_cnt + 1).values) == float(u[kept].min())
        del u, kept, mags
    # election / disjoint mean invariants (ties.py:55-72)
    nz = That != 0
    assert torch.equal(nz, elect.bool()), "T-hat is non-zero exactly on the elected survivors"
    assert not bool((elect.bool() & ~trim.bool()).any()), "elected entries survived the trim"
    signs = torch.sign(That)
    assert bool(((signs.max(dim=0).values - signs.min(dim=0).values) <= 1).all()), "one sign per column"
    cnt = nz.sum(dim=0)
    col = int(torch.argmax(cnt))
    ks = torch.nonzero(nz[:, col]).flatten()
    u_col = torch.stack([models[int(k)][col] - base[col] for k in ks])
    assert torch.equal(That[ks, col], u_col / float(len(ks))), "disjoint mean is an IEEE division by the count"
    del nz, signs, cnt
    # fused election + mean + per-layer lambda merge == materialised T-hat followed by the lambda merge
    layout = FlatLayout.from_shape_dict(shapes)
    seg_end, seg_group, keys = layout.device_blocks(True, base.device)
    w = torch.rand((len(keys), K), device="cuda") * 0.4 + 0.1
    two_step = merge_axpy(base, list(That.unbind(0)), w, _lib.MR_ORDER_SUM_FIRST, False, seg_end, seg_group)
    fused = merge_ties_lambda(base, models, density, w, seg_end, seg_group)
    assert torch.equal(fused, two_step)
    # merge_ties (trim after weighting, no election): equals base + sum of trimmed weighted updates
    wl = [0.5] * K
    mt = merge_ties(base, models, wl, density)
    s = torch.zeros(d, device="cuda")
    for k in range(K):      # equal positive weights scale magnitudes uniformly: same survivors as the unweighted trim
        s = s + torch.where(trim[k].bool(), torch.tensor(0.5, device="cuda") * (models[k] - base), torch.zeros((), device="cuda"))
    assert torch.equal(mt, base + s)


def test_recformer_large_config4_properties():
    """BASELINE config 4 merger shapes: Recformer-large, d = 433,610,754 (d mod 32 = 2; every tensor after the
    4,098-element position_ids starts at a flat offset that is 8- but not 16-byte aligned), K = 8, layer-wise G = 25."""
    shapes = synth.recformer_shapes()
    d = synth.total_numel(shapes)
    assert d == 433_610_754 and d % 32 == 2
    layout = FlatLayout.from_shape_dict(shapes)
    g = torch.Generator(device="cuda").manual_seed(7)
    base = torch.randn(d, generator=g, device="cuda") * 0.02
    models = [base + 1e-3 * torch.randn(d, generator=g, device="cuda") for _ in range(K)]
    seg_end, seg_group, keys = layout.device_blocks(True, base.device)
    assert len(keys) == 25
    # task arithmetic against torch's own elementwise kernels
    w = [0.3, 0.2, 0.1, 0.4, 0.25, 0.15, 0.35, 0.05]
    got = merge_task_vector(base, models, w)
    acc = base.clone()
    for m, wk in zip(models, w):
        acc = acc + torch.tensor(wk, dtype=torch.float32, device="cuda") * (m - base)
    assert torch.equal(got, acc)
    del acc, got
    # TIES: exact count per model, and the fused election + mean + layer-wise merge equals the two-step path, including
    # the 2-column interleaved tail of the flat vector (K = 8 >= 5) and blocks that straddle 128-bit quads
    density = 0.2
    k_cnt = ties_topk_count(density, d)
    assert k_cnt == 86_722_150
    That, trim, elect, cut = get_ties_vectors(base, models, density, return_masks=True)
    assert bool((trim.sum(dim=1) == k_cnt).all())
    assert torch.equal(That != 0, elect)
    del trim, elect
    # the stream-ordered fast select must hold at this size too (3.2 M keys fall inside the sample bracket: two
    # refinement levels over the collected keys), and agree with the public path's cut
    from mergerec_b200.merger.algorithms.ties import select_kth_largest
    fast_cut, status = select_kth_largest(base, models, k_cnt, None, defer_status=True)
    assert status.cpu().tolist() == [1] * K, "fast select fell back at d = 433,610,754"
    assert torch.equal(fast_cut, cut)
    lam = torch.rand((len(keys), K), device="cuda") * 0.4 + 0.1
    two_step = merge_axpy(base, list(That.unbind(0)), lam, _lib.MR_ORDER_SUM_FIRST, False, seg_end, seg_group)
    del That
    torch.cuda.empty_cache()
    fused = merge_ties_lambda(base, models, density, lam, seg_end, seg_group, cut=cut)
    assert torch.equal(fused, two_step)
    # the last two flat columns use torch.sum's 4-way interleaved order (K >= 5): recompute them by hand in that order
    T2 = torch.stack([m[d - 2:] - base[d - 2:] for m in models])          # (K, 2) raw task vectors of the tail
    lam_tail = lam[int(seg_group[-1])]
    tw = merge_axpy(base, [m - base for m in models], lam[:1].contiguous(), _lib.MR_ORDER_SUM_FIRST, False)   # task-wise
    prod = lam[0][:, None] * T2
    p0 = (prod[0] + prod[4]); p1 = (prod[1] + prod[5]); p2 = (prod[2] + prod[6]); p3 = (prod[3] + prod[7])
    assert torch.equal(tw[d - 2:], base[d - 2:] + (((p0 + p1) + p2) + p3)), "interleaved tail order (task-wise, whole-vector block)"
    del lam_tail
