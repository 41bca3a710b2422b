"""GPU: the distillation kernels (csrc/distill.cu) through the C ABI against the fp64 oracle and the golden vectors of
the reference's own loss classes (tests/golden/distill.npz).  Floating-point path: tolerance 2e-5 relative to the largest
magnitude of the compared array (losses, gradients), as stated in include/mergerec_b200.h."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from helpers import golden
from mergerec_b200 import _lib, synth
from mergerec_b200.module.distiller import (BatchDistillationSequence, DistillSequenceModule, TeacherScores,
                                            make_score_embeddings)
from mergerec_b200.module.distiller.sequence.module import distill_logits, fused_distill_losses, normalize_rows
from mergerec_b200.module.recommender import loss_fn as lf
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 2e-5


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def loss_object(name, kw):
    cls = {"CE": lf.DistillCELoss, "KD": lf.DistillKDLoss, "MSE": lf.DistillMSELoss, "ADAMERGING": lf.DistillAdaMergingLoss,
           "ADAMERGING_KD": lf.DistillAdaMergingKDLoss, "MERGED_PSEUDO_LABEL": lf.MergedPseudoLabelLoss,
           "MERGED_PSEUDO_LABEL_KD": lf.MergedPseudoLabelKDLoss, "SINGLE_PSEUDO_LABEL": lf.SinglePseudoLabelLoss,
           "SINGLE_PSEUDO_LABEL_KD": lf.SinglePseudoLabelKDLoss, "PAIRWISE": lf.DistillPairwiseLoss,
           "LISTNET": lf.DistillListNetLoss}[name]
    return cls(**kw)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("case", gc.DISTILL_CASES, ids=lambda c: c["name"])
def test_fused_step_matches_oracle_and_reference(case):
    g = golden("distill")
    c = synth.make_distill_case(case["B"], case["E"], case["rows"], case["n_seq"], case["seed"], planted=case["scale"])
    tables = [dev(t) for t in c["tables"]]
    teacher = TeacherScores([dev(t) for t in c["teacher_items"]], [dev(t) for t in c["teacher_seqs"]])
    for d, m in enumerate(teacher.scores):
        assert rel(m.cpu().numpy(), g[f"{case['name']}/score_embeddings{d}"]) < 1e-5
    trows64 = [orc.teacher_scores(c["teacher_items"][d], c["teacher_seqs"][d])[s]
               for d, s in zip(c["dataset_indexes"], c["sequence_ids"])]
    for lname, kw in gc.DISTILL_LOSSES:
        spec = loss_object(lname, kw).spec
        rep = dev(c["rep"]).requires_grad_(True)
        _, ptrs = teacher.rows(c["dataset_indexes"], c["sequence_ids"])
        losses = fused_distill_losses(rep, tables, c["dataset_indexes"], ptrs if spec.needs_teacher else None, spec)
        loss = losses.mean()
        loss.backward()
        o_losses, o_loss, o_grad = orc.distill_step(c["rep"], c["tables"], c["dataset_indexes"], trows64, lname, **kw)
        assert rel(losses.detach().cpu().numpy(), o_losses) < TOL, lname
        assert rel(rep.grad.cpu().numpy(), o_grad) < TOL, lname
        assert rel(losses.detach().cpu().numpy(), g[f"{case['name']}/{lname}/losses"]) < TOL, lname
        assert abs(float(loss) - float(g[f"{case['name']}/{lname}/loss"])) < TOL * max(1.0, abs(o_loss)), lname
        assert rel(rep.grad.cpu().numpy(), g[f"{case['name']}/{lname}/grad_rep"]) < TOL, lname


def test_logits_every_width_and_ragged_tables():
    rng = np.random.default_rng(5)
    for E in (4, 64, 132, 256, 388, 512, 768, 1024):
        rows = [1, 17, 33, 250]
        tables = [rng.standard_normal((n, E)).astype(np.float32) for n in rows]
        dom = [3, 0, 3, 1, 3, 3, 3, 2, 1, 3, 3]          # domain 3 has 7 samples -> two sample groups
        rep = rng.standard_normal((len(dom), E)).astype(np.float32)
        out = distill_logits(dev(rep), [dev(t) for t in tables], dom).cpu().numpy()
        for b, d in enumerate(dom):
            want = tables[d].astype(np.float64) @ rep[b].astype(np.float64)
            assert np.abs(out[b, :rows[d]] - want).max() < 2e-5 * max(1.0, np.abs(want).max()), (E, b)


@pytest.mark.parametrize("lname,kw", gc.DISTILL_LOSSES, ids=lambda x: x if isinstance(x, str) else "")
def test_loss_classes_drop_in_on_logit_rows(lname, kw):
    """`loss_fn(merged (R, N), single (R, N))` with autograd, R = 1 (how the reference calls it) and R = 5."""
    rng = np.random.default_rng(11)
    for R, N in ((1, 1000), (5, 333)):
        z = (3.0 * rng.standard_normal((R, N))).astype(np.float32)
        t = (z + rng.standard_normal((R, N))).astype(np.float32)
        zt = dev(z).requires_grad_(True)
        loss = loss_object(lname, kw)(zt, dev(t))
        loss.backward()
        want, grads = [], []
        for r in range(R):
            lv, gz = orc.distill_loss(z[r], t[r], lname, **kw)
            want.append(lv)
            grads.append(gz / R)
        assert abs(float(loss) - np.mean(want)) < TOL * max(1.0, abs(np.mean(want))), lname
        assert rel(zt.grad.cpu().numpy(), np.stack(grads)) < TOL, lname


def test_argmax_ties_take_the_first_maximum():
    z = dev(np.array([[0.1, 0.9, 0.3, 0.2]], np.float32)).requires_grad_(True)
    t = dev(np.array([[1.0, 2.0, 2.0, 0.0]], np.float32))
    lf.DistillCELoss()(z, t).backward()
    g = z.grad.cpu().numpy()[0]
    assert g[1] < 0 and g[2] > 0
    z.grad = None
    lf.DistillPairwiseLoss(10.0)(z, t).backward()
    assert z.grad.cpu().numpy()[0].tolist() == [0.0, -1.0, 1.0, 0.0]
    z1 = dev(np.array([[0.5]], np.float32)).requires_grad_(True)     # single item: pos == neg == 0 in the reference
    loss = lf.DistillPairwiseLoss(0.25)(z1, dev(np.array([[1.0]], np.float32)))
    assert float(loss) == 0.25


class _ToyEncoder(torch.nn.Module):
    def __init__(self, E):
        super().__init__()
        self.lin = torch.nn.Linear(E, E, bias=False)

    def forward(self, x):
        return self.lin(x)

    def trainable_parameters(self, **_):
        return list(self.parameters())


@pytest.mark.parametrize("on_the_fly", [False, True])
def test_distill_sequence_module_end_to_end(on_the_fly):
    """`DistillSequenceModule._forward_distill` + `configure_optimizers` against the same step in fp64 numpy."""
    case = gc.DISTILL_CASES[0]
    c = synth.make_distill_case(case["B"], case["E"], case["rows"], case["n_seq"], case["seed"], planted=case["scale"])
    torch.manual_seed(3)
    enc = _ToyEncoder(case["E"]).cuda()
    teacher = TeacherScores([dev(t) for t in c["teacher_items"]], [dev(t) for t in c["teacher_seqs"]], on_the_fly=on_the_fly)
    mod = DistillSequenceModule(enc, teacher, lf.DistillKDLoss(2.0), similarity="cosine", learning_rate=1e-2)
    mod.item_embeddings = [dev(t) for t in c["tables"]]
    batch = BatchDistillationSequence(sequence=dev(c["rep"]), dataset_indexes=c["dataset_indexes"], sequence_ids=c["sequence_ids"])
    opt = mod.configure_optimizers()
    loss = mod(batch)
    loss.backward()
    W = enc.lin.weight.detach().cpu().numpy().astype(np.float64)
    x = c["rep"].astype(np.float64)
    y = x @ W.T
    nrm = np.linalg.norm(y, axis=1, keepdims=True)
    rep = y / nrm
    trows = [orc.teacher_scores(c["teacher_items"][d], c["teacher_seqs"][d])[s] for d, s in zip(c["dataset_indexes"], c["sequence_ids"])]
    _, o_loss, o_grad = orc.distill_step(rep, c["tables"], c["dataset_indexes"], trows, "KD", temperature=2.0)
    gy = (o_grad - rep * (o_grad * rep).sum(1, keepdims=True)) / nrm     # through the cosine normalisation
    gW = gy.T @ x
    assert abs(float(loss) - o_loss) < 5e-5 * max(1.0, abs(o_loss))
    assert rel(enc.lin.weight.grad.cpu().numpy(), gW) < 1e-4
    opt.step()
    with pytest.raises(ValueError):
        mod(object())


def test_teacher_scores_and_normalize_rows():
    rng = np.random.default_rng(2)
    items = rng.standard_normal((1001, 96)).astype(np.float32)
    seqs = rng.standard_normal((37, 96)).astype(np.float32)
    got = make_score_embeddings(dev(items), dev(seqs)).cpu().numpy()
    assert rel(got, orc.teacher_scores(items, seqs)) < 1e-5
    n = normalize_rows(dev(items)).cpu().numpy()
    assert rel(n, items / np.linalg.norm(items.astype(np.float64), axis=1, keepdims=True)) < 1e-6


def test_errors():
    with pytest.raises(_lib.MergeRecLibraryError):
        fused_distill_losses(torch.zeros(2, 8), [torch.zeros(4, 8)], [0, 0], None, lf.DistillAdaMergingLoss().spec)
    rep = torch.zeros(2, 8, device="cuda")
    with pytest.raises(ValueError):
        fused_distill_losses(rep, [torch.zeros(4, 8, device="cuda")], [0, 1], None, lf.DistillAdaMergingLoss().spec)
    with pytest.raises(ValueError):
        fused_distill_losses(rep, [torch.zeros(4, 12, device="cuda")], [0, 0], None, lf.DistillAdaMergingLoss().spec)
    with pytest.raises(ValueError):
        lf.DistillKDLoss(2.0)(torch.zeros(1, 8, device="cuda"), torch.zeros(1, 9, device="cuda"))


def test_full_size_properties():
    """BASELINE config-3 shape of the step (16 samples, 8 domains, 25,000 items, E = 768): beyond the oracle's comfort,
    so size-independent properties -- softmax-type gradients sum to zero per row, the representation gradient is linear in
    the upstream gradient, two launches are bit-identical (deterministic reductions), and a sampled set of logits equals
    fp64 dot products."""
    rng = np.random.default_rng(9)
    E, D, N, B = 768, 8, 25000, 16
    tables = [torch.randn(N + 13 * d, E, device="cuda") * 0.05 for d in range(D)]
    dom = [b % D for b in range(B)]
    rep = torch.randn(B, E, device="cuda")
    teacher = [torch.randn(N + 13 * d, device="cuda") for d in dom]
    ptrs = [t.data_ptr() for t in teacher]
    logits = distill_logits(rep, tables, dom)
    for b in (0, 7, 15):
        idx = rng.integers(0, tables[dom[b]].shape[0], 50)
        want = tables[dom[b]][idx].double() @ rep[b].double()
        assert (logits[b, idx].double() - want).abs().max().item() < 1e-5
    spec = lf.DistillKDLoss(2.0).spec
    r1 = rep.clone().requires_grad_(True)
    l1 = fused_distill_losses(r1, tables, dom, ptrs, spec)
    l1.sum().backward()
    r2 = rep.clone().requires_grad_(True)
    l2 = fused_distill_losses(r2, tables, dom, ptrs, spec)
    (2.0 * l2.sum()).backward()
    assert torch.equal(l1, l2)
    assert torch.equal(2.0 * r1.grad, r2.grad)
    loss, gz = lf.launch_distill_loss(logits, ptrs, [t.shape[0] for t in teacher], spec, want_grad=True)
    for b in range(B):
        n = teacher[b].shape[0]
        assert abs(gz[b, :n].double().sum().item()) < 1e-4
    assert torch.isfinite(l1).all() and (l1 >= -1e-5).all()      # KL divergence is non-negative


def test_item_distill_module_shares_the_step():
    """`DistillModule` (reference item/module.py): same kernels, item-side batch fields (`items`, `item_ids`)."""
    from mergerec_b200.module.distiller import BatchDistillationItem, DistillModule
    case = gc.DISTILL_CASES[0]
    c = synth.make_distill_case(case["B"], case["E"], case["rows"], case["n_seq"], case["seed"], planted=case["scale"])
    torch.manual_seed(3)
    enc = _ToyEncoder(case["E"]).cuda()
    teacher = TeacherScores([dev(t) for t in c["teacher_items"]], [dev(t) for t in c["teacher_seqs"]])
    seq_mod = DistillSequenceModule(enc, teacher, lf.DistillKDLoss(2.0), similarity="cosine")
    item_mod = DistillModule(enc, teacher, lf.DistillKDLoss(2.0), similarity="cosine")
    seq_mod.item_embeddings = item_mod.item_embeddings = [dev(t) for t in c["tables"]]
    a = seq_mod(BatchDistillationSequence(sequence=dev(c["rep"]), dataset_indexes=c["dataset_indexes"], sequence_ids=c["sequence_ids"]))
    b = item_mod(BatchDistillationItem(items=dev(c["rep"]), dataset_indexes=c["dataset_indexes"], item_ids=c["sequence_ids"]))
    assert torch.equal(a, b)
    from types import SimpleNamespace
    enc_only = item_mod(SimpleNamespace(items=dev(c["rep"])))
    assert enc_only.shape == (case["B"], case["E"]) and torch.allclose(enc_only.norm(dim=-1), torch.ones(case["B"], device="cuda"), atol=1e-5)
    with pytest.raises(ValueError):
        item_mod(object())


def test_teacher_rows_are_validated_and_host_resident_matrices_match():
    """A teacher matrix narrower than the item table must raise (it would be read out of bounds by the loss kernel);
    teacher matrices kept in pinned host memory (large catalogs) give the same losses as device-resident ones."""
    from mergerec_b200.module.distiller.sequence.module import DistillSequenceModule
    c = synth.make_distill_case(6, 64, [37, 130, 257], 5, seed=81, planted=8.0)
    tables = [dev(t) for t in c["tables"]]
    scores = [torch.from_numpy(orc.teacher_scores(i, s).astype(np.float32)) for i, s in zip(c["teacher_items"], c["teacher_seqs"])]
    on_dev = TeacherScores.from_scores(scores)
    on_host = TeacherScores.from_scores(scores, max_device_bytes=0)
    assert on_dev.scores[0].is_cuda and not on_host.scores[0].is_cuda and on_host.scores[0].is_pinned()
    spec = loss_object("KD", dict(temperature=2.0)).spec
    out = []
    for teacher in (on_dev, on_host):
        keep, ptrs = teacher.rows(c["dataset_indexes"], c["sequence_ids"], tables)
        out.append(fused_distill_losses(dev(c["rep"]), tables, c["dataset_indexes"], ptrs, spec).cpu().numpy())
        del keep
    assert np.array_equal(out[0], out[1])
    narrow = TeacherScores.from_scores([s[:, :-1] for s in scores])
    with pytest.raises(ValueError, match="columns"):
        narrow.rows(c["dataset_indexes"], c["sequence_ids"], tables)
    with pytest.raises(IndexError):
        on_dev.rows(c["dataset_indexes"], [10 ** 6] * len(c["sequence_ids"]), tables)
    assert DistillSequenceModule is not None
