"""CPU: the fp64 oracle of the distillation step (oracle/oracle.py: distill_loss / distill_step / teacher_scores)
against tests/golden/distill.npz -- outputs of the reference's own loss classes (loss_fn.py) driven by the
`_forward_distill` loop in fp32 on the CPU (tests/golden/make_golden.py: gen_distill).  Plus the host-side behaviour of
the loss factory and the argument checks of the C ABI that need no GPU."""
import ctypes as C

import numpy as np
import pytest

import golden_cases as gc
from helpers import golden
from mergerec_b200 import _lib, synth
from mergerec_b200.merger.enums import LossType
from mergerec_b200.module.recommender import loss_fn as lf
from oracle import oracle as orc

TOL = 2e-5   # fp32 reference vs fp64 oracle, relative to the largest magnitude of the compared array


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("case", gc.DISTILL_CASES, ids=lambda c: c["name"])
def test_oracle_matches_reference_losses_and_gradients(case):
    g = golden("distill")
    c = synth.make_distill_case(case["B"], case["E"], case["rows"], case["n_seq"], case["seed"], planted=case["scale"])
    for d, (i, s) in enumerate(zip(c["teacher_items"], c["teacher_seqs"])):
        assert rel(orc.teacher_scores(i, s), g[f"{case['name']}/score_embeddings{d}"]) < 1e-6
    trows = [g[f"{case['name']}/score_embeddings{d}"][s] for d, s in zip(c["dataset_indexes"], c["sequence_ids"])]
    for lname, kw in gc.DISTILL_LOSSES:
        losses, loss, grad = orc.distill_step(c["rep"], c["tables"], c["dataset_indexes"], trows, lname, **kw)
        assert rel(losses, g[f"{case['name']}/{lname}/losses"]) < TOL, lname
        assert abs(loss - float(g[f"{case['name']}/{lname}/loss"])) < TOL * max(1.0, abs(loss)), lname
        assert rel(grad, g[f"{case['name']}/{lname}/grad_rep"]) < TOL, lname


def test_oracle_argmax_takes_first_maximum():
    z = np.array([0.1, 0.9, 0.3, 0.2], np.float32)
    t = np.array([1.0, 2.0, 2.0, 0.0], np.float32)          # two equal teacher maxima -> label 1 (torch.argmax)
    lv, gz = orc.distill_loss(z, t, "CE")
    assert gz[1] < 0 and gz[2] > 0
    lv, gz = orc.distill_loss(z, t, "PAIRWISE", margin=10.0)  # positive = 1, negative = 2
    assert gz[1] == -1.0 and gz[2] == 1.0


def test_loss_factory_mirrors_the_reference():
    assert isinstance(lf.distill_loss_factory(LossType.CE), lf.DistillCELoss)
    kd = lf.distill_loss_factory(LossType.KD, temperature=3.0)
    assert isinstance(kd, lf.DistillKDLoss) and kd.spec.temperature == 3.0
    with pytest.raises(ValueError, match="Temperature must be provided for KDLoss"):
        lf.distill_loss_factory(LossType.KD)
    with pytest.raises(ValueError, match="Coefficient must be provided"):
        lf.distill_loss_factory(LossType.ADAMERGING_KD, temperature=1.0)
    m = lf.distill_loss_factory(LossType.SINGLE_PSEUDO_LABEL_KD, temperature=2.0, coefficient=0.25)
    assert m.spec == lf.LossSpec(lf.MR_LOSS_SINGLE_PSEUDO_LABEL_KD, 2.0, 0.25, 0.0)
    with pytest.raises(ValueError, match="Unknown loss type"):
        lf.distill_loss_factory("nope")
    assert not lf.DistillAdaMergingLoss().spec.needs_teacher and lf.DistillCELoss().spec.needs_teacher


def test_distill_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    rows = (C.c_int64 * 1)(10)
    ptrs = (C.c_void_p * 1)(64)
    dom = (C.c_int32 * 2)(0, 3)
    assert lib.mr_distill_logits(C.c_void_p(64), 2, 6, ptrs, rows, 1, dom, C.c_void_p(64), 16, None) == -1
    assert b"multiple of 4" in lib.mr_last_error()
    assert lib.mr_distill_logits(C.c_void_p(64), 2, 8, ptrs, rows, 1, dom, C.c_void_p(64), 16, None) == -1
    assert b"dataset index 3" in lib.mr_last_error()
    n = (C.c_int64 * 1)(10)
    assert lib.mr_distill_loss(C.c_void_p(64), 16, None, n, 1, lf.MR_LOSS_KD, 2.0, 0.0, 0.0, C.c_void_p(64), None, 0, None) == -1
    assert b"needs teacher rows" in lib.mr_last_error()
    assert lib.mr_distill_loss(C.c_void_p(64), 16, None, n, 1, 99, 1.0, 0.0, 0.0, C.c_void_p(64), None, 0, None) == -1
    assert b"unknown loss type" in lib.mr_last_error()
    assert lib.mr_distill_grad_workspace_bytes(768) > 0


def test_no_cpu_path_for_the_losses():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MergeRecLibraryError):
        lf.DistillCELoss()(torch.zeros(1, 8), torch.zeros(1, 8))


def test_item_encoding_callbacks_cpu():
    """`encode_items` / `MultiDatasetItemEncodingCallback` (reference callbacks.py:18-50, 85-118): eval mode during the
    encoding, training flag restored, one contiguous fp32 table per dataset, encoded once."""
    import torch
    from types import SimpleNamespace
    from mergerec_b200.module.callbacks import ItemEncoderMixin, MultiDatasetItemEncodingCallback
    from mergerec_b200.module.distiller import DistillSequenceModule
    from mergerec_b200.module.recommender.loss_fn import DistillAdaMergingLoss

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(8, 8, bias=False)
            self.drop = torch.nn.Dropout(0.5)

        def forward(self, x):
            return self.drop(self.lin(x))

        def trainable_parameters(self, **_):
            return list(self.parameters())

    torch.manual_seed(0)
    mod = DistillSequenceModule.__new__(DistillSequenceModule)      # the constructor moves teacher scores to the GPU
    torch.nn.Module.__init__(mod)
    mod.merged_model, mod.similarity, mod.item_embeddings = Enc(), "cosine", None
    mod.loss_fn, mod.learning_rate, mod.trainable_args_kwargs = DistillAdaMergingLoss(), 1e-3, {}
    mod.train()
    loaders = [[SimpleNamespace(items=torch.randn(5, 8)), SimpleNamespace(items=torch.randn(3, 8))],
               [SimpleNamespace(items=torch.randn(4, 8))]]
    cb = MultiDatasetItemEncodingCallback(loaders)
    cb.on_train_epoch_start(None, mod)
    assert mod.training and len(mod.item_embeddings) == 2
    assert mod.item_embeddings[0].shape == (8, 8) and mod.item_embeddings[1].shape == (4, 8)
    assert not mod.item_embeddings[0].requires_grad and mod.item_embeddings[0].is_contiguous()
    want = torch.nn.functional.normalize(mod.merged_model.lin(torch.cat([b.items for b in loaders[0]])), dim=-1)
    assert torch.allclose(mod.item_embeddings[0], want, atol=1e-6)          # dropout was off: eval mode during encoding
    first = mod.item_embeddings[0]
    cb.on_train_epoch_start(None, mod)                                       # already injected: skipped
    assert mod.item_embeddings[0] is first
    assert ItemEncoderMixin.encode_items(loaders[1], mod).shape == (4, 8)
