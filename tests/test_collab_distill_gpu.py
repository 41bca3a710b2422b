"""GPU: stack B of SURVEY.md section 3 end to end through this package -- `load_merging_module` (lambda merge kernel,
parameter views) -> toy encoder -> `DistillSequenceModule._forward_distill` (catalogue-logits / loss / gradient kernels)
-> backward through the encoder -> lambda-gradient kernel -> Adam -- against tests/golden/collab_distill.npz, the same
three steps run with the UNMODIFIED reference (`load_merging_module`, its loss classes, the `_forward_distill` loop) in
fp32 on the CPU.  Adam normalises the gradient, so fp32 noise in it moves a weight by at most ~lr per step: tolerance 2e-4."""
import os
import sys

import numpy as np
import pytest
import torch

import golden_cases as gc
from helpers import golden
from mergerec_b200.merger.enums import LearnType, MergeType
from mergerec_b200.merger.weight_learning.module import load_merging_module
from mergerec_b200.module.distiller import BatchDistillationSequence, DistillSequenceModule, TeacherScores

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
pytestmark = pytest.mark.gpu


def _inputs(case):
    """Same generator as tests/golden/make_golden.py: collab_distill_inputs (kept in sync by hand: the golden script
    imports the reference and cannot be imported on the GPU box)."""
    rng = np.random.Generator(np.random.PCG64(case["seed"] + 9))
    D, E, B, steps, n_seq = len(case["rows"]), 24, 6, 3, 5
    tables = [rng.standard_normal((n, E), dtype=np.float32) for n in case["rows"]]
    tables = [(t / np.linalg.norm(t, axis=-1, keepdims=True)).astype(np.float32) for t in tables]
    t_items = [rng.standard_normal((n, E), dtype=np.float32) for n in case["rows"]]
    t_seqs = [rng.standard_normal((n_seq, E), dtype=np.float32) for _ in case["rows"]]
    ids = rng.integers(0, 37, size=(steps, B, 5)).astype(np.int64)
    dom = rng.integers(0, D, size=(steps, B)).astype(np.int64)
    seq = rng.integers(0, n_seq, size=(steps, B)).astype(np.int64)
    return tables, t_items, t_seqs, ids, dom, seq


@pytest.mark.parametrize("case", gc.COLLAB_DISTILL_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("on_the_fly", [False, True])
def test_collaborative_distillation_steps_match_the_reference(case, on_the_fly):
    from test_distill_gpu import loss_object
    from toy_model import ToyEncoder, make_toy_state_dicts
    g = golden("collab_distill")
    pre, fts = make_toy_state_dicts(case["K"], seed=case["seed"])
    torch.manual_seed(1234)
    merged = load_merging_module(MergeType.TASK_VECTOR, LearnType[case["learn_type"]], ToyEncoder(), pre, fts,
                                 ignore_keys=set(), initial_per_weight=0.3, disable_softmax=True)
    tables, t_items, t_seqs, ids, dom, seq = _inputs(case)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    teacher = TeacherScores([dev(t) for t in t_items], [dev(t) for t in t_seqs], on_the_fly=on_the_fly)
    module = DistillSequenceModule(merged, teacher, loss_object(case["loss"], case["kw"]), similarity="cosine",
                                   learning_rate=1e-2, trainable_args_kwargs=dict(freeze_global_weight=True,
                                                                                  freeze_global_bias=True, freeze_per_weight=False))
    module.item_embeddings = [dev(t) for t in tables]
    opt = module.configure_optimizers()
    traj, losses = [], []
    for step in range(ids.shape[0]):
        opt.zero_grad()
        batch = BatchDistillationSequence(sequence=dev(ids[step]), dataset_indexes=dom[step].tolist(), sequence_ids=seq[step].tolist())
        loss = module.training_step(batch)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
        traj.append(np.stack([merged.per_weights[k].detach().cpu().numpy().copy() for k in merged.per_weights.keys()]))
    np.testing.assert_allclose(np.asarray(losses), g[f"{case['name']}/losses"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(np.stack(traj), g[f"{case['name']}/per_weights_traj"], rtol=0, atol=2e-4)
    assert not np.allclose(traj[-1], 0.3)          # lambda moved
