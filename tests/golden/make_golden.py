"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference)
on seeded synthetic inputs.  Run in the build container only (the reference does not travel):

    python tests/golden/make_golden.py

Inputs are regenerated in the tests from the same numpy (PCG64) seeds via mergerec_b200.synth, so
only reference OUTPUTS are stored.  The oracle process is pinned to ATEN_CPU_CAPABILITY=avx2 so the
bits do not depend on the host CPU.
"""
import os
import sys

os.environ.setdefault("ATEN_CPU_CAPABILITY", "avx2")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mergerec_b200 import synth  # noqa: E402
from rec_retrieval.evaluator import Evaluator  # noqa: E402
from rec_retrieval.evaluator.metrics import NDCG, Recall  # noqa: E402
from rec_retrieval.merger import ModelMerger  # noqa: E402
from rec_retrieval.merger.algorithms.linear import merge_linear  # noqa: E402
from rec_retrieval.merger.algorithms.localize_and_stitch import (  # noqa: E402
    get_localize_and_stitch_vectors, merge_localize_and_stitch)
from rec_retrieval.merger.algorithms.task_vector import get_task_vectors, merge_task_vector  # noqa: E402
from rec_retrieval.merger.algorithms.ties import get_ties_vectors, merge_ties  # noqa: E402
from rec_retrieval.merger.enums import LearnType, MergeType  # noqa: E402
from rec_retrieval.merger.weight_learning import load_merging_module  # noqa: E402
from rec_retrieval.merger.weight_learning.module.layer_wise import TaskVectorMergingModuleLayerWise  # noqa: E402
from rec_retrieval.merger.weight_learning.module.task_wise import TaskVectorMergingModuleTaskWise  # noqa: E402

import golden_cases as gc  # noqa: E402  (tests/golden_cases.py: the shared case definitions)

T = torch.from_numpy


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  {name}.npz  {os.path.getsize(path) / 1024:.1f} KiB")


def gen_merge_flat():
    out = {}
    for case in gc.MERGE_FLAT_CASES:
        base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"])
        tb, tm = T(base), [T(m) for m in models]
        out[f"{case['name']}/task_vector"] = merge_task_vector(tb, tm, case["weights"]).numpy()
        out[f"{case['name']}/linear"] = merge_linear(models=tm, weights=case["weights"]).numpy()
        out[f"{case['name']}/task_vectors"] = get_task_vectors(tb, tm).numpy()
    save("merge_flat", **out)


def gen_model_merger():
    out = {}
    for case in gc.MODEL_MERGER_CASES:
        shapes = synth.tiny_shapes(recformer=case["recformer"])
        base, models = synth.make_state_dicts(shapes, case["K"], seed=case["seed"], sigma=1e-2)
        tbase = {k: T(v) for k, v in base.items()}
        tmodels = [{k: T(v) for k, v in m.items()} for m in models]
        merger = ModelMerger(tmodels, tbase)  # align_key_order=True -> sorted keys
        out[f"{case['name']}/keys"] = np.array(list(merger.shape_dict.keys()))
        out[f"{case['name']}/base_flat"] = merger.base_model.numpy()
        for mt, w, kw in case["merges"]:
            sd = merger.merge(mt, w, **kw)
            flat = torch.cat([v.reshape(-1) for v in sd.values()]).numpy()
            out[f"{case['name']}/{mt}"] = flat
    save("model_merger", **out)


class _NoModel(torch.nn.Module):
    def forward(self, x):
        return x


def gen_lambda():
    out = {}
    for case in gc.LAMBDA_CASES:
        shapes = synth.tiny_shapes(recformer=case["recformer"])
        base, models = synth.make_state_dicts(shapes, case["K"], seed=case["seed"], sigma=1e-2)
        tbase = {k: T(v) for k, v in base.items()}
        tmodels = [{k: T(v) for k, v in m.items()} for m in models]
        merger = ModelMerger(tmodels, tbase, align_key_order=False)
        tv = get_task_vectors(merger.base_model, merger.models)
        K = case["K"]
        rng = np.random.Generator(np.random.PCG64(case["seed"] + 100))
        for learn in ("task", "layer"):
            cls = TaskVectorMergingModuleTaskWise if learn == "task" else TaskVectorMergingModuleLayerWise
            for softmax in (False, True):
                mod = cls(merger.base_model, tv, _NoModel(), merger.shape_dict, disable_softmax=not softmax)
                keys = list(mod.per_weights.keys())
                with torch.no_grad():
                    for key in keys:
                        mod.per_weights[key].copy_(T(rng.uniform(0.1, 0.5, size=K).astype(np.float32)))
                        mod.global_weights[key].fill_(float(rng.uniform(0.8, 1.2)))
                        mod.global_biases[key].fill_(float(rng.uniform(-0.05, 0.05)))
                merged = mod._merge_task_vectors()
                g = T(rng.standard_normal(merged.numel(), dtype=np.float32))
                (merged * g).sum().backward()
                tag = f"{case['name']}/{learn}/softmax{int(softmax)}"
                out[f"{tag}/keys"] = np.array(keys)
                out[f"{tag}/per_weights"] = np.stack([mod.per_weights[k].detach().numpy() for k in keys])
                out[f"{tag}/global_weights"] = np.stack([mod.global_weights[k].detach().numpy() for k in keys])
                out[f"{tag}/global_biases"] = np.stack([mod.global_biases[k].detach().numpy() for k in keys])
                # the effective (G,K) lambda exactly as torch computed it (fed verbatim to kernels)
                ws = []
                for key in keys:
                    pw = mod.per_weights[key].detach()
                    if softmax:
                        pw = torch.softmax(pw, dim=0)
                    ws.append((mod.global_weights[key].detach() * pw + mod.global_biases[key].detach()).numpy())
                out[f"{tag}/w"] = np.stack(ws)
                out[f"{tag}/merged"] = merged.detach().numpy()
                out[f"{tag}/grad_out"] = g.numpy()
                out[f"{tag}/grad_per_weights"] = np.stack([mod.per_weights[k].grad.numpy() for k in keys])
                out[f"{tag}/grad_global_weights"] = np.stack([mod.global_weights[k].grad.numpy() for k in keys])
                out[f"{tag}/grad_global_biases"] = np.stack([mod.global_biases[k].grad.numpy() for k in keys])
    save("lambda_merge", **out)


def gen_ties():
    out = {}
    for case in gc.TIES_CASES:
        base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"], tie_free=case["tie_free"],
                                       quantize=case.get("quantize", 0.0))
        tb, tm = T(base), [T(m) for m in models]
        out[f"{case['name']}/ties_vectors"] = get_ties_vectors(tb, tm, case["density"]).numpy()
        out[f"{case['name']}/merge_ties"] = merge_ties(tb, tm, case["weights"], case["density"]).numpy()
    save("ties", **out)


def gen_lns():
    out = {}
    for case in gc.LNS_CASES:
        base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"], tie_free=case["tie_free"])
        tb, tm = T(base), [T(m) for m in models]
        out[f"{case['name']}/vectors"] = get_localize_and_stitch_vectors(tb, tm, case["density"]).numpy()
        out[f"{case['name']}/merged"] = merge_localize_and_stitch(tb, tm, case["weights"], case["density"]).numpy()
    save("lns", **out)


def gen_evaluator():
    out = {}
    table = [1 / (torch.log2(torch.tensor(r + 2)).item()) for r in range(1024)]
    out["ndcg_gain_table"] = np.asarray(table, np.float64)
    for case in gc.EVAL_CASES:
        users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
        scores = T(users) @ T(items).T
        tl = T(labels)
        ev = Evaluator(case["metrics"], case["ks"])
        res = ev(scores, tl, metric_prefix=case["prefix"])
        out[f"{case['name']}/raw_keys"] = np.array(list(res.keys()))
        out[f"{case['name']}/raw_values"] = np.asarray(list(res.values()), np.float64)
        out[f"{case['name']}/raw_topk"] = torch.topk(scores, max(case["ks"]), dim=1).indices.numpy()
        # canonical ids: stable descending sort of the reference's fp32 scores (lowest id wins ties)
        canon = torch.sort(scores, dim=1, descending=True, stable=True).indices[:, : max(case["ks"])]
        out[f"{case['name']}/canon_topk"] = canon.numpy().astype(np.int32)
        vals = {}
        for m in case["metrics"]:
            for k in case["ks"]:
                obj = {"RECALL": Recall, "NDCG": NDCG}[m](k)
                vals[case["prefix"] + obj.name] = obj(y_true=tl, y_pred=canon)
        out[f"{case['name']}/canon_keys"] = np.array(list(vals.keys()))
        out[f"{case['name']}/canon_values"] = np.asarray(list(vals.values()), np.float64)
    save("evaluator", **out)


def gen_evaluator_cfg4():
    """BASELINE config 4's evaluator shape: the reference Evaluator on its own CPU scores; canonical ids from a stable
    descending sort of those scores, metric floats from the reference's Recall / NDCG objects fed the canonical ids."""
    out = {}
    for case in gc.EVAL_CFG4_CASES:
        users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
        scores = T(users) @ T(items).T
        tl = T(labels)
        kmax = max(case["ks"])
        res = Evaluator(case["metrics"], case["ks"])(scores, tl, metric_prefix=case["prefix"])
        out[f"{case['name']}/raw_keys"] = np.array(list(res.keys()))
        out[f"{case['name']}/raw_values"] = np.asarray(list(res.values()), np.float64)
        order = torch.sort(scores, dim=1, descending=True, stable=True)
        canon = order.indices[:, :kmax]
        out[f"{case['name']}/canon_topk"] = canon.numpy().astype(np.int32)
        out[f"{case['name']}/canon_vals"] = order.values[:, :kmax].numpy()
        raw = torch.topk(scores, kmax, dim=1)
        assert torch.equal(raw.values, order.values[:, :kmax])       # same score multiset per row
        vals = {}
        for m in case["metrics"]:
            for k in case["ks"]:
                obj = {"RECALL": Recall, "NDCG": NDCG}[m](k)
                vals[case["prefix"] + obj.name] = obj(y_true=tl, y_pred=canon)
        out[f"{case['name']}/canon_keys"] = np.array(list(vals.keys()))
        out[f"{case['name']}/canon_values"] = np.asarray(list(vals.values()), np.float64)
    save("evaluator_cfg4", **out)


def gen_evaluator_bf16():
    """The reference's default bf16-mixed pipeline on grid catalogs: scores = (U.bf16 @ I.bf16.T) as torch computes them
    (bf16 result), canonical ids = stable descending sort, metrics from the reference's own Recall / NDCG objects."""
    out = {}
    for case in gc.EVAL_CASES:
        if case["kind"] != "grid":
            continue
        users, items, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
        scores = T(users).to(torch.bfloat16) @ T(items).to(torch.bfloat16).T
        assert scores.dtype == torch.bfloat16
        kmax = max(case["ks"])
        order = torch.sort(scores.float(), dim=1, descending=True, stable=True)
        canon = order.indices[:, :kmax]
        out[f"{case['name']}/canon_topk"] = canon.numpy().astype(np.int32)
        out[f"{case['name']}/canon_vals"] = order.values[:, :kmax].numpy()
        tl = T(labels)
        vals = {}
        for m in case["metrics"]:
            for k in case["ks"]:
                obj = {"RECALL": Recall, "NDCG": NDCG}[m](k)
                vals[case["prefix"] + obj.name] = obj(y_true=tl, y_pred=canon)
        out[f"{case['name']}/canon_keys"] = np.array(list(vals.keys()))
        out[f"{case['name']}/canon_values"] = np.asarray(list(vals.values()), np.float64)
        # raw torch.topk on the bf16 scores picks the same score multiset per row
        raw = torch.topk(scores, kmax, dim=1)
        assert torch.equal(raw.values.float(), order.values[:, :kmax])
    save("evaluator_bf16", **out)


def gen_module_e2e():
    """load_merging_module on the toy encoder + 3 Adam steps on lambda (stack B of SURVEY.md section 3)."""
    from toy_model import ToyEncoder, make_toy_state_dicts

    out = {}
    for case in gc.MODULE_CASES:
        pre, fts = make_toy_state_dicts(case["K"], seed=case["seed"])
        torch.manual_seed(1234)
        model = ToyEncoder()
        mod = load_merging_module(
            MergeType[case["merge_type"]], LearnType[case["learn_type"]], model, pre, fts, ignore_keys=set(),
            ties_density=case.get("density"), initial_per_weight=0.3, disable_softmax=case["disable_softmax"],
        )
        rng = np.random.Generator(np.random.PCG64(case["seed"] + 7))
        ids = T(rng.integers(0, 37, size=(3, 6, 5)).astype(np.int64))
        tgt = T(rng.standard_normal((3, 6, 24), dtype=np.float32))
        opt = torch.optim.Adam(mod.trainable_parameters(True, True, False), lr=1e-2)
        traj, losses = [], []
        out[f"{case['name']}/merged0"] = mod._merge_task_vectors().detach().numpy()
        for step in range(3):
            opt.zero_grad()
            rep = mod(ids[step])
            loss = ((rep - tgt[step]) ** 2).mean()
            loss.backward()
            opt.step()
            losses.append(float(loss))
            traj.append(np.stack([mod.per_weights[k].detach().numpy().copy() for k in mod.per_weights.keys()]))
        out[f"{case['name']}/keys"] = np.array(list(mod.per_weights.keys()))
        out[f"{case['name']}/per_weights_traj"] = np.stack(traj)
        out[f"{case['name']}/losses"] = np.asarray(losses, np.float64)
        sd = mod.get_state_dict()
        out[f"{case['name']}/final_flat"] = torch.cat([v.reshape(-1) for v in sd.values()]).detach().numpy()
        out[f"{case['name']}/sd_keys"] = np.array(list(sd.keys()))
    save("module_e2e", **out)


def _load_reference_loss_fn():
    """rec_retrieval/module/recommender/loss_fn.py without its package __init__ (which pulls in lightning / peft, absent
    here): the parent packages are registered as empty namespaces, the file itself is executed unmodified."""
    import importlib.util
    import types
    import rec_retrieval.merger.enums  # noqa: F401  (the real module the file imports)
    for name in ("rec_retrieval.module", "rec_retrieval.module.recommender"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = ["/root/reference/" + name.replace(".", "/")]
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("rec_retrieval.module.recommender.loss_fn",
                                                  "/root/reference/rec_retrieval/module/recommender/loss_fn.py")
    lf = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = lf
    spec.loader.exec_module(lf)
    return lf


def reference_loss(lf, name, kw):
    cls = {"CE": lf.DistillCELoss, "KD": lf.DistillKDLoss, "MSE": lf.DistillMSELoss, "ADAMERGING": lf.DistillAdaMergingLoss,
           "ADAMERGING_KD": lf.DistillAdaMergingKDLoss, "MERGED_PSEUDO_LABEL": lf.MergedPseudoLabelLoss,
           "MERGED_PSEUDO_LABEL_KD": lf.MergedPseudoLabelKDLoss, "SINGLE_PSEUDO_LABEL": lf.SinglePseudoLabelLoss,
           "SINGLE_PSEUDO_LABEL_KD": lf.SinglePseudoLabelKDLoss, "PAIRWISE": lf.DistillPairwiseLoss,
           "LISTNET": lf.DistillListNetLoss}[name]
    return cls(**kw)


def gen_distill():
    """The reference's own loss classes driven by the loop of DistillSequenceModule._forward_distill
    (sequence/module.py:59-76; the module itself needs lightning) with teacher logits built as merge_train.py:116-126
    does, in fp32 on the CPU: per-sample losses, the batch loss and its gradient w.r.t. the representations."""
    lf = _load_reference_loss_fn()
    out = {}
    for case in gc.DISTILL_CASES:
        c = synth.make_distill_case(case["B"], case["E"], case["rows"], case["n_seq"], case["seed"], planted=case["scale"])
        score_embeddings = []
        for item_embedding, sequence_embedding in zip(c["teacher_items"], c["teacher_seqs"]):
            item_embedding, sequence_embedding = T(item_embedding), T(sequence_embedding)
            item_embedding = item_embedding / item_embedding.norm(dim=-1, keepdim=True)
            sequence_embedding = sequence_embedding / sequence_embedding.norm(dim=-1, keepdim=True)
            score_embeddings.append(sequence_embedding @ item_embedding.T)
        for d, s in enumerate(score_embeddings):
            out[f"{case['name']}/score_embeddings{d}"] = s.numpy()
        item_embeddings = [T(t) for t in c["tables"]]
        for lname, kw in gc.DISTILL_LOSSES:
            loss_fn = reference_loss(lf, lname, kw)
            rep = T(c["rep"]).clone().requires_grad_(True)
            losses = []
            for i, (dataset_index, sequence_id) in enumerate(zip(c["dataset_indexes"], c["sequence_ids"])):
                merged_model_logit = rep[i] @ item_embeddings[dataset_index].T
                single_model_logit = score_embeddings[dataset_index][sequence_id]
                losses.append(loss_fn(merged_model_logit.unsqueeze(0), single_model_logit.unsqueeze(0)))
            loss = torch.stack(losses).mean()
            loss.backward()
            out[f"{case['name']}/{lname}/losses"] = torch.stack(losses).detach().numpy()
            out[f"{case['name']}/{lname}/loss"] = np.asarray(float(loss), np.float64)
            out[f"{case['name']}/{lname}/grad_rep"] = rep.grad.numpy()
    save("distill", **out)


def gen_pcb():
    from rec_retrieval.merger.algorithms.pcb import get_pcb_vectors, merge_pcb
    out = {}
    for case in gc.PCB_CASES:
        base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"])
        tb, tm = T(base), [T(m) for m in models]
        out[f"{case['name']}/vectors"] = get_pcb_vectors(tb, tm, density=case["density"]).numpy()
        out[f"{case['name']}/merged"] = merge_pcb(tb, tm, case["weights"], density=case["density"]).numpy()
    save("pcb", **out)


def gen_dare():
    from rec_retrieval.merger.algorithms.dare import merge_dare
    from torch.nn.functional import dropout
    out = {}
    for case in gc.DARE_CASES:
        base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"])
        tb, tm = T(base), [T(m) for m in models]
        torch.manual_seed(case["torch_seed"])
        out[f"{case['name']}/merged"] = merge_dare(tb, tm, case["weights"], density=case["density"]).numpy()
        # replay the generator: the mask of a dropout call depends on the shape and the RNG state only
        torch.manual_seed(case["torch_seed"])
        masks = np.stack([(dropout(torch.ones(case["d"]), p=case["density"], training=True) != 0).numpy() for _ in range(case["K"])])
        out[f"{case['name']}/masks"] = np.packbits(masks, axis=1)
    save("dare", **out)


def collab_distill_inputs(case):
    """Seeded inputs shared with tests/test_collab_distill_gpu.py."""
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


def gen_collab_distill():
    """The reference's load_merging_module + its loss classes + the `_forward_distill` loop (sequence/module.py:59-76,
    cosine similarity) + Adam on lambda, three steps on the toy encoder: lambda trajectory and losses."""
    from toy_model import ToyEncoder, make_toy_state_dicts
    lf = _load_reference_loss_fn()
    out = {}
    for case in gc.COLLAB_DISTILL_CASES:
        pre, fts = make_toy_state_dicts(case["K"], seed=case["seed"])
        torch.manual_seed(1234)
        mod = load_merging_module(MergeType.TASK_VECTOR, LearnType[case["learn_type"]], ToyEncoder(), pre, fts,
                                  ignore_keys=set(), initial_per_weight=0.3, disable_softmax=True)
        tables, t_items, t_seqs, ids, dom, seq = collab_distill_inputs(case)
        score_embeddings = []
        for ie, se in zip(t_items, t_seqs):
            ie, se = T(ie), T(se)
            ie = ie / ie.norm(dim=-1, keepdim=True)
            se = se / se.norm(dim=-1, keepdim=True)
            score_embeddings.append(se @ ie.T)
        item_embeddings = [T(t) for t in tables]
        loss_fn = reference_loss(lf, case["loss"], case["kw"])
        opt = torch.optim.Adam(mod.trainable_parameters(True, True, False), lr=1e-2, weight_decay=0.0)
        traj, losses = [], []
        for step in range(ids.shape[0]):
            opt.zero_grad()
            rep = torch.nn.functional.normalize(mod.forward(T(ids[step])), p=2, dim=-1)
            per = []
            for i, (d_i, s_i) in enumerate(zip(dom[step].tolist(), seq[step].tolist())):
                merged_model_logit = rep[i] @ item_embeddings[d_i].T
                single_model_logit = score_embeddings[d_i][s_i]
                per.append(loss_fn(merged_model_logit.unsqueeze(0), single_model_logit.unsqueeze(0)))
            loss = torch.stack(per).mean()
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
            traj.append(np.stack([mod.per_weights[k].detach().numpy().copy() for k in mod.per_weights.keys()]))
        out[f"{case['name']}/per_weights_traj"] = np.stack(traj)
        out[f"{case['name']}/losses"] = np.asarray(losses, np.float64)
    save("collab_distill", **out)


if __name__ == "__main__":
    torch.set_num_threads(4)
    print("torch", torch.__version__, "cpu capability", torch.backends.cpu.get_cpu_capability())
    only = set(sys.argv[1:])   # e.g. `make_golden.py lns` regenerates one file
    for name, fn in [("merge_flat", gen_merge_flat), ("model_merger", gen_model_merger), ("lambda_merge", gen_lambda),
                     ("ties", gen_ties), ("lns", gen_lns), ("evaluator", gen_evaluator), ("evaluator_bf16", gen_evaluator_bf16), ("evaluator_cfg4", gen_evaluator_cfg4),
                     ("module_e2e", gen_module_e2e), ("distill", gen_distill), ("pcb", gen_pcb), ("dare", gen_dare), ("collab_distill", gen_collab_distill)]:
        if not only or name in only:
            fn()
