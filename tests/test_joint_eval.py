"""`RecJointModule`-style per-batch evaluation and `RecModule`-style epoch evaluation (reference:
rec_retrieval/module/recommender/module.py:284-503) on the fused evaluator kernels.

CPU: the restated Lightning reduction (batch-size-weighted mean per logged key in float32, then `torch.stack(v).mean()` over
the dataloaders).  GPU: every per-batch metric float equals the oracle's evaluation of that batch (grid catalogs: exact
scores), the aggregate equals the same reduction applied to the oracle's per-batch values, and the catalog cross-entropy
matches torch on the materialised scores."""
import numpy as np
import pytest
import torch

from mergerec_b200 import synth
from oracle import oracle as orc


def test_weighted_mean_log_is_lightnings_mean_reduction():
    from mergerec_b200.module.recommender.module import WeightedMeanLog
    log = WeightedMeanLog()
    vals, sizes = [0.25, 0.5, 1.0 / 3.0], [64, 64, 17]
    for v, n in zip(vals, sizes):
        log.log("val/Recall@10/dataloader_idx_0", v, n)
    acc, cnt = torch.zeros((), dtype=torch.float32), torch.zeros((), dtype=torch.float32)
    for v, n in zip(vals, sizes):
        acc = acc + torch.tensor(v, dtype=torch.float32) * n
        cnt = cnt + n
    got = log.compute()["val/Recall@10/dataloader_idx_0"]
    assert got.dtype == torch.float32 and torch.equal(got, acc / cnt)
    assert abs(float(got) - (0.25 * 64 + 0.5 * 64 + 17 / 3.0) / 145) < 1e-6


@pytest.mark.gpu
def test_joint_evaluation_per_batch_and_aggregate():
    from mergerec_b200.evaluator import Evaluator
    from mergerec_b200.module.recommender.module import RecJointEvaluation, WeightedMeanLog
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    ev = Evaluator(["NDCG", "RECALL"], [1, 10, 50])
    joint = RecJointEvaluation(ev, similarity="dot", temperature=0.05)
    catalogs, batches = [], []
    for i, (n_items, sizes) in enumerate([(3001, [64, 64, 23]), (517, [40, 9])]):
        users, items, labels = synth.make_catalog(sum(sizes), n_items, 64, kind="grid", seed=60 + i)
        catalogs.append(items)
        off = 0
        for b in sizes:
            batches.append((i, users[off:off + b], labels[off:off + b]))
            off += b
    joint.set_item_embeddings([dev(c) for c in catalogs])
    joint.on_epoch_start()
    want_log = WeightedMeanLog()
    for i, u, lab in batches:
        loss = joint.step(dev(u), dev(lab), dataloader_idx=i, stage="test")
        scores = orc.scores_f32(u, catalogs[i])
        want = orc.evaluate(scores, lab, ["NDCG", "RECALL"], [1, 10, 50], "test/")
        assert ev.evaluate_embeddings(dev(u), dev(catalogs[i]), dev(lab), "test/") == want      # per-batch floats bit-exact
        ref_loss = torch.nn.functional.cross_entropy(torch.from_numpy(scores) / 0.05, torch.from_numpy(lab))
        assert abs(float(loss) - float(ref_loss)) <= 1e-5 * max(1.0, abs(float(ref_loss)))
        want_log.log_dict({f"{k}/dataloader_idx_{i}": v for k, v in want.items()}, batch_size=len(lab))
    got, ref = joint.logged_metrics(), want_log.compute()
    for k, v in ref.items():
        assert torch.equal(got[k], v), k
    assert "test/loss/dataloader_idx_1" in got
    agg = joint.on_epoch_end("test")
    assert set(agg) == {"test/NDCG@1", "test/NDCG@10", "test/NDCG@50", "test/Recall@1", "test/Recall@10", "test/Recall@50", "test/loss"}
    for name in ("NDCG@10", "Recall@50"):
        want_v = torch.stack([ref[f"test/{name}/dataloader_idx_0"], ref[f"test/{name}/dataloader_idx_1"]]).mean()
        assert torch.equal(agg[f"test/{name}"], want_v)
    with pytest.raises(RuntimeError):
        RecJointEvaluation(ev, "dot").step(dev(batches[0][1]), dev(batches[0][2]))


@pytest.mark.gpu
def test_rec_evaluation_epoch_equals_full_catalog_evaluation():
    from mergerec_b200.evaluator import Evaluator
    from mergerec_b200.module.recommender.module import RecEvaluation
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    users, items, labels = synth.make_catalog(150, 2000, 32, kind="grid", seed=71)
    ev = Evaluator(["NDCG", "RECALL"], [10, 50])
    rec = RecEvaluation(ev, similarity="dot", temperature=0.05)
    rec.set_item_embeddings(dev(items))
    rec.on_epoch_start()
    for a in range(0, 150, 64):
        rec.step(dev(users[a:a + 64]), dev(labels[a:a + 64]))
    got = rec.on_epoch_end("val")
    scores = orc.scores_f32(users, items)
    want = orc.evaluate(scores, labels, ["NDCG", "RECALL"], [10, 50], "val/")
    loss = got.pop("val/epoch_loss")
    assert got == want
    ref_loss = float(torch.nn.functional.cross_entropy(torch.from_numpy(scores) / 0.05, torch.from_numpy(labels)))
    assert abs(loss - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    # cosine similarity: rows are normalised before scoring (module.py:74-77)
    rc = RecEvaluation(ev, similarity="cosine")
    rc.set_item_embeddings(torch.nn.functional.normalize(dev(items), dim=-1))
    rc.on_epoch_start()
    rc.step(dev(users), dev(labels))
    m = rc.on_epoch_end("test")
    assert "test/loss" in m and 0.0 <= m["test/Recall@50"] <= 1.0
