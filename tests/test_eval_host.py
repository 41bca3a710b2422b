"""CPU tests of the evaluator's host logic: shard arithmetic, the metric finalisation that replicates the reference's
python arithmetic bit for bit (against tests/golden/evaluator.npz), and the multi-GPU exchange step exercised with
world_size-2 gloo on CPU tensors (the oracle stands in for the CUDA kernels, which need a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_cases as gc
from helpers import golden
from mergerec_b200 import synth
from mergerec_b200.evaluator import Evaluator, MetricType, NDCG, Recall, shard_bounds
from mergerec_b200.evaluator.metrics import _gain_table, ndcg_from_ranks, recall_from_ranks
from mergerec_b200.evaluator.sharded import exchange_topk
from oracle import oracle as orc


def test_shard_bounds_partition():
    for n in (0, 1, 7, 1000, 1_000_000, 20011):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_gain_table_matches_reference_expression():
    g = golden("evaluator")
    assert np.array_equal(_gain_table(1024), g["ndcg_gain_table"])


@pytest.mark.parametrize("case", gc.EVAL_CASES, ids=lambda c: c["name"])
def test_metric_finalisation_bit_exact(case):
    """ranks -> python floats must reproduce the reference Evaluator's values (stored golden) exactly."""
    g = golden("evaluator")
    _, _, labels = synth.make_catalog(case["Q"], case["N"], case["E"], kind=case["kind"], seed=case["seed"])
    ids = g[f"{case['name']}/canon_topk"]
    ranks = orc.label_rank(ids, labels)
    ev = Evaluator(case["metrics"], case["ks"])
    got = {case["prefix"] + m.name: m.from_ranks(ranks) for m in ev._metrics}
    assert list(got.keys()) == list(g[f"{case['name']}/canon_keys"])
    assert np.array_equal(np.asarray(list(got.values()), np.float64), g[f"{case['name']}/canon_values"])
    assert got == orc.evaluate_ids(ids, labels, case["metrics"], case["ks"], case["prefix"])


def test_metric_edge_cases():
    empty = np.zeros(0, np.int32)
    assert recall_from_ranks(empty, 10) == 0.0 and ndcg_from_ranks(empty, 10) == 0.0
    ranks = np.array([0, 3, -1, 9, 10], np.int32)
    assert recall_from_ranks(ranks, 10) == 3 / 5 and recall_from_ranks(ranks, 1) == 1 / 5 and recall_from_ranks(ranks, 11) == 4 / 5
    t = _gain_table(16)
    assert ndcg_from_ranks(ranks, 10) == sum([float(t[0]), float(t[3]), 0.0, float(t[9]), 0.0]) / 5  # python-float sum (compensated in 3.12)
    assert MetricType["RECALL"].metric_cls is Recall and MetricType["NDCG"].metric_cls is NDCG
    assert Recall(5).name == "Recall@5" and NDCG(50).name == "NDCG@50"
    with pytest.raises(KeyError):
        Evaluator(["MRR"], [10])


def test_fast_metric_finalisation_equals_the_row_loop():
    """`recall_from_ranks` / `ndcg_from_ranks` skip the rows that miss; the result must be the very float the reference's
    row loop `sum([...per row...]) / len` gives (metrics.py:49-61, 77-88), for any hit pattern."""
    rng = np.random.default_rng(5)
    for _ in range(60):
        n = int(rng.integers(0, 5000))
        k = int(rng.choice([1, 5, 10, 50, 100]))
        ranks = np.where(rng.random(n) < rng.random(), rng.integers(0, 120, size=n), -1).astype(np.int32)
        table = _gain_table(max(k, 1))
        rec, ndcg = [], []
        for r in ranks.tolist():                     # the reference's loop, one python float per row
            hit = 0 <= r < k
            rec.append(1.0 if hit else 0.0)
            ndcg.append(float(table[r]) if hit else 0.0)
        want_r = sum(rec) / len(rec) if rec else 0.0
        want_n = sum(ndcg) / len(ndcg) if ndcg else 0.0
        got_r, got_n = recall_from_ranks(ranks, k), ndcg_from_ranks(ranks, k)
        assert isinstance(got_r, float) and isinstance(got_n, float)
        assert got_r.hex() == float(want_r).hex() and got_n.hex() == float(want_n).hex()
        found = ranks[ranks >= 0]              # the compressed form `metrics_from_ids` uses
        assert Recall(k).from_found(found, n).hex() == float(want_r).hex() and NDCG(k).from_found(found, n).hex() == float(want_n).hex()


def test_float_sum_helper_is_the_interpreters_sum():
    """`mr_float_sum` must return exactly what the running interpreter's builtin `sum()` returns for a list of floats
    (Neumaier-compensated on CPython >= 3.12): random magnitudes, cancellation, long lists, the NDCG gain values."""
    from mergerec_b200.evaluator.metrics import _python_float_sum
    rng = np.random.default_rng(11)
    t = _gain_table(128)
    cases = [rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8) for n in (1, 2, 3, 17, 1000, 65536)]
    cases += [t[rng.integers(0, 100, size=40000)], np.array([1e16, 1.0, -1e16, 1.0]), np.array([0.1] * 1000),
              rng.standard_normal(5000) * np.exp(rng.uniform(-30, 30, 5000))]
    for x in cases:
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert _python_float_sum(x).hex() == float(sum(x.tolist())).hex()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_merge(vals, ids, k):
    v, i = orc.topk_merge(vals.numpy(), ids.numpy())
    return torch.from_numpy(v), torch.from_numpy(i)


def _exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Q, N, E, k = 40, 3001, 16, 25
        users, items, labels = synth.make_catalog(Q, N, E, kind="grid", seed=77)
        lo, hi = shard_bounds(N, world, rank)
        scores = orc.scores_f32(users, items[lo:hi])           # this rank's columns only
        v, i = orc.topk_rows(scores, k, id_base=lo)
        mv, mi = exchange_topk(torch.from_numpy(v), torch.from_numpy(i), k, dist.group.WORLD, merge=_oracle_merge)
        fv, fi = orc.topk_rows(orc.scores_f32(users, items), k)
        ok = np.array_equal(mi.numpy(), fi) and np.array_equal(mv.numpy().view(np.uint32), fv.view(np.uint32))
        ranks = orc.label_rank(mi.numpy(), labels)
        ev = Evaluator(["NDCG", "RECALL"], [5, 25])
        res = {m.name: m.from_ranks(ranks) for m in ev._metrics}
        ok = ok and res == orc.evaluate_ids(fi, labels, ["NDCG", "RECALL"], [5, 25])
        q.put((rank, bool(ok)))
    except Exception as e:  # noqa: BLE001 -- report instead of leaving the parent to time out
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_exchange_topk_gloo_world2():
    """Item table sharded over 2 ranks, per-rank top-K, allgather + merge: identical to the unsharded answer."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
