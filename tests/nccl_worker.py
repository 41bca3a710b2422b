"""Multi-GPU worker of tests/test_nccl_gpu.py: one process per GPU under torch.distributed.run, NCCL backend.

Checks, on real GPUs over NCCL/NVLink, the two sharded paths of SURVEY.md section 8(e):
  * evaluator -- item table row-sharded, fused scoring + local top-K per GPU, ONE all-gather of the per-GPU lists,
    `mr_topk_merge_packed`: ids, scores and metric floats must equal the single-GPU answer bit for bit (every rank also
    scores the whole table itself, and rank 0 checks the ids against the CPU oracle);
  * merger -- flat vector column-sharded, global TIES trim from all-reduced histograms: the merged slice must equal the
    corresponding columns of the single-GPU merge bit for bit.
Exit code 0 = all checks passed on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mergerec_b200 import synth  # noqa: E402
from mergerec_b200.evaluator import Evaluator, ShardedItemTable  # noqa: E402
from mergerec_b200.merger.algorithms import get_ties_vectors, merge_ties  # noqa: E402
from mergerec_b200.merger.sharded import flat_shard_bounds, get_ties_vectors_sharded, merge_ties_sharded  # noqa: E402


def main() -> int:
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    try:
        # ---- evaluator: grid (exact, plenty of ties) and gaussian catalogs, a ragged shard split
        for kind, Q, N, E, k in (("grid", 300, 20011, 64, 50), ("gauss", 520, 50021, 768, 100), ("grid", 70, 1000, 1024, 10)):
            users, items, labels = synth.make_catalog(Q, N, E, kind=kind, seed=31)
            tu, ti, tl = (torch.from_numpy(a).to(dev) for a in (users, items, labels))
            ev = Evaluator(["NDCG", "RECALL"], [1, 10, k])
            full_v, full_i = ev.topk_embeddings(tu, ti, k)                                   # single GPU, whole table
            table = ShardedItemTable.from_full(ti, group=dist.group.WORLD)
            assert table.world == world and table.n_total == N
            q = ev.prepare_queries(tu)
            for _ in range(2):                                                               # second call reuses the scratch
                sv, si = ev.topk_embeddings(q, table, k)
                ok &= bool(torch.equal(si, full_i)) and bool(torch.equal(sv.view(torch.int32), full_v.view(torch.int32)))
            # the replicated queries: 1/world per rank over PCIe + one all-gather == a plain copy
            from mergerec_b200.evaluator import replicate_from_host
            ok &= bool(torch.equal(replicate_from_host(torch.from_numpy(users), dist.group.WORLD), tu))
            ok &= bool(torch.equal(replicate_from_host(torch.from_numpy(labels), dist.group.WORLD), tl))
            # host-resident shard streamed in chunks behind the scoring, then the same exchange
            from mergerec_b200.evaluator import shard_bounds
            lo_i, hi_i = shard_bounds(N, world, rank)
            hv, hi_ids = ev.topk_embeddings_streamed(q, torch.from_numpy(items[lo_i:hi_i]), k, id_base=lo_i, n_total=N,
                                                     group=dist.group.WORLD, first_rows=997)
            ok &= bool(torch.equal(hi_ids, full_i)) and bool(torch.equal(hv.view(torch.int32), full_v.view(torch.int32)))
            m_sharded = ev.evaluate_embeddings(q, table, tl)
            m_full = ev.evaluate_embeddings(tu, ti, tl)
            ok &= m_sharded == m_full
            if rank == 0 and kind == "grid":
                from oracle import oracle as orc
                _, oi = orc.topk_rows(orc.scores_f32(users, items), k)
                ok &= bool(np.array_equal(si.cpu().numpy(), oi))
                ok &= m_sharded == orc.evaluate(orc.scores_f32(users, items), labels, ["NDCG", "RECALL"], [1, 10, k])
        # ---- merger: sharded TIES vectors / merge_ties vs the single-GPU kernels on the whole vector
        for d, K, quant in ((200_003, 5, 0.0), (131_072, 8, 2.5e-4)):
            base, models = synth.make_flat(d, K, seed=13, quantize=quant)
            tb, tm = torch.from_numpy(base).to(dev), [torch.from_numpy(m).to(dev) for m in models]
            lo, hi = flat_shard_bounds(d, world, rank)
            full = get_ties_vectors(tb, tm, 0.2)[:, :d]
            part = get_ties_vectors_sharded(tb[lo:hi].clone(), [m[lo:hi].clone() for m in tm], 0.2, d, dist.group.WORLD)
            ok &= bool(torch.equal(part[:, :hi - lo].view(torch.int32), full[:, lo:hi].view(torch.int32)))
            w = [0.3 + 0.1 * i for i in range(K)]
            fullm = merge_ties(tb, tm, w, 0.2)
            partm = merge_ties_sharded(tb[lo:hi].clone(), [m[lo:hi].clone() for m in tm], w, 0.2, d, dist.group.WORLD)
            ok &= bool(torch.equal(partm.view(torch.int32), fullm[lo:hi].view(torch.int32)))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    finally:
        dist.destroy_process_group()
    if rank == 0:
        print("nccl worker:", "ok" if ok else "MISMATCH", f"(world {world})")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
