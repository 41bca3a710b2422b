"""Small run of every kernel added in round 2, for `compute-sanitizer --tool memcheck python tools/sanitize_probe.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from mergerec_b200 import _lib, synth
from mergerec_b200.evaluator import Evaluator
from mergerec_b200.evaluator.sharded import topk_merge
from mergerec_b200.merger.algorithms import get_ties_vectors
from mergerec_b200.merger.algorithms.ties import merge_ties_lambda, select_kth_largest
from mergerec_b200.merger.layout import FlatLayout
from mergerec_b200.merger.sharded import DistSelect, flat_shard_bounds
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for K, d in ((8, 700_003), (3, 40_001), (16, 90_005)):          # one-pass select + build: sampled and unsampled, tails
    base, models = synth.make_flat(d, K, seed=K)
    tb, tm = dev(base), [dev(m) for m in models]
    T = get_ties_vectors(tb, tm, 0.2)
    cut = select_kth_largest(tb, tm, int(0.2 * d))
    w1 = torch.rand((1, K), device="cuda")
    merge_ties_lambda(tb, tm, 0.2, w1)
    merge_ties_lambda(tb, tm, 0.2, w1, one_pass=False)
    buf = torch.zeros((K + 1) * (d + 8), device="cuda")        # unaligned pointers: scalar instantiation
    views = [buf[i * (d + 8) + 1: i * (d + 8) + 1 + d] for i in range(K + 1)]
    views[0].copy_(tb)
    for v, m in zip(views[1:], tm):
        v.copy_(m)
    T2 = get_ties_vectors(views[0], views[1:], 0.2)
    assert torch.equal(T2[:, :d].view(torch.int32), T[:, :d].view(torch.int32))
shapes = synth.tiny_shapes(layers=3, hidden=40, ffn=72, vocab=3001, max_pos=18, recformer=True)
layout = FlatLayout.from_shape_dict(shapes)
base, models = synth.make_flat(layout.d, 8, seed=2)
tb, tm = dev(base), [dev(m) for m in models]
se, sg, keys = layout.device_blocks(True, tb.device)
merge_ties_lambda(tb, tm, 0.2, torch.rand((len(keys), 8), device="cuda"), se, sg)
d, K, world = 600_011, 5, 3                                       # phased sharded select, collectives by hand
base, models = synth.make_flat(d, K, seed=9)
fb, fm = dev(base), [dev(m) for m in models]
sels = []
for r in range(world):
    lo, hi = flat_shard_bounds(d, world, r)
    sels.append(DistSelect(fb[lo:hi].clone(), [m[lo:hi].clone() for m in fm], int(0.2 * d), d, lo, None))
gathered = None
for phase in range(DistSelect.PHASES):
    for s in sels:
        s.run(phase, gathered, world)
    if phase < 4:
        tot = torch.stack([s.counters for s in sels]).sum(0, dtype=torch.int32)
        for s in sels:
            s.counters.copy_(tot)
    elif phase == 4:
        gathered = torch.cat([s.survivors for s in sels])
assert torch.equal(sels[0].cut_global, select_kth_largest(fb, fm, int(0.2 * d)))
users, items, labels = synth.make_catalog(300, 5000, 64, kind="grid", seed=3)   # evaluator: streamed table, warp merge
ev = Evaluator(["NDCG", "RECALL"], [10, 50])
a = ev.evaluate_embeddings(dev(users), dev(items), dev(labels))
b = ev.evaluate_embeddings_streamed(dev(users), torch.from_numpy(items), dev(labels), first_rows=700)
assert a == b
vals = torch.sort(torch.randn(8, 50, 100, device="cuda"), dim=-1, descending=True).values
ids = torch.arange(8 * 50 * 100, device="cuda", dtype=torch.int32).reshape(8, 50, 100)
topk_merge(vals, ids, 100); topk_merge(torch.randn(8, 50, 100, device="cuda"), ids, 100); topk_merge(vals[:, :, :37].contiguous(), ids[:, :, :37].contiguous(), 64)
torch.cuda.synchronize()
print("sanitize probe ok")
