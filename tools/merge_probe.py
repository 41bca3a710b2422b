"""Probe: time mr_topk_merge / mr_topk_merge_packed on sorted and unsorted lists (Q = 65,536, K = 100)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mergerec_b200.evaluator.sharded import topk_merge, topk_merge_packed
Q, K = 65536, 100
g = torch.Generator(device="cuda").manual_seed(3)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
for L in (2, 4, 8):
    vals = torch.randn(L, Q, K, generator=g, device="cuda")
    ids = torch.stack([torch.argsort(torch.rand(Q, K * 4, generator=g, device="cuda"), dim=1)[:, :K].int() + l * 1000 for l in range(L)])
    svals, order = torch.sort(vals, dim=-1, descending=True)
    sids = torch.gather(ids, -1, order)
    packed = torch.stack([svals.view(torch.int32), sids], dim=1).contiguous()
    a = t(lambda: topk_merge(svals, sids, K)); b = t(lambda: topk_merge(vals, ids, K)); c = t(lambda: topk_merge_packed(packed, K))
    v1, i1 = topk_merge(svals, sids, K); v2, i2 = topk_merge(vals, ids, K)
    print(f"L={L}: sorted {a:.3f} ms, unsorted {b:.3f} ms, packed sorted {c:.3f} ms, equal {bool(torch.equal(i1, i2))}", flush=True)
