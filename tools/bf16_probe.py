"""Probe: bf16-compat scoring at BASELINE config 5's size; MR_SCORE_BF16_EPI_GROUPS=1|2 selects the epilogue layout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mergerec_b200.evaluator import Evaluator, ShardedItemTable, MR_SCORE_BF16
from mergerec_b200.evaluator.evaluator import score_topk
Q, N, E, K = 65536, 1_000_000, 768, 100
g = torch.Generator(device="cuda").manual_seed(5)
users = torch.nn.functional.normalize(torch.randn(Q, E, generator=g, device="cuda"), dim=-1)
items = torch.nn.functional.normalize(torch.randn(N, E, generator=g, device="cuda"), dim=-1)
ev = Evaluator(["RECALL"], [K])
table = ShardedItemTable(items, bf16=True)
q = ev.prepare_queries(users, mode=MR_SCORE_BF16)
for _ in range(2):
    v, i = score_topk(q.hi, None, table, K, MR_SCORE_BF16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    v, i = score_topk(q.hi, None, table, K, MR_SCORE_BF16)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"epi groups {os.environ.get('MR_SCORE_BF16_EPI_GROUPS', '2')}: {ms:.1f} ms, {2.0 * Q * N * E / ms / 1e9:.0f} TFLOP/s, checksum {int((i.long() * torch.arange(1, K + 1, device='cuda')).sum())}", flush=True)
