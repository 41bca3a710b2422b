"""Probe: time merge_ties_lambda (one-pass fused TIES + per-layer lambda merge) on K=8 BLaIR-base shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mergerec_b200 import synth
from mergerec_b200.merger.algorithms import ties as T
from mergerec_b200.merger.layout import FlatLayout
shapes = synth.roberta_shapes(); d = synth.total_numel(shapes); K = 8
g = torch.Generator(device="cuda").manual_seed(1)
base = torch.randn(d, generator=g, device="cuda") * 0.02
models = [base + 1e-3 * torch.randn(d, generator=g, device="cuda") for _ in range(K)]
layout = FlatLayout.from_shape_dict(shapes)
seg_end, seg_group, keys = layout.device_blocks(True, base.device)
w = torch.rand((len(keys), K), device="cuda") * 0.4 + 0.1
out = torch.empty_like(base)
for lw in (True, False):
    a = (seg_end, seg_group) if lw else (None, None)
    ww = w if lw else w[:1].contiguous()
    for _ in range(2):
        T.merge_ties_lambda(base, models, 0.2, ww, a[0], a[1], out=out, one_pass=bool(int(os.environ.get('MR_ONE_PASS', '1'))))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        T.merge_ties_lambda(base, models, 0.2, ww, a[0], a[1], out=out, one_pass=bool(int(os.environ.get('MR_ONE_PASS', '1'))))
    e1.record(); e1.synchronize()
    print("layer-wise" if lw else "task-wise", "merge_ties_lambda ms", e0.elapsed_time(e1) / 10, flush=True)
