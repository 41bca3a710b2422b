"""Probe: time get_pcb_vectors on K=8 BLaIR-base shapes (argument: repetitions, default 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mergerec_b200 import synth
from mergerec_b200.merger.algorithms.pcb import get_pcb_vectors
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
shapes = synth.roberta_shapes(); d = synth.total_numel(shapes); K = 8
g = torch.Generator(device="cuda").manual_seed(1)
base = torch.randn(d, generator=g, device="cuda") * 0.02
models = [base + 1e-3 * torch.randn(d, generator=g, device="cuda") for _ in range(K)]
for _ in range(2):
    out = get_pcb_vectors(base, models, 0.2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = get_pcb_vectors(base, models, 0.2)
e1.record(); e1.synchronize()
print("get_pcb_vectors ms", e0.elapsed_time(e1) / reps, "checksum", int(out.view(torch.int32).to(torch.int64).sum().item()), flush=True)
if len(sys.argv) <= 2:
    out2 = get_pcb_vectors(base, models, 0.2, force_ieee=True)
    print("prepared-reciprocal divisions == IEEE divides, bit for bit:", bool(torch.equal(out.view(torch.int32), out2.view(torch.int32))), flush=True)
