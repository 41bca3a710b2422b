"""Developer probe (GPU box): time TIES select / build variants separately."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mergerec_b200 import _lib, synth
from mergerec_b200.merger.algorithms import ties as T
from mergerec_b200.merger.layout import FlatLayout, alloc_rows

def t_ms(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n

shapes = synth.roberta_shapes(); d = synth.total_numel(shapes); K = 8
g = torch.Generator(device="cuda").manual_seed(1234)
base = torch.randn(d, generator=g, device="cuda") * 0.02
models = [base + 1e-3 * torch.randn(d, generator=g, device="cuda") for _ in range(K)]
layout = FlatLayout.from_shape_dict(shapes)
seg_end, seg_group, keys = layout.device_blocks(True, base.device)
w = torch.rand((len(keys), K), device="cuda") * 0.4 + 0.1
lib = _lib.load()
ws_bytes = int(lib.mr_ties_workspace_bytes(d, K)); print("ws MB", ws_bytes / 1e6)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
cut = torch.empty(K, dtype=torch.int64, device="cuda"); status = torch.zeros(K, dtype=torch.int32, device="cuda")
parr = _lib.ptr_array(models)
def sel():
    _lib.check(lib.mr_ties_select(_lib.dptr(base), parr, K, d, None, int(0.2 * d), _lib.dptr(cut), _lib.dptr(status), _lib.dptr(ws), ws_bytes, _lib.stream_handle()), "sel")
print("select fast ms", t_ms(sel), "status", status.tolist())
print("python ties_select ms", t_ms(lambda: T.ties_select(base, models, 0.2)))
out = torch.empty(d, device="cuda"); That = alloc_rows(K, d, base.device)
print("build VECTORS ms", t_ms(lambda: T._build(base, models, cut, _lib.MR_TIES_VECTORS, out=That, ldo=That.stride(0))))
print("build FUSED ms", t_ms(lambda: T._build(base, models, cut, _lib.MR_TIES_FUSED_MERGE, w=w, G=w.shape[0], seg_end=seg_end, seg_group=seg_group, out=out)))
print("build FUSED task-wise ms", t_ms(lambda: T._build(base, models, cut, _lib.MR_TIES_FUSED_MERGE, w=w[:1].contiguous(), G=1, out=out)))
print("build LNS ms", t_ms(lambda: T._build(base, models, cut, _lib.MR_TIES_LNS, out=That, ldo=That.stride(0))))
wt = torch.full((K,), 0.5, device="cuda")
print("build TRIMSUM ms", t_ms(lambda: T._build(base, models, cut, _lib.MR_TIES_TRIMSUM, w=wt, out=out)))
print("merge_ties_lambda ms", t_ms(lambda: T.merge_ties_lambda(base, models, 0.2, w, seg_end, seg_group, out=out)))
