"""Timing sweep of the fused scoring kernel (developer tool, GPU box).
    python tools/score_sweep.py [one]      # "one": a single launch of the profiling shape (for ncu)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mergerec_b200.evaluator import ShardedItemTable  # noqa: E402
from mergerec_b200.evaluator.evaluator import score_topk  # noqa: E402
from mergerec_b200.evaluator.sharded import split_tf32  # noqa: E402


def make(Q, N, E):
    g = torch.Generator(device="cuda").manual_seed(1)
    users = torch.nn.functional.normalize(torch.randn(Q, E, generator=g, device="cuda"), dim=-1)
    items = torch.nn.functional.normalize(torch.randn(N, E, generator=g, device="cuda"), dim=-1)
    return split_tf32(users), ShardedItemTable(items)


def timed(Q, N, E, K, mode=0, iters=3, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    (uh, ul), table = make(Q, N, E)
    score_topk(uh, ul, table, K, mode)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        score_topk(uh, ul, table, K, mode)
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / iters
    passes = 3 if mode == 0 else 1
    tf = passes * 2.0 * Q * N * E / (ms * 1e-3) / 1e12
    print(f"Q={Q:6d} N={N:8d} E={E:4d} K={K:3d} mode={mode} {env}: {ms:8.3f} ms  tensor {tf:7.1f} TFLOP/s  "
          f"logical {tf / passes:6.1f}", flush=True)
    for k in env:
        os.environ.pop(k, None)
    return ms


def stamps(Q, N, E, K, mode=0):
    """Per-tile clock64 stamps of block 0 (producer / mma / epilogue)."""
    from mergerec_b200 import _lib
    lib = _lib.load()
    buf = torch.zeros(4 * 64 * 4, dtype=torch.int64, device="cuda")
    (uh, ul), table = make(Q, N, E)
    score_topk(uh, ul, table, K, mode)
    lib.mr_score_topk_debug_buffer(_lib.dptr(buf), buf.numel() * 8)
    score_topk(uh, ul, table, K, mode)
    torch.cuda.synchronize()
    lib.mr_score_topk_debug_buffer(None, 0)
    b = buf.cpu().numpy().reshape(4, 64, 4)
    t0 = b[:3][b[:3] > 0].min()
    print(f"--- stamps Q={Q} N={N} E={E} K={K} mode={mode} (cycles since first stamp)")
    print("tile | prod: start emptyok lastkb | mma: start temptyok kb0 last | epi: start tfullok drained")
    for t in range(0, 40):
        if b[1, t, 0] == 0:
            break
        r = lambda x: int(x - t0) if x > 0 else -1  # noqa: E731
        print(f"{t:3d} | {r(b[0,t,0]):8d} {r(b[0,t,1]):8d} {r(b[0,t,2]):8d} | {r(b[1,t,0]):8d} {r(b[1,t,1]):8d} {r(b[1,t,2]):8d} "
              f"{r(b[1,t,3]):8d} | {r(b[2,t,0]):8d} {r(b[2,t,1]):8d} {r(b[2,t,2]):8d} | ld {b[3,t,0]:6d} flt {b[3,t,1]:6d} cmp {b[3,t,2]:7d} n {b[3,t,3]:3d}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        E = int(sys.argv[2]) if len(sys.argv) > 2 else 768
        K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
        (uh, ul), table = make(2048, 65536, E)
        for _ in range(2):
            score_topk(uh, ul, table, K, 0)
        torch.cuda.synchronize()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "stamps":
        stamps(2048, 65536, 128, 100)
        stamps(2048, 262144, 768, 100)
        sys.exit(0)
    timed(2048, 262144, 768, 100)
    timed(2048, 262144, 768, 100, MR_SCORE_BK=16)
    timed(2048, 500000, 768, 100)
    timed(2048, 500000, 768, 100, MR_SCORE_BK=16)
    timed(2048, 262144, 768, 10)
    timed(2048, 262144, 768, 10, MR_SCORE_BK=16)
    timed(8192, 131072, 768, 100)
    timed(8192, 131072, 768, 100, MR_SCORE_BK=16)
    timed(16384, 65536, 768, 100)
    timed(2048, 262144, 1024, 50)
    timed(2048, 262144, 1024, 50, MR_SCORE_BK=16)
    timed(2048, 262144, 768, 100, mode=1)
    timed(2048, 262144, 768, 100, MR_SCORE_CTA_GROUP=1)
    timed(2048, 262144, 768, 100, MR_SCORE_CTA_GROUP=1, MR_SCORE_BK=16)
    timed(256, 20000, 768, 10)
    timed(2048, 262144, 128, 100)
