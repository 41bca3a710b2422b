"""Diagnostic for the fused scoring kernel (developer tool, GPU box): runs small cases for both CTA-group
variants and prints where the results first differ from the oracle.  Usage: python tools/score_debug.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mergerec_b200 import synth  # noqa: E402
from mergerec_b200.evaluator import ShardedItemTable  # noqa: E402
from mergerec_b200.evaluator.evaluator import score_topk  # noqa: E402
from mergerec_b200.evaluator.sharded import split_tf32  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def run(Q, N, E, k, cg, mode=0, splits=0):
    os.environ["MR_SCORE_CTA_GROUP"] = str(cg)
    os.environ["MR_SCORE_SPLITS"] = str(splits)
    users, items, _ = synth.make_catalog(Q, N, E, kind="grid", seed=21)
    scores = orc.scores_f32(users, items)
    ov, oi = orc.topk_rows(scores, k)
    table = ShardedItemTable(torch.from_numpy(items).cuda())
    uh, ul = split_tf32(torch.from_numpy(users).cuda())
    try:
        v, i = score_topk(uh, ul, table, k, mode)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"Q={Q} N={N} E={E} k={k} cg={cg} mode={mode} splits={splits}: EXCEPTION {e}")
        return False
    v, i = v.cpu().numpy(), i.cpu().numpy()
    bad_i = (i != oi)
    bad_v = (v.view(np.uint32) != ov.view(np.uint32))
    print(f"Q={Q} N={N} E={E} k={k} cg={cg} mode={mode} splits={splits}: id mismatches {int(bad_i.sum())}/{i.size}, "
          f"value mismatches {int(bad_v.sum())}")
    if bad_i.any() or bad_v.any():
        r = int(np.argwhere(bad_i | bad_v)[0, 0])
        print("  first bad row", r, "\n   got ids ", i[r][:12], "\n   want ids", oi[r][:12], "\n   got v ", v[r][:8], "\n   want v", ov[r][:8])
        rows = np.unique(np.argwhere(bad_i | bad_v)[:, 0])
        print("  bad rows:", rows[:40], "count", len(rows))
        return False
    return True


if __name__ == "__main__":
    ok = True
    for cg in (1, 2):
        for shape in [(64, 256, 32, 5), (128, 512, 64, 10), (300, 3000, 96, 20), (256, 20000, 768, 10)]:
            ok &= run(*shape, cg=cg)
            if not ok:
                break
        ok &= run(128, 512, 64, 10, cg=cg, mode=1)
    print("ALL OK" if ok else "FAILURES")
