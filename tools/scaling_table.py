"""profiles/r02_scaling.txt from the committed default bench lines (profiles/r02_bench_default_n{1,2,4,8}.json)."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "profiles", f"r02_bench_default_n{n}.json")
    if os.path.exists(p):
        rows.append((n, json.load(open(p))))
t1 = rows[0][1]["ms_per_step"]
m1 = rows[0][1]["merger"]["ms_per_step"]
print("# BASELINE config 5 (evaluator, item table sharded over N GPUs, one NCCL all-gather + merge) and the nested merger line")
print("# `python bench.py --gpus N ` (N = 1: --steps 20 --warmup 5; N > 1: torch.distributed.run, --steps 10 --warmup 3); CUDA events, max over ranks")
print(f"{'N':>2} {'ms/step':>9} {'scores/s':>11} {'speed-up':>8} {'eff.':>6} {'e2e ms':>8} {'kernel ms':>10} {'gather+merge':>12} {'host ms':>8} {'sm MHz':>7}  checksum(topk_ids, score bits)")
for n, b in rows:
    st = b["stage_ms"]
    print(f"{n:2d} {b['ms_per_step']:9.2f} {b['value']:11.3e} {t1 / b['ms_per_step']:8.2f} {t1 / b['ms_per_step'] / n:6.2f} {b['e2e']['ms_per_step']:8.1f} "
          f"{st['score_topk_kernel_plus_split_merge']:10.2f} {st.get('all_gather_plus_shard_merge', 0.0):12.2f} {st['label_rank_d2h_and_host_metrics_wall']:8.2f} "
          f"{b['clocks']['sm_mhz']:7.0f}  {b['checksum']['topk_ids']}, {b['checksum']['topk_score_bits']}")
print()
print("# merger line of the same runs: N = 1 -> BASELINE config 2 (one GPU); N > 1 -> TIES merge with the flat vector sharded (ties_sharded)")
print(f"{'N':>2} {'ms/step':>9} {'GB/s (algorithmic)':>19}  merged-bits checksum")
for n, b in rows:
    m = b["merger"]
    print(f"{n:2d} {m['ms_per_step']:9.3f} {m['value']:19.0f}  {m.get('merged_bits_checksum')}")
print()
print("# metrics (identical for every N):", json.dumps(rows[0][1]["metrics"]))
