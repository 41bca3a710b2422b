export MR_BENCH_SKIP_ACCURACY=1
for w in 1 2 3 4 6; do
  MR_SCORE_PACE_TILES=$w timeout 300 python bench.py --workload eval_cfg5 --steps 4 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_pace_$w.json 2> gpurun_out/r2_pace_$w.err
  python -c "
import json; b=json.load(open('gpurun_out/r2_pace_$w.json')); print('pace $w ms/step', round(b['ms_per_step'],1), 'kernel', round(b['roofline']['ms_per_launch'],1), 'clk', b['clocks']['sm_mhz'], b['checksum']['topk_ids'])"
done
