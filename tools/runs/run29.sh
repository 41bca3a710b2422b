export MR_BENCH_SKIP_ACCURACY=1
for h in 1 2 3 4 1; do
  MR_SCORE_L2HINT=$h python bench.py --workload eval_cfg5 --steps 4 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_hint_$h.json 2> gpurun_out/r2_hint_$h.err
  echo "hint $h rc=$?"
  python -c "
import json; b=json.load(open('gpurun_out/r2_hint_$h.json')); print('hint $h ms/step', round(b['ms_per_step'],1), 'kernel', round(b['roofline']['ms_per_launch'],1), 'clk', b['clocks']['sm_mhz'])"
done
