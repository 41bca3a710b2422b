export MR_BENCH_SKIP_ACCURACY=1
for cfg in "1 1" "1 2" "1 3" "2 1"; do
  set -- $cfg
  MR_SCORE_PACE_TILES=$1 MR_SCORE_PACE_LEAD=$2 timeout 300 python bench.py --workload eval_cfg5 --steps 4 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_pl_$1_$2.json 2> gpurun_out/r2_pl_$1_$2.err
  python -c "
import json; b=json.load(open('gpurun_out/r2_pl_$1_$2.json')); print('W $1 lead $2 ms/step', round(b['ms_per_step'],1), 'kernel', round(b['roofline']['ms_per_launch'],1), 'clk', b['clocks']['sm_mhz'], b['checksum']['topk_ids'])"
done
MR_BENCH_EVAL_MODE=1 timeout 300 python bench.py --workload eval_cfg5 --steps 4 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_pl_x1.json 2> gpurun_out/r2_pl_x1.err
python -c "
import json; b=json.load(open('gpurun_out/r2_pl_x1.json')); print('tf32x1 ms/step', round(b['ms_per_step'],1), 'kernel', round(b['roofline']['ms_per_launch'],1), 'clk', b['clocks']['sm_mhz'])"
python tools/bf16_probe.py
