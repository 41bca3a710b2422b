# pacing lead sweep at config 5 + bf16-compat mode with / without pacing
export MR_BENCH_SKIP_ACCURACY=1
for l in 1 3; do
  MR_SCORE_PACE_LEAD=$l timeout 300 python bench.py --workload eval_cfg5 --steps 4 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_lead_$l.json 2> gpurun_out/r2_lead_$l.err
  python -c "
import json; b=json.load(open('gpurun_out/r2_lead_$l.json')); print('lead $l ms/step', round(b['ms_per_step'],1), 'clk', b['clocks']['sm_mhz'], b['checksum']['topk_ids'])"
done
MR_SCORE_PACE_LEAD=2 timeout 300 python bench.py --workload eval_cfg5 --steps 4 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_lead_2.json 2> gpurun_out/r2_lead_2.err
python -c "
import json; b=json.load(open('gpurun_out/r2_lead_2.json')); print('lead 2 ms/step', round(b['ms_per_step'],1), 'clk', b['clocks']['sm_mhz'], b['checksum']['topk_ids'])"
python tools/bf16_probe.py 2>&1 | tail -1
MR_SCORE_PACE_TILES=0 python tools/bf16_probe.py 2>&1 | tail -1
