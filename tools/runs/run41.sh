# pacing give-up rule: parity tests + config 4 / config 5 timing with and without it
export MR_BENCH_SKIP_ACCURACY=1
python -m pytest tests/test_eval_gpu.py -m gpu -q -x > gpurun_out/r2_pytest41.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest41.log
for g in 8 0; do
for w in cfg4 eval_cfg5; do
  MR_SCORE_PACE_GIVEUP=$g timeout 300 python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_gu_${w}_$g.json 2> gpurun_out/r2_gu_${w}_$g.err
  python -c "
import json; b=json.load(open('gpurun_out/r2_gu_${w}_$g.json')); print('$w giveup $g ms/step', round(b['ms_per_step'],2), 'kernel', round(b['roofline']['ms_per_launch'],2), 'clk', b['clocks']['sm_mhz'], b['checksum']['topk_ids'])"
done; done
