# wave-aligned unit order: parity tests + config 4 / config 5 timing
export MR_BENCH_SKIP_ACCURACY=1
python -m pytest tests/test_eval_gpu.py tests/test_property_gpu.py -m gpu -q -x -k "eval or score or topk" > gpurun_out/r2_pytest42.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest42.log
for w in cfg4 eval_cfg5 cfg1; do
  timeout 300 python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_wa_${w}.json 2> gpurun_out/r2_wa_${w}.err
  python -c "
import json; b=json.load(open('gpurun_out/r2_wa_${w}.json')); print('$w ms/step', round(b['ms_per_step'],3), 'kernel', round(b['roofline']['ms_per_launch'],3), 'clk', b['clocks']['sm_mhz'], b['checksum']['topk_ids'])"
done
