# last check of the shipped library: smoke + evaluator parity tests + one config-5 line
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests/test_eval_gpu.py -m gpu -q -x > gpurun_out/r2_pytest45.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2_pytest45.log
MR_BENCH_SKIP_ACCURACY=1 timeout 300 python bench.py --workload eval_cfg5 --steps 6 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_last_cfg5.json 2> gpurun_out/r2_last_cfg5.err
python -c "
import json; b=json.load(open('gpurun_out/r2_last_cfg5.json')); print('cfg5 ms/step', round(b['ms_per_step'],1), 'clk', b['clocks']['sm_mhz'], b['checksum']['topk_ids'], 'frac', round(b['roofline']['frac'],3))"
