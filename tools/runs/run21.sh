set -x
python -m pytest tests/test_ties_gpu.py tests/test_property_gpu.py tests/test_sharded_merger_gpu.py -m gpu -q --maxfail=5 > gpurun_out/r2_pytest21.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest21.log
python bench.py --workload ties_cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ties_e.json 2> gpurun_out/r2_bench_ties_e.err; echo "rc=$?"
MR_ONE_PASS=1 python tools/ties_fused_probe.py
export MR_BENCH_NO_GRAPH=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r2_launches_ties7.csv python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_l7.log 2>&1
