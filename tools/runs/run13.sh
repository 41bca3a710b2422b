set -x
python -m pytest tests/test_sharded_merger_gpu.py tests/test_ties_gpu.py -m gpu -q --maxfail=8 > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"
python bench.py --workload ties_sharded --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_sh1.json 2> gpurun_out/r2_bench_sh1.err; echo "rc=$?"
tail -3 gpurun_out/r2_pytest13.log
