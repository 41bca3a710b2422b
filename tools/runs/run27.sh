set -x
python -m pytest tests/test_ties_gpu.py tests/test_fullsize_gpu.py tests/test_pcb_gpu.py tests/test_sharded_merger_gpu.py tests/test_merge_gpu.py -m gpu -q --maxfail=5 > gpurun_out/r2_pytest27.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest27.log
python bench.py --workload ties_cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ties_g.json 2> gpurun_out/r2_bench_ties_g.err; echo "rc=$?"
MR_HYPOTHESIS_EXAMPLES=300 timeout 1200 python -m pytest tests/test_property_gpu.py -m gpu -q -x > gpurun_out/r2_soak.log 2>&1; echo "soak rc=$?"
tail -4 gpurun_out/r2_soak.log
