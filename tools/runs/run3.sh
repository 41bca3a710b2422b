set -x
python -m pytest tests/test_ties_gpu.py tests/test_fullsize_gpu.py tests/test_property_gpu.py tests/test_merge_gpu.py tests/test_distill_gpu.py tests/test_pcb_gpu.py -m gpu -q --maxfail=20 > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"
python bench.py --workload ties_cfg2 --steps 20 --warmup 5 > gpurun_out/r2_bench_ties_a.json 2> gpurun_out/r2_bench_ties_a.err; echo "rc=$?"
MR_BENCH_TIES_TWO_PASS=1 python bench.py --workload ties_cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ties_b.json 2> gpurun_out/r2_bench_ties_b.err; echo "rc=$?"
tail -5 gpurun_out/r2_pytest3.log
