set -x
python -m pytest tests/test_ties_gpu.py -m gpu -q --maxfail=5 -k "select_build or fused_select" > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?"
for m in 2 3 4; do
MR_TIES_SPEC_MINB=$m MR_BENCH_SKIP_ACCURACY=1 python bench.py --workload ties_cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ties_m$m.json 2> gpurun_out/r2_bench_ties_m$m.err; echo "rc=$?"
done
tail -3 gpurun_out/r2_pytest5.log
