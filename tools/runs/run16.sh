set -x
python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/r2_pytest16.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest16.log
python bench.py --workload cfg1 --steps 20 --warmup 3 > gpurun_out/r2_bench_cfg1b.json 2> gpurun_out/r2_bench_cfg1b.err; echo "cfg1 rc=$?"
