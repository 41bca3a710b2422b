set -x
python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench default rc=$?"
python bench.py --workload cfg1 --steps 20 --warmup 3 > gpurun_out/r2_bench_cfg1.json 2> gpurun_out/r2_bench_cfg1.err; echo "cfg1 rc=$?"
python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; echo "cfg4 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
nproc; free -g | head -2
