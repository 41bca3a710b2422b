# final 1-GPU validation of the round-2 build: smoke, all GPU tests, the default bench line and the reference arm
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_final.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err; echo "default rc=$?"
python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/r02_bench_cfg4.err; echo "cfg4 rc=$?"
