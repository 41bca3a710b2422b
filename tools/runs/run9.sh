set -x
python -m pytest tests/test_ties_gpu.py -m gpu -q --maxfail=5 > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?"
for q in 524288 1048576 2097152 4194304; do
MR_TIES_SPEC_SAMPLE_QUADS=$q python bench.py --workload ties_cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ties_q$q.json 2> gpurun_out/r2_bench_ties_q$q.err; echo "rc=$?"
done
tail -3 gpurun_out/r2_pytest9.log
export MR_BENCH_NO_GRAPH=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/r2_launches_ties5.csv python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_l5.log 2>&1
