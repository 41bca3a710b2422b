set -x
python -m pytest tests/test_ties_gpu.py -m gpu -q --maxfail=5 > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?"
for q in 131072 524288 1048576 2097152 4194304; do
MR_TIES_SPEC_SAMPLE_QUADS=$q python bench.py --workload ties_cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ties_q$q.json 2> gpurun_out/r2_bench_ties_q$q.err; echo "rc=$?"
done
tail -3 gpurun_out/r2_pytest8.log
export MR_BENCH_NO_GRAPH=1
python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_plain_ties8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ties_spec -s 2 -c 1 -o gpurun_out/r2_prof_spec2 python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_spec2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_ties4.csv python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_l4.log 2>&1
