set -x
export MR_BENCH_NO_GRAPH=1
python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_plain_ties.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ties_spec -s 2 -c 1 -o gpurun_out/r2_prof_spec python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_spec.log 2>&1
echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_ties.csv python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_l.log 2>&1
echo "ncu2 rc=$?"
