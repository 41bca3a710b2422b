# 1-GPU validation + bench lines + profiles (round 2)
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_final.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err; echo "default rc=$?"
python bench.py --workload ties_cfg2 --steps 20 --warmup 5 > gpurun_out/r02_bench_ties_cfg2.json 2> gpurun_out/r02_bench_ties_cfg2.err; echo "ties rc=$?"
python bench.py --workload cfg1 --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg1.json 2> gpurun_out/r02_bench_cfg1.err; echo "cfg1 rc=$?"
python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/r02_bench_cfg4.err; echo "cfg4 rc=$?"
python bench.py --workload collab_cfg3 --steps 10 --warmup 3 > gpurun_out/r02_bench_collab_cfg3.json 2> gpurun_out/r02_bench_collab_cfg3.err; echo "collab rc=$?"
python bench.py --workload distill_step --steps 20 --warmup 5 > gpurun_out/r02_bench_distill_step.json 2> gpurun_out/r02_bench_distill_step.err; echo "distill rc=$?"
export MR_BENCH_NO_GRAPH=1
python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_plain_ties_final.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ties_spec -s 2 -c 1 -o gpurun_out/r2_prof_spec_final python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_spec_final.log 2>&1
echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r2_launches_ties_final.csv python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_lf.log 2>&1
echo "ncu2 rc=$?"
