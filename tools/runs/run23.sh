set -x
python -m pytest tests/test_ties_gpu.py tests/test_property_gpu.py tests/test_sharded_merger_gpu.py tests/test_pcb_gpu.py "tests/test_fullsize_gpu.py" -m gpu -q --maxfail=5 > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest23.log
for q in 524288 1048576; do
MR_TIES_SPEC_SAMPLE_QUADS=$q python bench.py --workload ties_cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_ties_s$q.json 2> gpurun_out/r2_bench_ties_s$q.err; echo "rc=$?"
done
export MR_BENCH_NO_GRAPH=1
MR_TIES_SPEC_SAMPLE_QUADS=1048576 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r2_launches_ties8.csv python bench.py --workload ties_cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_l8.log 2>&1
