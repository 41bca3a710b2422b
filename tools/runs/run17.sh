set -x
python -m pytest tests/test_eval_gpu.py tests/test_lambda_gpu.py tests/test_joint_eval.py tests/test_loader_gpu.py -m gpu -q --maxfail=10 > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest17.log
python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg4b.json 2> gpurun_out/r2_bench_cfg4b.err; echo "cfg4 rc=$?"
export MR_BENCH_SKIP_ACCURACY=1
python bench.py --workload eval_cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_plain_eval.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_topk_kernel -s 1 -c 1 -o gpurun_out/r2_prof_score_cfg5 python bench.py --workload eval_cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_ncu_score.log 2>&1
echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_eval.csv python bench.py --workload eval_cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_ncu_le.log 2>&1
echo "ncu2 rc=$?"
