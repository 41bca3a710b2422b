export MR_BENCH_SKIP_ACCURACY=1
i=0
for cfg in "1 37 2" "1 18 4" "1 74 2" "1 25 2" "1 32 2" "1 37 4"; do
  set -- $cfg
  i=$((i+1))
  MR_SCORE_L2HINT=$1 MR_SCORE_QGROUP=$2 MR_SCORE_SPLITS=$3 python bench.py --workload eval_cfg5 --steps 4 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_sweepb_$i.json 2> gpurun_out/r2_sweepb_$i.err
  echo "$cfg rc=$?"
done
MR_SCORE_L2HINT=1 MR_SCORE_QGROUP=37 MR_SCORE_SPLITS=2 python bench.py --workload eval_cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_plain_eval2.log 2>&1 &&
MR_SCORE_L2HINT=1 MR_SCORE_QGROUP=37 MR_SCORE_SPLITS=2 ncu --set full --clock-control none -k regex:score_topk_kernel -s 1 -c 1 -o gpurun_out/r2_prof_score_cfg5_b python bench.py --workload eval_cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_ncu_score2.log 2>&1
