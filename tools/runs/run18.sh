export MR_BENCH_SKIP_ACCURACY=1
i=0
for cfg in "0 16 0" "1 16 0" "1 8 4" "1 8 9" "0 8 4" "1 37 2" "1 4 18" "1 16 4" "0 8 9"; do
  set -- $cfg
  i=$((i+1))
  MR_SCORE_L2HINT=$1 MR_SCORE_QGROUP=$2 MR_SCORE_SPLITS=$3 python bench.py --workload eval_cfg5 --steps 4 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_sweep_$i.json 2> gpurun_out/r2_sweep_$i.err
  echo "$cfg rc=$?"
done
