export MR_BENCH_SKIP_ACCURACY=1
for cfg in "37 2" "18 4" "9 8"; do
  set -- $cfg
  MR_SCORE_QGROUP=$1 MR_SCORE_SPLITS=$2 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:score_topk_kernel -s 1 -c 1 --csv --log-file gpurun_out/r2_dram_$1_$2.csv python bench.py --workload eval_cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_dram_$1_$2.log 2>&1
  echo "$cfg rc=$?"; grep -h "score_topk" gpurun_out/r2_dram_$1_$2.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
