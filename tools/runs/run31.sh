export MR_BENCH_SKIP_ACCURACY=1
python -m pytest tests/test_eval_gpu.py -m gpu -q -x > gpurun_out/r2_pytest31.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest31.log
for w in 0 4 8 16; do
  MR_SCORE_PACE_TILES=$w timeout 300 python bench.py --workload eval_cfg5 --steps 4 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_pace_$w.json 2> gpurun_out/r2_pace_$w.err
  echo "pace $w rc=$?"
  python -c "
import json; b=json.load(open('gpurun_out/r2_pace_$w.json')); print('pace $w ms/step', round(b['ms_per_step'],1), 'kernel', round(b['roofline']['ms_per_launch'],1), 'clk', b['clocks']['sm_mhz'], b['checksum']['topk_ids'])"
done
MR_SCORE_PACE_TILES=8 timeout 300 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:score_topk_kernel -s 1 -c 1 --csv --log-file gpurun_out/r2_dram_pace8.csv python bench.py --workload eval_cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-companion > gpurun_out/r2_dram_pace8.log 2>&1
grep -h "score_topk" gpurun_out/r2_dram_pace8.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
