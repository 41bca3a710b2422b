# refreshed 1-GPU evidence: ncu of the paced config-5 launch, bench lines of the other workloads, PCB launch list
set -x
export MR_BENCH_SKIP_ACCURACY=1
bash tools/ncu_capture.sh r2_score_cfg5_paced score_topk_kernel 3 -- python bench.py --workload eval_cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-companion
unset MR_BENCH_SKIP_ACCURACY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_pcb.csv python tools/pcb_probe.py 1 noieee > gpurun_out/r2_ncu_pcb.log 2>&1; echo "ncu pcb rc=$?"
python bench.py --workload cfg1 --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg1.json 2> gpurun_out/r02_bench_cfg1.err; echo "cfg1 rc=$?"
python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/r02_bench_cfg4.err; echo "cfg4 rc=$?"
python bench.py --workload ties_cfg2 --steps 20 --warmup 5 > gpurun_out/r02_bench_ties_cfg2.json 2> gpurun_out/r02_bench_ties_cfg2.err; echo "ties rc=$?"
