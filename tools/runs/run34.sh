# 1-GPU validation of HEAD (pacing) + fresh bench lines + PCB / fused-TIES launch lists
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_final.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err; echo "default rc=$?"
python tools/pcb_probe.py > gpurun_out/r2_pcb_probe.log 2>&1; echo "pcb rc=$?"; tail -3 gpurun_out/r2_pcb_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_pcb.csv python tools/pcb_probe.py 1 > gpurun_out/r2_ncu_pcb.log 2>&1; echo "ncu pcb rc=$?"
python tools/ties_fused_probe.py > gpurun_out/r2_fused_probe.log 2>&1; tail -2 gpurun_out/r2_fused_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_fused.csv python tools/ties_fused_probe.py > gpurun_out/r2_ncu_fused.log 2>&1; echo "ncu fused rc=$?"
