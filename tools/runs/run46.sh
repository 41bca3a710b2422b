# PCB: clustered + hashed sample -- tests, timing, launch list
python -m pytest tests/test_pcb_gpu.py -m gpu -q -x > gpurun_out/r2_pytest46.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2_pytest46.log
python -m pytest tests/test_property_gpu.py -m gpu -q -x -k pcb > gpurun_out/r2_pytest46b.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2_pytest46b.log
python tools/pcb_probe.py > gpurun_out/r2_pcb_probe.log 2>&1; echo "pcb rc=$?"; tail -2 gpurun_out/r2_pcb_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_pcb.csv python tools/pcb_probe.py 1 noieee > gpurun_out/r2_ncu_pcb.log 2>&1; echo "ncu pcb rc=$?"
