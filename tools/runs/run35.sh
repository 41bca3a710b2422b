# PCB rework: tests + timing + launch list
set -x
python -m pytest tests/test_pcb_gpu.py tests/test_cabi.py -m gpu -q -x > gpurun_out/r2_pytest35.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest35.log
python -m pytest tests/test_property_gpu.py -m gpu -q -x -k pcb > gpurun_out/r2_pytest35b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest35b.log
python tools/pcb_probe.py > gpurun_out/r2_pcb_probe.log 2>&1; echo "pcb rc=$?"; tail -3 gpurun_out/r2_pcb_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_pcb.csv python tools/pcb_probe.py 1 > gpurun_out/r2_ncu_pcb.log 2>&1; echo "ncu pcb rc=$?"
