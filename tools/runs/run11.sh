set -x
MR_TIES_FUSED_BLOCKED=1 python tools/ties_fused_probe.py
MR_TIES_FUSED_BLOCKED=0 python tools/ties_fused_probe.py
MR_TIES_FUSED_BLOCKED=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_fused.csv python tools/ties_fused_probe.py > gpurun_out/r2_ncu_f.log 2>&1
