set -x
python tools/ties_fused_probe.py > gpurun_out/r2_plain_f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ties_spec -s 3 -c 1 -o gpurun_out/r2_prof_specf python tools/ties_fused_probe.py > gpurun_out/r2_ncu_specf.log 2>&1
