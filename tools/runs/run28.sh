set -x
python tools/sanitize_probe.py > gpurun_out/r2_sanitize_plain.log 2>&1 && timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_probe.py > gpurun_out/r2_sanitize.log 2>&1
echo "sanitizer rc=$?"
tail -6 gpurun_out/r2_sanitize.log
