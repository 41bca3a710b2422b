set -x
MR_HYPOTHESIS_EXAMPLES=400 timeout 1500 python -m pytest tests/test_property_gpu.py -m gpu -q -x > gpurun_out/r2_soak.log 2>&1; echo "soak rc=$?"
tail -5 gpurun_out/r2_soak.log
