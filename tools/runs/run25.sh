set -x
python -m pytest tests/test_eval_gpu.py tests/test_property_gpu.py tests/test_joint_eval.py -m gpu -q --maxfail=5 > gpurun_out/r2_pytest25.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest25.log
python tools/merge_probe.py
