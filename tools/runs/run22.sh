set -x
python -m pytest tests/test_eval_gpu.py -m gpu -q --maxfail=5 > gpurun_out/r2_pytest22.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest22.log
MR_SCORE_BF16_EPI_GROUPS=1 python tools/bf16_probe.py
MR_SCORE_BF16_EPI_GROUPS=2 python tools/bf16_probe.py
