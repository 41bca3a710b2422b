# sanity of the clean rebuild: smoke + the test files of the kernels touched last
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests/test_eval_gpu.py tests/test_pcb_gpu.py tests/test_cabi.py tests/test_schedule.py -q -x > gpurun_out/r2_pytest47.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2_pytest47.log
