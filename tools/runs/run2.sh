set -x
python -m pytest tests/test_nccl_gpu.py tests/test_pcb_gpu.py "tests/test_fullsize_gpu.py::test_recformer_large_config4_properties" -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2_pytest2.log
