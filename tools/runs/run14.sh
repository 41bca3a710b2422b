set -x
python -m pytest tests/test_nccl_gpu.py tests/test_sharded_merger_gpu.py -m gpu -q > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2_pytest14.log
