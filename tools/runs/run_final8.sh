set -x
python -m pytest tests/test_nccl_gpu.py -m gpu -q > gpurun_out/r2_pytest_nccl8.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r2_pytest_nccl8.log
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_default_n$n.json 2> gpurun_out/r02_bench_default_n$n.err; echo "bench n$n rc=$?"
done
