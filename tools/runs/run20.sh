set -x
python -m pytest tests/test_eval_gpu.py tests/test_nccl_gpu.py -m gpu -q --maxfail=10 > gpurun_out/r2_pytest20.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest20.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default2.json 2> gpurun_out/r2_bench_default2.err; echo "bench rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2c.json 2> gpurun_out/r2_bench_n2c.err; echo "bench n2 rc=$?"
