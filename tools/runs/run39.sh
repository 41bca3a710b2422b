# 2-GPU validation of the paced scoring kernel: NCCL tests + the default bench line
set -x
python -m pytest tests/test_nccl_gpu.py -m gpu -q > gpurun_out/r2_pytest_nccl2.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r2_pytest_nccl2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_default_n2.json 2> gpurun_out/r02_bench_default_n2.err; echo "bench n2 rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/r02_bench_default_n2.json')); print('n2 ms/step', round(b['ms_per_step'],1), 'e2e', round(b['e2e']['ms_per_step'],1), b['checksum']['topk_ids'], b['stage_ms']); m=b['merger']; print('merger', round(m['ms_per_step'],3), m.get('checksum'))"
