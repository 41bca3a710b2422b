# N-GPU bench line of the paced build (N from $1)
set -x
n=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02_bench_default_n$n.json 2> gpurun_out/r02_bench_default_n$n.err; echo "bench n$n rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/r02_bench_default_n$n.json')); print('n$n ms/step', round(b['ms_per_step'],1), 'e2e', round(b['e2e']['ms_per_step'],1), b['checksum']['topk_ids'], b['stage_ms']); m=b['merger']; print('merger', round(m['ms_per_step'],3), m.get('checksum'))"
