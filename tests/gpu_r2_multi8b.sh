# round-2 (GPU box, --gpus 8): configs C4 and C5 on ONE process / ONE context over 8 GPUs; bench at N = 8 with the 64-bit host products
mkdir -p gpurun_out
timeout 900 python tests/gpu_multi8.py 8 c4 prove20 c5 2>&1 | tee gpurun_out/r02_multi8_one_process.log | tail -30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_n8b.json 2> gpurun_out/r02_bench_n8b.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n8b.json').read().strip().splitlines()[-1])
print('N=8 value', d['value'], 'e2e', d['e2e']['value'], 'sha', d['proof_sha256'], d['sharded_check']); print(d['step_phases_ms'][-1])" || tail -5 gpurun_out/r02_bench_n8b.err
