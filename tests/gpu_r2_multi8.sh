# round-2 (GPU box, --gpus 8): sharded tests at 4 and 8 GPUs, bench at N = 8 and 4, configs C4 and C5 on one multi-GPU context
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "8gpu or 4gpu" 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_multi_n8.log
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N value', d['value'], 'e2e', d['e2e']['value'], 'sha', d['proof_sha256'], d['sharded_check']); print(d['step_phases_ms'][-1])" || tail -5 gpurun_out/r02_bench_n$N.err
done
timeout 900 python tests/gpu_multi8.py 8 c4 prove20 c5 2>&1 | tee gpurun_out/r02_multi8_one_process.log | tail -30
