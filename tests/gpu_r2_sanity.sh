# round-2 last sanity (GPU box): the final tree's GPU tests (without the two full-size cases) and a short bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu_sanity.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_sanity.json 2> gpurun_out/r02_bench_n1_sanity.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1_sanity.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'sha', d['proof_sha256'], 'roofline', d['roofline']['frac'])" || tail -3 gpurun_out/r02_bench_n1_sanity.err
