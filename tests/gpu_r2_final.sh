# round-2 final (GPU box): full GPU test suite, the default bench line (with one real full-size CPU proof), launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu_final.log
python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -c 600 gpurun_out/r02_bench_n1_final.json; tail -3 gpurun_out/r02_bench_n1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
python tests/gpu_sc_kernels.py 2>&1 | tail -12
