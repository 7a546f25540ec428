# round-2 sweep 6 (GPU box): new defaults (affine rounds R=4, K=256, G1 too), full GPU test suite, bench line, timelines, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err; tail -c 3000 gpurun_out/r02_bench_n1_a.json
SB_TAG=_r2f python tests/gpu_timeline.py 20 2>&1 | tail -40
export SB_NO_TIMELINE=1
for V in "SB_MSM_AFFINE_K=512" "SB_MSM_AFFINE_ROUNDS=3" "SB_MSM_SPLIT=3" "SB_MSM_AFFINE_LOG2=20"; do
  env $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
for V in "SB_X=0" "SB_MSM_AFFINE_LOG2=19 SB_MSM_AFFINE_ROUNDS=3"; do
  env $V python tests/gpu_timeline.py 18 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
unset SB_NO_TIMELINE
SB_TAG=_r2f python tests/gpu_timeline.py 17 2>&1 | tail -40
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_racecheck.log
