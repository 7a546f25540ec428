# round-2 sweep 9 (GPU box): bench line with the new roofline object, window / affine knobs at small sizes, ncu captures
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; tail -c 1500 gpurun_out/r02_bench_n1_b.json; tail -3 gpurun_out/r02_bench_n1_b.err
export SB_NO_TIMELINE=1
for V in "SB_MSM_C_OFFSET=-2" "SB_MSM_C_OFFSET=-1" "SB_MSM_C_OFFSET=0"; do
  env $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
for V in "SB_X=1" "SB_MSM_C_OFFSET=0" "SB_MSM_C_OFFSET=0 SB_MSM_AFFINE_LOG2=19 SB_MSM_AFFINE_ROUNDS=2" "SB_MSM_C_OFFSET=0 SB_MSM_AFFINE_LOG2=19 SB_MSM_AFFINE_ROUNDS=3" "SB_MSM_AFFINE_LOG2=19 SB_MSM_AFFINE_ROUNDS=3" "SB_MSM_C_OFFSET=0 SB_MSM_AFFINE_LOG2=19 SB_MSM_AFFINE_ROUNDS=4"; do
  env $V python tests/gpu_timeline.py 17 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_affine_round -s 20 -c 1 -f -o gpurun_out/r02_ncu_affine_round_v3 python tests/gpu_timeline.py 20 > gpurun_out/ncu_a.log 2>&1; tail -2 gpurun_out/ncu_a.log
$NCU -k regex:k_bucket_reduce1 -s 6 -c 1 -f -o gpurun_out/r02_ncu_bucket_reduce1_quad python tests/gpu_timeline.py 20 > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
