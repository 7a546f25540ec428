# scheduling / chunking sweep (run on the GPU box): steady prove time per knob setting, one process each
mkdir -p gpurun_out
run() { # log_n tag env...
  L=$1; T=$2; shift 2
  env SB_NO_TIMELINE=1 "$@" python tests/gpu_timeline.py $L > gpurun_out/sched_${L}_$T.log 2>&1
}
python tests/gpu_timeline.py 20 > gpurun_out/sched_20_default.log 2>&1
run 20 b32 SB_MSM_S0_BIG=32
run 20 b64 SB_MSM_S0_BIG=64
run 20 s1_3 SB_MSM_S1=3
run 20 t1 SB_MSM_TAILS=1
run 20 t2 SB_MSM_TAILS=2
python tests/gpu_timeline.py 17 > gpurun_out/sched_17_default.log 2>&1
run 17 sm16 SB_MSM_S0_SMALL=16
run 17 sm32 SB_MSM_S0_SMALL=32
run 17 s1_3 SB_MSM_S1=3
run 17 t1 SB_MSM_TAILS=1
grep -h STEADY gpurun_out/sched_*.log
