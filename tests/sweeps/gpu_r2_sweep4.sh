# round-2 sweep 4 (GPU box): software-pipelined MSM groups (tails on high-priority streams), G1 parts, restructured affine rounds
mkdir -p gpurun_out
SB_MSM_AFFINE_LOG2=0 SB_MSM_AFFINE_ROUNDS=2 SB_MSM_AFFINE_G1=1 SB_MSM_AFFINE_K=3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_matches or structured or commit_and_open or prove_bytes or adversarial" 2>&1 | tail -2
SB_MSM_AFFINE_LOG2=12 SB_MSM_AFFINE_ROUNDS=4 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not 20]" 2>&1 | tail -2
export SB_NO_TIMELINE=1
A="SB_MSM_AFFINE_ROUNDS=4 SB_MSM_AFFINE_K=128"
for V in "SB_MSM_SPLIT=1 SB_MSM_G1_PARTS=1" "SB_MSM_SPLIT=2 SB_MSM_G1_PARTS=2" "SB_MSM_SPLIT=3 SB_MSM_G1_PARTS=2" "SB_MSM_SPLIT=4 SB_MSM_G1_PARTS=4" \
         "SB_MSM_SPLIT=1 SB_MSM_G1_PARTS=1 $A" "SB_MSM_SPLIT=2 SB_MSM_G1_PARTS=2 $A" "SB_MSM_SPLIT=3 SB_MSM_G1_PARTS=2 $A" \
         "SB_MSM_SPLIT=2 SB_MSM_G1_PARTS=2 SB_MSM_AFFINE_ROUNDS=3 SB_MSM_AFFINE_K=256" "SB_MSM_SPLIT=2 SB_MSM_G1_PARTS=2 SB_MSM_AFFINE_ROUNDS=5 SB_MSM_AFFINE_K=256" \
         "SB_MSM_SPLIT=2 SB_MSM_G1_PARTS=2 SB_MSM_AFFINE_ROUNDS=4 SB_MSM_AFFINE_K=256 SB_MSM_AFFINE_G1=1"; do
  env $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-400
done
for V in "SB_MSM_SPLIT=1 SB_MSM_G1_PARTS=1" "SB_MSM_SPLIT=2 SB_MSM_G1_PARTS=2" "SB_MSM_SPLIT=3 SB_MSM_G1_PARTS=2" "SB_MSM_SPLIT=2 SB_MSM_G1_PARTS=2 SB_MSM_AFFINE_ROUNDS=2 SB_MSM_AFFINE_LOG2=17" "SB_MSM_SPLIT=2 SB_MSM_G1_PARTS=2 SB_MSM_AFFINE_ROUNDS=3 SB_MSM_AFFINE_LOG2=17"; do
  env $V python tests/gpu_timeline.py 17 2>&1 | grep -E "STEADY|Error|error" | cut -c1-400
done
unset SB_NO_TIMELINE
SB_MSM_SPLIT=2 SB_MSM_G1_PARTS=2 $A SB_TAG=_r2d python tests/gpu_timeline.py 20 2>&1 | tail -30
export SB_NO_TIMELINE=1
NCU="ncu --set full --clock-control none --import-source on"
env SB_MSM_SPLIT=1 SB_MSM_G1_PARTS=1 $A $NCU -k regex:k_affine_round -s 8 -c 1 -f -o gpurun_out/r02_ncu_affine_round python tests/gpu_timeline.py 20 > gpurun_out/ncu_affine.log 2>&1; tail -2 gpurun_out/ncu_affine.log
env SB_MSM_SPLIT=1 SB_MSM_G1_PARTS=1 $NCU -k regex:k_seg_accum -s 24 -c 1 -f -o gpurun_out/r02_ncu_seg_accum python tests/gpu_timeline.py 20 > gpurun_out/ncu_accum.log 2>&1; tail -2 gpurun_out/ncu_accum.log
env SB_MSM_SPLIT=1 SB_MSM_G1_PARTS=1 $NCU -k regex:k_bucket_reduce1 -s 7 -c 1 -f -o gpurun_out/r02_ncu_bucket_reduce1 python tests/gpu_timeline.py 20 > gpurun_out/ncu_red.log 2>&1; tail -2 gpurun_out/ncu_red.log
