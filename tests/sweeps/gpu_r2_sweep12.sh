# round-2 sweep 12 (GPU box): register-lean second pass of the affine rounds (A/B)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not knobs" 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "16]" 2>&1 | tail -3
SB_TAG=_r2k python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|affine|accum_mixed"
SB_MSM_LEAN=0 SB_TAG=_r2k0 python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY"
export SB_NO_TIMELINE=1
for V in "SB_MSM_AFFINE_ROUNDS=5" "SB_MSM_AFFINE_CTAS=4" "SB_MSM_AFFINE_CTAS=5" "SB_MSM_AFFINE_LOG2=19 SB_MSM_AFFINE_ROUNDS=3"; do
  env $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
for V in "SB_X=1" "SB_MSM_AFFINE_LOG2=19 SB_MSM_AFFINE_ROUNDS=2" "SB_MSM_AFFINE_LOG2=19 SB_MSM_AFFINE_ROUNDS=3"; do
  env $V python tests/gpu_timeline.py 17 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_affine_round -s 20 -c 1 -f -o gpurun_out/r02_ncu_affine_round_v5_lean python tests/gpu_timeline.py 20 > gpurun_out/ncu_a.log 2>&1; tail -2 gpurun_out/ncu_a.log
