# round-2 sweep 3 (GPU box): batched-affine pairwise rounds with the binary-GCD inversion
mkdir -p gpurun_out
SUBSET="msm_matches or structured or commit_and_open or prove_bytes or adversarial"
echo "== affine forced on small inputs"
SB_MSM_AFFINE_LOG2=0 SB_MSM_AFFINE_ROUNDS=2 SB_MSM_AFFINE_G1=1 SB_MSM_AFFINE_K=3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SUBSET" 2>&1 | tail -2
SB_MSM_AFFINE_LOG2=0 SB_MSM_AFFINE_ROUNDS=5 SB_MSM_AFFINE_G1=1 SB_MSM_S0=3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SUBSET" 2>&1 | tail -2
echo "== 2^16 bit-exact with 3 rounds"
SB_MSM_AFFINE_LOG2=12 SB_MSM_AFFINE_ROUNDS=3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size and 16" 2>&1 | tail -2
export SB_NO_TIMELINE=1
for V in "SB_MSM_AFFINE_ROUNDS=0" "SB_MSM_AFFINE_ROUNDS=1" "SB_MSM_AFFINE_ROUNDS=2" "SB_MSM_AFFINE_ROUNDS=3" "SB_MSM_AFFINE_ROUNDS=4" "SB_MSM_AFFINE_ROUNDS=5" \
         "SB_MSM_AFFINE_ROUNDS=3 SB_MSM_AFFINE_K=32" "SB_MSM_AFFINE_ROUNDS=3 SB_MSM_AFFINE_K=128" "SB_MSM_AFFINE_ROUNDS=4 SB_MSM_AFFINE_K=128" \
         "SB_MSM_AFFINE_ROUNDS=3 SB_MSM_AFFINE_G1=1 SB_MSM_AFFINE_K=128" "SB_MSM_AFFINE_ROUNDS=3 SB_MSM_AFFINE_G1=1 SB_MSM_AFFINE_K=256"; do
  env SB_MSM_SPLIT=1 $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-360
done
for V in "SB_MSM_AFFINE_ROUNDS=0" "SB_MSM_AFFINE_ROUNDS=2 SB_MSM_AFFINE_LOG2=17" "SB_MSM_AFFINE_ROUNDS=3 SB_MSM_AFFINE_LOG2=17" "SB_MSM_AFFINE_ROUNDS=4 SB_MSM_AFFINE_LOG2=17"; do
  env SB_MSM_SPLIT=1 $V python tests/gpu_timeline.py 17 2>&1 | grep -E "STEADY|Error|error" | cut -c1-360
done
unset SB_NO_TIMELINE
SB_MSM_SPLIT=1 SB_MSM_AFFINE_ROUNDS=3 SB_TAG=_r2c python tests/gpu_timeline.py 20 2>&1 | tail -22
