# round-2 sweep 8 (GPU box): device-side indexer, prefetch in the affine rounds (A/B), occupancy-sized G1 rounds, reduction knobs
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not knobs" 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "16]" 2>&1 | tail -3
python tests/gpu_scale.py 20 noverify 2>&1 | grep -E "index|keygen|prove" | head -4
SB_TAG=_r2h python tests/gpu_timeline.py 20 2>&1 | tail -42
export SB_NO_TIMELINE=1
for V in "SB_MSM_PREFETCH=0" "SB_MSM_AFFINE_CTAS=4" "SB_MSM_PREFETCH=0 SB_MSM_AFFINE_CTAS=4" "SB_MSM_S0_SMALL=16" "SB_MSM_S0_SMALL=12" "SB_MSM_AFFINE_ROUNDS=5"; do
  env $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
unset SB_NO_TIMELINE
SB_MSM_RED_L=8 SB_TAG=_r2h_l8 python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|reduce"
SB_MSM_C_OFFSET=0 SB_TAG=_r2h_c0 python tests/gpu_timeline.py 17 2>&1 | tail -22
