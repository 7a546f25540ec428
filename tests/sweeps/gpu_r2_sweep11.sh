# round-2 sweep 11 (GPU box): cp.async staging in the affine rounds (A/B), reduction CTA shapes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not knobs" 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "16]" 2>&1 | tail -3
SB_TAG=_r2j python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|affine"
SB_MSM_STAGED=0 SB_TAG=_r2j0 python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|affine"
export SB_NO_TIMELINE=1
for V in "SB_MSM_RED_QUADS=32 SB_MSM_RED_L=8" "SB_MSM_RED_QUADS=32 SB_MSM_RED_L=4" "SB_MSM_RED_QUADS=16 SB_MSM_RED_L=8" "SB_MSM_RED_QUADS=16 SB_MSM_RED_L=16" "SB_MSM_RED_QUADS=32 SB_MSM_RED_L=16"; do
  env $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
  env $V python tests/gpu_timeline.py 17 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
python tests/gpu_timeline.py 17 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_affine_round -s 20 -c 1 -f -o gpurun_out/r02_ncu_affine_round_v4 python tests/gpu_timeline.py 20 > gpurun_out/ncu_a.log 2>&1; tail -2 gpurun_out/ncu_a.log
