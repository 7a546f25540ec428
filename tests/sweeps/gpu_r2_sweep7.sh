# round-2 sweep 7 (GPU box): quad-cooperative bucket reduction, sumcheck CTA sizes, window-size sweep
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not knobs" 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "16]" 2>&1 | tail -3
python tests/gpu_sc_kernels.py 2>&1 | tail -12
SB_TAG=_r2g python tests/gpu_timeline.py 20 2>&1 | tail -42
SB_TAG=_r2g python tests/gpu_timeline.py 17 2>&1 | tail -22
export SB_NO_TIMELINE=1
for V in "SB_MSM_C_MAX=17" "SB_MSM_C_MAX=18" "SB_MSM_C_MAX=19" "SB_MSM_C_MAX=18 SB_MSM_C_OFFSET=-2" "SB_MSM_C_MAX=18 SB_MSM_C_OFFSET=-1" "SB_MSM_C_MAX=19 SB_MSM_C_OFFSET=0" "SB_MSM_RED_L=2" "SB_MSM_RED_L=8"; do
  env $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
for V in "SB_MSM_C_OFFSET=-2" "SB_MSM_C_OFFSET=-1" "SB_MSM_C_OFFSET=0" "SB_MSM_C_OFFSET=1" "SB_MSM_C_OFFSET=0 SB_MSM_S0=16" "SB_MSM_C_OFFSET=0 SB_MSM_S0=32" "SB_MSM_S0=32"; do
  env $V python tests/gpu_timeline.py 17 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
