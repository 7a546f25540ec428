# round-2 sweep 2 (GPU box): multi-CTA plan kernels, concurrent groups, mailbox sumcheck
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not 20]" --durations=3 2>&1 | tail -6
export SB_NO_TIMELINE=1
for V in "SB_MSM_SPLIT=1" "SB_MSM_SPLIT=2" "SB_MSM_SPLIT=3" "SB_MSM_SPLIT=4" "SB_MSM_SPLIT=2 SB_MSM_S0_BIG=64 SB_MSM_S0_SMALL=32" "SB_MSM_SPLIT=2 SB_MSM_LEVELS=4"; do
  env $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
for V in "SB_MSM_SPLIT=1" "SB_MSM_SPLIT=2" "SB_MSM_SPLIT=3" "SB_MSM_SPLIT=2 SB_MSM_S0_SMALL=16" "SB_MSM_SPLIT=2 SB_MSM_S0_SMALL=32"; do
  env $V python tests/gpu_timeline.py 17 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
unset SB_NO_TIMELINE
SB_TAG=_r2b python tests/gpu_timeline.py 20 2>&1 | tail -25
for C in 2 4 8 16; do echo "== SB_SC_CTAS_PER_SM=$C"; SB_SC_CTAS_PER_SM=$C python tests/gpu_sc_kernels.py 2>&1 | grep -E "2\^20|2\^22"; done
