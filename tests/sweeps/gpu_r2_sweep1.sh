# round-2 sweep 1 (run on the GPU box): parity of the batched MSM pipeline, then timing under its knobs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12
export SB_NO_TIMELINE=1
for V in "SB_X=default" "SB_MSM_SPLIT=1" "SB_MSM_SPLIT=3" "SB_MSM_SPLIT=4" "SB_MSM_S0_BIG=32" "SB_MSM_S0_BIG=64" "SB_MSM_LEVELS=4" "SB_MSM_SORTED=0"; do
  env $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-400
done
for V in "SB_X=default" "SB_MSM_SPLIT=1" "SB_MSM_SPLIT=3"; do
  env $V python tests/gpu_timeline.py 17 2>&1 | grep -E "STEADY|Error|error" | cut -c1-400
done
unset SB_NO_TIMELINE
SB_TAG=_r2a python tests/gpu_timeline.py 20 2>&1 | tail -40
