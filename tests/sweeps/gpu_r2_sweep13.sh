# round-2 sweep 13 (GPU box): L1-bypassing loads in the affine rounds x register-lean second pass; 64-bit host field products
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not knobs" 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "16]" 2>&1 | tail -3
export SB_NO_TIMELINE=1
for V in "SB_MSM_LEAN=0 SB_MSM_STREAMING=1" "SB_MSM_LEAN=1 SB_MSM_STREAMING=1" "SB_MSM_LEAN=0 SB_MSM_STREAMING=0" "SB_MSM_LEAN=1 SB_MSM_STREAMING=0"; do
  env $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
for V in "SB_MSM_LEAN=0 SB_MSM_STREAMING=1" "SB_MSM_LEAN=1 SB_MSM_STREAMING=0"; do
  env $V python tests/gpu_timeline.py 17 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
unset SB_NO_TIMELINE
SB_MSM_LEAN=0 SB_MSM_STREAMING=1 SB_TAG=_r2l python tests/gpu_timeline.py 20 2>&1 | grep -E "affine|accum_mixed|reduce"
NCU="ncu --set full --clock-control none --import-source on"
SB_MSM_LEAN=0 SB_MSM_STREAMING=1 $NCU -k regex:k_affine_round -s 20 -c 1 -f -o gpurun_out/r02_ncu_affine_round_v6_stream python tests/gpu_timeline.py 20 > gpurun_out/ncu_a.log 2>&1; tail -2 gpurun_out/ncu_a.log
