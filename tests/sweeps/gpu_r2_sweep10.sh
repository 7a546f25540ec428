# round-2 sweep 10 (GPU box): quad operations on shared-memory operands, group-dependent window rule
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not knobs" 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "16]" 2>&1 | tail -3
SB_TAG=_r2i python tests/gpu_timeline.py 20 2>&1 | tail -42
SB_TAG=_r2i python tests/gpu_timeline.py 17 2>&1 | tail -24
export SB_NO_TIMELINE=1
for V in "SB_MSM_RED_L=2" "SB_MSM_RED_L=8"; do
  env $V python tests/gpu_timeline.py 20 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
  env $V python tests/gpu_timeline.py 17 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
done
python tests/gpu_timeline.py 18 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
python tests/gpu_timeline.py 16 2>&1 | grep -E "STEADY|Error|error" | cut -c1-420
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_bucket_reduce1 -s 6 -c 1 -f -o gpurun_out/r02_ncu_bucket_reduce1_quad_v2 python tests/gpu_timeline.py 20 > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log
