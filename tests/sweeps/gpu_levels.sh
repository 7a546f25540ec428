# per-level serialised kernel durations and steady prove time for a few chunk sizes (run on the GPU box)
mkdir -p gpurun_out
for B in 16 20 24 28 48; do
  SB_MSM_S0_BIG=$B python tests/gpu_levels.py 20 > gpurun_out/levels_b$B.log 2>&1
  SB_NO_TIMELINE=1 SB_MSM_S0_BIG=$B python tests/gpu_timeline.py 20 > gpurun_out/steady_b$B.log 2>&1
done
for B in 16 20 24 28 48; do echo "== $B"; grep -h "^commit\|^2^19\|^2^18\|^2^17\|^2^16\|^SUM" gpurun_out/levels_b$B.log; grep -h STEADY gpurun_out/steady_b$B.log | cut -c1-120; done
