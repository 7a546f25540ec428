# Round-2 starting point (run on the GPU box after `make -C r1cs-spartan_b200 lazy` here): parity and timing of the two
# experiments prepared at the end of round 1 and never run on a GPU --
#   (1) LAZY_Y3 point additions (libspartan_b200_lazy.so, selected with SB_LIB_PATH)
#   (2) cooperative batched-affine rounds (SB_MSM_AFFINE_COOP=1) for MSMs of >= 2^SB_MSM_AFFINE_LOG2 entries
# Each line prints the pytest tail or the STEADY prove time; nothing here changes a default.
mkdir -p gpurun_out
SUBSET="msm_matches or structured or commit_and_open or prove_bytes or adversarial or (full_size and 16)"
export SB_NO_TIMELINE=1
echo "== baseline";            python tests/gpu_timeline.py 20 2>&1 | grep STEADY | cut -c1-200
if [ -f r1cs-spartan_b200/libspartan_b200_lazy.so ]; then
  echo "== lazy parity";       SB_LIB_PATH=$PWD/r1cs-spartan_b200/libspartan_b200_lazy.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SUBSET" 2>&1 | tail -2
  echo "== lazy timing";       SB_LIB_PATH=$PWD/r1cs-spartan_b200/libspartan_b200_lazy.so python tests/gpu_timeline.py 20 2>&1 | grep STEADY | cut -c1-200
fi
echo "== coop affine parity (forced on small inputs)"
SB_MSM_AFFINE_LOG2=4 SB_MSM_AFFINE_COOP=1 SB_MSM_AFFINE_K=3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SUBSET" 2>&1 | tail -2
SB_MSM_AFFINE_LOG2=4 SB_MSM_AFFINE_COOP=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SUBSET" 2>&1 | tail -2
for K in 32 64 128; do
  echo "== coop affine timing K=$K (levels with >= 2^21 entries)"
  SB_MSM_AFFINE_LOG2=21 SB_MSM_AFFINE_COOP=1 SB_MSM_AFFINE_K=$K python tests/gpu_timeline.py 20 2>&1 | grep STEADY | cut -c1-200
done
# (3) sumcheck round kernels alone: more, finer CTAs for the large rounds (north-star target: >= 60 % of the IMAD ceiling at 2^20)
for C in 2 4 8 16; do echo "== SB_SC_CTAS_PER_SM=$C"; SB_SC_CTAS_PER_SM=$C python tests/gpu_sc_kernels.py 2>&1 | grep -E "2\^20|2\^22"; done
# (4) one more window bit (fewer mixed additions, twice the buckets) and the first-level chunk size, now that the
#     accumulation order is length-sorted
for V in "SB_MSM_C_OFFSET=-2 SB_MSM_C_MAX=17" "SB_MSM_S0_BIG=32" "SB_MSM_S0_BIG=64" "SB_MSM_S0_BIG=96"; do
  echo "== $V"; env $V python tests/gpu_timeline.py 20 2>&1 | grep STEADY | cut -c1-200
done
