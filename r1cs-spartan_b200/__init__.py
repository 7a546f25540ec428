"""r1cs-spartan_b200 -- B200-native (sm_100a) implementation of the r1cs-spartan prover hot path.

The directory name follows the project naming; Python code imports it as `r1cs_spartan_b200`
(the shim package next to it).  Only what the hot path needs lives here:

    csrc/        CUDA kernels, the host driver and the extern "C" surface (include/spartan_b200.h)
    api.py       ctypes mirror of the reference's prover-facing API
    workload.py  the reference's synthetic benchmark circuit
    dist.py      one-process-per-GPU plumbing (torch.distributed) for the sharded prover
    wire.py      arkworks CanonicalSerialize layouts of Proof / IndexPK / PublicParameter (exchange with a Rust build)
"""
from .api import (  # noqa: F401
    Context, CudaError, InvalidArgument, IndexPK, MLArgumentForR1CS, MLPolyCommit, MLProofForR1CS, PublicParameter,
    ProveTrace, Witness, default_context, device_count, eq_extension, load_library, multi_scalar_mul, LIB_PATH, EXPORTS,
)
from .workload import SyntheticR1CS  # noqa: F401
