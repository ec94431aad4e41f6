"""crescent_credentials_b200 -- B200-native Groth16 prover (BN254) behind the ark-groth16 API surface that
microsoft/crescent-credentials' `prove` step uses.  The compute path is libg16b200.so (hand-written CUDA for sm_100a,
see csrc/); this package is the thin host-side mirror of the reference interface on top of the C ABI in
include/g16_b200.h.  There is no CPU fallback: importing works anywhere, computing needs the library and a GPU."""
from . import ffi  # noqa: F401
from .groth16 import (CircomReduction, ConstraintMatrices, Groth16, LibsnarkReduction, Proof, ProvingKey,  # noqa: F401
                      VerifyingKey)

__all__ = ["ffi", "Groth16", "ProvingKey", "VerifyingKey", "Proof", "ConstraintMatrices", "LibsnarkReduction",
           "CircomReduction"]
