"""quasimodo_b200 -- B200-native implementation of QuasiModo's read-level hot path
(align -> pile up -> classify SNPs).  All computation lives in libquasimodo_b200.so (hand-written
sm_100a CUDA behind a C-ABI, include/quasimodo_b200.h); this package is the host-side mirror."""
from ._lib import QmError, SO_PATH  # noqa: F401
from .api import Context  # noqa: F401

__all__ = ["Context", "QmError", "SO_PATH"]
