"""fftoptionpricing_b200 -- B200 (sm_100a) batched Fourier option pricing.

Host-side mirror of danielhstahl/FFTOptionPricing's pricing path over a hand-written CUDA
library (``libfftop_b200.so``, C ABI in ``include/fftop_b200.h``).  Importing the package loads
the shared library; it raises ImportError when it has not been built.  No CPU fallback exists.
"""
from ._lib import LIB_PATH, declared_symbols, load_library  # noqa: F401
from .engine import (  # noqa: F401
    CALL_DELTA, CALL_GAMMA, CALL_PRICE, CALL_THETA, CGMY, FSTS_DELTA, FSTS_GAMMA, FSTS_PRICE,
    FSTS_THETA, GAUSS, MERTON, PAYOFF_CALL, PAYOFF_PUT, PUT_DELTA, PUT_GAMMA, PUT_PRICE, PUT_THETA,
    TC_CIR, TC_CIR_DIFFUSION, TC_NONE, VG, Engine, FopError, num_params,
)

load_library()
