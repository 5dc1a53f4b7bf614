"""dumphfdl_b200 -- B200 (sm_100a) front-end for dumphfdl's multichannel HFDL hot path.

The product is the C-ABI shared library ``dumphfdl_b200/libhfdl_b200.so`` (include/hfdl_b200.h), built by
``__graft_entry__.build()`` with nvcc.  This package is only the thin ctypes view of that ABI used by the tests
and bench.py; it contains no signal processing and there is no CPU fallback: ``load()`` raises if the
CUDA library is missing."""
from .api import Frontend, Geometry, Pdu, load, bind, fft_forward, fec_decode, viterbi27, pdu_len, LIB_PATH  # noqa: F401
