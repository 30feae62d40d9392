"""audiowmark_b200 -- B200-native (sm_100a) implementation of audiowmark's spectral
watermark hot path (embed, sync search, block decode, Viterbi) behind a C ABI.

Layout
  csrc/   CUDA kernels + C ABI (include/awm_b200.h)  -> lib/libawm_b200.so
  host/   C++ host side mirroring the reference's classes / CLI -> lib/libawm_host.so, bin/audiowmark
  capi.py ctypes binding of the C ABI (tests, bench)
"""
from . import capi  # noqa: F401
