"""spades_for_blackbird_b200 — B200-native (sm_100a) graph-construction front end of SPAdes 3.15.4.

Layout:
  csrc/   CUDA kernels + the C-ABI shared library (libspades_b200.so); include/sb200.h is the contract
  host/   host-side mirror of the reference's interfaces for this path (ctypes binding, synthetic read generator)
"""
__version__ = "0.1.0"
