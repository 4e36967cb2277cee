"""crdr_b200: B200-native (sm_100a) implementation of the CRDR codec hot path.

Host Python mirrors the reference's operator interface; the arithmetic lives in hand-written CUDA
behind the C ABI declared in ``include/crdr_b200.h``.
"""
__version__ = "0.1.0"
