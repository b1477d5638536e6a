"""adafocus_b200 -- B200-native (sm_100a) implementation of the AdaFocus offline-inference hot path.

Host side: Python mirror of the reference's models/ interface (adafocus_b200.models) over a C-ABI CUDA library
(include/adafocus_b200.h, built by adafocus_b200.build).  There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
