"""cute_nucleotides_b200 -- B200 (sm_100a) implementation of cute-nucleotides' 2-bit nucleotide codec.

Layout: csrc/ holds the CUDA kernels and the C ABI (built to libcute_nucleotides_cuda.so next to this
file); n_to_bits.py mirrors the reference's `n_to_bits` module; sharded.py is the multi-GPU host logic.
"""
from .n_to_bits import (  # noqa: F401
    LengthError, bits_to_n_cuda, decode_device, encode_checked_device, encode_device, generate_device,
    generate_words_device, n_to_bits_checked_cuda, n_to_bits_cuda, words_for_len,
)
from .n_to_bits2 import (  # noqa: F401
    bits_to_n2_cuda, decode2_device, encode2_device, generate2_device, n_to_bits2_cuda, words2_for_len,
)

__version__ = "0.1.0"
