"""cute_nucleotides_b200 -- B200 (sm_100a) implementation of cute-nucleotides' 2-bit nucleotide codec.

Layout: csrc/ holds the CUDA kernels and the C ABI (built to libcute_nucleotides_cuda.so next to this
file); n_to_bits.py mirrors the reference's `n_to_bits` module; sharded.py is the multi-GPU host logic.
"""
from .n_to_bits import (  # noqa: F401
    LengthError, bits_to_n_cuda, decode_device, encode_checked_device, encode_device, generate_device,
    generate_words_device, n_to_bits_checked_cuda, n_to_bits_cuda, words_for_len,
    ENC_COUNT, ENC_LUT_EXACT, ENC_PLAIN, encode_ex_device, n_to_bits_ex_cuda, n_to_bits_lut_cuda,
    bits_to_n_batch_cuda, bits_to_n_cuda_async, n_to_bits_batch_cuda, n_to_bits_cuda_async, decode_segmented_device, encode_segmented_device, segment_word_offsets,
)
from .n_to_bits2 import (  # noqa: F401
    bits_to_n2_cuda, decode2_device, encode2_device, generate2_device, n_to_bits2_cuda, words2_for_len,
    encode2_ex_device, n_to_bits2_ex_cuda, n_to_bits2_lut_cuda, bits_to_n2_batch_cuda, n_to_bits2_batch_cuda,
)
from .packed_ops import (  # noqa: F401
    complement_cuda, complement_device, hamming_cuda, hamming_device, reverse_complement_cuda, reverse_complement_device,
)
from .multi import (  # noqa: F401
    decode_sharded_devices, enable_peer_access, encode_gather_to_root, encode_sharded_devices, get_devices,
    scatter_decode_from_root, set_devices, shard_bounds_c, synchronize_devices,
)

__version__ = "0.1.0"
