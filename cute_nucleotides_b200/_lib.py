"""ctypes binding of libcute_nucleotides_cuda.so (the C ABI in include/cute_nucleotides_cuda.h).

There is deliberately no fallback: if the shared library is missing, or a call fails, an exception is
raised.  Nothing in this package imports or calls the CPU oracle under oracle/.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CN_CUDA_LIB", os.path.join(_HERE, "libcute_nucleotides_cuda.so"))   # override: A/B builds only

CN_OK, CN_ERR_LENGTH, CN_ERR_CUDA, CN_ERR_ARG, CN_ERR_NOMEM = 0, 1, 2, 3, 4
CN_DIR_ENCODE, CN_DIR_DECODE = 0, 1
ABI_VERSION = 2
CN_ENC_PLAIN, CN_ENC_COUNT, CN_ENC_LUT_EXACT = 0, 1, 2

# name -> (restype, argtypes); mirrors include/cute_nucleotides_cuda.h one to one
PROTOTYPES = {
    "cn_abi_version": (c_int, []),
    "cn_init": (c_int, [c_int]),
    "cn_device_count": (c_int, [POINTER(c_int)]),
    "cn_last_error": (c_char_p, []),
    "cn_length_panic_message": (c_char_p, []),
    "cn_words_for_len": (c_size_t, [c_size_t]),
    "cn_launch_count": (c_uint64, []),
    "cn_n_to_bits_host": (c_int, [c_void_p, c_size_t, c_void_p]),
    "cn_bits_to_n_host": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p]),
    "cn_n_to_bits_host_batch": (c_int, [POINTER(c_void_p), POINTER(c_size_t), c_size_t, POINTER(c_void_p)]),
    "cn_bits_to_n_host_batch": (c_int, [POINTER(c_void_p), POINTER(c_size_t), c_size_t, POINTER(c_void_p)]),
    "cn_n_to_bits_host_async": (c_int, [c_void_p, c_size_t, c_void_p, POINTER(c_void_p)]),
    "cn_bits_to_n_host_async": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, POINTER(c_void_p)]),
    "cn_wait": (c_int, [c_void_p]),
    "cn_encode_device": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "cn_decode_device": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "cn_encode_segmented_device": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p]),
    "cn_decode_segmented_device": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p]),
    "cn_encode_multi_device": (c_int, [c_void_p, c_size_t, POINTER(c_void_p), c_int, c_void_p]),
    "cn_encode2_multi_device": (c_int, [c_void_p, c_size_t, POINTER(c_void_p), c_int, c_void_p]),
    "cn_ipc_export": (c_int, [c_void_p, c_void_p, POINTER(c_size_t)]),
    "cn_ipc_open": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "cn_ipc_close": (c_int, [c_void_p, c_size_t]),
    "cn_n_to_bits_checked_host": (c_int, [c_void_p, c_size_t, c_void_p, POINTER(c_uint64)]),
    "cn_encode_checked_device": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "cn_n_to_bits_ex_host": (c_int, [c_void_p, c_size_t, c_void_p, c_int, POINTER(c_uint64)]),
    "cn_encode_ex_device": (c_int, [c_void_p, c_size_t, c_void_p, c_int, c_void_p, c_void_p]),
    "cn_n_to_bits2_ex_host": (c_int, [c_void_p, c_size_t, c_void_p, c_int, POINTER(c_uint64)]),
    "cn_encode2_ex_device": (c_int, [c_void_p, c_size_t, c_void_p, c_int, c_void_p, c_void_p]),
    "cn_set_devices": (c_int, [POINTER(c_int), c_int]),
    "cn_get_devices": (c_int, [POINTER(c_int), c_int, POINTER(c_int)]),
    "cn_shard_bounds": (c_int, [c_size_t, c_int, c_int, c_size_t, POINTER(c_size_t), POINTER(c_size_t)]),
    "cn_encode_sharded": (c_int, [c_int, POINTER(c_int), POINTER(c_void_p), POINTER(c_size_t), POINTER(c_void_p), POINTER(c_void_p)]),
    "cn_decode_sharded": (c_int, [c_int, POINTER(c_int), POINTER(c_void_p), POINTER(c_size_t), POINTER(c_size_t), POINTER(c_void_p), POINTER(c_void_p)]),
    "cn_enable_peer_access": (c_int, [POINTER(c_int), c_int]),
    "cn_synchronize_devices": (c_int, [POINTER(c_int), c_int]),
    "cn_hamming_device": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "cn_complement_device": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "cn_reverse_complement_device": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "cn_hamming_host": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, POINTER(c_uint64)]),
    "cn_complement_host": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p]),
    "cn_reverse_complement_host": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p]),
    "cn_words2_for_len": (c_size_t, [c_size_t]),
    "cn_n_to_bits2_host": (c_int, [c_void_p, c_size_t, c_void_p]),
    "cn_bits_to_n2_host": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p]),
    "cn_n_to_bits2_host_batch": (c_int, [POINTER(c_void_p), POINTER(c_size_t), c_size_t, POINTER(c_void_p)]),
    "cn_bits_to_n2_host_batch": (c_int, [POINTER(c_void_p), POINTER(c_size_t), c_size_t, POINTER(c_void_p)]),
    "cn_encode2_device": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "cn_decode2_device": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "cn_generate2_device": (c_int, [c_void_p, c_size_t, c_size_t, c_uint64, c_int, c_void_p]),
    "cn_generate_device": (c_int, [c_void_p, c_size_t, c_size_t, c_uint64, c_int, c_void_p]),
    "cn_generate_words_device": (c_int, [c_void_p, c_size_t, c_size_t, c_uint64, c_void_p]),
    "cn_device_malloc": (c_int, [POINTER(c_void_p), c_size_t]),
    "cn_device_free": (c_int, [c_void_p]),
    "cn_host_malloc_pinned": (c_int, [POINTER(c_void_p), c_size_t]),
    "cn_host_free_pinned": (c_int, [c_void_p]),
    "cn_memcpy_h2d": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "cn_memcpy_d2h": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "cn_stream_synchronize": (c_int, [c_void_p]),
    "cn_time_encode_device": (c_int, [c_void_p, c_size_t, c_void_p, c_int, POINTER(c_float)]),
    "cn_time_decode_device": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_int, POINTER(c_float)]),
    "cn_set_tuning": (c_int, [c_int, c_int, c_int, c_int]),
    "cn_get_tuning": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "cn_set_host_strategy": (c_int, [c_int, c_size_t]),
    "cn_set_host_chunks": (c_int, [c_size_t, c_size_t]),
    "cn_set_host_threads": (c_int, [c_int]),
}


class CuteNucleotidesError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[cn status {code}] {message}")
        self.code = code
        self.message = message


class LengthError(CuteNucleotidesError, ValueError):
    """len > 32 * nwords -- where the reference panics (src/n_to_bits.rs:52-54)."""


_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once) and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C cute_nucleotides_b200/csrc`.  There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library ever diverge
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.cn_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI mismatch: library {lib.cn_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status == CN_OK:
        return
    msg = (load().cn_last_error() or b"").decode("utf-8", "replace")
    if status == CN_ERR_LENGTH:
        raise LengthError(status, msg)
    raise CuteNucleotidesError(status, msg)
