"""Operations on packed 2-bit words: Hamming distance, complement, reverse complement.

PARITY UNPINNED: the reference motivates these (README.md:21-23, 45, 415-418) but ships no code for them.  They are
defined through the reference's own codec -- hamming(a, b, len) counts the i < len where bits_to_n_lut(a)[i] !=
bits_to_n_lut(b)[i]; complement / reverse_complement are n_to_bits_lut of the decoded sequence with A<->T, C<->G
(reversed) -- and checked against exactly that definition in tests/test_packed_ops*.py."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import check
from .n_to_bits import _stream_ptr, words_for_len

__all__ = ["hamming_cuda", "complement_cuda", "reverse_complement_cuda", "hamming_device", "complement_device",
           "reverse_complement_device"]


def hamming_cuda(a, b, length: int) -> int:
    """Number of positions i < length at which the two packed sequences differ (host arrays of uint64)."""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    res = ctypes.c_uint64(0)
    check(_lib.load().cn_hamming_host(a.ctypes.data, b.ctypes.data, min(a.size, b.size), length, ctypes.byref(res)))
    return int(res.value)


def _complement_host(bits, length: int, fn) -> np.ndarray:
    w = np.ascontiguousarray(bits, dtype=np.uint64)
    out = np.empty(words_for_len(length), dtype=np.uint64)
    check(fn(w.ctypes.data, w.size, length, out.ctypes.data))
    return out


def complement_cuda(bits, length: int) -> np.ndarray:
    return _complement_host(bits, length, _lib.load().cn_complement_host)


def reverse_complement_cuda(bits, length: int) -> np.ndarray:
    return _complement_host(bits, length, _lib.load().cn_reverse_complement_host)


def hamming_device(a, b, length: int, result=None, stream=None):
    """a, b: 8-byte-element CUDA tensors of packed words.  `result` (one-element int64 CUDA tensor) is INCREMENTED by the
    distance and returned; a fresh zeroed one is made when omitted.  Asynchronous."""
    import torch
    if result is None:
        result = torch.zeros(1, dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        check(_lib.load().cn_hamming_device(a.data_ptr(), b.data_ptr(), min(a.numel(), b.numel()), length, result.data_ptr(),
                                            _stream_ptr(stream)))
    return result


def _complement_device(bits, length: int, out, stream, fn):
    import torch
    if out is None:
        out = torch.empty(words_for_len(length), dtype=torch.int64, device=bits.device)
    with torch.cuda.device(bits.device):
        check(fn(bits.data_ptr(), bits.numel(), length, out.data_ptr(), _stream_ptr(stream)))
    return out


def complement_device(bits, length: int, out=None, stream=None):
    return _complement_device(bits, length, out, stream, _lib.load().cn_complement_device)


def reverse_complement_device(bits, length: int, out=None, stream=None):
    return _complement_device(bits, length, out, stream, _lib.load().cn_reverse_complement_device)
