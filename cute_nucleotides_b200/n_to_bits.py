"""Host-side mirror of the reference module `cute_nucleotides::n_to_bits` (src/n_to_bits.rs).

The reference's operator API is a naming family -- `n_to_bits_<variant>(n) -> Vec<u64>` and
`bits_to_n_<variant>(bits, len) -> Vec<u8>`; a new backend is a new suffix.  This module adds the
`_cuda` suffix with the same argument meaning and error behaviour:

    n_to_bits_cuda(n)            mirrors  pub fn n_to_bits_lut(n: &[u8]) -> Vec<u64>            (:34)
    bits_to_n_cuda(bits, len)    mirrors  pub fn bits_to_n_lut(bits: &[u64], len) -> Vec<u8>    (:51)

plus the device-resident pair used for the roofline runs (torch tensors already in HBM).  Every call
goes through the C ABI (include/cute_nucleotides_cuda.h); there is no CPU path here.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import LengthError, check

__all__ = [
    "n_to_bits_cuda", "bits_to_n_cuda", "words_for_len", "n_to_bits_checked_cuda", "encode_checked_device",
    "n_to_bits_lut_cuda", "n_to_bits_ex_cuda", "encode_ex_device", "n_to_bits_batch_cuda", "bits_to_n_batch_cuda", "n_to_bits_cuda_async", "bits_to_n_cuda_async",
    "encode_segmented_device", "decode_segmented_device", "segment_word_offsets",
    "encode_device", "decode_device", "generate_device", "generate_words_device", "LengthError",
    "ENC_PLAIN", "ENC_COUNT", "ENC_LUT_EXACT",
]

ENC_PLAIN, ENC_COUNT, ENC_LUT_EXACT = _lib.CN_ENC_PLAIN, _lib.CN_ENC_COUNT, _lib.CN_ENC_LUT_EXACT


def words_for_len(length: int) -> int:
    """ceil(len / 32), the number of u64 words an encoder returns (src/n_to_bits.rs:35)."""
    return (length >> 5) + (1 if length & 31 else 0)


def _as_u8(n) -> np.ndarray:
    if isinstance(n, np.ndarray):
        if n.dtype != np.uint8:
            raise TypeError("nucleotides must be bytes or a uint8 array")
        return n if n.flags.c_contiguous else np.ascontiguousarray(n)
    return np.frombuffer(n, dtype=np.uint8)   # bytes / bytearray / memoryview, zero-copy, any alignment


def n_to_bits_cuda(n) -> np.ndarray:
    """Encode ASCII {A,C,G,T/U} (any case) to 2-bit codes, 32 per u64.  Returns a uint64 array."""
    src = _as_u8(n)
    out = np.empty(words_for_len(src.size), dtype=np.uint64)   # caller-allocated, like Vec::with_capacity
    if src.size:
        check(_lib.load().cn_n_to_bits_host(src.ctypes.data, src.size, out.ctypes.data))
    return out


def n_to_bits_checked_cuda(n):
    """Like n_to_bits_cuda, and also returns how many bytes were outside {A,C,G,T,U,a,c,g,t,u} (the reference
    maps those silently).  Returns (words, invalid_count)."""
    src = _as_u8(n)
    out = np.empty(words_for_len(src.size), dtype=np.uint64)
    invalid = ctypes.c_uint64(0)
    check(_lib.load().cn_n_to_bits_checked_host(src.ctypes.data, src.size, out.ctypes.data, ctypes.byref(invalid)))
    return out, int(invalid.value)


def n_to_bits_ex_cuda(n, mode: int):
    """Encode with an explicit treatment of bytes outside {A,C,G,T,U,a,c,g,t,u} (include/cute_nucleotides_cuda.h):
    ENC_PLAIN -> (byte >> 1) & 3 like the reference's SIMD encoders; ENC_COUNT -> same words + count;
    ENC_LUT_EXACT -> code 0 like n_to_bits_lut's BYTE_LUT (src/n_to_bits.rs:8-21).  Returns (words, invalid_count)."""
    src = _as_u8(n)
    out = np.empty(words_for_len(src.size), dtype=np.uint64)
    invalid = ctypes.c_uint64(0)
    check(_lib.load().cn_n_to_bits_ex_host(src.ctypes.data, src.size, out.ctypes.data, mode, ctypes.byref(invalid)))
    return out, int(invalid.value)


def n_to_bits_lut_cuda(n) -> np.ndarray:
    """Bit-exact n_to_bits_lut (src/n_to_bits.rs:34-47) on EVERY input, including bytes outside the alphabet (-> 0)."""
    return n_to_bits_ex_cuda(n, ENC_LUT_EXACT)[0]


class AsyncResult:
    """Handle of a cn_*_host_async call.  wait() blocks until the call is done and returns its result (a uint64 array for
    an encode, bytes for a decode); it must be called exactly once.  The input is kept alive by the handle."""

    def __init__(self, request, keep, result, finish):
        self._request, self._keep, self._result, self._finish = request, keep, result, finish

    def wait(self):
        req, self._request = self._request, None
        if req is None:
            raise RuntimeError("wait() was already called on this request")
        check(_lib.load().cn_wait(req))
        return self._finish(self._result)


def n_to_bits_cuda_async(n, out=None) -> AsyncResult:
    """n_to_bits_cuda on a library thread (cn_n_to_bits_host_async): returns at once; .wait() yields the words."""
    src = _as_u8(n)
    res = np.empty(words_for_len(src.size), dtype=np.uint64) if out is None else out
    req = ctypes.c_void_p()
    check(_lib.load().cn_n_to_bits_host_async(src.ctypes.data, src.size, res.ctypes.data, ctypes.byref(req)))
    return AsyncResult(req, src, res, lambda r: r)


def bits_to_n_cuda_async(bits, length: int, out=None) -> AsyncResult:
    """bits_to_n_cuda on a library thread (cn_bits_to_n_host_async): .wait() yields the bytes (or `out` when given)."""
    words = np.ascontiguousarray(bits, dtype=np.uint64)
    res = np.empty(length, dtype=np.uint8) if out is None else out
    req = ctypes.c_void_p()
    check(_lib.load().cn_bits_to_n_host_async(words.ctypes.data, words.size, length, res.ctypes.data, ctypes.byref(req)))
    return AsyncResult(req, words, res, (lambda r: r.tobytes()) if out is None else (lambda r: r))


def _ptr_array(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_void_p))


def _encode_batch(buf, offsets, out, group, fn):
    src = _as_u8(buf)
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    count = offs.size - 1
    lens = np.ascontiguousarray(np.diff(offs))
    wcount = (lens + np.uint64(group - 1)) // np.uint64(group)
    woffs = np.zeros(count + 1, dtype=np.uint64)
    np.cumsum(wcount, out=woffs[1:])
    words = np.empty(int(woffs[-1]), dtype=np.uint64) if out is None else out[: int(woffs[-1])]
    if words.dtype != np.uint64 or words.size != int(woffs[-1]) or not words.flags.c_contiguous:
        raise ValueError("out must be a contiguous uint64 array with room for every sequence's words")
    in_ptrs = np.uint64(src.ctypes.data) + offs[:-1]
    out_ptrs = np.uint64(words.ctypes.data) + woffs[:-1] * np.uint64(8)
    lens_c = lens.ctypes.data_as(ctypes.POINTER(ctypes.c_size_t))
    check(fn(_ptr_array(in_ptrs), lens_c, count, _ptr_array(out_ptrs)))
    return words, woffs


def _decode_batch(words, word_offsets, lengths, out, group, fn):
    w = np.ascontiguousarray(words, dtype=np.uint64)
    woffs = np.ascontiguousarray(word_offsets, dtype=np.uint64)
    lens = np.ascontiguousarray(lengths, dtype=np.uint64)
    count = lens.size
    if np.any(lens > np.diff(woffs) * np.uint64(group)):
        raise LengthError(_lib.CN_ERR_LENGTH, _lib.load().cn_length_panic_message().decode())
    offs = np.zeros(count + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    out = np.empty(int(offs[-1]), dtype=np.uint8) if out is None else out[: int(offs[-1])]
    if out.dtype != np.uint8 or out.size != int(offs[-1]) or not out.flags.c_contiguous:
        raise ValueError("out must be a contiguous uint8 array with room for every sequence")
    in_ptrs = np.uint64(w.ctypes.data) + woffs[:-1] * np.uint64(8)
    out_ptrs = np.uint64(out.ctypes.data) + offs[:-1]
    lens_c = lens.ctypes.data_as(ctypes.POINTER(ctypes.c_size_t))
    check(fn(_ptr_array(in_ptrs), lens_c, count, _ptr_array(out_ptrs)))
    return out, offs


def n_to_bits_batch_cuda(buf, offsets, out=None):
    """Encode many independent sequences in ONE call (cn_n_to_bits_host_batch).  Sequence i is buf[offsets[i]:offsets[i+1]]
    (offsets: count+1 non-decreasing integers).  Returns (words, word_offsets): sequence i's packed words -- exactly what
    n_to_bits_cuda(sequence i) returns -- are words[word_offsets[i]:word_offsets[i+1]].  `out`: a uint64 array to reuse."""
    return _encode_batch(buf, offsets, out, 32, _lib.load().cn_n_to_bits_host_batch)


def bits_to_n_batch_cuda(words, word_offsets, lengths, out=None):
    """Inverse of n_to_bits_batch_cuda: returns (buf, offsets) with sequence i at buf[offsets[i]:offsets[i+1]].
    `out`: a uint8 array to reuse."""
    return _decode_batch(words, word_offsets, lengths, out, 32, _lib.load().cn_bits_to_n_host_batch)


def bits_to_n_cuda(bits, length: int) -> bytes:
    """Decode `length` nucleotides from packed words.  Raises LengthError where the reference panics."""
    words = np.ascontiguousarray(bits, dtype=np.uint64)
    if length < 0:
        raise ValueError("length must be non-negative")
    out = np.empty(length, dtype=np.uint8)
    # the length check lives in C as well (status CN_ERR_LENGTH), so the panic text comes from one place
    check(_lib.load().cn_bits_to_n_host(words.ctypes.data, words.size, length, out.ctypes.data))
    return out.tobytes()


# ---- device-resident API (torch tensors; torch is plumbing for memory and streams only) -------------

def _stream_ptr(stream) -> int:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def encode_device(n, out=None, stream=None):
    """n: uint8 CUDA tensor of ASCII nucleotides -> int64 CUDA tensor of ceil(len/32) packed words
    (bit pattern of the reference's u64s; torch has no general uint64 kernels, so the carrier is int64)."""
    import torch
    if n.dtype != torch.uint8 or not n.is_cuda or not n.is_contiguous():
        raise TypeError("encode_device expects a contiguous uint8 CUDA tensor")
    length = n.numel()
    if out is None:
        out = torch.empty(words_for_len(length), dtype=torch.int64, device=n.device)
    elif out.numel() < words_for_len(length) or out.element_size() != 8 or not out.is_contiguous():
        raise ValueError("out must be a contiguous 8-byte-element tensor with ceil(len/32) elements")
    with torch.cuda.device(n.device):
        check(_lib.load().cn_encode_device(n.data_ptr(), length, out.data_ptr(), _stream_ptr(stream)))
    return out


def encode_checked_device(n, counter, out=None, stream=None):
    """encode_device + validation: `counter` (one-element int64 CUDA tensor) is incremented by the number of
    bytes outside the alphabet.  Asynchronous like encode_device; read the counter after synchronising."""
    import torch
    if n.dtype != torch.uint8 or not n.is_cuda or not n.is_contiguous():
        raise TypeError("encode_checked_device expects a contiguous uint8 CUDA tensor")
    if counter.element_size() != 8 or counter.numel() < 1 or not counter.is_cuda:
        raise TypeError("counter must be a one-element 8-byte CUDA tensor")
    length = n.numel()
    if out is None:
        out = torch.empty(words_for_len(length), dtype=torch.int64, device=n.device)
    with torch.cuda.device(n.device):
        check(_lib.load().cn_encode_checked_device(n.data_ptr(), length, out.data_ptr(), counter.data_ptr(), _stream_ptr(stream)))
    return out


def encode_ex_device(n, mode: int, counter=None, out=None, stream=None):
    """encode_device with an explicit mode (ENC_PLAIN / ENC_COUNT / ENC_LUT_EXACT); `counter` (one-element int64 CUDA
    tensor, incremented) is required for ENC_COUNT and optional for ENC_LUT_EXACT."""
    import torch
    if n.dtype != torch.uint8 or not n.is_cuda or not n.is_contiguous():
        raise TypeError("encode_ex_device expects a contiguous uint8 CUDA tensor")
    if counter is not None and (counter.element_size() != 8 or counter.numel() < 1 or not counter.is_cuda):
        raise TypeError("counter must be a one-element 8-byte CUDA tensor")
    length = n.numel()
    if out is None:
        out = torch.empty(words_for_len(length), dtype=torch.int64, device=n.device)
    with torch.cuda.device(n.device):
        check(_lib.load().cn_encode_ex_device(n.data_ptr(), length, out.data_ptr(), mode,
                                              counter.data_ptr() if counter is not None else None, _stream_ptr(stream)))
    return out


def segment_word_offsets(offsets):
    """offsets: int64 CUDA (or CPU) tensor of count+1 byte offsets -> word offsets of the padded packed layout
    (sequence i gets ceil(len_i / 32) words), same device."""
    import torch
    lens = offsets[1:] - offsets[:-1]
    woff = torch.zeros_like(offsets)
    torch.cumsum((lens + 31) >> 5, dim=0, out=woff[1:])
    return woff


def encode_segmented_device(n, offsets, word_offsets=None, out=None, stream=None):
    """Many sequences resident in HBM, tightly concatenated in the uint8 CUDA tensor `n`: sequence i = n[offsets[i]:offsets[i+1]]
    (offsets: int64 CUDA tensor, count+1 entries, offsets[0] == 0, offsets[-1] == n.numel()).  Returns (words, word_offsets):
    sequence i's packed words -- what n_to_bits_lut(sequence i) yields -- are words[word_offsets[i]:word_offsets[i+1]]."""
    import torch
    if n.dtype != torch.uint8 or not n.is_cuda or offsets.dtype != torch.int64 or not offsets.is_cuda:
        raise TypeError("encode_segmented_device expects a uint8 CUDA tensor and int64 CUDA offsets")
    if word_offsets is None:
        word_offsets = segment_word_offsets(offsets)
    count = offsets.numel() - 1
    total_words = int(word_offsets[-1].item())
    if out is None:
        out = torch.empty(total_words, dtype=torch.int64, device=n.device)
    with torch.cuda.device(n.device):
        check(_lib.load().cn_encode_segmented_device(n.data_ptr(), offsets.data_ptr(), word_offsets.data_ptr(), count, n.numel(),
                                                     total_words, out.data_ptr(), _stream_ptr(stream)))
    return out, word_offsets


def decode_segmented_device(words, offsets, word_offsets, out=None, stream=None):
    """Inverse of encode_segmented_device: returns the uint8 CUDA tensor of offsets[-1] ASCII bytes."""
    import torch
    count = offsets.numel() - 1
    total_len = int(offsets[-1].item())
    if out is None:
        out = torch.empty(total_len, dtype=torch.uint8, device=words.device)
    with torch.cuda.device(words.device):
        check(_lib.load().cn_decode_segmented_device(words.data_ptr(), offsets.data_ptr(), word_offsets.data_ptr(), count, total_len,
                                                     words.numel(), out.data_ptr(), _stream_ptr(stream)))
    return out


def decode_device(bits, length: int, out=None, stream=None):
    """bits: 8-byte-element CUDA tensor of packed words -> uint8 CUDA tensor of `length` ASCII bytes."""
    import torch
    if bits.element_size() != 8 or not bits.is_cuda or not bits.is_contiguous():
        raise TypeError("decode_device expects a contiguous int64/uint64 CUDA tensor")
    if out is None:
        if length > bits.numel() * 32:
            raise LengthError(_lib.CN_ERR_LENGTH, _lib.load().cn_length_panic_message().decode())
        out = torch.empty(length, dtype=torch.uint8, device=bits.device)
    elif out.numel() < length or out.dtype != torch.uint8 or not out.is_contiguous():
        raise ValueError("out must be a contiguous uint8 tensor with at least `length` elements")
    with torch.cuda.device(bits.device):
        check(_lib.load().cn_decode_device(bits.data_ptr(), bits.numel(), length, out.data_ptr(), _stream_ptr(stream)))
    return out


def generate_device(out, offset: int, seed: int, alphabet: int = 4, stream=None):
    """Fill a uint8 CUDA tensor with the shared synthetic sequence (global nucleotide index offset+i)."""
    import torch
    with torch.cuda.device(out.device):
        check(_lib.load().cn_generate_device(out.data_ptr(), offset, out.numel(), seed, alphabet, _stream_ptr(stream)))
    return out


def generate_words_device(out, first_word: int, seed: int, stream=None):
    import torch
    with torch.cuda.device(out.device):
        check(_lib.load().cn_generate_words_device(out.data_ptr(), first_word, out.numel(), seed, _stream_ptr(stream)))
    return out
