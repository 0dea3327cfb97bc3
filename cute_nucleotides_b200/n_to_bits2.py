"""Host-side mirror of the reference module `cute_nucleotides::n_to_bits2` (src/n_to_bits2.rs): the base-5
codec for {A, C, T/U, G, N} -- three nucleotides per 7 bits, 27 per u64.

    n_to_bits2_cuda(n)           mirrors  pub fn n_to_bits2_lut(n: &[u8]) -> Vec<u64>           (:37)
    bits_to_n2_cuda(bits, len)   mirrors  pub fn bits_to_n2_lut(bits: &[u64], len) -> Vec<u8>   (:78)

Same conventions as n_to_bits.py: every call goes through the C ABI, no CPU path, LengthError where the
reference panics (:79-81)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import LengthError, check
from .n_to_bits import _as_u8, _decode_batch, _encode_batch, _stream_ptr

__all__ = ["n_to_bits2_cuda", "bits_to_n2_cuda", "words2_for_len", "encode2_device", "decode2_device", "generate2_device",
           "n_to_bits2_ex_cuda", "n_to_bits2_lut_cuda", "encode2_ex_device", "n_to_bits2_batch_cuda", "bits_to_n2_batch_cuda"]


def words2_for_len(length: int) -> int:
    """(len / 27) + (len % 27 != 0)  (src/n_to_bits2.rs:38)."""
    return length // 27 + (1 if length % 27 else 0)


def n_to_bits2_cuda(n) -> np.ndarray:
    src = _as_u8(n)
    out = np.empty(words2_for_len(src.size), dtype=np.uint64)
    if src.size:
        check(_lib.load().cn_n_to_bits2_host(src.ctypes.data, src.size, out.ctypes.data))
    return out


def n_to_bits2_ex_cuda(n, mode: int):
    """Base-5 encode with an explicit treatment of bytes outside {A,C,G,T,U,N,a,c,g,t,u,n}: ENC_PLAIN -> digit looked up by
    the low 3 bits like n_to_bits2_pext (src/n_to_bits2.rs:127-136); ENC_COUNT -> same words + count; ENC_LUT_EXACT ->
    digit 0 like n_to_bits2_lut's BYTE_LUT (:8-23).  Returns (words, invalid_count)."""
    src = _as_u8(n)
    out = np.empty(words2_for_len(src.size), dtype=np.uint64)
    invalid = ctypes.c_uint64(0)
    check(_lib.load().cn_n_to_bits2_ex_host(src.ctypes.data, src.size, out.ctypes.data, mode, ctypes.byref(invalid)))
    return out, int(invalid.value)


def n_to_bits2_lut_cuda(n) -> np.ndarray:
    """Bit-exact n_to_bits2_lut (src/n_to_bits2.rs:37-74) on every input."""
    return n_to_bits2_ex_cuda(n, _lib.CN_ENC_LUT_EXACT)[0]


def n_to_bits2_batch_cuda(buf, offsets, out=None):
    """Many independent sequences in ONE call (cn_n_to_bits2_host_batch); see n_to_bits_batch_cuda.  27 nucleotides per word."""
    return _encode_batch(buf, offsets, out, 27, _lib.load().cn_n_to_bits2_host_batch)


def bits_to_n2_batch_cuda(words, word_offsets, lengths, out=None):
    return _decode_batch(words, word_offsets, lengths, out, 27, _lib.load().cn_bits_to_n2_host_batch)


def bits_to_n2_cuda(bits, length: int) -> bytes:
    words = np.ascontiguousarray(bits, dtype=np.uint64)
    if length < 0:
        raise ValueError("length must be non-negative")
    out = np.empty(length, dtype=np.uint8)
    check(_lib.load().cn_bits_to_n2_host(words.ctypes.data, words.size, length, out.ctypes.data))
    return out.tobytes()


def encode2_device(n, out=None, stream=None):
    import torch
    if n.dtype != torch.uint8 or not n.is_cuda or not n.is_contiguous():
        raise TypeError("encode2_device expects a contiguous uint8 CUDA tensor")
    length = n.numel()
    if out is None:
        out = torch.empty(words2_for_len(length), dtype=torch.int64, device=n.device)
    elif out.numel() < words2_for_len(length) or out.element_size() != 8 or not out.is_contiguous():
        raise ValueError("out must be a contiguous 8-byte-element tensor with ceil(len/27) elements")
    with torch.cuda.device(n.device):
        check(_lib.load().cn_encode2_device(n.data_ptr(), length, out.data_ptr(), _stream_ptr(stream)))
    return out


def encode2_ex_device(n, mode: int, counter=None, out=None, stream=None):
    import torch
    if n.dtype != torch.uint8 or not n.is_cuda or not n.is_contiguous():
        raise TypeError("encode2_ex_device expects a contiguous uint8 CUDA tensor")
    length = n.numel()
    if out is None:
        out = torch.empty(words2_for_len(length), dtype=torch.int64, device=n.device)
    with torch.cuda.device(n.device):
        check(_lib.load().cn_encode2_ex_device(n.data_ptr(), length, out.data_ptr(), mode,
                                               counter.data_ptr() if counter is not None else None, _stream_ptr(stream)))
    return out


def decode2_device(bits, length: int, out=None, stream=None):
    import torch
    if bits.element_size() != 8 or not bits.is_cuda or not bits.is_contiguous():
        raise TypeError("decode2_device expects a contiguous int64/uint64 CUDA tensor")
    if out is None:
        if length > bits.numel() * 27:
            raise LengthError(_lib.CN_ERR_LENGTH, _lib.load().cn_length_panic_message().decode())
        out = torch.empty(length, dtype=torch.uint8, device=bits.device)
    elif out.numel() < length or out.dtype != torch.uint8 or not out.is_contiguous():
        raise ValueError("out must be a contiguous uint8 tensor with at least `length` elements")
    with torch.cuda.device(bits.device):
        check(_lib.load().cn_decode2_device(bits.data_ptr(), bits.numel(), length, out.data_ptr(), _stream_ptr(stream)))
    return out


def generate2_device(out, offset: int, seed: int, alphabet: int = 5, stream=None):
    import torch
    with torch.cuda.device(out.device):
        check(_lib.load().cn_generate2_device(out.data_ptr(), offset, out.numel(), seed, alphabet, _stream_ptr(stream)))
    return out
