"""numpy restatement of the reference's scalar codec -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It exists as a second, independently written statement of n_to_bits_lut /
bits_to_n_lut (src/n_to_bits.rs:34-69) and n_to_bits2_lut / bits_to_n2_lut (src/n_to_bits2.rs:37-107)
so that the C oracle (cn_oracle.c) is checked against something other than itself, in addition to
the reference's own known-answer tests.  The reference cannot be compiled here (no rustc/cargo).
"""
from __future__ import annotations

import numpy as np

PANIC_TEXT = "The length is greater than the number of nucleotides!"   # src/n_to_bits.rs:53

# src/n_to_bits.rs:8-21 -- 128 entries, unlisted bytes -> 0
BYTE_LUT = np.zeros(128, dtype=np.uint8)
for _ch, _code in ((b"a", 0), (b"t", 2), (b"u", 2), (b"c", 1), (b"g", 3),
                   (b"A", 0), (b"T", 2), (b"U", 2), (b"C", 1), (b"G", 3)):
    BYTE_LUT[_ch[0]] = _code
# src/n_to_bits.rs:23-30
BITS_LUT = np.frombuffer(b"ACTG", dtype=np.uint8)


def words_for_len(length: int) -> int:
    return (length >> 5) + (0 if length & 31 == 0 else 1)        # src/n_to_bits.rs:35


def n_to_bits_lut(n: bytes | np.ndarray) -> np.ndarray:
    """src/n_to_bits.rs:34-47: res[i >> 5] |= BYTE_LUT[n[i]] << ((i & 31) << 1)."""
    a = np.frombuffer(n, dtype=np.uint8) if not isinstance(n, np.ndarray) else n
    if a.size and int(a.max()) >= 128:
        raise ValueError("byte >= 0x80: out-of-bounds LUT read in the reference (undefined)")
    words = words_for_len(a.size)
    codes = np.zeros(words * 32, dtype=np.uint64)
    codes[: a.size] = BYTE_LUT[a]
    shifts = (np.arange(32, dtype=np.uint64) << np.uint64(1))
    return np.bitwise_or.reduce(codes.reshape(words, 32) << shifts, axis=1).astype(np.uint64) if words else np.zeros(0, np.uint64)


def bits_to_n_lut(bits: np.ndarray, length: int) -> bytes:
    """src/n_to_bits.rs:51-69: out[i] = BITS_LUT[(bits[i >> 5] >> ((i & 31) << 1)) & 3]."""
    bits = np.asarray(bits, dtype=np.uint64)
    if length > (bits.size << 5):
        raise ValueError(PANIC_TEXT)
    i = np.arange(length, dtype=np.uint64)
    cur = bits[(i >> np.uint64(5)).astype(np.int64)] if length else np.zeros(0, np.uint64)
    codes = (cur >> ((i & np.uint64(31)) << np.uint64(1))) & np.uint64(3)
    return BITS_LUT[codes.astype(np.int64)].tobytes()


def canonical(n: bytes) -> bytes:
    """What decode(encode(n)) returns: upper case, U -> T."""
    a = np.frombuffer(n, dtype=np.uint8)
    return BITS_LUT[BYTE_LUT[a]].tobytes()


# ---- base-5 codec (src/n_to_bits2.rs) ---------------------------------------------------------------
BYTE_LUT2 = np.zeros(128, dtype=np.uint8)                              # src/n_to_bits2.rs:8-23
for _ch, _code in ((b"a", 0), (b"c", 1), (b"t", 2), (b"u", 2), (b"g", 3), (b"n", 4),
                   (b"A", 0), (b"C", 1), (b"T", 2), (b"U", 2), (b"G", 3), (b"N", 4)):
    BYTE_LUT2[_ch[0]] = _code
BITS_LUT2 = np.frombuffer(b"ACTGN", dtype=np.uint8)                    # src/n_to_bits2.rs:25-33


def words2_for_len(length: int) -> int:
    return length // 27 + (0 if length % 27 == 0 else 1)              # src/n_to_bits2.rs:38


def n_to_bits2_lut(n: bytes | np.ndarray) -> np.ndarray:
    """src/n_to_bits2.rs:37-74: triplet -> a + 5b + 25c (7 bits), 9 triplets per u64; a ragged last
    triplet is encoded from the bytes that exist (:59-70)."""
    a = np.frombuffer(n, dtype=np.uint8) if not isinstance(n, np.ndarray) else n
    words = words2_for_len(a.size)
    digits = np.zeros(words * 27, dtype=np.uint64)
    digits[: a.size] = BYTE_LUT2[a]
    tri = digits.reshape(words * 9, 3)
    enc = tri[:, 0] + tri[:, 1] * np.uint64(5) + tri[:, 2] * np.uint64(25)
    shifts = np.arange(9, dtype=np.uint64) * np.uint64(7)
    return np.bitwise_or.reduce(enc.reshape(words, 9) << shifts, axis=1).astype(np.uint64) if words else np.zeros(0, np.uint64)


def bits_to_n2_lut(bits: np.ndarray, length: int) -> bytes:
    """src/n_to_bits2.rs:78-107."""
    bits = np.asarray(bits, dtype=np.uint64)
    if length > bits.size * 27:
        raise ValueError(PANIC_TEXT)
    shifts = np.arange(9, dtype=np.uint64) * np.uint64(7)
    enc = (bits[:, None] >> shifts) & np.uint64(0x7F)
    a = enc % np.uint64(5)
    b = (enc // np.uint64(5)) % np.uint64(5)
    c = enc // np.uint64(25)
    digits = np.stack([a, b, c], axis=2).reshape(-1)[:length]
    digits = np.minimum(digits, np.uint64(4))    # enc > 124 cannot come from the encoder; keep indexing defined
    return BITS_LUT2[digits.astype(np.int64)].tobytes()
