"""numpy-facing wrapper of oracle/libcn_oracle.so.  TEST INFRASTRUCTURE (see oracle/cn_oracle.c)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_int, c_size_t, c_uint64, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libcn_oracle.so")

ENCODERS = ("lut", "pext", "shift", "movemask", "mul")
DECODERS = ("lut", "shuffle", "pdep", "clmul")
ENCODERS2 = ("lut", "pext")          # src/n_to_bits2.rs
DECODERS2 = ("lut", "pdep")


def build_oracle() -> str:
    src = os.path.join(ORACLE_DIR, "cn_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ORACLE_SO


class Oracle:
    def __init__(self):
        self.lib = ctypes.CDLL(build_oracle())
        L = self.lib
        L.oracle_words_for_len.restype = c_size_t
        L.oracle_words_for_len.argtypes = [c_size_t]
        L.oracle_count_invalid.restype = c_size_t
        L.oracle_count_invalid.argtypes = [c_void_p, c_size_t]
        for v in ENCODERS:
            f = getattr(L, f"oracle_n_to_bits_{v}")
            f.restype, f.argtypes = None, [c_void_p, c_size_t, c_void_p]
        for v in DECODERS:
            f = getattr(L, f"oracle_bits_to_n_{v}")
            f.restype, f.argtypes = c_int, [c_void_p, c_size_t, c_size_t, c_void_p]
        L.oracle_encode_mt.restype, L.oracle_encode_mt.argtypes = c_int, [c_int, c_void_p, c_size_t, c_void_p, c_int]
        L.oracle_decode_mt.restype, L.oracle_decode_mt.argtypes = c_int, [c_int, c_void_p, c_size_t, c_size_t, c_void_p, c_int]
        L.oracle_generate.restype, L.oracle_generate.argtypes = c_int, [c_void_p, c_size_t, c_size_t, c_uint64, c_int]
        L.oracle_generate_mt.restype, L.oracle_generate_mt.argtypes = c_int, [c_void_p, c_size_t, c_size_t, c_uint64, c_int, c_int]
        L.oracle_generate_words.restype, L.oracle_generate_words.argtypes = None, [c_void_p, c_size_t, c_size_t, c_uint64]
        L.oracle_canonical.restype, L.oracle_canonical.argtypes = None, [c_void_p, c_size_t, c_void_p]
        L.oracle_count_invalid2.restype = c_size_t
        L.oracle_count_invalid2.argtypes = [c_void_p, c_size_t]
        L.oracle_hamming.restype, L.oracle_hamming.argtypes = c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]
        L.oracle_complement.restype, L.oracle_complement.argtypes = c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_int]
        L.oracle_time_small_calls.restype = ctypes.c_double
        L.oracle_time_small_calls.argtypes = [c_void_p, c_void_p, c_size_t, c_int, c_int]
        L.oracle_cpu_ok.restype = c_int
        L.oracle_words2_for_len.restype, L.oracle_words2_for_len.argtypes = c_size_t, [c_size_t]
        for v in ENCODERS2:
            f = getattr(L, f"oracle_n_to_bits2_{v}")
            f.restype, f.argtypes = None, [c_void_p, c_size_t, c_void_p]
        for v in DECODERS2:
            f = getattr(L, f"oracle_bits_to_n2_{v}")
            f.restype, f.argtypes = c_int, [c_void_p, c_size_t, c_size_t, c_void_p]
        L.oracle_generate2.restype, L.oracle_generate2.argtypes = c_int, [c_void_p, c_size_t, c_size_t, c_uint64, c_int]
        L.oracle_canonical2.restype, L.oracle_canonical2.argtypes = None, [c_void_p, c_size_t, c_void_p]
        self.simd_ok = bool(L.oracle_cpu_ok())
        self.threads = os.cpu_count() or 1

    # -- helpers ---------------------------------------------------------------------------------
    @staticmethod
    def _u8(n) -> np.ndarray:
        if isinstance(n, np.ndarray):
            assert n.dtype == np.uint8
            return np.ascontiguousarray(n)
        return np.frombuffer(bytes(n), dtype=np.uint8)

    def words_for_len(self, length: int) -> int:
        return self.lib.oracle_words_for_len(length)

    # -- reference function family ------------------------------------------------------------------
    def n_to_bits(self, n, variant: str = "lut") -> np.ndarray:
        a = self._u8(n)
        out = np.empty(self.words_for_len(a.size), dtype=np.uint64)
        getattr(self.lib, f"oracle_n_to_bits_{variant}")(a.ctypes.data, a.size, out.ctypes.data)
        return out

    def bits_to_n(self, bits, length: int, variant: str = "lut") -> bytes:
        w = np.ascontiguousarray(bits, dtype=np.uint64)
        cap = length if variant == "lut" else max(length, w.size * 32)   # SIMD decoders write whole words
        out = np.empty(cap, dtype=np.uint8)
        rc = getattr(self.lib, f"oracle_bits_to_n_{variant}")(w.ctypes.data, w.size, length, out.ctypes.data)
        if rc == 1:
            raise ValueError("The length is greater than the number of nucleotides!")
        return out[:length].tobytes()

    # -- multi-threaded drivers (timed CPU baseline; also used to check multi-GiB device results) ----
    def encode_mt(self, n: np.ndarray, variant: str = "movemask", threads: int | None = None, out=None) -> np.ndarray:
        kind = {"movemask": 0, "lut": 2}[variant]
        if out is None:
            out = np.empty(self.words_for_len(n.size), dtype=np.uint64)
        rc = self.lib.oracle_encode_mt(kind, n.ctypes.data, n.size, out.ctypes.data, threads or self.threads)
        assert rc == 0, rc
        return out

    def decode_mt(self, bits: np.ndarray, length: int, variant: str = "shuffle", threads: int | None = None, out=None) -> np.ndarray:
        kind = {"shuffle": 1, "lut": 3}[variant]
        if out is None:
            out = np.empty(length, dtype=np.uint8)
        rc = self.lib.oracle_decode_mt(kind, bits.ctypes.data, bits.size, length, out.ctypes.data, threads or self.threads)
        if rc == 1:
            raise ValueError("The length is greater than the number of nucleotides!")
        assert rc == 0, rc
        return out

    # -- synthetic data -----------------------------------------------------------------------------
    def generate(self, length: int, seed: int, alphabet: int = 4, offset: int = 0, out=None) -> np.ndarray:
        if out is None:
            out = np.empty(length, dtype=np.uint8)
        rc = self.lib.oracle_generate_mt(out.ctypes.data, offset, length, seed, alphabet, self.threads)
        assert rc == 0, rc
        return out

    def generate_words(self, nwords: int, seed: int, first_word: int = 0) -> np.ndarray:
        out = np.empty(nwords, dtype=np.uint64)
        self.lib.oracle_generate_words(out.ctypes.data, first_word, nwords, seed)
        return out

    def canonical(self, n) -> bytes:
        a = self._u8(n)
        out = np.empty(a.size, dtype=np.uint8)
        self.lib.oracle_canonical(a.ctypes.data, a.size, out.ctypes.data)
        return out.tobytes()

    def count_invalid(self, n) -> int:
        a = self._u8(n)
        return self.lib.oracle_count_invalid(a.ctypes.data, a.size)

    # -- operations on packed words, defined through the codec (parity unpinned: no reference code) -------
    def hamming(self, a, b, length: int) -> int:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        res = ctypes.c_uint64(0)
        rc = self.lib.oracle_hamming(a.ctypes.data, b.ctypes.data, min(a.size, b.size), length, ctypes.byref(res))
        if rc == 1:
            raise ValueError("The length is greater than the number of nucleotides!")
        return int(res.value)

    def complement(self, bits, length: int, reverse: bool = False) -> np.ndarray:
        w = np.ascontiguousarray(bits, dtype=np.uint64)
        out = np.empty(self.words_for_len(length), dtype=np.uint64)
        rc = self.lib.oracle_complement(w.ctypes.data, w.size, length, out.ctypes.data, int(reverse))
        if rc == 1:
            raise ValueError("The length is greater than the number of nucleotides!")
        return out

    def time_small_calls(self, buf, offsets, reps: int = 1, roundtrip: bool = True) -> float:
        """seconds for `reps` passes of one-call-per-sequence movemask (+ shuffle) on ONE thread, allocation inside"""
        a = self._u8(buf)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        return float(self.lib.oracle_time_small_calls(a.ctypes.data, o.ctypes.data, o.size - 1, reps, int(roundtrip)))

    def count_invalid2(self, n) -> int:
        a = self._u8(n)
        return self.lib.oracle_count_invalid2(a.ctypes.data, a.size)

    # -- base-5 codec (src/n_to_bits2.rs) --------------------------------------------------------------
    def words2_for_len(self, length: int) -> int:
        return self.lib.oracle_words2_for_len(length)

    def n_to_bits2(self, n, variant: str = "lut") -> np.ndarray:
        a = self._u8(n)
        out = np.empty(self.words2_for_len(a.size), dtype=np.uint64)
        getattr(self.lib, f"oracle_n_to_bits2_{variant}")(a.ctypes.data, a.size, out.ctypes.data)
        return out

    def bits_to_n2(self, bits, length: int, variant: str = "lut") -> bytes:
        w = np.ascontiguousarray(bits, dtype=np.uint64)
        out = np.empty(27 * w.size + 8, dtype=np.uint8)          # decoders write whole triplets / 32-byte stores
        rc = getattr(self.lib, f"oracle_bits_to_n2_{variant}")(w.ctypes.data, w.size, length, out.ctypes.data)
        if rc == 1:
            raise ValueError("The length is greater than the number of nucleotides!")
        return out[:length].tobytes()

    def generate2(self, length: int, seed: int, alphabet: int = 5, offset: int = 0) -> np.ndarray:
        out = np.empty(length, dtype=np.uint8)
        assert self.lib.oracle_generate2(out.ctypes.data, offset, length, seed, alphabet) == 0
        return out

    def canonical2(self, n) -> bytes:
        a = self._u8(n)
        out = np.empty(a.size, dtype=np.uint8)
        self.lib.oracle_canonical2(a.ctypes.data, a.size, out.ctypes.data)
        return out.tobytes()
