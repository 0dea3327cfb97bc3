"""GPU parity tests (run on the B200 box with `-m gpu`).  Every product call goes through the C ABI
(include/cute_nucleotides_cuda.h) via cute_nucleotides_b200; the oracle is only the checker.

Bar: bit-exact (integer/byte work).  Layout mirrors the reference's own tests (src/n_to_bits.rs:408-470):
known-answer vectors first, then the cases the reference never exercises (SURVEY section 4 "gaps"):
lower case, U, body+tail lengths, truncating len, the len panic, empty input, unaligned slices, > 4 GiB.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
    GOLD = json.load(f)

SIZES = [1, 3, 4, 5, 15, 16, 17, 31, 32, 33, 47, 48, 49, 63, 64, 65, 127, 128, 129, 255, 1000, 4095, 4096, 4097,
         16383, 40000, 65536 + 17, (1 << 20), (1 << 20) + 17, 3 * (1 << 20) + 5]
TUNINGS = [(v, u, t) for v in (16, 32) for u in (1, 2, 4, 8) for t in (64, 128, 256, 512)]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a GPU; there is no CPU fallback to test"
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(autouse=True)
def default_tuning(cn):
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    saved = []
    for d in (0, 1):
        import ctypes
        v, u, t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        lib.cn_get_tuning(d, ctypes.byref(v), ctypes.byref(u), ctypes.byref(t))
        saved.append((d, v.value, u.value, t.value))
    yield
    for d, v, u, t in saved:
        lib.cn_set_tuning(d, v, u, t)
    lib.cn_set_host_strategy(0, 16 << 20)


def to_np_u64(t):
    return t.cpu().numpy().view(np.uint64)


# ---------------------------------------------------------------------------------------------------
# the reference's own known-answer tests, through the drop-in signatures
# ---------------------------------------------------------------------------------------------------
def test_reference_kat_encode(cn):
    for case in GOLD["encode"]:                      # src/n_to_bits.rs:414-416 (and :427-450)
        got = cn.n_to_bits_cuda(case["input"].encode())
        assert got.dtype == np.uint64 and [str(int(x)) for x in got] == case["words"]


def test_reference_kat_decode(cn):
    for case in GOLD["decode"]:                      # src/n_to_bits.rs:421-422 (and :455-468)
        words = np.array([int(w) for w in case["words"]], dtype=np.uint64)
        assert cn.bits_to_n_cuda(words, case["len"]) == case["output"].encode()


def test_reference_bench_fixture(cn):
    fx = GOLD["bench_fixture"]                       # benches/bench_n_to_bits.rs:68-78
    n = (fx["unit"] * fx["repeat"]).encode()
    words = cn.n_to_bits_cuda(n)
    assert words.size == fx["nwords"] and np.all(words == np.uint64(int(fx["every_word"])))
    assert cn.bits_to_n_cuda(words, len(n)) == n


def test_length_panic(cn):
    with pytest.raises(cn.LengthError, match=GOLD["panic_text"]):
        cn.bits_to_n_cuda(np.zeros(2, dtype=np.uint64), 65)
    assert cn.bits_to_n_cuda(np.zeros(2, dtype=np.uint64), 64) == b"A" * 64


def test_empty(cn, torch_cuda):
    torch = torch_cuda
    assert cn.n_to_bits_cuda(b"").size == 0
    assert cn.bits_to_n_cuda(np.zeros(0, dtype=np.uint64), 0) == b""
    e = cn.encode_device(torch.empty(0, dtype=torch.uint8, device="cuda"))
    assert e.numel() == 0
    d = cn.decode_device(torch.empty(4, dtype=torch.int64, device="cuda"), 0)
    assert d.numel() == 0


# ---------------------------------------------------------------------------------------------------
# host-slice path (the literal drop-in) vs the scalar oracle
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", SIZES)
def test_host_encode_decode_vs_oracle(cn, oracle, size):
    n = oracle.generate(size, seed=size, alphabet=10)            # mixed case with U
    ref = oracle.n_to_bits(n, "lut")
    got = cn.n_to_bits_cuda(n)
    assert np.array_equal(got, ref)
    assert cn.bits_to_n_cuda(got, size) == oracle.bits_to_n(ref, size, "lut") == oracle.canonical(n)
    # truncating len: fewer nucleotides than the words hold
    for shorter in {size - 1, size // 2, max(size - 33, 0)}:
        if shorter >= 0:
            assert cn.bits_to_n_cuda(got, shorter) == oracle.bits_to_n(ref, shorter, "lut")


@pytest.mark.parametrize("offset", [1, 2, 3, 5, 8, 13, 15])
def test_host_unaligned_slices(cn, oracle, offset):
    base = oracle.generate(100000 + offset, seed=77, alphabet=10)
    n = base[offset:]                                            # a &[u8] can start anywhere
    assert np.array_equal(cn.n_to_bits_cuda(n), oracle.n_to_bits(n, "lut"))
    words = oracle.n_to_bits(n, "lut")
    assert cn.bits_to_n_cuda(words, n.size) == oracle.canonical(n)


@pytest.mark.parametrize("strategy,chunk", [(0, 4096), (0, 1 << 16), (1, 4096), (1, 1 << 20)])
def test_host_pipeline_chunking(cn, oracle, strategy, chunk):
    """Many small chunks through the slot ring (wrap-around, ragged last chunk), both strategies."""
    from cute_nucleotides_b200 import _lib
    assert _lib.load().cn_set_host_strategy(strategy, chunk) == 0
    for size in (chunk * 9 + 17, chunk * 4, chunk - 1, 5):
        n = oracle.generate(size, seed=size ^ 0x55, alphabet=10)
        ref = oracle.n_to_bits(n, "lut")
        got = cn.n_to_bits_cuda(n)
        assert np.array_equal(got, ref)
        assert cn.bits_to_n_cuda(got, size) == oracle.canonical(n)


def test_host_pinned_buffers(cn, oracle, torch_cuda):
    """Pinned caller buffers take the direct-DMA / zero-copy branches of the host pipeline."""
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    size = (1 << 22) + 1234
    n = oracle.generate(size, seed=9, alphabet=10)
    ref = oracle.n_to_bits(n, "lut")
    src = torch.from_numpy(n).pin_memory()
    out = torch.empty(ref.size, dtype=torch.int64).pin_memory()
    dec = torch.empty(size + 64, dtype=torch.uint8).pin_memory()
    for strategy in (0, 1):
        assert lib.cn_set_host_strategy(strategy, 1 << 20) == 0
        out.zero_(); dec.zero_()
        _lib.check(lib.cn_n_to_bits_host(src.data_ptr(), size, out.data_ptr()))
        assert np.array_equal(out.numpy().view(np.uint64), ref)
        _lib.check(lib.cn_bits_to_n_host(out.data_ptr(), ref.size, size, dec.data_ptr()))
        assert dec.numpy()[:size].tobytes() == oracle.canonical(n)
        # unaligned pinned sub-slices
        _lib.check(lib.cn_n_to_bits_host(src.data_ptr() + 3, size - 3, out.data_ptr()))
        assert np.array_equal(out.numpy().view(np.uint64)[: (size - 3 + 31) // 32], oracle.n_to_bits(n[3:], "lut"))
        _lib.check(lib.cn_bits_to_n_host(out.data_ptr(), (size - 3 + 31) // 32, size - 3, dec.data_ptr() + 5))
        assert dec.numpy()[5 : 5 + size - 3].tobytes() == oracle.canonical(n[3:])


# ---------------------------------------------------------------------------------------------------
# device-resident path vs the oracle, across every kernel configuration the library can select
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("vec,unroll,threads", TUNINGS)
def test_device_all_tunings(cn, oracle, torch_cuda, vec, unroll, threads):
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    assert lib.cn_set_tuning(0, vec, unroll, threads) == 0 and lib.cn_set_tuning(1, vec, unroll, threads) == 0
    tile = vec * unroll * threads
    for size in (1, 31, 33, tile - 1, tile, tile + 1, 3 * tile + 37, (1 << 20) + 17):
        n = oracle.generate(size, seed=size + vec, alphabet=10)
        ref = oracle.n_to_bits(n, "lut")
        d_n = torch.from_numpy(n).cuda()
        d_bits = cn.encode_device(d_n)
        assert np.array_equal(to_np_u64(d_bits), ref), (size,)
        d_out = cn.decode_device(d_bits, size)
        assert d_out.cpu().numpy().tobytes() == oracle.canonical(n), (size,)


@pytest.mark.parametrize("in_off", [0, 1, 4, 7, 9, 15, 16, 17])
@pytest.mark.parametrize("size", [0, 1, 14, 15, 16, 17, 31, 32, 33, 50, 4096, 100003])
def test_device_unaligned_pointers(cn, oracle, torch_cuda, in_off, size):
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    n = oracle.generate(size + 64, seed=size * 31 + in_off, alphabet=10)
    d_all = torch.from_numpy(n).cuda()
    sl = n[in_off : in_off + size]
    ref = oracle.n_to_bits(sl, "lut")
    d_bits = torch.full((ref.size + 1,), -1, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.cn_encode_device(d_all.data_ptr() + in_off, size, d_bits.data_ptr(), stream))
    got = to_np_u64(d_bits)
    assert np.array_equal(got[: ref.size], ref)
    assert got[ref.size] == np.uint64(0xFFFFFFFFFFFFFFFF)        # nothing written past ceil(len/32) words
    # decode into an unaligned destination, with guard bytes on both sides
    d_out = torch.full((size + 64,), 0x7E, dtype=torch.uint8, device="cuda")
    _lib.check(lib.cn_decode_device(d_bits.data_ptr(), ref.size, size, d_out.data_ptr() + in_off, stream))
    h = d_out.cpu().numpy()
    assert h[in_off : in_off + size].tobytes() == oracle.canonical(sl)
    assert np.all(h[:in_off] == 0x7E) and np.all(h[in_off + size :] == 0x7E)


def test_device_decode_arbitrary_words_vs_oracle(cn, oracle, torch_cuda):
    """BASELINE config 4 shape at small scale: any bit pattern is a valid packed word."""
    torch = torch_cuda
    nwords = (1 << 18) + 3
    words = oracle.generate_words(nwords, seed=4)
    d_bits = torch.from_numpy(words.view(np.int64)).cuda()
    for length in (32 * nwords, 32 * nwords - 1, 32 * nwords - 31, 32 * (nwords - 1) + 1, 12345):
        got = cn.decode_device(d_bits, length).cpu().numpy().tobytes()
        assert got == oracle.decode_mt(words, length, "lut").tobytes()
    with pytest.raises(cn.LengthError, match=GOLD["panic_text"]):
        cn.decode_device(d_bits, 32 * nwords + 1)


def test_device_generator_matches_host_generator(cn, oracle, torch_cuda):
    torch = torch_cuda
    for alphabet in (4, 10):
        for size, offset in ((1, 0), (7, 8), (100003, 0), (65536, 40000)):
            d = cn.generate_device(torch.empty(size, dtype=torch.uint8, device="cuda"), offset, 0xC0FFEE, alphabet)
            assert np.array_equal(d.cpu().numpy(), oracle.generate(size, 0xC0FFEE, alphabet, offset))
    w = cn.generate_words_device(torch.empty(1000, dtype=torch.int64, device="cuda"), 17, 99)
    assert np.array_equal(to_np_u64(w), oracle.generate_words(1000, 99, first_word=17))


def test_invalid_bytes_follow_the_simd_encoders(cn, oracle):
    """Outside the alphabet the kernel computes (b >> 1) & 3 like the reference's pext/shift/movemask/mul
    body (src/n_to_bits.rs:85,100); documented, not parity-relevant (SURVEY 8a notes)."""
    n = np.arange(128, dtype=np.uint8).repeat(32)              # every 7-bit byte, one word each
    got = cn.n_to_bits_cuda(n)
    if oracle.simd_ok:
        assert np.array_equal(got, oracle.n_to_bits(n, "mul"))
    codes = (np.arange(128) >> 1) & 3
    expect = np.array([sum(int(c) << (2 * k) for k in range(32)) for c in codes], dtype=np.uint64)
    assert np.array_equal(got, expect)


# ---------------------------------------------------------------------------------------------------
# BASELINE configs at full size
# ---------------------------------------------------------------------------------------------------
def test_config2_one_gib_bit_exact(cn, oracle, torch_cuda):
    """BASELINE config 2: 1 GiB random ACGT, encode + decode on one B200, bit-exact vs n_to_bits_lut
    on the identical host-generated buffer (all 2^25 words), decoded output equals the input."""
    torch = torch_cuda
    size = 1 << 30
    n = oracle.generate(size, seed=0xC0FFEE, alphabet=4)
    d_n = cn.generate_device(torch.empty(size, dtype=torch.uint8, device="cuda"), 0, 0xC0FFEE, 4)
    assert np.array_equal(d_n[: 1 << 24].cpu().numpy(), n[: 1 << 24])
    d_bits = cn.encode_device(d_n)
    ref = oracle.encode_mt(n, "lut")
    assert d_bits.numel() == 1 << 25
    assert np.array_equal(to_np_u64(d_bits), ref)
    d_out = cn.decode_device(d_bits, size)
    assert torch.equal(d_out, d_n)                                # alphabet 4 is already canonical
    assert np.array_equal(d_out.cpu().numpy(), n)


def _fold(t):
    """order-sensitive 64-bit checksum of a CUDA tensor's bytes, computed on device with torch"""
    import torch
    v = t.view(torch.int64) if t.dtype != torch.int64 else t
    idx = torch.arange(v.numel(), device=v.device, dtype=torch.int64)
    return int(((v ^ (idx * -7046029254386353131)) * 0x100000001B3).sum().item())   # wraps mod 2^64


def test_config3_config4_ten_gib_properties(cn, oracle, torch_cuda):
    """BASELINE configs 3 and 4 at full size (10 GiB, mixed case with U; > 4 GiB exercises 64-bit indexing).
    Size-independent properties: (i) sampled windows are bit-exact vs the oracle, including windows that
    straddle the 4 GiB and 8 GiB boundaries and the ragged end; (ii) decode(encode(x)) == canonical(x)
    over the whole buffer; (iii) encode is shard-linear: encoding a window alone gives the same words
    (checksum of checksums)."""
    torch = torch_cuda
    size = 10 * (1 << 30) + 21                                   # ragged: exercises the edge warp at > 2^33
    seed = 0xBADC0DE
    d_n = cn.generate_device(torch.empty(size, dtype=torch.uint8, device="cuda"), 0, seed, 10)
    d_bits = cn.encode_device(d_n)
    nwords = cn.words_for_len(size)
    assert d_bits.numel() == nwords
    win = 1 << 22
    starts = [0, (1 << 32) - win // 2, (1 << 33) - win // 2, 5 * (1 << 30) + 32 * 12345, size - win - 21]
    for s in starts:
        s -= s % 32
        host = oracle.generate(min(win, size - s), seed, 10, offset=s)
        assert np.array_equal(d_n[s : s + host.size].cpu().numpy(), host)
        ref = oracle.encode_mt(host, "lut")
        got = to_np_u64(d_bits[s // 32 : s // 32 + ref.size])
        assert np.array_equal(got, ref), hex(s)
        # (iii) the same window encoded on its own
        alone = cn.encode_device(d_n[s : s + host.size])
        assert _fold(alone) == _fold(d_bits[s // 32 : s // 32 + ref.size])
    # last word: high bits beyond the 21 ragged nucleotides are zero
    last = int(to_np_u64(d_bits[-1:])[0])
    assert last >> (2 * (size % 32)) == 0
    # (ii) whole-buffer round trip against the canonical form, chunked so torch temporaries stay small
    d_out = cn.decode_device(d_bits, size)
    lut = torch.zeros(256, dtype=torch.uint8, device="cuda")
    for ch, canon in zip(b"ACGTUacgtu", b"ACGTTACGTT"):
        lut[ch] = canon
    step = 1 << 28
    for s in range(0, size, step):
        e = min(size, s + step)
        assert torch.equal(d_out[s:e], lut[d_n[s:e].long()])
    del d_out, d_n
    # config 4: decode of arbitrary words, sampled windows vs bits_to_n_lut
    d_w = cn.generate_words_device(torch.empty(nwords, dtype=torch.int64, device="cuda"), 0, seed + 1)
    length = 32 * nwords - 7
    d_dec = cn.decode_device(d_w, length)
    for s in (0, (1 << 32) - 4096, (1 << 33) - 4096, length - (1 << 20)):
        s -= s % 32
        cnt = min(1 << 20, length - s)
        words = oracle.generate_words((cnt + 31) // 32, seed + 1, first_word=s // 32)
        assert d_dec[s : s + cnt].cpu().numpy().tobytes() == oracle.decode_mt(words, cnt, "lut").tobytes()
    # idempotence: encode(decode(w)) == w for whole words
    re = cn.encode_device(d_dec[: 32 * (nwords - 1)])
    assert torch.equal(re, d_w[: nwords - 1])


def test_host_api_is_thread_safe(cn, oracle):
    """The reference functions are pure and callable from any thread (SURVEY 8b "Threading"); the C ABI keeps
    per-thread staging.  Eight threads encode/decode different inputs concurrently (ctypes drops the GIL),
    including threads that exit and are replaced (their staging must be released, not corrupted)."""
    import threading
    errors = []

    def worker(seed):
        try:
            for rep in range(3):
                size = 100003 + 4099 * seed + (1 << 20) * (rep == 2)
                n = oracle.generate(size, seed=seed * 10 + rep, alphabet=10)
                ref = oracle.n_to_bits(n, "lut")
                got = cn.n_to_bits_cuda(n)
                if not np.array_equal(got, ref) or cn.bits_to_n_cuda(got, size) != oracle.canonical(n):
                    errors.append((seed, rep))
                m = oracle.generate2(size // 3, seed=seed, alphabet=12)
                if cn.bits_to_n2_cuda(cn.n_to_bits2_cuda(m), m.size) != oracle.canonical2(m):
                    errors.append((seed, rep, "base5"))
        except Exception as e:                      # noqa: BLE001
            errors.append((seed, repr(e)))

    for wave in range(2):
        threads = [threading.Thread(target=worker, args=(wave * 8 + i,)) for i in range(8)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    assert not errors, errors


def test_pageable_host_slices_above_4gib(cn, oracle):
    """The literal drop-in shape at scale: pageable (numpy / Vec-like) buffers on both sides, > 2^32 nucleotides, so the
    staged pipeline's 64-bit offsets, the copier pool and thousands of ring wrap-arounds are all on the path."""
    size = (1 << 32) + (1 << 20) + 21
    n = oracle.generate(size, seed=0xABCD, alphabet=10)
    got = cn.n_to_bits_cuda(n)
    ref = oracle.encode_mt(n, "lut")
    assert np.array_equal(got, ref)
    del ref
    out = np.frombuffer(cn.bits_to_n_cuda(got, size), dtype=np.uint8)
    lut = np.zeros(256, dtype=np.uint8)
    for ch, canon in zip(b"ACGTUacgtu", b"ACGTTACGTT"):
        lut[ch] = canon
    for s in range(0, size, 1 << 28):
        assert np.array_equal(out[s:s + (1 << 28)], lut[n[s:s + (1 << 28)]]), s


# ---------------------------------------------------------------------------------------------------
# encode + validation in one pass
# ---------------------------------------------------------------------------------------------------
def _with_garbage(oracle, size, seed, n_bad):
    rng = np.random.default_rng(seed)
    n = oracle.generate(size, seed=seed, alphabet=10).copy()
    if size and n_bad:
        pos = rng.choice(size, size=min(n_bad, size), replace=False)
        junk = np.frombuffer(b"NnXx-.*0 \n@[`{BbDdHhVvSsWw\x00\x7f\x80\xff\xc3", dtype=np.uint8)
        n[pos] = junk[rng.integers(0, junk.size, size=pos.size)]
    return n


@pytest.mark.parametrize("size,n_bad", [(0, 0), (1, 1), (31, 3), (32, 0), (33, 33), (1000, 17), (40000, 0), (40000, 1),
                                        ((1 << 20) + 17, 5000), (3 * (1 << 20) + 5, 1)])
def test_checked_encode_host(cn, oracle, size, n_bad):
    n = _with_garbage(oracle, size, size + n_bad, n_bad)
    words, invalid = cn.n_to_bits_checked_cuda(n)
    assert invalid == oracle.count_invalid(n)
    assert np.array_equal(words, cn.n_to_bits_cuda(n))                 # same packed words as the unchecked call
    if invalid == 0:
        assert np.array_equal(words, oracle.n_to_bits(n, "lut"))


def test_checked_encode_every_byte_value(cn, oracle):
    """All 256 byte values, each in its own 32-byte word among valid neighbours: exactly 246 are invalid."""
    n = np.full(256 * 32, ord("A"), dtype=np.uint8)
    n[np.arange(256) * 32 + (np.arange(256) % 32)] = np.arange(256, dtype=np.uint8)
    words, invalid = cn.n_to_bits_checked_cuda(n)
    assert invalid == 246 == oracle.count_invalid(n)
    for b in range(256):
        alone = np.full(64, ord("c"), dtype=np.uint8)
        alone[b % 64] = b
        assert cn.n_to_bits_checked_cuda(alone)[1] == (0 if bytes([b]) in [bytes([c]) for c in b"ACGTUacgtu"] else 1), b


@pytest.mark.parametrize("in_off", [0, 3, 16, 21])
def test_checked_encode_device(cn, oracle, torch_cuda, in_off):
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    for size, n_bad in ((5, 2), (4096 + 7, 9), ((1 << 22) + 3, 12345), ((1 << 22), 0)):
        n = _with_garbage(oracle, size + 32, size, n_bad)
        sl = n[in_off: in_off + size]
        d_all = torch.from_numpy(n).cuda()
        d_bits = torch.empty(cn.words_for_len(size), dtype=torch.int64, device="cuda")
        counter = torch.zeros(1, dtype=torch.int64, device="cuda")
        _lib.check(lib.cn_encode_checked_device(d_all.data_ptr() + in_off, size, d_bits.data_ptr(), counter.data_ptr(), st))
        assert int(counter.item()) == oracle.count_invalid(sl)
        plain = torch.empty_like(d_bits)
        _lib.check(lib.cn_encode_device(d_all.data_ptr() + in_off, size, plain.data_ptr(), st))
        assert torch.equal(d_bits, plain)
    # the tensor-level wrapper accumulates across calls
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    n = _with_garbage(oracle, 100000, 5, 77)
    cn.encode_checked_device(torch.from_numpy(n).cuda(), counter)
    cn.encode_checked_device(torch.from_numpy(n).cuda(), counter)
    assert int(counter.item()) == 2 * oracle.count_invalid(n)


# ---------------------------------------------------------------------------------------------------
# error behaviour of the C ABI: status codes + messages, never an abort
# ---------------------------------------------------------------------------------------------------
def test_abi_error_paths(cn, torch_cuda):
    torch = torch_cuda
    import ctypes
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    d = torch.zeros(256, dtype=torch.uint8, device="cuda")
    w = torch.zeros(32, dtype=torch.int64, device="cuda")
    # misaligned packed side
    assert lib.cn_encode_device(d.data_ptr(), 64, w.data_ptr() + 4, st) == _lib.CN_ERR_ARG
    assert b"8-byte aligned" in lib.cn_last_error()
    assert lib.cn_decode_device(w.data_ptr() + 4, 2, 64, d.data_ptr(), st) == _lib.CN_ERR_ARG
    assert lib.cn_encode2_device(d.data_ptr(), 54, w.data_ptr() + 4, st) == _lib.CN_ERR_ARG
    # null pointers with non-zero length
    assert lib.cn_encode_device(None, 64, w.data_ptr(), st) == _lib.CN_ERR_ARG
    assert lib.cn_n_to_bits_host(None, 64, w.data_ptr()) == _lib.CN_ERR_ARG
    assert lib.cn_bits_to_n_host(None, 2, 64, d.data_ptr()) == _lib.CN_ERR_ARG
    assert lib.cn_encode_checked_device(d.data_ptr(), 64, w.data_ptr(), None, st) == _lib.CN_ERR_ARG
    # the panic condition is a status code with the reference's text, for both codecs, host and device
    assert lib.cn_decode_device(w.data_ptr(), 2, 65, d.data_ptr(), st) == _lib.CN_ERR_LENGTH
    assert lib.cn_last_error() == lib.cn_length_panic_message()
    assert lib.cn_decode2_device(w.data_ptr(), 2, 55, d.data_ptr(), st) == _lib.CN_ERR_LENGTH
    assert lib.cn_bits_to_n2_host(w.data_ptr(), 2, 55, d.data_ptr()) == _lib.CN_ERR_LENGTH
    # bad enums / devices
    assert lib.cn_init(99) == _lib.CN_ERR_ARG and b"out of range" in lib.cn_last_error()
    assert lib.cn_generate_device(d.data_ptr(), 0, 16, 1, 7, st) == _lib.CN_ERR_ARG
    assert lib.cn_generate_device(d.data_ptr(), 3, 16, 1, 4, st) == _lib.CN_ERR_ARG
    assert lib.cn_set_host_strategy(2, 0) == _lib.CN_ERR_ARG and lib.cn_set_host_strategy(0, 1000) == _lib.CN_ERR_ARG
    count = ctypes.c_int()
    assert lib.cn_device_count(ctypes.byref(count)) == 0 and count.value >= 1
    # after all those failures the library still works
    assert cn.n_to_bits_cuda(b"ATCG").tolist() == [0xD8]
    torch.cuda.synchronize()


def test_harness_allocators(cn):
    """cn_device_malloc / cn_host_malloc_pinned / cn_memcpy_* / cn_time_*: the helpers the C++ / Rust harnesses use."""
    import ctypes
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    n = (b"ATCGatcgUu" * 1000)[:9999]
    d_n, d_w, h_w = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    words = cn.words_for_len(len(n))
    _lib.check(lib.cn_device_malloc(ctypes.byref(d_n), len(n)))
    _lib.check(lib.cn_device_malloc(ctypes.byref(d_w), words * 8))
    _lib.check(lib.cn_host_malloc_pinned(ctypes.byref(h_w), words * 8))
    _lib.check(lib.cn_memcpy_h2d(d_n, n, len(n), None))
    ms = ctypes.c_float()
    _lib.check(lib.cn_time_encode_device(d_n, len(n), d_w, 3, ctypes.byref(ms)))
    assert ms.value > 0
    _lib.check(lib.cn_memcpy_d2h(h_w, d_w, words * 8, None))
    _lib.check(lib.cn_stream_synchronize(None))
    got = np.ctypeslib.as_array(ctypes.cast(h_w, ctypes.POINTER(ctypes.c_uint64)), shape=(words,)).copy()
    assert np.array_equal(got, cn.n_to_bits_cuda(n))
    _lib.check(lib.cn_time_decode_device(d_w, words, len(n), d_n, 2, ctypes.byref(ms)))
    back = ctypes.create_string_buffer(len(n))
    _lib.check(lib.cn_memcpy_d2h(back, d_n, len(n), None))
    _lib.check(lib.cn_stream_synchronize(None))
    assert back.raw == n.upper().replace(b"U", b"T")
    for p in (d_n, d_w):
        _lib.check(lib.cn_device_free(p))
    _lib.check(lib.cn_host_free_pinned(h_w))


def test_fuzz_sizes_and_alignments(cn, oracle, torch_cuda):
    """300 random (length, source offset, destination offset) triples through the device entry points of both
    codecs, with guard bytes, bit-exact against the oracle.  Seeded: failures reproduce."""
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(20261017)
    pool = oracle.generate(1 << 21, seed=123, alphabet=10)
    pool5 = oracle.generate2(1 << 21, seed=321, alphabet=12)
    d_pool = torch.from_numpy(pool).cuda()
    d_pool5 = torch.from_numpy(pool5).cuda()
    d_words = torch.empty((1 << 17) + 8, dtype=torch.int64, device="cuda")
    d_out = torch.empty((1 << 21) + 128, dtype=torch.uint8, device="cuda")
    for it in range(300):
        scale = [64, 4096, 1 << 17, 1 << 20][it % 4]
        size = int(rng.integers(0, scale))
        src_off = int(rng.integers(0, 64))
        dst_off = int(rng.integers(0, 64))
        base5 = bool(it & 1)
        host = (pool5 if base5 else pool)[src_off: src_off + size]
        d_src = (d_pool5 if base5 else d_pool).data_ptr() + src_off
        ref = oracle.n_to_bits2(host, "lut") if base5 else oracle.n_to_bits(host, "lut")
        d_words.fill_(-1)
        d_out.fill_(0x7E)
        enc = lib.cn_encode2_device if base5 else lib.cn_encode_device
        dec = lib.cn_decode2_device if base5 else lib.cn_decode_device
        _lib.check(enc(d_src, size, d_words.data_ptr(), st))
        got = d_words[: ref.size + 1].cpu().numpy().view(np.uint64)
        assert np.array_equal(got[: ref.size], ref), (it, size, src_off, base5)
        assert got[ref.size] == np.uint64(0xFFFFFFFFFFFFFFFF), (it, size, src_off, base5)
        _lib.check(dec(d_words.data_ptr(), ref.size, size, d_out.data_ptr() + dst_off, st))
        h = d_out[: dst_off + size + 64].cpu().numpy()
        canon = oracle.canonical2(host) if base5 else oracle.canonical(host)
        assert h[dst_off: dst_off + size].tobytes() == canon, (it, size, dst_off, base5)
        assert np.all(h[:dst_off] == 0x7E) and np.all(h[dst_off + size:] == 0x7E), (it, size, dst_off, base5)


def test_device_calls_are_cuda_graph_capturable(cn, oracle, torch_cuda):
    """cn_encode_device / cn_decode_device / cn_encode_checked_device / base-5 pair neither allocate nor synchronise:
    a whole round trip can be captured once into a CUDA graph and replayed on new data (launch-bound small shards)."""
    torch = torch_cuda
    size = 40000 + 13                                             # the reference's bench size, ragged
    d_n = torch.empty(size, dtype=torch.uint8, device="cuda")
    d_bits = torch.empty(cn.words_for_len(size), dtype=torch.int64, device="cuda")
    d_out = torch.empty(size, dtype=torch.uint8, device="cuda")
    d_bits2 = torch.empty(cn.words2_for_len(size), dtype=torch.int64, device="cuda")
    d_out2 = torch.empty(size, dtype=torch.uint8, device="cuda")
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    first = oracle.generate(size, seed=1, alphabet=10)
    d_n.copy_(torch.from_numpy(first))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                                 # warm-up outside capture
        cn.encode_device(d_n, out=d_bits)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        cn.encode_checked_device(d_n, counter, out=d_bits)
        cn.decode_device(d_bits, size, out=d_out)
        cn.encode2_device(d_n, out=d_bits2)
        cn.decode2_device(d_bits2, size, out=d_out2)
    for seed in (2, 3, 4):
        n = oracle.generate(size, seed=seed, alphabet=10)
        n[seed] = ord("N")                                        # one invalid byte per replay
        d_n.copy_(torch.from_numpy(n))
        graph.replay()
        torch.cuda.synchronize()
        assert np.array_equal(d_bits.cpu().numpy().view(np.uint64), cn.n_to_bits_cuda(n))
        assert d_out.cpu().numpy().tobytes() == cn.bits_to_n_cuda(cn.n_to_bits_cuda(n), size)
        assert np.array_equal(d_bits2.cpu().numpy().view(np.uint64), oracle.n_to_bits2(n, "lut"))
        assert d_out2.cpu().numpy().tobytes() == oracle.canonical2(n)
    assert int(counter.item()) == 3
