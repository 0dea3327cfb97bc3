"""GPU parity tests of the encode MODES (include/cute_nucleotides_cuda.h, CN_ENC_*), through the C ABI.

The reference's variants disagree on bytes outside the alphabet: n_to_bits_lut maps them to 0 (BYTE_LUT,
src/n_to_bits.rs:8-21, :42), the SIMD encoders to (b >> 1) & 3 (:85, :100).  Mode CN_ENC_LUT_EXACT must be
bit-exact against n_to_bits_lut -- the oracle north_star names -- on EVERY input; mode CN_ENC_PLAIN against
n_to_bits_mul on every input; CN_ENC_COUNT adds the exact count.  Same for the base-5 codec
(n_to_bits2_lut, src/n_to_bits2.rs:8-23, vs the low-3-bit LUT of n_to_bits2_pext, :127-136)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

JUNK = np.frombuffer(b"NnXx-.*0 \n@[`{BbDdHhVvSsWw\x00\x7f\x80\xff\xc3RYKMrykm>", dtype=np.uint8)
SIZES = [0, 1, 5, 31, 32, 33, 64, 65, 1000, 4095, 4097, 40000, 65536 + 17, (1 << 20) + 17, 3 * (1 << 20) + 5]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(autouse=True)
def default_host_strategy(cn):
    from cute_nucleotides_b200 import _lib
    yield
    _lib.load().cn_set_host_strategy(0, 16 << 20)
    _lib.load().cn_set_host_chunks(16 << 20, 4 << 20)


def dirty(gen, size, seed, frac):
    """valid sequence with a fraction `frac` of junk bytes (frac >= 1: every byte uniformly random)"""
    rng = np.random.default_rng(seed)
    if frac >= 1:
        return rng.integers(0, 256, size=size, dtype=np.uint8)
    n = gen(size).copy()
    if size and frac > 0:
        k = max(1, int(size * frac))
        pos = rng.choice(size, size=min(k, size), replace=False)
        n[pos] = JUNK[rng.integers(0, JUNK.size, size=pos.size)]
    return n


def simd_body_expect(n):
    """(b >> 1) & 3 for EVERY byte -- the body of the reference's SIMD encoders (src/n_to_bits.rs:85, :100) applied to the
    whole input (their <32-byte tail goes through the LUT instead, which is why this is not simply oracle 'mul')."""
    codes = ((n >> 1) & 3).astype(np.uint64)
    codes = np.concatenate([codes, np.zeros((-n.size) % 32, dtype=np.uint64)]).reshape(-1, 32)
    return (codes << (np.arange(32, dtype=np.uint64) * np.uint64(2))).sum(axis=1, dtype=np.uint64)


# ---------------------------------------------------------------------------------------------------
# 2-bit codec
# ---------------------------------------------------------------------------------------------------
def test_lut_exact_every_byte_value(cn, oracle):
    """Each of the 256 byte values at each position of a 32-nt word: n_to_bits_lut says 0 for all but the 10 letters."""
    for pos in range(32):
        n = np.full(256 * 32, ord("G"), dtype=np.uint8)
        n[np.arange(256) * 32 + pos] = np.arange(256, dtype=np.uint8)
        words, invalid = cn.n_to_bits_ex_cuda(n, cn.ENC_LUT_EXACT)
        assert np.array_equal(words, oracle.n_to_bits(n, "lut")), pos
        assert invalid == 246 == oracle.count_invalid(n)
    assert np.array_equal(cn.n_to_bits_lut_cuda(b"ACGTNNNNacgtnnnn"), oracle.n_to_bits(b"ACGTNNNNacgtnnnn", "lut"))
    assert np.array_equal(cn.n_to_bits_cuda(b"ACGTNNNNacgtnnnn"), simd_body_expect(np.frombuffer(b"ACGTNNNNacgtnnnn", dtype=np.uint8)))


@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("frac", [0.0, 0.001, 0.3, 1.0])
def test_lut_exact_host(cn, oracle, size, frac):
    n = dirty(lambda s: oracle.generate(s, seed=size + 1, alphabet=10), size, size, frac)
    words, invalid = cn.n_to_bits_ex_cuda(n, cn.ENC_LUT_EXACT)
    assert np.array_equal(words, oracle.n_to_bits(n, "lut"))
    assert invalid == oracle.count_invalid(n)
    plain, zero = cn.n_to_bits_ex_cuda(n, cn.ENC_PLAIN)
    assert zero == 0 and np.array_equal(plain, simd_body_expect(n))              # the SIMD encoders' behaviour, on any bytes
    if size % 32 == 0 and oracle.simd_ok:
        assert np.array_equal(plain, oracle.n_to_bits(n, "mul"))
    counted, invalid2 = cn.n_to_bits_ex_cuda(n, cn.ENC_COUNT)
    assert invalid2 == invalid and np.array_equal(counted, plain)


def test_lut_exact_chunked_ring(cn, oracle):
    """4 KiB chunks: every slot of the ring is reused many times, per-slot counters are summed."""
    from cute_nucleotides_b200 import _lib
    _lib.check(_lib.load().cn_set_host_strategy(0, 4096))
    n = dirty(lambda s: oracle.generate(s, seed=9, alphabet=10), (1 << 20) + 333, 77, 0.01)
    words, invalid = cn.n_to_bits_ex_cuda(n, cn.ENC_LUT_EXACT)
    assert np.array_equal(words, oracle.n_to_bits(n, "lut")) and invalid == oracle.count_invalid(n)


@pytest.mark.parametrize("in_off", [0, 1, 3, 15, 16, 21, 31])
def test_lut_exact_device_any_alignment(cn, oracle, torch_cuda, in_off):
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    for size, frac in ((7, 0.5), (4096 + 7, 0.01), ((1 << 22) + 3, 0.002), ((1 << 22), 0.0), (100003, 1.0)):
        n = dirty(lambda s: oracle.generate(s, seed=size, alphabet=10), size + 64, size + in_off, frac)
        sl = n[in_off: in_off + size]
        d_all = torch.from_numpy(n).cuda()
        d_bits = torch.full((cn.words_for_len(size) + 1,), -1, dtype=torch.int64, device="cuda")
        counter = torch.zeros(1, dtype=torch.int64, device="cuda")
        _lib.check(lib.cn_encode_ex_device(d_all.data_ptr() + in_off, size, d_bits.data_ptr(), cn.ENC_LUT_EXACT, counter.data_ptr(), st))
        assert np.array_equal(d_bits[:-1].cpu().numpy().view(np.uint64), oracle.n_to_bits(sl, "lut"))
        assert int(d_bits[-1].item()) == -1                                   # guard word untouched
        assert int(counter.item()) == oracle.count_invalid(sl)
        # the counter is optional in this mode
        d2 = torch.empty(cn.words_for_len(size), dtype=torch.int64, device="cuda")
        _lib.check(lib.cn_encode_ex_device(d_all.data_ptr() + in_off, size, d2.data_ptr(), cn.ENC_LUT_EXACT, None, st))
        assert torch.equal(d2, d_bits[:-1])
    # tensor-level wrapper
    n = dirty(lambda s: oracle.generate(s, seed=4, alphabet=10), 100000, 5, 0.05)
    got = cn.encode_ex_device(torch.from_numpy(n).cuda(), cn.ENC_LUT_EXACT)
    assert np.array_equal(got.cpu().numpy().view(np.uint64), oracle.n_to_bits(n, "lut"))


def test_misaligned_encode_reads_nothing_before_the_buffer(cn, oracle, torch_cuda):
    """The buffer starts `off` bytes into a fresh allocation whose preceding bytes are poison (0xFF would encode as 3):
    results must only depend on the caller's bytes.  (compute-sanitizer runs of this test check the loads themselves.)"""
    torch = torch_cuda
    for off in (1, 7, 15):
        size = 5000
        n = oracle.generate(size, seed=off, alphabet=10)
        raw = torch.full((size + 64,), 0xFF, dtype=torch.uint8, device="cuda")
        raw[off: off + size] = torch.from_numpy(n).cuda()
        got = cn.encode_device(raw[off: off + size])
        assert np.array_equal(got.cpu().numpy().view(np.uint64), oracle.n_to_bits(n, "lut"))
        m = oracle.generate2(size, seed=off, alphabet=12)
        raw[off: off + size] = torch.from_numpy(m).cuda()
        got2 = cn.encode2_device(raw[off: off + size])
        assert np.array_equal(got2.cpu().numpy().view(np.uint64), oracle.n_to_bits2(m, "lut"))


def test_mode_argument_errors(cn, torch_cuda):
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    d = torch.zeros(64, dtype=torch.uint8, device="cuda")
    o = torch.zeros(2, dtype=torch.int64, device="cuda")
    assert lib.cn_encode_ex_device(d.data_ptr(), 64, o.data_ptr(), 7, None, None) == _lib.CN_ERR_ARG
    assert lib.cn_encode_ex_device(d.data_ptr(), 64, o.data_ptr(), cn.ENC_COUNT, None, None) == _lib.CN_ERR_ARG
    assert lib.cn_encode2_ex_device(d.data_ptr(), 64, o.data_ptr(), cn.ENC_COUNT, None, None) == _lib.CN_ERR_ARG
    buf = np.zeros(64, dtype=np.uint8)
    out = np.zeros(2, dtype=np.uint64)
    assert lib.cn_n_to_bits_ex_host(buf.ctypes.data, 64, out.ctypes.data, 9, None) == _lib.CN_ERR_ARG
    assert lib.cn_n_to_bits_ex_host(buf.ctypes.data, 64, out.ctypes.data, cn.ENC_COUNT, None) == _lib.CN_ERR_ARG


# ---------------------------------------------------------------------------------------------------
# base-5 codec
# ---------------------------------------------------------------------------------------------------
SIZES2 = [0, 1, 2, 3, 26, 27, 28, 80, 81, 82, 864, 3455, 3456, 3457, 27649, 100003, (1 << 20) + 5, 3 * (1 << 20) + 11]


def test_base5_every_byte_value(cn, oracle):
    for pos in (0, 1, 2, 13, 26):
        n = np.full(256 * 27, ord("N"), dtype=np.uint8)
        n[np.arange(256) * 27 + pos] = np.arange(256, dtype=np.uint8)
        words, invalid = cn.n_to_bits2_ex_cuda(n, cn.ENC_LUT_EXACT)
        assert np.array_equal(words, oracle.n_to_bits2(n, "lut")), pos
        assert invalid == 244 == oracle.count_invalid2(n)
        counted, invalid2 = cn.n_to_bits2_ex_cuda(n, cn.ENC_COUNT)
        assert invalid2 == 244 and np.array_equal(counted, cn.n_to_bits2_cuda(n))


@pytest.mark.parametrize("size", SIZES2)
@pytest.mark.parametrize("frac", [0.0, 0.001, 0.3, 1.0])
def test_base5_modes_host(cn, oracle, size, frac):
    n = dirty(lambda s: oracle.generate2(s, seed=size + 2, alphabet=12), size, size + 5, frac)
    words, invalid = cn.n_to_bits2_ex_cuda(n, cn.ENC_LUT_EXACT)
    assert np.array_equal(words, oracle.n_to_bits2(n, "lut"))
    assert invalid == oracle.count_invalid2(n)
    counted, invalid2 = cn.n_to_bits2_ex_cuda(n, cn.ENC_COUNT)
    assert invalid2 == invalid and np.array_equal(counted, cn.n_to_bits2_cuda(n))
    if invalid == 0:
        assert np.array_equal(counted, words)


@pytest.mark.parametrize("in_off", [0, 1, 5, 16, 23])
def test_base5_modes_device_any_alignment(cn, oracle, torch_cuda, in_off):
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    for size, frac in ((30, 0.5), (3456 * 3 + 11, 0.01), ((1 << 22) + 3, 0.002), (3456 * 40, 0.0), (100003, 1.0)):
        n = dirty(lambda s: oracle.generate2(s, seed=size, alphabet=12), size + 64, size + in_off, frac)
        sl = n[in_off: in_off + size]
        d_all = torch.from_numpy(n).cuda()
        for out_off in (0, 1):                                   # packed side 32-byte aligned / only 8-byte aligned (scalar kernel)
            d_bits = torch.full((cn.words2_for_len(size) + 2,), -1, dtype=torch.int64, device="cuda")
            counter = torch.zeros(1, dtype=torch.int64, device="cuda")
            _lib.check(lib.cn_encode2_ex_device(d_all.data_ptr() + in_off, size, d_bits.data_ptr() + 8 * out_off, cn.ENC_LUT_EXACT,
                                                counter.data_ptr(), st))
            got = d_bits[out_off: out_off + cn.words2_for_len(size)].cpu().numpy().view(np.uint64)
            assert np.array_equal(got, oracle.n_to_bits2(sl, "lut"))
            assert int(d_bits[-1].item()) == -1
            assert int(counter.item()) == oracle.count_invalid2(sl)
            counter.zero_()
            _lib.check(lib.cn_encode2_ex_device(d_all.data_ptr() + in_off, size, d_bits.data_ptr() + 8 * out_off, cn.ENC_COUNT,
                                                counter.data_ptr(), st))
            assert int(counter.item()) == oracle.count_invalid2(sl)
            plain = cn.encode2_device(d_all[in_off: in_off + size])
            assert torch.equal(d_bits[out_off: out_off + cn.words2_for_len(size)], plain)
