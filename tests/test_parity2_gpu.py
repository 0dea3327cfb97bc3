"""GPU parity tests of the base-5 codec (n_to_bits2 / bits_to_n2, src/n_to_bits2.rs), through the C ABI,
bit-exact against the oracle.  Mirrors the reference's tests (src/n_to_bits2.rs:270-299) and adds ragged
lengths (len % 27, len % 3), case/U handling, truncating len, the panic, unaligned buffers, tile boundaries
(3456 nt per warp tile) and a > 4 GiB run."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
    GOLD = json.load(f)

SIZES = [1, 2, 3, 4, 5, 26, 27, 28, 53, 54, 55, 80, 81, 82, 863, 864, 865, 3455, 3456, 3457, 6911, 6912, 6913,
         27647, 27648, 27649, 40000, 100003, (1 << 20) + 5, 3 * (1 << 20) + 11]


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


def test_reference_kat2(cn):
    for case in GOLD["encode2"]:                      # src/n_to_bits2.rs:275-279, 289-293
        got = cn.n_to_bits2_cuda(case["input"].encode())
        assert [str(int(x)) for x in got] == case["words"]
    for case in GOLD["decode2"]:                      # src/n_to_bits2.rs:282-286, 296-298
        words = np.array([int(w) for w in case["words"]], dtype=np.uint64)
        assert cn.bits_to_n2_cuda(words, case["len"]) == case["output"].encode()


def test_length_panic2(cn, torch_cuda):
    torch = torch_cuda
    with pytest.raises(cn.LengthError, match=GOLD["panic_text"]):
        cn.bits_to_n2_cuda(np.zeros(2, dtype=np.uint64), 55)
    assert cn.bits_to_n2_cuda(np.zeros(2, dtype=np.uint64), 54) == b"A" * 54
    with pytest.raises(cn.LengthError, match=GOLD["panic_text"]):
        cn.decode2_device(torch.zeros(2, dtype=torch.int64, device="cuda"), 55)
    assert cn.n_to_bits2_cuda(b"").size == 0 and cn.bits_to_n2_cuda(np.zeros(0, dtype=np.uint64), 0) == b""


@pytest.mark.parametrize("size", SIZES)
def test_host2_vs_oracle(cn, oracle, size):
    n = oracle.generate2(size, seed=size, alphabet=12)
    ref = oracle.n_to_bits2(n, "lut")
    got = cn.n_to_bits2_cuda(n)
    assert np.array_equal(got, ref)
    canon = oracle.canonical2(n)
    assert cn.bits_to_n2_cuda(got, size) == oracle.bits_to_n2(ref, size, "lut") == canon
    for shorter in {size - 1, size // 2, max(size - 29, 0)}:
        assert cn.bits_to_n2_cuda(got, shorter) == canon[:shorter]


@pytest.mark.parametrize("size", SIZES)
def test_device2_vs_oracle(cn, oracle, torch_cuda, size):
    torch = torch_cuda
    n = oracle.generate2(size, seed=size + 1, alphabet=12)
    ref = oracle.n_to_bits2(n, "lut")
    d_bits = cn.encode2_device(torch.from_numpy(n).cuda())
    assert np.array_equal(d_bits.cpu().numpy().view(np.uint64), ref)
    d_out = cn.decode2_device(d_bits, size)
    assert d_out.cpu().numpy().tobytes() == oracle.canonical2(n)


@pytest.mark.parametrize("off", [1, 3, 8, 15])
@pytest.mark.parametrize("size", [5, 27, 100, 3456 + 7, 40000])
def test_device2_unaligned(cn, oracle, torch_cuda, off, size):
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    n = oracle.generate2(size + 64, seed=size * 7 + off, alphabet=12)
    sl = n[off: off + size]
    ref = oracle.n_to_bits2(sl, "lut")
    d_all = torch.from_numpy(n).cuda()
    d_bits = torch.full((ref.size + 1,), -1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.cn_encode2_device(d_all.data_ptr() + off, size, d_bits.data_ptr(), st))
    got = d_bits.cpu().numpy().view(np.uint64)
    assert np.array_equal(got[: ref.size], ref) and got[ref.size] == np.uint64(0xFFFFFFFFFFFFFFFF)
    d_out = torch.full((size + 64,), 0x7E, dtype=torch.uint8, device="cuda")
    _lib.check(lib.cn_decode2_device(d_bits.data_ptr(), ref.size, size, d_out.data_ptr() + off, st))
    h = d_out.cpu().numpy()
    assert h[off: off + size].tobytes() == oracle.canonical2(sl)
    assert np.all(h[:off] == 0x7E) and np.all(h[off + size:] == 0x7E)


def test_device2_guard_bytes_aligned(cn, oracle, torch_cuda):
    """The tiled path must not write past len (ASCII) or past ceil(len/27) words."""
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    for size in (3456, 3456 * 9 + 13, 27648 * 3 + 26):
        n = oracle.generate2(size, seed=size, alphabet=5)
        ref = oracle.n_to_bits2(n, "lut")
        d_n = torch.from_numpy(n).cuda()
        d_bits = torch.full((ref.size + 2,), -1, dtype=torch.int64, device="cuda")
        _lib.check(lib.cn_encode2_device(d_n.data_ptr(), size, d_bits.data_ptr(), st))
        got = d_bits.cpu().numpy().view(np.uint64)
        assert np.array_equal(got[: ref.size], ref) and np.all(got[ref.size:] == np.uint64(0xFFFFFFFFFFFFFFFF))
        d_out = torch.full((size + 64,), 0x7E, dtype=torch.uint8, device="cuda")
        _lib.check(lib.cn_decode2_device(d_bits.data_ptr(), ref.size, size, d_out.data_ptr(), st))
        h = d_out.cpu().numpy()
        assert h[:size].tobytes() == n.tobytes() and np.all(h[size:] == 0x7E)


@pytest.mark.parametrize("strategy,chunk", [(0, 4096), (1, 4096), (0, 1 << 16)])
def test_host2_pipeline_chunking(cn, oracle, strategy, chunk):
    from cute_nucleotides_b200 import _lib
    assert _lib.load().cn_set_host_strategy(strategy, chunk) == 0
    for size in (chunk * 9 + 17, chunk * 3, chunk - 1, 5, (1 << 20) + 3):
        n = oracle.generate2(size, seed=size ^ 0x33, alphabet=12)
        ref = oracle.n_to_bits2(n, "lut")
        got = cn.n_to_bits2_cuda(n)
        assert np.array_equal(got, ref)
        assert cn.bits_to_n2_cuda(got, size) == oracle.canonical2(n)


def test_device2_arbitrary_valid_words(cn, oracle, torch_cuda):
    """Decode of words whose 7-bit fields are all valid (< 125): re-encoding returns the same words."""
    torch = torch_cuda
    rng = np.random.default_rng(11)
    fields = rng.integers(0, 125, size=(50000, 9), dtype=np.uint64)
    words = np.bitwise_or.reduce(fields << (np.arange(9, dtype=np.uint64) * np.uint64(7)), axis=1)
    d_w = torch.from_numpy(words.view(np.int64)).cuda()
    d_out = cn.decode2_device(d_w, 27 * words.size)
    assert d_out.cpu().numpy().tobytes() == oracle.bits_to_n2(words, 27 * words.size, "lut")
    assert torch.equal(cn.encode2_device(d_out), d_w)


def test_device2_generator_matches_host(cn, oracle, torch_cuda):
    torch = torch_cuda
    for alphabet in (5, 12):
        for size, offset in ((1, 0), (100003, 0), (65536, 40001)):
            d = cn.generate2_device(torch.empty(size, dtype=torch.uint8, device="cuda"), offset, 77, alphabet)
            assert np.array_equal(d.cpu().numpy(), oracle.generate2(size, 77, alphabet, offset))


def test_base5_five_gib_properties(cn, oracle, torch_cuda):
    """> 4 GiB (64-bit indexing): sampled windows bit-exact vs the oracle, whole-buffer round trip."""
    torch = torch_cuda
    size = 5 * (1 << 30) + 17
    d_n = cn.generate2_device(torch.empty(size, dtype=torch.uint8, device="cuda"), 0, 99, 12)
    d_bits = cn.encode2_device(d_n)
    assert d_bits.numel() == cn.words2_for_len(size)
    win = 27 * 4096 * 8
    for s in (0, (1 << 32) - win // 2, size - win - 17):
        s -= s % 27
        host = oracle.generate2(min(win, size - s), 99, 12, offset=s)
        assert np.array_equal(d_n[s: s + host.size].cpu().numpy(), host)
        ref = oracle.n_to_bits2(host, "lut")
        assert np.array_equal(d_bits[s // 27: s // 27 + ref.size].cpu().numpy().view(np.uint64), ref), hex(s)
    d_out = cn.decode2_device(d_bits, size)
    lut = torch.zeros(256, dtype=torch.uint8, device="cuda")
    for ch, canon in zip(b"ACGTUNacgtun", b"ACGTTNACGTTN"):
        lut[ch] = canon
    step = 1 << 28
    for s in range(0, size, step):
        e = min(size, s + step)
        assert torch.equal(d_out[s:e], lut[d_n[s:e].long()])
