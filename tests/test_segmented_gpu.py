"""GPU parity tests of the device-resident batch (cn_encode_segmented_device / cn_decode_segmented_device): many
sequences tightly concatenated in HBM; every sequence must come out exactly as n_to_bits_lut / bits_to_n_lut would
produce it on its own (src/n_to_bits.rs:34-69)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


def expected_words(oracle, buf, lens):
    """every sequence padded to whole words with zero bytes (code 0), then ONE oracle encode of the concatenation"""
    lens = np.asarray(lens, dtype=np.int64)
    words = (lens + 31) >> 5
    woff = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(words, out=woff[1:])
    off = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    padded = np.zeros(int(woff[-1]) * 32, dtype=np.uint8)
    if buf.size:
        pos = np.repeat(woff[:-1] * 32 - off[:-1], lens) + np.arange(int(off[-1]))
        padded[pos] = buf
    return oracle.encode_mt(padded, "lut") if padded.size else np.zeros(0, dtype=np.uint64), off, woff


def run_case(cn, oracle, torch, lens, seed, in_shift=0, out_shift=0):
    lens = np.asarray(lens, dtype=np.int64)
    total = int(lens.sum())
    buf = oracle.generate(total, seed=seed, alphabet=10) if total else np.zeros(0, dtype=np.uint8)
    ref, off, woff = expected_words(oracle, buf, lens)
    raw = torch.full((total + 64,), 0xFF, dtype=torch.uint8, device="cuda")           # poison around the buffer
    raw[in_shift: in_shift + total] = torch.from_numpy(buf).cuda()
    d_n = raw[in_shift: in_shift + total]
    d_off = torch.from_numpy(off).cuda()
    out = torch.full((int(woff[-1]) + 2,), -1, dtype=torch.int64, device="cuda")
    words, d_woff = cn.encode_segmented_device(d_n, d_off, out=out[1:-1] if woff[-1] else out[1:1])
    torch.cuda.synchronize()
    assert np.array_equal(d_woff.cpu().numpy(), woff)
    assert np.array_equal(words.cpu().numpy().view(np.uint64), ref), (lens[:8], in_shift)
    assert int(out[0].item()) == -1 and int(out[-1].item()) == -1                      # guard words untouched
    back_raw = torch.full((total + 64,), 0x5A, dtype=torch.uint8, device="cuda")
    back = cn.decode_segmented_device(words, d_off, d_woff, out=back_raw[out_shift: out_shift + total])
    torch.cuda.synchronize()
    assert back.cpu().numpy().tobytes() == oracle.canonical(buf), (lens[:8], out_shift)
    assert bool((back_raw[:out_shift] == 0x5A).all()) and bool((back_raw[out_shift + total:] == 0x5A).all())
    # a handful of sequences individually against the scalar oracle
    for i in np.random.default_rng(seed).choice(lens.size, size=min(20, lens.size), replace=False):
        s, e = int(off[i]), int(off[i + 1])
        assert np.array_equal(ref[int(woff[i]): int(woff[i + 1])], oracle.n_to_bits(buf[s:e], "lut"))


def test_segmented_kat(cn, torch_cuda):
    torch = torch_cuda
    buf = torch.tensor(list(b"ATCG" * 8 + b"ATCG"), dtype=torch.uint8, device="cuda")   # src/n_to_bits.rs:414-416
    off = torch.tensor([0, 32, 36], dtype=torch.int64, device="cuda")
    words, woff = cn.encode_segmented_device(buf, off)
    assert [hex(int(w) & (2 ** 64 - 1)) for w in words.cpu()] == ["0xd8d8d8d8d8d8d8d8", "0xd8"] and woff.tolist() == [0, 1, 2]
    assert cn.decode_segmented_device(words, off, woff).cpu().numpy().tobytes() == b"ATCG" * 9


@pytest.mark.parametrize("in_shift,out_shift", [(0, 0), (1, 3), (7, 16), (15, 5), (16, 31)])
def test_segmented_shapes_and_alignment(cn, oracle, torch_cuda, in_shift, out_shift):
    torch = torch_cuda
    rng = np.random.default_rng(in_shift * 100 + out_shift)
    run_case(cn, oracle, torch, [4], 1, in_shift, out_shift)
    run_case(cn, oracle, torch, [32, 4, 0, 33, 1, 31, 64, 65, 0, 0, 150, 151, 40000, 0], 2, in_shift, out_shift)
    run_case(cn, oracle, torch, rng.integers(1, 40, size=5000), 3, in_shift, out_shift)          # many sequences per CTA
    run_case(cn, oracle, torch, rng.integers(0, 3, size=3000), 4, in_shift, out_shift)           # mostly empty / 1-2 nt
    run_case(cn, oracle, torch, rng.integers(150, 301, size=20000), 5, in_shift, out_shift)      # short reads
    run_case(cn, oracle, torch, [8192 * 3 + 5], 6, in_shift, out_shift)                          # one long sequence
    run_case(cn, oracle, torch, [8191, 8193, 8192, 1, 8192 + 31], 7, in_shift, out_shift)        # CTA-sized pieces


def test_segmented_large(cn, oracle, torch_cuda):
    rng = np.random.default_rng(9)
    lens = rng.integers(100, 400, size=400000)
    lens[rng.choice(lens.size, 200, replace=False)] = rng.integers(10000, 100000, size=200)
    run_case(cn, oracle, torch_cuda, lens, 8, 3, 9)


def test_segmented_argument_errors(cn, torch_cuda):
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    d = torch.zeros(64, dtype=torch.uint8, device="cuda")
    off = torch.tensor([0, 64], dtype=torch.int64, device="cuda")
    woff = torch.tensor([0, 2], dtype=torch.int64, device="cuda")
    o = torch.zeros(2, dtype=torch.int64, device="cuda")
    assert lib.cn_encode_segmented_device(d.data_ptr(), None, woff.data_ptr(), 1, 64, 2, o.data_ptr(), None) == _lib.CN_ERR_ARG
    assert lib.cn_encode_segmented_device(d.data_ptr(), off.data_ptr(), woff.data_ptr(), 1, 65, 2, o.data_ptr(), None) == _lib.CN_ERR_LENGTH
    assert lib.cn_encode_segmented_device(d.data_ptr(), off.data_ptr(), woff.data_ptr(), 0, 0, 0, o.data_ptr(), None) == 0
