"""GPU tests of the operations on packed words (cn_hamming_*, cn_complement_*, cn_reverse_complement_*).

PARITY UNPINNED: the reference has no code for these (README.md:21-23, 45, 415-418 only motivate them).  The checker
is the definition through the reference's own codec, computed via ASCII by the oracle (decode with bits_to_n_lut,
operate on letters, re-encode with n_to_bits_lut), plus size-independent properties: complement and reverse complement
are involutions, hamming(a, a) == 0, hamming(a, complement(a)) == len, reverse complement commutes with the codec."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 160, 161, 255, 256, 257, 1000, 4095, 4096, 4097, 32 * 1026, 32 * 1027 - 5,
         32 * 1031 - 1, 32 * 1029, 40000, (1 << 20) + 17, 3 * (1 << 20) + 5]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


def mutate(oracle, n, seed, frac):
    rng = np.random.default_rng(seed)
    m = n.copy()
    if n.size:
        pos = rng.choice(n.size, size=max(1, int(n.size * frac)), replace=False)
        m[pos] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=pos.size)]
    return m


@pytest.mark.parametrize("size", SIZES)
def test_packed_ops_host_vs_definition(cn, oracle, size):
    n = oracle.generate(size, seed=size + 3, alphabet=10)
    m = mutate(oracle, n, size, 0.1)
    a, b = oracle.n_to_bits(n, "lut"), oracle.n_to_bits(m, "lut")
    assert cn.hamming_cuda(a, b, size) == oracle.hamming(a, b, size) == int(np.count_nonzero(
        np.frombuffer(oracle.canonical(n), dtype=np.uint8) != np.frombuffer(oracle.canonical(m), dtype=np.uint8)))
    assert cn.hamming_cuda(a, a, size) == 0
    comp = cn.complement_cuda(a, size)
    assert np.array_equal(comp, oracle.complement(a, size))
    assert cn.hamming_cuda(a, comp, size) == size                       # every nucleotide differs from its complement
    assert np.array_equal(cn.complement_cuda(comp, size), a)            # involution
    rc = cn.reverse_complement_cuda(a, size)
    assert np.array_equal(rc, oracle.complement(a, size, reverse=True))
    assert np.array_equal(cn.reverse_complement_cuda(rc, size), a)      # involution
    # through ASCII with the product's own codec: revcomp(encode(x)) == encode(revcomp_ascii(x))
    table = np.zeros(256, dtype=np.uint8)
    for ch, co in zip(b"ACGT", b"TGCA"):
        table[ch] = co
    ascii_rc = table[np.frombuffer(oracle.canonical(n), dtype=np.uint8)][::-1].copy()
    assert np.array_equal(rc, cn.n_to_bits_cuda(ascii_rc))


def test_only_the_first_len_nucleotides_count(cn, oracle):
    """Truncating len: garbage in the words' unused fields and in trailing words is ignored and never copied out."""
    size = 1000
    a = oracle.generate_words(40, seed=1)
    b = oracle.generate_words(40, seed=2)
    for length in (0, 1, 31, 32, 33, 999, 1000, 1279, 1280):
        assert cn.hamming_cuda(a, b, length) == oracle.hamming(a, b, length)
        assert np.array_equal(cn.complement_cuda(a, length), oracle.complement(a, length))
        assert np.array_equal(cn.reverse_complement_cuda(a, length), oracle.complement(a, length, reverse=True))
    with pytest.raises(cn.LengthError):
        cn.hamming_cuda(a, b, 40 * 32 + 1)
    with pytest.raises(cn.LengthError):
        cn.reverse_complement_cuda(a, 40 * 32 + 1)


@pytest.mark.parametrize("word_off", [0, 1, 3])
def test_packed_ops_device_alignment_and_guards(cn, oracle, torch_cuda, word_off):
    torch = torch_cuda
    for size in (5, 129, 4096 * 32 + 7, (1 << 22) + 31, 1 << 22):
        W = cn.words_for_len(size)
        wa, wb = oracle.generate_words(W, seed=size), oracle.generate_words(W, seed=size + 1)
        if size % 32:
            keep = np.uint64((1 << (2 * (size % 32))) - 1)
            wa[-1] &= keep
            wb[-1] &= keep
        da = torch.zeros(W + word_off + 1, dtype=torch.int64, device="cuda")
        db = torch.zeros(W + word_off + 1, dtype=torch.int64, device="cuda")
        da[word_off: word_off + W] = torch.from_numpy(wa.view(np.int64)).cuda()
        db[word_off: word_off + W] = torch.from_numpy(wb.view(np.int64)).cuda()
        va, vb = da[word_off: word_off + W], db[word_off: word_off + W]
        res = cn.hamming_device(va, vb, size)
        cn.hamming_device(va, vb, size, result=res)                     # accumulates
        assert int(res.item()) == 2 * oracle.hamming(wa, wb, size)
        out = torch.full((W + word_off + 1,), -1, dtype=torch.int64, device="cuda")
        cn.complement_device(va, size, out=out[word_off: word_off + W])
        assert np.array_equal(out[word_off: word_off + W].cpu().numpy().view(np.uint64), oracle.complement(wa, size))
        assert int(out[-1].item()) == -1 and (word_off == 0 or int(out[0].item()) == -1)
        out.fill_(-1)
        cn.reverse_complement_device(va, size, out=out[word_off: word_off + W])
        assert np.array_equal(out[word_off: word_off + W].cpu().numpy().view(np.uint64), oracle.complement(wa, size, reverse=True))
        assert int(out[-1].item()) == -1 and (word_off == 0 or int(out[0].item()) == -1)
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    assert lib.cn_reverse_complement_device(da.data_ptr(), 4, 100, da.data_ptr(), None) == _lib.CN_ERR_ARG   # in place
    assert lib.cn_hamming_device(da.data_ptr(), db.data_ptr(), 4, 100, None, None) == _lib.CN_ERR_ARG


def test_packed_ops_at_scale_properties(cn, oracle, torch_cuda):
    """2^33 + 21 nucleotides (2 GiB of packed words per operand): properties that need no host reference."""
    torch = torch_cuda
    size = (1 << 33) + 21
    W = cn.words_for_len(size)
    a = cn.generate_words_device(torch.empty(W, dtype=torch.int64, device="cuda"), 0, 11)
    a[-1] &= (1 << 42) - 1
    comp = cn.complement_device(a, size)
    assert int(cn.hamming_device(a, comp, size).item()) == size
    assert int(cn.hamming_device(a, a, size).item()) == 0
    rc = cn.reverse_complement_device(a, size)
    back = cn.reverse_complement_device(rc, size)
    assert torch.equal(back, a)
    # a window of the reverse complement against the definition
    win = 1 << 20
    tail = oracle.complement(a[-(win // 32 + 1):].cpu().numpy().view(np.uint64), win + 21, reverse=True)
    assert np.array_equal(rc[: win // 32].cpu().numpy().view(np.uint64), tail[: win // 32])
    del comp, rc, back


def test_round2_device_calls_are_cuda_graph_capturable(cn, oracle, torch_cuda):
    """The round-2 device entry points (_lut-exact encode, base-5 validating encode, Hamming, complement, reverse
    complement) neither allocate nor synchronise either: captured once, replayed on new data."""
    torch = torch_cuda
    size = 40000 + 13
    W = cn.words_for_len(size)
    d_n = torch.empty(size, dtype=torch.uint8, device="cuda")
    d_bits = torch.empty(W, dtype=torch.int64, device="cuda")
    d_comp = torch.empty(W, dtype=torch.int64, device="cuda")
    d_rc = torch.empty(W, dtype=torch.int64, device="cuda")
    d_bits2 = torch.empty(cn.words2_for_len(size), dtype=torch.int64, device="cuda")
    counters = torch.zeros(3, dtype=torch.int64, device="cuda")
    d_n.copy_(torch.from_numpy(oracle.generate(size, seed=1, alphabet=10)))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                                 # warm-up outside capture
        cn.encode_ex_device(d_n, cn.ENC_LUT_EXACT, out=d_bits)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        cn.encode_ex_device(d_n, cn.ENC_LUT_EXACT, counter=counters[0:1], out=d_bits)
        cn.encode2_ex_device(d_n, cn.ENC_COUNT, counter=counters[1:2], out=d_bits2)
        cn.complement_device(d_bits, size, out=d_comp)
        cn.reverse_complement_device(d_bits, size, out=d_rc)
        cn.hamming_device(d_bits, d_comp, size, result=counters[2:3])
    for seed in (2, 3, 4):
        n = oracle.generate(size, seed=seed, alphabet=10)
        n[seed * 1000] = ord("N")                                 # off the 2-bit alphabet, on the base-5 one
        n[seed * 1000 + 1] = ord("-")                             # off both
        d_n.copy_(torch.from_numpy(n))
        graph.replay()
        torch.cuda.synchronize()
        ref = oracle.n_to_bits(n, "lut")
        assert np.array_equal(d_bits.cpu().numpy().view(np.uint64), ref)
        assert np.array_equal(d_bits2.cpu().numpy().view(np.uint64), cn.n_to_bits2_cuda(n))
        assert np.array_equal(d_comp.cpu().numpy().view(np.uint64), oracle.complement(ref, size))
        assert np.array_equal(d_rc.cpu().numpy().view(np.uint64), oracle.complement(ref, size, reverse=True))
    assert counters.tolist() == [6, 3, 3 * size]
