"""GPU parity tests of the batched small-sequence calls (cn_n_to_bits_host_batch / cn_bits_to_n_host_batch): every
sequence of a batch must come out exactly as the reference's n_to_bits_lut / bits_to_n_lut would produce it on its own
(src/n_to_bits.rs:34-69) -- the shape of the reference's own bench (benches/bench_n_to_bits.rs:10-19, 38-47: one 40 000-nt
string per call), many at a time."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def default_chunks(cn):
    from cute_nucleotides_b200 import _lib
    yield
    _lib.load().cn_set_host_chunks(16 << 20, 4 << 20)


def check_batch(cn, oracle, lens, seed):
    lens = np.asarray(lens, dtype=np.uint64)
    offs = np.zeros(lens.size + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    buf = oracle.generate(int(offs[-1]), seed=seed, alphabet=10)
    words, woffs = cn.n_to_bits_batch_cuda(buf, offs)
    assert woffs[-1] == sum((int(n) + 31) // 32 for n in lens)
    # per-sequence parity on a sample (all of them when the batch is small), against the scalar oracle
    idx = range(lens.size) if lens.size <= 2000 else np.random.default_rng(seed).choice(lens.size, 2000, replace=False)
    for i in idx:
        s, e = int(offs[i]), int(offs[i + 1])
        ref = oracle.n_to_bits(buf[s:e], "lut")
        assert np.array_equal(words[int(woffs[i]): int(woffs[i + 1])], ref), (i, e - s)
    # whole-batch check without sampling: pad every sequence to whole words and encode the concatenation at once
    if lens.size > 2000:
        padded = np.zeros(int(woffs[-1]) * 32, dtype=np.uint8)
        starts = (woffs[:-1] * np.uint64(32)).astype(np.int64)
        pos = np.repeat(starts - offs[:-1].astype(np.int64), lens.astype(np.int64)) + np.arange(int(offs[-1]))
        padded[pos] = buf
        assert np.array_equal(words, oracle.encode_mt(padded, "lut"))
    out, ooffs = cn.bits_to_n_batch_cuda(words, woffs, lens)
    assert np.array_equal(ooffs, offs)
    assert out.tobytes() == oracle.canonical(buf)


def test_batch_basic_shapes(cn, oracle):
    check_batch(cn, oracle, [4], 1)                                           # the reference's tail-only KAT shape
    check_batch(cn, oracle, [32, 4, 0, 33, 1, 31, 64, 65, 0, 150, 151, 40000], 2)
    check_batch(cn, oracle, [40000] * 64, 3)                                  # the reference's bench shape x 64
    check_batch(cn, oracle, [0, 0, 0], 4)
    words, woffs = cn.n_to_bits_batch_cuda(b"", [0])
    assert words.size == 0


def test_batch_kat(cn):
    buf = b"ATCG" * 8 + b"ATCG"                                               # src/n_to_bits.rs:414-416
    words, woffs = cn.n_to_bits_batch_cuda(buf, [0, 32, 36])
    assert [hex(int(w)) for w in words] == ["0xd8d8d8d8d8d8d8d8", "0xd8"]
    out, _ = cn.bits_to_n_batch_cuda(words, woffs, [32, 4])
    assert out.tobytes() == buf
    with pytest.raises(cn.LengthError):
        cn.bits_to_n_batch_cuda(words, woffs, [33, 4])


def test_batch_many_reads(cn, oracle):
    """>= 10^5 sequences of 150 .. 40 000 nt in one call (several chunks, ring wrap-around, ragged last words)."""
    rng = np.random.default_rng(7)
    lens = rng.integers(150, 300, size=120000)
    lens[rng.choice(lens.size, 300, replace=False)] = rng.integers(1000, 40001, size=300)
    check_batch(cn, oracle, lens, 5)


def test_batch_with_sequences_larger_than_a_chunk(cn, oracle):
    from cute_nucleotides_b200 import _lib
    _lib.check(_lib.load().cn_set_host_chunks(0, 64 << 10))                    # 64 KiB chunks: 100 000-nt sequences do not fit
    check_batch(cn, oracle, [100, 100000, 5, 70000, 65536, 65537, 31, 200001, 77], 6)


def test_batch_fans_out_over_the_device_set(cn, oracle):
    """With cn_set_devices the batch is cut into runs of consecutive sequences, one per device (device 0 twice when only
    one GPU is visible: a worker thread plus the caller)."""
    import torch
    devs = list(range(torch.cuda.device_count())) if torch.cuda.device_count() > 1 else [0, 0]
    cn.set_devices(devs)
    try:
        rng = np.random.default_rng(11)
        check_batch(cn, oracle, rng.integers(100, 2000, size=40000), 8)          # ~42 M nt: split in two
        check_batch(cn, oracle, [40000] * 1000, 9)
        check_batch(cn, oracle, [5, 0, 77], 10)                                   # too small to split: runs on devs[0]
    finally:
        cn.set_devices([])


def test_batch_base5(cn, oracle):
    """The same batched calls for the base-5 codec (27 nucleotides per word): per sequence == n_to_bits2_lut / bits_to_n2_lut
    (src/n_to_bits2.rs:37-107), incl. ragged triplets, empty sequences and sequences longer than a staging chunk."""
    from cute_nucleotides_b200 import _lib
    rng = np.random.default_rng(21)
    for lens in ([5, 35, 0, 27, 28, 1, 2, 26, 54, 55, 40000], rng.integers(1, 400, size=30000), [3456 * 3 + 1] * 50):
        lens = np.asarray(lens, dtype=np.uint64)
        offs = np.zeros(lens.size + 1, dtype=np.uint64)
        np.cumsum(lens, out=offs[1:])
        buf = oracle.generate2(int(offs[-1]), seed=int(lens.size), alphabet=12)
        words, woffs = cn.n_to_bits2_batch_cuda(buf, offs)
        for i in (range(lens.size) if lens.size <= 100 else rng.choice(lens.size, 300, replace=False)):
            s, e = int(offs[i]), int(offs[i + 1])
            assert np.array_equal(words[int(woffs[i]): int(woffs[i + 1])], oracle.n_to_bits2(buf[s:e], "lut")), (i, e - s)
        out, ooffs = cn.bits_to_n2_batch_cuda(words, woffs, lens)
        assert np.array_equal(ooffs, offs) and out.tobytes() == oracle.canonical2(buf)
    _lib.check(_lib.load().cn_set_host_chunks(0, 64 << 10))
    lens = np.asarray([100, 100000, 5, 70000, 27 * 2048, 27 * 2048 + 1, 31], dtype=np.uint64)
    offs = np.zeros(lens.size + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    buf = oracle.generate2(int(offs[-1]), seed=3, alphabet=12)
    words, woffs = cn.n_to_bits2_batch_cuda(buf, offs)
    assert np.array_equal(words, np.concatenate([oracle.n_to_bits2(buf[int(offs[i]): int(offs[i + 1])], "lut") for i in range(lens.size)]))
    assert cn.bits_to_n2_batch_cuda(words, woffs, lens)[0].tobytes() == oracle.canonical2(buf)


def test_batch_from_threads(cn, oracle):
    import threading
    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            check_batch(cn, oracle, rng.integers(1, 5000, size=3000), seed)
        except Exception as e:      # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
