"""Seeded fuzz of the staged host pipeline (csrc/host_pipeline.cuh): random lengths, staging chunk sizes (so the 8-slot
ring wraps anywhere from never to thousands of times), pinned / pageable sides in every combination, both codecs, all
encode modes, both host strategies and the device fan-out -- every result bit-exact against the oracle.  The first
version of the asynchronous schedule retired chunks it had not issued yet; small chunks with slow copies exposed it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(autouse=True)
def defaults_afterwards(cn):
    from cute_nucleotides_b200 import _lib
    yield
    lib = _lib.load()
    lib.cn_set_host_strategy(0, 16 << 20)
    lib.cn_set_host_chunks(16 << 20, 4 << 20)
    cn.set_devices([])


class Buf:
    """a host buffer of `count` elements, pinned (cudaHostAlloc through torch) or pageable (numpy), optionally misaligned"""

    def __init__(self, torch, dtype, count, pinned, offset=0):
        item = np.dtype(dtype).itemsize
        if pinned:
            self._keep = torch.empty(count * item + 64, dtype=torch.uint8, pin_memory=True)
            raw = self._keep.numpy()
        else:
            raw = np.empty(count * item + 64, dtype=np.uint8)
        off = offset if item == 1 else 8 * (offset % 4)
        self.a = raw[off: off + count * item].view(dtype)
        self.ptr = self.a.ctypes.data


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_host_calls(cn, oracle, torch_cuda, seed):
    import ctypes
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1000 + seed)
    for _ in range(6):
        size = int(rng.choice([rng.integers(1, 300000), rng.integers(300000, 1 << 22), rng.integers(1 << 22, 1 << 25)]))
        chunk = int(rng.choice([4096, 8192, 65536, 1 << 20, 4 << 20])) if size < (1 << 23) else int(rng.choice([65536, 1 << 20, 4 << 20]))
        strategy = int(rng.integers(0, 2))
        _lib.check(lib.cn_set_host_strategy(strategy, chunk))
        cn.set_devices([] if rng.integers(0, 2) else [0, 0])
        base5 = rng.integers(0, 3) == 0
        what = (size, chunk, strategy, bool(base5))
        W = cn.words2_for_len(size) if base5 else cn.words_for_len(size)
        n_buf = Buf(torch, np.uint8, size, bool(rng.integers(0, 2)), int(rng.integers(0, 32)))
        w_buf = Buf(torch, np.uint64, W, bool(rng.integers(0, 2)), int(rng.integers(0, 4)))
        invalid = ctypes.c_uint64(0)
        if base5:
            n_buf.a[:] = oracle.generate2(size, seed=seed + size, alphabet=12)
            _lib.check(lib.cn_n_to_bits2_host(n_buf.ptr, size, w_buf.ptr))
            assert np.array_equal(w_buf.a, oracle.n_to_bits2(n_buf.a, "lut")), ("enc2",) + what
            trunc = size - int(rng.integers(0, min(size, 100)))
            o_buf = Buf(torch, np.uint8, trunc, bool(rng.integers(0, 2)), int(rng.integers(0, 32)))
            _lib.check(lib.cn_bits_to_n2_host(w_buf.ptr, W, trunc, o_buf.ptr))
            assert o_buf.a.tobytes() == oracle.canonical2(n_buf.a[:trunc]), ("dec2",) + what
            continue
        n_buf.a[:] = oracle.generate(size, seed=seed + size, alphabet=10)
        mode = int(rng.integers(0, 3))
        if mode:
            bad = rng.choice(size, size=min(size, int(rng.integers(0, 50))), replace=False)
            n_buf.a[bad] = rng.integers(0, 256, size=bad.size, dtype=np.uint8)
        _lib.check(lib.cn_n_to_bits_ex_host(n_buf.ptr, size, w_buf.ptr, mode, ctypes.byref(invalid)))
        lut = oracle.encode_mt(n_buf.a.copy(), "lut")
        if mode == cn.ENC_LUT_EXACT or invalid.value == 0:
            assert np.array_equal(w_buf.a, lut), ("enc", mode) + what
        if mode:
            assert invalid.value == oracle.count_invalid(n_buf.a), ("count", mode) + what
        w_buf.a[:] = lut
        trunc = size - int(rng.integers(0, min(size, 100)))
        o_buf = Buf(torch, np.uint8, trunc, bool(rng.integers(0, 2)), int(rng.integers(0, 32)))
        _lib.check(lib.cn_bits_to_n_host(w_buf.ptr, W, trunc, o_buf.ptr))
        assert o_buf.a.tobytes() == oracle.bits_to_n(lut, trunc, "lut"), ("dec",) + what
