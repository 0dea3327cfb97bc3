"""GPU tests of the asynchronous host calls (cn_n_to_bits_host_async / cn_bits_to_n_host_async / cn_wait): same results as
the synchronous calls, several requests in flight at once, statuses and the reference's panic delivered correctly."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_async_matches_sync(cn, oracle):
    for size in (0, 1, 33, 40000, (1 << 20) + 17, (1 << 25) + 5):
        n = oracle.generate(size, seed=size + 1, alphabet=10)
        words = cn.n_to_bits_cuda_async(n).wait()
        assert np.array_equal(words, oracle.encode_mt(n, "lut") if size else np.zeros(0, dtype=np.uint64))
        assert cn.bits_to_n_cuda_async(words, size).wait() == oracle.canonical(n)


def test_several_requests_in_flight(cn, oracle):
    """encode of batch k+1 overlapped with decode of batch k from ONE thread -- the duplex schedule of bench.py"""
    batches = [oracle.generate((1 << 24) + 1000 * k, seed=k, alphabet=10) for k in range(4)]
    pending = cn.n_to_bits_cuda_async(batches[0])
    decoded = []
    for k in range(len(batches)):
        words = pending.wait()
        if k + 1 < len(batches):
            pending = cn.n_to_bits_cuda_async(batches[k + 1])
        decoded.append(cn.bits_to_n_cuda(words, batches[k].size))
    for n, d in zip(batches, decoded):
        assert d == oracle.canonical(n)
    # six at once (more than the pool has threads): they queue, none is lost
    reqs = [cn.n_to_bits_cuda_async(b) for b in batches + batches[:2]]
    for r, b in zip(reqs, batches + batches[:2]):
        assert np.array_equal(r.wait(), oracle.encode_mt(b, "lut"))


def test_async_errors(cn):
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    with pytest.raises(cn.LengthError):                       # the reference's panic, raised at submission
        cn.bits_to_n_cuda_async(np.zeros(2, dtype=np.uint64), 65)
    req = ctypes.c_void_p()
    buf = np.zeros(64, dtype=np.uint8)
    assert lib.cn_n_to_bits_host_async(buf.ctypes.data, 64, None, ctypes.byref(req)) == _lib.CN_ERR_ARG and not req.value
    assert lib.cn_wait(None) == _lib.CN_ERR_ARG
    r = cn.n_to_bits_cuda_async(b"ATCG" * 8)
    assert [hex(int(w)) for w in r.wait()] == ["0xd8d8d8d8d8d8d8d8"]
    with pytest.raises(RuntimeError):
        r.wait()
