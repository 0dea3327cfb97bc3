"""GPU tests of multi-GPU inside ONE process (include/cute_nucleotides_cuda.h: cn_set_devices, cn_encode_sharded,
cn_decode_sharded, cn_enable_peer_access), bit-exact against the oracle.

With one visible GPU the same machinery is exercised by listing device 0 twice (a worker thread plus the calling
thread, two staging rings on one GPU); with >= 2 GPUs every listed device is distinct and the peer-memory paths
(encode + gather-to-root, scatter + decode in one kernel over NVLink) run as well."""
import ctypes
import threading

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
def single_device_afterwards(cn):
    yield
    cn.set_devices([])
    from cute_nucleotides_b200 import _lib
    _lib.load().cn_set_host_strategy(0, 16 << 20)
    _lib.load().cn_set_host_chunks(16 << 20, 4 << 20)


def device_lists(torch):
    n = torch.cuda.device_count()
    lists = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [list(range(n)), [1, 0]]
    return lists


def test_set_get_devices(cn, torch_cuda):
    from cute_nucleotides_b200 import _lib
    assert cn.get_devices() == []
    cn.set_devices([0, 0])
    assert cn.get_devices() == [0, 0]
    cn.set_devices([])
    assert cn.get_devices() == []
    bad = (ctypes.c_int * 1)(torch_cuda.cuda.device_count())
    assert _lib.load().cn_set_devices(bad, 1) == _lib.CN_ERR_ARG


@pytest.mark.parametrize("pinned", [False, True])
def test_host_calls_fan_out(cn, oracle, torch_cuda, pinned):
    """One cn_n_to_bits_host / cn_bits_to_n_host call split over the device list == the oracle, for sizes below and
    above the fan-out threshold (8 Mi nt per device), ragged ends, both codecs and the counted modes."""
    torch = torch_cuda
    for devs in device_lists(torch):
        cn.set_devices(devs)
        for size in (1000, (1 << 23) + 5, 3 * (1 << 23) + 12345, (1 << 26) + 77):
            n = oracle.generate(size, seed=size % 1000, alphabet=10)
            if pinned:
                h = torch.empty(size, dtype=torch.uint8).pin_memory()
                h.numpy()[:] = n
                src = h.numpy()
            else:
                src = n
            ref = oracle.encode_mt(n, "lut")
            got = cn.n_to_bits_cuda(src)
            assert np.array_equal(got, ref), (devs, size)
            assert cn.bits_to_n_cuda(got, size) == oracle.canonical(n), (devs, size)
        size = 3 * (1 << 23) + 999
        n = oracle.generate(size, seed=5, alphabet=10).copy()
        n[::1000003] = ord("N")
        n[(1 << 23) + 1] = 0xFF
        words, invalid = cn.n_to_bits_ex_cuda(n, cn.ENC_LUT_EXACT)
        assert invalid == oracle.count_invalid(n) and np.array_equal(words, oracle.encode_mt(n, "lut"))
        words, invalid = cn.n_to_bits_checked_cuda(n)
        assert invalid == oracle.count_invalid(n)
        m = oracle.generate2(size, seed=6, alphabet=12)
        w2 = cn.n_to_bits2_cuda(m)
        assert np.array_equal(w2, oracle.n_to_bits2(m, "lut")), devs
        assert cn.bits_to_n2_cuda(w2, size) == oracle.canonical2(m), devs
    assert torch.cuda.current_device() == 0                 # the caller's current device is left alone


def test_fan_out_from_several_threads(cn, oracle, torch_cuda):
    """Concurrent fanned-out calls share the per-device workers and the copier pool."""
    torch = torch_cuda
    cn.set_devices(device_lists(torch)[-2] if torch.cuda.device_count() >= 2 else [0, 0])
    errors = []

    def worker(seed):
        try:
            size = (1 << 24) + seed * 1001
            n = oracle.generate(size, seed=seed, alphabet=10)
            got = cn.n_to_bits_cuda(n)
            if not np.array_equal(got, oracle.encode_mt(n, "lut", threads=2)):
                errors.append((seed, "encode"))
            if cn.bits_to_n_cuda(got, size) != oracle.canonical(n):
                errors.append((seed, "decode"))
        except Exception as e:          # noqa: BLE001
            errors.append((seed, repr(e)))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_encode_decode_sharded_device(cn, oracle, torch_cuda):
    """cn_encode_sharded / cn_decode_sharded: one launch per device, shards planned by cn_shard_bounds."""
    torch = torch_cuda
    ngpu = torch.cuda.device_count()
    for nshards in (1, 2, 5):
        devices = [k % ngpu for k in range(nshards)]
        total = (1 << 24) + 77
        whole = oracle.generate(total, seed=nshards, alphabet=10)
        ref = oracle.encode_mt(whole, "lut")
        bounds = [cn.shard_bounds_c(total, nshards, k, 1 << 20) for k in range(nshards)]
        shards = [torch.from_numpy(whole[s:e]).to(f"cuda:{d}") for (s, e), d in zip(bounds, devices)]
        outs = cn.encode_sharded_devices(shards)
        cn.synchronize_devices(sorted(set(devices)))
        got = np.concatenate([o.cpu().numpy().view(np.uint64) for o in outs])
        assert np.array_equal(got, ref)
        dec = [torch.empty(e - s, dtype=torch.uint8, device=f"cuda:{d}") for (s, e), d in zip(bounds, devices)]
        cn.decode_sharded_devices([o.data_ptr() for o in outs], [o.numel() for o in outs], [e - s for s, e in bounds], dec)
        cn.synchronize_devices(sorted(set(devices)))
        assert b"".join(t.cpu().numpy().tobytes() for t in dec) == oracle.canonical(whole)
    # the reference's panic, before anything is launched
    from cute_nucleotides_b200 import _lib
    with pytest.raises(cn.LengthError):
        cn.decode_sharded_devices([outs[0].data_ptr()], [1], [33], [dec[0]])
    assert torch.cuda.current_device() == 0


def test_gather_to_root_and_scatter_decode_over_peer_memory(cn, oracle, torch_cuda):
    """Encode + gather-to-root in ONE kernel per device (stores into the root's buffer over NVLink), then scatter +
    decode in one kernel per device (loads from the root's buffer).  Needs >= 2 GPUs."""
    torch = torch_cuda
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    devices = list(range(min(ngpu, 8)))
    cn.enable_peer_access(devices)
    cn.enable_peer_access(devices)                       # idempotent
    for total in ((1 << 26) + 77, 100003, 1 << 28):
        whole = oracle.generate(total, seed=21, alphabet=10)
        bounds = [cn.shard_bounds_c(total, len(devices), k, 1 << 20 if total > (1 << 22) else 32) for k in range(len(devices))]
        shards = [torch.from_numpy(whole[s:e]).to(f"cuda:{d}") for (s, e), d in zip(bounds, devices)]
        for root in (0, devices[-1]):
            full = torch.full((cn.words_for_len(total),), -1, dtype=torch.int64, device=f"cuda:{root}")
            cn.synchronize_devices(devices)
            cn.encode_gather_to_root(shards, full)
            cn.synchronize_devices(devices)
            assert np.array_equal(full.cpu().numpy().view(np.uint64), oracle.encode_mt(whole, "lut")), (total, root)
            outs = [torch.empty(e - s, dtype=torch.uint8, device=f"cuda:{d}") for (s, e), d in zip(bounds, devices)]
            cn.scatter_decode_from_root(full, bounds, outs)
            cn.synchronize_devices(devices)
            assert b"".join(t.cpu().numpy().tobytes() for t in outs) == oracle.canonical(whole), (total, root)


def test_multi_destination_encode_both_codecs(cn, oracle, torch_cuda):
    """cn_encode_multi_device / cn_encode2_multi_device: the shard is read once and every packed word is stored to each
    destination.  Destinations live on every visible GPU (peer access) -- or all on device 0 when only one is visible."""
    import ctypes
    torch = torch_cuda
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    ngpu = torch.cuda.device_count()
    devices = list(range(min(ngpu, 8)))
    if ngpu > 1:
        cn.enable_peer_access(devices)
    dest_devs = devices if ngpu > 1 else [0, 0, 0]
    st = torch.cuda.current_stream().cuda_stream
    for size in (100, 3456 * 5 + 13, (1 << 22) + 77):
        # 2-bit
        n = oracle.generate(size, seed=size, alphabet=10)
        d_n = torch.from_numpy(n).cuda()
        outs = [torch.full((cn.words_for_len(size) + 4,), -1, dtype=torch.int64, device=f"cuda:{d}") for d in dest_devs]
        ptrs = (ctypes.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        _lib.check(lib.cn_encode_multi_device(d_n.data_ptr(), size, ptrs, len(outs), st))
        cn.synchronize_devices(devices)
        ref = oracle.n_to_bits(n, "lut")
        for o in outs:
            assert np.array_equal(o[:-4].cpu().numpy().view(np.uint64), ref) and int(o[-1].item()) == -1
        # base-5
        m = oracle.generate2(size, seed=size + 1, alphabet=12)
        d_m = torch.from_numpy(m).cuda()
        outs = [torch.full((cn.words2_for_len(size) + 4,), -1, dtype=torch.int64, device=f"cuda:{d}") for d in dest_devs]
        ptrs = (ctypes.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        _lib.check(lib.cn_encode2_multi_device(d_m.data_ptr(), size, ptrs, len(outs), st))
        cn.synchronize_devices(devices)
        ref2 = oracle.n_to_bits2(m, "lut")
        for o in outs:
            assert np.array_equal(o[:-4].cpu().numpy().view(np.uint64), ref2) and int(o[-1].item()) == -1
    bad = (ctypes.c_void_p * 1)(outs[0].data_ptr() + 8)
    assert lib.cn_encode2_multi_device(d_m.data_ptr(), 100, bad, 1, st) == _lib.CN_ERR_ARG      # destination not 32-byte aligned
