"""N-GPU test of the sharded path over NCCL (skipped unless >= 2 GPUs are visible): every rank generates and
encodes only its shard on its own GPU through the C ABI, the packed shards are assembled with the one
collective the path can use, and every rank decodes its word range; all compared bit-exact with the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, granule, result_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    import cute_nucleotides_b200 as cn
    from cute_nucleotides_b200 import sharded
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        orc = _oracle.Oracle()
        start, end = sharded.shard_bounds(total, world, rank, granule)
        d_shard = cn.generate_device(torch.empty(end - start, dtype=torch.uint8, device="cuda"), start, 7, 10)
        full = sharded.encode_sharded(d_shard, total, assemble=True, granule=granule)
        whole = orc.generate(total, seed=7, alphabet=10)
        ok = np.array_equal(full.cpu().numpy().view(np.uint64), orc.encode_mt(whole, "lut"))
        dec = sharded.decode_sharded(full, total, granule=granule)
        ok = ok and dec.cpu().numpy().tobytes() == orc.canonical(whole[start:end])
        # gather-to-root / scatter over NCCL (the collectives north_star names), then decode what was scattered
        ws, we = sharded.word_bounds(total, world, rank, granule)
        for root in (0, world - 1):
            gathered = sharded.gather_packed(full[ws:we].clone(), total, root=root, granule=granule)
            if rank == root:
                ok = ok and bool(torch.equal(gathered, full))
            mine = sharded.scatter_packed(gathered, total, root=root, granule=granule, like=full)
            ok = ok and bool(torch.equal(mine, full[ws:we]))
            ok = ok and cn.decode_device(mine, end - start).cpu().numpy().tobytes() == orc.canonical(whole[start:end])
        # the base-5 codec through the same sharded entry points (27-nt words; granule keeps word slices 32-byte aligned)
        total5 = total // 3 + 7
        g5 = sharded.default_granule("base5") * 64
        s5, e5 = sharded.shard_bounds(total5, world, rank, g5, 27)
        whole5 = orc.generate2(total5, seed=8, alphabet=12)
        d5 = torch.from_numpy(whole5[s5:e5]).cuda()
        full5 = sharded.encode_sharded(d5, total5, assemble=True, granule=g5, codec="base5")
        ok = ok and np.array_equal(full5.cpu().numpy().view(np.uint64), orc.n_to_bits2(whole5, "lut"))
        dec5 = sharded.decode_sharded(full5, total5, granule=g5, codec="base5")
        ok = ok and dec5.cpu().numpy().tobytes() == orc.canonical2(whole5[s5:e5])
        open(os.path.join(result_dir, f"rank{rank}"), "w").write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total,granule", [((1 << 26) + 77, 1 << 20), (100003, 32), (1 << 28, 1 << 20)])
def test_sharded_encode_assemble_decode_nccl(tmp_path, total, granule):
    import torch
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 8)
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), total, granule, str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"rank{r}").read() for r in range(world)] == ["ok"] * world


def test_one_process_two_gpus_via_cn_init(tmp_path):
    """The host-slice calls run on the device the calling thread selected with cn_init(): two threads of ONE process
    drive two GPUs (and two PCIe links) at once, each with its own thread-local staging."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    import cute_nucleotides_b200 as cn
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    orc = _oracle.Oracle()
    total = (1 << 24) + 77
    whole = orc.generate(total, seed=3, alphabet=10)
    from cute_nucleotides_b200 import sharded
    out = np.zeros(cn.words_for_len(total), dtype=np.uint64)
    errors = []

    def worker(dev):
        try:
            _lib.check(lib.cn_init(dev))
            s, e = sharded.shard_bounds(total, 2, dev, 1 << 20)
            ws, we = sharded.word_bounds(total, 2, dev, 1 << 20)
            out[ws:we] = cn.n_to_bits_cuda(whole[s:e])
        except Exception as ex:      # noqa: BLE001
            errors.append(repr(ex))

    threads = [threading.Thread(target=worker, args=(d,)) for d in (0, 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert np.array_equal(out, orc.encode_mt(whole, "lut"))


def _fused_worker(rank, world, port, total, granule, result_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    import cute_nucleotides_b200 as cn
    from cute_nucleotides_b200 import sharded
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        orc = _oracle.Oracle()
        start, end = sharded.shard_bounds(total, world, rank, granule)
        d_shard = cn.generate_device(torch.empty(end - start, dtype=torch.uint8, device="cuda"), start, 11, 10)
        asm = sharded.PeerAssembly(total, granule=granule)
        asm.full.fill_(-1)
        torch.cuda.synchronize()
        dist.barrier()
        for _ in range(2):                                  # twice: the mapping is reusable
            asm.encode(d_shard)
            full = asm.finish()
        whole = orc.generate(total, seed=11, alphabet=10)
        ok = np.array_equal(full.cpu().numpy().view(np.uint64), orc.encode_mt(whole, "lut"))
        nccl = sharded.encode_sharded(d_shard, total, assemble=True, granule=granule)
        ok = ok and bool(torch.equal(nccl, full))
        # decode this rank's range straight out of ANOTHER rank's buffer (peer loads over NVLink)
        remote = asm.decode_from((rank + 1) % world)
        torch.cuda.synchronize()
        ok = ok and remote.cpu().numpy().tobytes() == orc.canonical(whole[start:end])
        # encode + gather-to-ROOT only: every rank stores its words into the root's buffer and nowhere else
        root = world - 1
        dist.barrier()                                       # peers may still be reading this rank's buffer (decode_from above)
        asm.full.fill_(-1)
        torch.cuda.synchronize()
        dist.barrier()
        asm.encode(d_shard, dests=[root])
        asm.finish()
        if rank == root:
            ok = ok and np.array_equal(asm.full.cpu().numpy().view(np.uint64), orc.encode_mt(whole, "lut"))
        else:
            ok = ok and bool((asm.full == -1).all())
        scattered = asm.decode_from(root)                    # scatter + decode in one kernel, straight from the root
        torch.cuda.synchronize()
        ok = ok and scattered.cpu().numpy().tobytes() == orc.canonical(whole[start:end])
        asm.close()
        # the same fused kernels for the base-5 codec (cn_encode2_multi_device; 27-nt words, shards on 108-nt boundaries)
        total5 = total // 2 + 11
        g5 = sharded.BASE5_ALIGNED_GRANULE * (64 if total5 > (1 << 20) else 1)
        s5, e5 = sharded.shard_bounds(total5, world, rank, g5, 27)
        whole5 = orc.generate2(total5, seed=13, alphabet=12)
        d5 = torch.from_numpy(whole5[s5:e5]).cuda()
        asm5 = sharded.PeerAssembly(total5, granule=g5, codec="base5")
        asm5.full.fill_(-1)
        torch.cuda.synchronize()
        dist.barrier()
        asm5.encode(d5)
        full5 = asm5.finish()
        ok = ok and np.array_equal(full5.cpu().numpy().view(np.uint64), orc.n_to_bits2(whole5, "lut"))
        back5 = asm5.decode_from((rank + 1) % world)
        torch.cuda.synchronize()
        ok = ok and back5.cpu().numpy().tobytes() == orc.canonical2(whole5[s5:e5])
        asm5.close()
        open(os.path.join(result_dir, f"rank{rank}"), "w").write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total,granule", [((1 << 26) + 77, 1 << 20), (100003, 32), (1 << 28, 1 << 20)])
def test_fused_encode_assemble_over_peer_memory(tmp_path, total, granule):
    """cn_encode_multi_device: one kernel per rank writes its packed shard into every rank's buffer over NVLink;
    result identical to encode + NCCL all-gather and to the oracle."""
    import torch
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 8)
    import torch.multiprocessing as mp
    mp.spawn(_fused_worker, args=(world, _free_port(), total, granule, str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"rank{r}").read() for r in range(world)] == ["ok"] * world
