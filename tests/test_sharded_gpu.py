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
