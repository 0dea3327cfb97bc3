"""CPU tests of the N > 1 host logic: shard planning, and the assemble collective on world_size-2 gloo.
The shards themselves are encoded by the ORACLE here (the product has no CPU path); what is under test is
that ranges cover the sequence on word boundaries and that assembling shard outputs reproduces the
single-shot encoding bit for bit -- i.e. that sharding by offset is exact for this codec."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cute_nucleotides_b200 import sharded  # noqa: E402


@pytest.mark.parametrize("total", [0, 1, 31, 32, 33, 64, 1000, 40000, (1 << 20) + 17, 10 * (1 << 30) + 21, 80 * (1 << 30)])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("granule", [32, 1 << 20])
def test_shard_bounds_cover_on_word_boundaries(total, world, granule):
    spans = sharded.all_bounds(total, world, granule)
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
        assert e0 == s1 and s0 <= e0
    for s, e in spans:
        assert s % 32 == 0 or s == total
        assert e % granule == 0 or e == total
    sizes = [e - s for s, e in spans]
    assert max(sizes) - min(sizes) < 2 * granule      # one unit of imbalance plus the ragged last unit
    words = [sharded.word_bounds(total, world, r, granule) for r in range(world)]
    assert sum(e - s for s, e in words) == sharded.words_for_len(total)


@pytest.mark.parametrize("total", [0, 1, 26, 27, 28, 1000, 40000, 27 * 4096 * 3 + 5, 10 * (1 << 30) + 21])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("granule", [27, 27 * 128, 27 * (1 << 15)])
def test_shard_bounds_base5(total, world, granule):
    """Same planner for the base-5 codec: boundaries on multiples of 27 nucleotides (whole words)."""
    spans = sharded.all_bounds(total, world, granule, group=27)
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
        assert e0 == s1 and s0 <= e0
    for s, e in spans:
        assert s % 27 == 0 or s == total
    words = [sharded.word_bounds(total, world, r, granule, group=27) for r in range(world)]
    assert sum(e - s for s, e in words) == sharded.words_for_len(total, 27)
    with pytest.raises(ValueError):
        sharded.shard_bounds(total, world, 0, 32, group=27)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, granule, result_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = _oracle.Oracle()
        start, end = sharded.shard_bounds(total, world, rank, granule)
        shard = orc.generate(end - start, seed=7, alphabet=10, offset=start)     # each rank generates only its range
        local = torch.from_numpy(orc.n_to_bits(shard, "lut").view(np.int64).copy())
        full = sharded.assemble_packed(local, total, granule=granule)
        whole = orc.n_to_bits(orc.generate(total, seed=7, alphabet=10), "lut")
        ok = np.array_equal(full.numpy().view(np.uint64), whole)
        # decode side: every rank decodes its own word range of the assembled tensor (oracle as the decoder)
        ws, we = sharded.word_bounds(total, world, rank, granule)
        dec = orc.bits_to_n(full.numpy().view(np.uint64)[ws:we], end - start, "lut")
        ok = ok and dec == orc.canonical(shard)
        # gather-to-root and scatter (the collectives north_star names), both roots
        for root in (0, world - 1):
            gathered = sharded.gather_packed(local, total, root=root, granule=granule)
            ok = ok and ((gathered is None) == (rank != root))
            if rank == root:
                ok = ok and np.array_equal(gathered.numpy().view(np.uint64), whole)
            mine = sharded.scatter_packed(gathered, total, root=root, granule=granule, like=local)
            ok = ok and torch.equal(mine, local)
        open(os.path.join(result_dir, f"rank{rank}"), "w").write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


def _worker5(rank, world, port, total, granule, result_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = _oracle.Oracle()
        start, end = sharded.shard_bounds(total, world, rank, granule, group=27)
        shard = orc.generate2(end - start, seed=9, alphabet=12, offset=start)
        local = torch.from_numpy(orc.n_to_bits2(shard, "lut").view(np.int64).copy())
        full = sharded.assemble_packed(local, total, granule=granule, codec_group=27)
        whole = orc.n_to_bits2(orc.generate2(total, seed=9, alphabet=12), "lut")
        ok = np.array_equal(full.numpy().view(np.uint64), whole)
        open(os.path.join(result_dir, f"rank{rank}"), "w").write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


def test_assemble_base5_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker5, args=(2, _free_port(), 27 * 5000 + 13, 27 * 64, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"rank{r}").read() for r in range(2)] == ["ok", "ok"]


@pytest.mark.parametrize("total,granule", [(64 * 1000, 32), (100003, 32), ((1 << 21) + 77, 1 << 20), (31, 32)])
def test_assemble_world2_gloo(tmp_path, total, granule):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, granule, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"rank{r}").read() for r in range(2)] == ["ok", "ok"]
