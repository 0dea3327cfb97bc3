"""Multi-GPU host logic: shard one logical sequence by offset, one process per GPU (torch.distributed).

The codec needs no exchange: packed word w depends only on nucleotides 32w..32w+31 (src/n_to_bits.rs:39-42),
so rank r encodes / decodes its own contiguous range and NO collective sits on the data path.  The only
collective the path can use is assembling the packed shards afterwards (`assemble_packed`, an all-gather);
bench.py times it separately.  This module is backend-agnostic (nccl on GPUs, gloo in the CPU tests): it
only plans ranges and moves already-encoded tensors, it never computes codes itself.
"""
from __future__ import annotations

from typing import List, Tuple

GRANULE = 32          # nucleotides per packed u64 of the 2-bit codec: shard boundaries must be multiples of this
GROUP_2BIT, GROUP_BASE5 = 32, 27      # nucleotides per u64: n_to_bits (src/n_to_bits.rs:35) / n_to_bits2 (src/n_to_bits2.rs:38)


def shard_bounds(total_len: int, world: int, rank: int, granule: int = GRANULE, group: int = GROUP_2BIT) -> Tuple[int, int]:
    """Contiguous range [start, end) of nucleotides owned by `rank`.

    Boundaries are multiples of `granule` (itself a multiple of `group`, the nucleotides per packed word: 32 for the
    2-bit codec, 27 for the base-5 codec) so every rank writes whole, aligned words; ranges are balanced to within one
    granule; the ragged global tail lands on the last non-empty rank.  Ranks beyond the data get an empty range."""
    if granule % group:
        raise ValueError(f"granule must be a multiple of {group} nucleotides")
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    units = -(-total_len // granule)                       # ceil
    base, extra = divmod(units, world)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    start = min(first * granule, total_len)
    end = min((first + count) * granule, total_len)
    return start, end


def all_bounds(total_len: int, world: int, granule: int = GRANULE, group: int = GROUP_2BIT) -> List[Tuple[int, int]]:
    return [shard_bounds(total_len, world, r, granule, group) for r in range(world)]


def words_for_len(length: int, group: int = GROUP_2BIT) -> int:
    return length // group + (1 if length % group else 0)


def word_bounds(total_len: int, world: int, rank: int, granule: int = GRANULE, group: int = GROUP_2BIT) -> Tuple[int, int]:
    """Range of packed words produced by `rank` (start is exact because shard starts are multiples of `group`)."""
    start, end = shard_bounds(total_len, world, rank, granule, group)
    return start // group, start // group + words_for_len(end - start, group)


def assemble_packed(local_words, total_len: int, group=None, granule: int = GRANULE, codec_group: int = GROUP_2BIT):
    """All-gather the packed shards: every rank returns the full ceil(total_len/32)-word tensor.

    `local_words` is this rank's 8-byte-element tensor (CPU for gloo, CUDA for nccl) holding exactly the words
    of its shard.  Equal shards use all_gather_into_tensor (one NCCL collective, in place); ragged shards fall
    back to all_gather on padded buffers."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    spans = [word_bounds(total_len, world, r, granule, codec_group) for r in range(world)]
    counts = [e - s for s, e in spans]
    if local_words.numel() != counts[rank]:
        raise ValueError(f"rank {rank}: expected {counts[rank]} words, got {local_words.numel()}")
    total_words = words_for_len(total_len, codec_group)
    full = torch.empty(total_words, dtype=local_words.dtype, device=local_words.device)
    if len(set(counts)) == 1 and counts[0] * world == total_words:
        dist.all_gather_into_tensor(full, local_words.contiguous(), group=group)
        return full
    width = max(counts)
    padded = torch.zeros(width, dtype=local_words.dtype, device=local_words.device)
    padded[: counts[rank]] = local_words
    pieces = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(pieces, padded, group=group)
    for (s, e), piece in zip(spans, pieces):
        full[s:e] = piece[: e - s]
    return full


# The base-5 tiled decode wants its packed input 32-byte aligned: shard boundaries that are multiples of 4 words (108 nt)
# keep every rank's slice of the assembled tensor on the fast path.
BASE5_ALIGNED_GRANULE = 27 * 4


def _codec(codec: str):
    """(nucleotides per word, encode, decode) of "2bit" (n_to_bits) or "base5" (n_to_bits2)."""
    from . import n_to_bits, n_to_bits2
    if codec == "2bit":
        return GROUP_2BIT, n_to_bits.encode_device, n_to_bits.decode_device
    if codec == "base5":
        return GROUP_BASE5, n_to_bits2.encode2_device, n_to_bits2.decode2_device
    raise ValueError("codec must be '2bit' or 'base5'")


def default_granule(codec: str) -> int:
    return GRANULE if codec == "2bit" else BASE5_ALIGNED_GRANULE


def encode_sharded(n_shard, total_len: int, group=None, assemble: bool = False, granule: int = None, codec: str = "2bit"):
    """Encode this rank's shard on its GPU (cn_encode_device / cn_encode2_device through the C ABI); optionally assemble."""
    import torch.distributed as dist

    grp, enc, _ = _codec(codec)
    granule = default_granule(codec) if granule is None else granule
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    start, end = shard_bounds(total_len, world, rank, granule, grp)
    if n_shard.numel() != end - start:
        raise ValueError(f"rank {rank}: shard must hold nucleotides [{start}, {end})")
    words = enc(n_shard)
    return assemble_packed(words, total_len, group, granule, grp) if assemble else words


def decode_sharded(full_words, total_len: int, group=None, granule: int = None, codec: str = "2bit"):
    """Decode this rank's range of a packed tensor every rank already holds (e.g. after assemble/broadcast)."""
    import torch.distributed as dist

    grp, _, dec = _codec(codec)
    granule = default_granule(codec) if granule is None else granule
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    start, end = shard_bounds(total_len, world, rank, granule, grp)
    ws, we = word_bounds(total_len, world, rank, granule, grp)
    return dec(full_words[ws:we], end - start)


def gather_packed(local_words, total_len: int, root: int = 0, group=None, granule: int = GRANULE, codec_group: int = GROUP_2BIT):
    """Gather-to-root of the packed shards (the collective north_star names): `root` returns the full tensor, the
    other ranks None.  NCCL: grouped send/recv; gloo: the same calls on CPU tensors."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    spans = [word_bounds(total_len, world, r, granule, codec_group) for r in range(world)]
    if local_words.numel() != spans[rank][1] - spans[rank][0]:
        raise ValueError(f"rank {rank}: expected {spans[rank][1] - spans[rank][0]} words, got {local_words.numel()}")
    if rank != root:
        if local_words.numel():
            for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local_words.contiguous(), root, group)]):
                q.wait()
        return None
    full = torch.empty(words_for_len(total_len, codec_group), dtype=local_words.dtype, device=local_words.device)
    s, e = spans[root]
    full[s:e] = local_words
    ops = [dist.P2POp(dist.irecv, full[s:e], r, group) for r, (s, e) in enumerate(spans) if r != root and e > s]
    if ops:                                              # one batched (grouped) NCCL call: the receives run concurrently
        for q in dist.batch_isend_irecv(ops):
            q.wait()
    return full


def scatter_packed(full_words, total_len: int, root: int = 0, group=None, granule: int = GRANULE, codec_group: int = GROUP_2BIT,
                   like=None):
    """Scatter of the packed words from `root` (which holds `full_words`): every rank returns its own word range.
    `like` gives dtype/device on the non-root ranks."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    spans = [word_bounds(total_len, world, r, granule, codec_group) for r in range(world)]
    s, e = spans[rank]
    if rank == root:
        ops = [dist.P2POp(dist.isend, full_words[a:b], r, group) for r, (a, b) in enumerate(spans) if r != root and b > a]
        reqs = dist.batch_isend_irecv(ops) if ops else []
        mine = full_words[s:e].clone()
        for q in reqs:
            q.wait()
        return mine
    ref = like if like is not None else full_words
    mine = torch.empty(e - s, dtype=ref.dtype, device=ref.device)
    if e > s:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.irecv, mine, root, group)]):
            q.wait()
    return mine


class PeerAssembly:
    """Encode + assemble in ONE kernel per rank over NVLink peer memory (SURVEY 8e "fused option").

    Every rank owns a full-size packed buffer; the buffers are mapped into every other rank's address space with
    CUDA IPC (handles exchanged through torch.distributed), and `encode(n_shard)` launches a single kernel that reads
    the local shard once and stores each packed word into ALL ranks' buffers at the shard's word offset -- no
    intermediate packed shard, no NCCL collective on the data path.  `finish()` is the only synchronisation: it
    waits for the local kernel and then for every rank's (a barrier), after which `self.full` holds the whole
    packed sequence on every rank.  Needs one process per GPU on one NVLink-connected box (world <= 8)."""

    def __init__(self, total_len: int, group=None, granule: int = None, codec: str = "2bit"):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _lib

        self._lib, self._ctypes, self._torch, self._dist = _lib, ctypes, torch, dist
        self.lib = _lib.load()
        self.codec = codec
        self.grp = GROUP_2BIT if codec == "2bit" else GROUP_BASE5
        if codec not in ("2bit", "base5"):
            raise ValueError("codec must be '2bit' or 'base5'")
        if granule is None:
            granule = 1 << 20 if codec == "2bit" else BASE5_ALIGNED_GRANULE * 8192
        if codec == "base5" and granule % BASE5_ALIGNED_GRANULE:
            raise ValueError("base-5 shards must start on multiples of 108 nucleotides (4 words) so that every destination stays 32-byte aligned")
        self._enc = self.lib.cn_encode_multi_device if codec == "2bit" else self.lib.cn_encode2_multi_device
        self._dec = self.lib.cn_decode_device if codec == "2bit" else self.lib.cn_decode2_device
        self.group, self.total_len, self.granule = group, total_len, granule
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("PeerAssembly supports at most 8 ranks (one NVSwitch domain)")
        self.full = torch.empty(words_for_len(total_len, self.grp), dtype=torch.int64, device="cuda")
        handle = ctypes.create_string_buffer(64)
        offset = ctypes.c_size_t()
        _lib.check(self.lib.cn_ipc_export(self.full.data_ptr(), handle, ctypes.byref(offset)))
        mine = (bytes(handle.raw), int(offset.value))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._opened = []
        self.peer_ptrs = []
        for r, (h, off) in enumerate(everyone):
            if r == self.rank:
                self.peer_ptrs.append(self.full.data_ptr())
                continue
            p = ctypes.c_void_p()
            _lib.check(self.lib.cn_ipc_open(h, off, ctypes.byref(p)))
            self._opened.append((p.value, off))
            self.peer_ptrs.append(p.value)
        dist.barrier(group=group)

    def encode(self, n_shard, stream=None, dests=None):
        """Launch the fused kernel for this rank's shard (asynchronous).  `dests` = ranks whose buffers receive the packed
        words: all of them by default (encode + all-gather), `[root]` for encode + gather-to-root."""
        torch, ctypes = self._torch, self._ctypes
        start, end = shard_bounds(self.total_len, self.world, self.rank, self.granule, self.grp)
        if n_shard.numel() != end - start or n_shard.dtype != torch.uint8 or not n_shard.is_cuda:
            raise ValueError(f"rank {self.rank}: shard must be the uint8 CUDA tensor of nucleotides [{start}, {end})")
        word_off = (start // self.grp) * 8
        ranks = list(range(self.world)) if dests is None else list(dests)
        outs = (ctypes.c_void_p * len(ranks))(*[self.peer_ptrs[r] + word_off for r in ranks])
        s = stream if stream is not None else torch.cuda.current_stream()
        self._lib.check(self._enc(n_shard.data_ptr(), end - start, outs, len(ranks), s.cuda_stream))

    def decode_from(self, source_rank: int, stream=None, out=None):
        """Scatter/broadcast + decode in one kernel: decode THIS rank's range of the packed sequence reading the words
        straight out of `source_rank`'s buffer (peer loads over NVLink) -- no copy of the packed words first."""
        torch = self._torch
        start, end = shard_bounds(self.total_len, self.world, self.rank, self.granule, self.grp)
        ws, we = word_bounds(self.total_len, self.world, self.rank, self.granule, self.grp)
        if out is None:
            out = torch.empty(end - start, dtype=torch.uint8, device="cuda")
        s = stream if stream is not None else torch.cuda.current_stream()
        self._lib.check(self._dec(self.peer_ptrs[source_rank] + ws * 8, we - ws, end - start, out.data_ptr(), s.cuda_stream))
        return out

    def finish(self):
        """Wait until every rank's stores have landed everywhere; returns the assembled tensor."""
        self._torch.cuda.synchronize()
        self._dist.barrier(group=self.group)
        return self.full

    def close(self):
        self._torch.cuda.synchronize()
        self._dist.barrier(group=self.group)
        for p, off in self._opened:
            self.lib.cn_ipc_close(p, off)
        self._opened = []
