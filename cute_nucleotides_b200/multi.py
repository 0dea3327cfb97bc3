"""Single-process multi-GPU mirror of the C ABI's sharded entry points (include/cute_nucleotides_cuda.h,
"multi-GPU inside ONE process").  One Python process drives every GPU: no torch.distributed, no NCCL.

    set_devices([0, 1, ...])            the host-slice calls (n_to_bits_cuda, bits_to_n_cuda, ...) fan out over these GPUs
    encode_sharded_devices(shards)      one launch per device on device-resident shards (cn_encode_sharded)
    encode_gather_to_root(shards, r)    the same launches, but every shard's words are stored straight into ONE buffer on
                                        device r over NVLink peer memory: encode + gather-to-root with no collective
    scatter_decode_from_root(full, ..)  the inverse: every device decodes its word range reading device r's buffer

The split is by sequence offset at word boundaries (word w depends only on nucleotides 32w..32w+31,
src/n_to_bits.rs:39-42), so results are identical to a single-device call.
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

from . import _lib
from ._lib import check
from .n_to_bits import words_for_len

__all__ = ["set_devices", "get_devices", "shard_bounds_c", "encode_sharded_devices", "decode_sharded_devices",
           "enable_peer_access", "synchronize_devices", "encode_gather_to_root", "scatter_decode_from_root"]


def _int_array(values: Sequence[int]):
    return (ctypes.c_int * len(values))(*values)


def set_devices(devices: Sequence[int]) -> None:
    """Devices the host-slice calls fan out over; an empty list restores the single-device behaviour."""
    devices = list(devices)
    check(_lib.load().cn_set_devices(_int_array(devices) if devices else None, len(devices)))


def get_devices() -> List[int]:
    buf = (ctypes.c_int * 64)()
    count = ctypes.c_int(0)
    check(_lib.load().cn_get_devices(buf, 64, ctypes.byref(count)))
    return [buf[i] for i in range(count.value)]


def shard_bounds_c(total_len: int, nshards: int, shard: int, granule: int = 1 << 20) -> Tuple[int, int]:
    """The library's own planner (cn_shard_bounds); identical to sharded.shard_bounds."""
    s, e = ctypes.c_size_t(), ctypes.c_size_t()
    check(_lib.load().cn_shard_bounds(total_len, nshards, shard, granule, ctypes.byref(s), ctypes.byref(e)))
    return int(s.value), int(e.value)


def enable_peer_access(devices: Sequence[int]) -> None:
    devices = list(devices)
    check(_lib.load().cn_enable_peer_access(_int_array(devices), len(devices)))


def synchronize_devices(devices: Sequence[int]) -> None:
    devices = list(devices)
    check(_lib.load().cn_synchronize_devices(_int_array(devices), len(devices)))


def _ptrs(values: Sequence[int]):
    return (ctypes.c_void_p * len(values))(*values)


def _sizes(values: Sequence[int]):
    return (ctypes.c_size_t * len(values))(*values)


def _streams(tensors):
    import torch
    return _ptrs([torch.cuda.current_stream(t.device).cuda_stream for t in tensors])


def encode_sharded_devices(shards, outs=None, out_ptrs=None):
    """shards[k]: uint8 CUDA tensor on its own device.  Returns the list of packed int64 tensors (one per device), or
    writes to `out_ptrs[k]` (raw device addresses, possibly on another device) when given.  Asynchronous."""
    import torch
    devices = [t.device.index for t in shards]
    lens = [t.numel() for t in shards]
    if out_ptrs is None:
        if outs is None:
            outs = [torch.empty(words_for_len(n), dtype=torch.int64, device=t.device) for t, n in zip(shards, lens)]
        out_ptrs = [o.data_ptr() for o in outs]
    check(_lib.load().cn_encode_sharded(len(shards), _int_array(devices), _ptrs([t.data_ptr() for t in shards]), _sizes(lens),
                                        _ptrs(out_ptrs), _streams(shards)))
    return outs


def decode_sharded_devices(word_ptrs, nwords, lens, outs):
    """Decode shard k (nwords[k] words at raw device address word_ptrs[k], possibly on another device) into the uint8
    CUDA tensor outs[k] on its own device.  Asynchronous."""
    devices = [o.device.index for o in outs]
    check(_lib.load().cn_decode_sharded(len(outs), _int_array(devices), _ptrs(word_ptrs), _sizes(nwords), _sizes(lens),
                                        _ptrs([o.data_ptr() for o in outs]), _streams(outs)))
    return outs


def encode_gather_to_root(shards, full):
    """Encode every device's shard and store the packed words directly into `full` (an int64 tensor on the root device
    holding ceil(total/32) words) -- one kernel per device, the stores of the non-root devices travel over NVLink.
    All shards but the last must be multiples of 32 nucleotides.  Call enable_peer_access first.  Asynchronous."""
    offs, pos = [], 0
    for t in shards[:-1]:
        if t.numel() % 32:
            raise ValueError("every shard but the last must be a multiple of 32 nucleotides")
    for t in shards:
        offs.append(pos // 32 * 8)
        pos += t.numel()
    if full.numel() < words_for_len(pos):
        raise ValueError("root buffer too small")
    encode_sharded_devices(shards, out_ptrs=[full.data_ptr() + o for o in offs])
    return full


def scatter_decode_from_root(full, bounds, outs):
    """Inverse of encode_gather_to_root: device k decodes nucleotides [bounds[k][0], bounds[k][1]) reading its word range
    straight out of `full` on the root device (peer loads over NVLink)."""
    word_ptrs = [full.data_ptr() + (s // 32) * 8 for s, _ in bounds]
    lens = [e - s for s, e in bounds]
    nwords = [words_for_len(n) for n in lens]
    return decode_sharded_devices(word_ptrs, nwords, lens, outs)
