"""Sweep the host-slice pipeline on pinned and pageable buffers; prints JSON lines.
Usage (GPU box): python tools/bench_host.py [GiB] [pinned|pageable|all]
pinned:   strategy x chunk size
pageable: copier threads (ascending: the pool only grows) x staging chunk size"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import cute_nucleotides_b200 as cn
from cute_nucleotides_b200 import _lib
lib = _lib.load()
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4
what = sys.argv[2] if len(sys.argv) > 2 else "all"
L = int(gib * (1 << 30))
W = cn.words_for_len(L)
tmp = cn.generate_device(torch.empty(L, dtype=torch.uint8, device="cuda"), 0, 3, 10)
pin_n = torch.empty(L, dtype=torch.uint8).pin_memory(); pin_n.copy_(tmp)
pin_bits = torch.empty(W, dtype=torch.int64).pin_memory()
pin_out = torch.empty(L, dtype=torch.uint8).pin_memory()
pg_n = pin_n.numpy().copy(); pg_bits = np.empty(W, dtype=np.uint64); pg_out = np.empty(L, dtype=np.uint8)
pg_bits[:] = 0; pg_out[:] = 0                      # pre-fault
del tmp
torch.cuda.synchronize()

def run(name, enc, dec, reps=3):
    enc(); dec()
    te = td = 1e30
    for _ in range(reps):
        a = time.perf_counter(); enc(); b = time.perf_counter(); dec(); c = time.perf_counter()
        te, td = min(te, b - a), min(td, c - b)
    return {"case": name, "encode_gnt_s": round(L / te / 1e9, 2), "decode_gnt_s": round(L / td / 1e9, 2),
            "roundtrip_gnt_s": round(L / (te + td) / 1e9, 2)}

if what in ("all", "pinned"):
    for strategy in (0, 1):
        for chunk_mib in (4, 16, 64):
            _lib.check(lib.cn_set_host_strategy(strategy, chunk_mib << 20))
            r = run("pinned", lambda: _lib.check(lib.cn_n_to_bits_host(pin_n.data_ptr(), L, pin_bits.data_ptr())),
                    lambda: _lib.check(lib.cn_bits_to_n_host(pin_bits.data_ptr(), W, L, pin_out.data_ptr())))
            r.update(strategy=strategy, chunk_mib=chunk_mib); print(json.dumps(r), flush=True)
    _lib.check(lib.cn_set_host_strategy(0, 16 << 20))
if what in ("all", "pageable"):
    ok = None
    for threads in (2, 4, 6, 8, 10, 12, 16):
        if threads > (os.cpu_count() or 1):
            break
        _lib.check(lib.cn_set_host_threads(threads))
        for chunk_kib in (512, 1024, 2048, 4096, 8192, 16384):
            _lib.check(lib.cn_set_host_chunks(0, chunk_kib << 10))
            r = run("pageable (pre-faulted)", lambda: _lib.check(lib.cn_n_to_bits_host(pg_n.ctypes.data, L, pg_bits.ctypes.data)),
                    lambda: _lib.check(lib.cn_bits_to_n_host(pg_bits.ctypes.data, W, L, pg_out.ctypes.data)), reps=2)
            if ok is None:
                ok = bool(np.array_equal(pg_bits.view(np.int64), pin_bits.numpy())) if what == "all" else True
            r.update(threads=threads, chunk_kib=chunk_kib, verified=ok); print(json.dumps(r), flush=True)
