"""One pinned host call fanned out over all visible GPUs: host strategy (copy engines vs zero-copy kernels) x chunk size.
Harness for the N-link host ceiling (profiles/host_fanout_r02_n8.jsonl).  Usage: python tools/bench_fanout.py [GiB per GPU]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import cute_nucleotides_b200 as cn
from cute_nucleotides_b200 import _lib
lib = _lib.load()
ngpu = torch.cuda.device_count()
per = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
L = int(per * (1 << 30)) * ngpu
W = cn.words_for_len(L)
h_n = torch.empty(L, dtype=torch.uint8, pin_memory=True)
h_bits = torch.empty(W, dtype=torch.int64, pin_memory=True)
h_out = torch.empty(L, dtype=torch.uint8, pin_memory=True)
tmp = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:0")
for s in range(0, L, tmp.numel()):
    e = min(L, s + tmp.numel())
    cn.generate_device(tmp[: e - s], s, 5, 10)
    h_n[s:e].copy_(tmp[: e - s])
del tmp
torch.cuda.synchronize()
for devs in ([0], list(range(min(2, ngpu))), list(range(min(4, ngpu))), list(range(ngpu))):
    if len(devs) > 1 and devs == [0]:
        continue
    cn.set_devices(devs)
    n = int(per * (1 << 30)) * len(devs)
    for strategy in (0, 1):
        for chunk_mib in (4, 16, 64):
            _lib.check(lib.cn_set_host_strategy(strategy, chunk_mib << 20))
            enc = lambda: _lib.check(lib.cn_n_to_bits_host(h_n.data_ptr(), n, h_bits.data_ptr()))
            dec = lambda: _lib.check(lib.cn_bits_to_n_host(h_bits.data_ptr(), cn.words_for_len(n), n, h_out.data_ptr()))
            enc(); dec()
            te = td = 1e30
            for _ in range(3):
                a = time.perf_counter(); enc(); b = time.perf_counter(); dec(); c = time.perf_counter()
                te, td = min(te, b - a), min(td, c - b)
            print(json.dumps({"gpus": len(devs), "strategy": "zero-copy kernels" if strategy else "copy engines", "chunk_mib": chunk_mib,
                              "encode_gnt_s": round(n / te / 1e9, 1), "decode_gnt_s": round(n / td / 1e9, 1),
                              "roundtrip_gnt_s": round(n / (te + td) / 1e9, 1)}), flush=True)
cn.set_devices([])
