"""Sweep the kernel launch configurations the library can select (vec x unroll x threads) on one GPU and
print one JSON line per configuration.  Timing: cn_time_*_device = CUDA events on the launching stream.
Usage (on the GPU box): python tools/tune.py [--nucleotides N] [--iters K] > gpurun_out/tune.jsonl"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cute_nucleotides_b200 as cn  # noqa: E402
from cute_nucleotides_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nucleotides", type=int, default=10 << 30)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
lib = _lib.load()
L = args.nucleotides
W = cn.words_for_len(L)
d_n = cn.generate_device(torch.empty(L, dtype=torch.uint8, device="cuda"), 0, 1, 10)
d_bits = torch.empty(W, dtype=torch.int64, device="cuda")
d_out = torch.empty(L, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ms = ctypes.c_float()
results = []
for vec in (16, 32):
    for unroll in (1, 2, 4, 8):
        for threads in (64, 128, 256, 512):
            row = {"vec": vec, "unroll": unroll, "threads": threads}
            for name, d in (("encode", 0), ("decode", 1)):
                _lib.check(lib.cn_set_tuning(d, vec, unroll, threads))
                best = 1e30
                for _ in range(args.reps + 1):
                    if d == 0:
                        _lib.check(lib.cn_time_encode_device(d_n.data_ptr(), L, d_bits.data_ptr(), args.iters, ctypes.byref(ms)))
                    else:
                        _lib.check(lib.cn_time_decode_device(d_bits.data_ptr(), W, L, d_out.data_ptr(), args.iters, ctypes.byref(ms)))
                    best = min(best, ms.value / args.iters)
                row[name + "_ms"] = round(best, 4)
                row[name + "_gbs"] = round(1.25 * L / (best * 1e-3) / 1e9, 1)
            results.append(row)
            print(json.dumps(row), flush=True)
be = max(results, key=lambda r: r["encode_gbs"])
bd = max(results, key=lambda r: r["decode_gbs"])
print(json.dumps({"best_encode": be, "best_decode": bd}), flush=True)
