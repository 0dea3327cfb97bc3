"""Time the base-5 codec kernels (n_to_bits2 / bits_to_n2) on device-resident buffers.
Algorithmic traffic: 27 ASCII bytes + 8 packed bytes per 27 nucleotides = 35/27 = 1.296 B/nt per direction."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cute_nucleotides_b200 as cn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nucleotides", type=int, default=10 << 30)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
L = args.nucleotides
W = cn.words2_for_len(L)
d_n = cn.generate2_device(torch.empty(L, dtype=torch.uint8, device="cuda"), 0, 1, 12)
d_bits = torch.empty(W, dtype=torch.int64, device="cuda")
d_out = torch.empty(L, dtype=torch.uint8, device="cuda")
for _ in range(3):
    cn.encode2_device(d_n, out=d_bits)
    cn.decode2_device(d_bits, L, out=d_out)
torch.cuda.synchronize()
res = {}
for name, fn in (("encode2", lambda: cn.encode2_device(d_n, out=d_bits)), ("decode2", lambda: cn.decode2_device(d_bits, L, out=d_out))):
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / args.iters)
    bytes_ = L + 8 * W
    res[name] = {"ms": round(best, 4), "nt_per_s": L / (best * 1e-3), "gbs": round(bytes_ / (best * 1e-3) / 1e9, 1)}
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
for k in res:
    res[k]["frac_of_measured_copy_peak"] = round(res[k]["gbs"] / peak, 4)
print(json.dumps({"codec": "base-5 (n_to_bits2)", "nucleotides": L, "bytes_per_nt": 35 / 27, **res}))
