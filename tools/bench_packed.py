"""Time the packed-word ops (device-resident, CUDA events) -- harness for A/B runs (CN_HAMMING_UNROLL=1|2|4)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cute_nucleotides_b200 as cn
L = 10 << 30
W = cn.words_for_len(L)
a = cn.generate_words_device(torch.empty(W, dtype=torch.int64, device="cuda"), 0, 1)
b = cn.generate_words_device(torch.empty(W, dtype=torch.int64, device="cuda"), 0, 2)
res = torch.zeros(1, dtype=torch.int64, device="cuda")
def timed(fn, iters=30):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
out = {"unroll": os.environ.get("CN_HAMMING_UNROLL", "2")}
for name, fn in (("hamming", lambda: cn.hamming_device(a, b, L, result=res)), ("complement", lambda: cn.complement_device(a, L, out=b)),
                 ("reverse_complement", lambda: cn.reverse_complement_device(a, L, out=b))):
    ms = timed(fn)
    out[name] = {"ms": round(ms, 4), "gbs": round(0.5 * L / (ms * 1e-3) / 1e9, 1)}
print(json.dumps(out))
