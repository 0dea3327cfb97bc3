"""Cost of fusing validation into the encode pass: time cn_encode_device vs cn_encode_checked_device at 10 GiB."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cute_nucleotides_b200 as cn
L = 10 << 30
d_n = cn.generate_device(torch.empty(L, dtype=torch.uint8, device="cuda"), 0, 1, 10)
d_bits = torch.empty(cn.words_for_len(L), dtype=torch.int64, device="cuda")
counter = torch.zeros(1, dtype=torch.int64, device="cuda")
res = {}
for name, fn in (("encode", lambda: cn.encode_device(d_n, out=d_bits)), ("encode_checked", lambda: cn.encode_checked_device(d_n, counter, out=d_bits))):
    for _ in range(3):
        fn()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    res[name] = {"ms": round(best, 4), "gbs": round(1.25 * L / (best * 1e-3) / 1e9, 1)}
res["invalid_counted"] = int(counter.item())
print(json.dumps(res))
