"""One launch of each segmented kernel on ~1 G nucleotides, two shapes, for ncu.  Harness only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cute_nucleotides_b200 as cn
dev = torch.device("cuda", 0)
for lo, hi, n_seq in ((40000, 40001, 25000), (150, 301, 4000000)):
    lens = torch.randint(lo, hi, (n_seq,), device=dev, dtype=torch.int64)
    offs = torch.zeros(n_seq + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, dim=0, out=offs[1:])
    total = int(offs[-1].item())
    d_n = cn.generate_device(torch.empty(total, dtype=torch.uint8, device=dev), 0, 1, 10)
    woffs = cn.segment_word_offsets(offs)
    for _ in range(2):
        words, _w = cn.encode_segmented_device(d_n, offs, woffs)
        back = cn.decode_segmented_device(words, offs, woffs)
    torch.cuda.synchronize()
print("done")
