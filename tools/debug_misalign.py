import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch, _oracle
import cute_nucleotides_b200 as cn
from cute_nucleotides_b200 import _lib
lib = _lib.load(); orc = _oracle.Oracle()
size = (1 << 22) + 1234
n = orc.generate(size, seed=9, alphabet=10)
sl = n[3:]; L = sl.size
ref = orc.n_to_bits(sl, 'lut'); canon = np.frombuffer(orc.canonical(sl), dtype=np.uint8)
def first_bad(got):
    bad = np.nonzero(got != canon)[0]
    return (int(bad[0]), int(bad.size)) if bad.size else None
# device
d_bits = torch.from_numpy(ref.view(np.int64)).cuda()
d_out = torch.zeros(L + 64, dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for off in (0, 5, 11, 16):
    _lib.check(lib.cn_decode_device(d_bits.data_ptr(), ref.size, L, d_out.data_ptr() + off, st))
    print('device off', off, first_bad(d_out[off:off+L].cpu().numpy()))
h_bits = torch.from_numpy(ref.view(np.int64)).pin_memory()
h_out = torch.zeros(L + 64, dtype=torch.uint8).pin_memory()
for strategy in (0, 1):
    for off in (0, 5):
        lib.cn_set_host_strategy(strategy, 1 << 20)
        h_out.zero_()
        _lib.check(lib.cn_bits_to_n_host(h_bits.data_ptr(), ref.size, L, h_out.data_ptr() + off))
        print('host strategy', strategy, 'off', off, first_bad(h_out.numpy()[off:off+L]))
# zero copy direct device API on pinned pointers
for off in (0, 5):
    h_out.zero_()
    _lib.check(lib.cn_decode_device(h_bits.data_ptr(), ref.size, L, h_out.data_ptr() + off, st)); torch.cuda.synchronize()
    print('zero-copy device API off', off, first_bad(h_out.numpy()[off:off+L]))
