"""One launch of every kernel family on a 2 GiB workload, for `ncu --set full` (tools/profile_r02.sh).  Harness only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cute_nucleotides_b200 as cn

L = 2 << 30
dev = torch.device("cuda", 0)
d_n = cn.generate_device(torch.empty(L, dtype=torch.uint8, device=dev), 0, 1, 10)
d_bits = torch.empty(cn.words_for_len(L), dtype=torch.int64, device=dev)
d_out = torch.empty(L, dtype=torch.uint8, device=dev)
counter = torch.zeros(1, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
cn.encode_device(d_n, out=d_bits)                                   # encode_kernel<32,1,256,0,0>   plain
cn.decode_device(d_bits, L, out=d_out)                              # decode_kernel<32,1,256,0>
cn.encode_checked_device(d_n, counter, out=d_bits)                  # encode_kernel<32,2,256,0,1>   + count
cn.encode_ex_device(d_n, cn.ENC_LUT_EXACT, out=d_bits)              # encode_kernel<32,2,256,0,2>   n_to_bits_lut-exact
other = cn.generate_words_device(torch.empty_like(d_bits), 0, 5)
res = cn.hamming_device(d_bits, other, L)                           # hamming_kernel<2>
cn.complement_device(d_bits, L, out=other)                          # complement_kernel
cn.reverse_complement_device(d_bits, L, out=other)                  # reverse_complement_vec_kernel
cn.generate2_device(d_n, 0, 1, 12)
d_bits2 = torch.empty(cn.words2_for_len(L), dtype=torch.int64, device=dev)
cn.encode2_device(d_n, out=d_bits2)                                 # b5_encode_kernel<1,0,0>
cn.encode2_ex_device(d_n, cn.ENC_COUNT, counter=counter, out=d_bits2)   # b5_encode_kernel<1,1,0>
cn.decode2_device(d_bits2, L, out=d_out)                            # b5_decode_kernel<1>
torch.cuda.synchronize()
print("done")
