"""Per-kernel timings of the ViT-S-CvSt block at BASELINE config 3 shapes (B=256, 197 tokens, 6 heads x 64):
hand-written attention forward/backward next to torch's scaled_dot_product_attention (library), CUDA events,
L2 flushed between repetitions by the 155 MB hidden tensors in flight.   python profiles/vit_bench.py"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import revisiting_at_b200  # noqa: E402,F401
from revisiting_at_b200 import _abi, ops  # noqa: E402

dev = torch.device('cuda:0')
B, N, H, D = 256, 197, 6, 384
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


ONCE = '--once' in sys.argv          # one launch per kernel (for ncu)


def timeit(fn, reps=10):
    if ONCE:
        fn(); torch.cuda.synchronize(); return float('nan')
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


g = torch.Generator(device='cuda').manual_seed(0)
qkv = torch.randn(B, N, 3 * D, generator=g, device=dev).to(torch.bfloat16)
d_o = torch.randn(B, N, D, generator=g, device=dev).to(torch.bfloat16)
o = torch.empty(B, N, D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B * H * N, device=dev, dtype=torch.float32)
dqkv = torch.empty_like(qkv)
fl_f = 4 * B * H * N * N * 64
print(f'{"kernel":44s} {"us":>8s} {"TFLOP/s":>8s}')
t = timeit(lambda: _abi.attn_fwd(qkv, o, lse, H, 0.125))
print(f'{"attn_fwd (mma.sync, hand-written)":44s} {t:8.1f} {fl_f / t / 1e6:8.1f}')
t = timeit(lambda: _abi.attn_bwd(qkv, o, d_o, lse, dqkv, H, 0.125))
print(f'{"attn_bwd (mma.sync, hand-written)":44s} {t:8.1f} {2.5 * fl_f / t / 1e6:8.1f}')
if ONCE:
    sys.exit(0)
q, k, v = (z.contiguous().requires_grad_() for z in qkv.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0))
t = timeit(lambda: F.scaled_dot_product_attention(q, k, v))
print(f'{"torch SDPA fwd (library, pre-permuted q/k/v)":44s} {t:8.1f} {fl_f / t / 1e6:8.1f}')
out = F.scaled_dot_product_attention(q, k, v)
go = torch.randn_like(out)
t = timeit(lambda: torch.autograd.grad(out, (q, k, v), go, retain_graph=True))
print(f'{"torch SDPA bwd (library)":44s} {t:8.1f} {2.5 * fl_f / t / 1e6:8.1f}')
M = B * N
x = torch.randn(M, D, generator=g, device=dev).to(torch.bfloat16)
w = {n: (torch.randn(s, generator=g, device=dev) * 0.05).to(torch.bfloat16) for n, s in
     (('qkv', (3 * D, D)), ('proj', (D, D)), ('w1', (4 * D, D)), ('w2', (D, 4 * D)))}
for n, a in (('qkv', x), ('proj', x), ('w1', x), ('w2', torch.randn(M, 4 * D, generator=g, device=dev).to(torch.bfloat16))):
    c = torch.empty(M, w[n].shape[0], device=dev, dtype=torch.bfloat16)
    fl = 2 * M * w[n].shape[0] * w[n].shape[1]
    t = timeit(lambda: _abi.gemm_bf16(a, w[n], c))
    t2 = timeit(lambda: torch.matmul(a, w[n].t(), out=c))
    print(f'{"gemm " + n + " tcgen05 / cuBLAS":44s} {t:8.1f} {fl / t / 1e6:8.1f}   | {t2:8.1f} {fl / t2 / 1e6:8.1f}')
