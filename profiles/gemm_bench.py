"""tcgen05 GEMM vs cuBLAS (torch.matmul) on the pwconv shapes of ConvNeXt-T at batch 128 (bf16, fp32 accumulate)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import revisiting_at_b200  # noqa: F401
from revisiting_at_b200 import _abi

dev = torch.device('cuda:0')
BF16 = torch.bfloat16


def t(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


print(f'{"M":>8} {"N":>6} {"K":>6} | {"tcgen05 us":>10} {"TF/s":>7} | {"cuBLAS us":>10} {"TF/s":>7} | {"fused gelu us":>13} {"cuBLAS+gelu us":>14}')
for M, C in ((401408, 96), (100352, 192), (25088, 384), (6272, 768)):
    for (N, K) in ((4 * C, C), (C, 4 * C)):
        a = torch.randn(M, K, device=dev).to(BF16)
        w = torch.randn(N, K, device=dev).to(BF16)
        bias = torch.randn(N, device=dev)
        c = torch.empty(M, N, device=dev, dtype=BF16)
        c2 = torch.empty_like(c)
        fl = 2.0 * M * N * K
        t1 = t(lambda: _abi.gemm_bf16(a, w, c, _abi.EPI_NONE))
        t2 = t(lambda: torch.matmul(a, w.t(), out=c))
        t3 = t(lambda: _abi.gemm_bf16(a, w, c, _abi.EPI_BIAS_GELU, bias=bias, c2=c2))
        def unfused():
            torch.matmul(a, w.t(), out=c)
            _abi.bias_gelu_fwd(c, bias, c2)
        t4 = t(unfused)
        # the backward's pair: dz = (dout W2g) * GELU'(z) fused in the epilogue vs tcgen05 GEMM + bias_gelu_bwd kernel
        z0 = torch.zeros(N, device=dev)
        t5 = t(lambda: _abi.gemm_bf16(a, w, c, _abi.EPI_GELU_GRAD, aux=c2))
        def unfused_bwd():
            _abi.gemm_bf16(a, w, c, _abi.EPI_NONE)
            _abi.bias_gelu_bwd(c, c2, z0, c, None)
        t6 = t(unfused_bwd) if N > K else float('nan')
        print(f'{M:8d} {N:6d} {K:6d} | {t1 * 1e3:10.1f} {fl / t1 / 1e9:7.1f} | {t2 * 1e3:10.1f} {fl / t2 / 1e9:7.1f} | {t3 * 1e3:13.1f} {t4 * 1e3:14.1f}'
              f' | gelu_grad fused {t5 * 1e3:7.1f} us, gemm + gelu_bwd {t6 * 1e3:7.1f} us')
