"""Per-kernel timings of the ConvNeXt-T-CvSt engine at the bench shapes (batch 128, 224 px): every hand-written
layer kernel at each of the four stage shapes, with the figure that bounds it.

    python profiles/ops_bench.py [--batch 128] [--iters 10] [--once]

--once runs every kernel exactly one time after one warm-up launch (the target of an `ncu --set full` capture).
HBM-bound kernels report algorithmic GB/s (bytes every element must move once), the depthwise conv also its
fp32 FMA rate, the GEMMs TFLOP/s.  CUDA events on the launching stream; buffers rotate through > L2 of data.
"""
import argparse
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
import revisiting_at_b200  # noqa: E402,F401
from revisiting_at_b200 import _abi  # noqa: E402

BF16 = torch.bfloat16
dev = torch.device('cuda:0')
STAGES = ((56, 96), (28, 192), (14, 384), (7, 768))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--once', action='store_true')
    ap.add_argument('--only', default='')
    a = ap.parse_args()
    B = a.batch
    rows = []

    def timeit(name, fn, bytes_alg=0., flops=0., fma=0.):
        if a.only and not re.search(a.only, name):
            return
        if a.once:
            fn()
            return
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / a.iters * 1e3
        rows.append((name, us, bytes_alg / us / 1e3 if bytes_alg else 0., flops / us / 1e6 if flops else 0.,
                     fma / us / 1e6 if fma else 0.))

    g = torch.Generator(device='cuda').manual_seed(0)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g, device=dev) * scale).to(BF16)

    for H, C in STAGES:
        M = B * H * H
        tag = f'{H}x{H}x{C}'
        x, dy, y = rnd(B, H, H, C), rnd(B, H, H, C), torch.empty(B, H, H, C, device=dev, dtype=BF16)
        wt = torch.randn(49, C, generator=g, device=dev) * 0.1
        bias = torch.randn(C, generator=g, device=dev)
        el = M * C
        timeit(f'dwconv7_fwd {tag}', lambda: _abi.dwconv7_fwd(x, wt, bias, y), 4. * el, fma=49. * el)
        timeit(f'dwconv7_dgrad+add {tag}', lambda: _abi.dwconv7_fwd(dy, wt, None, y, add=x), 6. * el, fma=49. * el)
        dwt, db = torch.zeros(49, C, device=dev), torch.zeros(C, device=dev)
        timeit(f'dwconv7_wgrad {tag}', lambda: _abi.dwconv7_wgrad(x, dy, dwt, db), 4. * el, fma=49. * el)
        w, b = torch.randn(C, generator=g, device=dev), torch.randn(C, generator=g, device=dev)
        mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
        timeit(f'ln_fwd {tag}', lambda: _abi.ln_fwd(x, w, b, y, mean, rstd, 1e-6, False), 4. * el)
        timeit(f'ln_bwd(dx) {tag}', lambda: _abi.ln_bwd(dy, x, w, b, mean, rstd, y, None, None, False), 6. * el)
        dw_, db_ = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        timeit(f'ln_bwd(dx,dw,db) {tag}', lambda: _abi.ln_bwd(dy, x, w, b, mean, rstd, y, dw_, db_, False), 6. * el)
        N4 = 4 * C
        z, hh, dh = rnd(M, N4), torch.empty(M, N4, device=dev, dtype=BF16), rnd(M, N4)
        b4 = torch.randn(N4, generator=g, device=dev)
        timeit(f'bias_gelu_fwd {tag}x4', lambda: _abi.bias_gelu_fwd(z, b4, hh), 4. * M * N4)
        timeit(f'bias_gelu_bwd {tag}x4', lambda: _abi.bias_gelu_bwd(dh, z, b4, hh), 6. * M * N4)
        dbias = torch.zeros(N4, device=dev)
        timeit(f'bias_gelu_bwd+dbias {tag}x4', lambda: _abi.bias_gelu_bwd(dh, z, b4, hh, dbias), 6. * M * N4)
        timeit(f'colsum {tag}', lambda: _abi.colsum_bf16(x.view(M, C), db), 2. * el)
        w1 = rnd(N4, C, scale=0.05)
        w2 = rnd(C, N4, scale=0.05)
        x2, xr = x.view(M, C), rnd(M, C)
        fl = 2. * M * C * N4
        out = torch.empty(M, C, device=dev, dtype=BF16)
        timeit(f'gemm pwconv1 NONE [{M}x{N4}x{C}]', lambda: _abi.gemm_bf16(x2, w1, hh, _abi.EPI_NONE), 2. * (M * C + M * N4), fl)
        timeit(f'cublas pwconv1     [{M}x{N4}x{C}]', lambda: torch.matmul(x2, w1.t(), out=hh), 2. * (M * C + M * N4), fl)
        timeit(f'gemm pwconv1 BIAS_GELU(+z) [{M}x{N4}x{C}]',
               lambda: _abi.gemm_bf16(x2, w1, hh, _abi.EPI_BIAS_GELU, bias=b4, c2=z), 2. * (M * C + 2 * M * N4), fl)
        timeit(f'gemm pwconv2 RESIDUAL [{M}x{C}x{N4}]',
               lambda: _abi.gemm_bf16(dh, w2, out, _abi.EPI_RESIDUAL, bias=bias, aux=xr), 2. * (M * N4 + 2 * M * C), fl)
        timeit(f'cublas pwconv2       [{M}x{C}x{N4}]', lambda: torch.matmul(dh, w2.t(), out=out), 2. * (M * N4 + M * C), fl)
        timeit(f'gemm dz GELU_GRAD [{M}x{N4}x{C}]',
               lambda: _abi.gemm_bf16(x2, w1, hh, _abi.EPI_GELU_GRAD, aux=z), 2. * (M * C + 2 * M * N4), fl)
        timeit(f'gemm dgrad1 NONE [{M}x{C}x{N4}]', lambda: _abi.gemm_bf16(dh, w2, out, _abi.EPI_NONE), 2. * (M * N4 + M * C), fl)
        if _abi.mlp_fused_supported(C):
            w2t = w2.t().contiguous()       # [4C, C]
            w1t = w1.t().contiguous()       # [C, 4C]
            timeit(f'mlp fused fwd (z out) [{M}x{C}x{N4}]',
                   lambda: _abi.mlp_fused(x2, w1, w2, b4, z, out, bias2=bias, residual=xr), 2. * (3 * M * C + M * N4), 2 * fl)
            timeit(f'mlp fused fwd (z,a out) [{M}x{C}x{N4}]',
                   lambda: _abi.mlp_fused(x2, w1, w2, b4, z, out, bias2=bias, residual=xr, p_out=hh), 2. * (3 * M * C + 2 * M * N4), 2 * fl)
            timeit(f'mlp fused bwd (z in) [{M}x{C}x{N4}]',
                   lambda: _abi.mlp_fused(x2, w2t, w1t, b4, z, out, backward=True), 2. * (2 * M * C + M * N4), 2 * fl)
            timeit(f'mlp fused bwd (z in, dz out) [{M}x{C}x{N4}]',
                   lambda: _abi.mlp_fused(x2, w2t, w1t, b4, z, out, p_out=hh, backward=True), 2. * (2 * M * C + 2 * M * N4), 2 * fl)
        del x, dy, y, z, hh, dh
        torch.cuda.empty_cache()

    # stem LayerNorm + GELU shapes
    for H, C in ((112, 48), (56, 96)):
        M = B * H * H
        x, y = rnd(M, C), torch.empty(M, C, device=dev, dtype=BF16)
        w, b = torch.randn(C, generator=g, device=dev), torch.randn(C, generator=g, device=dev)
        mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
        timeit(f'ln_fwd+gelu stem {H}x{H}x{C}', lambda: _abi.ln_fwd(x, w, b, y, mean, rstd, 1e-6, True), 4. * M * C)
        timeit(f'ln_bwd+gelu stem {H}x{H}x{C}', lambda: _abi.ln_bwd(y, x, w, b, mean, rstd, y, None, None, True), 6. * M * C)
    # fused first stem stage (normalise -> conv3x3 s2 -> LN -> GELU) and its input gradient
    H, C0 = 224, 48
    x = torch.rand(B, 3, H, H, generator=g, device=dev)
    wk = torch.randn(27, C0, generator=g, device=dev) * 0.2
    cb, lw, lb = (torch.randn(C0, generator=g, device=dev) * 0.1 for _ in range(3))
    y = torch.empty(B, H // 2, H // 2, C0, device=dev, dtype=BF16)
    dy, dx = rnd(B, H // 2, H // 2, C0), torch.empty_like(x)
    mean3, std3 = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    npx = B * (H // 2) ** 2
    timeit(f'stem0_fwd 3x{H}x{H} -> {H // 2}x{H // 2}x{C0}', lambda: _abi.stem0_fwd(x, mean3, std3, wk, cb, lw, lb, y),
           4. * x.numel() + 2. * npx * C0, fma=27. * C0 * npx)
    timeit(f'stem0_bwd_input {H // 2}x{H // 2}x{C0} -> 3x{H}x{H}',
           lambda: _abi.stem0_bwd_input(dy, x, mean3, std3, wk, cb, lw, lb, dx), 8. * x.numel() + 2. * npx * C0,
           fma=2 * 27. * C0 * npx)

    # second stem convolution (3x3 s2, 48 -> 96 at 112 x 112): implicit GEMM on the tcgen05 kernel vs the library convolution
    from revisiting_at_b200 import ops as _ops
    Hs, Ci, Co = 112, 48, 96
    xs = rnd(B, Hs, Hs, Ci)
    cw = torch.randn(Co, Ci, 3, 3, generator=g, device=dev) * 0.05
    wk2 = _ops._conv3x3s2_wk(cw)
    ys = torch.empty(B, Hs // 2, Hs // 2, Co, device=dev, dtype=BF16)
    cwb = cw.to(BF16).contiguous(memory_format=torch.channels_last)
    xs_nchw = xs.permute(0, 3, 1, 2)
    npo = B * (Hs // 2) ** 2
    fl2 = 2. * npo * Co * 9 * Ci
    timeit(f'conv3x3s2 implicit GEMM {Hs}x{Hs}x{Ci} -> {Co}', lambda: _abi.conv3x3s2_fwd(xs, wk2, ys), 2. * xs.numel() + 2. * npo * Co, fl2)
    timeit(f'conv3x3s2 library       {Hs}x{Hs}x{Ci} -> {Co}', lambda: F.conv2d(xs_nchw, cwb, None, stride=2, padding=1), 2. * xs.numel() + 2. * npo * Co, fl2)

    if not a.once:
        print(f'{"kernel":58s} {"us":>9} {"GB/s alg":>9} {"TFLOP/s":>8} {"TFMA/s":>7}')
        for name, us, gbs, tf, tfma in rows:
            print(f'{name:58s} {us:9.1f} {gbs:9.0f} {tf:8.1f} {tfma:7.2f}')


if __name__ == '__main__':
    main()
