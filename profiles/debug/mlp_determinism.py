"""Run-to-run bit-exactness of the fused MLP kernels and exact agreement with the unfused kernels, per shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import revisiting_at_b200  # noqa
from revisiting_at_b200 import _abi as abi
BF16 = torch.bfloat16
dev = torch.device('cuda:0')
for C, M in ((96, 148 * 128 * 2 + 64), (128, 128 * 150 + 9), (192, 148 * 128 + 128 * 77 + 5), (128, 148 * 128 * 3 + 17)):
    g = torch.Generator(device='cuda').manual_seed(C + M)
    r = lambda *s, scale=1.: (torch.randn(*s, generator=g, device=dev) * scale).to(BF16)
    a_in, x, w1, w2 = r(M, C), r(M, C), r(4 * C, C, scale=C ** -0.5), r(C, 4 * C, scale=(4 * C) ** -0.5)
    b1, b2 = torch.randn(4 * C, generator=g, device=dev) * 0.5, torch.randn(C, generator=g, device=dev)
    z_in = r(M, 4 * C, scale=1.5)
    w2t, w1t = w2.t().contiguous(), w1.t().contiguous()
    e = lambda *s: torch.empty(*s, device=dev, dtype=BF16)
    # unfused
    zu, au, ou = e(M, 4 * C), e(M, 4 * C), e(M, C)
    abi.gemm_bf16(a_in, w1, zu, abi.EPI_NONE); abi.bias_gelu_fwd(zu, b1, au); abi.gemm_bf16(au, w2, ou, abi.EPI_RESIDUAL, bias=b2, aux=x)
    dau, dzu, dtu = e(M, 4 * C), e(M, 4 * C), e(M, C)
    abi.gemm_bf16(a_in, w2t, dau, abi.EPI_NONE); abi.bias_gelu_bwd(dau, z_in, b1, dzu, None); abi.gemm_bf16(dzu, w1t, dtu, abi.EPI_NONE)
    torch.cuda.synchronize()
    first = None
    for it in range(8):
        z, a, o = e(M, 4 * C), e(M, 4 * C), e(M, C)
        abi.mlp_fused(a_in, w1, w2, b1, z, o, bias2=b2, residual=x, p_out=a)
        dz, dt = e(M, 4 * C), e(M, C)
        abi.mlp_fused(a_in, w2t, w1t, b1, z_in, dt, p_out=dz, backward=True)
        dt_b = e(M, C)
        abi.mlp_fused(a_in, w2t, w1t, b1, z_in, dt_b, backward=True)
        torch.cuda.synchronize()
        cur = (z, a, o, dz, dt, dt_b)
        if first is None:
            first = cur
            ne = lambda p, q: int((p != q).sum())
            md = lambda p, q: float((p.float() - q.float()).abs().max())
            print(f'C={C} M={M}: vs unfused -- z != {ne(z, zu)}, a != {ne(a, au)}, out != {ne(o, ou)} (max {md(o, ou):.4g}); '
                  f'dz != {ne(dz, dzu)} (max {md(dz, dzu):.4g}), dt2 != {ne(dt, dtu)} (max {md(dt, dtu):.4g}), dt2(no dz out) vs dt2 != {ne(dt, dt_b)}', flush=True)
        else:
            same = [bool(torch.equal(p, q)) for p, q in zip(cur, first)]
            if not all(same):
                print(f'   repeat {it}: NOT identical to repeat 0: {same}', flush=True)
    print(f'   8 repeats done', flush=True)
