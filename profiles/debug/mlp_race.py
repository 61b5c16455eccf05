"""Repeat the fused-MLP backward on one shape and print where results differ from the unfused kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import revisiting_at_b200  # noqa
from revisiting_at_b200 import _abi as abi
BF16 = torch.bfloat16
dev = torch.device('cuda:0')
C, M = 96, 148 * 128 * 2 + 64
g = torch.Generator(device='cuda').manual_seed(1)
r = lambda *s, scale=1.: (torch.randn(*s, generator=g, device=dev) * scale).to(BF16)
dout, w1, w2 = r(M, C), r(4 * C, C, scale=C ** -0.5), r(C, 4 * C, scale=(4 * C) ** -0.5)
b1 = torch.randn(4 * C, generator=g, device=dev) * 0.5
z = r(M, 4 * C, scale=1.5)
w2t, w1t = w2.t().contiguous(), w1.t().contiguous()
da_u, dz_u, dt2_u = torch.empty(M, 4 * C, device=dev, dtype=BF16), torch.empty(M, 4 * C, device=dev, dtype=BF16), torch.empty(M, C, device=dev, dtype=BF16)
abi.gemm_bf16(dout, w2t, da_u, abi.EPI_NONE)
abi.bias_gelu_bwd(da_u, z, b1, dz_u, None)
abi.gemm_bf16(dz_u, w1t, dt2_u, abi.EPI_NONE)
torch.cuda.synchronize()
for mode in ('dz_out', 'no_dz', 'fwd'):
    for it in range(10):
        dz = torch.full((M, 4 * C), float('nan'), device=dev, dtype=BF16)
        dt2 = torch.full((M, C), float('nan'), device=dev, dtype=BF16)
        if mode == 'fwd':
            zz = torch.full((M, 4 * C), float('nan'), device=dev, dtype=BF16)
            abi.mlp_fused(dout, w1, w2, b1, zz, dt2, bias2=None, residual=dout, p_out=dz)
            torch.cuda.synchronize()
            ref_z = torch.empty_like(zz); abi.gemm_bf16(dout, w1, ref_z, abi.EPI_NONE)
            ref_a = torch.empty_like(zz); abi.bias_gelu_fwd(ref_z, b1, ref_a)
            bad = ((zz.float() - ref_z.float()).abs() > 0.02 + 0.01 * ref_z.float().abs()) | ((dz.float() - ref_a.float()).abs() > 0.02 + 0.02 * ref_a.float().abs())
            bad2 = torch.zeros(M, C, dtype=torch.bool, device=dev)
        else:
            abi.mlp_fused(dout, w2t, w1t, b1, z, dt2, p_out=dz if mode == 'dz_out' else None, backward=True)
            torch.cuda.synchronize()
            bad = ((dz.float() - dz_u.float()).abs() > 0.02 + 0.02 * dz_u.float().abs()) if mode == 'dz_out' else torch.zeros(M, 4 * C, dtype=torch.bool, device=dev)
            bad2 = (dt2.float() - dt2_u.float()).abs() > 0.02 + 0.02 * dt2_u.float().abs()
        nb, nb2 = int(bad.sum()), int(bad2.sum())
        msg = f'{mode} it {it}: hidden bad {nb}, out bad {nb2}'
        if nb:
            idx = bad.nonzero()
            rows, cols = idx[:, 0], idx[:, 1]
            tiles = torch.unique(rows // 128).tolist()
            msg += f' | tiles {tiles[:8]} (local {[t // 148 for t in tiles[:8]]}) rows%128 {torch.unique(rows % 128).tolist()[:40]} chunks {torch.unique(cols // 64).tolist()} cols%64 {torch.unique(cols % 64).tolist()[:40]}'
        if nb2:
            idx = bad2.nonzero()
            msg += f' | out tiles {torch.unique(idx[:, 0] // 128).tolist()[:8]} rows%128 {torch.unique(idx[:, 0] % 128).tolist()[:20]}'
        print(msg, flush=True)

# ---- what do the wrong values look like?
print('--- provenance of wrong dz values')
for it in range(8):
    dz = torch.full((M, 4 * C), float('nan'), device=dev, dtype=BF16)
    dt2 = torch.full((M, C), float('nan'), device=dev, dtype=BF16)
    abi.mlp_fused(dout, w2t, w1t, b1, z, dt2, p_out=dz, backward=True)
    torch.cuda.synchronize()
    bad = ((dz.float() - dz_u.float()).abs() > 0.02 + 0.02 * dz_u.float().abs())
    if not bad.any():
        continue
    idx = bad.nonzero()
    rows, cols = idx[:, 0], idx[:, 1]
    got = dz[rows, cols].float()
    def frac(cand):
        return float(((got - cand.float()).abs() <= 0.02 + 0.02 * cand.float().abs()).float().mean())
    nan_frac = float(torch.isnan(got).float().mean())
    cands = {'same tile chunk-2 (prev P[1])': dz_u[rows, cols - 128], 'same tile chunk+2': dz_u[rows, (cols + 128) % (4 * C)],
             'prev tile last chunk1 (g=5)': dz_u[(rows - 148 * 128).clamp(min=0), cols + 128 if True else cols],
             'da (no gelu grad)': da_u[rows, cols], 'zero': torch.zeros_like(got)}
    # recompute with z of other chunks (stale z) and da of other chunks
    zz = z.float(); daf = da_u.float()
    def dzf(da_, z_, c_):
        zb = (z_ + b1[c_]).requires_grad_()
        (gp,) = torch.autograd.grad(torch.nn.functional.gelu(zb).sum(), zb)
        return (da_ * gp)
    cands['da(9) with z of chunk-2'] = dzf(daf[rows, cols], zz[rows, cols - 128], cols)
    cands['da(9) with z of chunk+2'] = dzf(daf[rows, cols], zz[rows, (cols + 128) % (4 * C)], cols)
    cands['da of chunk-2 with z(9)'] = dzf(daf[rows, cols - 128], zz[rows, cols], cols)
    cands['da of chunk+2 with z(9)'] = dzf(daf[rows, (cols + 128) % (4 * C)], zz[rows, cols], cols)
    print(f'it {it}: {int(bad.sum())} bad, nan {nan_frac:.2f}; ' + '; '.join(f'{k}: {frac(v):.2f}' for k, v in cands.items()), flush=True)
