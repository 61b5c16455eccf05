"""GPU: the hand-written ConvNeXt layer kernels against plain PyTorch fp32 references of the same op
(bf16 storage tolerance written at each assert), and the whole ConvNeXt-T-CvSt engine against the CPU
model oracle."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


@pytest.fixture(scope='module')
def ops(cuda_dev):
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import ops as o
    return o


def _close(a, b, atol, rtol=2e-2):
    a, b = a.float(), b.float()
    err = (a - b).abs()
    ok = err <= atol + rtol * b.abs()
    assert bool(ok.all()), f'max err {err.max().item():.4g} (atol {atol}, rtol {rtol}), frac bad {(~ok).float().mean().item():.2e}'


@pytest.mark.parametrize('C', [48, 96, 144, 192, 384, 768, 1536])
@pytest.mark.parametrize('gelu', [False, True])
def test_layernorm_fwd_bwd(ops, cuda_dev, C, gelu):
    g = torch.Generator(device='cuda').manual_seed(C)
    M = 1031
    x = (torch.randn(M, C, generator=g, device=cuda_dev) * 2 + 0.5).to(BF16)
    w = torch.randn(C, generator=g, device=cuda_dev).requires_grad_()
    b = torch.randn(C, generator=g, device=cuda_dev).requires_grad_()
    dy = torch.randn(M, C, generator=g, device=cuda_dev).to(BF16)
    xr = x.float().requires_grad_()
    ref = F.layer_norm(xr, (C,), w, b, 1e-6)
    if gelu:
        ref = F.gelu(ref)
    rdx, rdw, rdb = torch.autograd.grad(ref, [xr, w, b], dy.float())
    xt = x.clone().requires_grad_()
    out = ops.layer_norm(xt, w, b, 1e-6, gelu)
    _close(out, ref, atol=2e-2)                    # bf16 output: half an ulp at |y| ~ 4 is 1.6e-2
    dx, dw, db = torch.autograd.grad(out, [xt, w, b], dy)
    _close(dx, rdx, atol=3e-2)
    _close(dw, rdw, atol=0.5, rtol=1e-2)           # sums of 1031 bf16-rounded products
    _close(db, rdb, atol=0.5, rtol=1e-2)
    (dx2,) = torch.autograd.grad(ops.layer_norm(xt, w, b, 1e-6, gelu), [xt], dy)   # input-grad only path
    assert torch.equal(dx2, dx)


@pytest.mark.parametrize('M,C', [(1031, 48), (200000, 96), (4100, 144)])
def test_layernorm_with_the_convolution_bias_folded_in(ops, cuda_dev, M, C):
    """y = GELU(LN(x + pre_bias)): the CvSt stem's conv bias rides in the LayerNorm kernels (plain and pipelined forms);
    its gradient is the column sum of dx (utils_architecture.py:205-211)."""
    g = torch.Generator(device='cuda').manual_seed(C + 1)
    x = (torch.randn(M, C, generator=g, device=cuda_dev) * 2 + 0.5).to(BF16)
    pb = torch.randn(C, generator=g, device=cuda_dev).requires_grad_()
    w = torch.randn(C, generator=g, device=cuda_dev).requires_grad_()
    b = torch.randn(C, generator=g, device=cuda_dev).requires_grad_()
    dy = torch.randn(M, C, generator=g, device=cuda_dev).to(BF16)
    xr = x.float().requires_grad_()
    ref = F.gelu(F.layer_norm(xr + pb, (C,), w, b, 1e-6))
    rdx, rdw, rdb, rdp = torch.autograd.grad(ref, [xr, w, b, pb], dy.float())
    xt = x.clone().requires_grad_()
    out = ops.layer_norm(xt, w, b, 1e-6, True, pre_bias=pb)
    _close(out, ref, atol=2e-2)
    dx, dw, db, dp = torch.autograd.grad(out, [xt, w, b, pb], dy)
    _close(dx, rdx, atol=3e-2)
    scale = M ** 0.5
    _close(dw, rdw, atol=2e-2 * scale, rtol=1e-2)
    _close(db, rdb, atol=2e-2 * scale, rtol=1e-2)
    _close(dp, rdp, atol=3e-2 * scale, rtol=2e-2)       # column sum of the bf16-rounded dx


@pytest.mark.parametrize('shape', [(2, 56, 56, 96), (3, 28, 28, 192), (2, 14, 14, 384), (5, 7, 7, 768), (1, 20, 23, 32),
                                   (2, 80, 80, 32), (3, 40, 40, 64), (9, 10, 10, 64), (2, 33, 47, 16), (1, 5, 3, 32),
                                   (2, 17, 96, 32)])
def test_dwconv7_fwd_dgrad_wgrad(ops, cuda_dev, shape):
    B, H, W, C = shape
    g = torch.Generator(device='cuda').manual_seed(H * C)
    x = torch.randn(B, H, W, C, generator=g, device=cuda_dev).to(BF16)
    w = (torch.randn(C, 1, 7, 7, generator=g, device=cuda_dev) * 0.1).requires_grad_()
    b = torch.randn(C, generator=g, device=cuda_dev).requires_grad_()
    dy = torch.randn(B, H, W, C, generator=g, device=cuda_dev).to(BF16)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_()
    ref = F.conv2d(xr, w, b, padding=3, groups=C)
    rdx, rdw, rdb = torch.autograd.grad(ref, [xr, w, b], dy.float().permute(0, 3, 1, 2))
    xt = x.clone().requires_grad_()
    out = ops._DwConv7.apply(xt, w, b)
    _close(out, ref.permute(0, 2, 3, 1), atol=2e-2)
    if C % 32:                                     # the weight-gradient kernel works on 32-channel groups
        with ops.input_grad_only():
            out = ops._DwConv7.apply(xt, w, b)
        (dx,) = torch.autograd.grad(out, [xt], dy)
        _close(dx, rdx.permute(0, 2, 3, 1), atol=2e-2)
        return
    dx, dw, db = torch.autograd.grad(out, [xt, w, b], dy)
    _close(dx, rdx.permute(0, 2, 3, 1), atol=2e-2)
    _close(dw, rdw, atol=1e-2 * (B * H * W) ** 0.5, rtol=1e-2)
    _close(db, rdb, atol=1e-2 * (B * H * W) ** 0.5, rtol=1e-2)


@pytest.mark.parametrize('shape', [(3, 56, 56, 96), (5, 14, 14, 64), (9, 7, 7, 32), (2, 80, 80, 16)])
def test_dwconv7_input_gradient_with_the_residual_join(cuda_dev, shape):
    """b200at_dwconv7_fwd(dy, flipped taps, add=dout): the block's backward join (models/convnext.py:49) in the same pass"""
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import _abi
    B, H, W, C = shape
    g = torch.Generator(device='cuda').manual_seed(7 * H + C)
    dy = torch.randn(B, H, W, C, generator=g, device=cuda_dev).to(BF16)
    res = torch.randn(B, H, W, C, generator=g, device=cuda_dev).to(BF16)
    w = torch.randn(C, 1, 7, 7, generator=g, device=cuda_dev) * 0.1
    wtf = w.reshape(C, 49).t().flip(0).contiguous()
    out = torch.full_like(dy, float('nan'))
    _abi.dwconv7_fwd(dy, wtf, None, out, add=res)
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w, padding=3, groups=C).permute(0, 2, 3, 1) + res.float()
    _close(out, ref, atol=3e-2)
    assert bool(torch.isfinite(out.float()).all())


@pytest.mark.parametrize('shape', ['2 14 14 64', '5 7 7 32 add', '9 10 10 48 add', '1 5 3 16', '4 14 14 384 add'])
def test_tensor_core_dwconv_on_the_narrow_maps_it_does_not_serve_by_default(cuda_dev, shape):
    """Maps narrower than 20 columns are dispatched to the FMA kernel (faster there); the tensor-core kernel still covers
    them (NB > 1 tiles, image masks) -- forced here through B200AT_DWM_MINW=1 in a fresh process (the switch is read once)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200AT_DWM_MINW='1')
    out = subprocess.run([sys.executable, os.path.join(root, 'profiles', 'debug', 'dwm_probe.py')] + shape.split(),
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'bad count 0 ' in out.stdout and "nan 0" in out.stdout, out.stdout


def test_bias_gelu_and_scale_residual(ops, cuda_dev):
    g = torch.Generator(device='cuda').manual_seed(1)
    M, N = 999, 384
    z = (torch.randn(M, N, generator=g, device=cuda_dev) * 2).to(BF16)
    bias = torch.randn(N, generator=g, device=cuda_dev).requires_grad_()
    dh = torch.randn(M, N, generator=g, device=cuda_dev).to(BF16)
    zr = z.float().requires_grad_()
    ref = F.gelu(zr + bias)
    rdz, rdb = torch.autograd.grad(ref, [zr, bias], dh.float())
    zt = z.clone().requires_grad_()
    out = ops._BiasGelu.apply(zt, bias)
    _close(out, ref, atol=1e-2)
    dz, db = torch.autograd.grad(out, [zt, bias], dh)
    _close(dz, rdz, atol=1e-2)
    _close(db, rdb, atol=0.3, rtol=1e-2)

    gamma = torch.randn(N, generator=g, device=cuda_dev).requires_grad_()
    res = torch.randn(M, N, generator=g, device=cuda_dev).to(BF16)
    rr = res.float().requires_grad_()
    ref = rr + gamma * (zr + bias)
    rg = torch.autograd.grad(ref, [zr, bias, gamma, rr], dh.float())
    rt = res.clone().requires_grad_()
    out = ops._ScaleResidual.apply(zt, bias, gamma, rt)
    _close(out, ref, atol=3e-2)
    got = torch.autograd.grad(out, [zt, bias, gamma, rt], dh)
    _close(got[0], rg[0], atol=2e-2)
    _close(got[1], rg[1], atol=0.5, rtol=1e-2)
    _close(got[2], rg[2], atol=0.5, rtol=1e-2)
    _close(got[3], rg[3], atol=1e-6)


def test_colsum(cuda_dev):
    from revisiting_at_b200 import _abi
    g = torch.Generator(device='cuda').manual_seed(5)
    for M, N in ((1000, 96), (4097, 384), (33, 3072), (1, 8)):
        a = torch.randn(M, N, generator=g, device=cuda_dev).to(BF16)
        out = torch.zeros(N, device=cuda_dev)
        _abi.colsum_bf16(a, out)
        _close(out, a.float().sum(0), atol=1e-3 * M ** 0.5, rtol=1e-5)


@pytest.mark.parametrize('C0', [48, 64, 96])
@pytest.mark.parametrize('shape,norm', [((2, 3, 64, 64), True), ((3, 3, 37, 45), False), ((1, 3, 224, 224), True)])
def test_stem0_direct_kernels(ops, cuda_dev, C0, shape, norm):
    """fused first stem stage (normalise -> conv3x3 s2 p1 -> LN -> GELU) and its input gradient vs fp32 torch"""
    g = torch.Generator(device='cuda').manual_seed(C0 + shape[2])
    x = torch.rand(*shape, generator=g, device=cuda_dev)
    cw = (torch.randn(C0, 3, 3, 3, generator=g, device=cuda_dev) * 0.3)
    cb = torch.randn(C0, generator=g, device=cuda_dev) * 0.1
    lw = 1 + 0.2 * torch.randn(C0, generator=g, device=cuda_dev)
    lb = 0.2 * torch.randn(C0, generator=g, device=cuda_dev)
    mean = torch.tensor([0.485, 0.456, 0.406], device=cuda_dev).view(1, 3, 1, 1) if norm else None
    std = torch.tensor([0.229, 0.224, 0.225], device=cuda_dev).view(1, 3, 1, 1) if norm else None
    xr = x.clone().requires_grad_()
    h = F.conv2d((xr - mean) / std if norm else xr, cw, cb, stride=2, padding=1).permute(0, 2, 3, 1)
    ref = F.gelu(F.layer_norm(h, (C0,), lw, lb, 1e-6))
    dy = torch.randn(ref.shape, generator=g, device=cuda_dev).to(BF16)
    (rdx,) = torch.autograd.grad(ref, [xr], dy.float())
    xt = x.clone().requires_grad_()
    with ops.input_grad_only():
        out = ops.stem_layer(xt, cw, cb, lw, lb, 2, True, mean, std)
    assert out.dtype == BF16 and out.shape == ref.shape
    assert out.grad_fn is not None and type(out.grad_fn).__name__.startswith('_Stem0')
    _close(out, ref, atol=1e-2)                     # fp32 math, one bf16 rounding of the result
    (dx,) = torch.autograd.grad(out, [xt], dy)
    assert dx.dtype == torch.float32 and dx.shape == x.shape
    _close(dx, rdx, atol=2e-3 * rdx.abs().max().item(), rtol=1e-3)   # fp32 math on the same bf16 dy
    with torch.no_grad():
        assert torch.equal(ops.stem_layer(x, cw, cb, lw, lb, 2, True, mean, std), out)


@pytest.mark.parametrize('tc', ['', 'residual,dgrad1', 'residual,dgrad1,gelu,gelu_grad'])
@pytest.mark.parametrize('shape', [(2, 28, 28, 96), (3, 7, 7, 192)])
def test_block_function_fwd_bwd(ops, cuda_dev, shape, tc, monkeypatch):
    """whole-block autograd node vs the fp32 reference block (all gradients), and the input-grad-only mode;
    `tc` = which pwconv GEMMs run on the tcgen05 kernel (the others: cuBLAS + elementwise kernels)"""
    monkeypatch.setattr(ops, 'TCGEN05', set(filter(None, tc.split(','))))
    B, H, W, C = shape
    g = torch.Generator(device='cuda').manual_seed(C)
    rnd = lambda *s, k=1.0: (torch.randn(*s, generator=g, device=cuda_dev) * k)
    x = rnd(B, H, W, C).to(BF16)
    P = dict(dw_w=rnd(C, 1, 7, 7, k=0.1), dw_b=rnd(C, k=0.1), ln_w=1 + rnd(C, k=0.1), ln_b=rnd(C, k=0.1),
             w1=rnd(4 * C, C, k=0.05), b1=rnd(4 * C, k=0.1), w2=rnd(C, 4 * C, k=0.05), b2=rnd(C, k=0.1),
             gamma=rnd(C, k=0.5))
    P = {k: v.requires_grad_() for k, v in P.items()}
    dout = rnd(B, H, W, C).to(BF16)

    xr = x.float().requires_grad_()
    h = F.conv2d(xr.permute(0, 3, 1, 2), P['dw_w'], P['dw_b'], padding=3, groups=C).permute(0, 2, 3, 1)
    h = F.layer_norm(h, (C,), P['ln_w'], P['ln_b'], 1e-6)
    h = F.linear(F.gelu(F.linear(h, P['w1'], P['b1'])), P['w2'], P['b2'])
    ref = xr + P['gamma'] * h
    names = list(P)
    rg = torch.autograd.grad(ref, [xr] + [P[n] for n in names], dout.float())

    xt = x.clone().requires_grad_()
    out = ops.convnext_block(xt, *[P[n] for n in names])
    _close(out, ref, atol=4e-2)
    got = torch.autograd.grad(out, [xt] + [P[n] for n in names], dout)
    _close(got[0], rg[0], atol=4e-2)
    for n, a, r in zip(names, got[1:], rg[1:]):
        cs = F.cosine_similarity(a.flatten().float(), r.flatten(), dim=0).item()
        assert cs > 0.995, (n, cs)
        assert abs(a.float().norm().item() / r.norm().item() - 1) < 0.03, n
    with ops.input_grad_only():
        out2 = ops.convnext_block(xt, *[P[n] for n in names])
    (dx2,) = torch.autograd.grad(out2, [xt], dout)
    assert torch.equal(dx2, got[0])


@pytest.mark.parametrize('shape,Co', [((2, 56, 56, 96), 192), ((3, 14, 14, 384), 768), ((1, 6, 10, 128), 256)])
def test_downsample_gemm_matches_reference(ops, cuda_dev, shape, Co):
    """LayerNorm -> Conv2d(k=2, s=2) (models/convnext.py:79-82) through the patch-layout LN kernel + tcgen05 GEMM"""
    B, H, W, C = shape
    g = torch.Generator(device='cuda').manual_seed(H + C)
    rnd = lambda *s, k=1.0: (torch.randn(*s, generator=g, device=cuda_dev) * k)
    x = rnd(B, H, W, C).to(BF16)
    P = dict(ln_w=1 + rnd(C, k=0.1), ln_b=rnd(C, k=0.1), cw=rnd(Co, C, 2, 2, k=0.05), cb=rnd(Co, k=0.1))
    P = {k: v.requires_grad_() for k, v in P.items()}
    names = list(P)
    dout = rnd(B, H // 2, W // 2, Co).to(BF16)
    xr = x.float().requires_grad_()
    t = F.layer_norm(xr, (C,), P['ln_w'], P['ln_b'], 1e-6)
    ref = F.conv2d(t.permute(0, 3, 1, 2), P['cw'], P['cb'], stride=2).permute(0, 2, 3, 1)
    rg = torch.autograd.grad(ref, [xr] + [P[n] for n in names], dout.float())
    xt = x.clone().requires_grad_()
    out = ops.downsample(xt, *[P[n] for n in names])
    assert out.shape == ref.shape
    _close(out, ref, atol=4e-2)
    got = torch.autograd.grad(out, [xt] + [P[n] for n in names], dout)
    _close(got[0], rg[0], atol=4e-2, rtol=3e-2)
    for n, a, r in zip(names, got[1:], rg[1:]):
        assert a.shape == r.shape, n
        cs = F.cosine_similarity(a.flatten().float(), r.flatten(), dim=0).item()
        assert cs > 0.995, (n, cs)
        assert abs(a.float().norm().item() / r.norm().item() - 1) < 0.03, n
    with ops.input_grad_only():
        out2 = ops.downsample(xt, *[P[n] for n in names])
    (dx2,) = torch.autograd.grad(out2, [xt], dout)
    assert torch.equal(dx2, got[0])


def test_convnext_engine_matches_oracle(cuda_dev):
    """ConvNeXt-T-CvSt, same seed-0 weights: bf16 engine on the GPU vs fp32 oracle on the CPU.
    Tolerances: logits 5e-2 absolute (|logit| ~ 1, ~40 bf16 layers); input gradient cosine >= 0.98."""
    from revisiting_at_b200 import convnext
    from oracle import convnext_oracle as co
    o = co.build('convnext_tiny', normalize=True, seed=0)
    m = convnext.build('convnext_tiny', normalize=True, seed=1)
    m.load_state_dict(o.state_dict())
    m = m.to(cuda_dev).eval()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(4, 3, 64, 64, generator=g)
    y = torch.randint(0, 1000, (4,), generator=g)
    xo = x.clone().requires_grad_()
    lo = o(xo)
    (go,) = torch.autograd.grad(F.cross_entropy(lo, y, reduction='sum'), xo)
    xm = x.to(cuda_dev).requires_grad_()
    lm = m(xm)
    (gm,) = torch.autograd.grad(F.cross_entropy(lm.float(), y.to(cuda_dev), reduction='sum'), xm)
    assert (lm.float().cpu() - lo).abs().max() <= 5e-2, (lm.float().cpu() - lo).abs().max()
    cos = F.cosine_similarity(gm.cpu().flatten(1), go.flatten(1)).min().item()
    assert cos >= 0.98, cos
    assert all(p.grad is None for p in m.parameters())
    # the attack's evaluation mode (input-grad only: fused first stem stage, no weight gradients)
    from revisiting_at_b200 import ops as O
    xa = x.to(cuda_dev).requires_grad_()
    with O.input_grad_only():
        la = m(xa)
    (ga,) = torch.autograd.grad(F.cross_entropy(la.float(), y.to(cuda_dev), reduction='sum'), xa)
    assert (la.float().cpu() - lo).abs().max() <= 5e-2
    assert F.cosine_similarity(ga.cpu().flatten(1), go.flatten(1)).min().item() >= 0.98
    assert all(p.grad is None for p in m.parameters())
    # full backward (outer training step): every parameter receives a finite gradient
    m.train()
    loss = F.cross_entropy(m(x.to(cuda_dev)).float(), y.to(cuda_dev))
    loss.backward()
    o.train()
    F.cross_entropy(o(x), y).backward()
    od = dict(o.named_parameters())
    for n, p in m.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n
        ref = od[n].grad
        cs = F.cosine_similarity(p.grad.flatten().cpu().float(), ref.flatten(), dim=0).item()
        assert cs >= 0.95 or ref.abs().max() < 1e-7, (n, cs)


def test_normalize_nhwc_matches_the_eager_expression_bit_for_bit(cuda_dev):
    """(x - mean) / std -> bf16 -> channels_last in one kernel == torch's four passes (utils_architecture.py:86-98)"""
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import _abi
    g = torch.Generator(device='cuda').manual_seed(3)
    mean = torch.tensor([0.485, 0.456, 0.406], device=cuda_dev).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=cuda_dev).view(1, 3, 1, 1)
    for shape in ((5, 3, 33, 47), (16, 3, 224, 224)):
        x = torch.rand(*shape, generator=g, device=cuda_dev)
        y = torch.full((shape[0], shape[2], shape[3], 3), float('nan'), device=cuda_dev, dtype=torch.bfloat16)
        _abi.normalize_nhwc_bf16(x, [float(v) for v in mean.flatten()], [float(v) for v in std.flatten()], y)
        want = ((x - mean) / std).to(torch.bfloat16).permute(0, 2, 3, 1)
        assert torch.equal(y, want)
        _abi.normalize_nhwc_bf16(x, None, None, y)
        assert torch.equal(y, x.to(torch.bfloat16).permute(0, 2, 3, 1))


@pytest.mark.parametrize('shape,Co', [((2, 112, 112, 48), 96), ((3, 40, 40, 64), 96), ((1, 160, 160, 48), 96),
                                      ((2, 24, 56, 16), 32), ((5, 8, 8, 8), 16)])
def test_conv3x3s2_implicit_gemm(ops, cuda_dev, shape, Co):
    """the stems' second convolution (utils_architecture.py:205-211) as an implicit GEMM on the tcgen05 kernel against the
    library convolution of the same bf16 operands; gradients flow through the library calls of its backward"""
    g = torch.Generator(device='cuda').manual_seed(sum(shape) + Co)
    B, H, W, Ci = shape
    x = torch.randn(shape, generator=g, device=cuda_dev).to(torch.bfloat16).requires_grad_()
    w = (torch.randn(Co, Ci, 3, 3, generator=g, device=cuda_dev) * (9 * Ci) ** -0.5).requires_grad_()
    assert ops._conv3x3s2_ok(x, w)
    y = ops._Conv3x3S2.apply(x, w)
    xr = x.detach().float().requires_grad_()
    wr = w.detach().to(torch.bfloat16).float().requires_grad_()
    yr = F.conv2d(xr.permute(0, 3, 1, 2), wr, None, stride=2, padding=1).permute(0, 2, 3, 1)
    assert y.shape == yr.shape
    err = (y.detach().float() - yr.detach()).abs()
    assert float(err.max()) <= 2e-2 + 1e-2 * float(yr.detach().abs().max()), float(err.max())
    dy = torch.randn(y.shape, generator=g, device=cuda_dev).to(torch.bfloat16)
    y.backward(dy)
    yr.backward(dy.float())
    assert torch.nn.functional.cosine_similarity(x.grad.float().flatten(), xr.grad.flatten(), dim=0) > 0.999
    assert torch.nn.functional.cosine_similarity(w.grad.float().flatten(), wr.grad.flatten(), dim=0) > 0.999


@pytest.mark.parametrize('C0', [48, 96])
@pytest.mark.parametrize('shape', [(2, 3, 64, 64), (1, 3, 224, 224)])
def test_stem0_training_forward_keeps_what_the_backward_needs(ops, cuda_dev, C0, shape):
    """the fused first stem stage in the training forward (detached fp32 input, parameters require grad): output and all
    parameter gradients against fp32 torch autograd of conv -> LN -> GELU on the normalised input"""
    g = torch.Generator(device='cuda').manual_seed(C0 + shape[2] + 1)
    x = torch.rand(*shape, generator=g, device=cuda_dev)
    P = [(torch.randn(C0, 3, 3, 3, generator=g, device=cuda_dev) * 0.3), torch.randn(C0, generator=g, device=cuda_dev) * 0.1,
         1 + 0.2 * torch.randn(C0, generator=g, device=cuda_dev), 0.2 * torch.randn(C0, generator=g, device=cuda_dev)]
    mean = torch.tensor([0.485, 0.456, 0.406], device=cuda_dev).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=cuda_dev).view(1, 3, 1, 1)
    Pr = [p.clone().requires_grad_() for p in P]
    h = F.conv2d((x - mean) / std, Pr[0], Pr[1], stride=2, padding=1).permute(0, 2, 3, 1)
    ref = F.gelu(F.layer_norm(h, (C0,), Pr[2], Pr[3], 1e-6))
    dy = torch.randn(ref.shape, generator=g, device=cuda_dev).to(BF16)
    rg = torch.autograd.grad(ref, Pr, dy.float())
    Pt = [p.clone().requires_grad_() for p in P]
    out = ops.stem_layer(x, Pt[0], Pt[1], Pt[2], Pt[3], 2, True, mean, std)
    assert type(out.grad_fn).__name__.startswith('_Stem0Train')
    _close(out, ref, atol=1e-2)
    tg = torch.autograd.grad(out, Pt, dy)
    for name, a, b in zip(('conv weight', 'conv bias', 'ln weight', 'ln bias'), tg, rg):
        cos = F.cosine_similarity(a.float().flatten(), b.flatten(), dim=0).item()
        assert cos > 0.999, (name, cos)
        assert abs(a.float().norm().item() / b.norm().item() - 1) < 2e-2, name
