"""GPU: the ViT-S-CvSt path (BASELINE config 3) -- the hand-written attention kernel and the whole
transformer block against plain PyTorch fp32 references of the same op, the engine against the CPU model oracle
(oracle/vit_oracle.py; timm is un-vendored, so this parity is unpinned by the reference), and `apgd_train` on it."""
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


def _ref_attention(qkv, heads, scale):
    B, N, C3 = qkv.shape
    D = C3 // 3
    q, k, v = qkv.reshape(B, N, 3, heads, D // heads).permute(2, 0, 3, 1, 4).unbind(0)
    attn = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
    return (attn @ v).transpose(1, 2).reshape(B, N, D)


@pytest.mark.parametrize('B,N,H', [(3, 197, 6), (2, 208, 2), (2, 16, 1), (1, 1, 1), (2, 50, 3), (1, 130, 6)])
def test_attention_fwd_bwd(ops, cuda_dev, B, N, H):
    """softmax(q k^T / 8) v and its gradient; bf16 probabilities => 2e-2 absolute on O(1) outputs."""
    g = torch.Generator(device='cuda').manual_seed(N * 7 + H)
    qkv = torch.randn(B, N, 3 * H * 64, generator=g, device=cuda_dev).to(BF16)
    d_o = torch.randn(B, N, H * 64, generator=g, device=cuda_dev).to(BF16)
    qr = qkv.float().requires_grad_()
    ref = _ref_attention(qr, H, 0.125)
    (rd,) = torch.autograd.grad(ref, qr, d_o.float())
    qt = qkv.clone().requires_grad_()
    out = ops.attention(qt, H)
    _close(out, ref, atol=2e-2)
    (dq,) = torch.autograd.grad(out, qt, d_o)
    _close(dq, rd, atol=3e-2, rtol=3e-2)
    assert bool(torch.isfinite(dq.float()).all())


def test_attention_peaked_rows(ops, cuda_dev):
    """large score range (one dominant key per row, scores ~ +-60): the online softmax must not overflow"""
    B, N, H = 1, 197, 1
    g = torch.Generator(device='cuda').manual_seed(5)
    qkv = (torch.randn(B, N, 3 * 64, generator=g, device=cuda_dev) * 4).to(BF16)
    qr = qkv.float().requires_grad_()
    ref = _ref_attention(qr, H, 0.125)
    out = ops.attention(qkv.clone(), H)
    assert bool(torch.isfinite(out.float()).all())
    _close(out, ref, atol=6e-2, rtol=3e-2)


def test_vit_block_matches_reference(ops, cuda_dev):
    B, N, D, H = 3, 197, 384, 6
    g = torch.Generator(device='cuda').manual_seed(11)
    rnd = lambda *s, k=1.0: (torch.randn(*s, generator=g, device=cuda_dev) * k)
    x = rnd(B, N, D).to(BF16)
    P = dict(n1w=1 + rnd(D, k=0.1), n1b=rnd(D, k=0.1), wqkv=rnd(3 * D, D, k=0.05), bqkv=rnd(3 * D, k=0.1),
             wproj=rnd(D, D, k=0.05), bproj=rnd(D, k=0.1), n2w=1 + rnd(D, k=0.1), n2b=rnd(D, k=0.1),
             w1=rnd(4 * D, D, k=0.05), b1=rnd(4 * D, k=0.1), w2=rnd(D, 4 * D, k=0.03), b2=rnd(D, k=0.1))
    P = {k: v.requires_grad_() for k, v in P.items()}
    names = list(P)
    dout = rnd(B, N, D).to(BF16)

    xr = x.float().requires_grad_()
    t = F.layer_norm(xr, (D,), P['n1w'], P['n1b'], 1e-6)
    a = _ref_attention(F.linear(t, P['wqkv'], P['bqkv']), H, 0.125)
    x1 = xr + F.linear(a, P['wproj'], P['bproj'])
    t2 = F.layer_norm(x1, (D,), P['n2w'], P['n2b'], 1e-6)
    ref = x1 + F.linear(F.gelu(F.linear(t2, P['w1'], P['b1'])), P['w2'], P['b2'])
    rg = torch.autograd.grad(ref, [xr] + [P[n] for n in names], dout.float())

    xt = x.clone().requires_grad_()
    out = ops.vit_block(xt, *[P[n] for n in names], H)
    _close(out, ref, atol=6e-2)
    got = torch.autograd.grad(out, [xt] + [P[n] for n in names], dout)
    _close(got[0], rg[0], atol=8e-2, rtol=3e-2)
    for n, a_, r in zip(names, got[1:], rg[1:]):
        cs = F.cosine_similarity(a_.flatten().float(), r.flatten(), dim=0).item()
        assert cs > 0.995, (n, cs)
        assert abs(a_.float().norm().item() / r.norm().item() - 1) < 0.03, n
    with ops.input_grad_only():
        out2 = ops.vit_block(xt, *[P[n] for n in names], H)
    (dx2,) = torch.autograd.grad(out2, [xt], dout)
    assert torch.equal(dx2, got[0])


def test_vit_engine_matches_oracle(cuda_dev):
    """ViT-S-CvSt, same seed-0 weights: bf16 engine on the GPU vs the fp32 restatement on the CPU.
    Tolerances: logits 5e-2 absolute, input-gradient cosine >= 0.98, weight-gradient cosine >= 0.95."""
    from revisiting_at_b200 import vit, ops as O
    from oracle import vit_oracle as vo
    o = vo.build(normalize=True, seed=0)
    m = vit.build(normalize=True, seed=1)
    m.load_state_dict(o.state_dict())
    m = m.to(cuda_dev).eval()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(3, 3, 224, 224, generator=g)
    y = torch.randint(0, 1000, (3,), generator=g)
    xo = x.clone().requires_grad_()
    lo = o(xo)
    (go,) = torch.autograd.grad(F.cross_entropy(lo, y, reduction='sum'), xo)
    for ctx in (None, O.input_grad_only()):
        xm = x.to(cuda_dev).requires_grad_()
        if ctx is None:
            lm = m(xm)
        else:
            with ctx:
                lm = m(xm)
        (gm,) = torch.autograd.grad(F.cross_entropy(lm.float(), y.to(cuda_dev), reduction='sum'), xm)
        assert (lm.float().cpu() - lo).abs().max() <= 5e-2, (lm.float().cpu() - lo).abs().max()
        cos = F.cosine_similarity(gm.cpu().flatten(1), go.flatten(1)).min().item()
        assert cos >= 0.98, cos
        assert all(p.grad is None for p in m.parameters())
    m.train()
    F.cross_entropy(m(x.to(cuda_dev)).float(), y.to(cuda_dev)).backward()
    o.train()
    F.cross_entropy(o(x), y).backward()
    od = dict(o.named_parameters())
    for n, p in m.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n
        ref = od[n].grad
        cs = F.cosine_similarity(p.grad.flatten().cpu().float(), ref.flatten(), dim=0).item()
        assert cs >= 0.95 or ref.abs().max() < 1e-7, (n, cs)


def test_apgd_on_vit_stays_in_ball(cuda_dev):
    """apgd_train through the ViT engine: result inside the fp32 eps-ball and [0,1], masks of the right type
    (the `check_imgs` invariants of utils_eval.py:67-81)."""
    import autopgd_train_clean as product
    from revisiting_at_b200 import vit
    m = vit.build(normalize=True, seed=0).to(cuda_dev).eval()
    g = torch.Generator().manual_seed(4)
    x = torch.rand(4, 3, 224, 224, generator=g).to(cuda_dev)
    y = torch.randint(0, 1000, (4,), generator=g).to(cuda_dev)
    eps = 4. / 255.
    xb, acc, lb, xa = product.apgd_train(m, x, y, 'Linf', eps, n_iter=2)
    e32 = torch.tensor(eps, dtype=torch.float32).item()
    assert bool(((xb - x).abs() <= e32 + 1e-7).all()) and bool((xb >= 0).all()) and bool((xb <= 1).all())
    assert acc.dtype == torch.bool and lb.shape == (4,) and bool(torch.isfinite(lb).all())
    assert (xb - x).abs().max() > 0.5 * eps
