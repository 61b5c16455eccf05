"""GPU: the hand-written tcgen05/TMEM/TMA GEMM (b200at_gemm_bf16) against torch fp32 matmul of the same
bf16 operands, every fused epilogue, shapes of the ConvNeXt pwconv layers plus ragged M / K / N tails.
Tolerance: the result is rounded once to bf16 (rel 2^-8) on top of fp32 accumulation-order noise."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16

SHAPES = [(256, 96, 96), (1000, 384, 96), (4096, 96, 384), (512, 768, 192), (300, 3072, 768), (6272, 768, 3072),
          (128, 112, 40), (777, 400, 200), (50176, 384, 96)]


@pytest.fixture(scope='module')
def abi(cuda_dev):
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import _abi
    return _abi


def _check(got, want, what):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    tol = 1e-2 * want.abs() + 2e-2
    bad = err > tol
    assert not bool(bad.any()), f'{what}: max err {err.max().item():.4g}, {int(bad.sum())} bad of {bad.numel()}, first bad at {bad.nonzero()[0].tolist()}'


@pytest.mark.parametrize('M,N,K', SHAPES)
def test_gemm_all_epilogues(abi, cuda_dev, M, N, K):
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g, device=cuda_dev).to(BF16)
    b = (torch.randn(N, K, generator=g, device=cuda_dev) * K ** -0.5).to(BF16)
    bias = torch.randn(N, generator=g, device=cuda_dev)
    aux = torch.randn(M, N, generator=g, device=cuda_dev).to(BF16)
    acc = a.float() @ b.float().t()
    c = torch.full((M, N), float('nan'), device=cuda_dev, dtype=BF16)

    abi.gemm_bf16(a, b, c, abi.EPI_NONE)
    torch.cuda.synchronize()
    _check(c, acc, 'none')
    abi.gemm_bf16(a, b, c, abi.EPI_BIAS, bias=bias)
    _check(c, acc + bias, 'bias')
    c2 = torch.empty_like(c)
    abi.gemm_bf16(a, b, c, abi.EPI_BIAS_GELU, bias=bias, c2=c2)
    _check(c2, acc + bias, 'pre-activation')
    _check(c, F.gelu(acc + bias), 'bias_gelu')
    abi.gemm_bf16(a, b, c, abi.EPI_RESIDUAL, bias=bias, aux=aux)
    _check(c, aux.float() + acc + bias, 'residual')
    abi.gemm_bf16(a, b, c, abi.EPI_GELU_GRAD, aux=aux)
    z = aux.float().requires_grad_()
    (gp,) = torch.autograd.grad(F.gelu(z).sum(), z)
    _check(c, acc * gp, 'gelu_grad')
    torch.cuda.synchronize()


def test_gemm_back_to_back_tiles_reuse_tmem(abi, cuda_dev):
    """more output tiles than SMs: every CTA walks several tiles through both TMEM accumulator stages"""
    M, N, K = 148 * 128 * 3 + 64, 192, 768
    g = torch.Generator(device='cuda').manual_seed(5)
    a = torch.randn(M, K, generator=g, device=cuda_dev).to(BF16)
    b = (torch.randn(N, K, generator=g, device=cuda_dev) * K ** -0.5).to(BF16)
    c = torch.empty(M, N, device=cuda_dev, dtype=BF16)
    abi.gemm_bf16(a, b, c, abi.EPI_NONE)
    _check(c, a.float() @ b.float().t(), 'multi-tile')


def test_gemm_gelu_epilogues_over_many_tiles(abi, cuda_dev):
    """the GELU / GELU' epilogues (16 epilogue warps, 32-column passes, aux prefetched one pass ahead) on more output tiles
    than SMs, a ragged last row tile, and without the optional pre-activation output"""
    M, N, K = 148 * 128 * 2 + 77, 512, 192
    g = torch.Generator(device='cuda').manual_seed(11)
    a = torch.randn(M, K, generator=g, device=cuda_dev).to(BF16)
    b = (torch.randn(N, K, generator=g, device=cuda_dev) * K ** -0.5).to(BF16)
    bias = torch.randn(N, generator=g, device=cuda_dev)
    aux = torch.randn(M, N, generator=g, device=cuda_dev).to(BF16)
    acc = a.float() @ b.float().t()
    c = torch.full((M, N), float('nan'), device=cuda_dev, dtype=BF16)
    abi.gemm_bf16(a, b, c, abi.EPI_BIAS_GELU, bias=bias)
    _check(c, F.gelu(acc + bias), 'bias_gelu without c2')
    c2 = torch.full((M, N), float('nan'), device=cuda_dev, dtype=BF16)
    abi.gemm_bf16(a, b, c, abi.EPI_BIAS_GELU, bias=bias, c2=c2)
    _check(c2, acc + bias, 'pre-activation')
    _check(c, F.gelu(acc + bias), 'bias_gelu')
    c.fill_(float('nan'))
    abi.gemm_bf16(a, b, c, abi.EPI_GELU_GRAD, aux=aux)
    z = aux.float().requires_grad_()
    (gp,) = torch.autograd.grad(F.gelu(z).sum(), z)
    _check(c, acc * gp, 'gelu_grad')


@pytest.mark.parametrize('M,N,K', [(148 * 128 + 77, 512, 192), (300, 3072, 768), (777, 400, 200)])
def test_gemm_gelu_grad_with_column_sums(abi, cuda_dev, M, N, K):
    """the GELU' GEMM that also accumulates the column sums of its (bf16-rounded) output: same output as the plain
    epilogue, sums equal to a separate pass over it"""
    g = torch.Generator(device='cuda').manual_seed(M + N)
    a = torch.randn(M, K, generator=g, device=cuda_dev).to(BF16)
    b = (torch.randn(N, K, generator=g, device=cuda_dev) * K ** -0.5).to(BF16)
    aux = torch.randn(M, N, generator=g, device=cuda_dev).to(BF16)
    want = torch.empty(M, N, device=cuda_dev, dtype=BF16)
    abi.gemm_bf16(a, b, want, abi.EPI_GELU_GRAD, aux=aux)
    got = torch.full((M, N), float('nan'), device=cuda_dev, dtype=BF16)
    col = torch.full((N,), 0.5, device=cuda_dev)                     # accumulated INTO
    abi.gemm_gelu_grad_colsum(a, b, got, aux, col)
    assert torch.equal(got, want)
    ref = want.float().sum(0) + 0.5
    assert torch.allclose(col, ref, atol=1e-3 * M ** 0.5, rtol=1e-4), float((col - ref).abs().max())
