"""Layer functions of the ConvNeXt-CvSt / ViT-S-CvSt forward and backward (reference math: models/convnext.py:37-50,
utils_architecture.py:57-81) on NHWC bf16 activations.

Each layer group is one autograd Function over the C ABI of include/b200at_model.h:
  * depthwise 7x7 conv (fwd / input-grad with the residual join / weight-grad), per-pixel LayerNorm (fwd [+GELU] /
    input-grad / gamma-beta grads), bias+GELU: the hand-written kernels of csrc/b200at_convnext.cu;
  * the block's MLP (pwconv1 -> GELU -> pwconv2 + layer scale + bias + residual) and its input gradient: ONE tcgen05
    kernel per direction with the hidden activation kept on chip (csrc/b200at_mlp.cu) for C in {96, 128, 192}, the
    tcgen05 GEMM (csrc/b200at_gemm.cu) + bias/GELU kernels for the wider stages;
  * first stem stage fused (csrc/b200at_stem.cu), downsample = patch-layout LayerNorm + tcgen05 GEMM, attention
    (csrc/b200at_attention.cu).
The Functions compute weight gradients only when autograd will ask for them (the attack's backward is input-grad
only: autopgd_train_clean.py:185).  Library calls that remain: cuBLAS for the weight-gradient GEMMs (contraction over
the M rows), cuDNN for the strided 3x3 stem convolutions outside the fused first stage.
"""
import os
import weakref

import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import _abi

BF16 = torch.bfloat16

# Custom autograd Functions cannot see which gradients a particular torch.autograd.grad call wants
# (`needs_input_grad` only says which inputs require grad), so the attack -- whose backward is input-grad
# only (autopgd_train_clean.py:185,283) -- declares it here and the Functions skip every weight gradient.
_INPUT_GRAD_ONLY = [False]


class input_grad_only:
    """Context manager: forwards run under it promise that only dL/dx will be asked of their graph."""
    def __enter__(self):
        self.prev = _INPUT_GRAD_ONLY[0]
        _INPUT_GRAD_ONLY[0] = True

    def __exit__(self, *exc):
        _INPUT_GRAD_ONLY[0] = self.prev
        return False


def _wants(ctx, *idx):
    return ctx.param_grads and any(ctx.needs_input_grad[i] for i in idx)


def _need_cuda(x):
    if not x.is_cuda:
        raise _abi.B200atError('the ConvNeXt-CvSt engine runs on CUDA only (no CPU path); got a CPU tensor')


class _LayerNorm(Function):
    """y = LN(x [+ pre_bias]) [-> GELU].  `pre_bias` is the bias of a library convolution in front of the LayerNorm (the
    CvSt stems, utils_architecture.py:205-211): added inside the LayerNorm kernels instead of by a separate full pass, its
    gradient is the column sum of dx."""
    @staticmethod
    def forward(ctx, x, w, b, eps, gelu, pre_bias=None):
        x = x.contiguous()
        y = torch.empty_like(x)
        M = x.numel() // x.shape[-1]
        mean = torch.empty(M, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        wf, bf = w.detach().float().contiguous(), b.detach().float().contiguous()
        pbf = None if pre_bias is None else pre_bias.detach().float().contiguous()
        _abi.ln_fwd(x, wf, bf, y, mean, rstd, eps, gelu, pre_bias=pbf)
        ctx.save_for_backward(x, wf, bf, mean, rstd, pbf)
        ctx.gelu = gelu
        ctx.param_grads = not _INPUT_GRAD_ONLY[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wf, bf, mean, rstd, pbf = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        pg = _wants(ctx, 1, 2)
        dw = torch.zeros_like(wf) if pg else None
        db = torch.zeros_like(bf) if pg else None
        _abi.ln_bwd(dy, x, wf, bf, mean, rstd, dx, dw, db, ctx.gelu, pre_bias=pbf)
        dpb = None
        if pbf is not None and ctx.param_grads and ctx.needs_input_grad[5]:
            dpb = torch.zeros_like(pbf)
            _abi.colsum_bf16(dx.view(-1, dx.shape[-1]), dpb)
        return dx, dw, db, None, None, dpb


def layer_norm(x, w, b, eps=1e-6, gelu=False, pre_bias=None):
    """x: [..., C] bf16 NHWC."""
    return _LayerNorm.apply(x, w, b, eps, gelu, pre_bias)


class _DwConv7(Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous()
        C = x.shape[-1]
        wt = _taps(w)
        y = torch.empty_like(x)
        _abi.dwconv7_fwd(x, wt, b.detach().float().contiguous(), y)
        ctx.save_for_backward(x, _taps_flipped(w))
        ctx.param_grads = not _INPUT_GRAD_ONLY[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wtf = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _abi.dwconv7_fwd(dy, wtf, None, dx)                            # correlation with the flipped taps
        if _wants(ctx, 1, 2):
            C = x.shape[-1]
            dwt = torch.zeros(49, C, device=x.device, dtype=torch.float32)
            db = torch.zeros(C, device=x.device, dtype=torch.float32)
            _abi.dwconv7_wgrad(x, dy, dwt, db)
            dw = dwt.t().reshape(C, 1, 7, 7)
        return dx, dw, db


class _BiasGelu(Function):
    @staticmethod
    def forward(ctx, z, bias):
        z = z.contiguous()
        bf = bias.detach().float().contiguous()
        h = torch.empty_like(z)
        _abi.bias_gelu_fwd(z, bf, h)
        ctx.save_for_backward(z, bf)
        ctx.param_grads = not _INPUT_GRAD_ONLY[0]
        return h

    @staticmethod
    def backward(ctx, dh):
        z, bf = ctx.saved_tensors
        dz = torch.empty_like(z)
        db = torch.zeros_like(bf) if _wants(ctx, 1) else None
        _abi.bias_gelu_bwd(dh.contiguous(), z, bf, dz, db)
        return dz, db


class _ScaleResidual(Function):
    """out = res + gamma * (z + bias)"""
    @staticmethod
    def forward(ctx, z, bias, gamma, res):
        z, res = z.contiguous(), res.contiguous()
        bf, gf = bias.detach().float().contiguous(), gamma.detach().float().contiguous()
        out = torch.empty_like(z)
        _abi.scale_residual_fwd(z, bf, gf, res, out)
        ctx.param_grads = (not _INPUT_GRAD_ONLY[0]) and (bias.requires_grad or gamma.requires_grad)
        ctx.save_for_backward(z if ctx.param_grads else None, bf, gf)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, bf, gf = ctx.saved_tensors
        dout = dout.contiguous()
        dz = torch.empty_like(dout)
        _abi.scale_bwd(dout, gf, dz)
        dbias = dgamma = None
        if _wants(ctx, 1, 2):
            d32 = dout.float()
            col = d32.sum(0)
            dbias = col * gf
            dgamma = (d32 * z.float()).sum(0) + col * bf
        return dz, dbias, dgamma, dout


def _f32(t):
    return t.detach().float().contiguous()


# which pwconv GEMMs run on the hand-written tcgen05 kernel (the rest: cuBLAS + the elementwise kernels).
#   'residual'  pwconv2 forward with layer-scale + bias + residual fused in the epilogue
#   'dgrad1'    d(t2) = dz W1                     (no epilogue; same speed as cuBLAS, one library call less)
#   'fc1'       pwconv1 forward, plain (bias rides in the GELU kernel): 3-9 % faster than cuBLAS at these shapes
#   'dgrad2'    da = dout (gamma W2)              (plain; same shapes as 'fc1')
#   'gelu'      pwconv1 forward with bias + GELU fused (+ pre-activation saved): 50 vs 32 + 33 us at 25088 x 1536 x 384
#   'gelu_grad' dz = (dout W2g) * GELU'(z) fused: 55 vs 33 + 44 us            (profiles/r02_ops_bench_gemm_epilogues.txt)
#   'mlp'       the whole MLP (pwconv1 -> GELU -> pwconv2 + scale + bias + residual, and its input gradient) as ONE
#               kernel per direction with the 4C hidden kept on chip (csrc/b200at_mlp.cu), for C in {96, 128, 192}
TCGEN05 = set(filter(None, os.environ.get('B200AT_TCGEN05', 'residual,dgrad1,fc1,dgrad2,mlp,gelu,gelu_grad').split(',')))
_MLP_OK = {}
GELU_GRAD_COLSUM = os.environ.get('B200AT_GELU_GRAD_COLSUM', '1') == '1'   # pwconv1 bias gradient in the GELU' GEMM's epilogue


def _mlp_fused(C):
    if 'mlp' not in TCGEN05:
        return False
    if C not in _MLP_OK:
        _MLP_OK[C] = _abi.mlp_fused_supported(C)
    return _MLP_OK[C]

_PCACHE = {}     # (ids of the source parameters, tag) -> (stamp, value, weakrefs of the parameters)


def _stamp(params):
    return tuple((p._version, p.data_ptr()) for p in params)


def _derived_multi(params, tag, fn):
    """Kernel-side copy derived from one or several parameters (cast / transposed / flipped / folded), rebuilt only
    when one of them changes; shared by the 4 forwards + 3 backwards of a step."""
    key = (tuple(id(p) for p in params), tag)
    hit = _PCACHE.get(key)
    if hit is not None and hit[0] == _stamp(params) and all(r() is p for r, p in zip(hit[2], params)):
        return hit[1]                                        # same live tensor objects, same storage, not written since
    with torch.no_grad():
        val = fn(*[p.detach() for p in params])
    if len(_PCACHE) > 8192:                                  # models that came and went (tests)
        for k in [k for k, v in _PCACHE.items() if any(r() is None for r in v[2])]:
            del _PCACHE[k]
    _PCACHE[key] = (_stamp(params), val, tuple(weakref.ref(p) for p in params))
    return val


def _derived(param, tag, fn):
    return _derived_multi((param,), tag, fn)


def snapshot_derived():
    """The current cache entries (used right after a CUDA-graph capture: they are the graph's own buffers)."""
    return dict(_PCACHE)


def install_derived(snapshot):
    """Declare the snapshot's values current for the parameters as they are now.  Only valid when something has
    just recomputed them in place from the current parameters -- i.e. right after replaying the CUDA graph whose
    capture produced them; the training forward that follows then reuses the graph's weight copies instead of
    deriving its own."""
    for key, (_, val, refs) in snapshot.items():
        params = [r() for r in refs]
        if all(p is not None for p in params):
            _PCACHE[key] = (_stamp(params), val, refs)


def invalidate_derived(keep_tags=('host3',)):
    """Forget every cached kernel-side parameter copy (except host-side constants): the next use rebuilds
    it.  REQUIRED after any parameter write that does not move the tensor's version counter -- `p.data.copy_()` /
    `p.data.mul_()` (swapping EMA weights in for an evaluation, manual weight surgery), fused optimisers outside
    `torch.optim`'s step hook -- or the model silently runs on the previous bf16 / transposed copies
    (`load_state_dict`, `copy_`, `add_` on the parameter itself DO move it and need nothing).  Called right before a CUDA-graph capture so that the cast / transpose / fold kernels are recorded
    INSIDE the graph and every replay re-derives them from the (in-place updated) fp32 master parameters."""
    for k in [k for k in _PCACHE if k[1] not in keep_tags]:
        del _PCACHE[k]


# Fused optimisers (torch._fused_adamw_ & co.) update the parameters in place WITHOUT moving their version counters, so
# the version check above cannot see an optimiser step: every optimiser's step() therefore drops the cache through
# torch's global post-step hook.  (copy_ / add_ / load_state_dict / foreach optimisers do bump the counters.)
def _after_optimizer_step(*_args, **_kw):
    invalidate_derived()


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _register_post_hook
    _OPT_HOOK = _register_post_hook(_after_optimizer_step)
except ImportError:                                          # very old torch: callers must invalidate themselves
    _OPT_HOOK = None


def _taps(dw_w):
    """depthwise weight [C,1,7,7] -> tap-major fp32 [49][C]"""
    return _derived(dw_w, 'taps', lambda w: w.float().reshape(w.shape[0], 49).t().contiguous())


def _taps_flipped(dw_w):
    """taps of the input-gradient correlation (180-degree rotated kernel)"""
    return _derived(dw_w, 'taps_flip', lambda w: w.float().reshape(w.shape[0], 49).t().flip(0).contiguous())


def _bf16(p):
    return _derived(p, 'bf16', lambda w: w.to(BF16).contiguous())


def _prepared(w1, w2, b2, gamma):
    """bf16 / transposed / layer-scale-folded copies of the block's MLP weights."""
    def build(w1, w2, b2, gamma):
        if w1.is_cuda and gamma.numel() % 32 == 0:
            return _abi.prepare_mlp_weights(w1, w2, b2, gamma)      # one launch instead of seven
        gf = gamma.float()
        w1b = w1.to(BF16).contiguous()                                  # [4C, C]   pwconv1: t2 @ w1b^T
        w1t = w1b.t().contiguous()                                      # [C, 4C]   dt2 = dz @ w1b  == dz @ w1t^T
        w2g = (gf[:, None] * w2.float()).to(BF16).contiguous()          # [C, 4C]   gamma folded: a @ w2g^T
        w2gt = w2g.t().contiguous()                                     # [4C, C]   da = dout @ w2g == dout @ w2gt^T
        b2g = (gf * b2.float()).contiguous()
        return dict(w1b=w1b, w1t=w1t, w2g=w2g, w2gt=w2gt, b2g=b2g, gf=gf.contiguous())
    return _derived_multi((w1, w2, b2, gamma), 'convnext_mlp', build)


def _wgrad(dy2, x2):
    """weight gradient dy2^T x2 (contraction over the M rows): library GEMM with fp32 output -- the split-K reduction
    writes fp32 directly instead of bf16 followed by a cast kernel."""
    return torch.mm(dy2.t(), x2, out_dtype=torch.float32)


def _zeros_split(dev, *sizes):
    """fp32 accumulators for the kernels that atomically add into them: ONE fill, views of the given sizes"""
    buf = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
    out, o = [], 0
    for n in sizes:
        out.append(buf[o:o + n])
        o += n
    return out


def _gemm(a, w, epi=_abi.EPI_NONE, bias=None, aux=None, c2=None):
    c = torch.empty(a.shape[0], w.shape[0], device=a.device, dtype=BF16)
    _abi.gemm_bf16(a, w, c, epi, bias=bias, aux=aux, c2=c2)
    return c


class _ConvNeXtBlock(Function):
    """One whole block (models/convnext.py:37-50) as a single autograd node on NHWC bf16:

        t1 = dwconv7(x) ; t2 = LN(t1) ; z = t2 W1^T + b1 ; a = GELU(z) ; out = x + a (gamma W2)^T + gamma b2

    Owning the whole backward lets the input-gradient pass (all the attack ever asks for) keep only
    {t1, LN stats, z}, fold the residual-gradient join into the depthwise input-gradient kernel, and skip
    every weight gradient; the outer training step additionally saves {x, t2, a} for the weight grads.
    Layer scale is folded into the second weight matrix (gamma W2, gamma b2), so `out` is one GEMM with a
    residual epilogue and the backward needs no separate scale pass.
    """

    @staticmethod
    def forward(ctx, x, dw_w, dw_b, ln_w, ln_b, w1, b1, w2, b2, gamma):
        x = x.contiguous()
        B, H, W, C = x.shape
        M = B * H * W
        pg = not _INPUT_GRAD_ONLY[0]
        P = _prepared(w1, w2, b2, gamma)
        wt, wtf = _taps(dw_w), _taps_flipped(dw_w)                      # tap-major [49][C]
        lnw, lnb, b1f = _f32(ln_w), _f32(ln_b), _f32(b1)
        t1 = torch.empty_like(x)
        _abi.dwconv7_fwd(x, wt, _f32(dw_b), t1)
        t2 = torch.empty_like(x)
        mean = torch.empty(M, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        _abi.ln_fwd(t1, lnw, lnb, t2, mean, rstd, 1e-6, False)
        fused = _mlp_fused(C)
        if fused:
            z = torch.empty(M, 4 * C, device=x.device, dtype=BF16)      # pre-activation WITHOUT the bias
            a = torch.empty_like(z) if pg else None                     # only the weight gradients need it
            out = torch.empty_like(x)
            _abi.mlp_fused(t2.view(M, C), P['w1b'], P['w2g'], b1f, z, out.view(M, C), bias2=P['b2g'],
                           residual=x.view(M, C), p_out=a)
            zb = b1f
        elif 'gelu' in TCGEN05:
            z = torch.empty(M, 4 * C, device=x.device, dtype=BF16)      # pre-activation INCLUDING the bias
            a = _gemm(t2.view(M, C), P['w1b'], _abi.EPI_BIAS_GELU, bias=b1f, c2=z)
            zb = None
        else:
            if 'fc1' in TCGEN05:
                z = _gemm(t2.view(M, C), P['w1b'])                      # bias added inside the GELU kernels
            else:
                z = t2.view(M, C) @ P['w1b'].t()                        # cuBLAS
            a = torch.empty_like(z)
            _abi.bias_gelu_fwd(z, b1f, a)
            zb = b1f
        if fused:
            pass
        elif 'residual' in TCGEN05:
            out = _gemm(a, P['w2g'], _abi.EPI_RESIDUAL, bias=P['b2g'], aux=x.view(M, C)).view(B, H, W, C)
        else:
            z2 = a @ P['w2g'].t()
            out = torch.empty_like(x)
            _abi.scale_residual_fwd(z2, P['b2g'], torch.ones_like(P['gf']), x.view(M, C), out.view(M, C))
        ctx.param_grads = pg
        ctx.zb = zb
        ctx.prep = P
        keep = (t1, mean, rstd, z, wtf, lnw, lnb)
        ctx.save_for_backward(*(keep + ((x, t2, a, w2.detach(), _f32(b2)) if pg else ())))
        return out

    @staticmethod
    def backward(ctx, dout):
        sv = ctx.saved_tensors
        t1, mean, rstd, z, wtf, lnw, lnb = sv[:7]
        P = ctx.prep
        dout = dout.contiguous()
        B, H, W, C = dout.shape
        M = B * H * W
        pg = ctx.param_grads and any(ctx.needs_input_grad[1:])
        d2 = dout.view(M, C)
        if pg:
            db1, ddw, ddb, col, dlnw, dlnb = _zeros_split(dout.device, 4 * C, 49 * C, C, C, C, C)
            ddw = ddw.view(49, C)
        else:
            db1 = dlnw = dlnb = None
        dt2 = None
        if _mlp_fused(C) and ctx.zb is not None:
            dz = torch.empty_like(z) if pg else None                    # only the weight gradients need it
            dt2 = torch.empty_like(dout)
            _abi.mlp_fused(d2, P['w2gt'], P['w1t'], ctx.zb, z, dt2.view(M, C), p_out=dz, backward=True)
            if pg:
                _abi.colsum_bf16(dz, db1)
        elif 'gelu_grad' in TCGEN05 and ctx.zb is None:
            if pg and GELU_GRAD_COLSUM:                                  # bias gradient in the GEMM's epilogue
                dz = torch.empty(M, 4 * C, device=dout.device, dtype=BF16)
                _abi.gemm_gelu_grad_colsum(d2, P['w2gt'], dz, z, db1)
            else:
                dz = _gemm(d2, P['w2gt'], _abi.EPI_GELU_GRAD, aux=z)
                if pg:
                    _abi.colsum_bf16(dz, db1)
        else:
            da = _gemm(d2, P['w2gt']) if 'dgrad2' in TCGEN05 else d2 @ P['w2g']   # [M,4C]  (layer scale folded)
            dz = torch.empty_like(da)
            zero = ctx.zb if ctx.zb is not None else torch.zeros(4 * C, device=dout.device, dtype=torch.float32)
            _abi.bias_gelu_bwd(da, z, zero, dz, db1)                    # pwconv1 bias gradient rides along
        if dt2 is not None:
            pass
        elif 'dgrad1' in TCGEN05:
            dt2 = _gemm(dz, P['w1t']).view(B, H, W, C)
        else:
            dt2 = (dz @ P['w1b']).view(B, H, W, C)
        dt1 = torch.empty_like(dt2)
        _abi.ln_bwd(dt2, t1, lnw, lnb, mean, rstd, dt1, dlnw, dlnb, False)
        dx = torch.empty_like(dout)
        _abi.dwconv7_fwd(dt1, wtf, None, dx, add=dout)                  # + residual gradient
        if not pg:
            return (dx,) + (None,) * 9
        x, t2, a, w2, b2f = sv[7:]
        gf = P['gf']
        _abi.dwconv7_wgrad(x, dt1, ddw, ddb)
        dw1 = _wgrad(dz, t2.view(M, C))
        dw2g = _wgrad(d2, a)                                            # gradient w.r.t. gamma-folded W2
        _abi.colsum_bf16(d2, col)
        dw2, db2, dgamma = _abi.finish_mlp_grads(dw2g, w2.float().contiguous(), col, b2f, gf)
        return dx, ddw.t().reshape(C, 1, 7, 7), ddb, dlnw, dlnb, dw1, db1, dw2, db2, dgamma


def convnext_block(x, dw_w, dw_b, ln_w, ln_b, w1, b1, w2, b2, gamma):
    """x: [B,H,W,C] bf16 NHWC -> same.  models/convnext.py:37-50."""
    return _ConvNeXtBlock.apply(x, dw_w, dw_b, ln_w, ln_b, w1, b1, w2, b2, gamma)


def _cast(p):
    """bf16 copy of a parameter for a library call.  When the parameter needs a gradient (outer training step)
    the cast stays in the autograd graph; for the attack's input-grad-only evaluations it is cached."""
    if _INPUT_GRAD_ONLY[0] or not torch.is_grad_enabled() or not p.requires_grad:
        return _bf16(p)
    return p.to(BF16)


class _Stem0(Function):
    """First stem stage in one kernel per direction (csrc/b200at_stem.cu): normalise -> conv3x3 s2 -> LN -> GELU.
    Input-gradient only (the attack's evaluations); the forward saves nothing but its input."""

    @staticmethod
    def forward(ctx, x, wk, cb, lw, lb, mean3, std3):
        B, _, H, W = x.shape
        y = torch.empty(B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, cb.numel(), device=x.device, dtype=BF16)
        _abi.stem0_fwd(x, mean3, std3, wk, cb, lw, lb, y)
        ctx.save_for_backward(x, wk, cb, lw, lb)
        ctx.norm = (mean3, std3)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wk, cb, lw, lb = ctx.saved_tensors
        dx = torch.empty_like(x)
        _abi.stem0_bwd_input(dy.contiguous(), x, ctx.norm[0], ctx.norm[1], wk, cb, lw, lb, dx)
        return dx, None, None, None, None, None, None


class _Stem0Train(Function):
    """The same first stem stage in the TRAINING forward (x: the attack's detached fp32 output): the fused kernel also
    leaves the convolution output (without bias) and the LayerNorm statistics, so the backward is the LN+GELU backward
    kernel with the conv bias as `pre_bias`, its column sum for the conv bias, and the library weight gradient of the
    convolution on the re-normalised input.  No input gradient (x does not require one)."""

    @staticmethod
    def forward(ctx, x, cw, cb, lw, lb, mean3, std3):
        B, _, H, W = x.shape
        C0 = cw.shape[0]
        wk = _derived(cw, 'stem0_wk', lambda w: w.float().reshape(w.shape[0], 27).t().contiguous())   # [27][C0]
        cbf, lwf, lbf = _f32(cb), _f32(lw), _f32(lb)
        shp = (B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C0)
        y = torch.empty(shp, device=x.device, dtype=BF16)
        y_pre = torch.empty(shp, device=x.device, dtype=BF16)
        mean = torch.empty(shp[0] * shp[1] * shp[2], device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        _abi.stem0_fwd_save(x, mean3, std3, wk, cbf, lwf, lbf, y, y_pre, mean, rstd)
        ctx.save_for_backward(x, y_pre, mean, rstd, cw, cbf, lwf, lbf)
        ctx.norm = (mean3, std3)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y_pre, mean, rstd, cw, cbf, lwf, lbf = ctx.saved_tensors
        C0 = cw.shape[0]
        dpre = torch.empty_like(y_pre)
        dlw, dlb, dcb = _zeros_split(dy.device, C0, C0, C0)
        _abi.ln_bwd(dy.contiguous(), y_pre, lwf, lbf, mean, rstd, dpre, dlw, dlb, True, pre_bias=cbf)
        _abi.colsum_bf16(dpre.view(-1, C0), dcb)
        B, _, H, W = x.shape
        t = torch.empty(B, H, W, 3, device=x.device, dtype=BF16)
        _abi.normalize_nhwc_bf16(x, ctx.norm[0], ctx.norm[1], t)
        _, dcw, _ = torch.ops.aten.convolution_backward(dpre.permute(0, 3, 1, 2), t.permute(0, 3, 1, 2),
                                                        _bf16(cw).contiguous(memory_format=torch.channels_last), None,
                                                        (2, 2), (1, 1), (1, 1), False, (0, 0), 1, (False, True, False))
        return None, dcw.to(cw.dtype), dcb, dlw, dlb, None, None


def _host3(t):
    """3 python floats of the normaliser's mean / std buffer (one D2H copy per buffer object and version)."""
    if t is None:
        return None
    return _derived(t, 'host3', lambda v: tuple(float(a) for a in v.flatten().cpu()))


STEM0_KERNEL = os.environ.get('B200AT_STEM0', '1') == '1'
STEM_CONV_GEMM = os.environ.get('B200AT_STEM_CONV', 'gemm') == 'gemm'
STEM0_TRAIN = os.environ.get('B200AT_STEM0_TRAIN', '1') == '1'     # fused first stage in the training forward too


def _conv3x3s2_wk(w):
    """[Co, Ci, 3, 3] -> [Co, 9 * 64] bf16: tap-major, the channels of a tap padded to one 64-wide k-block"""
    Co, Ci = w.shape[0], w.shape[1]
    wk = torch.zeros(Co, 9, 64, device=w.device, dtype=BF16)
    wk[:, :, :Ci] = w.permute(0, 2, 3, 1).reshape(Co, 9, Ci).to(BF16)
    return wk.view(Co, 576)


class _Conv3x3S2(Function):
    """Conv2d(k=3, s=2, p=1, no bias) on NHWC bf16 (the later convolutions of the CvSt stems,
    utils_architecture.py:205-211): forward = implicit GEMM on the tcgen05 kernel (b200at_conv3x3s2_fwd); the input and
    weight gradients stay library calls."""

    @staticmethod
    def forward(ctx, x, cw):
        x = x.contiguous()
        B, H, W, _ = x.shape
        wk = _derived(cw, 'conv3x3s2_wk', _conv3x3s2_wk)
        y = torch.empty(B, H // 2, W // 2, cw.shape[0], device=x.device, dtype=BF16)
        if not _abi.conv3x3s2_fwd(x, wk, y):             # a shape the kernel does not take (shared-memory budget)
            y = F.conv2d(x.permute(0, 3, 1, 2), _bf16(cw), None, stride=2, padding=1).permute(0, 2, 3, 1).contiguous()
        ctx.param_grads = not _INPUT_GRAD_ONLY[0]
        ctx.save_for_backward(x, cw)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, cw = ctx.saved_tensors
        xc = x.permute(0, 3, 1, 2)                       # NCHW views of channels_last storage
        dyc = dy.contiguous().permute(0, 3, 1, 2)
        wb = _bf16(cw).contiguous(memory_format=torch.channels_last)
        want_w = ctx.param_grads and ctx.needs_input_grad[1]
        dx, dw, _ = torch.ops.aten.convolution_backward(dyc, xc, wb, None, (2, 2), (1, 1), (1, 1), False, (0, 0), 1,
                                                        (ctx.needs_input_grad[0], want_w, False))
        if dx is not None:
            dx = dx.permute(0, 2, 3, 1)                  # channels_last result viewed as NHWC
        if dw is not None:
            dw = dw.to(cw.dtype)
        return dx, dw


def _conv3x3s2_ok(x, cw):
    B, H, W, Ci = x.shape
    return (STEM_CONV_GEMM and x.dtype == BF16 and H % 2 == 0 and W % 2 == 0 and Ci % 8 == 0 and Ci <= 64
            and cw.shape[0] % 16 == 0 and cw.shape[0] <= 256 and W // 2 <= 128 and tuple(cw.shape[2:]) == (3, 3))


def stem_layer(x, cw, cb, lw, lb, stride, first, mean=None, std=None):
    """conv3x3 -> LN over C + GELU.  First layer: x is fp32 NCHW in [0,1] (normalised here when mean/std are
    given); later layers: x is NHWC bf16.  Returns NHWC bf16.  The first layer of an attack evaluation (input
    gradient only / no gradient) is the fused direct kernel; everything else is the library convolution
    followed by the fused LN+GELU kernel."""
    _need_cuda(x)
    if (first and STEM0_KERNEL and stride == 2 and cw.shape[0] in (48, 64, 96) and x.dtype == torch.float32
            and (_INPUT_GRAD_ONLY[0] or not torch.is_grad_enabled())):
        wk = _derived(cw, 'stem0_wk', lambda w: w.float().reshape(w.shape[0], 27).t().contiguous())   # [27][C0]
        return _Stem0.apply(x.contiguous(), wk, _f32(cb), _f32(lw), _f32(lb), _host3(mean), _host3(std))
    if (first and STEM0_TRAIN and STEM0_KERNEL and stride == 2 and cw.shape[0] in (48, 64, 96) and x.dtype == torch.float32
            and x.is_contiguous() and x.shape[1] == 3 and not x.requires_grad and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0):
        # training forward on the attack's (detached) output: the fused kernel, keeping what the backward needs
        return _Stem0Train.apply(x, cw, cb, lw, lb, _host3(mean), _host3(std))
    if first and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 3 and not x.requires_grad:
        # training forward on the attack's (detached) output: normalise + cast + NHWC in one pass
        t = torch.empty(x.shape[0], x.shape[2], x.shape[3], 3, device=x.device, dtype=BF16)
        _abi.normalize_nhwc_bf16(x, _host3(mean), _host3(std), t)
        x = t.permute(0, 3, 1, 2)                        # NHWC storage viewed as NCHW channels_last
    elif first:
        if mean is not None:
            x = (x - mean) / std
        x = x.to(BF16).contiguous(memory_format=torch.channels_last)
    else:
        if stride == 2 and cw.shape[0] % 8 == 0 and _conv3x3s2_ok(x, cw):
            y = _Conv3x3S2.apply(x, cw)
            return layer_norm(y, lw, lb, 1e-6, gelu=True, pre_bias=cb)
        x = x.permute(0, 3, 1, 2)                        # NHWC storage viewed as NCHW channels_last
    # the library convolution runs WITHOUT its bias: torch adds a conv bias as a separate full pass over the output (the
    # largest activations of the network: 0.5 ms per step) and reduces its gradient in another; both ride in the
    # LayerNorm kernels instead (`pre_bias`), which need C % 8 == 0
    fold = cw.shape[0] % 8 == 0
    y = F.conv2d(x, _cast(cw), None if fold else _cast(cb), stride=stride, padding=1)
    y = y.permute(0, 2, 3, 1)                            # -> NHWC view of the channels_last result
    return layer_norm(y, lw, lb, 1e-6, gelu=True, pre_bias=cb if fold else None)


class _Downsample(Function):
    """LayerNorm -> Conv2d(kernel 2, stride 2) (models/convnext.py:79-82) as one autograd node: the LN kernel writes
    its result in the 2x2-patch layout, so the convolution is the tcgen05 GEMM [B*H/2*W/2, 4C] x [Cout, 4C]^T with
    the bias in its epilogue, and the input gradient is the GEMM with the transposed weight followed by the LN
    backward reading the patch layout.  No cuDNN, no separate bias-add / layout kernels."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, cw, cb):
        x = x.contiguous()
        B, H, W, C = x.shape
        Co = cw.shape[0]
        M = B * H * W
        lnw, lnb = _f32(ln_w), _f32(ln_b)
        wk = _derived(cw, 'patch2', lambda w: w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(BF16).contiguous())
        t = torch.empty(M // 4, 4 * C, device=x.device, dtype=BF16)
        mean = torch.empty(M, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        _abi.ln_fwd_patch2(x, lnw, lnb, t, mean, rstd, 1e-6)
        y = _gemm(t, wk, _abi.EPI_BIAS, bias=_f32(cb))
        ctx.param_grads = not _INPUT_GRAD_ONLY[0]
        ctx.save_for_backward(x, mean, rstd, lnw, lnb, cw, t if ctx.param_grads else None)
        return y.view(B, H // 2, W // 2, Co)

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, lnw, lnb, cw, t = ctx.saved_tensors
        B, H, W, C = x.shape
        Co = cw.shape[0]
        dy2 = dy.contiguous().view(-1, Co)
        wkt = _derived(cw, 'patch2_t', lambda w: w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(BF16).t().contiguous())
        dt = _gemm(dy2, wkt)                                            # [M/4, 4C], patch layout
        pg = _wants(ctx, 1, 2, 3, 4)
        dlw = torch.zeros_like(lnw) if pg else None
        dlb = torch.zeros_like(lnb) if pg else None
        dx = torch.empty_like(x)
        _abi.ln_bwd_patch2(dt, x, lnw, lnb, mean, rstd, dx, dlw, dlb)
        if not pg:
            return dx, None, None, None, None
        dwk = _wgrad(dy2, t)                                            # [Co, (kh, kw, Cin)]
        dcw = dwk.view(Co, 2, 2, C).permute(0, 3, 1, 2)
        dcb = torch.zeros(Co, device=dy.device, dtype=torch.float32)
        _abi.colsum_bf16(dy2, dcb)
        return dx, dlw, dlb, dcw, dcb


DOWNSAMPLE_GEMM = os.environ.get('B200AT_DOWNSAMPLE', 'gemm') == 'gemm'


def downsample(x, ln_w, ln_b, cw, cb):
    """LN over C -> conv2x2 s2.  NHWC bf16 in/out."""
    B, H, W, C = x.shape
    if DOWNSAMPLE_GEMM and H % 2 == 0 and W % 2 == 0 and cw.shape[0] % 16 == 0 and C % 8 == 0:
        return _Downsample.apply(x, ln_w, ln_b, cw, cb)
    y = layer_norm(x, ln_w, ln_b, 1e-6).permute(0, 3, 1, 2)            # odd sizes: library convolution
    y = F.conv2d(y, _cast(cw), _cast(cb), stride=2)
    return y.permute(0, 2, 3, 1)


def head(x, ln_w, ln_b, fw, fb):
    """global mean-pool -> LN -> Linear (tiny; library ops).  x NHWC bf16 -> logits bf16."""
    p = x.float().mean((1, 2))
    p = F.layer_norm(p, (p.shape[-1],), ln_w.float(), ln_b.float(), 1e-6)
    return F.linear(p.to(BF16), _cast(fw), _cast(fb))


# ------------------------------------------------------------------------------------------------ ViT-S-CvSt
# timm 0.8 `vision_transformer.Block` (un-vendored; call sites utils_architecture.py:271-301):
#     x = x + proj(attention(qkv(norm1(x)))) ;  x = x + fc2(GELU(fc1(norm2(x))))
# on token rows [B*N, D] bf16.  LayerNorm / bias+GELU: the kernels above; qkv (+bias), proj (+bias+residual),
# fc2 (+bias+residual) and every input-gradient GEMM: the tcgen05 kernel; attention: csrc/b200at_attention.cu.
ATTENTION_KERNEL = os.environ.get('B200AT_ATTN', 'kernel')     # 'sdpa' = library attention (A/B measurements only)


def _vit_prepared(wqkv, wproj, w1, w2):
    def build(*ws):
        prep = {}
        for name, w in zip(('qkv', 'proj', 'w1', 'w2'), ws):
            wb = w.to(BF16).contiguous()
            prep[name] = wb                                   # [out, in]: y = x @ wb^T
            prep[name + '_t'] = wb.t().contiguous()           # [in, out]: dx = dy @ wb == dy @ (wb^T)^T
        return prep
    return _derived_multi((wqkv, wproj, w1, w2), 'vit_block', build)


class _Attention(Function):
    """softmax(q k^T scale) v per head on the packed qkv rows (csrc/b200at_attention.cu)."""
    @staticmethod
    def forward(ctx, qkv, heads, scale):
        B, N, _ = qkv.shape
        qkv = qkv.contiguous()
        o = torch.empty(B, N, heads * 64, device=qkv.device, dtype=BF16)
        lse = torch.empty(B * heads * N, device=qkv.device, dtype=torch.float32)
        _abi.attn_fwd(qkv, o, lse, heads, scale)
        ctx.save_for_backward(qkv, o, lse)
        ctx.heads, ctx.scale = heads, scale
        return o

    @staticmethod
    def backward(ctx, d_o):
        qkv, o, lse = ctx.saved_tensors
        dqkv = torch.empty_like(qkv)
        _abi.attn_bwd(qkv, o, d_o.contiguous(), lse, dqkv, ctx.heads, ctx.scale)
        return dqkv, None, None


def attention(qkv, heads, scale=None):
    """qkv: [B,N,3*heads*64] bf16 -> [B,N,heads*64]."""
    _need_cuda(qkv)
    return _Attention.apply(qkv, heads, 64 ** -0.5 if scale is None else scale)


class _ViTBlock(Function):
    @staticmethod
    def forward(ctx, x, n1w, n1b, wqkv, bqkv, wproj, bproj, n2w, n2b, w1, b1, w2, b2, heads):
        x = x.contiguous()
        B, N, D = x.shape
        M = B * N
        pg = not _INPUT_GRAD_ONLY[0]
        P = _vit_prepared(wqkv, wproj, w1, w2)
        scale = (D // heads) ** -0.5
        n1wf, n1bf, n2wf, n2bf = _f32(n1w), _f32(n1b), _f32(n2w), _f32(n2b)
        x2 = x.view(M, D)
        t = torch.empty_like(x2)
        mean1 = torch.empty(M, device=x.device, dtype=torch.float32)
        rstd1 = torch.empty_like(mean1)
        _abi.ln_fwd(x2, n1wf, n1bf, t, mean1, rstd1, 1e-6, False)
        qkv = _gemm(t, P['qkv'], _abi.EPI_BIAS, bias=_f32(bqkv))
        o = torch.empty(B, N, D, device=x.device, dtype=BF16)
        lse = torch.empty(B * heads * N, device=x.device, dtype=torch.float32)
        _abi.attn_fwd(qkv.view(B, N, 3 * D), o, lse, heads, scale)
        x1 = _gemm(o.view(M, D), P['proj'], _abi.EPI_RESIDUAL, bias=_f32(bproj), aux=x2)
        t2 = torch.empty_like(x1)
        mean2 = torch.empty_like(mean1)
        rstd2 = torch.empty_like(mean1)
        _abi.ln_fwd(x1, n2wf, n2bf, t2, mean2, rstd2, 1e-6, False)
        b1f = _f32(b1)
        if 'gelu' in TCGEN05:
            z = torch.empty(M, P['w1'].shape[0], device=x.device, dtype=BF16)   # pre-activation INCLUDING the bias
            a = _gemm(t2, P['w1'], _abi.EPI_BIAS_GELU, bias=b1f, c2=z)
            b1f = torch.zeros_like(b1f)                                 # what the backward adds to the saved z
        else:
            z = _gemm(t2, P['w1'])                                      # bias added inside the GELU kernels
            a = torch.empty_like(z)
            _abi.bias_gelu_fwd(z, b1f, a)
        out = _gemm(a, P['w2'], _abi.EPI_RESIDUAL, bias=_f32(b2), aux=x1)
        ctx.param_grads, ctx.prep, ctx.heads, ctx.scale = pg, P, heads, scale
        keep = (x2, mean1, rstd1, qkv, o, lse, x1, mean2, rstd2, z, n1wf, n1bf, n2wf, n2bf, b1f)
        ctx.save_for_backward(*(keep + ((t, t2, a) if pg else ())))
        return out.view(B, N, D)

    @staticmethod
    def backward(ctx, dout):
        sv = ctx.saved_tensors
        x2, mean1, rstd1, qkv, o, lse, x1, mean2, rstd2, z, n1wf, n1bf, n2wf, n2bf, b1f = sv[:15]
        P, heads = ctx.prep, ctx.heads
        B, N, D = dout.shape
        M = B * N
        dev = dout.device
        pg = ctx.param_grads and any(ctx.needs_input_grad[1:13])
        d2 = dout.contiguous().view(M, D)
        # ---- MLP branch
        db1 = torch.zeros(4 * D, device=dev, dtype=torch.float32) if pg else None
        if 'gelu' in TCGEN05 and 'gelu_grad' in TCGEN05:                # the saved z includes the bias
            if pg and GELU_GRAD_COLSUM:
                dz = torch.empty(M, z.shape[1], device=dev, dtype=BF16)
                _abi.gemm_gelu_grad_colsum(d2, P['w2_t'], dz, z, db1)
            else:
                dz = _gemm(d2, P['w2_t'], _abi.EPI_GELU_GRAD, aux=z)    # [M,4D]
                if pg:
                    _abi.colsum_bf16(dz, db1)
        else:
            da = _gemm(d2, P['w2_t'])
            dz = torch.empty_like(da)
            _abi.bias_gelu_bwd(da, z, b1f, dz, db1)
        dt2 = _gemm(dz, P['w1_t'])                                      # [M,D]
        dn2w = torch.zeros(D, device=dev, dtype=torch.float32) if pg else None
        dn2b = torch.zeros(D, device=dev, dtype=torch.float32) if pg else None
        dx1 = torch.empty_like(dt2)
        _abi.ln_bwd(dt2, x1, n2wf, n2bf, mean2, rstd2, dx1, dn2w, dn2b, False)
        _abi.add_bf16(dx1, d2, dx1)                                     # + residual gradient
        # ---- attention branch
        do = _gemm(dx1, P['proj_t'])                                    # [M,D]
        dqkv = torch.empty_like(qkv)
        _abi.attn_bwd(qkv.view(B, N, 3 * D), o, do.view(B, N, D), lse, dqkv.view(B, N, 3 * D), heads, ctx.scale)
        dt = _gemm(dqkv, P['qkv_t'])                                    # [M,D]
        dn1w = torch.zeros(D, device=dev, dtype=torch.float32) if pg else None
        dn1b = torch.zeros(D, device=dev, dtype=torch.float32) if pg else None
        dx = torch.empty_like(dt)
        _abi.ln_bwd(dt, x2, n1wf, n1bf, mean1, rstd1, dx, dn1w, dn1b, False)
        _abi.add_bf16(dx, dx1, dx)
        dx = dx.view(B, N, D)
        if not pg:
            return (dx,) + (None,) * 13
        t, t2, a = sv[15:]

        def colsum(m):
            out = torch.zeros(m.shape[1], device=dev, dtype=torch.float32)
            _abi.colsum_bf16(m, out)
            return out
        dwqkv = _wgrad(dqkv, t)
        dwproj = _wgrad(dx1, o.view(M, D))
        dw1 = _wgrad(dz, t2)
        dw2 = _wgrad(d2, a)
        return (dx, dn1w, dn1b, dwqkv, colsum(dqkv), dwproj, colsum(dx1), dn2w, dn2b, dw1, db1, dw2, colsum(d2), None)


def vit_block(x, n1w, n1b, wqkv, bqkv, wproj, bproj, n2w, n2b, w1, b1, w2, b2, heads):
    """x: [B,N,D] bf16 tokens -> same (timm vision_transformer.Block, no layer scale, no drop path in eval)."""
    _need_cuda(x)
    return _ViTBlock.apply(x, n1w, n1b, wqkv, bqkv, wproj, bproj, n2w, n2b, w1, b1, w2, b2, heads)


class _Linear1x1(Function):
    """1x1 convolution / Linear on NHWC rows through the tcgen05 GEMM (last layer of the ViT conv stem,
    utils_architecture.py:139: `nn.Conv2d(planes*8, fin_dim, kernel_size=1)`)."""
    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous()
        M = x.numel() // x.shape[-1]
        wb = _derived(w, 'bf16_2d', lambda v: v.to(BF16).reshape(v.shape[0], -1).contiguous())
        y = _gemm(x.view(M, -1), wb, _abi.EPI_BIAS, bias=_f32(b))
        ctx.param_grads = not _INPUT_GRAD_ONLY[0]
        ctx.save_for_backward(x if ctx.param_grads else None, w)
        ctx.wshape = w.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy2 = dy.contiguous().view(-1, dy.shape[-1])
        wt = _derived(w, 'bf16_2d_t', lambda v: v.to(BF16).reshape(v.shape[0], -1).t().contiguous())
        dx = _gemm(dy2, wt).view(*dy.shape[:-1], wt.shape[0])
        if not _wants(ctx, 1, 2):
            return dx, None, None
        dw = _wgrad(dy2, x.view(dy2.shape[0], -1)).view(ctx.wshape)
        db = torch.zeros(dy2.shape[1], device=dy.device, dtype=torch.float32)
        _abi.colsum_bf16(dy2, db)
        return dx, dw, db


def linear_rows(x, w, b):
    """y = x W^T + b over the last dimension of a bf16 row tensor (W may be a [out,in,1,1] conv weight)."""
    _need_cuda(x)
    return _Linear1x1.apply(x, w, b)
