"""ctypes binding of libb200at.so (C ABI: include/b200at.h).

Passes raw device pointers, sizes and the caller's current CUDA stream; every wrapper checks
device / dtype / density so that a wrong buffer fails here and not inside a kernel.  The library
is the product: if it is missing this module raises -- there is no fallback of any kind.
"""
import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libb200at.so')
ABI_VERSION = 1

# state rows (csrc/b200at_math.cuh)
ST_STEP, ST_LOSS_BEST, ST_LOSS_BEST_LAST, ST_REDUCED_LAST, ST_ACC, ST_FLAGS, ST_LOSS_CUR = range(7)
ST_TOPK, ST_SP_OLD, ST_SP_BEST, ST_SP_ADV, ST_PRED = 7, 8, 9, 10, 11
ST_IDX_CUR, ST_IDX_OLD, ST_IDX_BEST, ST_IDX_BEST_ADV, ST_GIDX_CUR, ST_GIDX_BEST = 12, 13, 14, 15, 16, 17
ST_ROWS = 20
LOG_MAX_SLOTS = 8
F_IMPROVED, F_WRITE_ADV, F_RESTORE = 1, 2, 4
NORMS = {'Linf': 0, 'L2': 1, 'L1': 2}
LOSSES = {'ce': 0, 'dlr': 1, 'dlr-targeted': 2}
_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}

_lib = None
LAUNCHES = {'count': 0}   # kernels launched through this ABI (bench.py reports it as gpu_launches)


# optional per-launch device timing (bench.py): CUDA events recorded on the launching stream around
# each kernel; nothing is synchronised here, the reader calls elapsed_time after its own sync
TIMING = {'enabled': False, 'events': [], 'names': None}    # names: only these kernels are timed (None = all)


class B200atError(RuntimeError):
    pass


class _Timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.on = TIMING['enabled'] and (TIMING['names'] is None or self.name in TIMING['names'])
        if self.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.b.record()
            TIMING['events'].append((self.name, self.a, self.b))
        LAUNCHES['count'] += 1
        return False


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200atError(
                f'{LIB_PATH} is not built; run `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a).  There is no CPU fallback for the attack kernels.')
        L = ctypes.CDLL(LIB_PATH)
        L.b200at_abi_version.restype = c_int
        if L.b200at_abi_version() != ABI_VERSION:
            raise B200atError(f'ABI mismatch: library {L.b200at_abi_version()} != binding {ABI_VERSION}')
        _declare(L)
        _lib = L
    return _lib


def _declare(L):
    P, F, I, I64 = c_void_p, c_float, c_int, c_int64
    sig = {
        'b200at_apgd_init': [P, P, P, I64, I64, F, F, P],
        'b200at_linf_step': [P, P, P, P, P, P, P, P, P, I64, I64, F, F, P],
        'b200at_flush_best': [P, P, P, P, I64, I64, P],
        'b200at_linf_step_log': [P, P, P, I, P, P, I64, I64, F, F, P],
        'b200at_gather_best': [P, I, P, P, P, I64, I64, P],
        'b200at_l2_step': [P, P, P, P, P, P, P, P, P, P, I64, I64, F, F, P],
        'b200at_l1_step': [P, P, P, P, P, P, P, P, P, I64, I64, F, P],
        'b200at_loss_bookkeep': [P, I, P, P, P, P, P, P, I64, I64, I, I, I, I, I, F, F, I64, P],
        'b200at_loss_bookkeep_targeted': [P, I, P, P, P, P, P, P, I64, I64, I, I, I, I, F, F, I64, P],
        'b200at_fgsm_start': [P, P, P, I64, F, F, I, P],
        'b200at_fgsm_step': [P, P, P, P, I64, F, F, I, P],
    }
    sig.update({
        'b200at_ln_fwd': [P, P, P, P, P, P, I64, I64, F, I, P],
        'b200at_ln_bwd': [P, P, P, P, P, P, P, P, P, I64, I64, I, P],
        'b200at_ln_fwd_bias': [P, P, P, P, P, P, P, I64, I64, F, I, P],
        'b200at_ln_bwd_bias': [P, P, P, P, P, P, P, P, P, P, I64, I64, I, P],
        'b200at_ln_fwd_patch2': [P, P, P, P, P, P, I64, I64, I64, I64, F, P],
        'b200at_ln_bwd_patch2': [P, P, P, P, P, P, P, P, P, I64, I64, I64, I64, P],
        'b200at_bias_gelu_fwd': [P, P, P, I64, I64, P],
        'b200at_bias_gelu_bwd': [P, P, P, P, P, I64, I64, P],
        'b200at_colsum_bf16': [P, P, I64, I64, P],
        'b200at_scale_residual_fwd': [P, P, P, P, P, I64, I64, P],
        'b200at_scale_bwd': [P, P, P, I64, I64, P],
        'b200at_add_bf16': [P, P, P, I64, P],
        'b200at_dwconv7_fwd': [P, P, P, P, P, I64, I64, I64, I64, P],
        'b200at_dwconv7_wgrad': [P, P, P, P, I64, I64, I64, I64, P],
        'b200at_gemm_bf16': [P, P, P, P, P, P, I64, I64, I64, I, P],
        'b200at_conv3x3s2_fwd': [P, P, P, I64, I64, I64, I64, I64, P],
        'b200at_gemm_gelu_grad_colsum': [P, P, P, P, P, I64, I64, I64, P],
        'b200at_mlp_fused_supported': [I64],
        'b200at_mlp_fused': [P, P, P, P, P, P, P, P, P, I64, I64, I, P],
        'b200at_normalize_nhwc_bf16': [P, P, P, P, I64, I64, I64, P],
        'b200at_prepare_mlp_weights': [P, P, P, P, P, P, P, P, P, I64, P],
        'b200at_finish_mlp_grads': [P, P, P, P, P, P, P, P, I64, P],
        'b200at_stem0_fwd': [P, P, P, P, P, P, P, P, I64, I64, I64, I64, F, P],
        'b200at_stem0_fwd_save': [P, P, P, P, P, P, P, P, P, P, P, I64, I64, I64, I64, F, P],
        'b200at_stem0_bwd_input': [P, P, P, P, P, P, P, P, P, I64, I64, I64, I64, F, P],
        'b200at_attn_fwd': [P, P, P, I64, I64, I64, F, P],
        'b200at_attn_bwd': [P, P, P, P, P, I64, I64, I64, F, P],
    })
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = c_int
    for name in OPTIONAL:
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.argtypes = OPTIONAL[name]
            fn.restype = c_int


OPTIONAL = {}


def exported_symbols():
    """Names include/*.h declare; tests check each one resolves in the .so."""
    import re
    names = set()
    inc = os.path.join(os.path.dirname(_HERE), 'include')
    for h in sorted(os.listdir(inc)):
        if h.endswith('.h'):
            with open(os.path.join(inc, h)) as f:
                names |= set(re.findall(r'^\s*int\s+(b200at_\w+)\s*\(', f.read(), flags=re.M))
    return sorted(names)


def _check(err, what):
    if err != 0:
        raise B200atError(f'{what}: CUDA error {err}')


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _dense(t):
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def _img(t, like=None, name='tensor'):
    if not t.is_cuda:
        raise B200atError(f'{name} must be a CUDA tensor (no CPU path)')
    if t.dtype != torch.float32 or not _dense(t):
        raise B200atError(f'{name} must be dense fp32, got {t.dtype} strides {t.stride()}')
    if like is not None and (t.shape != like.shape or t.stride() != like.stride()):
        raise B200atError(f'{name}: layout differs from x ({t.shape}/{t.stride()} vs {like.shape}/{like.stride()})')
    return c_void_p(t.data_ptr())


def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def apgd_init(x, x_adv, state, step0, topk0):
    B, n = x.shape[0], x[0].numel() if x.shape[0] else 0
    with _Timed('apgd_init'):
        _check(lib().b200at_apgd_init(_img(x, name='x'), _img(x_adv, x, 'x_adv'), _p(state), B, n, step0, topk0,
                                      _stream()), 'apgd_init')


def linf_step(x, x_adv, x_old, x_new, grad, x_best, grad_best, x_best_adv, state, eps, a):
    B, n = x.shape[0], x[0].numel() if x.shape[0] else 0
    args = (_img(x, name='x'), _img(x_adv, x, 'x_adv'), _img(x_old, x, 'x_old'), _img(x_new, x, 'x_new'),
            _img(grad, x, 'grad'), _img(x_best, x, 'x_best'), _img(grad_best, x, 'grad_best'),
            _img(x_best_adv, x, 'x_best_adv'), _p(state), B, n, eps, a, _stream())
    with _Timed('linf_step'):
        _check(lib().b200at_linf_step(*args), 'linf_step')


def flush_best(x_adv, x_best, x_best_adv, state):
    B, n = x_adv.shape[0], x_adv[0].numel() if x_adv.shape[0] else 0
    with _Timed('flush_best'):
        _check(lib().b200at_flush_best(_img(x_adv, name='x_adv'), _img(x_best, x_adv, 'x_best'),
                                       _img(x_best_adv, x_adv, 'x_best_adv'), _p(state), B, n, _stream()),
               'flush_best')


def loss_bookkeep(logits, y, dlogits, loss_out, state, loss_steps, it, n_iter, ckpt_k, norm, loss, step_full,
                  step_min, n_fts, y_target=None):
    if not logits.is_cuda or logits.dim() != 2 or not logits.is_contiguous() or logits.dtype not in _DT:
        raise B200atError(f'logits must be a contiguous CUDA [B,C] fp32/bf16/fp16 tensor, got '
                          f'{tuple(logits.shape)} {logits.dtype}')
    B, C = logits.shape
    if y.dtype == torch.int64 and y.dim() == 1:
        yh, ys = _p(y.contiguous()), c_void_p(0)
    elif y.dim() == 2 and y.shape == logits.shape:
        y = y.to(torch.float32).contiguous()
        yh, ys = c_void_p(0), _p(y)
    else:
        raise B200atError(f'target must be int64 [B] or soft [B,C], got {tuple(y.shape)} {y.dtype}')
    if dlogits is not None and (dlogits.dtype != logits.dtype or dlogits.shape != logits.shape
                                or not dlogits.is_contiguous()):
        raise B200atError('dlogits must match logits')
    if loss == 'dlr-targeted':
        if y_target is None or y_target.dtype != torch.int64 or y_target.shape != (B,) or y.dim() != 1:
            raise B200atError("loss 'dlr-targeted' needs hard labels and an int64 [B] y_target")
        with _Timed('loss_bookkeep'):
            _check(lib().b200at_loss_bookkeep_targeted(_p(logits), _DT[logits.dtype], yh, _p(y_target.contiguous()),
                                                       _p(dlogits), _p(loss_out), _p(state), _p(loss_steps), B, C, it,
                                                       n_iter, ckpt_k, NORMS[norm], step_full, step_min, n_fts,
                                                       _stream()), 'loss_bookkeep_targeted')
        return
    with _Timed('loss_bookkeep'):
        _check(lib().b200at_loss_bookkeep(_p(logits), _DT[logits.dtype], yh, ys, _p(dlogits), _p(loss_out),
                                          _p(state), _p(loss_steps), B, C, it, n_iter, ckpt_k, NORMS[norm],
                                          LOSSES[loss], step_full, step_min, n_fts, _stream()), 'loss_bookkeep')


def fgsm_start(x, noise, x_adv, eps, noise_level, skip_projection):
    with _Timed('fgsm_start'):
        _check(lib().b200at_fgsm_start(_img(x, name='x'), _img(noise, x, 'noise'), _img(x_adv, x, 'x_adv'),
                                       x.numel(), eps, noise_level, int(bool(skip_projection)), _stream()),
               'fgsm_start')


def fgsm_step(x, x_adv, grad, out, eps, step, skip_projection):
    with _Timed('fgsm_step'):
        _check(lib().b200at_fgsm_step(_img(x, name='x'), _img(x_adv, x, 'x_adv'), _img(grad, x, 'grad'),
                                      _img(out, x, 'out'), x.numel(), eps, step, int(bool(skip_projection)),
                                      _stream()), 'fgsm_step')


def l2_step(x, x_adv, x_old, x_new, grad, x_best, grad_best, x_best_adv, state, eps, a, scratch=None):
    B, n = x.shape[0], x[0].numel() if x.shape[0] else 0
    if scratch is None:
        scratch = torch.empty(3 * 32 * max(B, 1), device=x.device, dtype=torch.float32)
    args = (_img(x, name='x'), _img(x_adv, x, 'x_adv'), _img(x_old, x, 'x_old'), _img(x_new, x, 'x_new'),
            _img(grad, x, 'grad'), _img(x_best, x, 'x_best'), _img(grad_best, x, 'grad_best'),
            _img(x_best_adv, x, 'x_best_adv'), _p(state), _p(scratch), B, n, eps, a, _stream())
    with _Timed('l2_step'):
        _check(lib().b200at_l2_step(*args), 'l2_step')             # one cluster launch (four with B200AT_L2_PHASES=4)
    return scratch


L1_LAUNCHES = 3 + 1 + 7 + 1    # histogram x3, sums, sectioning x7, final (+ one memset)


def l1_step(x, x_adv, x_new, grad, x_best, grad_best, x_best_adv, state, eps, scratch=None):
    B, n = x.shape[0], x[0].numel() if x.shape[0] else 0
    if x_new.data_ptr() == x_adv.data_ptr():
        raise B200atError('l1_step: x_new must not alias x_adv')
    if scratch is None:
        scratch = torch.empty((3 * 2048 + 32 + 64 + 2048) * max(B, 1), device=x.device, dtype=torch.int32)
    args = (_img(x, name='x'), _img(x_adv, x, 'x_adv'), _img(x_new, x, 'x_new'), _img(grad, x, 'grad'),
            _img(x_best, x, 'x_best'), _img(grad_best, x, 'grad_best'), _img(x_best_adv, x, 'x_best_adv'),
            _p(state), _p(scratch), B, n, eps, _stream())
    with _Timed('l1_step'):
        _check(lib().b200at_l1_step(*args), 'l1_step')
    LAUNCHES['count'] += L1_LAUNCHES - 1
    return scratch


# ---------------------------------------------------------------------------------------------------
# model-op kernels (include/b200at_model.h).  Activations: contiguous CUDA bf16; parameters: fp32.
def _act(t, name):
    if not t.is_cuda or t.dtype != torch.bfloat16 or not t.is_contiguous():
        raise B200atError(f'{name} must be a contiguous CUDA bf16 tensor, got {t.dtype} {tuple(t.shape)} '
                          f'strides {t.stride()} on {t.device} (no CPU path)')
    return c_void_p(t.data_ptr())


def _par(t, name, n=None):
    if t is None:
        return c_void_p(0)
    if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or (n is not None and t.numel() != n):
        raise B200atError(f'{name} must be a contiguous CUDA fp32 tensor' + (f' of {n} elements' if n else ''))
    return c_void_p(t.data_ptr())


def ln_fwd(x, w, b, y, mean, rstd, eps, gelu, pre_bias=None):
    """y = LN(x [+ pre_bias]) [-> GELU]; pre_bias: the bias of a library convolution in front (fp32 [C])"""
    C = x.shape[-1]
    M = x.numel() // C
    with _Timed('ln_fwd'):
        if pre_bias is None:
            _check(lib().b200at_ln_fwd(_act(x, 'x'), _par(w, 'w', C), _par(b, 'b', C), _act(y, 'y'),
                                       _par(mean, 'mean', M), _par(rstd, 'rstd', M), M, C, eps, int(gelu), _stream()),
                   'ln_fwd')
        else:
            _check(lib().b200at_ln_fwd_bias(_act(x, 'x'), _par(pre_bias, 'pre_bias', C), _par(w, 'w', C), _par(b, 'b', C),
                                            _act(y, 'y'), _par(mean, 'mean', M), _par(rstd, 'rstd', M), M, C, eps,
                                            int(gelu), _stream()), 'ln_fwd_bias')


def ln_bwd(dy, x, w, b, mean, rstd, dx, dw, db, gelu, pre_bias=None):
    C = x.shape[-1]
    M = x.numel() // C
    with _Timed('ln_bwd'):
        if pre_bias is None:
            _check(lib().b200at_ln_bwd(_act(dy, 'dy'), _act(x, 'x'), _par(w, 'w', C), _par(b, 'b', C),
                                       _par(mean, 'mean', M), _par(rstd, 'rstd', M), _act(dx, 'dx'), _par(dw, 'dw', C),
                                       _par(db, 'db', C), M, C, int(gelu), _stream()), 'ln_bwd')
        else:
            _check(lib().b200at_ln_bwd_bias(_act(dy, 'dy'), _act(x, 'x'), _par(pre_bias, 'pre_bias', C), _par(w, 'w', C),
                                            _par(b, 'b', C), _par(mean, 'mean', M), _par(rstd, 'rstd', M), _act(dx, 'dx'),
                                            _par(dw, 'dw', C), _par(db, 'db', C), M, C, int(gelu), _stream()),
                   'ln_bwd_bias')


def ln_fwd_patch2(x, w, b, y, mean, rstd, eps):
    """LayerNorm over C of NHWC x [B,H,W,C], written in the 2x2-patch layout y [B*H/2*W/2, 4C]."""
    B, H, W, C = x.shape
    M = B * H * W
    if y.numel() != x.numel():
        raise B200atError('ln_fwd_patch2: y size')
    with _Timed('ln_fwd_patch2'):
        _check(lib().b200at_ln_fwd_patch2(_act(x, 'x'), _par(w, 'w', C), _par(b, 'b', C), _act(y, 'y'),
                                          _par(mean, 'mean', M), _par(rstd, 'rstd', M), B, H, W, C, eps, _stream()),
               'ln_fwd_patch2')


def ln_bwd_patch2(dy, x, w, b, mean, rstd, dx, dw, db):
    B, H, W, C = x.shape
    M = B * H * W
    if dy.numel() != x.numel() or dx.shape != x.shape:
        raise B200atError('ln_bwd_patch2: sizes')
    with _Timed('ln_bwd_patch2'):
        _check(lib().b200at_ln_bwd_patch2(_act(dy, 'dy'), _act(x, 'x'), _par(w, 'w', C), _par(b, 'b', C),
                                          _par(mean, 'mean', M), _par(rstd, 'rstd', M), _act(dx, 'dx'),
                                          _par(dw, 'dw', C), _par(db, 'db', C), B, H, W, C, _stream()), 'ln_bwd_patch2')


def bias_gelu_fwd(z, bias, h):
    M, N = z.shape
    with _Timed('bias_gelu_fwd'):
        _check(lib().b200at_bias_gelu_fwd(_act(z, 'z'), _par(bias, 'bias', N), _act(h, 'h'), M, N, _stream()), 'bias_gelu_fwd')


def bias_gelu_bwd(dh, z, bias, dz, dbias=None):
    """dz = dh * GELU'(z + bias); dbias (fp32 [N], zeroed by the caller) also receives the column sums of dz."""
    M, N = z.shape
    with _Timed('bias_gelu_bwd'):
        _check(lib().b200at_bias_gelu_bwd(_act(dh, 'dh'), _act(z, 'z'), _par(bias, 'bias', N), _act(dz, 'dz'),
                                          _par(dbias, 'dbias', N), M, N, _stream()), 'bias_gelu_bwd')


def prepare_mlp_weights(w1, w2, b2, gamma):
    """One launch: {w1b, w1t, w2g, w2gt (bf16), b2g, gf (fp32)} of a ConvNeXt block (ops._prepared)."""
    C = gamma.numel()
    dev = w1.device
    w1, w2, b2, gamma = (t.detach().float().contiguous() for t in (w1, w2, b2, gamma))
    out = dict(w1b=torch.empty(4 * C, C, device=dev, dtype=torch.bfloat16), w1t=torch.empty(C, 4 * C, device=dev, dtype=torch.bfloat16),
               w2g=torch.empty(C, 4 * C, device=dev, dtype=torch.bfloat16), w2gt=torch.empty(4 * C, C, device=dev, dtype=torch.bfloat16),
               b2g=torch.empty(C, device=dev, dtype=torch.float32), gf=gamma)
    with _Timed('prepare_mlp_weights'):
        _check(lib().b200at_prepare_mlp_weights(_par(w1, 'w1', 4 * C * C), _par(w2, 'w2', 4 * C * C), _par(b2, 'b2', C),
                                                _par(gamma, 'gamma', C), _act(out['w1b'], 'w1b'), _act(out['w1t'], 'w1t'),
                                                _act(out['w2g'], 'w2g'), _act(out['w2gt'], 'w2gt'), _par(out['b2g'], 'b2g', C),
                                                C, _stream()), 'prepare_mlp_weights')
    return out


def finish_mlp_grads(dw2g, w2, col, b2, gamma):
    """(dW2, db2, dgamma) from the gradient w.r.t. the layer-scale-folded pwconv2 weight, one launch."""
    C = gamma.numel()
    dw2 = torch.empty_like(dw2g)
    db2 = torch.empty(C, device=dw2g.device, dtype=torch.float32)
    dgamma = torch.empty_like(db2)
    with _Timed('finish_mlp_grads'):
        _check(lib().b200at_finish_mlp_grads(_par(dw2g, 'dw2g', 4 * C * C), _par(w2, 'w2', 4 * C * C), _par(col, 'col', C),
                                             _par(b2, 'b2', C), _par(gamma, 'gamma', C), _par(dw2, 'dw2', 4 * C * C),
                                             _par(db2, 'db2', C), _par(dgamma, 'dgamma', C), C, _stream()), 'finish_mlp_grads')
    return dw2, db2, dgamma


def colsum_bf16(a, out):
    """out[n] += sum_m a[m][n]"""
    M, N = a.shape
    with _Timed('colsum_bf16'):
        _check(lib().b200at_colsum_bf16(_act(a, 'a'), _par(out, 'out', N), M, N, _stream()), 'colsum_bf16')


def scale_residual_fwd(z, bias, gamma, res, out):
    M, N = z.shape
    with _Timed('scale_residual_fwd'):
        _check(lib().b200at_scale_residual_fwd(_act(z, 'z'), _par(bias, 'bias', N), _par(gamma, 'gamma', N),
                                               _act(res, 'res'), _act(out, 'out'), M, N, _stream()), 'scale_residual_fwd')


def scale_bwd(dout, gamma, dz):
    M, N = dout.shape
    with _Timed('scale_bwd'):
        _check(lib().b200at_scale_bwd(_act(dout, 'dout'), _par(gamma, 'gamma', N), _act(dz, 'dz'), M, N, _stream()),
               'scale_bwd')


def add_bf16(a, b, c):
    with _Timed('add_bf16'):
        _check(lib().b200at_add_bf16(_act(a, 'a'), _act(b, 'b'), _act(c, 'c'), a.numel(), _stream()), 'add_bf16')


def dwconv7_fwd(x, wt, bias, y, add=None):
    B, H, W, C = x.shape
    with _Timed('dwconv7'):
        _check(lib().b200at_dwconv7_fwd(_act(x, 'x'), _par(wt, 'wt', 49 * C), _par(bias, 'bias', C),
                                        _act(add, 'add') if add is not None else c_void_p(0), _act(y, 'y'),
                                        B, H, W, C, _stream()), 'dwconv7_fwd')


def dwconv7_wgrad(x, dy, dw, db):
    B, H, W, C = x.shape
    with _Timed('dwconv7_wgrad'):
        _check(lib().b200at_dwconv7_wgrad(_act(x, 'x'), _act(dy, 'dy'), _par(dw, 'dw', 49 * C), _par(db, 'db', C),
                                          B, H, W, C, _stream()), 'dwconv7_wgrad')


def _host3(v):
    return (c_float * 3)(*[float(t) for t in v]) if v is not None else None


def _x_nchw(x):
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3 or not x.is_contiguous():
        raise B200atError(f'stem0: x must be a contiguous CUDA fp32 NCHW image batch with 3 channels, got {x.dtype} '
                          f'{tuple(x.shape)} strides {x.stride()} on {x.device}')
    return c_void_p(x.data_ptr())


def stem0_fwd(x, mean3, std3, wk, bias, ln_w, ln_b, y, eps=1e-6):
    """y[B,Ho,Wo,C0] bf16 NHWC = GELU(LN(conv3x3s2((x - mean) / std) + bias)); mean3 / std3: 3 python floats or None."""
    B, _, H, W = x.shape
    C0 = bias.numel()
    if tuple(y.shape) != (B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C0):
        raise B200atError(f'stem0_fwd: y shape {tuple(y.shape)}')
    m, sd = _host3(mean3), _host3(std3)
    with _Timed('stem0_fwd'):
        _check(lib().b200at_stem0_fwd(_x_nchw(x), m, sd, _par(wk, 'wk', 27 * C0), _par(bias, 'bias', C0),
                                      _par(ln_w, 'ln_w', C0), _par(ln_b, 'ln_b', C0), _act(y, 'y'), B, H, W, C0, eps,
                                      _stream()), 'stem0_fwd')


def stem0_fwd_save(x, mean3, std3, wk, bias, ln_w, ln_b, y, y_pre, mean, rstd, eps=1e-6):
    """stem0_fwd for the training forward: also y_pre (conv output without bias, bf16) and the LayerNorm statistics."""
    B, _, H, W = x.shape
    C0 = bias.numel()
    shp = (B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C0)
    if tuple(y.shape) != shp or tuple(y_pre.shape) != shp:
        raise B200atError(f'stem0_fwd_save: y shape {tuple(y.shape)} / {tuple(y_pre.shape)}')
    M = shp[0] * shp[1] * shp[2]
    m, sd = _host3(mean3), _host3(std3)
    with _Timed('stem0_fwd_save'):
        _check(lib().b200at_stem0_fwd_save(_x_nchw(x), m, sd, _par(wk, 'wk', 27 * C0), _par(bias, 'bias', C0),
                                           _par(ln_w, 'ln_w', C0), _par(ln_b, 'ln_b', C0), _act(y, 'y'), _act(y_pre, 'y_pre'),
                                           _par(mean, 'mean', M), _par(rstd, 'rstd', M), B, H, W, C0, eps, _stream()),
               'stem0_fwd_save')


def normalize_nhwc_bf16(x, mean3, std3, y):
    """y[B,H,W,3] bf16 = (x[B,3,H,W] - mean) / std (fp32 arithmetic, one rounding); mean3 / std3: 3 python floats or None."""
    B, _, H, W = x.shape
    if tuple(y.shape) != (B, H, W, 3):
        raise B200atError(f'normalize_nhwc_bf16: y shape {tuple(y.shape)}')
    with _Timed('normalize_nhwc_bf16'):
        _check(lib().b200at_normalize_nhwc_bf16(_x_nchw(x), _host3(mean3), _host3(std3), _act(y, 'y'), B, H, W, _stream()),
               'normalize_nhwc_bf16')


def stem0_bwd_input(dy, x, mean3, std3, wk, bias, ln_w, ln_b, dx, eps=1e-6):
    B, _, H, W = x.shape
    C0 = bias.numel()
    if tuple(dy.shape) != (B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C0) or dx.shape != x.shape:
        raise B200atError(f'stem0_bwd_input: dy {tuple(dy.shape)} dx {tuple(dx.shape)}')
    m, sd = _host3(mean3), _host3(std3)
    with _Timed('stem0_bwd_input'):
        _check(lib().b200at_stem0_bwd_input(_act(dy, 'dy'), _x_nchw(x), m, sd, _par(wk, 'wk', 27 * C0),
                                            _par(bias, 'bias', C0), _par(ln_w, 'ln_w', C0), _par(ln_b, 'ln_b', C0),
                                            _x_nchw(dx), B, H, W, C0, eps, _stream()), 'stem0_bwd_input')


def _slot_ptrs(ts, like, name):
    return (c_void_p * len(ts))(*[_img(t, like, f'{name}[{i}]').value for i, t in enumerate(ts)])


def linf_step_log(x, x_slots, g_slots, x_new, state, eps, a):
    """x_slots / g_slots: python lists of the iterate / gradient tensors, slot k first."""
    B, n = x.shape[0], x[0].numel() if x.shape[0] else 0
    g_full = list(g_slots) + [g_slots[0]] * (len(x_slots) - len(g_slots))
    args = (_img(x, name='x'), _slot_ptrs(x_slots, x, 'x_slots'), _slot_ptrs(g_full, x, 'g_slots'), len(x_slots),
            _img(x_new, x, 'x_new'), _p(state), B, n, eps, a, _stream())
    with _Timed('linf_step_log_first' if a == 1.0 else 'linf_step_log'):
        _check(lib().b200at_linf_step_log(*args), 'linf_step_log')


def gather_best(x_slots, x_best, x_best_adv, state):
    B, n = x_best.shape[0], x_best[0].numel() if x_best.shape[0] else 0
    with _Timed('gather_best'):
        _check(lib().b200at_gather_best(_slot_ptrs(x_slots, x_best, 'x_slots'), len(x_slots), _img(x_best, name='x_best'),
                                        _img(x_best_adv, x_best, 'x_best_adv'), _p(state), B, n, _stream()),
               'gather_best')


EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_RESIDUAL, EPI_GELU_GRAD = range(5)


def gemm_bf16(a, b, c, epilogue=EPI_NONE, bias=None, aux=None, c2=None):
    """c[M,N] = epilogue(a[M,K] @ b[N,K]^T) on the tcgen05 kernel (include/b200at_model.h)."""
    M, K = a.shape
    N = b.shape[0]
    if b.shape[1] != K or tuple(c.shape) != (M, N):
        raise B200atError(f'gemm shapes: a {tuple(a.shape)} b {tuple(b.shape)} c {tuple(c.shape)}')
    for t, nm in ((aux, 'aux'), (c2, 'c2')):
        if t is not None and tuple(t.shape) != (M, N):
            raise B200atError(f'gemm {nm} must be [M,N]')
    with _Timed('gemm_bf16'):
        _check(lib().b200at_gemm_bf16(_act(a, 'a'), _act(b, 'b'), _act(c, 'c'),
                                      _act(c2, 'c2') if c2 is not None else c_void_p(0),
                                      _act(aux, 'aux') if aux is not None else c_void_p(0),
                                      _par(bias, 'bias', N), M, N, K, epilogue, _stream()), 'gemm_bf16')


def gemm_gelu_grad_colsum(a, b, c, aux, colsum):
    """c[M,N] = (a @ b^T) * GELU'(aux) and colsum[N] += column sums of c (include/b200at_model.h)."""
    M, K = a.shape
    N = b.shape[0]
    if b.shape[1] != K or tuple(c.shape) != (M, N) or tuple(aux.shape) != (M, N):
        raise B200atError(f'gemm_gelu_grad_colsum shapes: a {tuple(a.shape)} b {tuple(b.shape)} c {tuple(c.shape)}')
    with _Timed('gemm_gelu_grad_colsum'):
        _check(lib().b200at_gemm_gelu_grad_colsum(_act(a, 'a'), _act(b, 'b'), _act(c, 'c'), _act(aux, 'aux'),
                                                  _par(colsum, 'colsum', N), M, N, K, _stream()), 'gemm_gelu_grad_colsum')


def conv3x3s2_fwd(x, wk, y):
    """y[B,H/2,W/2,Co] = conv3x3 stride 2 pad 1 of x[B,H,W,Ci] (NHWC bf16, no bias) as an implicit GEMM on the tcgen05 kernel
    (include/b200at_model.h: b200at_conv3x3s2_fwd).  wk [Co, 9*64] bf16.  Returns False when the kernel does not take the shape."""
    B, H, W, Ci = x.shape
    Co = wk.shape[0]
    if tuple(wk.shape) != (Co, 576) or tuple(y.shape) != (B, H // 2, W // 2, Co):
        raise B200atError(f'conv3x3s2 shapes: x {tuple(x.shape)} wk {tuple(wk.shape)} y {tuple(y.shape)}')
    with _Timed('conv3x3s2_fwd'):
        rc = lib().b200at_conv3x3s2_fwd(_act(x, 'x'), _act(wk, 'wk'), _act(y, 'y'), B, H, W, Ci, Co, _stream())
    if rc == -1:
        return False
    _check(rc, 'conv3x3s2_fwd')
    return True


def mlp_fused_supported(C):
    return bool(lib().b200at_mlp_fused_supported(C))


def mlp_fused(a, wa, wb, bias1, z, out, bias2=None, residual=None, p_out=None, backward=False):
    """The block's MLP in one tcgen05 kernel (include/b200at_model.h: b200at_mlp_fused).  a [M,C], wa [4C,C], wb [C,4C],
    z [M,4C] (written forward / read backward), out [M,C]; p_out [M,4C] optional (GELU output / dz)."""
    M, C = a.shape
    if tuple(wa.shape) != (4 * C, C) or tuple(wb.shape) != (C, 4 * C) or tuple(z.shape) != (M, 4 * C) \
            or tuple(out.shape) != (M, C):
        raise B200atError(f'mlp_fused shapes: a {tuple(a.shape)} wa {tuple(wa.shape)} wb {tuple(wb.shape)} '
                          f'z {tuple(z.shape)} out {tuple(out.shape)}')
    for t, nm, shp in ((residual, 'residual', (M, C)), (p_out, 'p_out', (M, 4 * C))):
        if t is not None and tuple(t.shape) != shp:
            raise B200atError(f'mlp_fused {nm} must be {shp}')
    null = c_void_p(0)
    with _Timed('mlp_fused_bwd' if backward else 'mlp_fused_fwd'):
        _check(lib().b200at_mlp_fused(_act(a, 'a'), _act(wa, 'wa'), _act(wb, 'wb'), _par(bias1, 'bias1', 4 * C),
                                      _par(bias2, 'bias2', C) if bias2 is not None else null,
                                      _act(residual, 'residual') if residual is not None else null,
                                      _act(z, 'z'), _act(p_out, 'p_out') if p_out is not None else null,
                                      _act(out, 'out'), M, C, 1 if backward else 0, _stream()), 'mlp_fused')


def attn_fwd(qkv, o, lse, heads, scale):
    """o = softmax(q k^T scale) v per head; qkv [B,N,3*heads*64] bf16, o [B,N,heads*64], lse fp32 [B,heads,N]."""
    B, N, C3 = qkv.shape
    if C3 != 3 * heads * 64 or tuple(o.shape) != (B, N, heads * 64):
        raise B200atError(f'attention shapes: qkv {tuple(qkv.shape)} o {tuple(o.shape)} heads {heads} (head dim 64)')
    with _Timed('attn_fwd'):
        _check(lib().b200at_attn_fwd(_act(qkv, 'qkv'), _act(o, 'o'), _par(lse, 'lse', B * heads * N), B, N, heads,
                                     scale, _stream()), 'attn_fwd')


def attn_bwd(qkv, o, d_o, lse, dqkv, heads, scale):
    B, N, C3 = qkv.shape
    if C3 != 3 * heads * 64 or o.shape != d_o.shape or dqkv.shape != qkv.shape:
        raise B200atError('attention backward shapes')
    with _Timed('attn_bwd'):
        _check(lib().b200at_attn_bwd(_act(qkv, 'qkv'), _act(o, 'o'), _act(d_o, 'd_o'), _par(lse, 'lse', B * heads * N),
                                     _act(dqkv, 'dqkv'), B, N, heads, scale, _stream()), 'attn_bwd')


def l1_projection(x2, y2, eps1):
    """`L1_projection(x2, y2, eps1)` (autopgd_train_clean.py:24-91): the correction delta with x2 + y2 + delta inside
    {||. - x2||_1 <= eps1} and [0,1]^n, through the projection half of the l1 step kernels (a zero gradient makes the
    step half the identity, so the kernels see u = x2 + y2 and project it).  Tolerance-level (1e-6), like every l1
    result."""
    if not x2.is_cuda:
        raise B200atError('L1_projection: CUDA tensors only (no CPU path)')
    x = x2.detach().to(torch.float32).contiguous()
    u = (x + y2.detach().to(torch.float32)).contiguous()
    B = x.shape[0]
    state = torch.zeros(ST_ROWS, max(B, 1), device=x.device, dtype=torch.float32)
    out = torch.empty_like(x)
    l1_step(x, u, out, torch.zeros_like(x), out, out, out, state, float(eps1))
    return (out - u).view_as(x2)
