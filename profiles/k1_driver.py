"""Kernel-only driver for the attack-step kernels at BASELINE sizes (synthetic inputs of SURVEY.md 8d).

    python profiles/k1_driver.py [--batch 128] [--res 224] [--iters 20] [--json]

Times b200at_linf_step (and friends) with CUDA events after warm-up; working set 385 MB+ > L2, so every
launch streams from HBM.  Also the target of the `ncu --set full -k regex:linf_step` capture."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import revisiting_at_b200  # noqa: E402,F401
from revisiting_at_b200 import _abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--res', type=int, default=224)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--json', action='store_true')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    B, shape, eps = a.batch, (3, a.res, a.res), 4 / 255.
    n = 3 * a.res * a.res
    g = torch.Generator(device='cuda').manual_seed(1234)
    x = torch.rand(B, *shape, generator=g, device=dev)
    xa = (x + (torch.rand(B, *shape, generator=g, device=dev) * 2 - 1) * eps).clamp(0, 1)
    xo = (xa - (torch.rand(B, *shape, generator=g, device=dev) * 2 - 1) * eps).clamp(0, 1)
    gr = torch.randn(B, *shape, generator=g, device=dev) * 1e-3
    gr[torch.rand(B, *shape, generator=g, device=dev) < 0.1] = 0.
    xb, gb, xba = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    st = torch.zeros(_abi.ST_ROWS, B, device=dev)
    st[_abi.ST_STEP] = torch.tensor([2 * eps, eps, eps / 2] * B, device=dev)[:B]
    res = {}

    def timeit(name, fn, bytes_alg):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        res[name] = {'ms': ms, 'GBps_algorithmic': bytes_alg / ms / 1e6}

    flags = st[_abi.ST_FLAGS].view(torch.int32)
    flags.zero_()
    timeit('linf_step_steady(20B/elt)', lambda: _abi.linf_step(x, xa, xo, xo, gr, xb, gb, xba, st, eps, 0.75), 20. * B * n)
    flags.fill_(3)
    timeit('linf_step_first(seed best: 12B rd + 16B wr)', lambda: _abi.linf_step(x, xa, xa, xo, gr, xb, gb, xba, st, eps, 1.0), 28. * B * n)
    flags.copy_(torch.arange(B, device=dev, dtype=torch.int32) % 8)
    timeit('linf_step_mixed_flags(20B/elt credited)', lambda: _abi.linf_step(x, xa, xo, xo, gr, xb, gb, xba, st, eps, 0.75), 20. * B * n)
    xs = [xa, xo, xb]
    gs = [gr, gb]
    for r, v in ((_abi.ST_IDX_CUR, 1), (_abi.ST_IDX_OLD, 0), (_abi.ST_GIDX_CUR, 1)):
        st[r].view(torch.int32).fill_(v)
    gb.copy_(gr)
    timeit('linf_step_log_steady(20B/elt)', lambda: _abi.linf_step_log(x, xs, gs, xba, st, eps, 0.75), 20. * B * n)
    st[_abi.ST_IDX_OLD].view(torch.int32).fill_(1)
    timeit('linf_step_log_first(16B/elt)', lambda: _abi.linf_step_log(x, xs, gs, xba, st, eps, 1.0), 16. * B * n)
    st[_abi.ST_IDX_CUR].view(torch.int32).copy_(torch.arange(B, device=dev, dtype=torch.int32) % 2)
    st[_abi.ST_IDX_OLD].view(torch.int32).copy_((torch.arange(B, device=dev, dtype=torch.int32) + 1) % 2)
    timeit('linf_step_log_mixed_slots(20B/elt)', lambda: _abi.linf_step_log(x, xs, gs, xba, st, eps, 0.75), 20. * B * n)
    st[_abi.ST_IDX_BEST].view(torch.int32).copy_(torch.arange(B, device=dev, dtype=torch.int32) % 3)
    st[_abi.ST_IDX_BEST_ADV].view(torch.int32).copy_((torch.arange(B, device=dev, dtype=torch.int32) // 3) % 3)
    out1, out2 = torch.empty_like(x), torch.empty_like(x)
    timeit('gather_best(16B/elt)', lambda: _abi.gather_best(xs, out1, out2, st), 16. * B * n)
    flags.fill_(3)
    timeit('flush_best(all flagged, 12B/elt)', lambda: _abi.flush_best(xa, xb, xba, st), 12. * B * n)
    timeit('apgd_init(8B/elt)', lambda: _abi.apgd_init(x, xa, st, 2 * eps, 0.), 8. * B * n)
    timeit('torch_copy(8B/elt, reference point)', lambda: xb.copy_(x), 8. * B * n)
    z = torch.randn(B, 1000, device=dev).bfloat16()
    y = torch.randint(0, 1000, (B,), device=dev)
    dl = torch.empty_like(z)
    ls = torch.zeros(2, B, device=dev)
    timeit('loss_bookkeep(bf16 logits)', lambda: _abi.loss_bookkeep(z, y, dl, None, st, ls, 0, 2, 1, 'Linf', 'ce', 2 * eps, 0., n), 4. * B * 1000)
    # l2 / l1 moves (the same 20 B/element single-pass ideal is the denominator; the extra passes of the
    # dependent per-sample reductions count against the kernel)
    flags.zero_()
    st[_abi.ST_STEP] = torch.tensor([2 * 0.5, 0.5, 0.25] * B, device=dev)[:B]
    xn = torch.empty_like(x)
    sc = {'l2': None, 'l1': None}

    def l2():
        sc['l2'] = _abi.l2_step(x, xa, xo, xn, gr, xb, gb, xba, st, 0.5, 0.75, sc['l2'])
    timeit('l2_step_steady(20B/elt credited)', l2, 20. * B * n)
    st[_abi.ST_STEP] = 12.0
    st[_abi.ST_TOPK] = 0.05

    def l1():
        sc['l1'] = _abi.l1_step(x, xa, xn, gr, xb, gb, xba, st, 12.0, sc['l1'])
    timeit('l1_step_steady(20B/elt credited)', l1, 20. * B * n)
    if a.json:
        print(json.dumps({'batch': B, 'res': a.res, 'results': res}))
    else:
        for k, v in res.items():
            print(f'{k:50s} {v["ms"] * 1e3:9.1f} us  {v["GBps_algorithmic"]:8.1f} GB/s')


if __name__ == '__main__':
    main()
