"""Adversarial train step: the caller side of the attack.

`WrappedModel` mirrors /root/reference/main.py:260-301 (attack inside the forward of the wrapped
module: eval -> perturb -> train -> forward on element [0] of the attack's return).  `AdvTrainStep`
restates the body of `ImageNetTrainer.train_loop` (main.py:961-997): autocast forward (incl. attack),
loss, backward with DDP's bucketed NCCL gradient all-reduce, AdamW step, optional EMA.  Differences
from the reference, all deliberate: bf16 autocast without GradScaler instead of
fp16 + GradScaler; EMA kept on the device (the reference's `ModelEmaV2(device='cpu')` forces a
full-parameter D2H copy every step).
"""
from __future__ import annotations

import time
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _abi, ops
from .attack import apgd_train
from .fgsm import fgsm_train


class WrappedModel(nn.Module):
    """include the generation of the adversarial perturbation in the forward pass (main.py:260-301)."""

    def __init__(self, base_model, perturb, verbose=False):
        super().__init__()
        self.base_model = base_model
        self.perturb = perturb
        self.perturb_input = False
        self.verbose = verbose

    def forward(self, x, y=None):
        if self.perturb_input:
            assert y is not None
            self.base_model.eval()                         # attack runs in eval mode (main.py:279)
            if self.verbose:
                print('perturb input')
                t0 = time.time()
            z = self.perturb(self.base_model, x, y)
            if self.verbose:
                print(f'inference time={time.time() - t0:.5f}')
            self.base_model.train()
            if isinstance(z, (tuple, list)):
                z = z[0]                                   # x_best, not x_best_adv
            return self.base_model(z)
        if self.verbose:
            print('clean inference')
        return self.base_model(x)

    def set_perturb(self, mode):
        self.perturb_input = mode


def make_attack(attack='apgd', norm='Linf', eps=4. / 255., n_iter=2, verbose=False, mixup_fn=None, alpha=1.25,
                noise_level=1., skip_projection=0):
    """The wiring of main.py:831-842 (`adv.*` config keys -> functools.partial)."""
    if attack == 'apgd':
        return partial(apgd_train, norm=norm, eps=eps, n_iter=n_iter, verbose=verbose, mixup=mixup_fn)
    if attack == 'fgsm':
        return partial(fgsm_train, eps=eps, use_rs=True, alpha=alpha, noise_level=noise_level,
                       skip_projection=skip_projection == 1)
    if attack == 'none':
        return None
    raise ValueError(attack)


class GraphedAttack:
    """`perturb(model, x, y)` replayed from a CUDA graph.

    The attack is a fixed launch sequence: no host synchronisation, data-independent checkpoint schedule, every
    data-dependent branch of the reference is a per-sample predicate on the device (attack.py).  With n_iter=2
    it is ~750 launches of 10-150 us kernels per call, so at 1 process per GPU the Python/launch path is a visible
    share of the step; one `cudaGraphLaunch` removes it.  The first `warmup` calls per input signature run
    eagerly (cuDNN/cuBLAS heuristics, lazy module loading), the next one is captured -- after forgetting the
    cached bf16/transposed parameter copies, so their re-derivation from the fp32 master parameters is part of
    the graph and a replay after an optimiser step sees the new weights.  Inputs are copied into static buffers;
    the returned tensors are the graph's static outputs, valid until the next call with the same signature."""

    def __init__(self, perturb, warmup=2):
        self.perturb = perturb
        self.warmup = warmup
        self.seen = {}
        self.graphs = {}

    def __call__(self, model, x, y):
        key = (tuple(x.shape), x.dtype, tuple(x.stride()), tuple(y.shape), y.dtype, x.device.index, id(model))
        hit = self.graphs.get(key)
        if hit is None:
            n = self.seen.get(key, 0)
            self.seen[key] = n + 1
            if n < self.warmup or not x.is_cuda:
                return self.perturb(model, x, y)
            sx, sy = x.detach().clone(), y.detach().clone()
            ops.invalidate_derived()
            graph = torch.cuda.CUDAGraph()
            n0 = _abi.LAUNCHES['count']
            with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                out = self.perturb(model, sx, sy)
            hit = self.graphs[key] = (graph, sx, sy, out, _abi.LAUNCHES['count'] - n0, ops.snapshot_derived())
        graph, sx, sy, out, n_kernels, derived = hit
        sx.copy_(x, non_blocking=True)
        sy.copy_(y, non_blocking=True)
        graph.replay()
        ops.install_derived(derived)                      # the replay just re-derived these from the current parameters
        _abi.LAUNCHES['count'] += n_kernels               # kernels of this library inside the replayed graph
        return out


class GraphedStep:
    """The WHOLE training step -- attack, training forward, backward, gradient all-reduce, fused AdamW, EMA -- replayed from
    one CUDA graph (torch's whole-network capture recipe: static input buffers, `zero_grad(set_to_none=True)`, capturable
    fused optimiser).  The outer forward / backward is ~400 launches of 3-100 us kernels driven from Python autograd; at
    one process per GPU that launch path is a visible share of the step (the DistributedDataParallel experiment shows it:
    1.25 ms of pure host-side hook work appeared 1:1 in the step time, profiles/r02_ddp_n2_variants.txt).
    The first `warmup` calls per input signature run eagerly, the next one is captured.  The learning rate is a device
    tensor (`optimizer.param_groups[i]['lr']`), so schedules keep working: `set_lr` writes it."""

    def __init__(self, step, warmup=2):
        self.step, self.warmup = step, warmup
        self.seen, self.graphs = {}, {}

    def __call__(self, images, target):
        step = self.step
        key = (tuple(images.shape), images.dtype, tuple(target.shape), target.dtype)
        hit = self.graphs.get(key)
        if hit is None:
            n = self.seen.get(key, 0)
            self.seen[key] = n + 1
            if n < self.warmup or not images.is_cuda:
                return step.eager_step(images, target)
            sx, sy = images.detach().clone(), target.detach().clone()
            step.use_graph(False)                          # the attack is recorded inline: captures do not nest
            ops.invalidate_derived()
            step.optimizer.zero_grad(set_to_none=True)
            graph = torch.cuda.CUDAGraph()
            n0 = _abi.LAUNCHES['count']
            with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                loss = step.eager_step(sx, sy)
            hit = self.graphs[key] = (graph, sx, sy, loss, _abi.LAUNCHES['count'] - n0)
        graph, sx, sy, loss, n_kernels = hit
        sx.copy_(images, non_blocking=True)
        sy.copy_(target, non_blocking=True)
        graph.replay()
        ops.invalidate_derived()                           # eager consumers (validation) must re-derive from the new weights
        _abi.LAUNCHES['count'] += n_kernels
        return loss


class DevicePrefetcher:
    """Pinned host batches -> device, one batch ahead on a copy stream: the device side of the reference's input path
    (`DataLoader(pin_memory=True)` workers + `images.cuda(non_blocking=True)`, main.py:961-966), so that the 77 MB
    image copy of step i+1 runs under the compute of step i instead of in front of it."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device)
        self.queue = []

    def submit(self, *host_tensors):
        with torch.cuda.stream(self.stream):
            dev = tuple(t.to(self.device, non_blocking=True) for t in host_tensors)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((dev, ev))

    def get(self):
        dev, ev = self.queue.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)                       # allocated on the copy stream, consumed on the compute stream
        return dev


class DeviceEma:
    """EMA of the parameters kept on the device: one fused multi-tensor lerp per step
    (replaces timm ModelEmaV2(decay=0.9999, device='cpu'), main.py:882-887,996-997)."""

    def __init__(self, model, decay=0.9999):
        self.decay = decay
        self.model = model
        self.params = [p for p in model.parameters()]
        self.shadow = [p.detach().clone() for p in self.params]

    @torch.no_grad()
    def update(self):
        torch._foreach_lerp_(self.shadow, [p.detach() for p in self.params], 1. - self.decay)

    def state_dict(self):
        """EMA weights (+ the model's buffers) under the model's own key names: what timm's
        `get_state_dict(model_ema)` hands to `weights_ema_N.pt` (main.py:741)."""
        shadow = {id(p): s for p, s in zip(self.params, self.shadow)}
        out = {}
        for k, v in self.model.state_dict(keep_vars=True).items():
            out[k] = shadow.get(id(v), v).detach().clone()
        return out


class FlatGradAllReduce:
    """The step's one collective (main.py:890,992: DDP's mean of the fp32 weight gradients over the ranks) as ONE NCCL
    all-reduce over a flat fp32 buffer, issued when the backward has finished.

    Why not `DistributedDataParallel`: measured at 2 GPUs (profiles/r02_ddp_n2_variants.txt) the wrapper costs 1.25 ms per
    33 ms step whatever the bucket size (8 / 25 / 120 MB), gradient dtype (fp32 / bf16 hook) or NCCL CTA budget -- it is
    the reducer's per-parameter hooks and bucket bookkeeping on a launch-bound backward (186 parameters), not wire time:
    115 MB over NVLink 5 is ~0.2 ms.  Here: one multi-tensor copy of the gradients into the flat buffer (35 us), one
    all-reduce (AVG), and `p.grad` re-pointed at views of the buffer for the optimiser.  Same arithmetic as DDP's
    (sum, then / world size in fp32).  Like DDP's constructor, the parameters and buffers of rank 0 are broadcast once."""

    def __init__(self, module, buckets=1):
        import torch.distributed as dist
        self.dist = dist
        self.world = dist.get_world_size()
        self.params = [p for p in module.parameters() if p.requires_grad]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=torch.float32)
        self._layout()
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, 0)
        self.avg = dist.get_backend() == 'nccl'
        # Overlap (buckets > 1).  The backward produces the gradients in reverse execution order, and 56 % of
        # ConvNeXt-T's weights sit in the head and the last stage, whose backward is over after a tenth of the backward's
        # time.  The first step records the order in which the gradients arrive (post-accumulate hooks) and lays the flat
        # buffer out in that order, cut into `buckets` contiguous ranges of about equal bytes; from then on, when the last
        # gradient of a range has arrived the range is copied in and all-reduced on a side stream while the backward
        # goes on, and only the last range (the stem and the first stages: a few MB) is exposed.  Inside the whole-step
        # CUDA graph the hooks run at capture only (fork / join of the side stream become graph edges); eager, they
        # cost host time per parameter.  Opt-in (B200AT_FLAT_BUCKETS): measured, it does not beat the single all-reduce.
        self.nbuckets = buckets if (buckets > 1 and len(self.params) >= buckets) else 1
        self.buckets, self.side = [], None
        self.learning = self.nbuckets > 1
        self.order = []
        if self.nbuckets > 1:
            self.side = torch.cuda.Stream(device=dev) if dev.type == 'cuda' else None     # CPU (gloo tests): in line
            self.index_of = {id(p): i for i, p in enumerate(self.params)}
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)
        self.arrived, self.done, self.forked = [], [], False

    def _layout(self):
        self.views, self.offsets, off = [], [], 0
        for p in self.params:
            # same strides as the parameter (channels_last conv weights included): the fused optimiser wants parameter and
            # gradient in one layout; every parameter is dense, so its strides address exactly numel() elements
            self.views.append(self.flat[off:off + p.numel()].as_strided(p.shape, p.stride()))
            self.offsets.append(off)
            off += p.numel()

    def _hook(self, p):
        i = self.index_of[id(p)]
        if self.learning:
            self.order.append(i)
            return
        b = self.bucket_of[i]
        self.arrived[b] += 1
        lo, hi = self.buckets[b]
        if self.arrived[b] == hi - lo and not self.done[b]:
            self._reduce_range(lo, hi, overlap=True)
            self.done[b] = True

    def _finish_learning(self):
        """flat buffer in arrival order (gradients that never arrived: at the end), ranges of about equal bytes"""
        seen = set()
        order = [i for i in self.order if not (i in seen or seen.add(i))]
        order += [i for i in range(len(self.params)) if i not in seen]
        self.params = [self.params[i] for i in order]
        self.index_of = {id(p): i for i, p in enumerate(self.params)}
        self._layout()
        total, acc, start = self.flat.numel(), 0, 0
        for i, p in enumerate(self.params):
            acc += p.numel()
            if acc >= total * (len(self.buckets) + 1) / self.nbuckets or i == len(self.params) - 1:
                self.buckets.append((start, i + 1))
                start = i + 1
        self.bucket_of = {i: b for b, (lo, hi) in enumerate(self.buckets) for i in range(lo, hi)}
        self.arrived = [0] * len(self.buckets)
        self.done = [False] * len(self.buckets)
        self.learning = False

    @torch.no_grad()
    def _reduce_range(self, lo, hi, overlap):
        src, dst = [], []
        for p, v in zip(self.params[lo:hi], self.views[lo:hi]):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad if p.grad.dtype == torch.float32 else p.grad.float())
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)
        end = self.offsets[hi] if hi < len(self.params) else self.flat.numel()
        seg = self.flat[self.offsets[lo]:end]
        if overlap and self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())
            self.forked = True
            with torch.cuda.stream(self.side):
                self._all_reduce(seg)
        else:
            self._all_reduce(seg)

    def _all_reduce(self, t):
        if self.avg:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.AVG)
        else:                                                   # gloo (CPU tests): no AVG
            self.dist.all_reduce(t)
            t.div_(self.world)

    @torch.no_grad()
    def reduce(self):
        if self.learning:
            self._finish_learning()                             # this step: everything after the backward, new layout
            self._reduce_range(0, len(self.params), overlap=False)
        elif self.buckets:
            for b, (lo, hi) in enumerate(self.buckets):         # ranges whose last gradient never arrived
                if not self.done[b]:
                    self._reduce_range(lo, hi, overlap=False)
            if self.forked:
                torch.cuda.current_stream().wait_stream(self.side)
            self.arrived = [0] * len(self.buckets)
            self.done = [False] * len(self.buckets)
            self.forked = False
        else:
            self._reduce_range(0, len(self.params), overlap=False)
        for p, v in zip(self.params, self.views):
            p.grad = v


class AdvTrainStep:
    """One adversarial training step on one rank (main.py:961-997)."""

    def __init__(self, base_model, attack='apgd', norm='Linf', eps=4. / 255., n_iter=2, lr=1e-3, weight_decay=0.05,
                 label_smoothing=0., ema=False, distributed=False, device=None, autocast_dtype=torch.bfloat16,
                 channels_last=True, mixup_fn=None, perturb=None, graph_attack=False, param_groups=None,
                 optimizer='adamw', momentum=0.9, graph_step=False):
        self.device = device
        # `perturb` overrides the attack callable (same (model, x, y) contract as main.py:283)
        perturb = perturb if perturb is not None else make_attack(attack, norm, eps, n_iter, mixup_fn=mixup_fn)
        self.eager_perturb = perturb
        if graph_attack and perturb is not None and attack == 'apgd':      # fgsm draws fresh noise per call: stays eager
            perturb = GraphedAttack(perturb)
        self.graphed_perturb = perturb
        if channels_last:
            base_model = base_model.to(memory_format=torch.channels_last)      # misc.use_channel_last (main.py:815-817)
        model = WrappedModel(base_model, perturb) if perturb is not None else base_model
        model = model.to(device)
        self.raw = model
        self.ema = DeviceEma(model) if ema else None                           # created before the DDP wrap (main.py:884)
        self.flat_reduce = None
        import os
        if distributed and os.environ.get('B200AT_DDP', 'flat') == 'flat':
            # B200AT_FLAT_BUCKETS > 1: ranges of the flat buffer reduced while the backward runs.  Measured with the whole-step
            # graph (profiles/r02_flat_buckets.txt): 30.33 / 30.35 / 30.33 / 30.24 ms at 2 GPUs for 1 / 2 / 3 / 4 ranges,
            # 30.56-30.67 vs 30.70-30.72 ms at 8 GPUs for 1 vs 3 -- no gain (the all-reduce of 115 MB is ~0.3 ms of a
            # 30 ms step and NCCL's CTAs compete with the persistent kernels while it overlaps), so the default stays 1.
            nb = int(os.environ.get('B200AT_FLAT_BUCKETS', '1'))
            self.flat_reduce = FlatGradAllReduce(model, buckets=nb)
        elif distributed:
            ids = [device.index] if (device is not None and device.type == 'cuda') else None
            # main.py:890.  broadcast_buffers=False: the only buffers are the normaliser's constant mean/std, and the
            # per-forward re-broadcast would bump their version counters, i.e. force a device->host read of the 3+3
            # constants (a host synchronisation) in front of every fused first-stem-stage launch.
            # B200AT_DDP_BUCKET_MB / B200AT_DDP_BF16: measurement knobs for the all-reduce tail (profiles/r02_ddp_n2_*.txt);
            # defaults are torch's (25 MB buckets, fp32 gradients: main.py:890 uses DDP as is).  B200AT_DDP=torch selects it.
            bucket_mb = float(os.environ.get('B200AT_DDP_BUCKET_MB', '25'))
            model = nn.parallel.DistributedDataParallel(model, device_ids=ids, broadcast_buffers=False,
                                                        gradient_as_bucket_view=True, bucket_cap_mb=bucket_mb)
            if os.environ.get('B200AT_DDP_BF16') == '1':
                from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
                model.register_comm_hook(None, default_hooks.bf16_compress_hook)
        self.model = model
        self.perturb = perturb is not None
        if param_groups is not None:                                           # the driver's per-arch rule (main.py:395-452)
            groups = [g for g in param_groups(self.raw.named_parameters()) if g['params']]
        else:
            decay, no_decay = [], []
            for n, p in self.raw.named_parameters():
                (no_decay if p.ndim <= 1 else decay).append(p)
            groups = [{'params': decay, 'weight_decay': weight_decay}, {'params': no_decay, 'weight_decay': 0.}]
        on_gpu = device is not None and device.type == 'cuda'
        if optimizer == 'sgd':                                                 # main.py:454-457
            self.optimizer = torch.optim.SGD(groups, lr=lr, momentum=momentum)
        else:
            # graph_step: the optimiser is recorded in the step's CUDA graph -> capturable (device-side step counter and lr)
            cap = bool(graph_step and on_gpu)
            self.optimizer = torch.optim.AdamW(groups, lr=torch.tensor(float(lr), device=device) if cap else lr,
                                               betas=(0.9, 0.95), fused=on_gpu, capturable=cap)
        self.label_smoothing = label_smoothing
        self.mixup_fn = mixup_fn
        self.autocast_dtype = autocast_dtype
        self.graphed = GraphedStep(self) if (graph_step and on_gpu and optimizer != 'sgd' and
                                            (not distributed or self.flat_reduce is not None)) else None

    def set_lr(self, lr):
        """learning rate of every group (main.py:957-959 sets it per iteration); a device tensor under `graph_step`"""
        for g in self.optimizer.param_groups:
            if torch.is_tensor(g['lr']):
                g['lr'].fill_(float(lr))
            else:
                g['lr'] = lr

    def use_graph(self, on):
        """switch the attack between its CUDA-graph replay and the eager launch sequence (same kernels)"""
        if self.perturb:
            self.raw.perturb = self.graphed_perturb if on else self.eager_perturb

    def loss(self, output, target):
        if target.dim() == 2:                                                   # timm SoftTargetCrossEntropy (main.py:461-466)
            return torch.sum(-target * F.log_softmax(output.float(), dim=-1), dim=-1).mean()
        return F.cross_entropy(output.float(), target, label_smoothing=self.label_smoothing)

    def __call__(self, images, target):
        if self.graphed is not None:
            return self.graphed(images, target)
        return self.eager_step(images, target)

    def eager_step(self, images, target):
        self.model.train()
        if self.perturb:
            self.raw.set_perturb(True)
        if self.mixup_fn is not None:
            images, target = self.mixup_fn(images, target)
        self.optimizer.zero_grad(set_to_none=True)
        with torch.autocast(device_type=images.device.type, dtype=self.autocast_dtype):
            output = self.model(images, target) if self.perturb else self.model(images)
            loss = self.loss(output, target)
        loss.backward()                                                         # (B200AT_DDP=torch: all-reduce overlaps here)
        if self.flat_reduce is not None:
            self.flat_reduce.reduce()                                           # the step's only collective
        self.optimizer.step()
        ops.invalidate_derived()          # fused optimisers do not move the parameters' version counters (see ops.py)
        if self.ema is not None:
            self.ema.update()
        return loss.detach()
