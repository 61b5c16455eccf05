#!/usr/bin/env python
"""bench.py -- APGD adversarial-training throughput on B200 (BASELINE.json metric / configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one adversarial training step of ConvNeXt-T-CvSt on one synthetic batch of 128 images per
GPU (3x224x224, l-inf 4/255, APGD n_iter=2, bf16 autocast): 4 forwards + 2 input-grad backwards inside
`apgd_train`, one full backward with DDP's NCCL all-reduce, fused AdamW.  Rank 0 prints ONE JSON line.

  value     images/s over all ranks, inputs resident in HBM, K steps between CUDA events, max over ranks
  e2e       same, but every step starts from pinned HOST buffers (H2D inside the timed region) and ends
            with a D2H read of the loss
  roofline  the fused l-inf APGD update kernel (b200at_linf_step): 20 B/element x B x n_fts per launch
             / mean launch duration, measured live with CUDA events on the launching
            stream over the timed region; peak = MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own apgd_train + model files (staged under oracle/_ref by build(); the oracle
            port if absent) in the same step, fp32, on this box's host cores, bounded sample (steps of 32
            images = BASELINE configs[0]), rank 0 at N=1 only
  --impl reference   times that CPU implementation instead (rank 0 only)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = 'apgd_adv_train_images_per_sec'
UNIT = 'images/s'
BATCH_PER_GPU = 128
RES = 224
EPS = 4. / 255.
N_ITER = 2
ARCH = 'convnext_tiny'
N_FTS = 3 * RES * RES
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback, used only if MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--arch', default=ARCH, choices=sorted(WORKLOADS),
                    help='convnext_tiny = BASELINE configs[1] (the metric line); the others are secondary workloads')
    ap.add_argument('--batch', type=int, default=None, help='images per GPU (default: the BASELINE config of --arch)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ema', type=int, default=None, help='on-device EMA of the parameters (default: on for convnext_base = config 4)')
    ap.add_argument('--label-smoothing', type=float, default=None, help='default 0.1 for convnext_base (config 4), else 0')
    ap.add_argument('--no-graph', action='store_true', help='launch the attack kernel by kernel instead of replaying its CUDA graph')
    ap.add_argument('--graph-step', type=int, default=int(os.environ.get('B200AT_GRAPH_STEP', '1')), help='1: replay the WHOLE step (attack + training forward/backward + all-reduce + AdamW) from one CUDA graph')
    ap.add_argument('--cpu-seconds', type=float, default=20.0, help='budget of the cpu_baseline sample')
    ap.add_argument('--res', type=int, default=RES, help='image side (224 = the metric line; 320 = the secondary resolution of north_star)')
    return ap.parse_args()


# arch -> (display name, default batch per GPU, which BASELINE.json config it is)
WORKLOADS = {
    'convnext_tiny': ('ConvNeXt-T-CvSt', 128, 'BASELINE.json configs[1]'),
    'vit_small': ('ViT-S-CvSt', 256, 'BASELINE.json configs[2], single-GPU share'),
    'convnext_base': ('ConvNeXt-B-CvSt', 64, 'BASELINE.json configs[3] model, single-GPU share'),
}


def build_engine(arch):
    if arch == 'vit_small':
        from revisiting_at_b200 import vit
        return vit.build(normalize=True, seed=0)
    from revisiting_at_b200 import convnext
    return convnext.build(arch, normalize=True, seed=0)


def build_oracle(arch):
    if arch == 'vit_small':
        from oracle import vit_oracle
        return vit_oracle.build(normalize=True, seed=0)
    from oracle import convnext_oracle
    return convnext_oracle.build(arch, normalize=True, seed=0)


def workload_config(n_gpus, batch, arch=ARCH, ema=False, label_smoothing=0.):
    name, _, which = WORKLOADS[arch]
    return {'workload': f'{name} APGD l-inf 4/255 n_iter={N_ITER} adversarial train step, bf16 autocast, '
                        f'batch {batch}/GPU, 3x{RES}x{RES} ({which})',
            'arch': arch, 'batch_per_gpu': batch, 'global_batch': batch * n_gpus, 'resolution': RES,
            'norm': 'Linf', 'eps': '4/255', 'n_iter': N_ITER, 'parallelism': f'dp{n_gpus}',
            'ema': bool(ema), 'label_smoothing': label_smoothing,
            'l2_policy': f'inputs_exceed_l2 (per-step working set >> 126 MB; image-sized passes stream {20 * batch * N_FTS / 1e6:.0f} MB)'}


def synth_batch(batch, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, RES, RES, generator=g)
    y = torch.randint(0, 1000, (batch,), generator=g)
    return x, y


ENGINE_NOTE = {
    False: 'hand-written NHWC bf16 kernels: fused first stem stage (attack evaluations and training forward; tensor-core input '
           'gradient), second stem convolution forward as an implicit GEMM on the tcgen05 kernel, dwconv7 fwd/dgrad (tensor-core '
           'Toeplitz form on the wide maps) and wgrad, LayerNorm (+patch layout for the downsample), fused tcgen05 MLP kernel per '
           'direction (pwconv1 -> GELU -> pwconv2+scale+bias+residual, hidden on chip) for C <= 192, tcgen05 GEMM with fused '
           'bias+GELU / GELU-grad / residual epilogues for the wider stages and the downsample conv2x2s2; library calls: cuBLAS '
           'weight-gradient GEMMs, cuDNN input / weight gradients of the 3x3 stem convolutions',
    True: 'hand-written kernels: fused first stem stage, LayerNorm(+GELU), mma.sync attention fwd/bwd, tcgen05 GEMM (qkv+bias, '
          'proj/fc2+bias+residual, fc1 with bias+GELU, GELU-grad GEMM, all input-gradient GEMMs); cuBLAS for weight-gradient '
          'GEMMs; second stem convolution forward on the tcgen05 kernel, cuDNN for its gradients and for stem convs 3-4',
}


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arm
CPU_BATCH = 32            # BASELINE.json configs[0]: the reference's own CPU-runnable case


def cpu_step_factory(arch=ARCH):
    """(step, kind): the UNMODIFIED reference `apgd_train` + its own ConvNeXt-T-CvSt modules when staged
    (oracle/_ref, kind "reference"); the oracle port otherwise (other archs: timm is un-vendored) -- kind "port"."""
    from oracle import train_step_oracle
    torch.set_num_threads(os.cpu_count())
    if arch == ARCH and RES % 32 == 0:
        step = train_step_oracle.reference_train_step('Linf', EPS, N_ITER)
        if step is not None:
            return step, 'reference'
    return train_step_oracle.OracleTrainStep(build_oracle(arch), 'Linf', EPS, N_ITER), 'port'


_KIND_NOTE = {'reference': 'unmodified reference apgd_train + models/convnext.py + ConvBlock1 (oracle/_ref), train_loop body restated',
              'port': 'oracle/train_step_oracle.py port'}


def _cpu_batch(step, budget_s, n_steps):
    """CPU_BATCH images per step unless n_steps of them would not fit the budget (probe: one step of 4 images)."""
    x, y = synth_batch(4, 7)
    t0 = time.time(); step(x, y); t4 = time.time() - t0
    per_img = t4 / 4
    fit = int(budget_s / max(n_steps, 1) / max(per_img, 1e-4))
    return max(2, min(CPU_BATCH, fit))


def cpu_baseline(seconds, arch=ARCH):
    """The reference's CPU implementation of the step on the host cores, bounded sample: one probe step of 4 images
    (doubles as warm-up), then steps of CPU_BATCH images for ~`seconds`, timed with the wall clock."""
    step, kind = cpu_step_factory(arch)
    batch = _cpu_batch(step, seconds, 1)
    x, y = synth_batch(batch, 8)
    n, t0 = 0, time.time()
    while True:
        step(x, y); n += batch
        if time.time() - t0 > seconds * 0.6 or n >= 3 * CPU_BATCH:
            break
    dt = time.time() - t0
    return {'value': n / dt, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': kind,
            'sample': f'{n} images in steps of {batch} (fp32, {_KIND_NOTE[kind]}, same model/attack config), '
                      f'{dt:.1f} s wall after one warm-up step'}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (kind "reference": its unmodified
    apgd_train and model files staged under oracle/_ref at build time; "port" only if they are absent), all host
    threads, rank 0 only, every step a bounded sample (CPU_BATCH images) of the 128/GPU workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    step, kind = cpu_step_factory(args.arch)
    total = args.steps + args.warmup
    batch = _cpu_batch(step, 200.0, total)                   # whole run ends within a few minutes
    x, y = synth_batch(batch, 8)
    for _ in range(args.warmup):
        step(x, y)
    t0 = time.time()
    for _ in range(args.steps):
        step(x, y)
    dt = time.time() - t0
    v = batch * args.steps / dt
    cores = torch.get_num_threads()
    sample = (f'{args.steps} steps of {batch} images (bounded sample of the {args.batch}/GPU step), fp32, {cores} threads, '
              f'{_KIND_NOTE[kind]}')
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args.gpus, args.batch, args.arch, args.ema, args.label_smoothing),
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference)')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    distributed = world > 1
    if distributed:
        dist.init_process_group('nccl', device_id=dev)
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import _abi
    from revisiting_at_b200.train_step import AdvTrainStep
    _abi.lib()                                              # fail loudly if the CUDA library is missing
    torch.backends.cudnn.benchmark = True                   # main.py:25

    batch = args.batch
    model = build_engine(args.arch)
    step = AdvTrainStep(model, 'apgd', 'Linf', EPS, N_ITER, distributed=distributed, device=dev,
                        graph_attack=not args.no_graph, ema=bool(args.ema), label_smoothing=args.label_smoothing,
                        graph_step=bool(args.graph_step) and not args.no_graph)

    pool = 2
    host = [synth_batch(batch, 1234 + 17 * rank + i) for i in range(pool)]
    host = [(x.pin_memory(), y.pin_memory()) for x, y in host]
    resident = [(x.to(dev), y.to(dev)) for x, y in host]

    def sync_all():
        torch.cuda.synchronize(dev)
        if distributed:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def resident_step(i):
        x, y = resident[i % pool]
        step(x, y)

    sink = torch.zeros(1, pin_memory=True)

    from revisiting_at_b200.train_step import DevicePrefetcher
    prefetch = DevicePrefetcher(dev)

    def e2e_step(i):
        # every step's images + labels come from pinned HOST memory inside the timed region; the copy of step i+1
        # is issued on a copy stream before step i's kernels (one batch ahead, like a pin_memory DataLoader)
        if not prefetch.queue:
            prefetch.submit(*host[i % pool])
        x, y = prefetch.get()
        prefetch.submit(*host[(i + 1) % pool])
        loss = step(x, y)
        sink.copy_(loss.reshape(1), non_blocking=False)     # D2H read of the step's result

    n_warm = max(args.warmup, 3) + (0 if args.no_graph else 2) + (1 if args.graph_step else 0)   # graph mode: 2 eager calls, 1 capture, >= 2 replays
    for i in range(n_warm):
        resident_step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _abi.LAUNCHES['count']
    ms = timed(resident_step, args.steps)
    launches = _abi.LAUNCHES['count'] - launches0
    for i in range(2):
        e2e_step(i)
    prefetch.queue.clear()                                  # the timed region starts with nothing on the device
    ms_e2e = timed(e2e_step, args.steps)
    prefetch.queue.clear()
    # roofline of the fused l-inf update: the same K steps once more with the attack launched kernel by kernel
    # (a kernel inside a replayed CUDA graph cannot carry events), CUDA events on the launching stream around
    # every K1 launch.
    step.graphed = None                                     # (whole-step graph off for this part as well)
    step.use_graph(False)
    resident_step(0)
    # algorithmic bytes per element of each timed K1 launch (20 B = read x, x_adv, x_adv_old, grad +
    # write x_adv; the first move of a call has x_adv_old == x_adv, one stream fewer => credited 16 B)
    credit = {'linf_step': 20.0, 'linf_step_log': 20.0, 'linf_step_log_first': 16.0}
    _abi.TIMING['names'] = set(credit)                      # events only around the K1 launches (2 per attack call)
    _abi.TIMING['enabled'] = True
    _abi.TIMING['events'].clear()
    profile_range = os.environ.get('B200AT_PROFILE_RANGE') == '1'      # ncu --profile-from-start off
    if profile_range:
        torch.cuda.profiler.start()
    ms_eager = timed(resident_step, args.steps)
    if profile_range:
        torch.cuda.profiler.stop()
    _abi.TIMING['enabled'] = False
    k1 = [(a.elapsed_time(b), credit[name]) for name, a, b in _abi.TIMING['events'] if name in credit]
    _abi.TIMING['events'].clear()
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return
    imgs = batch * world * args.steps
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))['hbm_gbs'], 'MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)'
    else:
        peak, peak_src = FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'
    k1_ms = sum(t for t, _ in k1) / len(k1) if k1 else float('nan')
    alg_bytes = sum(c for _, c in k1) / len(k1) * batch * N_FTS if k1 else float('nan')   # mean per launch
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'k1_linf_step_traffic.json')
    if os.path.exists(tp) and RES == 224 and batch == 128 and args.arch == ARCH:   # the capture is of this shape
        traffic = json.load(open(tp)).get('dram_bytes_per_launch')
    line = {
        'metric': METRIC, 'value': imgs / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': n_warm, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic', 'config': workload_config(world, batch, args.arch, args.ema, args.label_smoothing),
        'e2e': {'value': imgs / (ms_e2e * 1e-3), 'unit': UNIT,
                'h2d_bytes_per_step': batch * N_FTS * 4 + batch * 8, 'd2h_bytes_per_step': 4,
                'ms_per_step': ms_e2e / args.steps,
                'input_path': 'pinned host batch -> device on a copy stream, one batch ahead (DevicePrefetcher); '
                              'K+1 copies for K steps inside the timed region'},
        'gpu_launches': launches,
        'roofline': {'kernel': 'b200at_linf_step_log (fused l-inf APGD update, iterate-log form)', 'bound': 'hbm',
                     'achieved': achieved,
                     'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                     'algorithmic_bytes_per_launch': alg_bytes, 'launch_ms_mean': k1_ms, 'launches_timed': len(k1),
                     'peak_source': peak_src, 'frac_of_nominal_8TBps': achieved / 8000.0,
                     'timed_region': f'{args.steps} more steps with the attack launched kernel by kernel '
                                     f'({ms_eager / args.steps:.2f} ms/step)'},
        'attack_launch': 'eager' if args.no_graph else ('cuda_graph of the WHOLE step (attack + training forward/backward + all-reduce + AdamW)' if args.graph_step else 'cuda_graph (one graph per apgd_train call: 3 forwards + 2 input-grad backwards + update/bookkeeping kernels)'),
        'clocks': clocks,
        'model_engine': ENGINE_NOTE[args.arch == 'vit_small'],
    }
    if world == 1 and not args.no_cpu_baseline:
        torch.cuda.empty_cache()
        line['cpu_baseline'] = cpu_baseline(args.cpu_seconds, args.arch)
    print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def main():
    global RES, N_FTS
    args = parse()
    RES, N_FTS = args.res, 3 * args.res * args.res
    if args.batch is None:
        args.batch = WORKLOADS[args.arch][1]
    if args.ema is None:
        args.ema = int(args.arch == 'convnext_base')
    if args.label_smoothing is None:
        args.label_smoothing = 0.1 if args.arch == 'convnext_base' else 0.
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
