#!/usr/bin/env python
"""Adversarial-training driver with the reference's command line.

    python main.py --data.num_workers=12 --data.in_memory=1 --data.train_dataset=<dir|synthetic[:N]> \
        --data.val_dataset=<dir|synthetic> --logging.folder=<dir> --adv.attack apgd --adv.n_iter 2 --adv.norm Linf \
        --training.distributed 1 --dist.world_size 8 --model.arch convnext_tiny --model.not_original 1 ...

Same sections, keys, defaults and `--section.key value` syntax as /root/reference/main.py:106-189 (run_train.sh:10-18
drives it unchanged), same wiring of `adv.*` into `functools.partial(apgd_train | fgsm_train)` (main.py:831-842),
same `WrappedModel` (main.py:260-301), per-iteration learning-rate interpolation (main.py:957-959,975-976),
parameter groups (main.py:395-452), checkpoint files (main.py:737-756) and one process per GPU (main.py:1131-1135).
The adversarial train step itself is `revisiting_at_b200.train_step.AdvTrainStep`, i.e. the sm_100a kernels.

Deliberately different: bf16 autocast without GradScaler (`training.precision` is accepted and ignored),
EMA on the device, models restricted to the CvSt families the engine builds (ConvNeXt-T/S/B/L, ViT-S with
`model.not_original 1`), `logging.save_freq` is honoured (the reference overwrites it with 1 before saving: main.py:733).  Outside the hot path and therefore minimal: the data pipeline (`synthetic[:N]` batches,
or a torchvision ImageFolder with random-resized-crop + flip; timm's RandAugment / RandomErasing are not rebuilt)
and validation (clean top-1 on the first batches, main.py:905-942).
"""
import json
import math
import os
import sys
import time
from datetime import datetime
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

try:                                                     # the pip package where it exists, else the in-tree subset
    from fastargs import Param, Section, get_current_config
    from fastargs.decorators import param
    from fastargs.validation import And, OneOf
    _make_config = None
except ImportError:
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200.fastargs_compat import And, OneOf, Param, Section, get_current_config, param
    from revisiting_at_b200.fastargs_compat import make_config as _make_config

# ---------------------------------------------------------------------------- configuration (main.py:106-189)
Section('model', 'model details').params(
    arch=Param(str, default='effnet_b0'),
    pretrained=Param(int, 'is pretrained? (1/0)', default=1),
    ckpt_path=Param(str, 'path to resume model', default=''),
    add_normalization=Param(int, '0 if no normalization, 1 otherwise', default=1),
    not_original=Param(int, 'conv stem (CvSt) instead of the patch stem', default=0),
    updated=Param(int, 'bigger conv stem', default=0),
    model_ema=Param(float, 'Use EMA?', default=0),
    freeze_some=Param(int, 'freeze some layers', default=0),
    early=Param(int, 'freeze early layers?', default=1))
Section('resolution', 'resolution scheduling').params(
    min_res=Param(int, 'the minimum (starting) resolution', default=160),
    max_res=Param(int, 'the maximum (starting) resolution', default=160),
    end_ramp=Param(int, 'when to stop interpolating resolution', default=0),
    start_ramp=Param(int, 'when to start interpolating resolution', default=0))
Section('data', 'data related stuff').params(
    train_dataset=Param(str, 'training set: directory or synthetic[:images per epoch and rank]', required=True),
    val_dataset=Param(str, 'validation set: directory or synthetic', required=True),
    num_workers=Param(int, 'The number of workers', required=True),
    in_memory=Param(int, 'does the dataset fit in memory? (1/0)', required=True),
    seed=Param(int, 'seed for training loader', default=0),
    augmentations=Param(int, 'Mixup/CutMix soft targets (+ flip)?', default=0))
Section('lr', 'lr scheduling').params(
    step_ratio=Param(float, 'learning rate step ratio', default=0.1),
    step_length=Param(int, 'learning rate step length', default=30),
    lr_schedule_type=Param(OneOf(['step', 'cyclic', 'cosine']), default='cosine'),
    lr=Param(float, 'learning rate', default=1e-3),
    lr_peak_epoch=Param(int, 'Epoch at which LR peaks', default=10))
Section('logging', 'how to log stuff').params(
    folder=Param(str, 'log location', default='./runs'),
    log_level=Param(int, '0 if only at end 1 otherwise', default=1),
    save_freq=Param(int, 'save models every nth epoch', default=2),
    addendum=Param(str, 'additional comments?', default=''))
Section('validation', 'Validation parameters stuff').params(
    batch_size=Param(int, 'The batch size for validation', default=64),
    resolution=Param(int, 'final resized validation image size', default=224),
    lr_tta=Param(int, 'should do lr flipping/avging at test time', default=0),
    precision=Param(str, 'np precision', default='fp16'))
Section('training', 'training hyper param stuff').params(
    eval_only=Param(int, 'eval only?', default=0),
    batch_size=Param(int, 'The batch size', default=512),
    optimizer=Param(And(str, OneOf(['sgd', 'adamw'])), 'The optimizer', default='adamw'),
    momentum=Param(float, 'SGD momentum', default=0.9),
    weight_decay=Param(float, 'weight decay', default=0.05),
    epochs=Param(int, 'number of epochs', default=100),
    label_smoothing=Param(float, 'label smoothing parameter', default=0.1),
    distributed=Param(int, 'is distributed?', default=0),
    use_blurpool=Param(int, 'use blurpool?', default=0),
    precision=Param(str, 'np precision', default='fp16'))
Section('dist', 'distributed training options').params(
    world_size=Param(int, 'number gpus', default=1),
    address=Param(str, 'address', default='localhost'),
    port=Param(str, 'port', default='12355'))
Section('adv', 'adversarial training options').params(
    attack=Param(str, 'if None standard training', default='none'),
    norm=Param(str, '', default='Linf'),
    eps=Param(float, '', default=4. / 255.),
    n_iter=Param(int, '', default=2),
    verbose=Param(int, '', default=0),
    noise_level=Param(float, '', default=1.),
    skip_projection=Param(int, '', default=0),
    alpha=Param(float, 'step size multiplier', default=1.))
Section('misc', 'other parameters').params(
    notes=Param(str, '', default=''),
    use_channel_last=Param(int, 'whether to use channel last memory format', default=1))


# ---------------------------------------------------------------------------- schedules (main.py:208-243)
@param('lr.lr')
@param('lr.step_ratio')
@param('lr.step_length')
@param('training.epochs')
def get_step_lr(epoch, lr, step_ratio, step_length, epochs):
    return 0 if epoch >= epochs else lr * step_ratio ** (epoch // step_length)


@param('lr.lr')
@param('training.epochs')
@param('lr.lr_peak_epoch')
def get_cyclic_lr(epoch, lr, epochs, lr_peak_epoch):
    return np.interp([epoch], [0, lr_peak_epoch, epochs], [1e-4 * lr, lr, 0])[0]


@param('lr.lr')
@param('training.epochs')
@param('lr.lr_peak_epoch')
def get_cosine_lr(epoch, lr, epochs, lr_peak_epoch):
    if epoch <= lr_peak_epoch:                                         # linear warm-up from 1e-4 lr
        return np.interp([epoch], [0, lr_peak_epoch], [1e-4 * lr, lr])[0]
    floor = 5e-6
    phase = (epoch - lr_peak_epoch) / (epochs - lr_peak_epoch)
    return floor + .5 * (lr - floor) * (1 + math.cos(math.pi * phase))


LR_SCHEDULES = {'cyclic': get_cyclic_lr, 'step': get_step_lr, 'cosine': get_cosine_lr}


def weight_decay_groups(named_parameters, arch, weight_decay):
    """main.py:395-452: for convnext / resnet names the no-decay set is "name contains 'bn' or '.bias'" (so LayerNorm
    weights and the layer scale ARE decayed there); for every other arch it is `ndim <= 1 or name ends with .bias`."""
    named = [(k, v) for k, v in named_parameters]
    if 'convnext' in arch or 'resnet' in arch:
        skip = lambda k, v: 'bn' in k or '.bias' in k
        named_ = named
    else:
        skip = lambda k, v: v.ndim <= 1 or k.endswith('.bias')
        named_ = [(k, v) for k, v in named if v.requires_grad]
    return [{'params': [v for k, v in named_ if skip(k, v)], 'weight_decay': 0.},
            {'params': [v for k, v in named_ if not skip(k, v)], 'weight_decay': weight_decay}]


# ---------------------------------------------------------------------------- data (outside the hot path: minimal)
class SyntheticLoader:
    """`images per epoch and rank` random images in [0,1] with random labels, generated once on the host (pinned),
    served in `batch_size` slices; `synthetic:N` in `data.train_dataset` sets N (default 8 batches)."""

    def __init__(self, spec, batch_size, res, seed, n_cls=1000):
        n = int(spec.split(':')[1]) if ':' in spec else 8 * batch_size
        self.n_batches = max(n // batch_size, 1)
        g = torch.Generator().manual_seed(seed)
        pool = min(self.n_batches, 2)
        self.batches = [(torch.rand(batch_size, 3, res, res, generator=g), torch.randint(0, n_cls, (batch_size,), generator=g))
                        for _ in range(pool)]
        if torch.cuda.is_available():
            self.batches = [(x.pin_memory(), y.pin_memory()) for x, y in self.batches]

    def __len__(self):
        return self.n_batches

    def __iter__(self):
        for i in range(self.n_batches):
            yield self.batches[i % len(self.batches)]


def folder_loader(path, batch_size, res, num_workers, world_size, rank, seed, train, flip):
    from torchvision import datasets, transforms
    if train:
        tf = [transforms.RandomResizedCrop(res, scale=(0.08, 1.0), ratio=(3. / 4., 4. / 3.),
                                           interpolation=transforms.InterpolationMode.BICUBIC)]
        tf += [transforms.RandomHorizontalFlip()] if flip else []
    else:
        tf = [transforms.Resize(int(res / 0.875), interpolation=transforms.InterpolationMode.BICUBIC), transforms.CenterCrop(res)]
    ds = datasets.ImageFolder(path, transforms.Compose(tf + [transforms.ToTensor()]))
    sampler = torch.utils.data.DistributedSampler(ds, num_replicas=world_size, rank=rank, shuffle=train, seed=seed)
    return torch.utils.data.DataLoader(ds, sampler=sampler, batch_size=batch_size, num_workers=num_workers,
                                       pin_memory=True, drop_last=train)


# ---------------------------------------------------------------------------- trainer (main.py:328-1152)
class ImageNetTrainer:
    @param('training.distributed')
    @param('training.eval_only')
    def __init__(self, gpu, distributed, eval_only):
        self.all_params = get_current_config()
        self.gpu = gpu
        self.device = torch.device('cuda', gpu)
        if not torch.cuda.is_available():
            raise SystemExit('main.py: no CUDA device (the adversarial train step has no CPU path)')
        torch.cuda.set_device(self.device)
        torch.backends.cudnn.benchmark = True                          # main.py:25
        if distributed:
            self.setup_distributed()
        self.train_loader, self.val_loader, self.mixup_fn = self.create_train_loader()
        self.step = self.create_model_and_scaler()
        self.model, self.optimizer = self.step.model, self.step.optimizer
        self.initialize_logger()

    @param('dist.address')
    @param('dist.port')
    @param('dist.world_size')
    def setup_distributed(self, address, port, world_size):
        os.environ.setdefault('MASTER_ADDR', address)
        os.environ.setdefault('MASTER_PORT', port)
        dist.init_process_group('nccl', rank=self.gpu, world_size=world_size, device_id=self.device)

    def cleanup_distributed(self):
        dist.destroy_process_group()

    @param('lr.lr_schedule_type')
    def get_lr(self, epoch, lr_schedule_type):
        return LR_SCHEDULES[lr_schedule_type](epoch)

    @param('resolution.min_res')
    @param('resolution.max_res')
    @param('resolution.end_ramp')
    @param('resolution.start_ramp')
    def get_resolution(self, epoch, min_res, max_res, end_ramp, start_ramp):
        assert min_res <= max_res
        if epoch <= start_ramp:
            return min_res
        if epoch >= end_ramp:
            return max_res
        interp = np.interp([epoch], [start_ramp, end_ramp], [min_res, max_res])
        return int(np.round(interp[0] / 32)) * 32                     # nearest multiple of 32

    @param('data.train_dataset')
    @param('data.val_dataset')
    @param('data.num_workers')
    @param('training.batch_size')
    @param('validation.batch_size', alias='val_batch_size')
    @param('training.distributed')
    @param('training.label_smoothing')
    @param('data.seed')
    @param('data.augmentations')
    @param('dist.world_size')
    def create_train_loader(self, train_dataset, val_dataset, num_workers, batch_size, val_batch_size, distributed,
                            label_smoothing, seed, augmentations, world_size):
        torch.manual_seed(seed)
        res = self.get_resolution(0)
        if self.get_resolution(10 ** 6) != res:
            # the reference retargets the decoder every epoch (main.py:716-720); this driver builds its loaders once
            raise SystemExit('progressive resizing (resolution.min_res != resolution.max_res) is not built: '
                             'pass equal --resolution.min_res / --resolution.max_res')
        world = world_size if distributed else 1
        if train_dataset.startswith('synthetic'):
            train = SyntheticLoader(train_dataset, batch_size, res, seed * 1000 + self.gpu)
        else:
            train = folder_loader(train_dataset, batch_size, res, num_workers, world, self.gpu, seed, True, bool(augmentations))
        if val_dataset.startswith('synthetic'):
            val = SyntheticLoader('synthetic:%d' % (2 * val_batch_size), val_batch_size, res, 77 + self.gpu)
        else:
            val = folder_loader(val_dataset, val_batch_size, res, num_workers, world, self.gpu, seed, False, False)
        mixup_fn = None
        if augmentations:                                              # parserr.py:27-32 through main.py:599-607
            from revisiting_at_b200.mixup import Mixup
            mixup_fn = Mixup(mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5, mode='batch',
                             label_smoothing=label_smoothing, num_classes=1000)
        return train, val, mixup_fn

    @param('model.arch')
    @param('model.pretrained')
    @param('model.not_original')
    @param('model.model_ema')
    @param('model.ckpt_path')
    @param('model.add_normalization')
    @param('training.distributed')
    @param('training.optimizer')
    @param('training.momentum')
    @param('training.weight_decay')
    @param('adv.attack')
    @param('adv.norm')
    @param('adv.eps')
    @param('adv.n_iter')
    @param('adv.verbose')
    @param('adv.alpha')
    @param('adv.noise_level')
    @param('adv.skip_projection')
    @param('misc.use_channel_last')
    def create_model_and_scaler(self, arch, pretrained, not_original, model_ema, ckpt_path, add_normalization, distributed,
                                optimizer, momentum, weight_decay, attack, norm, eps, n_iter, verbose, alpha, noise_level,
                                skip_projection, use_channel_last):
        from revisiting_at_b200 import checkpoint, convnext, vit
        from revisiting_at_b200.train_step import AdvTrainStep, make_attack
        arch = arch.replace('timm_', '')
        if pretrained:
            raise SystemExit('model.pretrained 1 needs timm weights from the network; pass --model.pretrained 0 '
                             '(and --model.ckpt_path for a checkpoint)')
        if not not_original:
            raise SystemExit('only the conv-stem (CvSt) models are built: pass --model.not_original 1')
        if arch in convnext.ARCHS:
            model = convnext.ConvNeXtCvSt(arch)
            model = convnext.Normalized(model) if add_normalization else model
        elif arch in ('vit_s', 'deit_s', 'vit_small'):
            model = vit.build(normalize=bool(add_normalization), seed=int(torch.initial_seed() % (2 ** 31)))
        else:
            raise SystemExit(f'model.arch {arch!r}: the engine builds {sorted(convnext.ARCHS)} and vit_s')
        if ckpt_path:
            # before the EMA shadow is cloned and before DDP wraps the model: the reference loads the checkpoint
            # (main.py:856-872) ahead of ModelEmaV2 (:881-887), so a resumed EMA starts from the loaded weights
            checkpoint.load_checkpoint(model, ckpt_path)
            print('checkpoint loaded')
        perturb = make_attack(attack, norm, eps, n_iter, verbose == 1, self.mixup_fn, alpha, noise_level, skip_projection)
        step = AdvTrainStep(model, attack=attack, perturb=perturb, distributed=bool(distributed), device=self.device,
                            ema=bool(model_ema), channels_last=bool(use_channel_last), mixup_fn=None,
                            graph_attack=attack == 'apgd' and self.mixup_fn is None,
                            # the whole step (attack, forward, backward, all-reduce, AdamW, EMA) as one CUDA graph; the
                            # host-side Mixup draw and the fgsm noise keep those configurations on the eager path
                            graph_step=(attack == 'apgd' and self.mixup_fn is None and optimizer == 'adamw' and
                                        os.environ.get('B200AT_GRAPH_STEP', '1') == '1'),
                            param_groups=lambda named: weight_decay_groups(named, arch, weight_decay),
                            optimizer=optimizer, momentum=momentum)
        return step

    def single_val(self):
        """clean top-1 on the first validation batches of this rank (main.py:905-942)."""
        self.model.eval()
        hit = n = 0
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
            for idx, (images, target) in enumerate(self.val_loader):
                images, target = images.to(self.device, non_blocking=True), target.to(self.device, non_blocking=True)
                out = self.step.raw.base_model(images) if self.step.perturb else self.step.raw(images)
                hit += (out.max(1)[1] == target).sum().item()
                n += target.shape[0]
                if idx >= 200:
                    break
        return hit / max(n, 1), n

    @param('logging.log_level')
    @param('adv.attack')
    def train_loop(self, epoch, log_level, attack):
        lr_start, lr_end = self.get_lr(epoch), self.get_lr(epoch + 1)
        iters = len(self.train_loader)
        lrs = np.interp(np.arange(iters), [0, iters], [lr_start, lr_end])
        losses, t0, seen = [], time.time(), 0
        for ix, (images, target) in enumerate(self.train_loader):
            images = images.to(self.device, non_blocking=True)
            target = target.to(self.device, non_blocking=True)
            if self.mixup_fn is not None:
                images, target = self.mixup_fn(images, target)
            self.step.set_lr(float(lrs[ix]))                          # a device tensor when the step is a CUDA graph
            loss = self.step(images, target)
            seen += images.shape[0]
            if log_level > 0:
                losses.append(loss)
                if log_level > 1 and self.gpu == 0:
                    print(f'ep={epoch}, iter={ix}, shape={tuple(images.shape)}, lr={lrs[ix]:.6f}, loss={loss.item():.3f}')
        torch.cuda.synchronize(self.device)
        self.images_per_sec = seen / max(time.time() - t0, 1e-9)
        if self.step.perturb:
            self.step.raw.set_perturb(False)                           # main.py:1020-1024
        return torch.stack(losses).mean() if losses else torch.zeros((), device=self.device)

    @param('training.epochs')
    @param('logging.log_level')
    @param('logging.save_freq')
    def train(self, epochs, log_level, save_freq):
        from revisiting_at_b200 import checkpoint
        acc, n = self.single_val()
        if log_level > 0 and self.gpu == 0:
            self.log({'Validation acc': acc, 'points': n})
        for epoch in range(epochs):
            sampler = getattr(self.train_loader, 'sampler', None)
            if hasattr(sampler, 'set_epoch'):
                sampler.set_epoch(epoch)                               # a new permutation / rank shard per epoch
            train_loss = self.train_loop(epoch)
            if log_level > 0:
                self.log({'train_loss': train_loss.item(), 'epoch': epoch, 'images_per_sec_rank0': self.images_per_sec})
            if train_loss.isnan():
                sys.exit('loss is NaN')
            if self.gpu == 0 and (epoch % max(save_freq, 1) == 0 or epoch == epochs - 1):
                ema = self.step.ema.state_dict() if self.step.ema is not None else None
                checkpoint.save_checkpoint(self.log_folder, epoch, self.step.raw, self.optimizer, ema, epochs)

    def eval_and_log(self):
        acc, n = self.single_val()
        self.log({'Validation acc': acc, 'points': n})

    @param('logging.folder')
    @param('model.arch')
    @param('adv.attack')
    @param('logging.addendum')
    def initialize_logger(self, folder, arch, attack, addendum):
        self.log_folder, self.start_time = None, time.time()
        if self.gpu != 0:
            return
        kind = f'adv_{addendum}' if attack != 'none' else f'clean_{addendum}'
        self.log_folder = (Path(folder) / f'model_{str(datetime.now())[:-7]}_{arch}_{kind}'.replace(' ', '_')).absolute()
        self.log_folder.mkdir(parents=True, exist_ok=True)
        print(f'=> Logging in {self.log_folder}')
        params = {'.'.join(k): self.all_params[k] for k in self.all_params.entries.keys()}
        with open(self.log_folder / 'params.json', 'w') as fh:
            json.dump(params, fh)

    def log(self, content):
        print(f'=> Log: {content}')
        if self.gpu != 0 or self.log_folder is None:
            return
        now = time.time()
        with open(self.log_folder / 'log', 'a') as fh:
            fh.write(json.dumps({'timestamp': now, 'relative_time': now - self.start_time, **content}) + '\n')

    # ------------------------------------------------------------------ launch (main.py:1128-1152)
    @classmethod
    @param('training.distributed')
    @param('dist.world_size')
    def launch_from_args(cls, distributed, world_size):
        if distributed and 'LOCAL_RANK' in os.environ:                 # already one process per GPU (torchrun)
            cls.exec(int(os.environ['LOCAL_RANK']))
        elif distributed:
            torch.multiprocessing.spawn(cls._exec_wrapper, args=(sys.argv[1:],), nprocs=world_size, join=True)
        else:
            cls.exec(0)

    @classmethod
    def _exec_wrapper(cls, gpu, argv):
        make_config(argv, quiet=True)
        cls.exec(gpu)

    @classmethod
    @param('training.distributed')
    @param('training.eval_only')
    def exec(cls, gpu, distributed, eval_only):
        trainer = cls(gpu=gpu)
        if eval_only:
            trainer.eval_and_log()
        else:
            trainer.train()
        if distributed:
            trainer.cleanup_distributed()


def make_config(argv=None, quiet=False):
    if _make_config is not None:
        return _make_config(argv, quiet)
    from argparse import ArgumentParser
    config = get_current_config()
    parser = ArgumentParser(description='Fast imagenet training')
    config.augment_argparse(parser)
    config.collect_argparse_args(parser)
    config.validate(mode='stderr')
    if not quiet:
        config.summary()
    return config


if __name__ == '__main__':
    make_config()
    ImageNetTrainer.launch_from_args()
