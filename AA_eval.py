#!/usr/bin/env python
"""Robust evaluation driver with the reference's command line (AA_eval.py:58-85, as launched by runner_aa_eval.py:8-16):

    python AA_eval.py --model_in <run folder | checkpoint file | random> --mod convnext_base --not-orig 1 --a100 1 \
        --full_aa 0 --l_norms Linf --batch_size 100 [--img_size 320] [--data_dir <ImageNet val folder | synthetic>]

Flow of the reference (AA_eval.py:87-252): fixed validation subset -> build the model (`get_new_model` + optional
normaliser, `params.json` of the run folder decides `add_normalization`) -> load the checkpoint (`module.` /
`base_model.` prefixes stripped) -> clean accuracy -> `AutoAttack(model, norm, eps, version='standard')` with
`attacks_to_run = ['apgd-ce', 'apgd-t']` unless `--full_aa 1` -> robust accuracy, logged to
`evaluated_logs_<norm>_<full_aa>_8_255.txt` in the run folder.  Here the model is the B200 engine and `AutoAttack` the
kernel-backed implementation (`autoattack/` shim), so the 100-iteration APGD-CE / APGD-T runs on the attack kernels.

As committed, the reference script cannot run (undefined `rann`, the runner passes an `--a100` flag the
parser does not define); this driver accepts the runner's flags as they are.  Differences: `robustbench.load_imagenet`
is absent in this image, so the validation subset comes from a torchvision ImageFolder with the reference's transform
(Resize(img/0.875, bicubic) -> CenterCrop -> ToTensor) or from `--data_dir synthetic` (`synthetic:self`: labels = the model's own predictions, so every attack runs); `--full_aa 1` (FAB-T, Square)
is refused; under `torchrun` the points are sharded over the ranks (the reference fans out one process per model).
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

eps_dict = {'imagenet': {'Linf': 4. / 255., 'L2': 2., 'L1': 75.}}          # AA_eval.py:33


class Logger:
    def __init__(self, log_path):
        self.log_path = log_path

    def log(self, str_to_log, verbose=False):
        print(str_to_log)
        if self.log_path is not None:
            with open(self.log_path, 'a') as f:
                f.write(str(str_to_log) + '\n')


def get_args_parser(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument('--batch_size', default=200, type=int)
    p.add_argument('--model', default='convnext_tiny', type=str)
    p.add_argument('--n_ex', type=int, default=5000)
    p.add_argument('--norm', type=str)
    p.add_argument('--eps', type=float)
    p.add_argument('--data_dir', type=str, default='synthetic')
    p.add_argument('--only_clean', action='store_true')
    p.add_argument('--save_imgs', action='store_true')
    p.add_argument('--precision', type=str, default='fp32')
    p.add_argument('--ckpt_path', type=str, default=None)
    p.add_argument('--mod', type=str)
    p.add_argument('--model_in', nargs='+')
    p.add_argument('--full_aa', type=int, default=0)
    p.add_argument('--init', type=str)
    p.add_argument('--add_normalization', action='store_true', default=False)
    p.add_argument('--l_norms', type=str, default='Linf')
    p.add_argument('--l_epss', type=str)
    p.add_argument('--get_stats', action='store_true')
    p.add_argument('--use_fixed_val_set', action='store_true', default=False)
    p.add_argument('--img_size', type=int, default=224, help='resolution to test the evaluation for')
    p.add_argument('--not_channel_last', action='store_false')
    p.add_argument('--not-original', type=int, default=1)       # the runner abbreviates it to --not-orig
    p.add_argument('--updated', action='store_true', default=False)
    p.add_argument('--a100', type=int, default=1)               # passed by runner_aa_eval.py:15
    return p.parse_args(argv)


def load_points(data_dir, n_ex, img_size, seed=0):
    """the fixed validation subset: [n, 3, S, S] in [0, 1] and labels, on the host"""
    if data_dir.startswith('synthetic'):                                # 'synthetic:self' -> labels = the model's own predictions
        g = torch.Generator().manual_seed(seed)
        return torch.rand(n_ex, 3, img_size, img_size, generator=g), torch.randint(0, 1000, (n_ex,), generator=g)
    from torchvision import datasets, transforms
    scale = int(math.floor(img_size / 0.875))
    tf = transforms.Compose([transforms.Resize(scale, interpolation=transforms.InterpolationMode.BICUBIC),
                             transforms.CenterCrop(img_size), transforms.ToTensor()])
    ds = datasets.ImageFolder(data_dir, tf)
    step = max(len(ds) // n_ex, 1)                                     # evenly spread, like a fixed class-balanced subset
    xs, ys = zip(*[ds[i] for i in range(0, step * n_ex, step)][:n_ex])
    return torch.stack(xs), torch.tensor(ys)


def resolve_checkpoint(model_in):
    """run folder -> weights_20.pt (AA_eval.py:124) or the latest weights_N.pt; file -> itself; 'random' -> None"""
    if model_in in (None, '', 'random'):
        return None, None
    if os.path.isdir(model_in):
        cand = os.path.join(model_in, 'weights_20.pt')
        if not os.path.exists(cand):
            files = [f for f in os.listdir(model_in) if f.startswith('weights_') and f.endswith('.pt') and 'ema' not in f]
            if not files:
                raise FileNotFoundError(f'no weights_N.pt in {model_in}')
            cand = os.path.join(model_in, max(files, key=lambda f: int(f[len('weights_'):-3])))
        return cand, model_in
    return model_in, os.path.dirname(model_in) or '.'


def build_model(arch, add_normalization, img_size):
    from revisiting_at_b200 import convnext, vit
    arch = arch.replace('timm_', '')
    if arch in convnext.ARCHS:
        m = convnext.ConvNeXtCvSt(arch)
        return convnext.Normalized(m) if add_normalization else m
    if arch in ('vit_s', 'deit_s', 'vit_small'):
        if img_size != 224:
            raise SystemExit('the ViT engine is built for 224 x 224 (197 tokens); evaluate ViT-S-CvSt at --img_size 224')
        return vit.build(normalize=add_normalization)
    raise SystemExit(f'--mod {arch!r}: the engine builds {sorted(convnext.ARCHS)} and vit_s')


@torch.no_grad()
def clean_accuracy(model, x, y, batch_size, device):
    hit = 0
    for i in range(0, x.shape[0], batch_size):
        xb, yb = x[i:i + batch_size].to(device), y[i:i + batch_size].to(device)
        hit += (model(xb).max(1)[1] == yb).sum().item()
    return hit / max(x.shape[0], 1)


def main(argv=None):
    args = get_args_parser(argv)
    if not torch.cuda.is_available():
        raise SystemExit('AA_eval.py: no CUDA device (the attack kernels have no CPU path)')
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import checkpoint
    from autoattack import AutoAttack
    distributed = 'LOCAL_RANK' in os.environ and int(os.environ.get('WORLD_SIZE', '1')) > 1
    local = int(os.environ.get('LOCAL_RANK', '0'))
    device = torch.device('cuda', local)
    torch.cuda.set_device(device)
    if distributed:
        torch.distributed.init_process_group('nccl', device_id=device)
    rank0 = not distributed or torch.distributed.get_rank() == 0

    model_in = ' '.join(args.model_in) if args.model_in else args.ckpt_path
    ckpt, savedir = resolve_checkpoint(model_in)
    if savedir is None:
        savedir = './results'
        os.makedirs(savedir, exist_ok=True)
    params_file = os.path.join(savedir, 'params.json')
    if os.path.exists(params_file):                                     # AA_eval.py:131-136
        with open(params_file) as f:
            params = json.load(f)
        if 'model.add_normalization' in params:
            args.add_normalization = params['model.add_normalization'] == 1
    arch = args.mod or args.model
    if args.eps is not None and args.eps > 1 and args.norm == 'Linf':
        args.eps /= 255.
    if args.full_aa:
        raise SystemExit('--full_aa 1 adds FAB-T and Square, which are outside the hot path; use --full_aa 0')

    x_test, y_test = load_points(args.data_dir, args.n_ex, args.img_size)
    print(f'{arch} has resolution : {args.img_size}')
    model = build_model(arch, args.add_normalization, args.img_size)
    if ckpt is not None:
        checkpoint.load_checkpoint(model, ckpt)
        print(f'loaded {ckpt}')
    model = model.to(device).eval()
    if args.data_dir == 'synthetic:self':                               # every point starts correctly classified: the attacks run
        with torch.no_grad():
            y_test = torch.cat([model(x_test[i:i + args.batch_size].to(device)).max(1)[1].cpu()
                                for i in range(0, x_test.shape[0], args.batch_size)])

    log_path = os.path.join(savedir, f'evaluated_logs_{args.l_norms}_{args.full_aa}_8_255.txt') if rank0 else None
    logger = Logger(log_path)
    acc = clean_accuracy(model, x_test, y_test, args.batch_size, device)
    logger.log(f'clean accuracy ({x_test.shape[0]} points): {acc:.2%}')
    if args.only_clean:
        return
    for nrm in [args.l_norms]:
        eps = args.eps if (args.eps is not None and args.norm == nrm) else eps_dict['imagenet'][nrm]
        adversary = AutoAttack(model, norm=nrm, eps=eps, version='standard', log_path=log_path, verbose=rank0, device=device)
        adversary.attacks_to_run = ['apgd-ce', 'apgd-t']                 # AA_eval.py:233-234
        assert not model.training
        x_adv = adversary.run_standard_evaluation(x_test, y_test, bs=args.batch_size)
        racc = clean_accuracy(model, x_adv, y_test, args.batch_size, device)
        logger.log(f'norm={nrm} eps={eps:.5f}\nrobust accuracy: {racc:.2%}')
        if args.save_imgs and rank0:
            torch.save(x_adv.cpu(), os.path.join(savedir, f'aa_short_1_{args.n_ex}_{nrm}_{eps:.5f}.pth'))
    if distributed:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
