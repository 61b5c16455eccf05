"""AutoAttack-compatible APGD-CE / APGD-T evaluation on the B200 attack kernels (BASELINE config 5).

The reference evaluates with the un-vendored pip package autoattack-0.1 (README.md:15):

    adversary = AutoAttack(model, norm=..., eps=..., version='standard')          AA_eval.py:226-231
    adversary.attacks_to_run = ['apgd-ce', 'apgd-t']                              AA_eval.py:233-234 (full_aa=0)
    x_adv = adversary.run_standard_evaluation(x, y, bs=bs)                        AA_eval.py:237-239

This module exposes the same classes, constructor arguments, attributes and methods (`AutoAttack`,
`APGDAttack`, `APGDAttack_targeted`; top-level shim package `autoattack/`), so AA_eval.py's calls bind to
it unchanged.  The package's source is not under /root/reference; its published algorithm
(`autoattack/autopgd_base.py`, `autoattack/autoattack.py`, fra31/auto-attack) is restated -- **parity
unpinned** except where the reference's own fork pins it: `apgd_train` (autopgd_train_clean.py:123-371) is
that file's `attack_single_run` with the random start removed and the last backward skipped, and
`dlr_loss` / `dlr_loss_targeted` (:99-111) are the two DLR losses.  Everything per-iteration therefore runs
on the golden-tested kernels through `attack.run_apgd` (copying form: fused l-inf / l2 / l1 step, one
loss+bookkeeping launch per forward, no host synchronisation inside a run); what this module adds is the
host-side protocol around a run:

  * random start (autopgd_base.py `attack_single_run`): x + eps * t / max|t|, t ~ U(-1,1) (l-inf);
    x + eps * t / ||t||_2, t ~ N(0,1) (l2); x + t + L1_projection(x, t, eps) (l1);
  * restarts and target classes only over the points that are still robust (the compaction between runs is
    the protocol's one data-dependent host decision per run);
  * `use_largereps` (l1, version='standard'): 3 runs at 3 eps / 2 eps / eps with 30 / 30 / 40 % of the
    iterations, each started from the previous run's highest-loss point re-projected on the new ball;
  * `run_standard_evaluation`: clean pass, then each attack over the still-robust points in batches of `bs`;
    with `torch.distributed` initialised the points are sharded by rank (contiguous shards, no
    communication during the attacks) and the flags / adversarial points are all-gathered once at the end.

'fab-t' and 'square' (version='standard' with full_aa=1) are outside the hot path and raise.
"""
from __future__ import annotations

import math
import time

import torch

from . import attack as _attack
from .compat import L1_projection


class APGDAttack:
    """autoattack.autopgd_base.APGDAttack (same constructor arguments and `perturb` contract)."""

    def __init__(self, predict, n_iter=100, norm='Linf', n_restarts=1, eps=None, seed=0, loss='ce', eot_iter=1,
                 rho=.75, topk=None, verbose=False, device=None, use_largereps=False, is_tf_model=False, logger=None,
                 backend=None):
        if eot_iter != 1:
            raise NotImplementedError('eot_iter != 1 (randomised defences) is outside the hot path')
        if is_tf_model:
            raise NotImplementedError('TensorFlow models are outside the hot path')
        if rho != .75:
            raise NotImplementedError('the checkpoint threshold is fixed at rho=0.75 in the kernels')
        self.model = predict
        self.n_iter = n_iter
        self.eps = eps
        self.norm = norm
        self.n_restarts = n_restarts
        self.seed = seed
        self.loss = loss
        self.eot_iter = eot_iter
        self.thr_decr = rho
        self.topk = topk
        self.verbose = verbose
        self.device = device
        self.use_rs = True
        self.use_largereps = use_largereps
        self.n_iter_orig = n_iter + 0
        self.eps_orig = eps + 0. if eps is not None else None
        self.is_tf_model = is_tf_model
        self.y_target = None
        self.logger = logger
        self._be = backend if backend is not None else _attack._CUDA
        assert self.norm in ['Linf', 'L2', 'L1']
        assert self.eps is not None

    # ---------------------------------------------------------------- pieces of attack_single_run
    # where the random starts are drawn: None = on the device of x (no H2D copy); 'cpu' = on the host and copied, as
    # the package does (`torch.rand(x.shape).to(device)`), which also makes runs comparable across devices
    rng_device = None

    def _generator(self, x):
        g = torch.Generator(device=self.rng_device if self.rng_device is not None else x.device)
        g.manual_seed(int(self.seed) & 0x7fffffffffffffff)
        return g

    def _random_start(self, x, eps, gen):
        B = x.shape[0]
        exp = (B,) + (1,) * (x.dim() - 1)
        if self.norm == 'Linf':
            t = (2 * torch.rand(x.shape, device=gen.device, dtype=x.dtype, generator=gen) - 1).to(x.device)
            return x + eps * t / (t.abs().reshape(B, -1).max(1)[0].view(exp) + 1e-12)
        t = torch.randn(x.shape, device=gen.device, dtype=x.dtype, generator=gen).to(x.device)
        if self.norm == 'L2':
            return x + eps * t / ((t ** 2).reshape(B, -1).sum(-1).sqrt().view(exp) + 1e-12)
        return x + t + self._l1_projection(x, t, eps)

    def _l1_projection(self, x, d, eps):
        if self._be is _attack._CUDA:
            return L1_projection(x, d, eps)
        return self._be.l1_projection(x, d, eps)

    def attack_single_run(self, x, y, x_init=None, gen=None, restart_state=False):
        """one APGD run from a random start (or `x_init`): (x_best, acc, loss_best, x_best_adv)"""
        if x_init is None:
            x_init = self._random_start(x, self.eps, gen if gen is not None else self._generator(x))
        was_training = getattr(self.model, 'training', False)
        if was_training:
            self.model.eval()
        try:
            return _attack.run_apgd(self._be, self.model, x, y, self.norm, self.eps, n_iter=self.n_iter, loss=self.loss,
                                    verbose=self.verbose, is_train=False, x_init=x_init, y_target=self.y_target,
                                    l1_restart_state=restart_state)
        finally:
            if was_training:
                self.model.train()

    def decr_eps_pgd(self, x, y, epss, iters, gen=None):
        """l1 large-eps schedule (autopgd_base.py `decr_eps_pgd`)"""
        assert len(epss) == len(iters) and self.norm == 'L1'
        gen = gen if gen is not None else self._generator(x)
        x_init = x + torch.randn(x.shape, device=gen.device, dtype=x.dtype, generator=gen).to(x.device)
        x_init = x_init + self._l1_projection(x, x_init - x, 1. * float(epss[0]))
        n_iter, eps = self.n_iter, self.eps
        try:
            for e, niter in zip(epss, iters):
                self.n_iter, self.eps = niter + 0, e + 0.
                x_init = x_init + self._l1_projection(x, x_init - x, 1. * e)
                x_init, acc, loss, x_adv = self.attack_single_run(x, y, x_init=x_init, restart_state=True)
        finally:
            self.n_iter, self.eps = n_iter, eps
        return x_init, acc, loss, x_adv

    def _schedule(self):
        epss = [3. * self.eps_orig, 2. * self.eps_orig, 1. * self.eps_orig]
        iters = [math.ceil(c) for c in (.3 * self.n_iter_orig, .3 * self.n_iter_orig, .4 * self.n_iter_orig)]
        iters[-1] = self.n_iter_orig - sum(iters[:-1])
        return epss, iters

    def _predict(self, x):
        with torch.no_grad(), _attack.input_grad_only():
            return self.model(x)

    def _run(self, x, y, gen):
        if self.use_largereps:
            epss, iters = self._schedule()
            return self.decr_eps_pgd(x, y, epss, iters, gen)
        return self.attack_single_run(x, y, gen=gen)

    def perturb(self, x, y=None, best_loss=False, x_init=None):
        """:param x: clean images  :param y: labels (None: the model's predictions)
        Returns the adversarial points found (the clean point where the attack failed)."""
        assert self.loss in ['ce', 'dlr']
        if best_loss:
            raise NotImplementedError('best_loss=True is not used by the standard evaluation')
        if y is not None and y.dim() == 0:
            x, y = x.unsqueeze(0), y.unsqueeze(0)
        x = x.detach().clone().float()
        y_pred = self._predict(x).max(1)[1]
        y = y_pred.detach().clone().long() if y is None else y.detach().clone().long().to(x.device)
        adv = x.clone()
        acc = y_pred == y
        gen = self._generator(x)
        for _ in range(self.n_restarts):
            ind_to_fool = acc.nonzero().flatten()                        # the protocol's host decision per run
            if ind_to_fool.numel() == 0:
                break
            best_curr, acc_curr, loss_curr, adv_curr = self._run(x[ind_to_fool].clone(), y[ind_to_fool].clone(), gen)
            ind_curr = (acc_curr == 0).nonzero().flatten()
            acc[ind_to_fool[ind_curr]] = False
            adv[ind_to_fool[ind_curr]] = adv_curr[ind_curr]
            if self.verbose:
                print('restart - robust accuracy: {:.2%}'.format(acc.float().mean().item()))
        return adv


class APGDAttack_targeted(APGDAttack):
    """autoattack.autopgd_base.APGDAttack_targeted: APGD on the targeted DLR loss against the
    `n_target_classes` most likely wrong classes of each point."""

    def __init__(self, predict, n_iter=100, norm='Linf', n_restarts=1, eps=None, seed=0, eot_iter=1, rho=.75,
                 topk=None, n_target_classes=9, verbose=False, device=None, use_largereps=False, is_tf_model=False,
                 logger=None, backend=None):
        super().__init__(predict, n_iter=n_iter, norm=norm, n_restarts=n_restarts, eps=eps, seed=seed,
                         loss='dlr-targeted', eot_iter=eot_iter, rho=rho, topk=topk, verbose=verbose, device=device,
                         use_largereps=use_largereps, is_tf_model=is_tf_model, logger=logger, backend=backend)
        self.y_target = None
        self.n_target_classes = n_target_classes

    def perturb(self, x, y=None, x_init=None):
        assert self.loss in ['dlr-targeted']
        if y is not None and y.dim() == 0:
            x, y = x.unsqueeze(0), y.unsqueeze(0)
        x = x.detach().clone().float()
        y_pred = self._predict(x).max(1)[1]
        y = y_pred.detach().clone().long() if y is None else y.detach().clone().long().to(x.device)
        adv = x.clone()
        acc = y_pred == y
        gen = self._generator(x)
        try:
            for target_class in range(2, self.n_target_classes + 2):
                for _ in range(self.n_restarts):
                    ind_to_fool = acc.nonzero().flatten()
                    if ind_to_fool.numel() == 0:
                        break
                    x_to_fool, y_to_fool = x[ind_to_fool].clone(), y[ind_to_fool].clone()
                    output = self._predict(x_to_fool)
                    self.y_target = output.float().sort(dim=1)[1][:, -target_class].contiguous()
                    best_curr, acc_curr, loss_curr, adv_curr = self._run(x_to_fool, y_to_fool, gen)
                    ind_curr = (acc_curr == 0).nonzero().flatten()
                    acc[ind_to_fool[ind_curr]] = False
                    adv[ind_to_fool[ind_curr]] = adv_curr[ind_curr]
                    if self.verbose:
                        print('target class {} - robust accuracy: {:.2%}'.format(target_class, acc.float().mean().item()))
        finally:
            self.y_target = None
        return adv


class AutoAttack:
    """autoattack.AutoAttack restricted to the attacks the reference runs (AA_eval.py:226-239)."""

    def __init__(self, model, norm='Linf', eps=.3, seed=None, verbose=True, attacks_to_run=[], version='standard',
                 is_tf_model=False, device='cuda', log_path=None, backend=None):
        self.model = model
        self.norm = norm
        assert norm in ['Linf', 'L2', 'L1']
        self.epsilon = eps
        self.seed = seed
        self.verbose = verbose
        self.attacks_to_run = list(attacks_to_run)
        self.version = version
        self.is_tf_model = is_tf_model
        self.device = device
        self.log_path = log_path
        if version in ['standard', 'plus', 'rand'] and self.attacks_to_run != []:
            raise ValueError('attacks_to_run will be overridden unless you use version=\'custom\'')
        if is_tf_model:
            raise NotImplementedError('TensorFlow models are outside the hot path')
        self.apgd = APGDAttack(self.model, n_restarts=5, n_iter=100, verbose=False, eps=self.epsilon, norm=self.norm,
                               eot_iter=1, rho=.75, seed=self.seed if self.seed is not None else 0, device=self.device,
                               backend=backend)
        self.apgd_targeted = APGDAttack_targeted(self.model, n_restarts=1, n_iter=100, verbose=False,
                                                 eps=self.epsilon, norm=self.norm, eot_iter=1, rho=.75,
                                                 seed=self.seed if self.seed is not None else 0, device=self.device,
                                                 backend=backend)
        if version in ['standard', 'plus', 'rand']:
            self.set_version(version)

    def set_version(self, version='standard'):
        if self.verbose:
            print('setting parameters for {} version'.format(version))
        if version == 'standard':
            self.attacks_to_run = ['apgd-ce', 'apgd-t', 'fab-t', 'square']
            if self.norm in ['Linf', 'L2']:
                self.apgd.n_restarts = 1
                self.apgd_targeted.n_target_classes = 9
            else:
                self.apgd.use_largereps = True
                self.apgd_targeted.use_largereps = True
                self.apgd.n_restarts = 5
                self.apgd_targeted.n_target_classes = 5
            self.apgd_targeted.n_restarts = 1
        else:
            raise NotImplementedError(f"version {version!r}: only 'standard' and 'custom' are on the reference's path")

    def get_logits(self, x):
        with torch.no_grad(), _attack.input_grad_only():
            return self.model(x)

    def get_seed(self):
        return time.time() if self.seed is None else self.seed

    def _log(self, msg):
        if self.verbose:
            print(msg)
        if self.log_path is not None:
            with open(self.log_path, 'a') as f:
                f.write(msg + '\n')

    def clean_accuracy(self, x_orig, y_orig, bs=250):
        acc = 0.
        for s in range(0, x_orig.shape[0], bs):
            x, y = x_orig[s:s + bs].to(self.device), y_orig[s:s + bs].to(self.device)
            acc += (self.get_logits(x).max(1)[1] == y).float().sum().item()
        self._log('clean accuracy: {:.2%}'.format(acc / x_orig.shape[0]))
        return acc / x_orig.shape[0]

    def _shard(self, n, shard):
        if shard and torch.distributed.is_available() and torch.distributed.is_initialized():
            w, r = torch.distributed.get_world_size(), torch.distributed.get_rank()
            per = (n + w - 1) // w
            return min(r * per, n), min((r + 1) * per, n), w
        return 0, n, 1

    def run_standard_evaluation(self, x_orig, y_orig, bs=250, return_labels=False, state_path=None, shard=True):
        """robust evaluation of `x_orig` ([N,3,H,W] in [0,1], any device): returns x_adv (and the predicted labels
        on it) with the clean point kept wherever every attack failed.  `shard`: split the N points over the ranks of
        an initialised process group (no communication until the final all-gather)."""
        if state_path is not None:
            raise NotImplementedError('state_path (resumable evaluation) is outside the hot path')
        for a in self.attacks_to_run:
            if a not in ('apgd-ce', 'apgd-t'):
                raise NotImplementedError(f'attack {a!r} is outside the hot path; set attacks_to_run = '
                                          "['apgd-ce', 'apgd-t'] as AA_eval.py does for full_aa=0")
        if self.verbose:
            print('using {} version including {}'.format(self.version, ', '.join(self.attacks_to_run)))
        N = x_orig.shape[0]
        lo, hi, world = self._shard(N, shard)
        xs, ys = x_orig[lo:hi], y_orig[lo:hi]
        n = hi - lo
        dev = torch.device(self.device) if not isinstance(self.device, torch.device) else self.device
        robust = torch.zeros(n, dtype=torch.bool, device=xs.device)
        y_adv = torch.empty(n, dtype=torch.long, device=xs.device)
        x_adv = xs.clone().detach()
        t0 = time.time()
        for s in range(0, n, bs):
            x, y = xs[s:s + bs].to(dev), ys[s:s + bs].to(dev)
            out = self.get_logits(x).max(dim=1)[1]
            y_adv[s:s + bs] = out.to(y_adv.device)
            robust[s:s + bs] = y.eq(out).to(robust.device)
        results = {'clean': self._global_fraction(robust, N, world)}
        self._log('initial accuracy: {:.2%}'.format(results['clean']))
        for name in self.attacks_to_run:
            idcs = robust.nonzero().flatten()
            for s in range(0, idcs.numel(), bs):
                bi = idcs[s:s + bs]
                x, y = xs[bi].clone().to(dev), ys[bi].clone().to(dev)
                if name == 'apgd-ce':
                    self.apgd.loss = 'ce'
                    self.apgd.seed = self.get_seed()
                    adv = self.apgd.perturb(x, y)
                else:
                    self.apgd_targeted.seed = self.get_seed()
                    adv = self.apgd_targeted.perturb(x, y)
                out = self.get_logits(adv).max(dim=1)[1]
                false_batch = (~y.eq(out)).to(robust.device)
                bad = bi[false_batch]
                robust[bad] = False
                x_adv[bad] = adv[false_batch.to(adv.device)].detach().to(x_adv.device)
                y_adv[bad] = out[false_batch.to(out.device)].to(y_adv.device)
                if self.verbose:
                    print('{} - {}/{} - {} out of {} successfully perturbed'.format(
                        name, s // bs + 1, (idcs.numel() + bs - 1) // bs, int(false_batch.sum()), x.shape[0]))
            results[name] = self._global_fraction(robust, N, world)
            self._log('robust accuracy after {}: {:.2%} (total time {:.1f} s)'.format(
                name.upper(), results[name], time.time() - t0))
        self.results = results
        if world > 1:
            x_adv, y_adv = self._all_gather(x_adv, N, world, dev), self._all_gather(y_adv, N, world, dev)
        self._log('robust accuracy: {:.2%}'.format(results[self.attacks_to_run[-1]] if self.attacks_to_run else results['clean']))
        return (x_adv, y_adv) if return_labels else x_adv

    @staticmethod
    def _global_fraction(flags, N, world):
        c = flags.sum().to(torch.float64)
        if world > 1:
            c = c.to(flags.device if flags.is_cuda else 'cpu')
            if torch.distributed.get_backend() == 'nccl' and not c.is_cuda:
                c = c.cuda()
            torch.distributed.all_reduce(c)
        return c.item() / N

    @staticmethod
    def _all_gather(t, N, world, dev):
        per = (N + world - 1) // world
        comm_dev = dev if torch.distributed.get_backend() == 'nccl' else torch.device('cpu')
        pad = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=comm_dev)
        pad[:t.shape[0]] = t.to(comm_dev)
        parts = [torch.empty_like(pad) for _ in range(world)]
        torch.distributed.all_gather(parts, pad)
        return torch.cat(parts)[:N].to(t.device)
