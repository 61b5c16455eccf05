"""CPU restatement of one adversarial training step -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Follows /root/reference/main.py: `WrappedModel.forward` (:273-293: eval -> apgd_train -> train ->
forward on x_best) inside the body of `train_loop` (:984-997: zero_grad, forward, CrossEntropy,
backward, AdamW betas .9/.95 step).  fp32 on the host cores (CUDA autocast/GradScaler have no CPU
meaning).  Used by bench.py's `cpu_baseline` leg and `--impl reference` arm (there with `attack=` the
UNMODIFIED reference `apgd_train` and `model` the reference's own modules when `oracle/_ref` is staged:
`reference_train_step`), and by tests/test_gpu_full_loop.py::test_train_step_matches_the_oracle_step; never by
the product."""
import torch
import torch.nn.functional as F

from .apgd_oracle import apgd_train_oracle


class OracleTrainStep:
    def __init__(self, model, norm='Linf', eps=4 / 255., n_iter=2, lr=1e-3, weight_decay=0.05, label_smoothing=0.,
                 attack=None):
        self.model = model
        self.attack = attack if attack is not None else apgd_train_oracle
        self.norm, self.eps, self.n_iter = norm, eps, n_iter
        decay = [p for p in model.parameters() if p.ndim > 1]
        no_decay = [p for p in model.parameters() if p.ndim <= 1]
        self.opt = torch.optim.AdamW([{'params': decay, 'weight_decay': weight_decay},
                                      {'params': no_decay, 'weight_decay': 0.}], lr=lr, betas=(0.9, 0.95))
        self.label_smoothing = label_smoothing

    def __call__(self, images, target):
        self.model.eval()
        x_best = self.attack(self.model, images, target, self.norm, self.eps, n_iter=self.n_iter)[0]
        self.model.train()
        self.opt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(self.model(x_best), target, label_smoothing=self.label_smoothing)
        loss.backward()
        self.opt.step()
        return loss.detach()


def reference_train_step(norm='Linf', eps=4 / 255., n_iter=2, seed=0):
    """The step with the reference's OWN `apgd_train` and ConvNeXt-T-CvSt modules (`oracle/_ref` or the mounted
    reference, loaded by file path; oracle seed-0 weights so both arms of the bench hold the same parameters), or
    None when no reference file is reachable.  Only the ~10 lines of `train_loop` around them are restated:
    main.py itself imports fastargs / timm / torchmetrics, none of which are in this image."""
    from . import convnext_oracle, ref_loader
    if not ref_loader.available():
        return None
    sd = convnext_oracle.build('convnext_tiny', normalize=False, seed=seed).state_dict()
    model = ref_loader.convnext_t_cvst_normalized(sd)
    return OracleTrainStep(model, norm, eps, n_iter, attack=ref_loader.attack_module().apgd_train)
