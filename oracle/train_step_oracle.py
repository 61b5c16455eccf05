"""CPU restatement of one adversarial training step -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Follows /root/reference/main.py: `WrappedModel.forward` (:273-293: eval -> apgd_train -> train ->
forward on x_best) inside the body of `train_loop` (:984-997: zero_grad, forward, CrossEntropy,
backward, AdamW betas .9/.95 step).  fp32 on the host cores (CUDA autocast/GradScaler have no CPU
meaning).  Used by bench.py's `cpu_baseline` leg and `--impl reference` arm, and by the train-step
parity test; never by the product."""
import torch
import torch.nn.functional as F

from .apgd_oracle import apgd_train_oracle


class OracleTrainStep:
    def __init__(self, model, norm='Linf', eps=4 / 255., n_iter=2, lr=1e-3, weight_decay=0.05, label_smoothing=0.):
        self.model = model
        self.norm, self.eps, self.n_iter = norm, eps, n_iter
        decay = [p for p in model.parameters() if p.ndim > 1]
        no_decay = [p for p in model.parameters() if p.ndim <= 1]
        self.opt = torch.optim.AdamW([{'params': decay, 'weight_decay': weight_decay},
                                      {'params': no_decay, 'weight_decay': 0.}], lr=lr, betas=(0.9, 0.95))
        self.label_smoothing = label_smoothing

    def __call__(self, images, target):
        self.model.eval()
        x_best = apgd_train_oracle(self.model, images, target, self.norm, self.eps, n_iter=self.n_iter)[0]
        self.model.train()
        self.opt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(self.model(x_best), target, label_smoothing=self.label_smoothing)
        loss.backward()
        self.opt.step()
        return loss.detach()
