"""Tiny real model for model-loop goldens -- TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class SmallCNN(nn.Module):
    """Logits and input-gradients depend on x; weights come from the seed / the fixture."""
    def __init__(self, C=10):
        super().__init__()
        self.c1 = nn.Conv2d(3, 8, 3, stride=2, padding=1)
        self.c2 = nn.Conv2d(8, 16, 3, stride=2, padding=1)
        self.fc = nn.Linear(16, C)

    def forward(self, x):
        h = F.gelu(self.c1(x))
        h = F.gelu(self.c2(h))
        return self.fc(h.mean((-2, -1))) * 20.


def from_fixture(g) -> SmallCNN:
    m = SmallCNN()
    sd = {k: torch.from_numpy(np.asarray(g['w_' + k.replace('.', '_')])) for k in m.state_dict()}
    m.load_state_dict(sd)
    return m.eval()


class SensitiveNet(nn.Module):
    """Small model whose predictions flip under perturbations of a few /255 (conv -> dense -> dense): exercises the
    restart / target-class compaction of the AutoAttack protocol, which the mean-pooled SmallCNN never triggers."""
    def __init__(self, C=10, hw=16, k=3.0):
        super().__init__()
        self.c1 = nn.Conv2d(3, 8, 3, stride=2, padding=1)
        self.fc1 = nn.Linear(8 * (hw // 2) ** 2, 32)
        self.fc2 = nn.Linear(32, C)
        self.k = k

    def forward(self, x):
        h = F.gelu(self.c1(x))
        h = F.gelu(self.fc1(h.flatten(1)))
        return self.fc2(h) * self.k
