"""GPU: the CUDA-graph replay of the attack (train_step.GraphedAttack) is the same computation as the eager
launch sequence -- bit-identical adversarial examples, also after the parameters change under the graph."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graphed_attack_equals_eager_and_tracks_weight_updates(cuda_dev):
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import convnext
    from revisiting_at_b200.train_step import GraphedAttack, make_attack
    m = convnext.build('convnext_tiny', normalize=True, seed=0).to(cuda_dev).eval()
    eager = make_attack('apgd', 'Linf', 4. / 255., 2)
    graphed = GraphedAttack(eager, warmup=1)
    g = torch.Generator().manual_seed(9)
    xs = [torch.rand(8, 3, 64, 64, generator=g).to(cuda_dev) for _ in range(4)]
    ys = [torch.randint(0, 1000, (8,), generator=g).to(cuda_dev) for _ in range(4)]
    for i in range(4):                       # call 0 eager warm-up, call 1 captures + replays, calls 2.. replay
        if i == 3:                           # an "optimiser step": parameters change in place under the graph
            with torch.no_grad():
                for p in m.parameters():
                    p.add_(0.05 * torch.randn_like(p))
        got = [t.clone() for t in graphed(m, xs[i], ys[i])]
        ref = eager(m, xs[i], ys[i])
        for a, b in zip(got, ref):
            assert torch.equal(a, b), i
    assert len(graphed.graphs) == 1


def test_train_step_with_graphed_attack_matches_eager(cuda_dev):
    """three optimiser steps, graph on vs off, same seeds: identical losses and parameters"""
    from revisiting_at_b200 import convnext
    from revisiting_at_b200.train_step import AdvTrainStep
    base = convnext.build('convnext_tiny', normalize=True, seed=0)
    g = torch.Generator().manual_seed(2)
    batches = [(torch.rand(8, 3, 64, 64, generator=g).to(cuda_dev), torch.randint(0, 1000, (8,), generator=g).to(cuda_dev))
               for _ in range(5)]
    out = []
    for graph in (False, True):
        step = AdvTrainStep(copy.deepcopy(base), 'apgd', 'Linf', 4. / 255., 2, device=cuda_dev, graph_attack=graph)
        losses = [step(x, y).item() for x, y in batches]
        out.append((losses, [p.detach().clone() for p in step.raw.parameters()]))
    assert out[0][0] == out[1][0], (out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert torch.equal(a, b)
