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
    """Inside real optimiser steps (bf16 autocast, channels_last parameters, fused AdamW): at every step the
    replayed attack returns bit for bit what the eager launch sequence returns on the same parameter state.
    (Whole training runs are not compared: the weight-gradient kernels accumulate with atomics, so two eager
    runs already differ in the last bits of the parameters after one step.)"""
    from revisiting_at_b200 import convnext
    from revisiting_at_b200.train_step import AdvTrainStep
    base = convnext.build('convnext_tiny', normalize=True, seed=0)
    g = torch.Generator().manual_seed(2)
    batches = [(torch.rand(8, 3, 64, 64, generator=g).to(cuda_dev), torch.randint(0, 1000, (8,), generator=g).to(cuda_dev))
               for _ in range(6)]
    step = AdvTrainStep(copy.deepcopy(base), 'apgd', 'Linf', 4. / 255., 2, device=cuda_dev, graph_attack=True)
    graphed, eager = step.graphed_perturb, step.eager_perturb
    checked = []

    def spy(model, x, y):
        ref = [t.clone() for t in eager(model, x, y)]
        out = graphed(model, x, y)
        checked.append(all(torch.equal(a, b) for a, b in zip(out, ref)))
        return out
    step.raw.perturb = spy
    losses = [step(x, y).item() for x, y in batches]
    assert checked == [True] * 6, checked
    assert len(graphed.graphs) == 1 and all(l == l and l < 20 for l in losses), losses
    # the forward really runs on the UPDATED parameters (fused AdamW does not move version counters: a stale
    # kernel-side weight copy would show up here): engine logits == fp32 oracle logits on the trained state
    from oracle import convnext_oracle as co
    o = co.build('convnext_tiny', normalize=True, seed=0)
    o.load_state_dict({k[len('base_model.'):]: v.detach().cpu().float().contiguous()
                       for k, v in step.raw.state_dict().items()})
    o0 = co.build('convnext_tiny', normalize=True, seed=0)
    x = batches[0][0][:4]
    step.raw.base_model.eval()
    with torch.no_grad():
        got = step.raw.base_model(x).float().cpu()
        want, stale = o.eval()(x.cpu()), o0.eval()(x.cpu())
    assert (got - want).abs().max() <= 5e-2, (got - want).abs().max()
    assert (want - stale).abs().max() > 0.2            # the six steps moved the logits far beyond that tolerance
