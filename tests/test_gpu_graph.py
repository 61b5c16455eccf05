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


def test_whole_step_graph_trains_like_the_eager_step(cuda_dev):
    """`graph_step=True`: attack + training forward / backward + fused AdamW (+ EMA) replayed from ONE CUDA graph.  Two
    trainers from the same weights see the same six batches; the weight-gradient kernels use atomics, so the parameters
    are compared with a tolerance, and the learning-rate tensor is changed between replays to show schedules still act."""
    from revisiting_at_b200 import convnext
    from revisiting_at_b200.train_step import AdvTrainStep
    base = convnext.build('convnext_tiny', normalize=True, seed=0)
    g = torch.Generator().manual_seed(4)
    batches = [(torch.rand(8, 3, 64, 64, generator=g).to(cuda_dev), torch.randint(0, 1000, (8,), generator=g).to(cuda_dev))
               for _ in range(6)]
    eager = AdvTrainStep(copy.deepcopy(base), 'apgd', 'Linf', 4. / 255., 2, device=cuda_dev, graph_attack=True, ema=True)
    whole = AdvTrainStep(copy.deepcopy(base), 'apgd', 'Linf', 4. / 255., 2, device=cuda_dev, graph_attack=True, ema=True,
                         graph_step=True)
    le, lw = [], []
    for i, (x, y) in enumerate(batches):
        lr = 1e-3 * (1 + i)
        eager.set_lr(lr); whole.set_lr(lr)
        le.append(eager(x, y).item()); lw.append(whole(x, y).item())
    assert len(whole.graphed.graphs) == 1
    assert all(abs(a - b) <= 2e-2 * abs(a) for a, b in zip(le, lw)), (le, lw)
    pe = torch.cat([p.detach().flatten() for p in eager.raw.parameters()])
    pw = torch.cat([p.detach().flatten() for p in whole.raw.parameters()])
    moved = (pe - torch.cat([p.detach().flatten() for p in base.parameters()]).to(cuda_dev)).abs().max().item()
    assert moved > 1e-3                                                   # six AdamW steps with growing lr did move them
    cos = torch.nn.functional.cosine_similarity(pe - pw.new_zeros(()), pw, dim=0).item()
    assert cos > 0.9999, cos
    # the replay really applies the CURRENT learning rate: a step with lr = 0 leaves the parameters untouched
    whole.set_lr(0.0)
    before = pw.clone()
    whole(*batches[0])
    after = torch.cat([p.detach().flatten() for p in whole.raw.parameters()])
    assert torch.equal(before, after)
    # the EMA shadow is updated inside the graph as well
    ee = torch.cat([t.flatten() for t in eager.ema.shadow]); ew = torch.cat([t.flatten() for t in whole.ema.shadow])
    assert (ee - ew).abs().max().item() <= 1e-4 and not torch.equal(ew, torch.cat([p.detach().flatten() for p in base.parameters()]).to(cuda_dev))
