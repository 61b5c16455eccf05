"""pinpoint where the graphed train step departs from the eager one (debug aid)"""
import copy, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import revisiting_at_b200  # noqa
from revisiting_at_b200 import convnext, ops
from revisiting_at_b200.train_step import AdvTrainStep
dev = torch.device('cuda:0')
base = convnext.build('convnext_tiny', normalize=True, seed=0)
g = torch.Generator().manual_seed(2)
batches = [(torch.rand(8, 3, 64, 64, generator=g).to(dev), torch.randint(0, 1000, (8,), generator=g).to(dev)) for _ in range(5)]

def run(graph, hook=None):
    step = AdvTrainStep(copy.deepcopy(base), 'apgd', 'Linf', 4. / 255., 2, device=dev, graph_attack=graph)
    rec = []
    orig = step.raw.perturb
    def spy(model, x, y):
        out = orig(model, x, y)
        rec.append([t.clone() for t in out])
        if hook: hook()
        return out
    step.raw.perturb = spy
    if graph:
        step.graphed_perturb = spy
    losses = []
    params = []
    for x, y in batches:
        losses.append(step(x, y).item())
        params.append([p.detach().clone() for p in step.raw.parameters()])
    return losses, rec, params

for name, kw in (('eager', dict(graph=False)), ('eager2', dict(graph=False)), ('graph', dict(graph=True)),
                 ('graph+invalidate-after-attack', dict(graph=True, hook=ops.invalidate_derived))):
    l, rec, params = run(**kw)
    if name == 'eager':
        l0, rec0, p0 = l, rec, params
    print(name, l)
    for i in range(5):
        same_attack = all(torch.equal(a, b) for a, b in zip(rec[i], rec0[i]))
        same_params = all(torch.equal(a, b) for a, b in zip(params[i], p0[i]))
        nbad = sum(int(not torch.equal(a, b)) for a, b in zip(params[i], p0[i]))
        print(f'   step {i}: attack outputs equal {same_attack}; params after step equal {same_params} ({nbad} tensors differ)')
