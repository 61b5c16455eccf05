"""where does the replayed attack depart from the eager one inside a train step? (debug aid)"""
import copy, os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import revisiting_at_b200  # noqa
from revisiting_at_b200 import convnext, ops
from revisiting_at_b200.train_step import GraphedAttack, make_attack
dev = torch.device('cuda:0')
base = convnext.build('convnext_tiny', normalize=True, seed=0)
g = torch.Generator().manual_seed(2)
xs = [torch.rand(8, 3, 64, 64, generator=g).to(dev) for _ in range(4)]
ys = [torch.randint(0, 1000, (8,), generator=g).to(dev) for _ in range(4)]

def diff(a, b):
    return [((p.float() - q.float()).abs().max().item(), int((p != q).sum().item())) for p, q in zip(a, b)]

for autocast, cl, order in itertools.product((False, True), (False, True), ('eager_first', 'graph_first')):
    m = copy.deepcopy(base)
    if cl:
        m = m.to(memory_format=torch.channels_last)
    m = m.to(dev).eval()
    eager = make_attack('apgd', 'Linf', 4. / 255., 2)
    graphed = GraphedAttack(eager, warmup=0)
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
        eager(m, xs[3], ys[3])
    res = []
    for i in range(3):
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
            if order == 'eager_first':
                ref = [t.clone() for t in eager(m, xs[i], ys[i])]
                out = [t.clone() for t in graphed(m, xs[i], ys[i])]
            else:
                out = [t.clone() for t in graphed(m, xs[i], ys[i])]
                ref = [t.clone() for t in eager(m, xs[i], ys[i])]
        res.append(diff(out, ref))
    print(f'autocast={autocast} channels_last={cl} {order}:', res)
