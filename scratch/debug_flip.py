import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import numpy as np, torch
import cases
import graphsage_b200
from graphsage_b200 import models as M
from graphsage_b200.graph import AdjCSR
dev = torch.device('cuda:0')
name = sys.argv[1] if len(sys.argv) > 1 else 'pubmed_max_unsup'
inp, fx = cases.load_fixture(name)
spec = inp['spec']
def rel(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
res = {}
for prec in ('fp32', 'tf32x3'):
    feats = torch.from_numpy(inp['feats']).to(dev)
    adj = AdjCSR(inp['rowptr'], inp['col'])
    model = M.GraphSage(spec['num_layers'], feats.shape[1], spec['hidden'], feats, adj, dev, gcn=spec['gcn'], agg_func=spec['agg'], seed=1, precision=prec).to(dev)
    with torch.no_grad():
        for i, w in enumerate(inp['weights']):
            getattr(model, f'sage_layer{i + 1}').weight.copy_(torch.from_numpy(w))
    batch = fx['batch']
    model.inject_samples([(c[0], c[1]) for c in fx['calls']])
    embs = model(batch)
    unsup = M.UnsupervisedLoss(adj, inp['train'], dev)
    npos = {int(n): [] for n in fx['pos_nodes']}; nneg = {int(n): [] for n in fx['neg_nodes']}
    for a, b in fx['pos_pairs']: npos[int(a)].append((int(a), int(b)))
    for a, b in fx['neg_pairs']: nneg[int(a)].append((int(a), int(b)))
    unsup.set_pairs(batch.tolist(), inp['seeds'].tolist(), npos, nneg)
    net = unsup.get_loss_margin(embs, batch) if spec['unsup_loss'] == 'margin' else unsup.get_loss_sage(embs, batch)
    net.backward()
    L = model._last_layers
    res[prec] = dict(h1=L[0].h.clone(), h2=L[1].h.clone(), arg2=None if L[1].argmax is None else L[1].argmax.clone(),
                     gw1=model.sage_layer1.weight.grad.clone(), gw2=model.sage_layer2.weight.grad.clone(), live1=int(L[0].num_rows.item()))
    print(prec, 'embs', rel(embs, fx['ref_embs']), 'gw1', rel(res[prec]['gw1'], fx['ref_grad_w1']), 'gw2', rel(res[prec]['gw2'], fx['ref_grad_w2']))
a, b = res['fp32'], res['tf32x3']
n1 = a['live1']
print('h1 mask diff', int(((a['h1'][:n1] > 0) != (b['h1'][:n1] > 0)).sum()), 'of', a['h1'][:n1].numel())
print('h2 mask diff', int(((a['h2'] > 0) != (b['h2'] > 0)).sum()))
if a['arg2'] is not None:
    print('argmax2 diff', int((a['arg2'] != b['arg2']).sum()), 'of', a['arg2'].numel())
d = (a['gw1'] - b['gw1']).abs()
print('gw1 diff max', float(d.max()), 'at', np.unravel_index(int(d.argmax()), d.shape), 'max|gw1|', float(a['gw1'].abs().max()), 'rows affected', int((d.max(1)[0] > 1e-6 * a['gw1'].abs().max()).sum()))
