"""Generate the golden fixtures under tests/golden/ by running the *reference itself*.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it (1) builds the inputs from `graphsage_b200.synth` (numpy PCG64 streams,
regenerable anywhere), (2) runs the unmodified reference classes imported from
/root/reference/src/models.py with `random.seed(824)` (the reference default,
src/main.py:18), recording every `_get_unique_neighs_list` call, (3) runs
`oracle/sage_oracle.py` under the same seed and asserts it reproduces the reference's
samples exactly and its tensors to fp32 round-off, then (4) writes a compressed `.npz`
with the recorded samples and the reference's outputs.  Topologies parsed from the
reference's shipped edge lists are stored as CSR (`*_topology.npz`).
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness, sage_oracle as so          # noqa: E402
import graphsage_b200.synth as synth                        # noqa: E402
sys.path.insert(0, HERE)
import cases                                                # noqa: E402

SEED = 824


def pack_sets(rows):
    """list of sets -> (ptr, col) in each set's own iteration order."""
    ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    col = []
    for i, s in enumerate(rows):
        col.extend(int(x) for x in s)
        ptr[i + 1] = len(col)
    return ptr, np.asarray(col, dtype=np.int64)


def build_case(name, topology):
    ref = ref_harness.load_reference_models()
    inp = cases.build_inputs(name, topology)
    spec = inp['spec']
    rowptr, col, feats, labels, train_nodes = inp['rowptr'], inp['col'], inp['feats'], inp['labels'], inp['train']
    weights, cls_w, cls_b, seeds = inp['weights'], inp['cls_w'], inp['cls_b'], inp['seeds']
    hidden, classes, gcn, agg, learn = spec['hidden'], spec['classes'], spec['gcn'], spec['agg'], spec['learn']
    unsup_loss, num_layers, num_neg, extend = spec['unsup_loss'], spec['num_layers'], spec['num_neg'], spec['extend']
    weight_seed = cases.WEIGHT_SEED
    adj = so.csr_to_adj_dict(rowptr, col)
    n, f = feats.shape
    feats_t = torch.from_numpy(feats)

    # ---------------- the reference, unmodified ----------------
    random.seed(SEED)
    model = ref.GraphSage(num_layers, f, hidden, feats_t, adj, torch.device('cpu'), gcn=gcn, agg_func=agg)
    cls = ref.Classification(hidden, classes)
    with torch.no_grad():
        for layer in range(num_layers):
            getattr(model, f'sage_layer{layer + 1}').weight.copy_(torch.from_numpy(weights[layer]))
        cls.layer[0].weight.copy_(torch.from_numpy(cls_w))
        cls.layer[0].bias.copy_(torch.from_numpy(cls_b))
    unsup = ref.UnsupervisedLoss(adj, train_nodes, torch.device('cpu'))
    if extend:
        batch = np.asarray(list(unsup.extend_nodes(seeds, num_neg=num_neg)))          # src/utils.py:149
    else:
        batch = np.asarray(seeds)
    with ref_harness.SampleRecorder(model) as rec:
        embs = model(batch)                                                           # src/utils.py:157
    logp = cls(embs)
    loss_sup = -torch.sum(logp[range(logp.size(0)), labels[batch]], 0) / len(batch)   # src/utils.py:162-163
    loss_net = None
    if learn != 'sup':
        loss_net = unsup.get_loss_margin(embs, batch) if unsup_loss == 'margin' else unsup.get_loss_sage(embs, batch)
    loss = {'sup': loss_sup, 'plus_unsup': None, 'unsup': loss_net}[learn]
    if learn == 'plus_unsup':
        loss = loss_sup + loss_net
    loss.backward()
    ref_out = dict(embs=embs.detach().numpy(), logp=logp.detach().numpy(), loss=loss.detach().numpy().reshape(-1),
                   loss_sup=loss_sup.detach().numpy().reshape(-1))
    if loss_net is not None:
        ref_out['loss_net'] = loss_net.detach().numpy().reshape(-1)
    for layer in range(num_layers):
        ref_out[f'grad_w{layer + 1}'] = getattr(model, f'sage_layer{layer + 1}').weight.grad.numpy()
    if learn != 'unsup':
        ref_out['grad_cls_w'] = cls.layer[0].weight.grad.numpy()
        ref_out['grad_cls_b'] = cls.layer[0].bias.grad.numpy()

    # ---------------- the oracle under the same seed: must agree ----------------
    random.seed(SEED)
    w_t = [torch.from_numpy(w.copy()).requires_grad_(True) for w in weights]
    cw_t = torch.from_numpy(cls_w.copy()).requires_grad_(True)
    cb_t = torch.from_numpy(cls_b.copy()).requires_grad_(True)
    pairs = so.PairSampler(adj, train_nodes)
    if extend:
        o_batch = np.asarray(list(pairs.extend_nodes(seeds, num_neg=num_neg)))
        assert np.array_equal(o_batch, batch), "oracle extend_nodes diverged from the reference"
        assert pairs.positive_pairs == unsup.positive_pairs and pairs.negtive_pairs == unsup.negtive_pairs
    o_rec = []
    o_embs = so.graphsage_forward(w_t, feats_t, adj, batch, gcn, agg, record=o_rec)
    assert len(o_rec) == len(rec.calls)
    for (n1, s1, u1), (n2, s2, u2) in zip(o_rec, rec.calls):
        assert list(n1) == list(n2) and s1 == s2 and u1 == u2, "oracle sampling diverged from the reference"
    o_logp = so.classification(cw_t, cb_t, o_embs)
    o_sup = so.supervised_loss(o_logp, labels[batch])
    o_net = None
    if learn != 'sup':
        o_net = so.loss_margin(pairs, o_embs, batch) if unsup_loss == 'margin' else so.loss_sage(pairs, o_embs, batch)
    o_loss = o_sup if learn == 'sup' else (o_net if learn == 'unsup' else o_sup + o_net)
    o_loss.backward()

    def close(a, b, what):
        a, b = np.asarray(a), np.asarray(b)
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
        assert err <= 1e-6, f"{name}: oracle vs reference {what}: {err:.3e}"
        return err

    errs = dict(embs=close(o_embs.detach().numpy(), ref_out['embs'], 'embs'),
                logp=close(o_logp.detach().numpy(), ref_out['logp'], 'logp'),
                loss=close(o_loss.detach().numpy().reshape(-1), ref_out['loss'], 'loss'))
    for layer in range(num_layers):
        errs[f'grad_w{layer + 1}'] = close(w_t[layer].grad.numpy(), ref_out[f'grad_w{layer + 1}'], f'grad_w{layer + 1}')
    if learn != 'unsup':
        errs['grad_cls_w'] = close(cw_t.grad.numpy(), ref_out['grad_cls_w'], 'grad_cls_w')
    # replay through the injection seam as well
    replay = so.graphsage_forward([w.detach() for w in w_t], feats_t, adj, batch, gcn, agg,
                                  injected=[(c[1], c[2]) for c in rec.calls])
    errs['replay'] = close(replay.numpy(), ref_out['embs'], 'replayed embs')

    # ---------------- fixture ----------------
    out = dict(batch=batch.astype(np.int64), seeds=np.asarray(seeds, dtype=np.int64),
               meta=np.asarray([num_layers, hidden, classes, int(gcn), weight_seed, num_neg], dtype=np.int64),
               agg=np.asarray(agg), learn=np.asarray(learn), unsup_loss=np.asarray(unsup_loss),
               feats_digest=np.asarray(synth.digest(feats)), cls_b=cls_b)
    for c, (nodes, samp, uniq) in enumerate(rec.calls):
        ptr, cc = pack_sets(samp)
        out[f'call{c}_nodes'] = np.asarray(nodes, dtype=np.int64)
        out[f'call{c}_ptr'] = ptr
        out[f'call{c}_col'] = cc
        out[f'call{c}_uniq'] = np.asarray(uniq, dtype=np.int64)
    if extend:
        out['pos_pairs'] = np.asarray(unsup.positive_pairs, dtype=np.int64).reshape(-1, 2)
        out['neg_pairs'] = np.asarray(unsup.negtive_pairs, dtype=np.int64).reshape(-1, 2)
        out['pos_nodes'] = np.asarray(list(unsup.node_positive_pairs.keys()), dtype=np.int64)
        out['neg_nodes'] = np.asarray(list(unsup.node_negtive_pairs.keys()), dtype=np.int64)
    out.update({f'ref_{k}': v for k, v in ref_out.items()})
    path = os.path.join(HERE, f'{name}.npz')
    np.savez_compressed(path, **out)
    print(f"{name}: |B|={len(batch)} calls={[len(c[0]) for c in rec.calls]} uniq={[len(c[2]) for c in rec.calls]} "
          f"loss={float(ref_out['loss'][0]):.6f} oracle-vs-ref max rel err={max(errs.values()):.2e} "
          f"-> {os.path.getsize(path) / 1e6:.2f} MB")


def main():
    torch.manual_seed(SEED)
    np.random.seed(SEED)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    topo = {'cora': ref_harness.cora_topology(), 'pubmed': ref_harness.pubmed_topology()}
    for k, (rowptr, col) in topo.items():
        np.savez_compressed(os.path.join(HERE, f'{k}_topology.npz'), rowptr=rowptr, col=col)
        print(k, len(rowptr) - 1, len(col), 'self-loops', cases.self_loop_nodes(rowptr, col))
    for name, spec in cases.CASES.items():
        build_case(name, topo[spec['topo']])


if __name__ == '__main__':
    main()
