"""CPU: the oracle (oracle/sage_oracle.py) against the committed golden fixtures, which
hold outputs of the reference itself (tests/golden/make_golden.py)."""
import random

import numpy as np
import pytest
import torch

import cases
from oracle import sage_oracle as so

TOL = 1e-5   # norm-relative, fp32; same-machine agreement is exact


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize('name', list(cases.CASES))
def test_oracle_reproduces_reference_outputs(name):
    inp, fx = cases.load_fixture(name)
    spec = inp['spec']
    adj = so.LazySetAdjacency(inp['rowptr'], inp['col'])
    w = [torch.from_numpy(x.copy()).requires_grad_(True) for x in inp['weights']]
    cw = torch.from_numpy(inp['cls_w'].copy()).requires_grad_(True)
    cb = torch.from_numpy(inp['cls_b'].copy()).requires_grad_(True)
    feats = torch.from_numpy(inp['feats'])
    batch = fx['batch']
    embs = so.graphsage_forward(w, feats, adj, batch, spec['gcn'], spec['agg'], injected=[(c[1], c[2]) for c in fx['calls']])
    assert rel(embs.detach().numpy(), fx['ref_embs']) <= TOL
    logp = so.classification(cw, cb, embs)
    assert rel(logp.detach().numpy(), fx['ref_logp']) <= TOL
    loss = so.supervised_loss(logp, inp['labels'][batch])
    assert rel(loss.detach().numpy().reshape(-1), fx['ref_loss_sup']) <= TOL
    if spec['learn'] == 'sup':
        loss.backward()
        for layer in range(spec['num_layers']):
            assert rel(w[layer].grad.numpy(), fx[f'ref_grad_w{layer + 1}']) <= TOL
        assert rel(cw.grad.numpy(), fx['ref_grad_cls_w']) <= TOL
        assert rel(cb.grad.numpy(), fx['ref_grad_cls_b']) <= TOL


@pytest.mark.parametrize('name', ['cora_gcn_margin', 'cora_max_plus', 'pubmed_max_unsup'])
def test_oracle_unsup_losses_and_grads(name):
    inp, fx = cases.load_fixture(name)
    spec = inp['spec']
    adj = so.LazySetAdjacency(inp['rowptr'], inp['col'])
    # same python RNG stream as the reference run => same pairs, same samples
    random.seed(824)
    pairs = so.PairSampler(adj, inp['train'])
    batch = np.asarray(list(pairs.extend_nodes(inp['seeds'], num_neg=spec['num_neg'])))
    assert np.array_equal(batch, fx['batch'])
    assert np.array_equal(np.asarray(pairs.positive_pairs).reshape(-1, 2), fx['pos_pairs'])
    assert np.array_equal(np.asarray(pairs.negtive_pairs).reshape(-1, 2), fx['neg_pairs'])
    w = [torch.from_numpy(x.copy()).requires_grad_(True) for x in inp['weights']]
    cw = torch.from_numpy(inp['cls_w'].copy()).requires_grad_(True)
    cb = torch.from_numpy(inp['cls_b'].copy()).requires_grad_(True)
    rec = []
    embs = so.graphsage_forward(w, torch.from_numpy(inp['feats']), adj, batch, spec['gcn'], spec['agg'], record=rec)
    for (nodes, samp, uniq), (n2, s2, u2) in zip(rec, fx['calls']):
        assert list(nodes) == list(n2) and samp == s2 and uniq == u2     # sampling stream identical
    net = so.loss_margin(pairs, embs, batch) if spec['unsup_loss'] == 'margin' else so.loss_sage(pairs, embs, batch)
    assert rel(net.detach().numpy().reshape(-1), fx['ref_loss_net']) <= TOL
    loss = net
    if spec['learn'] == 'plus_unsup':
        loss = so.supervised_loss(so.classification(cw, cb, embs), inp['labels'][batch]) + net
    loss.backward()
    for layer in range(spec['num_layers']):
        assert rel(w[layer].grad.numpy(), fx[f'ref_grad_w{layer + 1}']) <= TOL


def test_canonical_unique_remap_matches_reference_sets():
    inp, fx = cases.load_fixture('pubmed_selfloop_mean')
    for nodes, samp, uniq in fx['calls']:
        for drop_self in (True, False):
            U, self_idx, cols, cnt = so.canonical_unique_remap(nodes, samp, drop_self)
            assert set(U.tolist()) == set(uniq) and np.all(np.diff(U) > 0)
            assert np.array_equal(U[self_idx], np.asarray(nodes))
            for r, s in enumerate(samp):
                want = sorted(x for x in s if not (drop_self and x == nodes[r]))
                assert U[cols[r, :cnt[r]]].tolist() == want
                assert np.all(cols[r, cnt[r]:] == -1)


def test_sampler_semantics():
    rowptr, col = cases.load_topology('cora')
    adj = so.LazySetAdjacency(rowptr, col)
    deg = np.diff(rowptr)
    nodes = list(range(0, 2708, 7))
    samp = so.sample_neighbors(adj, nodes, 10, random.Random(1))
    for n, s in zip(nodes, samp):
        nb = adj[n]
        assert n in s
        got = s - {n}
        assert got <= nb
        assert len(got) == (len(nb) if deg[n] < 10 else 10)               # src/models.py:282


def test_train_classification_oracle_matches_the_reference_loop():
    """N4: oracle restatement of src/utils.py:90-107 against the weights the reference's own loop produced
    (tests/golden/make_golden_train_classification.py asserts a difference of exactly 0 when it writes them)."""
    import os
    import numpy as np
    import torch
    from oracle import sage_oracle as so
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_classification.npz"))
    w, b, loss = so.train_classification(torch.from_numpy(g["w0"]), torch.from_numpy(g["b0"]), torch.from_numpy(g["feats"]),
                                         g["labels"], list(g["orders"]))
    assert float((w - torch.from_numpy(g["w1"])).abs().max()) <= 1e-7
    assert float((b - torch.from_numpy(g["b1"])).abs().max()) <= 1e-7
    assert np.isfinite(float(loss))
    assert sorted(g["orders"][0].tolist()) == sorted(g["train"].tolist())          # shuffle permutes the train nodes (:91)
