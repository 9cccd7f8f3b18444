"""Input definitions of the golden cases, shared by make_golden.py (which runs the reference
on them) and the tests (which rebuild the same inputs and compare against the stored
reference outputs).  Inputs are regenerated from numpy PCG64 streams; each fixture stores a
digest of the feature table so a drifted generator is caught rather than silently compared."""
from __future__ import annotations

import os

import numpy as np

import graphsage_b200.synth as synth

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # cfg-1: Cora topology, 1433 binary feats, MEAN, sup, normal (num_neg 100), b_sz 20
    'cora_mean_sup': dict(topo='cora', feats=('binary', 1433, 0.0127, 11), classes=7, label_seed=12, split_seed=13,
                          hidden=128, gcn=False, agg='MEAN', learn='sup', unsup_loss='normal', b_sz=20),
    'cora_gcn_margin': dict(topo='cora', feats=('binary', 1433, 0.0127, 11), classes=7, label_seed=12, split_seed=13,
                            hidden=128, gcn=True, agg='MEAN', learn='plus_unsup', unsup_loss='margin', b_sz=20),
    'cora_max_plus': dict(topo='cora', feats=('normal', 64, 14), classes=7, label_seed=12, split_seed=13,
                          hidden=32, gcn=False, agg='MAX', learn='plus_unsup', unsup_loss='margin', b_sz=20),
    'cora_3layer_gcn_max': dict(topo='cora', feats=('normal', 64, 14), classes=7, label_seed=12, split_seed=13,
                                hidden=32, gcn=True, agg='MAX', learn='sup', unsup_loss='margin', b_sz=8, num_layers=3),
    # cfg-2: Pubmed topology, 500 TF-IDF-like feats, MAX, unsup, normal, b_sz 20
    'pubmed_max_unsup': dict(topo='pubmed', feats=('sparse', 500, 0.1, 21), classes=3, label_seed=22, split_seed=23,
                             hidden=128, gcn=False, agg='MAX', learn='unsup', unsup_loss='normal', b_sz=20),
    # Pubmed's three self-loop nodes in one batch, both gcn settings
    'pubmed_selfloop_mean': dict(topo='pubmed', feats=('normal', 48, 24), classes=3, label_seed=22, split_seed=23,
                                 hidden=32, gcn=False, agg='MEAN', learn='sup', unsup_loss='normal', b_sz=16,
                                 seeds='selfloops', extend=False),
    'pubmed_selfloop_gcn': dict(topo='pubmed', feats=('normal', 48, 24), classes=3, label_seed=22, split_seed=23,
                                hidden=32, gcn=True, agg='MEAN', learn='sup', unsup_loss='normal', b_sz=16,
                                seeds='selfloops', extend=False),
}

WEIGHT_SEED = 7


def load_topology(name):
    z = np.load(os.path.join(HERE, f'{name}_topology.npz'))
    return z['rowptr'], z['col']


def self_loop_nodes(rowptr, col):
    row = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
    return sorted(set(row[row == col].tolist()))


def build_inputs(name, topology=None):
    """-> dict(rowptr, col, feats, labels, train, weights[list], cls_w, cls_b, seeds, spec)."""
    spec = dict(CASES[name])
    rowptr, col = topology if topology is not None else load_topology(spec['topo'])
    n = len(rowptr) - 1
    kind = spec['feats'][0]
    if kind == 'binary':
        feats = synth.features_binary(n, spec['feats'][1], spec['feats'][2], seed=spec['feats'][3])
    elif kind == 'sparse':
        feats = synth.features_sparse_float(n, spec['feats'][1], spec['feats'][2], seed=spec['feats'][3])
    else:
        feats = synth.features_normal(n, spec['feats'][1], seed=spec['feats'][2])
    labels = synth.labels_uniform(n, spec['classes'], seed=spec['label_seed'])
    _, _, train = synth.split_nodes(n, seed=spec['split_seed'])
    num_layers = spec.get('num_layers', 2)
    hidden, gcn, f = spec['hidden'], spec['gcn'], feats.shape[1]
    wrng = np.random.default_rng(WEIGHT_SEED)
    weights = []
    for layer in range(num_layers):
        in_size = f if layer == 0 else hidden
        weights.append(synth.xavier_uniform_np(wrng, hidden, in_size if gcn else 2 * in_size))
    cls_w = synth.xavier_uniform_np(wrng, spec['classes'], hidden)
    cls_b = wrng.uniform(-0.05, 0.05, size=(spec['classes'],)).astype(np.float32)
    if spec.get('seeds') == 'selfloops':
        loops = self_loop_nodes(rowptr, col)
        seeds = np.asarray(loops + [int(x) for x in train[:spec['b_sz'] - len(loops)]], dtype=np.int64)
    else:
        seeds = train[:spec['b_sz']]
    spec['num_layers'] = num_layers
    spec['num_neg'] = 6 if spec['unsup_loss'] == 'margin' else 100        # src/utils.py:119-122
    spec.setdefault('extend', True)
    return dict(rowptr=rowptr, col=col, feats=feats, labels=labels, train=train, weights=weights, cls_w=cls_w,
                cls_b=cls_b, seeds=seeds, spec=spec)


def load_fixture(name):
    """-> (inputs dict, fixture dict with recorded calls and reference outputs)."""
    inp = build_inputs(name)
    z = np.load(os.path.join(HERE, f'{name}.npz'))
    fx = {k: z[k] for k in z.files}
    assert str(fx['feats_digest']) == synth.digest(inp['feats']), 'synthetic feature generator drifted'
    calls = []
    c = 0
    while f'call{c}_nodes' in fx:
        ptr, colv = fx[f'call{c}_ptr'], fx[f'call{c}_col']
        samp = [set(colv[ptr[i]:ptr[i + 1]].tolist()) for i in range(len(ptr) - 1)]
        calls.append((fx[f'call{c}_nodes'].tolist(), samp, fx[f'call{c}_uniq'].tolist()))
        c += 1
    fx['calls'] = calls
    return inp, fx
