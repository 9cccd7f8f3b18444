"""GPU: the reference's OWN train / evaluation loops, imported unchanged from the staged copy under
baseline/_ref/ (`python baseline/make_ref.py`; git-ignored, ships with gpurun), executed against the drop-in
classes of this package -- "the drop-in test" of SURVEY.md §1:

    src/utils.py:113-193  apply_model(dataCenter, ds, graphSage, classification, unsupervised_loss, b_sz, ...)
    src/utils.py:13-57    evaluate(dataCenter, ds, graphSage, classification, device, max_vali_f1, name, epoch)

Samplers on both sides are random with different streams, so the comparison is functional: the same loop, with
the same seeds for everything the loop itself draws (sklearn shuffle), trains (a) the reference's classes on the
CPU and (b) the drop-in on the GPU on a task with learnable labels; both must reach the same validation
micro-F1 within a few points, far above chance, and `evaluate` must pickle the live drop-in modules as it
pickles the reference's (src/utils.py:52)."""
import glob
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'baseline'))
import make_ref  # noqa: E402

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not make_ref.available(),
                               reason='baseline/_ref not staged (python baseline/make_ref.py in the build container)')


def _task(n_train=400):
    """Cora topology, 64 normal features, 7 classes that ARE learnable: label = argmax of the first 7 columns."""
    inp = cases.build_inputs('cora_max_plus')
    feats = inp['feats']
    labels = np.argmax(feats[:, :7], axis=1).astype(np.int64)
    n = feats.shape[0]
    perm = np.random.default_rng(3).permutation(n)
    dc = types.SimpleNamespace(cora_train=perm[:n_train].astype(np.int64), cora_val=perm[n_train:n_train + 300].astype(np.int64),
                               cora_test=perm[n_train + 300:n_train + 600].astype(np.int64), cora_labels=labels)
    return inp, feats, labels, dc


def _seed_all(s=824):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


@needs_ref
@pytest.mark.parametrize('learn_method,unsup_loss', [('sup', 'margin'), ('plus_unsup', 'margin'), ('plus_unsup', 'normal')])
def test_reference_apply_model_and_evaluate_run_unchanged_against_the_drop_in(tmp_path, monkeypatch, capsys, learn_method,
                                                                             unsup_loss):
    assert make_ref.verify(), 'baseline/_ref differs from what make_ref.py staged'
    ref_utils = make_ref.load('utils')
    ref_models = make_ref.load('models')
    import graphsage_b200  # noqa: F401
    from graphsage_b200 import models as our_models
    from oracle import sage_oracle as so
    monkeypatch.chdir(tmp_path)
    os.makedirs('models')
    inp, feats, labels, dc = _task(n_train=1000 if unsup_loss == 'normal' else 600)
    adj = so.csr_to_adj_dict(inp['rowptr'], inp['col'])          # the defaultdict(set) of src/dataCenter.py:33
    epochs, b_sz = (2 if unsup_loss == 'normal' else 3), 20
    f1 = {}
    for side, M, dev in (('reference', ref_models, torch.device('cpu')), ('drop-in', our_models, torch.device('cuda:0'))):
        _seed_all()
        features = torch.from_numpy(feats).to(dev)                                            # main.py:52
        graphSage = M.GraphSage(2, features.size(1), 128, features, adj, dev, gcn=False, agg_func='MEAN')
        graphSage.to(dev)                                                                     # main.py:54-55
        classification = M.Classification(128, 7)
        classification.to(dev)                                                                # main.py:58-59
        unsupervised_loss = M.UnsupervisedLoss(adj, dc.cora_train, dev)                       # main.py:61
        best = 0
        for epoch in range(epochs):                                                           # main.py:69-75
            graphSage, classification = ref_utils.apply_model(dc, 'cora', graphSage, classification, unsupervised_loss,
                                                              b_sz, unsup_loss, dev, learn_method)
            best = ref_utils.evaluate(dc, 'cora', graphSage, classification, dev, best, side, epoch)
        f1[side] = float(best)
        saved = glob.glob(f'models/model_best_{side}_ep*.torch')
        assert saved, 'evaluate() must have pickled the live modules (src/utils.py:52)'
        if side == 'drop-in':
            from graphsage_b200 import native
            assert native.launch_count() > 0                     # the loop really ran on the CUDA library
            back = torch.load(saved[-1], weights_only=False)
            assert [type(m).__name__ for m in back] == ['GraphSage', 'Classification']
            assert back[0].sage_layer1.weight.shape == graphSage.sage_layer1.weight.shape
            assert back[0].raw_features.shape == features.shape   # the pickle carries what the reference's does
    out = capsys.readouterr().out
    assert 'Step [1/' in out and 'Validation F1:' in out          # the reference's own prints, from its own loop
    if unsup_loss == 'margin':
        assert f1['reference'] > 0.45 and f1['drop-in'] > 0.45, f1    # chance is 1/7; the reference reaches ~0.6 here
        assert abs(f1['reference'] - f1['drop-in']) < 0.12, f1
    else:       # Q = 10 times the pair loss swamps the classifier in two epochs: both sides stay near 0.2-0.3
        assert abs(f1['reference'] - f1['drop-in']) < 0.15, f1
