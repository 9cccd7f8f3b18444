"""Golden vectors for N4 (src/utils.py:80-111, train_classification)  --  run in the BUILD container only.

Runs the REFERENCE's own `train_classification` loop (imported from /root/reference/src/utils.py) for 3 epochs on
synthetic frozen embeddings.  Its two callees that are not part of the loop body are replaced by stubs: the
embeddings come from a fixed array instead of `get_gnn_embeddings` (:88) and `evaluate` (:109) is a no-op; `shuffle`
(:91) is wrapped to record the order it returned.  Everything between -- batching, classifier forward, NLL, backward,
clip_grad_norm_, SGD -- is the reference's code on the reference's Classification module.  Output:
tests/golden/train_classification.npz; the oracle restatement is checked against it here (must be identical).

    python tests/golden/make_golden_train_classification.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness, sage_oracle as so      # noqa: E402


class _DC:          # the attribute surface train_classification reads (:86-87)
    pass


def main():
    utils = ref_harness.load_reference_utils()
    models = ref_harness.load_reference_models()
    rng = np.random.default_rng(17)
    n, dim, classes, epochs = 300, 128, 7, 3
    feats = torch.from_numpy(np.maximum(rng.standard_normal((n, dim)), 0).astype(np.float32))     # ReLU outputs
    labels = rng.integers(0, classes, size=n).astype(np.int64)
    train = rng.permutation(n)[:173].astype(np.int64)                  # 173 = 3 x 50 + 23: a ragged last batch
    torch.manual_seed(824)
    cls = models.Classification(dim, classes)
    w0 = cls.layer[0].weight.detach().clone()
    b0 = cls.layer[0].bias.detach().clone()
    dc = _DC()
    dc.x_train, dc.x_labels = train, labels
    orders = []
    orig_shuffle = utils.shuffle

    def recording_shuffle(a):
        out = orig_shuffle(a)
        orders.append(np.asarray(out).copy())
        return out

    utils.shuffle = recording_shuffle
    utils.get_gnn_embeddings = lambda gnn, d, ds: feats
    utils.evaluate = lambda *a, **k: 0.0
    np.random.seed(824)
    try:
        cls, _ = utils.train_classification(dc, None, cls, "x", torch.device("cpu"), 0.0, "golden", epochs=epochs)
    finally:
        utils.shuffle = orig_shuffle
    w1, b1 = cls.layer[0].weight.detach(), cls.layer[0].bias.detach()
    assert len(orders) == epochs
    ow, ob, _ = so.train_classification(w0, b0, feats, labels, orders)
    dw, db = float((ow - w1).abs().max()), float((ob - b1).abs().max())
    print("oracle vs reference: max |dW| =", dw, " max |db| =", db)
    assert dw == 0.0 and db == 0.0, "oracle restatement differs from the reference loop"
    np.savez_compressed(os.path.join(HERE, "train_classification.npz"), feats=feats.numpy(), labels=labels, train=train,
                        w0=w0.numpy(), b0=b0.numpy(), w1=w1.numpy(), b1=b1.numpy(), orders=np.stack(orders))
    print("moved:", float((w1 - w0).abs().max()))


if __name__ == "__main__":
    main()
