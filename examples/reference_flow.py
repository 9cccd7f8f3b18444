"""The reference's `src/main.py` flow, end to end, on the B200 path (needs a GPU):

    text files --datacache.DataCenter--> CSR + features --GraphSage / Classification--> supervised steps
    (SupervisedTrainer, one CUDA-graph replay each) --> inference.evaluate (src/utils.py:13-57) -->
    inference.get_gnn_embeddings + train_classification on the frozen embeddings (src/utils.py:59-111)

    python examples/reference_flow.py [dir with cora.content / cora.cites]      (default: tests/golden/datacenter)

Nothing here is a benchmark; it shows that a user of the reference finds the same pieces under the same names.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import graphsage_b200  # noqa: E402,F401
from graphsage_b200 import datacache, inference  # noqa: E402
from graphsage_b200.models import Classification, GraphSage  # noqa: E402
from graphsage_b200.trainer import SupervisedTrainer  # noqa: E402


def main(data_dir: str | None = None, epochs: int = 2, b_sz: int = 8, seed: int = 824, log=print) -> dict:
    data_dir = data_dir or os.path.join(ROOT, "tests", "golden", "datacenter")
    device = torch.device("cuda:0")
    np.random.seed(seed)                                                    # main.py:41
    torch.manual_seed(seed)                                                 # main.py:42
    config = {"file_path.cora_content": os.path.join(data_dir, "cora.content"),
              "file_path.cora_cite": os.path.join(data_dir, "cora.cites")}
    ds = "cora"
    dataCenter = datacache.DataCenter(config)                               # main.py:47-48
    dataCenter.load_dataSet(ds)
    features = torch.from_numpy(np.asarray(getattr(dataCenter, ds + "_feats"))).to(device)       # main.py:52
    labels = getattr(dataCenter, ds + "_labels")
    train, val, test = (getattr(dataCenter, f"{ds}_{part}") for part in ("train", "val", "test"))
    graphSage = GraphSage(2, features.size(1), 128, features, getattr(dataCenter, ds + "_adj_lists"), device,
                          gcn=False, agg_func="MEAN", seed=seed).to(device)                     # main.py:54-55
    num_labels = len(set(labels.tolist()))                                                      # main.py:57
    classification = Classification(128, num_labels).to(device)                                 # main.py:58-59
    trainer = SupervisedTrainer(graphSage, classification, labels, b_sz)
    max_vali_f1, losses = 0.0, []
    for epoch in range(epochs):                                                                 # main.py:69-75
        order = np.random.permutation(train)                                                    # utils.py:127
        for lo in range(0, len(order) - b_sz + 1, b_sz):                                        # full batches: the captured step is static
            losses.append(float(trainer.step(order[lo:lo + b_sz]).item()))
        vali_f1, test_f1, max_vali_f1 = inference.evaluate(val, test, labels, graphSage, classification, max_vali_f1)
        log(f"epoch {epoch}: loss {losses[-1]:.4f}  validation F1 {vali_f1:.4f}" +
            (f"  test F1 {test_f1:.4f}" if test_f1 is not None else ""))
    embeddings = inference.get_gnn_embeddings(graphSage)                                        # utils.py:88
    head = Classification(128, num_labels).to(device)
    inference.train_classification(embeddings, train, labels, head, epochs=3, b_sz=b_sz)        # utils.py:80-111
    pred = torch.argmax(head(embeddings), dim=1)
    train_acc = inference.micro_f1(torch.from_numpy(labels[train]).to(device), pred[torch.from_numpy(train).to(device)])
    log(f"classifier on frozen embeddings: train micro-F1 {train_acc:.4f}")
    return dict(losses=losses, max_vali_f1=max_vali_f1, embeddings=embeddings, train_acc=train_acc,
                num_nodes=int(features.shape[0]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
