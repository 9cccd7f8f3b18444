"""Forward-only use of the hot path: the reference's `get_gnn_embeddings` and the metric half of
`evaluate` (src/utils.py:59-78 and :13-57) on the device.

SURVEY.md §8(f) row N2.  The reference embeds all nodes 500 at a time through `gnn_model(nodes_batch)`
(utils.py:63-71) and scores the validation / test split with sklearn's micro-F1 of the arg-max class
(utils.py:26-47), which for single-label classification is the accuracy.  Here the same loop runs without
autograd, in batches as large as the caller likes (the sampler is per seed, so the batch size changes only
how many seeds share one launch), writes straight into one preallocated [N x out_size] tensor and never
leaves the device; the per-batch work is the same kernels as a training forward: sampler ->
unique/remap -> aggregation -> SageLayer GEMM.  At b_sz 8192 on the cfg-3 graph the layer-1 aggregation
launch gathers ~390 MB and runs at 0.8 of the HBM copy peak (bench.py --workload infer).

Out of scope, as in SURVEY.md §2: `torch.save` of the live modules (utils.py:52) and the console prints.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import native
from .models import Classification, GraphSage, _as_device_ids

__all__ = ["get_gnn_embeddings", "predict", "micro_f1", "evaluate"]


def _weights(gnn_model: GraphSage):
    ws = [getattr(gnn_model, f"sage_layer{i}").weight.detach() for i in range(1, gnn_model.num_layers + 1)]
    for w in ws:
        native.require_cuda(w, "GraphSage weights")
    return ws


@torch.no_grad()
def get_gnn_embeddings(gnn_model: GraphSage, nodes=None, b_sz: int = 500, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Embeddings of `nodes` (all nodes of the graph when None, like src/utils.py:62), `b_sz` per forward
    (reference: 500, utils.py:63).  Returns a detached fp32 [len(nodes) x out_size] tensor on the device,
    row i <-> nodes[i].  Neighbour sampling is active, as it is in the reference's inference."""
    csr, _, dev = gnn_model._state()
    if nodes is None:
        nodes_dev = torch.arange(csr.num_nodes, dtype=torch.int32, device=dev)
    else:
        nodes_dev = _as_device_ids(nodes, dev)
    n, width = int(nodes_dev.shape[0]), gnn_model.out_size
    if b_sz < 1:
        raise ValueError("b_sz must be positive")
    if out is None:
        out = torch.empty((n, width), dtype=torch.float32, device=dev)
    elif out.shape != (n, width) or out.dtype != torch.float32 or out.device != dev:
        raise ValueError(f"out must be a float32 [{n} x {width}] tensor on {dev}")
    weights = _weights(gnn_model)
    for lo in range(0, n, b_sz):
        batch = nodes_dev[lo:lo + b_sz]
        layers = gnn_model._run_forward(batch, weights, None)
        out[lo:lo + batch.shape[0]].copy_(layers[-1].h[:batch.shape[0], :width])
    return out


@torch.no_grad()
def predict(gnn_model: GraphSage, classification: Classification, nodes, b_sz: int = 8192) -> torch.Tensor:
    """arg-max class of `classification(gnn_model(nodes))` (src/utils.py:26-28), int64 on the device."""
    embs = get_gnn_embeddings(gnn_model, nodes, b_sz=b_sz)
    return torch.argmax(classification(embs), dim=1)


def micro_f1(labels_true: torch.Tensor, predicted: torch.Tensor) -> float:
    """sklearn.metrics.f1_score(average='micro') for single-label multi-class input (src/utils.py:32,45):
    TP, FP and FN are summed over classes, every miss is one FP and one FN, so F1 = accuracy."""
    if labels_true.shape != predicted.shape:
        raise ValueError("labels and predictions differ in length")        # the reference asserts the same (utils.py:30)
    if labels_true.numel() == 0:
        return 0.0
    return float((labels_true.to(predicted.device) == predicted).double().mean().item())


def evaluate(val_nodes, test_nodes, labels, gnn_model: GraphSage, classification: Classification,
             max_vali_f1: float = 0.0, b_sz: int = 8192) -> Tuple[float, Optional[float], float]:
    """The metric half of src/utils.py:13-57: validation micro-F1; when it beats `max_vali_f1` also the test
    micro-F1 (the reference then saves the modules, which stays with the caller).  Parameters keep their
    requires_grad flags: nothing here records autograd state.  Returns (vali_f1, test_f1 or None,
    new max_vali_f1)."""
    dev = gnn_model._state()[2]
    labels_dev = labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.asarray(labels, dtype=np.int64))
    labels_dev = labels_dev.to(dev)

    def score(nodes):
        ids = _as_device_ids(nodes, dev)
        return micro_f1(labels_dev[ids.long()], predict(gnn_model, classification, ids, b_sz=b_sz))

    vali_f1 = score(val_nodes)
    test_f1 = None
    if vali_f1 > max_vali_f1:
        max_vali_f1 = vali_f1
        test_f1 = score(test_nodes)
    return vali_f1, test_f1, max_vali_f1
