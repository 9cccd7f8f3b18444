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

`train_classification` (row N4, src/utils.py:80-111) trains the classifier on frozen embeddings: 800 epochs of
[50 x 128] . [128 x C] steps, a pure launch-latency workload.  One epoch (every batch: row gather, the one-launch
classifier + NLL + backward kernel, clip + SGD) is captured as ONE CUDA graph that reads the epoch's node order
from a device buffer, so an epoch costs one replay and one small H2D copy (or a device randperm).

Out of scope, as in SURVEY.md §2: `torch.save` of the live modules (utils.py:52) and the console prints.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import native, ops
from .models import Classification, GraphSage, _as_device_ids

__all__ = ["get_gnn_embeddings", "predict", "micro_f1", "evaluate", "train_classification"]


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


def train_classification(features: torch.Tensor, train_nodes, labels, classification: Classification, epochs: int = 800,
                         b_sz: int = 50, lr: float = 0.5, max_norm: float = 5.0, orders=None, on_epoch=None,
                         use_graph: bool = True, seed: int = 0) -> torch.Tensor:
    """src/utils.py:80-111 on the device.  `features`: frozen embeddings [N x emb] (what `get_gnn_embeddings`
    returned, :88); per epoch the train nodes are shuffled (:91), cut into batches of `b_sz` (:85, last one ragged),
    and every batch does classifier forward, NLL mean (:100-101), backward, clip_grad_norm_(5) (:106) and an SGD
    step with lr 0.5 (:82,107).  `orders` (optional): one node order per epoch, replacing the shuffle -- replaying
    the orders the reference's `shuffle` produced reproduces its weights (tests/golden/train_classification.npz).
    `on_epoch(epoch)` is called after every epoch (the reference evaluates there, :109).  Updates
    `classification` in place and returns the device loss of the last batch."""
    lin = classification.layer[0]
    native.require_cuda(lin.weight, "Classification parameters")
    dev = lin.weight.device
    native.require_cuda(features, "features")
    feats = features.detach()
    if feats.dtype != torch.float32 or feats.stride(1) != 1 or feats.stride(0) % 4:
        feats = feats.float().contiguous()
    dim, classes = int(lin.weight.shape[1]), int(lin.weight.shape[0])
    if feats.shape[1] != dim:
        raise ValueError(f"features have {feats.shape[1]} columns, the classifier expects {dim}")
    train_dev = _as_device_ids(train_nodes, dev)
    n_train = int(train_dev.shape[0])
    if n_train == 0 or epochs <= 0:
        return torch.zeros((1,), dtype=torch.float32, device=dev)
    labels_dev = labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.asarray(labels, dtype=np.int64))
    labels_dev = labels_dev.to(dev)
    if orders is not None:
        orders = [np.ascontiguousarray(np.asarray(o), dtype=np.int32) for o in orders]
        if len(orders) < epochs or any(len(o) != n_train for o in orders[:epochs]):
            raise ValueError("orders must hold one permutation of the train nodes per epoch")
    params = [lin.weight.data, lin.bias.data]
    grads = [torch.zeros_like(p) for p in params]
    tl = ops.TensorList(params, grads)
    order = torch.empty((n_train,), dtype=torch.int32, device=dev)      # this epoch's node order, read by the graph
    loss = torch.zeros((1,), dtype=torch.float32, device=dev)

    def epoch_body():
        for lo in range(0, n_train, b_sz):
            ids = order[lo:lo + b_sz]
            emb = feats.index_select(0, ids)                                           # :97
            ops.cls_nll_fwd_bwd(emb, dim, params[0], params[1], classes, labels_dev, ids, loss, None, grads[0], grads[1],
                                mask_relu_input=False)                                 # :99-104
            ops.clip_sgd(tl, max_norm, lr, 1.0, zero_grads=True)                       # :106-108

    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)

    def next_order(epoch):
        if orders is not None:
            order.copy_(torch.from_numpy(orders[epoch]), non_blocking=False)
        else:
            order.copy_(train_dev[torch.randperm(n_train, device=dev, generator=gen)])

    graph = None
    for epoch in range(epochs):
        next_order(epoch)
        if not use_graph:
            epoch_body()
        elif graph is None:
            saved = [p.clone() for p in params]                        # before the fork: the side stream is ordered behind it
            side = torch.cuda.Stream(device=dev)                       # warm-up off the capture (allocator, lazy loads)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                epoch_body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            for p, q in zip(params, saved):
                p.copy_(q)
            for g in grads:
                g.zero_()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                epoch_body()
            graph.replay()
        else:
            graph.replay()
        if on_epoch is not None:
            on_epoch(epoch)
    return loss
