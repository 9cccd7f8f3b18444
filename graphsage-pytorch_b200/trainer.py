"""Device-resident supervised train step: the body of the reference's `apply_model`
(src/utils.py:141-191, learn_method='sup') as one CUDA-graph replay.

    seeds -> sample -> unique/remap -> sample -> agg1 -> layer1 -> agg2 -> layer2 -> classifier ->
    NLL -> backward (classifier, layer2, scatter, layer1) -> all-reduce + clip + SGD (one kernel)

The eager drop-in classes (models.py) pay Python + autograd + launch overhead per kernel; at
b_sz=1024 the kernels themselves are ~100 us, so the step is captured once (all shapes are
static: frontier sizes that are only known on the device are passed as device counters, see
include/gsage_b200.h) and replayed.  The Philox offset is a device counter bumped inside the
graph so every replay draws fresh neighbours.  The gradients live in one flat buffer; the
exchange step of the data-parallel path and the update are ONE kernel
(gs_dp_allreduce_clip_sgd: push over NVLink peer memory, rank-ordered sum, per-model clip, SGD,
zero), captured in the same graph as forward/backward -- a step is a single graph replay at any
world size.  `exchange='nccl'` keeps the library all-reduce + separate update kernels as the
comparison baseline.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch

from . import native, ops
from .models import _PRECISIONS, Classification, GraphSage
from .peer import DpExchange


def flat_layout(shapes):
    """Offsets of every parameter's gradient inside the single flat buffer that is all-reduced
    (one collective per step).  Each view starts on a 16-byte boundary so the
    kernels' 128-bit accesses stay aligned.  Returns (offsets, total_elements)."""
    sizes = [int(np.prod(s)) if len(s) else 1 for s in shapes]
    offs = np.concatenate([[0], np.cumsum([(n + 3) & ~3 for n in sizes])]).astype(np.int64)
    return [int(o) for o in offs[:-1]], int(offs[-1])


def dp_allreduce_(flat_grad: torch.Tensor, world_size: int, group=None) -> torch.Tensor:
    """The one exchange step of the data-parallel path: sum the flat gradient over ranks.  The
    division by world_size happens inside the update (gs_clip_sgd's grad_div), after which every
    rank clips and steps identically, so replicas stay bit-identical."""
    if world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def shard_batches(train: np.ndarray, b_sz: int, steps: int, rank: int, world: int, seed: int) -> np.ndarray:
    """Disjoint b_sz slices of the shuffled train ids per (step, rank): rank r takes slice r of
    every global batch of world*b_sz seeds (src/utils.py:127,145 applied per rank)."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(train)
    need = b_sz * steps * world
    if need > len(perm):
        perm = np.concatenate([perm] * (need // len(perm) + 1))
    return np.ascontiguousarray(perm[:need].reshape(steps, world, b_sz)[:, rank, :])


class UnsupervisedTrainer:
    """N1 for every `learn_method` that needs the batch EXTENSION of src/utils.py:149 -- 'unsup', 'plus_unsup' and, with
    `learn_method='sup'`, the reference's own supervised step, whose NLL is averaged over the extended batch
    (src/utils.py:141-191, branches :159-181).  One step = extend the batch with random-walk positives and far
    negatives (device samplers), forward of the union, pair loss (`unsup_loss` = 'normal' -> get_loss_sage with num_neg
    100, 'margin' -> get_loss_margin with num_neg 6, utils.py:119-125; not for 'sup'), for 'plus_unsup' / 'sup' the
    classifier + NLL over the WHOLE extended batch (:161-164, :169-174), backward, exchange, clip_grad_norm_(5) per
    model and SGD(lr 0.7) (:185-187; in 'unsup' mode the classifier receives no gradient) -- without a host round trip:
    the size of the extended batch stays on the device (`UnsupervisedLoss.extend_device`,
    `GraphSage._run_prep(num_rows=...)`, `gs_cls_nll_fwd_bwd(num_rows_dev)`), where the drop-in `extend_nodes` has to
    hand a python list back to the reference's loop.
    Data parallel like SupervisedTrainer: the gradients live in one flat buffer and the fused exchange + clip + SGD
    kernel (peer.DpExchange) ends the step at any `world_size`; every rank trains on its own seed slice.
    `use_graph=True` captures the whole step as one CUDA graph: every buffer is sized by its static bound, the batch is
    copied into a static seed buffer, and the samplers add a device-resident step counter to their Philox offsets
    (`offset_dev`), so replay t draws what the eager step t draws.
    A layer-1 width that is not a multiple of 4 (Reddit's 602) runs on zero-padded copies of W1 / dW1 so the GEMMs keep
    their asynchronous tensor-core path (two tiny copies per step instead of register-staged operand loads)."""

    def __init__(self, model: GraphSage, unsupervised_loss, b_sz: int, *, unsup_loss: str = "normal",
                 learn_method: str = "unsup", classifier: Optional[Classification] = None, labels=None,
                 lr: float = 0.7, max_norm: float = 5.0, use_graph: bool = False, process_group=None, world_size: int = 1,
                 rank: int = 0):
        if unsup_loss not in ("normal", "margin"):
            raise ValueError("unsup_loss can be only 'margin' or 'normal'.")             # utils.py:124-125 (it exits)
        self.plus = learn_method in ("plus_unsup", "sup")   # a classifier head trained with the NLL of the extended batch
        self.pairs = learn_method != "sup"                  # anything but 'sup' trains on the pair loss (:165-181)
        if self.plus and (classifier is None or labels is None):
            raise ValueError(f"learn_method='{learn_method}' needs the classifier and the labels")
        self.model, self.unsup, self.b_sz = model, unsupervised_loss, int(b_sz)
        self.mode = 1 if unsup_loss == "margin" else 0
        self.num_neg = 6 if unsup_loss == "margin" else 100                              # utils.py:119-123
        self.lr, self.max_norm, self.world_size = lr, max_norm, int(world_size)
        _, _, dev = model._state()
        self.dev = dev
        self.weights: List[torch.Tensor] = [getattr(model, f'sage_layer{i}').weight for i in range(1, model.num_layers + 1)]
        params = list(self.weights)
        if self.plus:
            lin = classifier.layer[0]
            self.cls_w, self.cls_b = lin.weight, lin.bias
            params += [self.cls_w, self.cls_b]
            self.labels = (labels if isinstance(labels, torch.Tensor) else
                           torch.from_numpy(np.asarray(labels, dtype=np.int64))).to(dev)
            self.loss_sup = torch.zeros((1,), dtype=torch.float32, device=dev)
        for p in params:
            native.require_cuda(p, "parameters")
        if self.world_size > 1:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():          # replicas start identical: rank 0's parameters win
                for p in params:
                    dist.broadcast(p.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                                   group=process_group)
        n_sage = len(self.weights)
        offs, total = flat_layout([tuple(p.shape) for p in params])
        self.flat_grad = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.grads = [self.flat_grad[o:o + p.numel()].view_as(p) for o, p in zip(offs, params)]
        self.cls_grads = self.grads[n_sage:]
        # clip groups follow src/utils.py:185-186: model 0 = graphSage, model 1 = classification
        self.dp = DpExchange(self.flat_grad, [p.data for p in params], offs, [0] * n_sage + [1] * (len(params) - n_sage),
                             world=self.world_size, rank=rank, group=process_group)
        # zero-padded W1 / dW1 when the layer-1 width is not a multiple of 4 (see the class docstring)
        f = model.input_size
        self._pad = None
        if f % 4 and _PRECISIONS[model.precision] != native.PREC_FP32:
            f4 = ops.pad4(f)
            cols = f4 if model.gcn else 2 * f4
            self._pad = dict(f=f, f4=f4, w=torch.zeros((self.weights[0].shape[0], cols), dtype=torch.float32, device=dev),
                             g=torch.zeros((self.weights[0].shape[0], cols), dtype=torch.float32, device=dev))
        self.one = torch.ones((1,), dtype=torch.float32, device=dev)
        self.loss = torch.zeros((1,), dtype=torch.float32, device=dev)
        self.last_layers = None
        self.last_count = None
        self.use_graph = bool(use_graph)
        self.seeds = torch.zeros((self.b_sz,), dtype=torch.int32, device=dev)        # static input of the captured step
        self.step_counter = torch.zeros((1,), dtype=torch.int64, device=dev)          # Philox offset of the captured step
        self._loss_out = torch.zeros((1,), dtype=torch.float32, device=dev)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0

    def step_device(self, seeds) -> torch.Tensor:
        """One training step on `seeds` (device int32 tensor, numpy array or list of b_sz node ids).  Returns the
        device loss ([1]); nothing is copied to the host."""
        if not self.use_graph:
            before = native.launch_count()
            out = self._body(seeds, None)
            self.launches_per_step = native.launch_count() - before
            return out
        from .models import _as_device_ids
        ids = _as_device_ids(seeds, self.dev)
        if ids.shape[0] != self.b_sz:
            raise ValueError(f"the captured step takes batches of exactly {self.b_sz} seeds")
        self.seeds.copy_(ids, non_blocking=True)
        if self._graph is None:
            self._capture()
        self._graph.replay()
        self.loss = self._loss_out
        return self._loss_out

    def _capture(self):
        m, u = self.model, self.unsup
        params = [w.data for w in self.weights] + ([self.cls_w.data, self.cls_b.data] if self.plus else [])
        saved = [p.clone() for p in params]
        calls = (u._calls, m._calls)
        side = torch.cuda.Stream(device=self.dev)                 # warm-up off the capture: allocator, lazy module loads
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._body(self.seeds, self.step_counter)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.dev)
        for p, q in zip(params, saved):                            # the warm-up took a real step: undo it
            p.copy_(q)
        u._calls, m._calls = calls                                 # the captured launches carry call number calls + 1
        before = native.launch_count()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            loss = self._body(self.seeds, self.step_counter)
            self._loss_out.copy_(loss.reshape(1))
            self.step_counter.add_(1)
        self.launches_per_step = native.launch_count() - before

    def _effective_weights(self):
        """The weights the GEMMs read: W1 re-laid into its zero-padded copy when the input width needs it."""
        weights = [w.detach() for w in self.weights]
        if self._pad is not None:
            f, f4, wp = self._pad['f'], self._pad['f4'], self._pad['w']
            wp[:, :f].copy_(weights[0][:, :f])
            if not self.model.gcn:
                wp[:, f4:f4 + f].copy_(weights[0][:, f:])
            weights[0] = wp
        return weights

    def _body(self, seeds, offset_dev) -> torch.Tensor:
        m, u = self.model, self.unsup
        uniq, num_uniq = u.extend_device(seeds, self.num_neg, offset_dev=offset_dev)      # utils.py:149
        weights = self._effective_weights()
        layers = m._run_prep(uniq, None, offset_dev=offset_dev, num_rows=num_uniq)
        if self._pad is not None:
            layers[0].dim_in = self._pad['f4']                   # layer 1 contracts over the padded width (pad columns are zero)
        layers = m._run_compute(layers, weights)                 # utils.py:157
        self.last_layers, self.last_count = layers, num_uniq
        emb = layers[-1].h
        p = u._pairs
        gemb = torch.zeros_like(emb)
        loss = None
        if self.pairs:
            loss, coef_pos, coef_neg, num_active = ops.pair_loss_fwd(emb, m.out_size, p['seed_idx'], p['pos_ptr'], p['pos_idx'],
                                                                     p['neg_ptr'], p['neg_idx'], self.mode, float(u.Q),
                                                                     float(u.MARGIN))     # utils.py:169-180
        if self.plus:
            # classifier + NLL mean over the extended batch (:161-164); writes its grad_emb rows, the pair loss adds to them
            ops.cls_nll_fwd_bwd(emb, m.out_size, self.cls_w.detach(), self.cls_b.detach(), self.cls_w.shape[0], self.labels,
                                uniq, self.loss_sup, gemb, self.cls_grads[0], self.cls_grads[1],
                                precision=_PRECISIONS[m.precision], mask_relu_input=False, num_rows=num_uniq)
            loss = self.loss_sup if loss is None else loss + self.loss_sup                # :164 / :174
        if self.pairs:
            ops.pair_loss_bwd(emb, m.out_size, p['seed_idx'], p['pos_ptr'], p['pos_idx'], p['neg_ptr'], p['neg_idx'], coef_pos,
                              coef_neg, num_active, self.one, gemb)                       # utils.py:184
        n_sage = len(self.weights)
        grad_bufs = list(self.grads[:n_sage])
        if self._pad is not None:
            self._pad['g'].zero_()
            grad_bufs[0] = self._pad['g']
        m._run_backward(layers, gemb, weights, [True] * n_sage, grad_bufs=grad_bufs, own_grad=True)
        if self._pad is not None:                                # dW1 back into its place in the flat gradient
            f, f4, gp = self._pad['f'], self._pad['f4'], self._pad['g']
            self.grads[0][:, :f].copy_(gp[:, :f])
            if not m.gcn:
                self.grads[0][:, f:].copy_(gp[:, f4:f4 + f])
        self.dp.update(self.max_norm, self.lr, None)             # exchange + clip per model + SGD + zero (utils.py:185-191)
        self.loss = loss
        return loss

    def step(self, nodes_batch) -> torch.Tensor:
        return self.step_device(nodes_batch)

    def check(self):
        """Raise if the fused exchange ever timed out waiting for a peer (synchronises the stream)."""
        if self.world_size > 1:
            self.dp.status()


def _zero_beside(side: torch.cuda.Stream, loss: torch.Tensor, model: GraphSage, layers):
    """Fork: zero the loss accumulator and the backward's scatter targets on `side` while the forward
    GEMMs run on the current stream.  Returns (scatter_bufs, event to wait on before the loss kernel)."""
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        loss.zero_()
        bufs = model.zeroed_scatter_bufs(layers)
        ev = torch.cuda.Event()
        ev.record(side)
    for b in bufs:
        if b is not None:
            b.record_stream(main)
    return bufs, ev


class SupervisedTrainer:
    def __init__(self, model: GraphSage, classifier: Classification, labels, b_sz: int, *, lr: float = 0.7,
                 max_norm: float = 5.0, use_graph: bool = True, process_group=None, world_size: int = 1,
                 rank: int = 0, exchange: str = "peer"):
        csr, table, dev = model._state()
        self.model, self.classifier, self.dev, self.b_sz = model, classifier, dev, int(b_sz)
        self.lr, self.max_norm, self.world_size, self.pg = lr, max_norm, world_size, process_group
        self.labels = labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.asarray(labels, dtype=np.int64))
        self.labels = self.labels.to(dev)
        self.weights: List[torch.Tensor] = [getattr(model, f'sage_layer{i}').weight for i in range(1, model.num_layers + 1)]
        lin = classifier.layer[0]
        self.cls_w, self.cls_b = lin.weight, lin.bias
        for p in self.weights + [self.cls_w, self.cls_b]:
            native.require_cuda(p, "parameters")
        # one flat gradient buffer (single allreduce), views per tensor
        params = self.weights + [self.cls_w, self.cls_b]
        if world_size > 1:
            # replicas must start identical (the exchange only averages gradients): rank 0's parameters win
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                for p in params:
                    dist.broadcast(p.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                                   group=process_group)
        offs, total = flat_layout([tuple(p.shape) for p in params])
        self.flat_grad = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.grads = [self.flat_grad[o:o + p.numel()].view_as(p) for o, p in zip(offs, params)]
        n_sage = len(self.weights)
        # clip is per model (src/utils.py:185-186): graphSage parameters, then classification parameters
        self.tl_sage = ops.TensorList([p.data for p in self.weights], self.grads[:n_sage])
        self.tl_cls = ops.TensorList([self.cls_w.data, self.cls_b.data], self.grads[n_sage:])
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' (fused NVLink kernel) or 'nccl'")
        self.exchange = exchange
        self.dp: Optional[DpExchange] = None
        # Layer 1 on DENSE operands (GS_DENSE_X1=1, off by default): the aggregation writes the layer's input rows
        # [self | agg] and their low tf32 halves (ops.agg_fwd_x), the fused update keeps W1's low half current, and the
        # forward GEMM is fed by TMA alone (csrc/sage_gemm_tma.cu).  Measured at the headline configuration: the GEMM
        # drops from 13.8 to 10.1 us, but the aggregation moves 66 MB instead of 50 (18.5 us instead of 11.5), outlasts
        # the GEMM it hides behind and keeps the top-layer kernel's CTAs off the SMs: 73.9 us per step against 66.5 with
        # the self rows gathered inside the GEMMs (profiles/r2_dense_x1_sweep.txt).
        self.dense_x1 = (os.environ.get("GS_DENSE_X1", "0") == "1" and exchange == "peer" and model.dense_x_ok()
                         and _PRECISIONS[model.precision] != native.PREC_FP32)
        # W1's low tf32 half w - trunc_tf32(w), kept current by the fused update: the layer-1 forward GEMM takes it as a
        # TMA operand next to W1 itself (csrc/sage_gemm_tma.cu) instead of splitting the tile in shared memory
        self.weights_lo: List[Optional[torch.Tensor]] = [None] * n_sage
        if exchange == "peer" and _PRECISIONS[model.precision] == native.PREC_TF32X3 and self.weights[0].shape[1] % 4 == 0:
            self.weights_lo[0] = ops.split_lo(self.weights[0].data)
        self._lo_version = [w._version for w in self.weights]
        # replicas of the classifier gradients (GS_CLS_REPS=8, off by default): the top-layer kernel's 64+ CTAs spread
        # their atomic adds over them and the fused update folds them back in.  Measured: 0.8K of the kernel's 30K
        # cycles saved, the same again spent in the update kernel -- the adds are issue-bound, not contention-bound.
        self._cls_rep = None
        reps = int(os.environ.get("GS_CLS_REPS", "1"))
        if reps > 1 and exchange == "peer" and self.cls_w.numel() % 4 == 0 and self.cls_w.shape[1] == ops.TOP_H:
            self._cls_rep = (torch.zeros((reps - 1, self.cls_w.numel()), dtype=torch.float32, device=dev),
                             torch.zeros((reps - 1, ops.TOP_MAX_CLASSES), dtype=torch.float32, device=dev))
        if exchange == "peer":
            # clip groups follow src/utils.py:185-186: model 0 = graphSage, model 1 = classification
            self.dp = DpExchange(self.flat_grad, [p.data for p in params], offs, [0] * n_sage + [1, 1],
                                 world=world_size, rank=rank, group=process_group,
                                 params_lo=self.weights_lo + [None, None],
                                 extras=([None] * n_sage + list(self._cls_rep)) if self._cls_rep is not None else None)
        #: poll the exchange's sticky status word every this many steps (0 = only at flush()/check()); a timed-out
        #: peer wait would otherwise leave the ranks training different weights silently
        self.status_every = 256
        self._steps_since_check = 0
        self.seeds = torch.zeros((self.b_sz,), dtype=torch.int32, device=dev)
        self.seeds_pinned = torch.zeros((self.b_sz,), dtype=torch.int32).pin_memory()
        self.loss = torch.zeros((1,), dtype=torch.float32, device=dev)
        self.step_counter = torch.zeros((1,), dtype=torch.int64, device=dev)
        self.use_graph = use_graph
        self._graph_fb: Optional[torch.cuda.CUDAGraph] = None
        self._graph_up: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0
        self.last_layers = None
        self._side = torch.cuda.Stream(device=dev)      # dW GEMMs run beside the dX/scatter chain
        #: run the top layer + classifier + loss, forward and backward, as ONE launch (ops.sage_top_sup) when the
        #: configuration allows it (MEAN, hidden 128, <= 64 classes, tensor-core precision); GS_FUSED_TOP=0: never
        self.fused_top = os.environ.get("GS_FUSED_TOP", "1") != "0"
        self._top_ws = ops.sage_top_workspace(dev)
        #: two-layer steps: the classifier's weight gradient dlog^T . h runs as a third problem of the weight-gradient
        #: launch (a few row chunks) instead of one atomic add per CTA and element inside the top-layer kernel, which
        #: then only saves d(logits); GS_CLS_DW_GROUP=0: inside the top-layer kernel
        self.cls_dw_in_group = os.environ.get("GS_CLS_DW_GROUP", "1") != "0" and self._cls_rep is None
        self._dlog: Optional[torch.Tensor] = None

    # ---- the weight-dependent half of a step on prepared frontiers -------------------------------
    def _fused_top_ok(self, layers) -> bool:
        m = self.model
        return (self.fused_top and m.num_layers >= 2 and
                ops.sage_top_supported(m.out_size, m.out_size, int(self.cls_w.shape[0]), layers[-1].stride,
                                       _PRECISIONS[m.precision], m.agg_func == 'MEAN'))

    def _train_on(self, layers, seeds, loss_out: Optional[torch.Tensor] = None):
        """forward of the layers -> classifier -> NLL -> backward into self.grads (src/utils.py:157-163,184) for the
        batch `seeds` whose frontiers `layers` _run_prep has prepared.  Default path: the layers below the top one as
        tcgen05 GEMMs, then ONE launch for the whole top layer (gather + mean, SageLayer, classifier, log-softmax,
        NLL, their backward, dX and the scatter into the gradient of the layer below: ops.sage_top_sup), then the
        weight-gradient GEMMs -- the top layer's on the side stream beside the chain of the layers below."""
        m = self.model
        weights = [w.detach() for w in self.weights]
        n_sage = len(self.weights)
        loss_out = self.loss if loss_out is None else loss_out
        if not self._fused_top_ok(layers):
            return self._train_on_unfused(layers, seeds, weights, loss_out)
        L, H = m.num_layers, m.out_size
        prec = _PRECISIONS[m.precision]
        below, top = layers[L - 2], layers[L - 1]
        if below.gh is None:         # gradient w.r.t. the output of the layer below; zero-filled by that layer's GEMM
            below.gh = torch.empty((below.rows_max, H), dtype=torch.float32, device=self.dev)
        g_below = below.gh
        m._run_compute(layers, weights, upto=L - 1, zero_grad_of_last=g_below, weights_lo=self.weights_lo)
        top.table_in, top.dim_in = below.h, H
        cls_in_group = self.cls_dw_in_group and L == 2
        if cls_in_group and (self._dlog is None or self._dlog.shape[0] < top.rows_max):
            self._dlog = torch.zeros((top.rows_max, ops.TOP_MAX_CLASSES), dtype=torch.float32, device=self.dev)
        top.h, top.agg, top.dz = ops.sage_top_sup(below.h, top.nbr_idx, top.stride, top.cnt, top.self_idx, top.num_rows,
                                                  top.rows_max, weights[L - 1], m.gcn, self.cls_w.detach(),
                                                  self.cls_b.detach(), self.labels, seeds, loss_out,
                                                  None if cls_in_group else self.grads[n_sage],
                                                  self.grads[n_sage + 1], g_below, self._top_ws, prec, out_h=top.h,
                                                  out_agg=top.agg, out_dz=top.dz,
                                                  cls_w_rep=self._cls_rep[0] if self._cls_rep else None,
                                                  cls_b_rep=self._cls_rep[1] if self._cls_rep else None,
                                                  out_dlog=self._dlog if cls_in_group else None)
        top.argmax, dz = None, top.dz
        self.last_layers = layers
        if L == 2:
            # the weight gradients are leaves now: ONE grid, CTAs split in proportion to the work (csrc/sage_gemm_tc.cu)
            g = int(m.gcn)
            problems = [
                (None if m.gcn else below.table_in, below.self_idx, below.agg, below.dim_in, g_below, below.h, H,
                 below.num_rows, below.rows_max, self.grads[0], g, 0, 0),
                (None if m.gcn else below.h, top.self_idx, top.agg, H, dz, top.h, H, top.num_rows, top.rows_max,
                 self.grads[1], g, 0, 0)]
            if cls_in_group:         # grad Wc[c, :] = sum_r dlog[r, c] h[r, :]: a 'gcn' problem whose rows are h
                classes = int(self.cls_w.shape[0])
                problems.append((None, None, top.h, H, self._dlog, None, classes, top.num_rows, top.rows_max,
                                 self.grads[n_sage], 1, 0, (classes + 3) & ~3))
            ops.sage_gemm_bwd_w_group(problems, prec)
            return
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):                 # dW of the top layer: a leaf, beside the chain below
            ops.sage_gemm_bwd_w(None if m.gcn else below.h, top.self_idx, top.agg, H, dz, top.h, H, m.gcn, False,
                                top.num_rows, top.rows_max, self.grads[L - 1], precision=prec)
        for t in (dz, top.agg, top.h, below.h):
            t.record_stream(self._side)
        m._run_backward(layers[:L - 1], g_below, weights[:L - 1], [True] * (L - 1), grad_bufs=self.grads[:L - 1],
                        own_grad=True, top_masked=True, side_stream=None)
        main.wait_stream(self._side)

    def _train_on_unfused(self, layers, seeds, weights, loss_out):
        m = self.model
        n_sage = len(self.weights)
        scatter_bufs, zeroed = _zero_beside(self._side, loss_out, m, layers)
        layers = m._run_compute(layers, weights, weights_lo=self.weights_lo)
        self.last_layers = layers
        emb = layers[-1].h
        gemb = torch.empty_like(emb)
        torch.cuda.current_stream().wait_event(zeroed)
        ops.cls_nll_fwd_bwd(emb, m.out_size, self.cls_w.detach(), self.cls_b.detach(), self.cls_w.shape[0], self.labels,
                            seeds, loss_out, gemb, self.grads[n_sage], self.grads[n_sage + 1],
                            precision=_PRECISIONS[m.precision], mask_relu_input=True, zero_loss=False)   # utils.py:153,161-163
        m._run_backward(layers, gemb, weights, [True] * n_sage, grad_bufs=self.grads[:n_sage], own_grad=True,
                        top_masked=True, side_stream=self._side, scatter_bufs=scatter_bufs)

    # ---- the step, expressed once; runs eagerly or under capture -------------------------------
    def _forward_backward(self):
        layers = self.model._run_prep(self.seeds, None, offset_dev=self.step_counter, dense_x=self.dense_x1)
        self._train_on(layers, self.seeds)
        if self.dp is None:
            self.step_counter.add_(1)        # the fused update kernel bumps it otherwise

    def _update(self):
        if self.dp is not None:
            self.dp.update(self.max_norm, self.lr, self.step_counter)                                  # utils.py:184-191
            return
        div = float(self.world_size)
        ops.clip_sgd(self.tl_sage, self.max_norm, self.lr, div, zero_grads=True)                       # utils.py:185-191
        ops.clip_sgd(self.tl_cls, self.max_norm, self.lr, div, zero_grads=True)

    def _allreduce(self):
        if self.dp is None:
            dp_allreduce_(self.flat_grad, self.world_size, self.pg)

    def _capture(self):
        # warm-up on a side stream (allocator + lazy module loads), then capture
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._forward_backward()
                self._zero_grads()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.dev)
        before = native.launch_count()
        self._graph_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_fb):
            self._forward_backward()
            if self.dp is not None:
                self._update()               # the exchange is a kernel of ours: the whole step is one graph
        mid = native.launch_count()
        if self.dp is None:
            self._graph_up = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_up):
                self._update()
        self.launches_per_step = native.launch_count() - before
        self._fb_launches, self._up_launches = mid - before, native.launch_count() - mid
        self._zero_grads()

    def step_device(self, seeds_dev: torch.Tensor) -> torch.Tensor:
        """One step with the batch already in HBM (int32 [b_sz]).  Returns the device loss."""
        self.seeds.copy_(seeds_dev, non_blocking=True)
        return self._run()

    def step(self, nodes_batch) -> torch.Tensor:
        """One step from a HOST batch (numpy int64 as the reference's loop provides,
        src/utils.py:145): pinned staging copy -> H2D -> step.  Call .item() on the result for
        the loss (the reference does, src/utils.py:183)."""
        arr = np.asarray(nodes_batch)
        if arr.shape[0] != self.b_sz:
            raise ValueError(f"trainer was built for b_sz={self.b_sz}, got {arr.shape[0]}")
        self.seeds_pinned.numpy()[:] = arr
        self.seeds.copy_(self.seeds_pinned, non_blocking=True)
        return self._run()

    def _zero_grads(self):
        """Clear the flat gradient and the replicas beside it (after a warm-up pass that did not update)."""
        self.flat_grad.zero_()
        if self._cls_rep is not None:
            for t in self._cls_rep:
                t.zero_()

    def check(self):
        """Raise if the fused exchange ever timed out waiting for a peer (synchronises the stream)."""
        self._steps_since_check = 0
        if self.dp is not None and self.world_size > 1:
            self.dp.status()

    def weights_changed(self):
        """Call after writing the SageLayer weights from OUTSIDE the trainer (a checkpoint load): the low halves the
        layer-1 GEMM reads beside them are recomputed.  In-place torch writes to the parameters themselves are noticed
        through their version counters at the next step; writes through `.data` or raw pointers are not."""
        for i, lo in enumerate(self.weights_lo):
            if lo is not None:
                ops.split_lo(self.weights[i].data, lo)
            self._lo_version[i] = self.weights[i]._version

    def _count_steps(self, n: int = 1):
        if any(lo is not None and w._version != v for lo, w, v in zip(self.weights_lo, self.weights, self._lo_version)):
            self.weights_changed()
        self._steps_since_check += n
        if self.status_every and self.world_size > 1 and self._steps_since_check >= self.status_every:
            self.check()

    def _run(self) -> torch.Tensor:
        self._count_steps()
        if self.use_graph:
            if self._graph_fb is None:
                self._capture()
            self._graph_fb.replay()
            if self.dp is None:
                self._allreduce()
                self._graph_up.replay()
        else:
            before = native.launch_count()
            self._forward_backward()
            self._allreduce()
            self._update()
            self.launches_per_step = native.launch_count() - before
        return self.loss


class PipelinedTrainer(SupervisedTrainer):
    """Software-pipelined, device-resident supervised loop (src/utils.py:141-191).

    Sampling, unique/remap and the layer-1 aggregation of the raw features do not depend on the weights
    (`GraphSage._run_sample`, `_run_agg1`), so the graph of step n has TWO branches over THREE static frontier slots:

        training chain     layer-1 GEMM -> fused top layer -> weight-gradient GEMMs -> exchange + update   of batch n
        preparation chain  layer-1 aggregation of batch n+1,  then sampling + unique/remap of batch n+2

    The order inside the preparation chain matters: the aggregation (HBM-bound, fills every SM with resident warps)
    runs beside the layer-1 GEMM at the start of the step; run later it keeps the CTAs of the top-layer and
    weight-gradient kernels off the SMs until it has drained (measured: 73.6 us per step with sampling first,
    63.4 us with the aggregation first; scratch/branch_probe.py).  Every batch still gets exactly the same work and
    arithmetic as in `SupervisedTrainer` (tests/test_gpu_model.py::test_pipelined_trainer_matches_plain_trainer).

    Batches come from a device-side QUEUE (the reference slices every batch of an epoch from one shuffled array,
    src/utils.py:127,145): the top sampler launch of a preparation fetches the next row of the queued [rows x b_sz]
    array and advances the cursor, so consecutive replays need no host work in between and three steps share one
    graph launch.

        tr.set_queue(batches_int32_dev)        # or tr.feed(host_batch) per step (pinned H2D on a copy stream)
        tr.prime()                             # samples batches 0 and 1, aggregates batch 0
        loss = tr.run(n)                       # n steps: trains batch i | aggregates batch i+1, samples batch i+2
        loss = tr.flush()                      # trains on the (up to two) batches still in the pipeline
    `submit` / `submit_device` / `step` keep the one-call-per-batch form on top of the same machinery."""

    RING = 8          # rows of the internal staging ring used by feed()/submit(): two 3-step launches and the 2 in flight
    SLOTS = 3         # frontier slots: being trained on / aggregated / sampled

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.dp is None:
            raise ValueError("PipelinedTrainer needs exchange='peer' (the whole step must be one graph)")
        dev = self.dev
        self.slot_seeds = [torch.zeros((self.b_sz,), dtype=torch.int32, device=dev) for _ in range(self.SLOTS)]
        self.slot_layers = [None] * self.SLOTS
        self.slot_loss = torch.zeros((self.SLOTS,), dtype=torch.float32, device=dev)
        self._chain_fence = torch.zeros((1,), dtype=torch.float32, device=dev)      # see _both
        self.early_reads = os.environ.get("GS_EARLY_READS", "1") != "0"                   # A/B switch, see _compute
        self.sample_counter = torch.zeros((1,), dtype=torch.int64, device=dev)
        self.queue_desc = torch.zeros((4,), dtype=torch.int64, device=dev)        # {address, rows, next, ticket}
        self._queue = None
        self._ring = torch.zeros((self.RING, self.b_sz), dtype=torch.int32, device=dev)
        self._ring_pinned = torch.zeros((self.RING, self.b_sz), dtype=torch.int32).pin_memory()
        self._ring_events = [None] * self.RING
        self._fed = 0                          # batches written into the ring so far
        self._fetched = 0                      # batch fetches enqueued since the ring became the queue
        self._fetch_events = [None] * self.RING    # _fetch_events[j % RING]: recorded behind the j-th fetch
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._prep_stream = torch.cuda.Stream(device=dev)
        # diagnostics knobs (scratch/sweep.sh): occupancy cap of the background aggregation (CTAs/SM, 0 = none),
        # programmatic dependent launch inside the preparation chain, max-shared carveout for it, a high-priority
        # stream for the training chain.  None of them moves the step by more than 1 us once the aggregation runs
        # first (see the class docstring); the defaults are what bench.py measures.
        self.bg_agg_ctas = int(os.environ.get("GS_BG_AGG_CTAS", "0"))
        self.prep_pdl = os.environ.get("GS_PREP_PDL", "0") == "1"
        self.bg_carveout = os.environ.get("GS_BG_CARVE", "0") == "1"
        self._train_stream = (torch.cuda.Stream(device=dev, priority=-1)
                              if os.environ.get("GS_TRAIN_PRIO", "0") == "1" else None)
        self._graphs = [None] * self.SLOTS     # one step training on slot s
        self._graph_multi = None               # SLOTS steps (slots 0, 1, 2) in one launch
        self._cur: Optional[int] = None        # slot holding the oldest prepared, not yet trained batch
        self._pending = 0                      # batches sampled but not yet trained (2 in the steady state)
        self._agg_done = False                 # the batch in slot _cur has its layer-1 aggregation
        self._prep_call: Optional[int] = None  # the model call number every sampling launches with (see _sample)

    # ---- batch queue --------------------------------------------------------------------------------
    def set_queue(self, batches_dev: torch.Tensor):
        """Queue an int32 [rows, b_sz] array resident in HBM; steps consume its rows in order (and wrap)."""
        native.require_cuda(batches_dev, "batch queue")
        if batches_dev.dtype != torch.int32 or batches_dev.dim() != 2 or batches_dev.shape[1] != self.b_sz \
                or not batches_dev.is_contiguous():
            raise ValueError(f"batch queue must be a contiguous int32 [rows, {self.b_sz}] tensor")
        if self._pending:
            raise RuntimeError("flush() the pipeline before changing the batch queue")
        self._queue = batches_dev
        desc = torch.tensor([batches_dev.data_ptr(), batches_dev.shape[0], 0, 0], dtype=torch.int64)
        self.queue_desc.copy_(desc.to(self.dev), non_blocking=False)

    def feed(self, nodes_batch):
        """Stage one HOST batch (numpy/list) for a later step: pinned copy + H2D on the copy stream
        into the internal ring; the step that consumes it waits for the copy's event only."""
        arr = np.asarray(nodes_batch)
        if arr.shape[0] != self.b_sz:
            raise ValueError(f"trainer was built for b_sz={self.b_sz}, got {arr.shape[0]}")
        self._use_ring()
        r = self._fed % self.RING
        fetched = self._row_free(r)                       # raises when the ring is over-full
        if self._ring_events[r] is not None:
            self._ring_events[r].synchronize()            # the pinned row is free again
        self._ring_pinned[r].numpy()[:] = arr
        self._copy_stream.wait_event(fetched)             # ... and the device row has been fetched by its step
        with torch.cuda.stream(self._copy_stream):
            self._ring[r].copy_(self._ring_pinned[r], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._ring_events[r] = ev
        torch.cuda.current_stream().wait_event(ev)
        self._fed += 1

    def _use_ring(self):
        if self._queue is not self._ring:
            torch.cuda.current_stream().synchronize()
            self.set_queue(self._ring)
            self._fed = self._fetched = 0
            self._fetch_events = [None] * self.RING

    def _row_free(self, r: int) -> torch.cuda.Event:
        """Event behind the fetch that consumed the batch currently in ring row r (an already-signalled one when
        the row was never used).  Feeding more than RING batches ahead of the steps that fetch them would overwrite
        a batch no step has read yet: that is an error, not something to wait for (nothing is enqueued to wait on)."""
        prev = self._fed - self.RING                       # index of the batch being overwritten
        if prev < 0:
            ev = torch.cuda.Event()
            ev.record()
            return ev
        if self._fetched <= prev:
            raise RuntimeError(f"batch ring is full: {self._fed - self._fetched} batches fed ahead of the steps that "
                               f"fetch them (capacity {self.RING}); call run()/submit() between feeds")
        return self._fetch_events[prev % self.RING]

    def _mark_fetched(self, n: int = 1):
        """n more batch fetches are enqueued on the current stream: one event behind them covers all."""
        if self._queue is not self._ring:
            return
        ev = torch.cuda.Event()
        ev.record()
        for _ in range(n):
            self._fetch_events[self._fetched % self.RING] = ev
            self._fetched += 1

    def _unfetched(self) -> int:
        """Batches in the queue no sampling has fetched yet (a device queue wraps: never runs dry)."""
        return (self._fed - self._fetched) if self._queue is self._ring else (1 << 30)

    def feed_device(self, seeds_dev: torch.Tensor):
        self._use_ring()
        r = self._fed % self.RING
        torch.cuda.current_stream().wait_event(self._row_free(r))     # raises when the ring is over-full
        self._ring[r].copy_(seeds_dev, non_blocking=True)
        self._fed += 1

    # ---- the pieces of a step ---------------------------------------------------------------------
    def _sample(self, slot: int):
        # Philox offset of a sampling = (model call number << 8 | layer) + (sample_counter << 8).  The call number is
        # baked into a captured launch, so every sampling -- whichever graph or slot it was captured for, or eager --
        # carries the SAME one and the device counter (+1 per batch) alone tells consecutive batches apart.
        m = self.model
        if self._prep_call is None:
            self._prep_call = m._calls
        keep, m._calls = m._calls, self._prep_call
        # the top sampler launch fetches the next queued batch into slot_seeds[slot] itself (gs_sample_neighbors_ex)
        self.slot_layers[slot] = m._run_sample(self.slot_seeds[slot], None, offset_dev=self.sample_counter,
                                               reuse=self.slot_layers[slot], queue_desc=self.queue_desc)
        m._calls = max(keep, self._prep_call + 1)
        self.sample_counter.add_(1)

    def _aggregate(self, slot: int):
        self.model._run_agg1(self.slot_layers[slot], dense_x=self.dense_x1)

    def _compute(self, slot: int, update: bool = True, early: bool = False):
        # the loss of the batch trained from slot s lands in slot_loss[s]: the three steps of a multi-step graph
        # launch leave three losses behind (run() points self.loss at the last one)
        # early (from _both only): the frontiers' index lists and row counts come from the preparation branch --
        # another stream, joined by an event -- so the kernels of the training chain may read them before their PDL
        # wait (gs_set_early_reads); never when the lists were written by launches of THIS stream just before
        native.set_early_reads(early and self.early_reads)
        try:
            self._train_on(self.slot_layers[slot], self.slot_seeds[slot], loss_out=self.slot_loss[slot:slot + 1])
        finally:
            native.set_early_reads(False)
        if update:
            self.dp.update(self.max_norm, self.lr, None)

    def _background(self, on: bool):
        native.set_pdl((1 if self.prep_pdl else 0) if on else -1)
        native.set_agg_ctas(self.bg_agg_ctas if on else 0)
        native.set_background(self.bg_carveout and on)

    def _both(self, slot: int, update: bool = True):
        """train on `slot` | aggregate slot+1, sample into slot+2 -- a fork/join, eagerly or under capture"""
        main = torch.cuda.current_stream()
        if not torch.cuda.is_current_stream_capturing():
            # Eager calls may follow launches of this stream that wrote the frontiers (prime(), the warm-up).  A launch
            # WITHOUT the programmatic attribute is a full barrier in the chain: everything before it has completed
            # before anything after it is scheduled.  (Captured: the graph launch / the join edges are that barrier.)
            self._chain_fence.add_(0)
        self._prep_stream.wait_stream(main)
        try:
            with torch.cuda.stream(self._prep_stream):
                # programmatic dependent launch stays on for the (critical) training chain only: early-resident
                # dependents of the preparation chain would hold SM slots the GEMMs need
                self._background(True)
                self._aggregate((slot + 1) % self.SLOTS)
                self._sample((slot + 2) % self.SLOTS)
            self._background(False)
            if self._train_stream is not None:            # the critical chain on a high-priority stream (fork/join)
                self._train_stream.wait_stream(main)
                with torch.cuda.stream(self._train_stream):
                    self._compute(slot, update, early=True)
                main.wait_stream(self._train_stream)
            else:
                self._compute(slot, update, early=True)
        finally:
            self._background(False)
        main.wait_stream(self._prep_stream)

    def _capture_pipeline(self):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up: allocates every slot, loads modules; no update, cursors restored
            saved = (self.sample_counter.clone(), self.queue_desc.clone())
            for slot in range(self.SLOTS):
                self._sample(slot)
                self._aggregate(slot)
            for slot in range(self.SLOTS):
                self._both(slot, update=False)
                self._zero_grads()
            self.sample_counter.copy_(saved[0])
            self.queue_desc.copy_(saved[1])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.dev)
        for slot in range(self.SLOTS):
            before = native.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._both(slot)
            self._graphs[slot] = g
            self.launches_per_step = native.launch_count() - before
        self._graph_multi = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_multi):
            for slot in range(self.SLOTS):
                self._both(slot)
        self._zero_grads()

    # ---- public ------------------------------------------------------------------------------------
    def prime(self):
        """Fill the pipeline: sample the first two queued batches, aggregate the first (nothing to train on yet)."""
        if self._queue is None:
            raise RuntimeError("set_queue() or feed() a batch first")
        if self.use_graph and self._graphs[0] is None:
            self._capture_pipeline()
        if self._cur is None:
            if self._unfetched() < 2:
                raise RuntimeError("prime() needs two queued batches (feed() them first)")
            self._sample(0)
            self._aggregate(0)
            self._sample(1)
            self._mark_fetched(2)
            self._cur, self._pending, self._agg_done = 0, 2, True

    def run(self, n: int = 1) -> torch.Tensor:
        """n pipelined steps: step i trains on the oldest prepared batch, aggregates the next one and samples the
        one after from the queue.  Returns the device loss of the last trained batch."""
        if self._cur is None or self._pending < 2:
            raise RuntimeError("prime() the pipeline first")
        if self._unfetched() < n:
            raise RuntimeError(f"run({n}) needs {n} queued batches to sample, {self._unfetched()} fed")
        self._count_steps(n)
        while n > 0:
            if self.use_graph and self._cur == 0 and n >= self.SLOTS:
                self._graph_multi.replay()
                self._mark_fetched(self.SLOTS)
                n -= self.SLOTS
                continue
            if self.use_graph:
                self._graphs[self._cur].replay()
            else:
                before = native.launch_count()
                self._both(self._cur)
                self.launches_per_step = native.launch_count() - before
            self._mark_fetched()
            self._cur = (self._cur + 1) % self.SLOTS
            n -= 1
        last = (self._cur + self.SLOTS - 1) % self.SLOTS
        self.loss = self.slot_loss[last:last + 1]
        return self.loss

    def flush(self) -> Optional[torch.Tensor]:
        """Drain: sample what is still queued in the ring (at most what the free slots hold), then train on every
        prepared batch in order (nothing left to prepare beside them)."""
        if self._queue is self._ring and self._cur is None and self._unfetched() > 0:
            self._cur, self._pending, self._agg_done = 0, 0, False
        if self._queue is self._ring and self._cur is not None:
            while self._unfetched() > 0 and self._pending < self.SLOTS:
                self._sample((self._cur + self._pending) % self.SLOTS)
                self._mark_fetched()
                self._pending += 1
        if self._cur is None:
            return None
        self._count_steps(0)
        while self._pending > 0:
            if not self._agg_done:
                self._aggregate(self._cur)
            self._compute(self._cur)
            self.loss = self.slot_loss[self._cur:self._cur + 1]
            self._cur = (self._cur + 1) % self.SLOTS
            self._pending -= 1
            self._agg_done = False
        self._cur = None
        self.check()
        return self.loss

    # one call per batch, on top of the queue: the batch handed over is trained on two calls later
    def _submitted(self) -> Optional[torch.Tensor]:
        if self._cur is None:
            if self._unfetched() < 2:
                return None
            self.prime()
            return None
        return self.run(1)

    def submit_device(self, seeds_dev: torch.Tensor) -> Optional[torch.Tensor]:
        """Hand over the next batch (int32 [b_sz] in HBM).  Trains on the batch submitted two calls earlier (returns
        its device loss) while the later ones are prepared; the first two calls only prepare."""
        self.feed_device(seeds_dev)
        return self._submitted()

    def submit(self, nodes_batch) -> Optional[torch.Tensor]:
        """`submit_device` from a HOST batch (pinned staging + H2D on the copy stream)."""
        self.feed(nodes_batch)
        return self._submitted()

    def step_device(self, seeds_dev: torch.Tensor) -> torch.Tensor:
        self.submit_device(seeds_dev)
        return self.flush()

    def step(self, nodes_batch) -> torch.Tensor:
        self.submit(nodes_batch)
        return self.flush()
