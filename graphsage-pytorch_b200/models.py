"""Drop-in for the reference's `src/models.py`: GraphSage, SageLayer, Classification,
UnsupervisedLoss with the same constructor / forward signatures, attribute names and
state_dict keys (`sage_layer{i}.weight`, `layer.0.weight`, `layer.0.bias`), so the
reference's own training loop (`src/utils.py`) runs against these classes unchanged.

Everything numerical happens in hand-written sm_100a kernels behind the C ABI of
include/gsage_b200.h (see ops.py / native.py).  There is no CPU path: a model whose
tensors are not on a CUDA device raises.  Reference lines are cited per method.
"""
from __future__ import annotations

import os
import weakref
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native, ops
from .graph import AdjCSR, DeviceCSR, adj_to_csr
from .peer import ShardedTable

_PRECISIONS = {"fp32": native.PREC_FP32, "tf32": native.PREC_TF32, "tf32x3": native.PREC_TF32X3}

# One device CSR per adjacency object (GraphSage and UnsupervisedLoss receive the same dict,
# src/main.py:54,61): id(adj) -> (weakref-or-strong ref, DeviceCSR)
_CSR_CACHE: Dict[tuple, tuple] = {}


def _device_csr(adj_lists, num_nodes: Optional[int], device) -> DeviceCSR:
    if isinstance(adj_lists, DeviceCSR):            # already resident (device-generated graphs)
        if num_nodes is not None and adj_lists.num_nodes != num_nodes:
            raise ValueError(f"adjacency has {adj_lists.num_nodes} rows, feature table {num_nodes}")
        return adj_lists
    key = (id(adj_lists), str(device))
    hit = _CSR_CACHE.get(key)
    if hit is not None and hit[0] is adj_lists and (num_nodes is None or hit[1].num_nodes == num_nodes):
        return hit[1]
    if num_nodes is None:
        if isinstance(adj_lists, AdjCSR):
            num_nodes = adj_lists.num_nodes
        else:
            hi = -1
            for node, nbrs in adj_lists.items():
                hi = max(hi, int(node), max(nbrs) if nbrs else -1)
            num_nodes = int(hi) + 1
    csr = DeviceCSR.from_adj(adj_lists, num_nodes, device)
    _CSR_CACHE[key] = (adj_lists, csr)
    if len(_CSR_CACHE) > 8:
        _CSR_CACHE.pop(next(iter(_CSR_CACHE)))
    return csr


def _as_device_ids(nodes, device) -> torch.Tensor:
    """numpy int64 array (src/utils.py:149), python list (src/utils.py:67) or tensor -> int32 on device."""
    if isinstance(nodes, torch.Tensor):
        return nodes.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
    arr = np.ascontiguousarray(np.asarray(nodes), dtype=np.int32)
    return torch.from_numpy(arr).to(device, non_blocking=True)


def _padded_table(x: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous, rows padded with zeros to a multiple of 4 floats (128-bit loads)."""
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    f = x.shape[1]
    if f % 4:
        x = F.pad(x, (0, 4 - f % 4))
    return x.contiguous()


# ================================================================================================
# Classification                                                       src/models.py:8-27
# ================================================================================================
class _ClassifyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, embeds, weight, bias):
        emb = _padded_table(embeds) if (embeds.shape[1] % 4 or not embeds.is_contiguous()
                                        or embeds.dtype != torch.float32) else embeds.detach()
        w = weight.detach().contiguous()
        b = bias.detach().contiguous() if bias is not None else None
        logp = ops.cls_fwd(emb, embeds.shape[1], w, b, w.shape[0])
        ctx.save_for_backward(emb, w, logp)
        ctx.has_bias = bias is not None
        ctx.dim = embeds.shape[1]
        return logp

    @staticmethod
    def backward(ctx, grad_logp):
        emb, w, logp = ctx.saved_tensors
        need_e, need_w, need_b = ctx.needs_input_grad
        dim = ctx.dim
        grad_emb = torch.empty_like(emb) if need_e else None
        grad_w = torch.zeros_like(w) if need_w else None
        grad_b = torch.zeros((w.shape[0],), dtype=torch.float32, device=w.device) if (need_b and ctx.has_bias) else None
        ops.cls_bwd(grad_logp.contiguous(), logp, emb, dim, w, w.shape[0], grad_emb, grad_w, grad_b)
        if grad_emb is not None and grad_emb.shape[1] != dim:
            grad_emb = grad_emb[:, :dim]
        return grad_emb, grad_w, grad_b


class Classification(nn.Module):
    """log_softmax(Linear(emb_size -> num_classes)); src/models.py:8-27."""

    def __init__(self, emb_size, num_classes):
        super().__init__()
        self.layer = nn.Sequential(nn.Linear(emb_size, num_classes))        # :14-17
        self.init_params()

    def init_params(self):                                                   # :20-23
        for param in self.parameters():
            if len(param.size()) == 2:
                nn.init.xavier_uniform_(param)

    def forward(self, embeds):                                               # :25-27
        lin = self.layer[0]
        native.require_cuda(lin.weight, "Classification parameters")
        return _ClassifyFn.apply(embeds, lin.weight, lin.bias)


# ================================================================================================
# SageLayer                                                            src/models.py:189-220
# ================================================================================================
class _SageLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, self_feats, agg_feats, weight, gcn, precision):
        dim = agg_feats.shape[1]
        agg = _padded_table(agg_feats)
        selff = None if gcn else _padded_table(self_feats)
        w = weight.detach().contiguous()
        rows, out_dim = agg.shape[0], w.shape[0]
        out = ops.sage_gemm_fwd(selff, None, agg, dim, w, out_dim, gcn, None, rows, True, precision)
        ctx.save_for_backward(selff, agg, w, out)
        ctx.gcn, ctx.dim = gcn, dim
        return out[:, :out_dim] if out.shape[1] != out_dim else out

    @staticmethod
    def backward(ctx, grad_out):
        selff, agg, w, out = ctx.saved_tensors
        gcn, dim = ctx.gcn, ctx.dim
        rows, out_dim = agg.shape[0], w.shape[0]
        g = _padded_table(grad_out)
        need_s, need_a, need_w = ctx.needs_input_grad[:3]
        grad_w = None
        if need_w:
            grad_w = torch.zeros_like(w)
            ops.sage_gemm_bwd_w(selff, None, agg, dim, g, out, out_dim, gcn, True, None, rows, grad_w)
        gs = ga = None
        if need_a or (need_s and not gcn):
            gs, ga = ops.sage_gemm_bwd_x(g, out, w, dim, out_dim, gcn, True, None, rows)
            ga = ga[:, :dim]
            gs = gs[:, :dim] if gs is not None else None
        return gs, ga, grad_w, None, None


class SageLayer(nn.Module):
    """relu(W . cat[self, agg]^T)^T ; src/models.py:189-220 (no bias, ReLU on every layer)."""

    def __init__(self, input_size, out_size, gcn=False, precision: str = "tf32x3"):
        super().__init__()
        self.input_size = input_size
        self.out_size = out_size
        self.gcn = gcn
        self.precision = precision
        self.weight = nn.Parameter(torch.empty(out_size, self.input_size if self.gcn else 2 * self.input_size))   # :201
        self.init_params()

    def init_params(self):                                                   # :205-207
        for param in self.parameters():
            nn.init.xavier_uniform_(param)

    def forward(self, self_feats, aggregate_feats, neighs=None):             # :209-220
        native.require_cuda(self.weight, "SageLayer.weight")
        return _SageLayerFn.apply(self_feats, aggregate_feats, self.weight, self.gcn, _PRECISIONS[self.precision])


# ================================================================================================
# GraphSage                                                            src/models.py:222-330
# ================================================================================================
class _Frontier:
    """Per-layer device state of one forward pass (rows = destination nodes of that layer)."""
    __slots__ = ("nodes", "num_rows", "rows_max", "stride", "nbr", "cnt", "nbr_idx", "self_idx", "agg", "argmax", "h",
                 "table_in", "dim_in", "dz", "gh", "x", "x_lo")
    # x / x_lo: layer 1's dense input rows [self | agg] and their low tf32 halves (GraphSage._run_agg1(dense_x=True)).
    # dz: d(pre-activation) of this layer (A operand of its weight-gradient GEMM), gh: gradient w.r.t. its output h;
    # both only exist on the fused-top path of the trainers.  Buffers survive from one forward to the next when the
    # frontiers are reused (`_run_prep(reuse=...)`): static addresses for captured steps.

    def __init__(self):
        for name in self.__slots__:
            setattr(self, name, None)

    def inherit(self, old: "_Frontier", first_layer: bool):
        """Take over the weight-dependent buffers of the frontier this one replaces (same static shapes)."""
        self.h, self.dz, self.gh = old.h, old.dz, old.gh
        self.agg, self.argmax = old.agg, old.argmax
        self.x, self.x_lo = old.x, old.x_lo
        if first_layer and not isinstance(old.table_in, ShardedTable) and old.self_idx is None:
            self.table_in = old.table_in                 # the fp32 self rows a sharded gather emits (see _run_agg1)


class _GraphSageFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, nodes_dev, injected, *weights):
        layers = model._run_forward(nodes_dev, [w.detach() for w in weights], injected)
        ctx.model, ctx.layers = model, layers
        ctx.save_for_backward(*weights)
        h = layers[-1].h
        model._last_layers = layers
        return h[:, :model.out_size] if h.shape[1] != model.out_size else h

    @staticmethod
    def backward(ctx, grad_out):
        weights = [w.detach() for w in ctx.saved_tensors]
        grads = ctx.model._run_backward(ctx.layers, grad_out, weights, ctx.needs_input_grad[3:])
        return (None, None, None, *grads)


class GraphSage(nn.Module):
    """Same constructor and forward as src/models.py:224,241.  Extra keyword arguments
    (`num_sample`, `precision`, `seed`) default to the reference's behaviour."""

    def __init__(self, num_layers, input_size, out_size, raw_features, adj_lists, device, gcn=False, agg_func='MEAN',
                 *, num_sample: int = 10, precision: str = "tf32x3", seed: Optional[int] = None,
                 unique_algo: str = "auto", normalize: bool = False):
        super().__init__()
        if agg_func not in ('MEAN', 'MAX'):
            raise ValueError("agg_func must be 'MEAN' or 'MAX' (src/models.py:311,316)")
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {list(_PRECISIONS)}")
        if not 1 <= num_sample <= native.MAX_FANOUT:
            raise ValueError(f"num_sample must be in 1..{native.MAX_FANOUT}")
        self.input_size = input_size
        self.out_size = out_size
        self.num_layers = num_layers
        self.gcn = gcn
        self.device = device
        self.agg_func = agg_func
        self.num_sample = num_sample          # src/models.py:277 default argument
        if unique_algo not in ("auto", "bitmap", "radix"):
            raise ValueError("unique_algo must be 'auto', 'bitmap' or 'radix'")
        self.unique_algo = unique_algo        # K2 path: bitmap/rank (default when N <= 2^28) or radix sort
        self.precision = precision
        #: row-L2-normalise every layer's output in the GEMM epilogue (the original GraphSAGE's step; this reference stops
        #: at the ReLU, src/models.py:219, so the default is off).  Forward only: embeddings / inference.
        self.normalize = bool(normalize)
        self.raw_features = raw_features      # :234
        self.adj_lists = adj_lists            # :235
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0x7FFFFFFFFFFFFFFF
        self._calls = 0
        #: GS_L2_PREFETCH=1 (off by default): the layer-1 sampler requests the feature rows it drew (and the nodes' own
        #: rows) into L2 for the aggregation kernel that follows.  Measured: the aggregation drops from 13.1 to 10.6 us,
        #: the sampler grows from 8.1 to 12.6 us (it becomes the DRAM-bound kernel), the step from 64.7 to 69.7 us
        self.l2_prefetch = os.environ.get("GS_L2_PREFETCH", "0") == "1"
        self._native_state = None
        self._bitmap_ws = None
        self._injected = None
        self._last_layers = None
        for index in range(1, num_layers + 1):                                # :237-239
            layer_size = out_size if index != 1 else input_size
            setattr(self, 'sage_layer' + str(index), SageLayer(layer_size, out_size, gcn=self.gcn, precision=precision))

    # ---- pickling (src/utils.py:52 saves the live modules): native handles are rebuilt lazily
    def __getstate__(self):
        state = self.__dict__.copy()
        state['_native_state'] = None
        state['_bitmap_ws'] = None
        state['_last_layers'] = None
        state['_injected'] = None
        return state

    # ---- lazily built device state: CSR + padded feature table ---------------------------------
    def _state(self):
        if self._native_state is None:
            dev = self.sage_layer1.weight.device
            if dev.type != 'cuda':
                dev = torch.device(self.device) if self.device is not None else dev
            if torch.device(dev).type != 'cuda':
                raise RuntimeError("GraphSage (gsage_b200) runs only on a CUDA device; there is no CPU fallback. "
                                   "Construct it with device=torch.device('cuda') and call .to(device).")
            native.load()
            feats = self.raw_features
            if isinstance(feats, ShardedTable):        # row-partitioned bf16 table, peer shards read over NVLink
                table = feats
            else:
                if not isinstance(feats, torch.Tensor):
                    feats = torch.as_tensor(np.asarray(feats), dtype=torch.float32)
                table = _padded_table(feats.to(dev))
            csr = _device_csr(self.adj_lists, table.shape[0], dev)
            self._native_state = (csr, table, torch.device(dev))
            use_bitmap = self.unique_algo == "bitmap" or (self.unique_algo == "auto" and csr.num_nodes <= (1 << 28))
            self._bitmap_ws = ops.unique_bitmap_workspace(csr.num_nodes, dev) if use_bitmap else None
        return self._native_state

    def inject_samples(self, calls):
        """Injected-sample mode: `calls` is the per-call list of (nodes, samp_neighs) recorded
        at the `_get_unique_neighs_list` seam of a reference run (src/models.py:250); the next
        forward consumes them instead of drawing its own.  Pass None to clear."""
        self._injected = calls

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, nodes_batch):                                           # :241-269
        csr, table, dev = self._state()
        nodes_dev = _as_device_ids(nodes_batch, dev)
        weights = [getattr(self, 'sage_layer' + str(i)).weight for i in range(1, self.num_layers + 1)]
        for w in weights:
            native.require_cuda(w, "GraphSage weights")
        injected, self._injected = self._injected, None
        return _GraphSageFn.apply(self, nodes_dev, injected, *weights)

    def _list_stride(self) -> int:
        return self.num_sample + (1 if self.gcn else 0)

    def _injected_lists(self, nodes_host: np.ndarray, call, positional: bool):
        """Recorded python sets -> fixed-stride id lists in the canonical form the sampler
        writes (ascending, own id dropped for gcn=False / present once for gcn=True)."""
        rec_nodes, rec_samp = call[0], call[1]
        by_node = None if positional else {int(n): s for n, s in zip(rec_nodes, rec_samp)}
        rows = []
        for i, n in enumerate(nodes_host.tolist()):
            s = rec_samp[i] if positional else by_node[n]
            ids = sorted(int(x) for x in s if int(x) != n)
            if self.gcn:
                ids = sorted(ids + [n])
            rows.append(ids)
        stride = max(self._list_stride(), max((len(r) for r in rows), default=1))
        nbr = np.full((len(rows), stride), -1, dtype=np.int32)
        cnt = np.zeros(len(rows), dtype=np.int32)
        for i, ids in enumerate(rows):
            nbr[i, :len(ids)] = ids
            cnt[i] = len(ids)
        return nbr, cnt, stride

    def _run_forward(self, nodes_dev: torch.Tensor, weights: Sequence[torch.Tensor], injected=None,
                     offset_dev: Optional[torch.Tensor] = None) -> List[_Frontier]:
        return self._run_compute(self._run_prep(nodes_dev, injected, offset_dev), weights)

    def _run_prep(self, nodes_dev: torch.Tensor, injected=None, offset_dev: Optional[torch.Tensor] = None,
                  reuse: Optional[List[_Frontier]] = None, num_rows: Optional[torch.Tensor] = None,
                  queue_desc: Optional[torch.Tensor] = None, dense_x: bool = False) -> List[_Frontier]:
        """The weight-independent half of a forward pass: `_run_sample` then `_run_agg1`."""
        return self._run_agg1(self._run_sample(nodes_dev, injected, offset_dev, reuse, num_rows, queue_desc), dense_x)

    def _run_sample(self, nodes_dev: torch.Tensor, injected=None, offset_dev: Optional[torch.Tensor] = None,
                    reuse: Optional[List[_Frontier]] = None, num_rows: Optional[torch.Tensor] = None,
                    queue_desc: Optional[torch.Tensor] = None) -> List[_Frontier]:
        """Sampling + unique/remap of every layer (src/models.py:249-251): the integer half of the preparation.
        `reuse`: the frontiers of an earlier call whose buffers are overwritten in place (static
        addresses: the pipelined trainer prepares step n+1 in a graph branch beside step n).
        `queue_desc`: the batch is the next row of a device-side queue (ops.fetch_batch's descriptor); the top
        sampler launch fetches it itself and writes it into `nodes_dev`.
        With the bitmap unique, its mark / clear passes run inside the sampler launches either side of it, so a
        2-layer preparation is 5 launches: sample(+fetch, +mark), scan, emit/remap, sample(+clear), aggregate
        (the last one is `_run_agg1`)."""
        if num_rows is not None and injected is not None:
            raise ValueError("injected samples describe a batch of known size")
        if queue_desc is not None and (injected is not None or num_rows is not None):
            raise ValueError("a queued batch is sampled natively and has a static size")
        csr, table, dev = self._state()
        L, k = self.num_layers, self.num_sample
        self_mode = native.SELF_ONCE if self.gcn else native.SELF_DROP
        mode = native.AGG_MEAN if self.agg_func == 'MEAN' else native.AGG_MAX
        self._calls += 1
        layers: List[Optional[_Frontier]] = [None] * (L + 1)
        # `num_rows` (device int32, optional): only the first num_rows[0] entries of nodes_dev are a batch -- a batch
        # extended on the device (UnsupervisedTrainer) has a size the host never learns
        nodes, rows_max = nodes_dev, int(nodes_dev.shape[0])
        # ---- sampling phase, batch outward (src/models.py:249-251) ----
        for l in range(L, 0, -1):
            old = reuse[l - 1] if reuse is not None else None
            fr = _Frontier()
            if old is not None:
                fr.inherit(old, first_layer=(l == 1))
            fr.nodes, fr.num_rows, fr.rows_max = nodes, num_rows, rows_max
            if injected is not None:
                live = rows_max if num_rows is None else int(num_rows.item())
                host_nodes = nodes[:live].cpu().numpy()
                nbr_h, cnt_h, stride = self._injected_lists(host_nodes, injected[L - l], positional=(l == L))
                fr.stride = stride
                fr.nbr = torch.full((rows_max, stride), -1, dtype=torch.int32, device=dev)
                fr.cnt = torch.zeros((rows_max,), dtype=torch.int32, device=dev)
                fr.nbr[:live] = torch.from_numpy(nbr_h).to(dev)
                fr.cnt[:live] = torch.from_numpy(cnt_h).to(dev)
            else:
                fr.stride = self._list_stride()
                offset = (self._calls << 8) | l
                fuse = self._bitmap_ws is not None
                marks = fuse and l > 1                      # feeds a bitmap unique: mark while sampling
                clears = fuse and l == 1 and L > 1          # follows one and marks nothing itself: clear its words
                fr.nbr, fr.cnt = ops.sample_neighbors(csr.rowptr, csr.col, csr.num_nodes, nodes, num_rows, rows_max, k,
                                                      fr.stride, self_mode, self.seed, offset, offset_dev=offset_dev,
                                                      out_nbr=old.nbr if old else None, out_cnt=old.cnt if old else None,
                                                      queue_desc=queue_desc if l == L else None,
                                                      fetch_dst=nodes if (l == L and queue_desc is not None) else None,
                                                      mark_bitmap=self._bitmap_ws if marks else None,
                                                      clear_bitmap=self._bitmap_ws if clears else None,
                                                      prefetch_table=table if (l == 1 and self.l2_prefetch and
                                                                               not isinstance(table, ShardedTable)) else None,
                                                      prefetch_cols=self.input_size)
            if l > 1:   # unique + remap (src/models.py:286-288); the next frontier is U, ascending
                prev = reuse[l - 2] if reuse is not None else None
                outs = dict(uniq=prev.nodes, num_uniq=prev.num_rows, nbr_idx=old.nbr_idx, self_idx=old.self_idx) if old else {}
                if self._bitmap_ws is not None:
                    flags = 0
                    if injected is None:
                        flags = native.UNIQUE_MARKED | (native.UNIQUE_LEAVE_MARKS if l - 1 == 1 else 0)
                    uniq, num_uniq, fr.nbr_idx, fr.self_idx = ops.unique_remap_bitmap(
                        nodes, num_rows, rows_max, fr.nbr, fr.stride, csr.num_nodes, self._bitmap_ws, flags=flags, **outs)
                else:
                    uniq, num_uniq, fr.nbr_idx, fr.self_idx = ops.unique_remap(nodes, num_rows, rows_max, fr.nbr,
                                                                               fr.stride, csr.id_bits, **outs)
                nodes, num_rows = uniq, num_uniq
                rows_max = min(rows_max * (fr.stride + 1), max(csr.num_nodes, 1))
            else:       # layer 1 gathers straight from the feature table by node id: no U0, no remap
                fr.nbr_idx, fr.self_idx = fr.nbr, nodes
                # what _run_agg1 will leave behind, set here as well: under CUDA-graph replay the python objects of
                # the last captured sampling stand for every later batch of the slot
                fr.dim_in = self.input_size
                if isinstance(table, ShardedTable):
                    fr.self_idx = None                      # K4 reads the fp32 self rows the gather emits, in row order
                elif fr.x is not None:
                    self._bind_dense(fr)                    # dense input rows from an earlier _run_agg1(dense_x=True)
                else:
                    fr.table_in = table
            layers[l] = fr
        return layers[1:]

    def _bind_dense(self, fr: _Frontier):
        """Layer 1 reads its operands from the dense rows fr.x = [self | agg] (views, identity row index)."""
        f = self.input_size
        fr.table_in, fr.self_idx = (None, None) if self.gcn else (fr.x[:, :f], None)
        fr.agg, fr.argmax = (fr.x if self.gcn else fr.x[:, f:]), None

    def dense_x_ok(self) -> bool:
        """Whether layer 1 can take dense input rows (gs_agg_fwd_x): a plain fp32 table whose width is a multiple of 4."""
        _, table, _ = self._state()
        return not isinstance(table, ShardedTable) and self.input_size % 4 == 0

    def _run_agg1(self, layers: List[_Frontier], dense_x: bool = False) -> List[_Frontier]:
        """The layer-1 aggregation of the raw features (src/models.py:260, index 1) on sampled frontiers: the
        HBM-bound half of the preparation.  Output buffers a frontier already owns are overwritten in place.
        `dense_x`: write the layer's whole input row [self | agg] and its low tf32 halves (ops.agg_fwd_x) -- a train
        step's layer-1 GEMMs then read dense operands by TMA instead of gathering feature rows on the critical chain."""
        csr, table, dev = self._state()
        mode = native.AGG_MEAN if self.agg_func == 'MEAN' else native.AGG_MAX
        fr = layers[0]
        if dense_x and self.dense_x_ok():
            fr.dim_in = self.input_size
            fr.x, fr.x_lo = ops.agg_fwd_x(table, self.input_size, fr.nbr, fr.stride, fr.cnt, None if self.gcn else fr.nodes,
                                          fr.num_rows, fr.rows_max, mode, x=fr.x, x_lo=fr.x_lo)
            self._bind_dense(fr)
            return layers
        self_rows_buf = fr.table_in if (fr.table_in is not None and fr.table_in is not table) else None
        fr.table_in, fr.dim_in = table, self.input_size
        if isinstance(table, ShardedTable):
            # row-partitioned table: K3 also emits the fp32 self rows (:265), which become K4's self
            # operand with the identity index (forward and backward)
            fr.agg, self_rows = ops.agg_fwd_sharded(table, fr.nbr_idx, fr.stride, fr.cnt, fr.nodes, fr.num_rows,
                                                    fr.rows_max, want_self=not self.gcn, out=fr.agg,
                                                    out_self=self_rows_buf if not self.gcn else None, mode=mode)
            fr.argmax, fr.table_in, fr.self_idx = None, self_rows, None
        else:
            fr.agg, fr.argmax = ops.agg_fwd(table, self.input_size, fr.nbr_idx, fr.stride, fr.cnt, fr.num_rows,
                                            fr.rows_max, mode, out=fr.agg, argmax=fr.argmax)
        return layers

    def _run_compute(self, layers: List[_Frontier], weights: Sequence[torch.Tensor], upto: Optional[int] = None,
                     zero_grad_of_last: Optional[torch.Tensor] = None,
                     weights_lo: Optional[Sequence[Optional[torch.Tensor]]] = None) -> List[_Frontier]:
        """The weight-dependent half (src/models.py:255-267): SageLayer GEMM of every layer and the
        aggregations above layer 1.  `upto` (default: all layers): stop after that many layers -- a trainer that
        runs the top layer fused with the loss (ops.sage_top_sup) computes the layers below it here.
        `zero_grad_of_last`: buffer shaped like the last computed layer's output, zero-filled by that layer's
        GEMM epilogue (the backward scatter of the layer above accumulates into it)."""
        L = self.num_layers if upto is None else int(upto)
        mode = native.AGG_MEAN if self.agg_func == 'MEAN' else native.AGG_MAX
        prec = _PRECISIONS[self.precision]
        for l in range(1, L + 1):
            fr = layers[l - 1]
            if l > 1:
                prev = layers[l - 2]
                fr.table_in, fr.dim_in = prev.h, self.out_size
                fr.agg, fr.argmax = ops.agg_fwd(prev.h, self.out_size, fr.nbr_idx, fr.stride, fr.cnt, fr.num_rows,
                                                fr.rows_max, mode, out=fr.agg, argmax=fr.argmax)
            fr.h = ops.sage_gemm_fwd(None if self.gcn else fr.table_in, fr.self_idx, fr.agg, fr.dim_in, weights[l - 1],
                                     self.out_size, self.gcn, fr.num_rows, fr.rows_max, True, prec, out=fr.h,
                                     zero_out=zero_grad_of_last if l == L else None, l2_normalize=self.normalize,
                                     x_lo=fr.x_lo if l == 1 else None,
                                     weight_lo=weights_lo[l - 1] if weights_lo is not None else None)
        return layers

    def _run_backward(self, layers: List[_Frontier], grad_out: torch.Tensor, weights, needs,
                      *args, **kwargs):
        if self.normalize:
            raise NotImplementedError("normalize=True is a forward-only epilogue (embeddings / inference); "
                                      "the reference's training path has no normalisation (src/models.py:219)")
        return self._run_backward_impl(layers, grad_out, weights, needs, *args, **kwargs)

    def _run_backward_impl(self, layers: List[_Frontier], grad_out: torch.Tensor, weights, needs,
                      grad_bufs=None, own_grad: bool = False, top_masked: bool = False,
                      side_stream: Optional[torch.cuda.Stream] = None,
                      scatter_bufs: Optional[Sequence[Optional[torch.Tensor]]] = None) -> List[Optional[torch.Tensor]]:
        """Weight gradients of all layers.  `grad_bufs` (optional, pre-zeroed) receive them in
        place (static buffers of the captured train step); otherwise fresh tensors are returned.
        `own_grad`: grad_out is a scratch buffer of the caller and may be overwritten.
        `top_masked`: grad_out already carries the last layer's ReLU mask (gs_cls_nll_fwd_bwd).
        `side_stream`: the weight-gradient GEMM of every layer but the first is a leaf of the
        dependency graph (nothing downstream reads dW); with a side stream it runs beside the
        dX -> scatter chain instead of in front of it (a fork/join when captured into a CUDA graph).
        `scatter_bufs[i]` (optional): ZEROED [layers[i].rows_max x pad4(H)] buffer that receives the gradient
        w.r.t. layer i+1's output (see `zeroed_scatter_bufs`); without it the fill runs in line."""
        L, H = len(layers), self.out_size          # a trainer that ran the top layer fused passes the layers below it
        mode = native.AGG_MEAN if self.agg_func == 'MEAN' else native.AGG_MAX
        prec = _PRECISIONS[self.precision]
        # tensor-core path: dZ = grad * (h > 0) is formed once, in place, and the GEMMs run with
        # relu=False (they stream dZ with cp.async); the FFMA path masks while loading its tiles.
        premask = prec != native.PREC_FP32
        g = _padded_table(grad_out)
        if premask and not own_grad and g.data_ptr() == grad_out.data_ptr():
            g = g.clone()                # never overwrite a gradient tensor autograd handed us
        grads: List[Optional[torch.Tensor]] = [None] * L
        keep, forked = [], False         # tensors read on the side stream stay referenced until the join
        lowest = min((i for i in range(L) if needs[i]), default=None)
        if lowest is None:
            return grads
        masked = top_masked              # g already carries the ReLU mask of the layer it belongs to
        for l in range(L, 0, -1):
            fr = layers[l - 1]
            w = weights[l - 1]
            if premask and not masked:
                ops.relu_bwd_inplace(g, fr.h, H, fr.num_rows, fr.rows_max)
            if needs[l - 1]:
                gw = torch.zeros_like(w) if grad_bufs is None else grad_bufs[l - 1]
                fork = side_stream is not None and l - 1 > lowest and grad_bufs is not None
                if fork:
                    main = torch.cuda.current_stream()
                    side_stream.wait_stream(main)
                    keep.append(g)
                    with torch.cuda.stream(side_stream):
                        ops.sage_gemm_bwd_w(None if self.gcn else fr.table_in, fr.self_idx, fr.agg, fr.dim_in, g, fr.h, H,
                                            self.gcn, not premask, fr.num_rows, fr.rows_max, gw, precision=prec)
                    forked = True
                else:
                    ops.sage_gemm_bwd_w(None if self.gcn else fr.table_in, fr.self_idx, fr.agg, fr.dim_in, g, fr.h, H,
                                        self.gcn, not premask, fr.num_rows, fr.rows_max, gw, precision=prec)
                grads[l - 1] = gw
            if l - 1 <= lowest:          # nothing below needs a gradient (raw features never do)
                break
            gs, ga = ops.sage_gemm_bwd_x(g, fr.h, w, fr.dim_in, H, self.gcn, not premask, fr.num_rows, fr.rows_max,
                                         precision=prec)
            prev = layers[l - 2]
            if scatter_bufs is not None and scatter_bufs[l - 2] is not None:
                g_prev = scatter_bufs[l - 2]         # zero-filled by the caller, off the critical path
            else:
                g_prev = torch.zeros((prev.rows_max, ops.pad4(H)), dtype=torch.float32, device=g.device)
            # tensor-core path: the scatter applies layer l-1's ReLU mask itself (mask_table = its output)
            ops.agg_bwd(ga, gs, fr.dim_in, fr.nbr_idx, fr.stride, fr.cnt, fr.self_idx, fr.argmax, fr.num_rows,
                        fr.rows_max, mode, g_prev, mask_table=prev.h if premask else None)
            masked = premask
            g = g_prev
        if forked:
            torch.cuda.current_stream().wait_stream(side_stream)
        return grads

    def zeroed_scatter_bufs(self, layers: List[_Frontier]) -> List[Optional[torch.Tensor]]:
        """Zero-filled targets of the backward scatters (one per layer below the top), allocated on the
        current stream: a trainer calls this on a side stream at the start of the step so the fills run
        beside the forward GEMMs instead of between bwd_x and the scatter."""
        H = self.out_size
        return [torch.zeros((fr.rows_max, ops.pad4(H)), dtype=torch.float32, device=fr.nodes.device)
                for fr in layers[:-1]] + [None]

    # ---- compatibility methods of the reference (slow paths, host round trips) ------------------
    def _get_unique_neighs_list(self, nodes, num_sample=10):                  # :277-289
        """Same return triple as the reference: (list of python sets incl. the node itself,
        {node: index}, unique list).  Unique order is ascending id (the reference's is CPython
        set order; as sets they are equal)."""
        csr, _, dev = self._state()
        nodes_dev = _as_device_ids(nodes, dev)
        n = int(nodes_dev.shape[0])
        self._calls += 1
        if num_sample is None:
            raise NotImplementedError("num_sample=None (full neighbourhood) is not used by the reference forward")
        nbr, cnt = ops.sample_neighbors(csr.rowptr, csr.col, csr.num_nodes, nodes_dev, None, n, int(num_sample),
                                        int(num_sample) + 1, native.SELF_ONCE, self.seed, (self._calls << 8) | 0xFF)
        uniq, num_uniq, _, _ = ops.unique_remap(nodes_dev, None, n, nbr, int(num_sample) + 1, csr.id_bits,
                                                want_nbr_idx=False, want_self_idx=False)
        nbr_h, cnt_h = nbr.cpu().numpy(), cnt.cpu().numpy()
        uniq_list = uniq[:int(num_uniq.item())].cpu().tolist()
        samp = [set(nbr_h[i, :cnt_h[i]].tolist()) for i in range(n)]
        return samp, dict(zip(uniq_list, range(len(uniq_list)))), uniq_list

    def _nodes_map(self, nodes, hidden_embs, neighs):                         # :271-275
        layer_nodes, samp_neighs, layer_nodes_dict = neighs
        assert len(samp_neighs) == len(nodes)
        return [layer_nodes_dict[x] for x in nodes]

    def aggregate(self, nodes, pre_hidden_embs, pre_neighs, num_sample=10):   # :291-330
        """Reference signature; runs the K3 kernel on lists built from the given python sets."""
        unique_nodes_list, samp_neighs, unique_nodes = pre_neighs
        assert len(nodes) == len(samp_neighs)
        _, _, dev = self._state()
        rows = []
        for i, s in enumerate(samp_neighs):
            me = nodes[i]
            assert me in s                                                    # :295-296
            rows.append(sorted(unique_nodes[x] for x in s if self.gcn or x != me))
        stride = max(1, max(len(r) for r in rows))
        nbr = np.full((len(rows), stride), -1, dtype=np.int32)
        cnt = np.asarray([len(r) for r in rows], dtype=np.int32)
        for i, r in enumerate(rows):
            nbr[i, :len(r)] = r
        embed = pre_hidden_embs if len(pre_hidden_embs) == len(unique_nodes) \
            else pre_hidden_embs[torch.as_tensor(unique_nodes_list, device=pre_hidden_embs.device)]   # :300-303
        table = _padded_table(embed.to(dev))
        mode = native.AGG_MEAN if self.agg_func == 'MEAN' else native.AGG_MAX
        out, _ = ops.agg_fwd(table, embed.shape[1], torch.from_numpy(nbr).to(dev), stride, torch.from_numpy(cnt).to(dev),
                             None, len(rows), mode)
        return out[:, :embed.shape[1]]


# ================================================================================================
# UnsupervisedLoss                                                     src/models.py:45-186
# ================================================================================================
class _PairLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, embeddings, owner, mode):
        emb = _padded_table(embeddings)
        dim = embeddings.shape[1]
        p = owner._pairs
        loss, coef_pos, coef_neg, num_active = ops.pair_loss_fwd(emb, dim, p['seed_idx'], p['pos_ptr'], p['pos_idx'],
                                                                 p['neg_ptr'], p['neg_idx'], mode, float(owner.Q),
                                                                 float(owner.MARGIN))
        ctx.save_for_backward(emb, coef_pos, coef_neg, num_active)
        ctx.pairs, ctx.dim = p, dim
        return loss.reshape(()) if mode == 0 else loss        # get_loss_sage: 0-d, get_loss_margin: [1]  (:96, :128)

    @staticmethod
    def backward(ctx, grad_loss):
        emb, coef_pos, coef_neg, num_active = ctx.saved_tensors
        p = ctx.pairs
        grad_emb = torch.zeros_like(emb)
        ops.pair_loss_bwd(emb, ctx.dim, p['seed_idx'], p['pos_ptr'], p['pos_idx'], p['neg_ptr'], p['neg_idx'],
                          coef_pos, coef_neg, num_active, grad_loss.reshape(1).contiguous().float(), grad_emb)
        return (grad_emb[:, :ctx.dim] if grad_emb.shape[1] != ctx.dim else grad_emb), None, None


def negative_radius(nnz: int, num_nodes: int, max_hops: int, budget: int) -> int:
    """Largest h <= max_hops with mean_degree^h <= budget, at least 1 (UnsupervisedLoss.negative_hops)."""
    mean_deg = max(float(nnz) / max(num_nodes, 1), 1.0)
    h, ball = 0, 1.0
    while h < max_hops and ball * mean_deg <= budget:
        ball *= mean_deg
        h += 1
    return max(h, 1)


class UnsupervisedLoss(object):
    """Same surface as src/models.py:45-186.  Sampling (random-walk positives, far negatives,
    batch union) and both losses run on the device; the python pair stores
    (`positive_pairs`, `node_positive_pairs`, ...) are materialised lazily on attribute access."""

    def __init__(self, adj_lists, train_nodes, device, *, seed: Optional[int] = None):
        self.Q = 10                      # :49
        self.N_WALKS = 6                 # :50
        self.WALK_LEN = 1                # :51
        self.N_WALK_LEN = 5              # :52
        self.MARGIN = 3                  # :53
        self.adj_lists = adj_lists
        self.train_nodes = train_nodes
        self.device = device
        self.target_nodes = None
        self.unique_nodes_batch = []
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0x7FFFFFFFFFFFFFFF
        self._calls = 0
        self._dev = None
        self._train_distinct = False
        self._pairs = None
        self._host_pairs = None
        self.neg_hops = None             # None: negative_hops() decides (5 on Cora / Pubmed, as the reference)

    def _state(self):
        if self._dev is None:
            dev = torch.device(self.device)
            if dev.type != 'cuda':
                raise RuntimeError("UnsupervisedLoss (gsage_b200) runs only on a CUDA device; there is no CPU fallback")
            native.load()
            csr = _device_csr(self.adj_lists, None, dev)
            train = torch.from_numpy(np.ascontiguousarray(np.asarray(self.train_nodes), dtype=np.int32)).to(dev)
            is_train = torch.zeros((csr.num_nodes,), dtype=torch.uint8, device=dev)
            is_train[train.long()] = 1
            # the rejection sampler of gs_negative_sample_ex counts far nodes as |train| - |train in ball|: ids must be distinct
            self._train_distinct = int(is_train.sum().item()) == int(train.numel())
            self._dev = (csr, train, is_train, dev)
        return self._dev

    # ---- A8: batch extension (src/models.py:135-148) -------------------------------------------
    def extend_nodes(self, nodes, num_neg=6):
        uniq, num_uniq = self.extend_device(nodes, num_neg)
        self.target_nodes = nodes
        batch = uniq[:int(num_uniq.item())]
        self._pairs['batch'] = batch
        self.unique_nodes_batch = batch.cpu().tolist()
        return self.unique_nodes_batch

    def extend_device(self, nodes, num_neg=6, offset_dev=None):
        """The device half of extend_nodes: draws the pairs and returns (uniq buffer, device count) without a host
        round trip -- the first count[0] entries of the buffer are the extended batch, ascending.  The pair stores
        are installed as in extend_nodes (the python views of them materialise on first access).  `offset_dev`
        (device int64 step counter) is added, shifted by 8, to the samplers' Philox offsets: a captured step draws
        new pairs at every replay."""
        csr, train, is_train, dev = self._state()
        self._calls += 1
        seeds = _as_device_ids(nodes, dev)
        s = int(seeds.shape[0])
        n_pos = self.N_WALKS * self.WALK_LEN
        pos = ops.random_walk_pos(csr.rowptr, csr.col, csr.num_nodes, seeds, self.N_WALKS, self.WALK_LEN, is_train,
                                  self.seed, (self._calls << 8) | 1, offset_dev=offset_dev)      # :169-186
        neg, _ = ops.negative_sample(csr.rowptr, csr.col, csr.num_nodes, seeds, self.negative_hops(), int(num_neg), train,
                                     self.seed, (self._calls << 8) | 2, offset_dev=offset_dev,
                                     is_train=is_train if self._train_distinct else None)      # :153-167
        lists = torch.cat([pos, neg], dim=1).contiguous()
        stride = n_pos + int(num_neg)
        uniq, num_uniq, idx, seed_idx = ops.unique_remap(seeds, None, s, lists, stride, csr.id_bits)    # :146
        self._pairs = dict(seed_idx=seed_idx, pos_idx=idx[:, :n_pos].contiguous().view(-1),
                           neg_idx=idx[:, n_pos:].contiguous().view(-1),
                           pos_ptr=torch.arange(0, (s + 1) * n_pos, n_pos, dtype=torch.int32, device=dev),
                           neg_ptr=torch.arange(0, (s + 1) * int(num_neg), int(num_neg), dtype=torch.int32, device=dev),
                           seeds=seeds, pos=pos, neg=neg)
        self._host_pairs = None
        self.target_nodes = nodes
        return uniq, num_uniq

    #: budget of adjacency entries one seed's ball may span before the ball is shrunk (see negative_hops)
    NEG_BALL_BUDGET = 10_000

    def negative_hops(self) -> int:
        """Radius of the ball a seed's negatives must lie outside.  The reference excludes the N_WALK_LEN = 5 hop
        ball (src/models.py:155-162), which is what runs on Cora and Pubmed (mean degree ~4: ~10^3 entries per
        seed).  On a dense graph that ball is the whole graph -- the reference then dies on an empty `far_nodes`
        (cfg-3, cfg-4) and a per-seed BFS would walk every edge -- so the radius is the largest
        h <= N_WALK_LEN whose expected ball, mean_degree^h, stays within NEG_BALL_BUDGET entries, and at least 1
        (negatives are then train nodes outside the seed's own neighbourhood).  A documented deviation that only
        takes effect where the reference cannot run; set `neg_hops` to force a radius."""
        forced = getattr(self, 'neg_hops', None)
        if forced is not None:
            return int(forced)
        csr = self._state()[0]
        return negative_radius(csr.nnz, csr.num_nodes, self.N_WALK_LEN, self.NEG_BALL_BUDGET)

    def set_pairs(self, unique_nodes_batch, seeds, node_positive_pairs, node_negtive_pairs):
        """Injected-pair mode: install pair stores recorded from a reference run (same
        attribute semantics as src/models.py:59-63) instead of drawing them on the device."""
        _, _, _, dev = self._state()
        index = {int(n): i for i, n in enumerate(unique_nodes_batch)}
        seeds = [int(s) for s in seeds]
        pos_rows = [[index[int(b)] for _, b in node_positive_pairs.get(s, [])] for s in seeds]
        neg_rows = [[index[int(b)] for _, b in node_negtive_pairs.get(s, [])] for s in seeds]

        def csr(rows):
            ptr = np.zeros(len(rows) + 1, dtype=np.int32)
            ptr[1:] = np.cumsum([len(r) for r in rows])
            flat = np.asarray([x for r in rows for x in r] or [0], dtype=np.int32)
            return torch.from_numpy(ptr).to(dev), torch.from_numpy(flat).to(dev)

        pos_ptr, pos_idx = csr(pos_rows)
        neg_ptr, neg_idx = csr(neg_rows)
        seed_idx = torch.tensor([index[s] for s in seeds], dtype=torch.int32, device=dev)
        self._pairs = dict(seed_idx=seed_idx, pos_idx=pos_idx, neg_idx=neg_idx, pos_ptr=pos_ptr, neg_ptr=neg_ptr)
        self._host_pairs = dict(pos=[t for s in seeds for t in node_positive_pairs.get(s, [])],
                                neg=[t for s in seeds for t in node_negtive_pairs.get(s, [])],
                                npos=node_positive_pairs, nneg=node_negtive_pairs)
        self.target_nodes = seeds
        self.unique_nodes_batch = [int(n) for n in unique_nodes_batch]

    def get_positive_nodes(self, nodes):                                      # :150-151
        raise NotImplementedError("call extend_nodes(); positives are drawn on the device (gs_random_walk_pos)")

    def get_negtive_nodes(self, nodes, num_neg):                              # :153
        raise NotImplementedError("call extend_nodes(); negatives are drawn on the device (gs_negative_sample)")

    # ---- lazily materialised python pair stores (reference attribute names, :59-62) -----------
    def _host(self):
        if self._host_pairs is None:
            p = self._pairs
            if p is None:
                return dict(pos=[], neg=[], npos={}, nneg={})
            seeds = p['seeds'].cpu().tolist()
            pos, neg = p['pos'].cpu().tolist(), p['neg'].cpu().tolist()
            npos = {s: [(s, v) for v in row if v >= 0] for s, row in zip(seeds, pos)}
            nneg = {s: [(s, v) for v in row if v >= 0] for s, row in zip(seeds, neg)}
            self._host_pairs = dict(pos=[t for s in seeds for t in npos[s]], neg=[t for s in seeds for t in nneg[s]],
                                    npos=npos, nneg=nneg)
        return self._host_pairs

    positive_pairs = property(lambda self: self._host()['pos'])
    negtive_pairs = property(lambda self: self._host()['neg'])
    node_positive_pairs = property(lambda self: self._host()['npos'])
    node_negtive_pairs = property(lambda self: self._host()['nneg'])

    def _check(self, embeddings, nodes):
        assert len(embeddings) == len(self.unique_nodes_batch)                                  # :66 / :101
        assert np.array_equal(np.asarray(nodes), np.asarray(self.unique_nodes_batch))           # :67 / :102
        native.require_cuda(embeddings, "embeddings")

    # ---- A9 / A10 -------------------------------------------------------------------------------
    def get_loss_sage(self, embeddings, nodes):                               # :65-98
        self._check(embeddings, nodes)
        return _PairLossFn.apply(embeddings, self, 0)

    def get_loss_margin(self, embeddings, nodes):                             # :100-132
        self._check(embeddings, nodes)
        return _PairLossFn.apply(embeddings, self, 1)
