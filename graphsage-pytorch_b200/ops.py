"""Operator-level Python wrappers over the C ABI (one function per entry point of
include/gsage_b200.h).  They allocate outputs with torch (device memory is plumbing), pass
raw pointers and torch's current stream, and raise on any non-zero return code.  No wrapper
has a CPU or eager-PyTorch fallback."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import native
from .native import check, ptr, stream

I32 = torch.int32
F32 = torch.float32


def pad4(n: int) -> int:
    return (int(n) + 3) & ~3


def _lib():
    return native.load()


# ----------------------------------------------------------------------------------------------
# K1
# ----------------------------------------------------------------------------------------------
def sample_neighbors(rowptr, col, num_nodes: int, nodes, num_rows, max_rows: int, k: int, stride: int,
                     self_mode: int, seed: int, offset: int, out_nbr=None, out_cnt=None, offset_dev=None, *,
                     queue_desc=None, fetch_dst=None, mark_bitmap=None, clear_bitmap=None, prefetch_table=None,
                     prefetch_cols: int = 0):
    """src/models.py:279-285 on the device CSR -> (nbr [max_rows, stride] int32, cnt [max_rows] int32).
    Keyword extras fold neighbouring launches of a preparation chain into this one (gs_sample_neighbors_ex):
    `queue_desc` + `fetch_dst` = fetch_batch, `mark_bitmap` / `clear_bitmap` = the mark / clear passes of the bitmap
    unique that follows / preceded, `prefetch_table` (+ `prefetch_cols` floats per row) = an L2 prefetch of the
    table rows of every drawn id and of the node itself, for the gather kernel that follows."""
    dev = (nodes if nodes is not None else fetch_dst).device
    if dev.type != 'cuda':
        native.require_cuda(nodes if nodes is not None else fetch_dst, "nodes")
    if out_nbr is None:
        out_nbr = torch.empty((max_rows, stride), dtype=I32, device=dev)
    if out_cnt is None:
        out_cnt = torch.empty((max_rows,), dtype=I32, device=dev)
    check(_lib().gs_sample_neighbors_ex(ptr(rowptr), ptr(col), num_nodes, ptr(nodes) if queue_desc is None else None,
                                        ptr(num_rows), max_rows, k, stride, self_mode, seed & 0xFFFFFFFFFFFFFFFF,
                                        offset & 0xFFFFFFFFFFFFFFFF, ptr(offset_dev), ptr(out_nbr), ptr(out_cnt),
                                        ptr(queue_desc), ptr(fetch_dst), ptr(mark_bitmap), ptr(clear_bitmap),
                                        ptr(prefetch_table),
                                        prefetch_table.stride(0) * prefetch_table.element_size() if prefetch_table is not None else 0,
                                        pad4(prefetch_cols) * 4 if prefetch_table is not None else 0, stream()),
          "gs_sample_neighbors")
    return out_nbr, out_cnt


def fetch_batch(queue_desc, b_sz: int, dst):
    """dst <- row (next % rows) of the queued batch array; next += 1 (device-resident train loop)."""
    check(_lib().gs_fetch_batch(ptr(queue_desc), b_sz, ptr(dst), stream()), "gs_fetch_batch")
    return dst


# ----------------------------------------------------------------------------------------------
# K2
# ----------------------------------------------------------------------------------------------
def unique_workspace(max_rows: int, stride: int, device) -> torch.Tensor:
    n = int(_lib().gs_unique_workspace_bytes(max_rows, stride))
    return torch.empty((max(n, 16),), dtype=torch.uint8, device=device)


def unique_remap(nodes, num_rows, max_rows: int, nbr, stride: int, id_bits: int, *, uniq=None, num_uniq=None,
                 nbr_idx=None, self_idx=None, workspace=None, want_nbr_idx=True, want_self_idx=True):
    """src/models.py:286-288 (+ the lookups of :306 / :274) -> (uniq, num_uniq, nbr_idx, self_idx)."""
    native.require_cuda(nodes, "nodes")
    dev = nodes.device
    cap = max_rows * (stride + 1)
    if uniq is None:
        uniq = torch.empty((max(cap, 1),), dtype=I32, device=dev)
    if num_uniq is None:
        num_uniq = torch.empty((1,), dtype=I32, device=dev)
    if nbr_idx is None and want_nbr_idx and stride > 0:
        nbr_idx = torch.empty((max_rows, stride), dtype=I32, device=dev)
    if self_idx is None and want_self_idx:
        self_idx = torch.empty((max_rows,), dtype=I32, device=dev)
    if workspace is None:
        workspace = unique_workspace(max_rows, stride, dev)
    check(_lib().gs_unique_remap(ptr(nodes), ptr(num_rows), max_rows, ptr(nbr), stride, id_bits, ptr(uniq),
                                 ptr(num_uniq), ptr(nbr_idx), ptr(self_idx), ptr(workspace), workspace.numel(),
                                 stream()), "gs_unique_remap")
    return uniq, num_uniq, nbr_idx, self_idx


def unique_bitmap_workspace(num_nodes: int, device) -> torch.Tensor:
    """Zero-initialised scratch of the bitmap path; allocate once per graph and reuse (every
    call leaves the bitmap part zeroed again)."""
    n = int(_lib().gs_unique_bitmap_workspace_bytes(num_nodes))
    return torch.zeros((max(n, 16),), dtype=torch.uint8, device=device)


def unique_remap_bitmap(nodes, num_rows, max_rows: int, nbr, stride: int, num_nodes: int, workspace, *, uniq=None,
                        num_uniq=None, nbr_idx=None, self_idx=None, want_nbr_idx=True, want_self_idx=True, flags: int = 0):
    """Same contract and outputs as unique_remap, via the bitmap/rank path (gs_unique_remap_bitmap_ex).  `flags`:
    native.UNIQUE_MARKED / UNIQUE_LEAVE_MARKS when the neighbouring sampler launches do the mark / clear passes."""
    native.require_cuda(nodes, "nodes")
    dev = nodes.device
    cap = min(max_rows * (stride + 1), max(num_nodes, 1))
    if uniq is None:
        uniq = torch.empty((max(cap, 1),), dtype=I32, device=dev)
    if num_uniq is None:
        num_uniq = torch.empty((1,), dtype=I32, device=dev)
    if nbr_idx is None and want_nbr_idx and stride > 0:
        nbr_idx = torch.empty((max_rows, stride), dtype=I32, device=dev)
    if self_idx is None and want_self_idx:
        self_idx = torch.empty((max_rows,), dtype=I32, device=dev)
    check(_lib().gs_unique_remap_bitmap_ex(ptr(nodes), ptr(num_rows), max_rows, ptr(nbr), stride, num_nodes, ptr(uniq),
                                           ptr(num_uniq), ptr(nbr_idx), ptr(self_idx), ptr(workspace), workspace.numel(),
                                           int(flags), stream()), "gs_unique_remap_bitmap")
    return uniq, num_uniq, nbr_idx, self_idx


# ----------------------------------------------------------------------------------------------
# K3
# ----------------------------------------------------------------------------------------------
def agg_fwd(table, dim: int, nbr, stride: int, cnt, num_rows, max_rows: int, mode: int, out=None, argmax=None):
    """src/models.py:300-326 -> out [max_rows, pad4(dim)] (and argmax for MAX)."""
    native.require_cuda(table, "table")
    ld_out = pad4(dim)
    if out is None:
        out = torch.empty((max_rows, ld_out), dtype=F32, device=table.device)
    if mode == native.AGG_MAX and argmax is None:
        argmax = torch.empty((max_rows, ld_out), dtype=I32, device=table.device)
    check(_lib().gs_agg_fwd(ptr(table), table.stride(0), dim, ptr(nbr), stride, ptr(cnt), ptr(num_rows), max_rows, mode,
                            ptr(out), out.stride(0), ptr(argmax), argmax.stride(0) if argmax is not None else 0,
                            stream()), "gs_agg_fwd")
    return out, argmax


def pad8(n: int) -> int:
    return (int(n) + 7) & ~7


def agg_fwd_sharded(table, nbr, stride: int, cnt, self_nodes, num_rows, max_rows: int, want_self: bool = True,
                    out=None, out_self=None, mode: int = native.AGG_MEAN):
    """MEAN or MAX over a row-partitioned bf16 table (peer.ShardedTable), global node ids in `nbr`
    -> (agg fp32 [max_rows, pad8(dim)], self rows fp32 or None).  Remote shards are read over NVLink."""
    native.require_cuda(nbr, "nbr")
    ld = pad8(table.dim)
    if out is None:
        out = torch.empty((max_rows, ld), dtype=F32, device=nbr.device)
    if want_self and out_self is None:
        out_self = torch.empty((max_rows, ld), dtype=F32, device=nbr.device)
    check(_lib().gs_agg_fwd_bf16_sharded(table.bases, table.num_shards, table.rows_per_shard, table.ld, table.dim,
                                         ptr(nbr), stride, ptr(cnt), ptr(self_nodes) if want_self else None,
                                         ptr(num_rows), max_rows, ptr(out), out.stride(0),
                                         ptr(out_self) if want_self else None,
                                         out_self.stride(0) if want_self else 0, mode, stream()), "gs_agg_fwd_bf16_sharded")
    return out, (out_self if want_self else None)


def agg_bwd(grad_agg, grad_self, dim: int, nbr, stride: int, cnt, self_idx, argmax, num_rows, max_rows: int,
            mode: int, grad_table, mask_table=None):
    """Scatter backward of K3 (+ self rows).  `mask_table`: ReLU output of the rows scattered into;
    with it grad_table receives d(pre-activation) and no separate relu_bwd pass is needed."""
    check(_lib().gs_agg_bwd(ptr(grad_agg), grad_agg.stride(0) if grad_agg is not None else 0,
                            ptr(grad_self), grad_self.stride(0) if grad_self is not None else 0, dim,
                            ptr(nbr), stride, ptr(cnt), ptr(self_idx), ptr(argmax),
                            argmax.stride(0) if argmax is not None else 0, ptr(num_rows), max_rows, mode,
                            ptr(grad_table), grad_table.stride(0), ptr(mask_table),
                            mask_table.stride(0) if mask_table is not None else 0, stream()), "gs_agg_bwd")
    return grad_table


# ----------------------------------------------------------------------------------------------
# K4
# ----------------------------------------------------------------------------------------------
def agg_fwd_x(table, dim: int, nbr, stride: int, cnt, self_nodes, num_rows, max_rows: int, mode: int, x=None, x_lo=None,
              want_lo: bool = True):
    """K3 writing the layer's dense input row X = [table[self] | agg] (gs_agg_fwd_x) and its low halves.
    self_nodes None (gcn): X = agg only.  Returns (x [max_rows, K], x_lo or None) with K = pad4(dim) * (1 or 2)."""
    native.require_cuda(table, "table")
    d4 = pad4(dim)
    k = d4 if self_nodes is None else 2 * d4
    kp = (k + 31) & ~31          # row pitch: a multiple of 128 bytes, so the 128-byte pieces a TMA box reads are whole lines
    if x is None:
        x = torch.empty((max_rows, kp), dtype=F32, device=table.device)[:, :k]
    if x_lo is None and want_lo:
        x_lo = torch.empty((max_rows, kp), dtype=F32, device=table.device)[:, :k]
    check(_lib().gs_agg_fwd_x(ptr(table), table.stride(0), dim, ptr(nbr), stride, ptr(cnt), ptr(self_nodes), ptr(num_rows),
                              max_rows, mode, ptr(x), x.stride(0), 0 if self_nodes is None else d4, ptr(x_lo), stream()),
          "gs_agg_fwd_x")
    return x, x_lo


def split_lo(src, dst=None):
    """dst = src - trunc_tf32(src) (the low half of K4's 3-term split); the fused update kernel keeps it current."""
    if dst is None:
        dst = torch.empty_like(src)
    check(_lib().gs_split_lo(ptr(src), ptr(dst), src.numel(), stream()), "gs_split_lo")
    return dst


def sage_gemm_fwd(self_table, self_idx, agg, dim: int, weight, out_dim: int, gcn: bool, num_rows, max_rows: int,
                  relu: bool = True, precision: int = native.PREC_FP32, out=None, zero_out=None, x_lo=None, weight_lo=None,
                  l2_normalize: bool = False):
    """src/models.py:215-219 -> out [max_rows, pad4(out_dim)].  `zero_out` (optional, same shape as out) is
    zero-filled on the way: the buffer the backward of the layer above scatters d(out) into."""
    native.require_cuda(agg, "agg")
    if out is None:
        ld = pad4(out_dim)
        alloc = torch.empty if ld == out_dim else torch.zeros
        out = alloc((max_rows, ld), dtype=F32, device=agg.device)
    check(_lib().gs_sage_gemm_fwd_ex(ptr(self_table), self_table.stride(0) if self_table is not None else 0, ptr(self_idx),
                                     ptr(agg), agg.stride(0), dim, ptr(weight), weight.stride(0), out_dim, int(gcn),
                                     ptr(num_rows), max_rows, ptr(out), out.stride(0), int(relu), precision,
                                     ptr(zero_out), zero_out.stride(0) if zero_out is not None else 0, ptr(x_lo),
                                     ptr(weight_lo), int(l2_normalize), stream()),
          "gs_sage_gemm_fwd")
    return out


TOP_H, TOP_MAX_CLASSES, TOP_MAX_STRIDE = 128, 64, 16


def sage_top_supported(dim: int, out_dim: int, num_classes: int, stride: int, precision: int, mean: bool) -> bool:
    """Whether gs_sage_top_sup covers this top layer (otherwise it runs as separate kernels)."""
    return (mean and dim == TOP_H and out_dim == TOP_H and 1 <= num_classes <= TOP_MAX_CLASSES
            and 1 <= stride <= TOP_MAX_STRIDE and precision in (native.PREC_TF32, native.PREC_TF32X3))


def sage_top_workspace(device) -> torch.Tensor:
    return torch.zeros((int(_lib().gs_sage_top_workspace_bytes()),), dtype=torch.uint8, device=device)


def sage_top_sup(table, nbr_idx, stride: int, cnt, self_idx, num_rows, max_rows: int, weight, gcn: bool, cls_w, cls_b,
                 labels, label_index, loss, grad_cls_w, grad_cls_b, grad_table, workspace, precision: int, *,
                 out_h=None, out_agg=None, out_dz=None, logp=None, cls_w_rep=None, cls_b_rep=None, out_dlog=None):
    """The top SageLayer + classifier + NLL, forward and backward, in one launch (gs_sage_top_sup).
    Returns (h, agg, dz): the layer's output, and the B / A operands of its weight-gradient GEMM.
    `cls_w_rep` [R-1, C*128] / `cls_b_rep` [R-1, 64] (zeroed): more replicas of the classifier gradients the CTAs spread
    their atomic adds over; the fused update folds them in (peer.DpExchange(extras=...)).
    `out_dlog` [max_rows, 64] with grad_cls_w None: d(logits) is saved instead and the classifier's weight gradient is a
    problem of sage_gemm_bwd_w_group."""
    native.require_cuda(table, "table")
    dev = table.device
    H = TOP_H
    if out_h is None:
        out_h = torch.empty((max_rows, H), dtype=F32, device=dev)
    if out_agg is None:
        out_agg = torch.empty((max_rows, H), dtype=F32, device=dev)
    if out_dz is None:
        out_dz = torch.empty((max_rows, H), dtype=F32, device=dev)
    classes = int(cls_w.shape[0])
    check(_lib().gs_sage_top_sup(ptr(table), table.stride(0), ptr(nbr_idx), stride, ptr(cnt), ptr(self_idx), ptr(num_rows),
                                 max_rows, ptr(weight), weight.stride(0), H, H, int(gcn), ptr(cls_w), ptr(cls_b), classes,
                                 ptr(labels), ptr(label_index), ptr(out_h), out_h.stride(0), ptr(out_agg), out_agg.stride(0),
                                 ptr(out_dz), out_dz.stride(0), ptr(out_dlog), out_dlog.stride(0) if out_dlog is not None else 0,
                                 ptr(logp), ptr(loss), ptr(grad_cls_w), ptr(grad_cls_b),
                                 ptr(grad_table), grad_table.stride(0) if grad_table is not None else 0, ptr(workspace),
                                 workspace.numel(), precision, ptr(cls_w_rep), ptr(cls_b_rep),
                                 1 + (int(cls_w_rep.shape[0]) if cls_w_rep is not None else 0), stream()), "gs_sage_top_sup")
    return out_h, out_agg, out_dz


def sage_gemm_bwd_w(self_table, self_idx, agg, dim: int, grad_out, out, out_dim: int, gcn: bool, relu: bool, num_rows,
                    max_rows: int, grad_w, precision: int = native.PREC_FP32):
    check(_lib().gs_sage_gemm_bwd_w(ptr(self_table), self_table.stride(0) if self_table is not None else 0,
                                    ptr(self_idx), ptr(agg), agg.stride(0), dim, ptr(grad_out), grad_out.stride(0),
                                    ptr(out), out.stride(0) if out is not None else 0, out_dim, int(gcn), int(relu),
                                    ptr(num_rows), max_rows, ptr(grad_w), grad_w.stride(0), precision, stream()),
          "gs_sage_gemm_bwd_w")
    return grad_w


def sage_gemm_bwd_w_group(problems, precision: int):
    """Up to three weight-gradient problems in one launch (gs_sage_gemm_bwd_w_group).  `problems`: tuples
    (self_table, self_idx, agg, dim, grad_out, out, out_dim, num_rows, max_rows, grad_w, gcn, relu, grad_out_cols);
    the first ten as for sage_gemm_bwd_w, grad_out_cols = 0 or the zero-padded readable width of grad_out's rows."""
    import ctypes
    n = len(problems)
    assert 1 <= n <= 3
    vp, i64, i32 = ctypes.c_void_p * n, ctypes.c_int64 * n, ctypes.c_int32 * n
    st, si, ag, dim, go, out, od, nr, mr, gw, gcn, relu, cols = list(zip(*problems))
    ld = lambda ts: i64(*[int(t.stride(0)) if t is not None else 0 for t in ts])
    ints = lambda xs: i32(*[int(x) for x in xs])
    check(_lib().gs_sage_gemm_bwd_w_group(n, vp(*[ptr(t) for t in st]), ld(st), vp(*[ptr(t) for t in si]),
                                          vp(*[ptr(t) for t in ag]), ld(ag), ints(dim), vp(*[ptr(t) for t in go]), ld(go),
                                          vp(*[ptr(t) for t in out]), ld(out), ints(od), ints(cols), ints(gcn), ints(relu),
                                          vp(*[ptr(t) for t in nr]), ints(mr), vp(*[ptr(t) for t in gw]), ld(gw),
                                          precision, stream()), "gs_sage_gemm_bwd_w_group")


def sage_gemm_bwd_w_pair(problems, gcn: bool, relu: bool, precision: int):
    """Two weight-gradient problems with common gcn / relu in one launch.  `problems`: two tuples
    (self_table, self_idx, agg, dim, grad_out, out, out_dim, num_rows, max_rows, grad_w), as for sage_gemm_bwd_w."""
    assert len(problems) == 2
    sage_gemm_bwd_w_group([tuple(q) + (int(gcn), int(relu), 0) for q in problems], precision)


def sage_gemm_bwd_x(grad_out, out, weight, dim: int, out_dim: int, gcn: bool, relu: bool, num_rows, max_rows: int,
                    grad_self=None, grad_agg=None, precision: int = native.PREC_FP32):
    dev = grad_out.device
    ld = pad4(dim)
    alloc = torch.empty if ld == dim else torch.zeros
    if grad_agg is None:
        grad_agg = alloc((max_rows, ld), dtype=F32, device=dev)
    if not gcn and grad_self is None:
        grad_self = alloc((max_rows, ld), dtype=F32, device=dev)
    check(_lib().gs_sage_gemm_bwd_x(ptr(grad_out), grad_out.stride(0), ptr(out), out.stride(0) if out is not None else 0,
                                    ptr(weight), weight.stride(0), dim, out_dim, int(gcn), int(relu), ptr(num_rows),
                                    max_rows, ptr(grad_self), grad_self.stride(0) if grad_self is not None else 0,
                                    ptr(grad_agg), grad_agg.stride(0), precision, stream()), "gs_sage_gemm_bwd_x")
    return grad_self, grad_agg


def relu_bwd_inplace(grad, out, dim: int, num_rows, max_rows: int):
    """grad[r, c] = 0 where out[r, c] <= 0 (in place)."""
    check(_lib().gs_relu_bwd_inplace(ptr(grad), grad.stride(0), ptr(out), out.stride(0), dim, ptr(num_rows), max_rows,
                                     stream()), "gs_relu_bwd_inplace")
    return grad


# ----------------------------------------------------------------------------------------------
# classifier / loss / update
# ----------------------------------------------------------------------------------------------
def cls_fwd(emb, dim: int, weight, bias, num_classes: int, logp=None, precision: int = native.PREC_TF32X3):
    native.require_cuda(emb, "embeds")
    rows = emb.shape[0]
    if logp is None:
        logp = torch.empty((rows, num_classes), dtype=F32, device=emb.device)
    check(_lib().gs_cls_fwd(ptr(emb), emb.stride(0), rows, dim, ptr(weight), ptr(bias), num_classes, ptr(logp),
                            precision, stream()), "gs_cls_fwd")
    return logp


def cls_bwd(grad_logp, logp, emb, dim: int, weight, num_classes: int, grad_emb, grad_w, grad_b, scratch=None,
            precision: int = native.PREC_TF32X3):
    rows = emb.shape[0]
    if scratch is None:
        scratch = torch.empty((rows, num_classes), dtype=F32, device=emb.device)
    check(_lib().gs_cls_bwd(ptr(grad_logp), ptr(logp), ptr(emb), emb.stride(0), rows, dim, ptr(weight), num_classes,
                            ptr(grad_emb), grad_emb.stride(0) if grad_emb is not None else 0, ptr(grad_w), ptr(grad_b),
                            ptr(scratch), precision, stream()), "gs_cls_bwd")


def nll_fwd_bwd(logp, labels, loss=None, grad_logp=None, want_grad=True, label_index=None):
    rows, classes = logp.shape
    if loss is None:
        loss = torch.empty((1,), dtype=F32, device=logp.device)
    if grad_logp is None and want_grad:
        grad_logp = torch.empty_like(logp)
    check(_lib().gs_nll_fwd_bwd(ptr(logp), ptr(labels), ptr(label_index), rows, classes, ptr(loss), ptr(grad_logp),
                                stream()),
          "gs_nll_fwd_bwd")
    return loss, grad_logp


def cls_nll_fwd_bwd(emb, dim: int, weight, bias, num_classes: int, labels, label_index, loss, grad_emb, grad_w, grad_b,
                    logp=None, scratch=None, precision: int = native.PREC_TF32X3, mask_relu_input: bool = False,
                    zero_loss: bool = True, num_rows=None):
    """Classifier + NLL(mean) forward and backward in one call (src/models.py:25-27, src/utils.py:153,162-163).
    `zero_loss=False`: the caller zeroed `loss` already (a trainer does it beside the forward GEMMs).
    `num_rows` (device int32): only the first num_rows[0] rows of `emb` are a batch."""
    rows = emb.shape[0]
    if logp is None:
        logp = torch.empty((rows, num_classes), dtype=F32, device=emb.device)
    if scratch is None:
        scratch = torch.empty((rows, num_classes), dtype=F32, device=emb.device)
    check(_lib().gs_cls_nll_fwd_bwd(ptr(emb), emb.stride(0), rows, dim, ptr(weight), ptr(bias), num_classes, ptr(labels),
                                    ptr(label_index), ptr(logp), ptr(loss), ptr(grad_emb),
                                    grad_emb.stride(0) if grad_emb is not None else 0, ptr(grad_w), ptr(grad_b),
                                    ptr(scratch), int(mask_relu_input), int(zero_loss), ptr(num_rows), precision, stream()),
          "gs_cls_nll_fwd_bwd")
    return logp


class TensorList:
    """Device-side (param, grad, numel) table of one model for gs_clip_sgd."""

    def __init__(self, params, grads):
        dev = params[0].device
        self.params, self.grads = list(params), list(grads)
        self.p = torch.tensor([p.data_ptr() for p in params], dtype=torch.int64, device=dev)
        self.g = torch.tensor([g.data_ptr() for g in grads], dtype=torch.int64, device=dev)
        self.n = torch.tensor([p.numel() for p in params], dtype=torch.int64, device=dev)
        self.max_numel = max(p.numel() for p in params)
        self.scratch = torch.zeros((1,), dtype=F32, device=dev)


def clip_sgd(tl: TensorList, max_norm: float, lr: float, grad_div: float = 1.0, zero_grads: bool = False):
    check(_lib().gs_clip_sgd(ptr(tl.p), ptr(tl.g), ptr(tl.n), len(tl.params), tl.max_numel, float(max_norm), float(lr),
                             float(grad_div), int(zero_grads), ptr(tl.scratch), stream()), "gs_clip_sgd")


# ----------------------------------------------------------------------------------------------
# K5 / K6
# ----------------------------------------------------------------------------------------------
def random_walk_pos(rowptr, col, num_nodes: int, seeds, n_walks: int, walk_len: int, is_train, seed: int, offset: int,
                    out=None, offset_dev=None):
    n = seeds.shape[0]
    if out is None:
        out = torch.empty((n, n_walks * walk_len), dtype=I32, device=seeds.device)
    check(_lib().gs_random_walk_pos(ptr(rowptr), ptr(col), num_nodes, ptr(seeds), n, n_walks, walk_len, ptr(is_train),
                                    seed & 0xFFFFFFFFFFFFFFFF, offset & 0xFFFFFFFFFFFFFFFF, ptr(offset_dev), ptr(out),
                                    stream()), "gs_random_walk_pos")
    return out


def negative_workspace_bytes(num_nodes: int, num_seeds: int) -> int:
    return int(_lib().gs_negative_workspace_bytes(num_nodes, num_seeds))


def negative_sample(rowptr, col, num_nodes: int, seeds, hops: int, num_neg: int, train_nodes, seed: int, offset: int,
                    workspace=None, out=None, out_cnt=None, offset_dev=None, is_train=None):
    n = seeds.shape[0]
    dev = seeds.device
    if workspace is None:
        workspace = torch.empty((negative_workspace_bytes(num_nodes, n),), dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.empty((n, num_neg), dtype=I32, device=dev)
    if out_cnt is None:
        out_cnt = torch.empty((n,), dtype=I32, device=dev)
    check(_lib().gs_negative_sample_ex(ptr(rowptr), ptr(col), num_nodes, ptr(seeds), n, hops, num_neg, ptr(train_nodes),
                                       train_nodes.shape[0], ptr(is_train), seed & 0xFFFFFFFFFFFFFFFF,
                                       offset & 0xFFFFFFFFFFFFFFFF, ptr(offset_dev), ptr(out), ptr(out_cnt), ptr(workspace),
                                       workspace.numel(), stream()), "gs_negative_sample")
    return out, out_cnt


def pair_loss_fwd(emb, dim: int, seed_idx, pos_ptr, pos_idx, neg_ptr, neg_idx, mode: int, q: float, margin: float):
    dev = emb.device
    loss = torch.empty((1,), dtype=F32, device=dev)
    coef_pos = torch.empty((pos_idx.numel(),), dtype=F32, device=dev)
    coef_neg = torch.empty((neg_idx.numel(),), dtype=F32, device=dev)
    scratch = torch.empty((1,), dtype=F32, device=dev)
    num_active = torch.empty((1,), dtype=I32, device=dev)
    check(_lib().gs_pair_loss_fwd(ptr(emb), emb.stride(0), dim, ptr(seed_idx), seed_idx.shape[0], ptr(pos_ptr),
                                  ptr(pos_idx), ptr(neg_ptr), ptr(neg_idx), mode, float(q), float(margin), ptr(loss),
                                  ptr(coef_pos), ptr(coef_neg), ptr(scratch), ptr(num_active), stream()),
          "gs_pair_loss_fwd")
    return loss, coef_pos, coef_neg, num_active


def pair_loss_bwd(emb, dim: int, seed_idx, pos_ptr, pos_idx, neg_ptr, neg_idx, coef_pos, coef_neg, num_active,
                  grad_loss, grad_emb):
    check(_lib().gs_pair_loss_bwd(ptr(emb), emb.stride(0), dim, ptr(seed_idx), seed_idx.shape[0], ptr(pos_ptr),
                                  ptr(pos_idx), ptr(neg_ptr), ptr(neg_idx), ptr(coef_pos), ptr(coef_neg),
                                  ptr(num_active), ptr(grad_loss), ptr(grad_emb), grad_emb.stride(0), stream()),
          "gs_pair_loss_bwd")
    return grad_emb
