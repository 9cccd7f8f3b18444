"""Diagnostics: the two branches of the pipelined step, each replayed ALONE from a CUDA graph at the headline workload
(train chain on an already prepared slot; preparation chain on queued batches), and both together."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import graphsage_b200  # noqa
from graphsage_b200 import models, native
from graphsage_b200.graph import AdjCSR
from graphsage_b200.trainer import PipelinedTrainer
import graphsage_b200.synth as synth

dev = torch.device('cuda:0')
scale = float(os.environ.get('SCALE', '1.0'))
cfg, rowptr, col, feats, labels, train = bench.build_workload(scale)
torch.manual_seed(824)
model = models.GraphSage(2, cfg['feats'], cfg['hidden'], torch.from_numpy(feats).to(dev), AdjCSR(rowptr, col), dev, seed=824).to(dev)
cls = models.Classification(cfg['hidden'], cfg['classes']).to(dev)
tr = PipelinedTrainer(model, cls, labels, 1024, lr=0.0)
batches = torch.from_numpy(bench.batches_for(train, 1024, 64, 0, 1).astype(np.int32)).to(dev)
tr.set_queue(batches)
tr.prime()
tr.run(8)
torch.cuda.synchronize()

def timed(graph, n_inside, reps=20):
    graph.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        graph.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * n_inside)

N = 10
g_train = torch.cuda.CUDAGraph()
with torch.cuda.graph(g_train):
    for _ in range(N):
        tr._compute(0)
print(f"train chain alone : {timed(g_train, N):7.2f} us per step ({tr.launches_per_step} launches per full step)")
for pdl in (0, 1):
    g_prep = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_prep):
        native.set_pdl(pdl); native.set_background(True)
        for _ in range(N):
            tr._sample(1); tr._aggregate(1)
        native.set_pdl(-1); native.set_background(False)
    print(f"prep chain alone  : {timed(g_prep, N):7.2f} us per step (pdl {pdl}, background carveout)")
g_prep = torch.cuda.CUDAGraph()
with torch.cuda.graph(g_prep):
    native.set_pdl(1)
    for _ in range(N):
        tr._sample(1); tr._aggregate(1)
    native.set_pdl(-1)
print(f"prep chain alone  : {timed(g_prep, N):7.2f} us per step (pdl 1, default carveout)")
g_both = torch.cuda.CUDAGraph()
with torch.cuda.graph(g_both):
    for i in range(N):
        tr._both(i % 3)
print(f"both (fork/join)  : {timed(g_both, N):7.2f} us per step")
# individual kernels of the train chain, each as its own graph chain
from graphsage_b200 import ops
from graphsage_b200.models import _PRECISIONS
layers = tr.slot_layers[0]
csr, table, _ = model._state()
w = [x.detach() for x in tr.weights]
prec = _PRECISIONS[model.precision]
below, top = layers
def chain(fn, n=N):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    return timed(g, n)
print(f"  fwd1 gemm       : {chain(lambda: model._run_compute(layers, w, upto=1, zero_grad_of_last=below.gh, weights_lo=tr.weights_lo)):7.2f} us  (dense/TMA path: {tr.dense_x1})")
lib = native.load()
if hasattr(lib, 'gs_debug_tma_trace_read'):
    import ctypes
    model._run_compute(layers, w, upto=1, zero_grad_of_last=below.gh, weights_lo=tr.weights_lo); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 32)(); lib.gs_debug_tma_trace_read(buf, 32); t = list(buf)
    print('  TMA fwd trace (CTA 0, cycles from start): pdl', t[1] - t[0], 'setup', t[2] - t[0], 'stage ready at', [t[8 + i] - t[0] for i in range(7)],
          'acc', t[3] - t[0], 'staged', t[4] - t[0], 'done', t[5] - t[0])
print(f"  fwd1 gemm, gathered-operand kernel : {chain(lambda: ops.sage_gemm_fwd(table, below.nodes, below.agg.contiguous() if tr.dense_x1 else below.agg, 100, w[0], 128, False, below.num_rows, below.rows_max, True, prec)):7.2f} us")
print(f"  top kernel      : {chain(lambda: ops.sage_top_sup(below.h, top.nbr_idx, top.stride, top.cnt, top.self_idx, None, top.rows_max, w[1], False, tr.cls_w.detach(), tr.cls_b.detach(), tr.labels, tr.slot_seeds[0], tr.loss, tr.grads[2], tr.grads[3], below.gh, tr._top_ws, prec, out_h=top.h, out_agg=top.agg, out_dz=top.dz)):7.2f} us")
csr, table, _ = model._state()
print(f"  dW pair         : {chain(lambda: ops.sage_gemm_bwd_w_pair([(below.table_in, below.self_idx, below.agg, below.dim_in, below.gh, below.h, 128, below.num_rows, below.rows_max, tr.grads[0]), (below.h, top.self_idx, top.agg, 128, top.dz, top.h, 128, top.num_rows, top.rows_max, tr.grads[1])], False, False, prec)):7.2f} us")
print(f"  dW layer 1 only : {chain(lambda: ops.sage_gemm_bwd_w(below.table_in, below.self_idx, below.agg, below.dim_in, below.gh, below.h, 128, False, False, below.num_rows, below.rows_max, tr.grads[0], precision=prec)):7.2f} us")
print(f"  update          : {chain(lambda: tr.dp.update(5.0, 0.0, None)):7.2f} us")
csr, table, _ = model._state()
print(f"  sampler L2      : {chain(lambda: ops.sample_neighbors(csr.rowptr, csr.col, csr.num_nodes, tr.slot_seeds[0], None, 1024, 10, 10, native.SELF_DROP, 1, 5)):7.2f} us")
print(f"  sampler L1      : {chain(lambda: ops.sample_neighbors(csr.rowptr, csr.col, csr.num_nodes, below.nodes, below.num_rows, below.rows_max, 10, 10, native.SELF_DROP, 1, 5, out_nbr=below.nbr, out_cnt=below.cnt)):7.2f} us")
agg_plain = torch.empty((below.rows_max, 100), device=dev)
print(f"  agg L1          : {chain(lambda: ops.agg_fwd(table, 100, below.nbr, below.stride, below.cnt, below.num_rows, below.rows_max, native.AGG_MEAN, out=agg_plain)):7.2f} us")
xx, xlo = torch.empty((below.rows_max, 224), device=dev)[:, :200], torch.empty((below.rows_max, 224), device=dev)[:, :200]
print(f"  agg L1 -> X, X_lo: {chain(lambda: ops.agg_fwd_x(table, 100, below.nbr, below.stride, below.cnt, below.nodes, below.num_rows, below.rows_max, native.AGG_MEAN, x=xx, x_lo=xlo)):7.2f} us")
print(f"  agg L1 -> X only : {chain(lambda: ops.agg_fwd_x(table, 100, below.nbr, below.stride, below.cnt, below.nodes, below.num_rows, below.rows_max, native.AGG_MEAN, x=xx, want_lo=False)):7.2f} us")

# ---- which part of the preparation branch costs the train chain its time? ----
real_agg, real_sample, real_unique, real_aggx = ops.agg_fwd, ops.sample_neighbors, ops.unique_remap_bitmap, ops.agg_fwd_x
def both_graph():
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(N):
            tr._both(i % 3)
    return timed(g, N)
ops.agg_fwd = lambda table, dim, nbr, stride, cnt, num_rows, max_rows, mode, out=None, argmax=None: (out, argmax); ops.agg_fwd_x = lambda table, dim, nbr, stride, cnt, sn, num_rows, max_rows, mode, x=None, x_lo=None, want_lo=True: (x, x_lo)
print(f"both, prep without the aggregation launch : {both_graph():7.2f} us per step")
ops.agg_fwd = real_agg; ops.agg_fwd_x = real_aggx
def fake_sample(rowptr, col, n, nodes, num_rows, max_rows, k, stride, *a, out_nbr=None, out_cnt=None, **kw):
    return out_nbr, out_cnt
def fake_unique(nodes, num_rows, max_rows, nbr, stride, n, ws, uniq=None, num_uniq=None, nbr_idx=None, self_idx=None, **kw):
    return uniq, num_uniq, nbr_idx, self_idx
ops.sample_neighbors, ops.unique_remap_bitmap = fake_sample, fake_unique
print(f"both, prep = the aggregation launch only  : {both_graph():7.2f} us per step")
ops.agg_fwd = lambda table, dim, nbr, stride, cnt, num_rows, max_rows, mode, out=None, argmax=None: (out, argmax); ops.agg_fwd_x = lambda table, dim, nbr, stride, cnt, sn, num_rows, max_rows, mode, x=None, x_lo=None, want_lo=True: (x, x_lo)
print(f"both, empty prep branch (fork/join only)  : {both_graph():7.2f} us per step")
ops.agg_fwd, ops.sample_neighbors, ops.unique_remap_bitmap, ops.agg_fwd_x = real_agg, real_sample, real_unique, real_aggx

