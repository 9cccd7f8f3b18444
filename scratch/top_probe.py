"""Diagnostics (not a bench number): the fused top-layer kernel alone at the cfg-3 shape -- CUDA-event time of a
graph-replayed chain, and (library built with -DGS_TOP_TRACE) the SM-clock phase breakdown of CTA 0."""
import ctypes, os, sys, subprocess
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphsage_b200  # noqa
from graphsage_b200 import native, ops

dev = torch.device('cuda:0')
rows, n_prev, H, C, fan = int(os.environ.get('ROWS', 1024)), 11000, 128, 47, 10
g = torch.Generator(device=dev).manual_seed(0)
table = torch.relu(torch.randn((n_prev, H), generator=g, device=dev))
nbr = torch.randint(0, n_prev, (rows, fan), generator=g, device=dev, dtype=torch.int32)
cnt = torch.full((rows,), fan, dtype=torch.int32, device=dev)
self_idx = torch.randint(0, n_prev, (rows,), generator=g, device=dev, dtype=torch.int32)
w = torch.randn((H, 2 * H), generator=g, device=dev) * 0.1
cw = torch.randn((C, H), generator=g, device=dev) * 0.2
cb = torch.zeros((C,), device=dev)
labels = torch.randint(0, C, (rows,), generator=g, device=dev)
loss = torch.zeros((1,), device=dev)
gcw, gcb, gt = torch.zeros((C, H), device=dev), torch.zeros((C,), device=dev), torch.zeros((n_prev, H), device=dev)
ws = ops.sage_top_workspace(dev)
oh, oa, od = (torch.empty((rows, H), device=dev) for _ in range(3))
prec = native.PREC_TF32X3 if os.environ.get('PREC', 'x3') == 'x3' else native.PREC_TF32

reps = int(os.environ.get('REPS', 8))
rw = torch.zeros((reps - 1, C * H), device=dev) if reps > 1 else None
rb = torch.zeros((reps - 1, 64), device=dev) if reps > 1 else None

def launch():
    ops.sage_top_sup(table, nbr, fan, cnt, self_idx, None, rows, w, False, cw, cb, labels, None, loss, gcw, gcb, gt, ws, prec,
                     out_h=oh, out_agg=oa, out_dz=od, cls_w_rep=rw, cls_b_rep=rb)

for _ in range(3):
    launch()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(20):
        launch()
gr.replay(); torch.cuda.synchronize()
a.record()
for _ in range(5):
    gr.replay()
b.record(); torch.cuda.synchronize()
print(f"top kernel rows={rows} prec={os.environ.get('PREC', 'x3')}: {a.elapsed_time(b) * 1e3 / 100:.2f} us per launch (graph chain of 20, PDL)")
lib = native.load()
if hasattr(lib, 'gs_debug_top_trace_read'):
    launch(); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    lib.gs_debug_top_trace_read(buf, 16)
    t = list(buf)[:12]
    seq = [('pdl', 0, 1), ('idx', 1, 2), ('gather', 2, 3), ('w wait', 3, 4), ('fwd mma', 4, 5), ('logits mma', 5, 10),
           ('softmax', 10, 6), ('dh mma + dz', 6, 11), ('grad Wc/bc', 11, 7), ('dX mma', 7, 8), ('scatter', 8, 9)]
    print('phase cycles (CTA 0, first tile): ' + ', '.join(f"{n} {t[b] - t[a]}" for n, a, b in seq) + f"; total {t[9] - t[0]}")
