"""Diagnostics: the fused update kernel alone at the cfg-3 parameter shapes (world 1): event time of a graph chain and
(trace build) the phase breakdown of CTA 0."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphsage_b200  # noqa
from graphsage_b200 import native
from graphsage_b200.peer import DpExchange
from graphsage_b200.trainer import flat_layout
dev = torch.device('cuda:0')
shapes = [(128, 200), (128, 256), (47, 128), (47,)]
if os.environ.get('DP_SHAPES') == 'cfg1':
    shapes = [(128, 2866), (128, 256), (7, 128), (7,)]
params = [torch.randn(s, device=dev) for s in shapes]
offs, total = flat_layout(shapes)
flat = torch.randn((total,), device=dev)
dp = DpExchange(flat, params, offs, [0, 0, 1, 1])
for _ in range(3):
    dp.update(5.0, 0.7, None)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        dp.update(5.0, 0.7, None)
g.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    g.replay()
b.record(); torch.cuda.synchronize()
print(f"dp_update world 1, {total} floats: {a.elapsed_time(b) * 1e3 / 100:.2f} us per launch (graph chain of 20)")
lib = native.load()
if hasattr(lib, 'gs_debug_dp_trace_read'):
    dp.update(5.0, 0.7, None); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    lib.gs_debug_dp_trace_read(buf, 16)
    t = list(buf)[:8]
    names = ['pdl', 'epoch', 'push+wait', 'loads+partials', 'barrier', 'coef', 'sgd']
    print('phase cycles (CTA 0): ' + ', '.join(f"{n} {t[i + 1] - t[i]}" for i, n in enumerate(names)) + f"; total {t[7] - t[0]}")
