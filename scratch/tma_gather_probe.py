"""Diagnostics: the layer-1 forward GEMM at the cfg-3 shape, gathered-TMA kernel against the thread-staged one (set
GS_TMA_GATHER=2 / 0 in the environment): event time of a graph chain, and the result against torch fp64."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphsage_b200  # noqa
from graphsage_b200 import native, ops
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
n_tab, rows, mx, dim, H = 200000, 10500, 11264, 100, 128
tab = torch.from_numpy(rng.standard_normal((n_tab, dim)).astype(np.float32)).to(dev)
sidx = torch.from_numpy(rng.integers(0, n_tab, size=mx).astype(np.int32)).to(dev)
agg = torch.from_numpy(rng.standard_normal((mx, dim)).astype(np.float32)).to(dev)
w = torch.from_numpy((rng.standard_normal((H, 2 * dim)) * 0.1).astype(np.float32)).to(dev)
w_lo = ops.split_lo(w)
nr = torch.tensor([rows], dtype=torch.int32, device=dev)
out = torch.zeros((mx, H), device=dev); z = torch.ones((mx, H), device=dev)
for wl in (None, w_lo):
    out.zero_()
    ops.sage_gemm_fwd(tab, sidx, agg, dim, w, H, False, nr, mx, relu=True, precision=native.PREC_TF32X3, out=out, zero_out=z, weight_lo=wl)
    torch.cuda.synchronize()
    X = torch.cat([tab[sidx[:rows].long()], agg[:rows]], 1).double()
    want = torch.relu(X @ w.double().t())
    err = float((out[:rows].double() - want).abs().max() / want.abs().max())
    print(f"weight_lo {'given' if wl is not None else 'none '}: rel err {err:.2e}, zero_out cleared {float(z[:rows].abs().max()) == 0.0}, rows beyond live untouched {float(out[rows:].abs().max()) == 0.0}")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        ops.sage_gemm_fwd(tab, sidx, agg, dim, w, H, False, nr, mx, relu=True, precision=native.PREC_TF32X3, out=out, zero_out=z, weight_lo=w_lo)
g.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    g.replay()
b.record(); torch.cuda.synchronize()
print(f"GS_TMA_GATHER={os.environ.get('GS_TMA_GATHER', '1')}: {a.elapsed_time(b) * 10:.2f} us per launch (graph chain of 20, PDL)")
lib = native.load()
if hasattr(lib, 'gs_debug_tma_trace_read'):
    import ctypes
    ops.sage_gemm_fwd(tab, sidx, agg, dim, w, H, False, nr, mx, relu=True, precision=native.PREC_TF32X3, out=out, zero_out=z, weight_lo=w_lo)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 32)()
    lib.gs_debug_tma_trace_read(buf, 32)
    t = list(buf); t0 = t[0]
    print('CTA 0 cycles since kernel start: pdl done %d, setup done %d, acc ready %d, end %d' % (t[1] - t0, t[2] - t0, t[3] - t0, t[5] - t0))
    print('  stage issued :', [t[24 + i] - t0 for i in range(8)])
    print('  stage landed :', [t[16 + i] - t0 for i in range(8)])
    print('  stage mma    :', [t[8 + i] - t0 for i in range(8)])
