"""Time the five tcgen05 GEMM launches of a cfg-3 step for a given build of the library.
   GS_LIB=<path to .so> python scratch/time_tc.py     (scratch tool; event-timed chains of 20 launches)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import graphsage_b200
from graphsage_b200 import build as b
if os.environ.get("GS_LIB"):
    b.LIB = os.path.abspath(os.environ["GS_LIB"])
    b.is_current = lambda: True
from graphsage_b200 import native, ops as g
lib = native.load(build_if_missing=False)
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
def mk(rows, dim, out_dim, n_table):
    table = torch.from_numpy(rng.standard_normal((n_table, dim)).astype(np.float32)).to(dev)
    agg = torch.from_numpy(rng.standard_normal((rows, dim)).astype(np.float32)).to(dev)
    sidx = torch.from_numpy(rng.integers(0, n_table, size=rows).astype(np.int32)).to(dev)
    w = torch.from_numpy(rng.uniform(-0.2, 0.2, size=(out_dim, 2 * dim)).astype(np.float32)).to(dev)
    gout = torch.from_numpy(rng.standard_normal((rows, out_dim)).astype(np.float32)).to(dev)
    return table, agg, sidx, w, gout
def timeit(label, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n): fn()
    gr.replay(); torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): gr.replay()
    c.record(); torch.cuda.synchronize()
    us = a.elapsed_time(c) * 1e3 / (5 * n)
    print(f"  {label:12s} {us:7.2f} us", flush=True)
    return us
prec = int(os.environ.get("PREC", "2"))
tot = 0.0
table, agg, sidx, w, gout = mk(11264, 100, 128, 2_000_000)
out = g.sage_gemm_fwd(table, sidx, agg, 100, w, 128, False, None, 11264, True, 0)
ref = out.clone()
o2 = g.sage_gemm_fwd(table, sidx, agg, 100, w, 128, False, None, 11264, True, prec)
print(f"lib={os.environ.get('GS_LIB','default')} prec={prec}  fwd L1 max rel err vs fp32 path: {float((o2-ref).abs().max()/ref.abs().max()):.2e}")
tot += timeit("fwd L1", lambda: g.sage_gemm_fwd(table, sidx, agg, 100, w, 128, False, None, 11264, True, prec))
gw = torch.zeros_like(w)
tot += timeit("bwd_w L1", lambda: g.sage_gemm_bwd_w(table, sidx, agg, 100, gout, out, 128, False, False, None, 11264, gw, precision=prec))
gw0 = torch.zeros_like(w); g.sage_gemm_bwd_w(table, sidx, agg, 100, gout, out, 128, False, False, None, 11264, gw0, precision=0)
gw1 = torch.zeros_like(w); g.sage_gemm_bwd_w(table, sidx, agg, 100, gout, out, 128, False, False, None, 11264, gw1, precision=prec)
print(f"  bwd_w L1 max rel err: {float((gw1-gw0).abs().max()/gw0.abs().max()):.2e}")
table, agg, sidx, w, gout = mk(1024, 128, 128, 11264)
out = g.sage_gemm_fwd(table, sidx, agg, 128, w, 128, False, None, 1024, True, 0)
tot += timeit("fwd L2", lambda: g.sage_gemm_fwd(table, sidx, agg, 128, w, 128, False, None, 1024, True, prec))
tot += timeit("bwd_x L2", lambda: g.sage_gemm_bwd_x(gout, out, w, 128, 128, False, False, None, 1024, precision=prec))
gx0 = g.sage_gemm_bwd_x(gout, out, w, 128, 128, False, False, None, 1024, precision=0)
gx1 = g.sage_gemm_bwd_x(gout, out, w, 128, 128, False, False, None, 1024, precision=prec)
e = max(float((a1-a0).abs().max()/a0.abs().max()) for a0, a1 in zip(gx0, gx1)) if isinstance(gx0, (tuple, list)) else float((gx1-gx0).abs().max()/gx0.abs().max())
print(f"  bwd_x L2 max rel err: {e:.2e}")
gw = torch.zeros_like(w)
tot += timeit("bwd_w L2", lambda: g.sage_gemm_bwd_w(table, sidx, agg, 128, gout, out, 128, False, False, None, 1024, gw, precision=prec))
print(f"  sum          {tot:7.2f} us")
