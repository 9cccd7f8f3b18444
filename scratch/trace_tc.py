import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import graphsage_b200
from graphsage_b200 import build as b
b.LIB = os.path.join(ROOT, 'scratch', 'libgsage_trace.so')      # load the trace build instead
b.is_current = lambda: True
from graphsage_b200 import native, ops as g
lib = native.load(build_if_missing=False)
lib.gs_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
dev = torch.device('cuda:0')
def trace(label, fn, k_stages):
    for _ in range(3):
        fn(); torch.cuda.synchronize()
    buf = np.zeros(512, dtype=np.int64)
    lib.gs_debug_trace_read(buf.ctypes.data, 512)
    t0 = buf[0]
    rel = lambda i: int(buf[i] - t0)
    print(f"== {label}: setup_done {rel(1)}  acc_ready {rel(2)}  tmem2smem {rel(4)}  synced {rel(5)}  epi_done {rel(3)}")
    for ks in range(k_stages):
        print(f"   ks{ks}: loader_issued {rel(19+4*ks):7d}  landed {rel(16+4*ks):7d}  converted {rel(17+4*ks):7d}  mma_start {rel(18+4*ks):7d}")
rng = np.random.default_rng(0)
def mk(rows, dim, out_dim):
    n_table = 3000
    table = torch.from_numpy(rng.standard_normal((n_table, dim)).astype(np.float32)).to(dev)
    agg = torch.from_numpy(rng.standard_normal((rows, dim)).astype(np.float32)).to(dev)
    sidx = torch.from_numpy(rng.integers(0, n_table, size=rows).astype(np.int32)).to(dev)
    w = torch.from_numpy(rng.uniform(-0.2, 0.2, size=(out_dim, 2 * dim)).astype(np.float32)).to(dev)
    gout = torch.from_numpy(rng.standard_normal((rows, out_dim)).astype(np.float32)).to(dev)
    return table, agg, sidx, w, gout
for prec in (2, 1):
    table, agg, sidx, w, gout = mk(11264, 100, 128)
    out = g.sage_gemm_fwd(table, sidx, agg, 100, w, 128, False, None, 11264, True, 0)
    trace(f"fwd L1 prec={prec}", lambda: g.sage_gemm_fwd(table, sidx, agg, 100, w, 128, False, None, 11264, True, prec), 7)
    gw = torch.zeros_like(w)
    trace(f"bwd_w L1 prec={prec}", lambda: g.sage_gemm_bwd_w(table, sidx, agg, 100, gout, out, 128, False, False, None, 11264, gw, precision=prec), 5)
    table, agg, sidx, w, gout = mk(1024, 128, 128)
    out = g.sage_gemm_fwd(table, sidx, agg, 128, w, 128, False, None, 1024, True, 0)
    trace(f"fwd L2 prec={prec}", lambda: g.sage_gemm_fwd(table, sidx, agg, 128, w, 128, False, None, 1024, True, prec), 8)
    trace(f"bwd_x L2 prec={prec}", lambda: g.sage_gemm_bwd_x(gout, out, w, 128, 128, False, False, None, 1024, precision=prec), 4)
    gw = torch.zeros_like(w)
    trace(f"bwd_w L2 prec={prec}", lambda: g.sage_gemm_bwd_w(table, sidx, agg, 128, gout, out, 128, False, False, None, 1024, gw, precision=prec), 4)
