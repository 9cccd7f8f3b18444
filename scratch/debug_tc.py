import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphsage_b200
from graphsage_b200 import ops as g, native
native.load()
dev = torch.device('cuda:0')
def rel(a, b):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    return float((a-b).abs().max() / b.abs().max().clamp_min(1e-30))
def run(dim, out_dim, gcn, rows, live, precision):
    rng = np.random.default_rng(1)
    n_table = 3000
    ld = (dim + 3) & ~3
    table = torch.zeros((n_table, ld)); table[:, :dim] = torch.from_numpy(rng.standard_normal((n_table, dim)).astype(np.float32))
    agg = torch.zeros((rows, ld)); agg[:, :dim] = torch.from_numpy(rng.standard_normal((rows, dim)).astype(np.float32))
    self_idx = rng.integers(0, n_table, size=rows)
    w = torch.from_numpy(rng.uniform(-0.2, 0.2, size=(out_dim, dim if gcn else 2 * dim)).astype(np.float32))
    live_t = torch.tensor([live], dtype=torch.int32, device=dev)
    sidx_d = torch.from_numpy(self_idx.astype(np.int32)).to(dev)
    t_d, a_d, w_d = table.to(dev), agg.to(dev), w.to(dev)
    out32 = g.sage_gemm_fwd(None if gcn else t_d, sidx_d, a_d, dim, w_d, out_dim, gcn, live_t, rows, True, 0)
    out = g.sage_gemm_fwd(None if gcn else t_d, sidx_d, a_d, dim, w_d, out_dim, gcn, live_t, rows, True, precision)
    e_f = rel(out[:live, :out_dim], out32[:live, :out_dim])
    gout = torch.from_numpy(rng.standard_normal((rows, (out_dim + 3) & ~3)).astype(np.float32)).to(dev)
    gw32 = torch.zeros_like(w_d); g.sage_gemm_bwd_w(None if gcn else t_d, sidx_d, a_d, dim, gout, out32, out_dim, gcn, True, live_t, rows, gw32)
    gw = torch.zeros_like(w_d); g.sage_gemm_bwd_w(None if gcn else t_d, sidx_d, a_d, dim, gout, out32, out_dim, gcn, True, live_t, rows, gw, precision=precision)
    e_w = rel(gw, gw32)
    gs32, ga32 = g.sage_gemm_bwd_x(gout, out32, w_d, dim, out_dim, gcn, True, live_t, rows)
    gs, ga = g.sage_gemm_bwd_x(gout, out32, w_d, dim, out_dim, gcn, True, live_t, rows, precision=precision)
    e_a = rel(ga[:live, :dim], ga32[:live, :dim])
    e_s = rel(gs[:live, :dim], gs32[:live, :dim]) if not gcn else 0.0
    torch.cuda.synchronize()
    print(f"dim={dim} out={out_dim} gcn={gcn} rows={rows} live={live} prec={precision}: fwd {e_f:.2e} bwd_w {e_w:.2e} bwd_x agg {e_a:.2e} self {e_s:.2e}  |gw| {float(gw.abs().max()):.3e} |gw32| {float(gw32.abs().max()):.3e}", flush=True)
    return gw, gw32
for prec in (2, 1):
    for cfg in [(128, 128, False, 1024, 1024), (100, 128, False, 11264, 10900), (64, 32, False, 200, 130), (602, 128, True, 1000, 999), (50, 256, False, 400, 400), (1433, 128, False, 300, 257), (7, 5, False, 64, 33)]:
        try:
            gw, gw32 = run(*cfg, prec)
        except Exception as e:
            print('FAILED', cfg, prec, repr(e)[:300], flush=True)
gw, gw32 = run(128, 128, False, 128, 128, 1)
print(gw[:4, :8].cpu(), gw32[:4, :8].cpu())
