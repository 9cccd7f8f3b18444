"""Event-timed chains (CUDA graph of 20 launches) of the small kernels of the training chain at cfg-3 sizes."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import graphsage_b200
from graphsage_b200 import native, ops as g
native.load()
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
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
    print(f"  {label:34s} {a.elapsed_time(c) * 1e3 / (5 * n):7.2f} us", flush=True)
rows, dim, C = 1024, 128, 47
emb = torch.from_numpy(rng.standard_normal((rows, dim)).astype(np.float32)).clamp_(min=0).to(dev)
w = torch.from_numpy(rng.uniform(-0.2, 0.2, size=(C, dim)).astype(np.float32)).to(dev)
b = torch.zeros(C, device=dev)
labels = torch.from_numpy(rng.integers(0, C, size=5000)).to(dev)
idx = torch.from_numpy(rng.integers(0, 5000, size=rows).astype(np.int32)).to(dev)
loss = torch.zeros(1, device=dev); gemb = torch.empty_like(emb); gw = torch.zeros_like(w); gb = torch.zeros_like(b)
logp = torch.empty((rows, C), device=dev); scratch = torch.empty((rows, C), device=dev)
timeit("cls full (grad_w, grad_b)", lambda: g.cls_nll_fwd_bwd(emb, dim, w, b, C, labels, idx, loss, gemb, gw, gb, logp=logp, scratch=scratch, mask_relu_input=True, zero_loss=False))
timeit("cls without grad_w", lambda: g.cls_nll_fwd_bwd(emb, dim, w, b, C, labels, idx, loss, gemb, None, gb, logp=logp, scratch=scratch, mask_relu_input=True, zero_loss=False))
timeit("cls without grad_w, grad_b", lambda: g.cls_nll_fwd_bwd(emb, dim, w, b, C, labels, idx, loss, gemb, None, None, logp=logp, scratch=scratch, mask_relu_input=True, zero_loss=False))
timeit("cls without grad_w, grad_emb", lambda: g.cls_nll_fwd_bwd(emb, dim, w, b, C, labels, idx, loss, None, None, gb, logp=logp, scratch=scratch, mask_relu_input=True, zero_loss=False))
# layer-2 aggregation and its backward
R1, R2, k = 11264, 1024, 10
h1 = torch.from_numpy(rng.standard_normal((R1, dim)).astype(np.float32)).to(dev)
nbr = torch.from_numpy(np.sort(rng.integers(0, R1, size=(R2, k)), axis=1).astype(np.int32)).to(dev)
cnt = torch.full((R2,), k, dtype=torch.int32, device=dev)
sidx = torch.from_numpy(rng.integers(0, R1, size=R2).astype(np.int32)).to(dev)
out = torch.empty((R2, dim), device=dev)
timeit("agg_fwd L2", lambda: g.agg_fwd(h1, dim, nbr, k, cnt, None, R2, native.AGG_MEAN, out=out))
ga = torch.from_numpy(rng.standard_normal((R2, dim)).astype(np.float32)).to(dev)
gs_ = torch.from_numpy(rng.standard_normal((R2, dim)).astype(np.float32)).to(dev)
gt = torch.zeros((R1, dim), device=dev)
timeit("agg_bwd L2 (masked)", lambda: g.agg_bwd(ga, gs_, dim, nbr, k, cnt, sidx, None, None, R2, native.AGG_MEAN, gt, mask_table=h1))
timeit("agg_bwd L2 (no mask)", lambda: g.agg_bwd(ga, gs_, dim, nbr, k, cnt, sidx, None, None, R2, native.AGG_MEAN, gt))
timeit("zeros 11264x128", lambda: gt.zero_())
