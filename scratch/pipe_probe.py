import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import graphsage_b200
from graphsage_b200 import models, native, synth
from graphsage_b200.graph import AdjCSR
from graphsage_b200.trainer import PipelinedTrainer
native.load()
dev = torch.device('cuda:0')
cfg = synth.CONFIGS["cfg3_products"]
rowptr, col = synth.powerlaw_graph(cfg["n"], cfg["edges"], seed=0, cache_dir="/tmp/gsage_cache")
feats = torch.from_numpy(synth.features_normal(cfg["n"], 100, seed=1)).to(dev)
labels = synth.labels_uniform(cfg["n"], 47, seed=2)
model = models.GraphSage(2, 100, 128, feats, AdjCSR(rowptr, col), dev, seed=1).to(dev)
cls = models.Classification(128, 47).to(dev)
tr = PipelinedTrainer(model, cls, labels, 1024)
seeds = torch.randint(0, cfg["n"], (64, 1024), device=dev, dtype=torch.int32)
tr.submit_device(seeds[0]); tr.submit_device(seeds[1]); torch.cuda.synchronize()
def cap(fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); tr.flat_grad.zero_()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g
gP = cap(lambda: tr._prep(1))
gC = cap(lambda: tr._compute(0))
gPC = cap(lambda: tr._both(0))
def timeit(g, n=300):
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for name, g in (("prep", gP), ("compute", gC), ("both", gPC)):
    print(name, "us/replay", round(timeit(g), 1), flush=True)
tr.dp.status()
# ---- where do the 15 us between the probe (106) and the bench loop (121) go? ----
g0, g1 = tr._graphs[0], tr._graphs[1]
def loop(fn, n=400):
    for _ in range(10): fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
print("same graph, no copy     ", round(loop(lambda i: g0.replay()), 1))
print("alternating, no copy    ", round(loop(lambda i: (g0 if i & 1 else g1).replay()), 1))
def with_copy(i):
    tr.slot_seeds[i & 1].copy_(seeds[i % 64], non_blocking=True)
    (g0 if i & 1 else g1).replay()
print("alternating + seeds copy", round(loop(with_copy), 1))
print("submit_device           ", round(loop(lambda i: tr.submit_device(seeds[i % 64])), 1))
tr.dp.status()
