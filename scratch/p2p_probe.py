import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import subprocess
print(subprocess.run("nvidia-smi topo -m | head -8; nvidia-smi nvlink -s -i 0 | head -8", shell=True, capture_output=True, text=True).stdout, flush=True)
import numpy as np, torch
import graphsage_b200
from graphsage_b200 import native, ops, peer
native.load()
d0, d1 = torch.device('cuda:0'), torch.device('cuda:1')
print("can_access_peer", torch.cuda.can_device_access_peer(0, 1), flush=True)
a = torch.empty((1 << 28,), dtype=torch.float32, device=d0)   # 1 GiB
b = torch.empty((1 << 28,), dtype=torch.float32, device=d1)
for _ in range(2):
    b.copy_(a)
torch.cuda.synchronize(d0); torch.cuda.synchronize(d1)
t0 = time.time()
for _ in range(5):
    b.copy_(a)
torch.cuda.synchronize(d0); torch.cuda.synchronize(d1)
print("memcpy peer GB/s", 5 * a.numel() * 4 / (time.time() - t0) / 1e9, flush=True)
# gather kernel on cuda:0 reading a shard that lives on cuda:1 (same process, peer access enabled by torch)
torch.cuda.set_device(0)
rps, dim = 4_000_000, 128
s0 = torch.randn((rps, dim), device=d0).to(torch.bfloat16)
s1 = torch.randn((rps, dim), device=d1).to(torch.bfloat16)
rows, stride = 90000, 10
for name, bases, lo, hi in [("local", [s0.data_ptr(), s0.data_ptr()], 0, rps), ("remote", [s0.data_ptr(), s1.data_ptr()], rps, 2 * rps),
                            ("half", [s0.data_ptr(), s1.data_ptr()], 0, 2 * rps)]:
    table = peer.ShardedTable(bases, dim, rps, 2 * rps, dim, d0)
    nbr = torch.randint(lo, hi, (rows, stride), device=d0, dtype=torch.int32)
    cnt = torch.full((rows,), stride, dtype=torch.int32, device=d0)
    nodes = torch.randint(lo, hi, (rows,), device=d0, dtype=torch.int32)
    for _ in range(2):
        ops.agg_fwd_sharded(table, nbr, stride, cnt, nodes, None, rows)
    torch.cuda.synchronize(d0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.agg_fwd_sharded(table, nbr, stride, cnt, nodes, None, rows)
    e1.record(); torch.cuda.synchronize(d0)
    t = e0.elapsed_time(e1) / 5 * 1e-3
    print(name, "us", t * 1e6, "gathered GB/s", rows * (stride + 1) * dim * 2 / t / 1e9, flush=True)
