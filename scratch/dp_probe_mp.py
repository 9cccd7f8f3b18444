"""Diagnostics (torchrun, one rank per GPU): the fused exchange + update kernel alone at the cfg-3 parameter shapes, in
lockstep chains of 20 per graph replay; event time per launch (max over ranks) and, with the trace build, the phase
breakdown of CTA 0 of the last launch on every rank."""
import ctypes, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
import graphsage_b200  # noqa
from graphsage_b200 import native
from graphsage_b200.peer import DpExchange
from graphsage_b200.trainer import flat_layout
shapes = [(128, 200), (128, 256), (47, 128), (47,)]
g0 = torch.Generator().manual_seed(1)
params = [torch.randn(s, generator=g0).to(dev) for s in shapes]
offs, total = flat_layout(shapes)
flat = torch.randn((total,), device=dev)
dp = DpExchange(flat, params, offs, [0, 0, 1, 1], world=world, rank=rank)
for _ in range(3):
    dp.update(5.0, 0.7, None)
torch.cuda.synchronize(); dist.barrier()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        dp.update(5.0, 0.7, None)
g.replay(); torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    g.replay()
b.record(); torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) * 1e3 / 100], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"dp_update world {world}, {total} floats: {t.item():.2f} us per launch (lockstep graph chain of 20, max over ranks)", flush=True)
lib = native.load()
if hasattr(lib, 'gs_debug_dp_trace_read'):
    dist.barrier()
    g.replay(); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    lib.gs_debug_dp_trace_read(buf, 16)
    tt = list(buf)[:8]
    names = ['pdl', 'epoch', 'fold+push+wait', 'loads+partials', 'barrier', 'coef', 'sgd']
    for r in range(world):
        dist.barrier()
        if r == rank:
            print(f'rank {rank} phase cycles (CTA 0, last launch of a chain): ' + ', '.join(f"{n} {tt[i + 1] - tt[i]}" for i, n in enumerate(names)) + f"; total {tt[7] - tt[0]}", flush=True)
dist.barrier()
dist.destroy_process_group()
