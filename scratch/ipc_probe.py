import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
import graphsage_b200
from graphsage_b200 import native, ops, peer
native.load()
rps, dim = int(os.environ.get("RPS", "4000000")), 128
rows, stride = 90000, 10
other = 1 - rank

def bench(name, table, who):
    lo, hi = other * rps, (other + 1) * rps
    nbr = torch.randint(lo, hi, (rows, stride), device=dev, dtype=torch.int32)
    cnt = torch.full((rows,), stride, dtype=torch.int32, device=dev)
    nodes = torch.randint(lo, hi, (rows,), device=dev, dtype=torch.int32)
    dist.barrier(); torch.cuda.synchronize()
    if rank in who:
        for _ in range(2):
            ops.agg_fwd_sharded(table, nbr, stride, cnt, nodes, None, rows)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.agg_fwd_sharded(table, nbr, stride, cnt, nodes, None, rows)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5 * 1e-3
        print(f"[rank {rank}] {name} ranks={who}: {t * 1e6:.1f} us, remote gather {rows * (stride + 1) * dim * 2 / t / 1e9:.1f} GB/s", flush=True)
    dist.barrier()

# (a) cudaMalloc + CUDA IPC
shard = torch.randn((rps, dim), device=dev).to(torch.bfloat16)
tab_a = peer.ShardedTable.distributed(shard, rps * world)
bench("cudaMalloc+IPC", tab_a, [0])
bench("cudaMalloc+IPC", tab_a, [0, 1])
# (b) torch symmetric memory (cuMemCreate + fabric/fd handles + cuMemMap)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty((rps * dim,), dtype=torch.int16, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    t.view(torch.bfloat16).copy_(shard.view(-1))
    torch.cuda.synchronize(); dist.barrier()
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    print(f"[rank {rank}] symm ptrs {[hex(p) for p in ptrs]} local {hex(t.data_ptr())}", flush=True)
    tab_b = peer.ShardedTable(ptrs, dim, rps, rps * world, dim, dev)
    bench("symmetric_memory", tab_b, [0])
    bench("symmetric_memory", tab_b, [0, 1])
except Exception as e:
    print(f"[rank {rank}] symmetric memory failed: {e!r}", flush=True)
dist.barrier()
dist.destroy_process_group()
