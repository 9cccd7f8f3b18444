"""Run under torchrun with >= 2 ranks, one GPU each (see tests/test_gpu_peer.py and DESIGN.md §5):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 tests/mp_peer_check.py

Checks, over real CUDA-IPC peer memory:
  1. gs_dp_allreduce_clip_sgd == NCCL all-reduce(mean) + torch clip_grad_norm_ per model + SGD,
     for several consecutive steps (both slot parities), and the replicas stay BIT-identical;
  2. gs_agg_fwd_bf16_sharded over a table whose shards live on different GPUs == the dense
     kernel on a locally rebuilt copy of the whole table;
  3. a data-parallel SupervisedTrainer step (captured graph, fused exchange) keeps replicas identical.
Prints MP_PEER_CHECK_OK from rank 0 on success.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def all_equal_across_ranks(t, world):
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bufs, t.contiguous())
    return all(torch.equal(bufs[0], b) for b in bufs[1:])


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import graphsage_b200  # noqa: F401
    from graphsage_b200 import models, native, ops, peer
    from graphsage_b200.graph import AdjCSR
    from graphsage_b200.trainer import SupervisedTrainer, flat_layout, shard_batches
    import cases
    native.load()

    # ---- 1. fused all-reduce + clip + SGD --------------------------------------------------------
    shapes, groups = [(128, 200), (128, 256), (47, 128), (47,)], [0, 0, 1, 1]
    g0 = torch.Generator().manual_seed(1)                      # identical initial replicas
    params = [torch.randn(s, generator=g0).to(dev) for s in shapes]
    ref = [p.clone() for p in params]
    offs, total = flat_layout(shapes)
    flat = torch.zeros((total,), dtype=torch.float32, device=dev)
    views = [flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, params)]
    backends = ['symm', 'ipc']                                 # VMM symmetric memory (default) and legacy CUDA IPC regions
    os.environ['GS_PEER_BACKEND'] = backends[0]
    dp = peer.DpExchange(flat, params, offs, groups, world=world, rank=rank)
    gr = torch.Generator().manual_seed(100 + rank)             # different gradients per rank
    for step in range(10):
        if step == 5:                                          # second half: the same exchange over gs_peer_* IPC regions
            dp.close()
            os.environ['GS_PEER_BACKEND'] = backends[1]
            dp = peer.DpExchange(flat, params, offs, groups, world=world, rank=rank)

        scale = [0.01, 5.0, 1.0, 0.1, 3.0][step % 5]
        for v in views:
            v.copy_((torch.randn(v.shape, generator=gr) * scale).to(dev))
        mean = flat.clone()
        dist.all_reduce(mean)
        mean /= world
        dp.update(5.0, 0.7)
        epoch, status, _ = dp.status()
        assert (epoch, status) == (step % 5 + 1, 0), (epoch, status)
        # torch statement of src/utils.py:185-187 on the mean gradient
        ps = [torch.nn.Parameter(r.clone()) for r in ref]
        for p, o in zip(ps, offs):
            p.grad = mean[o:o + p.numel()].view_as(p).clone()
        for grp in (0, 1):
            torch.nn.utils.clip_grad_norm_([p for p, q in zip(ps, groups) if q == grp], 5.0)
        torch.optim.SGD(ps, lr=0.7).step()
        ref = [p.detach().clone() for p in ps]
        for p, r in zip(params, ref):
            assert rel(p, r) <= 1e-6, (step, rel(p, r))
            assert all_equal_across_ranks(p, world), f'replicas diverged at step {step}'
        assert float(flat.abs().max()) == 0.0
    os.environ['GS_PEER_BACKEND'] = 'symm'
    dist.barrier()

    # ---- 2. remote-shard gather ------------------------------------------------------------------
    dim, rps = 128, 4096
    n = rps * world

    def shard_rows(r):
        return torch.randn((rps, dim), generator=torch.Generator().manual_seed(500 + r))

    table = peer.ShardedTable.distributed(shard_rows(rank).to(dev), n)
    full = torch.cat([shard_rows(r) for r in range(world)]).to(dev).to(torch.bfloat16).float()
    rng = np.random.default_rng(7 + rank)
    rows, stride = 3000, 11
    cnt = rng.integers(1, stride + 1, size=rows).astype(np.int32)
    nbr = np.full((rows, stride), -1, dtype=np.int32)
    for r in range(rows):
        nbr[r, :cnt[r]] = np.sort(rng.choice(n, size=cnt[r], replace=False))
    nodes = rng.integers(0, n, size=rows).astype(np.int32)
    nbr_d, cnt_d, nodes_d = (torch.from_numpy(x).to(dev) for x in (nbr, cnt, nodes))
    agg, selfr = ops.agg_fwd_sharded(table, nbr_d, stride, cnt_d, nodes_d, None, rows)
    want, _ = ops.agg_fwd(full, dim, nbr_d, stride, cnt_d, None, rows, native.AGG_MEAN)
    torch.cuda.synchronize()
    assert rel(agg[:, :dim], want[:, :dim]) <= 1e-6
    assert torch.equal(selfr[:, :dim], full[nodes_d.long()])
    remote_frac = float((torch.from_numpy(nbr[nbr >= 0]) // rps != rank).float().mean())
    dist.barrier()

    # ---- 3. data-parallel trainer step over the sharded table ------------------------------------
    rowptr, col = cases.load_topology('cora')
    nn_ = len(rowptr) - 1
    rps2 = (nn_ + world - 1) // world
    feats_full = torch.randn((rps2 * world, 128), generator=torch.Generator().manual_seed(9))
    tab2 = peer.ShardedTable.distributed(feats_full[rank * rps2:(rank + 1) * rps2].to(dev), nn_)
    torch.manual_seed(0)
    model = models.GraphSage(2, 128, 64, tab2, AdjCSR(rowptr, col), dev, gcn=False, agg_func='MEAN', seed=50 + rank,
                             precision='fp32').to(dev)
    cls = models.Classification(64, 7).to(dev)
    labels = np.random.default_rng(3).integers(0, 7, size=nn_)
    tr = SupervisedTrainer(model, cls, labels, 64, world_size=world, rank=rank)
    batches = shard_batches(np.arange(nn_), 64, 6, rank, world, seed=2)
    losses = [float(tr.step(b).item()) for b in batches]
    tr.dp.status()
    assert all(np.isfinite(losses)), losses
    for p in list(model.parameters()) + list(cls.parameters()):
        assert all_equal_across_ranks(p.data, world), 'trainer replicas diverged'
    dist.barrier()
    if rank == 0:
        print(f'MP_PEER_CHECK_OK world={world} remote_rows={remote_frac:.2f} losses={losses}', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
