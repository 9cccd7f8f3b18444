"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path --
seed sharding, the flat gradient layout, the single all-reduce and the divide-by-world update
rule -- checked against a single-process evaluation of the global batch with the oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from oracle import sage_oracle as so


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


class FirstK:
    """Deterministic stand-in for `random` at the sampling seam: a node's draw depends on the node
    only, so a frontier node reached by both ranks gets the same neighbours as in one global batch
    (with a real RNG the two evaluations agree only in expectation)."""

    @staticmethod
    def sample(population, k):
        return sorted(population)[:k]


def _oracle_grads(inp, batch):
    w = [torch.from_numpy(x.copy()).requires_grad_(True) for x in inp['weights']]
    cw = torch.from_numpy(inp['cls_w'].copy()).requires_grad_(True)
    cb = torch.from_numpy(inp['cls_b'].copy()).requires_grad_(True)
    adj = so.LazySetAdjacency(inp['rowptr'], inp['col'])
    embs = so.graphsage_forward(w, torch.from_numpy(inp['feats']), adj, batch, inp['spec']['gcn'], inp['spec']['agg'],
                                rng=FirstK)
    loss = so.supervised_loss(so.classification(cw, cb, embs), inp['labels'][np.asarray(batch)])
    loss.backward()
    return [w[0].grad, w[1].grad, cw.grad, cb.grad]


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import graphsage_b200  # noqa: F401
    from graphsage_b200.trainer import dp_allreduce_, flat_layout, shard_batches
    torch.set_num_threads(1)
    inp = cases.build_inputs('pubmed_selfloop_mean')
    b_sz = 8
    shard = shard_batches(inp['train'], b_sz, steps=3, rank=rank, world=world, seed=824)
    grads = _oracle_grads(inp, shard[1])
    offs, total = flat_layout([tuple(g.shape) for g in grads])
    flat = torch.zeros(total)
    for o, g in zip(offs, grads):
        flat[o:o + g.numel()] = g.reshape(-1)
    dp_allreduce_(flat, world)                      # the one collective of the step
    flat /= world                                   # what gs_clip_sgd's grad_div does
    out_q.put((rank, shard, flat.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_equals_global_batch_gradient():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, shard, flat = q.get(timeout=240)
        got[rank] = (shard, flat)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards: disjoint per step, together they tile the global batch stream
    s0, s1 = got[0][0], got[1][0]
    assert s0.shape == s1.shape == (3, 8)
    for step in range(3):
        assert not set(s0[step]) & set(s1[step])
    assert len(set(s0.ravel()) | set(s1.ravel())) == 48
    # both ranks hold the same reduced gradient
    assert np.array_equal(got[0][1], got[1][1])
    # ... and it is the gradient of the mean loss over the union batch (equal shard sizes)
    inp = cases.build_inputs('pubmed_selfloop_mean')
    union = np.concatenate([s0[1], s1[1]])
    import graphsage_b200  # noqa: F401
    from graphsage_b200.trainer import flat_layout
    grads = _oracle_grads(inp, union)
    offs, total = flat_layout([tuple(g.shape) for g in grads])
    want = np.zeros(total, dtype=np.float32)
    for o, g in zip(offs, grads):
        want[o:o + g.numel()] = g.reshape(-1).numpy()
    err = np.abs(got[0][1] - want).max() / np.abs(want).max()
    assert err <= 1e-5, err


def test_flat_layout_is_16_byte_aligned():
    import graphsage_b200  # noqa: F401
    from graphsage_b200.trainer import flat_layout
    offs, total = flat_layout([(128, 200), (128, 256), (47, 128), (47,)])
    assert offs == [0, 25600, 58368, 64384] and total == 64432
    assert all(o % 4 == 0 for o in offs)
