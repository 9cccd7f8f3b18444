"""GPU: the multi-GPU pieces of the path exercised on ONE device, plus a
2-rank run over real CUDA-IPC peer memory when the box has two GPUs.

  * gs_dp_allreduce_clip_sgd at world 1 against torch's clip_grad_norm_ + SGD (src/utils.py:185-187);
  * the same kernel as rank 0 of a world of 2, with the peer emulated by pre-filling this rank's
    receive slot (so nothing waits), checking the sum order, the pushed copy, the slot being left
    empty again and the rewrite of a gradient word that carries the empty pattern;
  * gs_agg_fwd_bf16_sharded against the dense fp32 kernel on the bf16-rounded table;
  * GraphSage over a ShardedTable against GraphSage over the equivalent dense table.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _empty_region(lib, n_total, world, dev):
    """An exchange region as peer.DpExchange prepares it: zeroed header, receive slots filled with the EMPTY word."""
    nbytes, recv_off = int(lib.gs_dp_region_bytes(n_total, world)), int(lib.gs_dp_region_recv_offset())
    reg = torch.zeros((nbytes,), dtype=torch.uint8, device=dev)
    reg[recv_off:].fill_(0xFF)
    return reg, recv_off


@pytest.fixture(scope='module')
def P():
    import graphsage_b200  # noqa: F401
    from graphsage_b200 import native, peer
    native.load()
    return peer


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


def _flat_problem(dev, shapes, seed, scale=1.0):
    from graphsage_b200.trainer import flat_layout
    g = torch.Generator(device='cpu').manual_seed(seed)
    params = [torch.randn(s, generator=g).to(dev) for s in shapes]
    offs, total = flat_layout([tuple(s) for s in shapes])
    flat = torch.zeros((total,), dtype=torch.float32, device=dev)
    grads = [flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, params)]
    for gr in grads:
        gr.copy_((torch.randn(gr.shape, generator=g) * scale).to(dev))
    return params, offs, flat, grads


def _torch_update(params, grads, groups, max_norm, lr):
    """src/utils.py:185-187 with torch ops: clip each model separately, then SGD."""
    ps = [torch.nn.Parameter(p.detach().cpu().clone()) for p in params]
    for p, g in zip(ps, grads):
        p.grad = g.detach().cpu().clone()
    norms = []
    for grp in sorted(set(groups)):
        norms.append(float(torch.nn.utils.clip_grad_norm_([p for p, q in zip(ps, groups) if q == grp], max_norm)))
    torch.optim.SGD(ps, lr=lr).step()
    return [p.detach() for p in ps], norms


@pytest.mark.parametrize('scale', [0.01, 3.0])          # below and above the clip threshold
def test_dp_update_world1_matches_torch_clip_sgd(P, dev, scale):
    shapes = [(128, 200), (128, 256), (47, 128), (47,)]
    groups = [0, 0, 1, 1]
    params, offs, flat, grads = _flat_problem(dev, shapes, 5, scale)
    want, norms = _torch_update(params, grads, groups, 5.0, 0.7)
    dp = P.DpExchange(flat, params, offs, groups, world=1, rank=0)
    dp.update(5.0, 0.7)
    epoch, status, got_norms = dp.status()
    assert (epoch, status) == (1, 0)
    for p, w in zip(params, want):
        assert rel(p, w) <= 1e-6
    assert abs(got_norms[0] - norms[0]) <= 1e-5 * norms[0] and abs(got_norms[1] - norms[1]) <= 1e-5 * norms[1]
    assert float(flat.abs().max()) == 0.0                              # zero_grad, src/utils.py:189-191
    # a second step on fresh gradients reuses the state (epoch 2)
    for gr in grads:
        gr.normal_()
    want2, _ = _torch_update(params, grads, groups, 5.0, 0.7)
    dp.update(5.0, 0.7)
    assert dp.status()[0] == 2
    for p, w in zip(params, want2):
        assert rel(p, w) <= 1e-6


@pytest.mark.parametrize('world', [1, 2])
def test_dp_update_folds_gradient_replicas_and_keeps_weight_low_halves(P, dev, world):
    """seg_extra (replicas a producer spread its atomics over) are summed into the gradient before the norm / the
    exchange and cleared; seg_params_lo receives p - trunc_tf32(p) of the updated parameter."""
    from graphsage_b200 import native, ops
    lib = native.load()
    shapes = [(128, 200), (47, 128), (47,)]
    groups = [0, 1, 1]
    params, offs, flat, grads = _flat_problem(dev, shapes, 21, 1.0)
    gen = torch.Generator(device='cpu').manual_seed(4)
    ex_w = torch.randn((7, 47 * 128), generator=gen).to(dev)
    ex_b = torch.zeros((7, 64), device=dev)
    ex_b[:, :47] = torch.randn((7, 47), generator=gen).to(dev)
    w_lo = ops.split_lo(params[0])
    total = [grads[0].clone(), grads[1] + ex_w.sum(0).view(47, 128), grads[2] + ex_b[:, :47].sum(0)]
    kw = {}
    if world == 2:      # the emulated peer contributes zeros: the mean halves the local sum
        mine, recv_off = _empty_region(lib, flat.numel(), 2, dev)
        theirs, _ = _empty_region(lib, flat.numel(), 2, dev)
        mine[recv_off:].view(torch.float32).view(2, 2, flat.numel())[1, 1].zero_()      # epoch 1 -> parity 1, source rank 1
        kw = dict(region_ptrs=[mine.data_ptr(), theirs.data_ptr()])
        total = [t * 0.5 for t in total]
    want, _ = _torch_update(params, total, groups, 5.0, 0.7)
    dp = P.DpExchange(flat, params, offs, groups, world=world, rank=0, params_lo=[w_lo, None, None],
                      extras=[None, ex_w, ex_b], **kw)
    dp.update(5.0, 0.7)
    assert dp.status()[:2] == (1, 0)
    for p, w in zip(params, want):
        assert rel(p, w) <= 1e-6
    assert float(ex_w.abs().max()) == 0.0 and float(ex_b.abs().max()) == 0.0 and float(flat.abs().max()) == 0.0
    hi = (params[0].view(torch.int32) & -8192).view(torch.float32)
    assert torch.equal(hi + w_lo, params[0])
    if world == 2:      # what went over the wire already contained the replicas
        their_recv = theirs[recv_off:].view(torch.float32).view(2, 2, flat.numel())
        o = offs[1]
        assert rel(their_recv[1, 0, o:o + 47 * 128], (total[1] * 2).reshape(-1)) <= 1e-6


def test_dp_update_rank0_of_2_with_emulated_peer(P, dev):
    from graphsage_b200 import native
    lib = native.load()
    shapes = [(64, 130), (10, 64), (10,)]            # 130: a tensor whose numel is not a multiple of 4 columns wide
    groups = [0, 1, 1]
    params, offs, flat, grads = _flat_problem(dev, shapes, 11, 2.0)
    n_total = flat.numel()
    mine, recv_off = _empty_region(lib, n_total, 2, dev)
    theirs, _ = _empty_region(lib, n_total, 2, dev)
    dp = P.DpExchange(flat, params, offs, groups, world=2, rank=0, region_ptrs=[mine.data_ptr(), theirs.data_ptr()])
    my_recv = mine[recv_off:].view(torch.float32).view(2, 2, n_total)          # [parity][source rank][n]
    their_recv = theirs[recv_off:].view(torch.float32).view(2, 2, n_total)
    gen = torch.Generator(device='cpu').manual_seed(3)
    for epoch in (1, 2, 3, 4):
        par = epoch & 1
        local = (torch.randn((n_total,), generator=gen) * 2.0).to(dev)
        for gr, o in zip(grads, offs):
            gr.copy_(local[o:o + gr.numel()].view_as(gr))
        local = flat.clone()
        remote = torch.zeros_like(flat)
        for gr, o in zip(grads, offs):
            remote[o:o + gr.numel()] = (torch.randn((gr.numel(),), generator=gen) * 2.0).to(dev)
        assert torch.all(my_recv[par, 1].view(torch.int32) == -1)               # empty before the peer pushes
        my_recv[par, 1].copy_(remote)              # what rank 1 would have pushed
        mean = [((local[o:o + p.numel()] + remote[o:o + p.numel()]) * 0.5).view_as(p) for o, p in zip(offs, params)]
        want, _ = _torch_update(params, mean, groups, 5.0, 0.7)
        torch.cuda.synchronize()
        dp.update(5.0, 0.7)
        assert dp.status()[:2] == (epoch, 0)
        for p, w in zip(params, want):
            assert rel(p, w) <= 1e-6
        assert torch.equal(their_recv[par, 0], local)                           # my gradient landed in the peer's slot
        assert torch.all(my_recv[par, 1].view(torch.int32) == -1)               # what I consumed is empty again
        assert torch.all(my_recv[par ^ 1].view(torch.int32) == -1) and torch.all(my_recv[par, 0].view(torch.int32) == -1)
        assert float(flat.abs().max()) == 0.0
        their_recv[par, 0].view(torch.int32).fill_(-1)                          # the peer's own kernel would have done this


def test_dp_update_never_pushes_the_empty_pattern(P, dev):
    """A gradient word with the bits 0xffffffff (a NaN) must not look like 'nothing arrived' to the peer: it goes out
    as the canonical NaN."""
    from graphsage_b200 import native
    lib = native.load()
    shapes = [(16, 16)]
    params, offs, flat, grads = _flat_problem(dev, shapes, 5)
    mine, recv_off = _empty_region(lib, flat.numel(), 2, dev)
    theirs, _ = _empty_region(lib, flat.numel(), 2, dev)
    mine[recv_off:].view(torch.float32).view(2, 2, flat.numel())[1, 1].zero_()
    flat.view(torch.int32)[7] = -1
    dp = P.DpExchange(flat, params, offs, [0], world=2, rank=0, region_ptrs=[mine.data_ptr(), theirs.data_ptr()])
    dp.update(5.0, 0.7)
    assert dp.status()[:2] == (1, 0)
    sent = theirs[recv_off:].view(torch.int32).view(2, 2, flat.numel())[1, 0]
    assert int(sent[7]) == 0x7fffffff and not bool((sent == -1).any())


def test_dp_update_times_out_instead_of_hanging(P, dev):
    shapes = [(32, 32)]
    params, offs, flat, grads = _flat_problem(dev, shapes, 1)
    from graphsage_b200 import native
    lib = native.load()
    mine, _ = _empty_region(lib, flat.numel(), 2, dev)
    theirs, _ = _empty_region(lib, flat.numel(), 2, dev)
    dp = P.DpExchange(flat, params, offs, [0], world=2, rank=0, timeout_s=0.05,
                      region_ptrs=[mine.data_ptr(), theirs.data_ptr()])
    dp.update(5.0, 0.7)                                # the peer never publishes
    with pytest.raises(RuntimeError, match='timed out'):
        dp.status()


@pytest.mark.parametrize('dim,shards', [(128, 8), (128, 1), (100, 3), (264, 2)])
def test_agg_fwd_sharded_matches_dense_kernel(P, dev, dim, shards):
    from graphsage_b200 import native, ops
    rng = np.random.default_rng(dim + shards)
    n, rows, stride = 5000, 777, 11
    feats = torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)).to(dev)
    table = P.ShardedTable.from_full(feats, shards)
    dense = table.to_dense_fp32()
    assert torch.equal(dense, feats.to(torch.bfloat16).float())
    cnt = rng.integers(0, stride + 1, size=rows).astype(np.int32)
    cnt[:3] = [0, 1, stride]
    nbr = np.full((rows, stride), -1, dtype=np.int32)
    for r in range(rows):
        nbr[r, :cnt[r]] = np.sort(rng.choice(n, size=cnt[r], replace=False))
    nodes = rng.integers(0, n, size=rows).astype(np.int32)
    nodes[0], nodes[1] = 0, n - 1
    nbr_d, cnt_d, nodes_d = (torch.from_numpy(x).to(dev) for x in (nbr, cnt, nodes))
    live = torch.tensor([rows - 5], dtype=torch.int32, device=dev)
    agg, selfr = ops.agg_fwd_sharded(table, nbr_d, stride, cnt_d, nodes_d, live, rows)
    pad = torch.nn.functional.pad(dense, (0, (-dim) % 4)).contiguous()
    want, _ = ops.agg_fwd(pad, dim, nbr_d, stride, cnt_d, live, rows, native.AGG_MEAN)
    torch.cuda.synchronize()
    ok = cnt[:rows - 5] > 0
    got, ref = agg[:rows - 5, :dim].cpu().numpy(), want[:rows - 5, :dim].cpu().numpy()
    assert rel(got[ok], ref[ok]) <= 1e-6
    assert np.all(np.isnan(got[~ok]))                                   # 0/0 as in the reference (src/models.py:312-313)
    assert torch.equal(selfr[:rows - 5, :dim], dense[nodes_d[:rows - 5].long()])


@pytest.mark.parametrize('gcn', [False, True])
def test_graphsage_over_sharded_table_matches_dense_table(P, dev, gcn):
    from graphsage_b200 import models
    from graphsage_b200.graph import AdjCSR
    rowptr, col = cases.load_topology('cora')
    n = len(rowptr) - 1
    rng = np.random.default_rng(4)
    feats = torch.from_numpy(rng.standard_normal((n, 128)).astype(np.float32)).to(dev)
    table = P.ShardedTable.from_full(feats, 4)
    adj = AdjCSR(rowptr, col)
    torch.manual_seed(0)
    a = models.GraphSage(2, 128, 64, table, adj, dev, gcn=gcn, agg_func='MEAN', seed=9, precision='fp32').to(dev)
    b = models.GraphSage(2, 128, 64, table.to_dense_fp32(), adj, dev, gcn=gcn, agg_func='MEAN', seed=9, precision='fp32').to(dev)
    b.load_state_dict(a.state_dict())
    batch = np.arange(0, n, 9)
    ea, eb = a(batch), b(batch)                       # same Philox seed and call count => same samples
    assert rel(ea, eb) <= 1e-5
    ea.square().sum().backward()
    eb.square().sum().backward()
    for l in (1, 2):
        ga, gb = getattr(a, f'sage_layer{l}').weight.grad, getattr(b, f'sage_layer{l}').weight.grad
        assert rel(ga, gb) <= 1e-5


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_ranks_over_cuda_ipc():
    """Real peers: tests/mp_peer_check.py under torchrun on 2 GPUs (fused all-reduce/update against
    NCCL + torch, bit-identical replicas, remote-shard gather)."""
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', '29517', os.path.join(ROOT, 'tests', 'mp_peer_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'MP_PEER_CHECK_OK' in out.stdout


@pytest.mark.parametrize('agg_func', ['MEAN', 'MAX'])
def test_sharded_gather_matches_the_oracle_on_a_device_generated_powerlaw_graph(P, dev, agg_func):
    """The cfg-5 path against the ORACLE (oracle.aggregate, the reference's dense-mask algorithm, src/models.py:
    291-326) -- not against another kernel of this repo: a down-scaled `synth.device_powerlaw_csr` graph (the
    generator bench.py --workload cfg5 uses: rows repeat ids), bf16 table in 4 shards, native sampler draws, the
    drawn lists replayed through the oracle on the bf16-rounded table.  MEAN within 1e-6 (fp32 accumulation of the
    same bf16 values), MAX exactly."""
    from graphsage_b200 import native, ops, synth
    from oracle import sage_oracle as so
    n, dim, rows, k = 60000, 128, 1500, 10
    rowptr, col = synth.device_powerlaw_csr(n, 16.0, dev, seed=3)
    rng = np.random.default_rng(8)
    feats = torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)).to(dev)
    table = P.ShardedTable.from_full(feats, 4)
    rounded = table.to_dense_fp32().cpu()                       # what the kernel really reads
    nodes = torch.from_numpy(rng.choice(n, size=rows, replace=False).astype(np.int32)).to(dev)
    nbr, cnt = ops.sample_neighbors(rowptr, col, n, nodes, None, rows, k, k, native.SELF_DROP, 5, 9)
    mode = native.AGG_MEAN if agg_func == 'MEAN' else native.AGG_MAX
    agg, selfr = ops.agg_fwd_sharded(table, nbr, k, cnt, nodes, None, rows, mode=mode)
    torch.cuda.synchronize()
    nbr_h, cnt_h, nodes_h = nbr.cpu().numpy(), cnt.cpu().numpy(), nodes.cpu().tolist()
    keep = [i for i in range(rows) if cnt_h[i] > 0]              # the reference raises on an empty MAX row, NaN for MEAN
    assert len(keep) > rows * 0.9
    samp = [set(nbr_h[i, :cnt_h[i]].tolist()) | {nodes_h[i]} for i in keep]
    uniq = sorted(set().union(*samp))
    index_of = {v: j for j, v in enumerate(uniq)}
    want = so.aggregate([nodes_h[i] for i in keep], rounded, (uniq, samp, index_of), False, agg_func)
    got = agg[torch.as_tensor(keep, device=dev), :dim].cpu()
    if agg_func == 'MAX':
        assert torch.equal(got, want)
    else:
        assert rel(got, want) <= 1e-6
    assert torch.equal(selfr[:, :dim].cpu(), rounded[torch.as_tensor(nodes_h)])
    # repeated ids inside CSR rows (the generator draws targets with replacement) really occurred and were drawn as sets
    rp, cl = rowptr.cpu().numpy(), col.cpu().numpy()
    assert any(len(set(cl[rp[v]:rp[v + 1]].tolist())) < rp[v + 1] - rp[v] for v in nodes_h[:400])
    assert all(len(set(nbr_h[i, :cnt_h[i]].tolist())) == cnt_h[i] for i in range(rows))


def test_graphsage_max_over_a_sharded_table(P, dev):
    """agg_func='MAX' over a ShardedTable (src/models.py:316-326): forward and weight gradients equal the same model
    over the dense copy of the bf16 table."""
    from graphsage_b200 import models
    from graphsage_b200.graph import AdjCSR
    rowptr, col = cases.load_topology('cora')
    n = len(rowptr) - 1
    rng = np.random.default_rng(4)
    feats = torch.from_numpy(rng.standard_normal((n, 128)).astype(np.float32)).to(dev)
    table = P.ShardedTable.from_full(feats, 4)
    adj = AdjCSR(rowptr, col)
    torch.manual_seed(0)
    a = models.GraphSage(2, 128, 64, table, adj, dev, gcn=False, agg_func='MAX', seed=9, precision='fp32').to(dev)
    b = models.GraphSage(2, 128, 64, table.to_dense_fp32(), adj, dev, gcn=False, agg_func='MAX', seed=9, precision='fp32').to(dev)
    b.load_state_dict(a.state_dict())
    batch = np.arange(0, n, 9)
    ea, eb = a(batch), b(batch)
    assert rel(ea, eb) <= 1e-5
    ea.square().sum().backward()
    eb.square().sum().backward()
    for l in (1, 2):
        assert rel(getattr(a, f'sage_layer{l}').weight.grad, getattr(b, f'sage_layer{l}').weight.grad) <= 1e-5
