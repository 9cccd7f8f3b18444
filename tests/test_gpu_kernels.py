"""GPU: each kernel through the C ABI (ops.py -> libgsage_b200.so) against an independent CPU
statement of the same operation.  Integer outputs are compared bit-exactly; fp32 outputs with
the norm-relative 1e-5 bound of SURVEY.md §8(c)."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import sage_oracle as so

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope='module')
def g():
    import graphsage_b200  # noqa: F401
    from graphsage_b200 import native, ops
    native.load()
    return ops


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


# ------------------------------------------------------------------------------------------------
# K1 sampler
# ------------------------------------------------------------------------------------------------
def _csr_dev(rowptr, col, dev):
    return torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev)


@pytest.mark.parametrize('self_mode', [0, 1, 2])
def test_sampler_exact_semantics(g, dev, self_mode):
    rowptr, col = cases.load_topology('pubmed')          # has 3 self-loops and deg in 1..171
    n = len(rowptr) - 1
    rp, cl = _csr_dev(rowptr, col, dev)
    nodes = np.arange(n, dtype=np.int32)
    k, stride = 10, 11
    nbr, cnt = g.sample_neighbors(rp, cl, n, torch.from_numpy(nodes).to(dev), None, n, k, stride, self_mode, 824, 5)
    nbr, cnt = nbr.cpu().numpy(), cnt.cpu().numpy()
    deg = np.diff(rowptr)
    for v in range(0, n, 3):
        row = nbr[v, :cnt[v]]
        adj = set(col[rowptr[v]:rowptr[v + 1]].tolist())
        assert np.all(nbr[v, cnt[v]:] == -1)
        assert np.all(np.diff(row) > 0)                                        # ascending, distinct
        drawn = set(row.tolist())
        if self_mode == 0:
            assert drawn <= adj and len(drawn) == min(deg[v], k)               # src/models.py:282
        elif self_mode == 1:
            assert v not in drawn and drawn <= adj
            assert len(drawn) >= min(deg[v], k) - 1
            if v not in adj:
                assert len(drawn) == min(deg[v], k)
        else:
            assert v in drawn and (drawn - {v}) <= adj
    # determinism under a fixed (seed, offset) and sensitivity to the offset
    nbr2, _ = g.sample_neighbors(rp, cl, n, torch.from_numpy(nodes).to(dev), None, n, k, stride, self_mode, 824, 5)
    nbr3, _ = g.sample_neighbors(rp, cl, n, torch.from_numpy(nodes).to(dev), None, n, k, stride, self_mode, 824, 6)
    assert np.array_equal(nbr2.cpu().numpy(), nbr)
    assert not np.array_equal(nbr3.cpu().numpy(), nbr)


def test_sampler_distribution(g, dev):
    """Inclusion frequency of every neighbour -> k/deg (chi-square), pair inclusion ->
    k(k-1)/(deg(deg-1)); the native sampler cannot match random.sample's stream."""
    deg, k, trials = 37, 10, 20000
    rowptr = np.array([0, deg], dtype=np.int64)
    col = np.arange(1, deg + 1, dtype=np.int32)
    rp, cl = _csr_dev(rowptr, col, dev)
    nodes = torch.zeros((trials,), dtype=torch.int32, device=dev)
    nbr, cnt = g.sample_neighbors(rp, cl, deg + 1, nodes, None, trials, k, k, 0, 12345, 1)
    nbr = nbr.cpu().numpy()
    assert np.all(cnt.cpu().numpy() == k)
    counts = np.bincount(nbr.ravel(), minlength=deg + 1)[1:]
    expect = trials * k / deg
    chi2 = ((counts - expect) ** 2 / (expect * (1 - k / deg))).sum()
    assert chi2 < 80, chi2                                  # dof 36: mean 36, p(>80) ~ 3e-5
    pair = np.zeros((deg + 1, deg + 1))
    for a in range(k):
        for b in range(a + 1, k):
            np.add.at(pair, (nbr[:, a], nbr[:, b]), 1)
    pair = (pair + pair.T)[1:, 1:][np.triu_indices(deg, 1)]
    p2 = k * (k - 1) / (deg * (deg - 1))
    z = (pair - trials * p2) / np.sqrt(trials * p2 * (1 - p2))
    assert np.abs(z).max() < 5.5 and abs(z.mean()) < 0.3
    # rows are independent: consecutive rows share about k*k/deg members
    inter = np.mean([len(set(nbr[i]) & set(nbr[i + 1])) for i in range(0, 4000, 2)])
    assert abs(inter - k * k / deg) < 0.2


def test_sampler_device_row_count(g, dev):
    rowptr, col = cases.load_topology('cora')
    n = len(rowptr) - 1
    rp, cl = _csr_dev(rowptr, col, dev)
    nodes = torch.arange(100, dtype=torch.int32, device=dev)
    live = torch.tensor([37], dtype=torch.int32, device=dev)
    nbr, cnt = g.sample_neighbors(rp, cl, n, nodes, live, 100, 10, 10, 1, 1, 1)
    assert np.all(nbr[37:].cpu().numpy() == -1) and np.all(cnt[37:].cpu().numpy() == 0)
    assert np.all(cnt[:37].cpu().numpy() > 0)


# ------------------------------------------------------------------------------------------------
# K2 unique / remap: bit-exact
# ------------------------------------------------------------------------------------------------
def _check_unique(g, dev, nodes, nbr, id_bits, live=None):
    rows, stride = nbr.shape
    nodes_t, nbr_t = torch.from_numpy(nodes).to(dev), torch.from_numpy(nbr).to(dev)
    live_t = None if live is None else torch.tensor([live], dtype=torch.int32, device=dev)
    uniq, num, nbr_idx, self_idx = g.unique_remap(nodes_t, live_t, rows, nbr_t, stride, id_bits)
    m = rows if live is None else live
    ids = np.concatenate([nodes[:m], nbr[:m].ravel()])
    want = np.unique(ids[ids >= 0])
    n = int(num.item())
    assert n == len(want)
    got = uniq[:n].cpu().numpy()
    assert np.array_equal(got, want)
    ni, si = nbr_idx.cpu().numpy(), self_idx.cpu().numpy()
    ref_idx = np.where(nbr[:m] >= 0, np.searchsorted(want, np.maximum(nbr[:m], 0)), -1)
    assert np.array_equal(ni[:m], ref_idx)
    assert np.all(ni[m:] == -1)
    assert np.array_equal(si[:m], np.searchsorted(want, nodes[:m]))


@pytest.mark.parametrize('rows,stride,n_ids,bits', [(1, 11, 50, 6), (1024, 11, 2_449_029, 22), (2047, 11, 100_000_000, 27),
                                                    (2048, 11, 5000, 13), (37, 3, 1 << 31, 32)])
def test_unique_single_cta(g, dev, rows, stride, n_ids, bits):
    rng = np.random.default_rng(rows + stride)
    nodes = rng.integers(0, n_ids, size=rows).astype(np.int32)
    nbr = rng.integers(0, n_ids, size=(rows, stride)).astype(np.int32)
    nbr[rng.random((rows, stride)) < 0.2] = -1
    _check_unique(g, dev, nodes, nbr, bits)
    _check_unique(g, dev, nodes, nbr, bits, live=max(1, rows // 3))


@pytest.mark.parametrize('rows,stride,n_ids,bits', [(11264, 11, 2_449_029, 22), (90112, 11, 100_000_000, 27),
                                                    (3000, 106, 19717, 15)])
def test_unique_multi_cta(g, dev, rows, stride, n_ids, bits):
    rng = np.random.default_rng(rows)
    nodes = rng.integers(0, n_ids, size=rows).astype(np.int32)
    nbr = rng.integers(0, n_ids, size=(rows, stride)).astype(np.int32)
    nbr[rng.random((rows, stride)) < 0.1] = -1
    _check_unique(g, dev, nodes, nbr, bits)
    _check_unique(g, dev, nodes, nbr, bits, live=rows // 2 + 1)


@pytest.mark.parametrize('rows,stride,n_ids', [(1, 11, 50), (1024, 11, 2_449_029), (11264, 11, 2_449_029),
                                               (8192, 11, 100_000_000), (3000, 106, 19717), (500, 4, 4096 * 32 + 1)])
def test_unique_bitmap_path(g, dev, rows, stride, n_ids):
    """Bitmap/rank path: identical outputs to the radix path (np.unique oracle), bitmap left clean."""
    rng = np.random.default_rng(rows * 7 + stride)
    ws = g.unique_bitmap_workspace(n_ids, dev)
    words = (n_ids + 31) // 32
    for rep, live in enumerate((None, max(1, rows // 3), None)):
        nodes = rng.integers(0, n_ids, size=rows).astype(np.int32)
        nbr = rng.integers(0, n_ids, size=(rows, stride)).astype(np.int32)
        nbr[rng.random((rows, stride)) < 0.15] = -1
        if rep == 2:
            nodes[:], nbr[:] = n_ids - 1, n_ids - 1            # everything collides on the last id
        nodes_t, nbr_t = torch.from_numpy(nodes).to(dev), torch.from_numpy(nbr).to(dev)
        live_t = None if live is None else torch.tensor([live], dtype=torch.int32, device=dev)
        uniq, num, nbr_idx, self_idx = g.unique_remap_bitmap(nodes_t, live_t, rows, nbr_t, stride, n_ids, ws)
        m = rows if live is None else live
        ids = np.concatenate([nodes[:m], nbr[:m].ravel()])
        want = np.unique(ids[ids >= 0])
        n = int(num.item())
        assert n == len(want) and np.array_equal(uniq[:n].cpu().numpy(), want)
        ni, si = nbr_idx.cpu().numpy(), self_idx.cpu().numpy()
        assert np.array_equal(ni[:m], np.where(nbr[:m] >= 0, np.searchsorted(want, np.maximum(nbr[:m], 0)), -1))
        assert np.all(ni[m:] == -1) and np.array_equal(si[:m], np.searchsorted(want, nodes[:m]))
        assert int(ws[:words * 4].view(torch.int32).abs().max().item()) == 0       # bitmap cleared for the next call


@pytest.mark.parametrize('name', ['cora_mean_sup', 'pubmed_selfloop_gcn', 'cora_3layer_gcn_max'])
def test_unique_remap_matches_reference_calls(g, dev, name):
    """Injected-sample contract: on the reference's own recorded samples,
    U_dev == U_ref as sets and every remapped slot points at the recorded neighbour."""
    inp, fx = cases.load_fixture(name)
    gcn = inp['spec']['gcn']
    bits = int(len(inp['rowptr']) - 1).bit_length()
    for nodes, samp, uniq_ref in fx['calls']:
        U, self_idx, cols, cnt = so.canonical_unique_remap(nodes, samp, drop_self=not gcn)
        width = cols.shape[1]
        nbr = np.where(cols >= 0, U[np.maximum(cols, 0)], -1).astype(np.int32)     # canonical id lists
        nodes_np = np.asarray(nodes, dtype=np.int32)
        uniq, num, nbr_idx, sidx = g.unique_remap(torch.from_numpy(nodes_np).to(dev), None, len(nodes),
                                                  torch.from_numpy(nbr).to(dev), width, bits)
        n = int(num.item())
        assert set(uniq[:n].cpu().tolist()) == set(uniq_ref) and n == len(uniq_ref)
        assert np.array_equal(uniq[:n].cpu().numpy(), U.astype(np.int32))
        assert np.array_equal(nbr_idx.cpu().numpy(), cols)
        assert np.array_equal(sidx.cpu().numpy(), self_idx.astype(np.int32))


# ------------------------------------------------------------------------------------------------
# K3 aggregation
# ------------------------------------------------------------------------------------------------
def _agg_ref(table, nbr, cnt, mode):
    rows = []
    for r in range(nbr.shape[0]):
        ids = nbr[r, :cnt[r]]
        if len(ids) == 0:
            rows.append(torch.full((table.shape[1],), float('nan')))
        elif mode == 0:
            rows.append(table[ids].sum(0) / len(ids))
        else:
            rows.append(table[ids].max(0)[0])
    return torch.stack(rows)


@pytest.mark.parametrize('dim', [100, 128, 1433, 500, 7])
@pytest.mark.parametrize('mode', [0, 1])
def test_agg_fwd_bwd(g, dev, dim, mode):
    rng = np.random.default_rng(dim + mode)
    n_table, rows, stride = 3000, 777, 11
    ld = (dim + 3) & ~3
    table = torch.zeros((n_table, ld))
    table[:, :dim] = torch.from_numpy(rng.standard_normal((n_table, dim)).astype(np.float32))
    cnt = rng.integers(0, stride + 1, size=rows).astype(np.int32)
    cnt[:5] = [0, 1, stride, 2, 10]
    nbr = np.full((rows, stride), -1, dtype=np.int32)
    for r in range(rows):
        nbr[r, :cnt[r]] = np.sort(rng.choice(n_table, size=cnt[r], replace=False))
    t_dev = table.to(dev)
    nbr_d, cnt_d = torch.from_numpy(nbr).to(dev), torch.from_numpy(cnt).to(dev)
    out, argmax = g.agg_fwd(t_dev, dim, nbr_d, stride, cnt_d, None, rows, mode)
    want = _agg_ref(table[:, :dim], nbr, cnt, mode)
    got = out[:, :dim].cpu()
    empty = cnt == 0
    assert torch.isnan(got[empty]).all()                      # 0/0 as the reference (MEAN); MAX raises there
    assert rel(got[~empty], want[~empty]) <= TOL
    if ld != dim:
        assert torch.all(out[~torch.from_numpy(empty).to(dev)][:, dim:] == 0)
    # backward against autograd on the same expression (rows with cnt == 0 contribute nothing)
    tbl = table[:, :dim].clone().requires_grad_(True)
    keep = np.where(~empty)[0]
    ref = _agg_ref(tbl, nbr[keep], cnt[keep], mode)
    gout = torch.from_numpy(rng.standard_normal((rows, ld)).astype(np.float32))
    gout[:, dim:] = 0
    (ref * gout[keep, :dim]).sum().backward()
    gself = torch.from_numpy(rng.standard_normal((rows, ld)).astype(np.float32))
    gself[:, dim:] = 0
    self_idx = rng.integers(0, n_table, size=rows).astype(np.int32)
    want_g = tbl.grad.clone()
    want_g.index_add_(0, torch.from_numpy(self_idx).long(), gself[:, :dim])
    gt = torch.zeros((n_table, ld), device=dev)
    g.agg_bwd(gout.to(dev), gself.to(dev), dim, nbr_d, stride, cnt_d, torch.from_numpy(self_idx).to(dev), argmax, None,
              rows, mode, gt)
    assert rel(gt[:, :dim], want_g) <= TOL
    # mask_table = the ReLU output the table rows came from: the scatter yields d(pre-activation),
    # i.e. the unmasked result multiplied by (h > 0) -- what scatter + relu_bwd_inplace produced before
    h = torch.zeros((n_table, ld))
    h[:, :dim] = torch.from_numpy(rng.standard_normal((n_table, dim)).astype(np.float32)).clamp_(min=0)
    gm = torch.zeros((n_table, ld), device=dev)
    g.agg_bwd(gout.to(dev), gself.to(dev), dim, nbr_d, stride, cnt_d, torch.from_numpy(self_idx).to(dev), argmax, None,
              rows, mode, gm, mask_table=h.to(dev))
    assert rel(gm[:, :dim], want_g * (h[:, :dim] > 0)) <= TOL
    assert torch.all(gm[:, :dim].cpu()[h[:, :dim] <= 0] == 0)


# ------------------------------------------------------------------------------------------------
# K4 SageLayer GEMM (fp32 path) and Classification
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dim,out_dim,gcn', [(100, 128, False), (128, 128, False), (1433, 128, False), (602, 128, True),
                                             (64, 32, False), (7, 5, False)])
def test_sage_gemm_fwd_bwd(g, dev, dim, out_dim, gcn):
    rng = np.random.default_rng(dim * 3 + out_dim)
    n_table, rows = 1500, 1027
    ld = (dim + 3) & ~3
    table = torch.zeros((n_table, ld))
    table[:, :dim] = torch.from_numpy(rng.standard_normal((n_table, dim)).astype(np.float32))
    agg = torch.zeros((rows, ld))
    agg[:, :dim] = torch.from_numpy(rng.standard_normal((rows, dim)).astype(np.float32))
    self_idx = rng.integers(0, n_table, size=rows)
    w = torch.from_numpy(rng.uniform(-0.2, 0.2, size=(out_dim, dim if gcn else 2 * dim)).astype(np.float32))
    live = 1000
    live_t = torch.tensor([live], dtype=torch.int32, device=dev)
    sidx_d = torch.from_numpy(self_idx.astype(np.int32)).to(dev)
    out = g.sage_gemm_fwd(None if gcn else table.to(dev), sidx_d, agg.to(dev), dim, w.to(dev), out_dim, gcn, live_t, rows)
    # torch fp32 reference of src/models.py:215-219
    w_ref = w.clone().requires_grad_(True)
    tbl = table[:, :dim].clone().requires_grad_(True)
    agg_ref = agg[:live, :dim].clone().requires_grad_(True)
    comb = agg_ref if gcn else torch.cat([tbl[self_idx[:live]], agg_ref], 1)
    want = torch.relu(w_ref.mm(comb.t())).t()
    assert rel(out[:live, :out_dim], want) <= TOL
    gout = torch.from_numpy(rng.standard_normal((rows, (out_dim + 3) & ~3)).astype(np.float32))
    (want * gout[:live, :out_dim]).sum().backward()
    gw = torch.zeros_like(w, device=dev)
    g.sage_gemm_bwd_w(None if gcn else table.to(dev), sidx_d, agg.to(dev), dim, gout.to(dev), out, out_dim, gcn, True,
                      live_t, rows, gw)
    assert rel(gw, w_ref.grad) <= TOL
    gs, ga = g.sage_gemm_bwd_x(gout.to(dev), out, w.to(dev), dim, out_dim, gcn, True, live_t, rows)
    assert rel(ga[:live, :dim], agg_ref.grad) <= TOL
    if not gcn:
        want_gs = torch.zeros((n_table, dim)).index_add_(0, torch.from_numpy(self_idx[:live]), gs[:live, :dim].cpu())
        assert rel(want_gs, tbl.grad) <= TOL


TC_TOL = {1: 2e-3, 2: 1e-5}      # GS_PREC_TF32 (10-bit mantissa operands) / GS_PREC_TF32X3 (fp32-faithful split)


@pytest.mark.parametrize('precision', [2, 1])
@pytest.mark.parametrize('dim,out_dim,gcn,rows,live', [(100, 128, False, 11264, 10900), (128, 128, False, 1024, 1024),
                                                       (1433, 128, False, 300, 257), (602, 128, True, 1000, 999),
                                                       (64, 32, False, 200, 130), (50, 256, False, 400, 400),
                                                       (7, 5, False, 64, 33)])
def test_sage_gemm_tensor_core_path(g, dev, precision, dim, out_dim, gcn, rows, live):
    """tcgen05 kind::tf32 path of K4 (forward, bwd_x, bwd_w) against a float64 evaluation of
    src/models.py:215-219 and its autograd."""
    rng = np.random.default_rng(dim * 5 + out_dim + precision)
    n_table = 3000
    ld = (dim + 3) & ~3
    table = torch.zeros((n_table, ld))
    table[:, :dim] = torch.from_numpy(rng.standard_normal((n_table, dim)).astype(np.float32))
    agg = torch.zeros((rows, ld))
    agg[:, :dim] = torch.from_numpy(rng.standard_normal((rows, dim)).astype(np.float32))
    self_idx = rng.integers(0, n_table, size=rows)
    w = torch.from_numpy(rng.uniform(-0.2, 0.2, size=(out_dim, dim if gcn else 2 * dim)).astype(np.float32))
    live_t = torch.tensor([live], dtype=torch.int32, device=dev)
    sidx_d = torch.from_numpy(self_idx.astype(np.int32)).to(dev)
    t_d, a_d, w_d = table.to(dev), agg.to(dev), w.to(dev)
    out = g.sage_gemm_fwd(None if gcn else t_d, sidx_d, a_d, dim, w_d, out_dim, gcn, live_t, rows, True, precision)
    w_ref = w.double().requires_grad_(True)
    tbl = table[:, :dim].double().requires_grad_(True)
    agg_ref = agg[:live, :dim].double().requires_grad_(True)
    comb = agg_ref if gcn else torch.cat([tbl[self_idx[:live]], agg_ref], 1)
    want = torch.relu(w_ref.mm(comb.t())).t()
    tol = TC_TOL[precision]
    assert rel(out[:live, :out_dim], want) <= tol
    # backward uses the fp32-path forward output for the ReLU mask so both sides mask identically
    out32 = g.sage_gemm_fwd(None if gcn else t_d, sidx_d, a_d, dim, w_d, out_dim, gcn, live_t, rows, True, 0)
    gout = torch.from_numpy(rng.standard_normal((rows, (out_dim + 3) & ~3)).astype(np.float32))
    mask = (out32[:live, :out_dim] > 0).cpu().double()
    pre = w_ref.mm(comb.t()).t()
    (pre * mask * gout[:live, :out_dim].double()).sum().backward()
    gw = torch.zeros_like(w, device=dev)
    g.sage_gemm_bwd_w(None if gcn else t_d, sidx_d, a_d, dim, gout.to(dev), out32, out_dim, gcn, True, live_t, rows, gw,
                      precision=precision)
    assert rel(gw, w_ref.grad) <= tol
    gs, ga = g.sage_gemm_bwd_x(gout.to(dev), out32, w_d, dim, out_dim, gcn, True, live_t, rows, precision=precision)
    assert rel(ga[:live, :dim], agg_ref.grad) <= tol
    if not gcn:
        want_gs = torch.zeros((n_table, dim), dtype=torch.float64).index_add_(
            0, torch.from_numpy(self_idx[:live]), gs[:live, :dim].cpu().double())
        assert rel(want_gs, tbl.grad) <= tol
    # pre-masked gradient + relu=False: the cp.async-staged variants of the same kernels
    gm = gout.to(dev).clone()
    g.relu_bwd_inplace(gm, out32, out_dim, live_t, rows)
    gw2 = torch.zeros_like(w, device=dev)
    g.sage_gemm_bwd_w(None if gcn else t_d, sidx_d, a_d, dim, gm, out32, out_dim, gcn, False, live_t, rows, gw2,
                      precision=precision)
    assert rel(gw2, w_ref.grad) <= tol
    gs2, ga2 = g.sage_gemm_bwd_x(gm, out32, w_d, dim, out_dim, gcn, False, live_t, rows, precision=precision)
    assert rel(ga2[:live, :dim], agg_ref.grad) <= tol
    if not gcn:
        assert rel(gs2[:live, :dim], gs[:live, :dim]) <= tol


@pytest.mark.parametrize('rows,dim,classes', [(1024, 128, 47), (150, 32, 7), (33, 128, 3)])
def test_classifier_and_nll(g, dev, rows, dim, classes):
    rng = np.random.default_rng(classes)
    emb = torch.from_numpy(rng.standard_normal((rows, dim)).astype(np.float32))
    w = torch.from_numpy(rng.uniform(-0.3, 0.3, size=(classes, dim)).astype(np.float32))
    b = torch.from_numpy(rng.uniform(-0.1, 0.1, size=(classes,)).astype(np.float32))
    labels = torch.from_numpy(rng.integers(0, classes, size=rows))
    e_ref, w_ref, b_ref = emb.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    logp_ref = so.classification(w_ref, b_ref, e_ref)
    loss_ref = so.supervised_loss(logp_ref, labels.numpy())
    loss_ref.backward()
    logp = g.cls_fwd(emb.to(dev), dim, w.to(dev), b.to(dev), classes)
    assert rel(logp, logp_ref) <= TOL
    loss, glogp = g.nll_fwd_bwd(logp, labels.to(dev))
    assert rel(loss, loss_ref.reshape(1)) <= TOL
    ge, gw, gb = torch.empty((rows, dim), device=dev), torch.zeros((classes, dim), device=dev), torch.zeros((classes,), device=dev)
    g.cls_bwd(glogp, logp, emb.to(dev), dim, w.to(dev), classes, ge, gw, gb)
    assert rel(ge, e_ref.grad) <= TOL and rel(gw, w_ref.grad) <= TOL and rel(gb, b_ref.grad) <= TOL


def test_clip_sgd_matches_torch(g, dev):
    rng = np.random.default_rng(3)
    shapes = [(128, 200), (128, 256), (47, 128), (47,)]
    for scale in (0.01, 10.0):                                  # below and above the clip threshold
        ps = [torch.from_numpy(rng.standard_normal(s).astype(np.float32)) for s in shapes]
        gs_ = [torch.from_numpy((scale * rng.standard_normal(s)).astype(np.float32)) for s in shapes]
        ref = [p.clone().requires_grad_(True) for p in ps]
        for p, gr in zip(ref, gs_):
            p.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_(ref, 5)                  # src/utils.py:186
        torch.optim.SGD(ref, lr=0.7).step()                     # src/utils.py:136,187
        pd, gd = [p.to(dev) for p in ps], [x.to(dev) for x in gs_]
        tl = g.TensorList(pd, gd)
        g.clip_sgd(tl, 5.0, 0.7, 1.0, zero_grads=True)
        for a, b in zip(pd, ref):
            assert rel(a, b) <= 1e-6
        assert all(float(x.abs().max()) == 0.0 for x in gd)


# ------------------------------------------------------------------------------------------------
# supervised tail in one launch (classifier + log_softmax + NLL + backward), src/models.py:25-27, src/utils.py:153,162-163
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('rows,dim,classes,mask', [(1024, 128, 47, True), (1000, 128, 7, False), (37, 64, 3, True),
                                                    (513, 256, 64, False), (200, 128, 41, True), (64, 128, 100, True)])
def test_cls_nll_fwd_bwd_matches_torch(g, dev, rows, dim, classes, mask):
    from graphsage_b200 import native
    rng = np.random.default_rng(rows + classes)
    emb = np.maximum(rng.standard_normal((rows, dim)), 0).astype(np.float32)       # a ReLU output, as in the model
    w = rng.uniform(-0.3, 0.3, (classes, dim)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, (classes,)).astype(np.float32)
    n_nodes = 5000
    labels = rng.integers(0, classes, n_nodes).astype(np.int64)
    idx = rng.integers(0, n_nodes, rows).astype(np.int32)
    te = torch.from_numpy(emb).requires_grad_(True)
    tw = torch.from_numpy(w).requires_grad_(True)
    tb = torch.from_numpy(b).requires_grad_(True)
    logp = torch.log_softmax(te @ tw.t() + tb, dim=1)
    y = torch.from_numpy(labels[idx])
    loss = -torch.sum(logp[range(rows), y], 0) / rows                              # src/utils.py:162-163
    loss.backward()
    want_ge = te.grad * (te.detach() > 0) if mask else te.grad
    d = lambda a: torch.from_numpy(a).to(dev)
    e_d, w_d, b_d = d(emb), d(w), d(b)
    loss_d = torch.full((1,), 7.0, device=dev)
    ge = torch.full((rows, dim), 9.0, device=dev)
    gw, gb = torch.zeros((classes, dim), device=dev), torch.zeros((classes,), device=dev)
    lp = g.cls_nll_fwd_bwd(e_d, dim, w_d, b_d, classes, d(labels), d(idx), loss_d, ge, gw, gb,
                           precision=native.PREC_FP32, mask_relu_input=mask)
    assert rel(lp, logp) <= TOL
    assert abs(float(loss_d.item()) - float(loss)) <= TOL * abs(float(loss))
    assert rel(ge, want_ge) <= TOL
    assert rel(gw, tw.grad) <= TOL
    assert rel(gb, tb.grad) <= TOL


def test_sampler_treats_repeated_ids_in_a_row_as_a_set(g, dev):
    """Device-generated graphs may repeat a neighbour inside a CSR row; the reference's rows are sets
    (src/dataCenter.py:33), so the drawn list must be distinct, ascending and fully written."""
    rng = np.random.default_rng(0)
    n, k, stride = 64, 10, 11
    rows = [rng.integers(0, 8, size=rng.integers(1, 40)).astype(np.int32) for _ in range(n)]     # ids 0..7: many repeats
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum([len(r) for r in rows])
    col = np.concatenate(rows)
    rp, cl = _csr_dev(rowptr, col, dev)
    nodes = torch.arange(n, dtype=torch.int32, device=dev)
    for self_mode in (0, 1, 2):
        for off in range(20):
            out = torch.full((n, stride), 12345, dtype=torch.int32, device=dev)               # poison: every slot must be written
            cnt = torch.full((n,), -7, dtype=torch.int32, device=dev)
            g.sample_neighbors(rp, cl, n, nodes, None, n, k, stride, self_mode, 7, off, out_nbr=out, out_cnt=cnt)
            o, c = out.cpu().numpy(), cnt.cpu().numpy()
            for v in range(n):
                row = o[v, :c[v]]
                assert np.all(o[v, c[v]:] == -1) and np.all(np.diff(row) > 0)
                allowed = set(rows[v].tolist()) | ({v} if self_mode == 2 else set())
                assert set(row.tolist()) <= allowed
                if self_mode == 1:
                    assert v not in row
                if self_mode == 2:
                    assert v in row


# ------------------------------------------------------------------------------------------------
# the fused top layer (csrc/sage_top.cu): gather + mean + SageLayer + classifier + NLL + backward + scatter
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('gcn,classes,rows,live,precision,tol', [
    (False, 47, 1000, 777, 'tf32x3', 1e-5), (False, 47, 1024, None, 'tf32x3', 1e-5), (True, 41, 333, None, 'tf32x3', 1e-5),
    (False, 3, 16, 5, 'tf32x3', 1e-5), (False, 64, 2500, None, 'tf32x3', 1e-5), (False, 47, 1000, None, 'tf32', 3e-3)])
def test_sage_top_sup_matches_torch_autograd(g, dev, gcn, classes, rows, live, precision, tol):
    """One launch against torch fp64 autograd of the same top layer (src/models.py:209-220,25-27,300-314 and
    src/utils.py:162-163): outputs, loss, classifier gradients, dZ and the ReLU-gated scatter into the gradient of
    the layer below; `live` < rows exercises the device-side row count, rows % 16 != 0 the ragged last tile,
    2500 rows the persistent multi-tile loop, repeated / shared neighbours the atomics."""
    from graphsage_b200 import native, ops
    H, fan = 128, 10
    rng = np.random.default_rng(rows + classes)
    n_prev = max(64, rows * 3)
    table = np.maximum(rng.standard_normal((n_prev, H)), 0).astype(np.float32)          # a ReLU output: ~half zeros
    stride = fan + (1 if gcn else 0)
    nbr = np.full((rows, stride), -1, dtype=np.int32)
    cnt = rng.integers(1, fan + 1, size=rows).astype(np.int32)
    self_idx = rng.integers(0, n_prev, size=rows).astype(np.int32)
    for r in range(rows):
        ids = rng.choice(n_prev, size=cnt[r], replace=False)
        if gcn:
            ids = np.unique(np.concatenate([ids, [self_idx[r]]]))
            cnt[r] = len(ids)
        nbr[r, :cnt[r]] = np.sort(ids)
    k = H if gcn else 2 * H
    w = (rng.standard_normal((H, k)) * 0.1).astype(np.float32)
    cw = (rng.standard_normal((classes, H)) * 0.2).astype(np.float32)
    cb = (rng.standard_normal(classes) * 0.1).astype(np.float32)
    n_nodes = 5000
    node_of_row = rng.integers(0, n_nodes, size=rows).astype(np.int32)
    labels = rng.integers(0, classes, size=n_nodes).astype(np.int64)
    n_live = rows if live is None else live
    # ---- torch fp64 reference ----
    t = torch.tensor(table, dtype=torch.float64, requires_grad=True)
    W = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    CW = torch.tensor(cw, dtype=torch.float64, requires_grad=True)
    CB = torch.tensor(cb, dtype=torch.float64, requires_grad=True)
    aggs = torch.stack([t[torch.as_tensor(nbr[r, :cnt[r]].astype(np.int64))].mean(0) for r in range(n_live)])
    X = aggs if gcn else torch.cat([t[torch.as_tensor(self_idx[:n_live].astype(np.int64))], aggs], 1)
    z = X @ W.t()
    z.retain_grad()
    h = torch.relu(z)
    logp = torch.log_softmax(h @ CW.t() + CB, 1)
    y = torch.as_tensor(labels[node_of_row[:n_live]])
    loss = -logp[torch.arange(n_live), y].sum() / n_live
    loss.backward()
    want_gt = t.grad * (t.detach() > 0)
    # ---- device ----
    d = lambda a: torch.from_numpy(a).to(dev)
    table_d, g_table = d(table), torch.zeros((n_prev, H), device=dev)
    gcw, gcb = torch.zeros((classes, H), device=dev), torch.zeros((classes,), device=dev)
    loss_d = torch.full((1,), 123.0, device=dev)
    logp_d = torch.zeros((rows, classes), device=dev)
    num_rows = None if live is None else torch.tensor([live], dtype=torch.int32, device=dev)
    ws = ops.sage_top_workspace(dev)
    prec = {'tf32x3': native.PREC_TF32X3, 'tf32': native.PREC_TF32}[precision]
    # rows == 1024: the classifier gradients spread over 8 replicas, as the trainers launch the kernel
    rep_w = torch.zeros((7, classes * H), device=dev) if rows == 1024 else None
    rep_b = torch.zeros((7, 64), device=dev) if rows == 1024 else None
    for rep in range(2):                                   # twice: the ticket / partials of the workspace reset themselves
        g_table.zero_(); gcw.zero_(); gcb.zero_()
        out_h, out_agg, out_dz = ops.sage_top_sup(table_d, d(nbr), stride, d(cnt), d(self_idx), num_rows, rows, d(w), gcn,
                                                  d(cw), d(cb), d(labels), d(node_of_row), loss_d, gcw, gcb, g_table, ws,
                                                  prec, logp=logp_d, cls_w_rep=rep_w, cls_b_rep=rep_b)
        torch.cuda.synchronize()
        if rep_w is not None:
            assert float(rep_w.abs().max()) > 0                      # the replicas really took part of the adds
            gcw += rep_w.sum(0).view(classes, H); gcb += rep_b[:, :classes].sum(0)
            rep_w.zero_(); rep_b.zero_()
        assert rel(out_agg[:n_live], aggs) <= 1e-6
        assert rel(out_h[:n_live], h) <= tol
        assert rel(logp_d[:n_live], logp) <= tol
        assert rel(loss_d, loss.reshape(1)) <= tol
        if precision == 'tf32':       # single-pass tf32: pre-activations within 1e-3 of zero gate differently than in fp64,
            continue                  # which is a property of the reduced-precision mode, not of the kernel
        assert rel(out_dz[:n_live], z.grad) <= tol * 3
        assert rel(gcw, CW.grad) <= tol * 3 and rel(gcb, CB.grad) <= tol * 3
        assert rel(g_table, want_gt) <= tol * 3
    # the weight gradient of the layer from the saved operands (gs_sage_gemm_bwd_w, relu = 0)
    if precision == 'tf32':
        return
    gw = torch.zeros((H, k), device=dev)
    ops.sage_gemm_bwd_w(None if gcn else table_d, d(self_idx), out_agg, H, out_dz, out_h, H, gcn, False, num_rows, rows, gw,
                        precision=prec)
    assert rel(gw, W.grad) <= tol * 3
    # unsupported shapes are refused, not mis-computed
    assert not ops.sage_top_supported(64, 64, 7, 10, native.PREC_TF32X3, True)
    assert not ops.sage_top_supported(128, 128, 7, 10, native.PREC_FP32, True)
    assert not ops.sage_top_supported(128, 128, 100, 10, native.PREC_TF32X3, True)


@pytest.mark.parametrize('precision', ['tf32x3', 'fp32'])
def test_sage_gemm_bwd_w_pair_equals_two_launches(g, dev, precision):
    """gs_sage_gemm_bwd_w_pair: the weight gradients of two layers (cfg-3 shapes: 10.9K x 200 and 1024 x 256 into
    128 outputs) as ONE grid must equal the two separate launches and the fp64 products."""
    from graphsage_b200 import native
    prec = {'tf32x3': native.PREC_TF32X3, 'fp32': native.PREC_FP32}[precision]
    rng = np.random.default_rng(3)
    H = 128
    shapes = [(10900, 11264, 100, 30000), (1000, 1024, 128, 10900)]        # live rows, max rows, dim, table rows
    probs, want = [], []
    for live, mx, dim, n_tab in shapes:
        tab = torch.from_numpy(rng.standard_normal((n_tab, dim)).astype(np.float32)).to(dev)
        sidx = torch.from_numpy(rng.integers(0, n_tab, size=mx).astype(np.int32)).to(dev)
        agg = torch.from_numpy(rng.standard_normal((mx, dim)).astype(np.float32)).to(dev)
        dz = torch.from_numpy((rng.standard_normal((mx, H)) * (rng.random((mx, H)) > 0.5)).astype(np.float32)).to(dev)
        out = torch.ones((mx, H), device=dev)
        nr = torch.tensor([live], dtype=torch.int32, device=dev)
        gw = torch.zeros((H, 2 * dim), device=dev)
        probs.append((tab, sidx, agg, dim, dz, out, H, nr, mx, gw))
        X = torch.cat([tab[sidx[:live].long()], agg[:live]], 1).double()
        want.append(dz[:live].double().t() @ X)
    g.sage_gemm_bwd_w_pair(probs, False, False, prec)
    for (tab, sidx, agg, dim, dz, out, _, nr, mx, gw), w in zip(probs, want):
        single = torch.zeros_like(gw)
        g.sage_gemm_bwd_w(tab, sidx, agg, dim, dz, out, H, False, False, nr, mx, single, precision=prec)
        assert rel(gw, w) <= TOL and rel(single, w) <= TOL


@pytest.mark.parametrize('classes,live', [(47, 1000), (64, None), (3, 130)])
def test_classifier_weight_gradient_as_third_problem_of_the_group(g, dev, classes, live):
    """The step's three weight gradients as ONE grid (gs_sage_gemm_bwd_w_group): the two layers' and the classifier's,
    the latter from d(logits) saved by the top-layer kernel (out_dlog, 64 zero-padded columns, grad_cls_w = NULL) and the
    layer output h -- against the fp64 products; the rows of grad Wc beyond num_classes do not exist and what follows
    the buffer must stay untouched."""
    from graphsage_b200 import native, ops
    prec = native.PREC_TF32X3
    rng = np.random.default_rng(classes)
    H, fan, rows = 128, 10, 1024
    n_live = rows if live is None else live
    n_prev = 10900
    d = lambda a: torch.from_numpy(a).to(dev)
    table = np.maximum(rng.standard_normal((n_prev, H)), 0).astype(np.float32)
    nbr = np.full((rows, fan), -1, dtype=np.int32)
    cnt = rng.integers(1, fan + 1, size=rows).astype(np.int32)
    self_idx = rng.integers(0, n_prev, size=rows).astype(np.int32)
    for r in range(rows):
        nbr[r, :cnt[r]] = np.sort(rng.choice(n_prev, size=cnt[r], replace=False))
    w = (rng.standard_normal((H, 2 * H)) * 0.1).astype(np.float32)
    cw = (rng.standard_normal((classes, H)) * 0.2).astype(np.float32)
    cb = (rng.standard_normal(classes) * 0.1).astype(np.float32)
    labels = rng.integers(0, classes, size=5000).astype(np.int64)
    node_of_row = rng.integers(0, 5000, size=rows).astype(np.int32)
    num_rows = None if live is None else torch.tensor([live], dtype=torch.int32, device=dev)
    table_d, g_table = d(table), torch.zeros((n_prev, H), device=dev)
    guard = torch.full((classes * H + 64,), 7.0, device=dev)          # grad Wc followed by a guard zone
    gcw = guard[:classes * H].view(classes, H)
    gcw.zero_()
    gcb, loss_d = torch.zeros((classes,), device=dev), torch.zeros((1,), device=dev)
    dlog = torch.full((rows, 64), 9.0, device=dev)
    ws = ops.sage_top_workspace(dev)
    out_h, out_agg, out_dz = ops.sage_top_sup(table_d, d(nbr), fan, d(cnt), d(self_idx), num_rows, rows, d(w), False, d(cw),
                                              d(cb), d(labels), d(node_of_row), loss_d, None, gcb, g_table, ws, prec,
                                              out_dlog=dlog)
    torch.cuda.synchronize()
    assert float(gcw.abs().max()) == 0.0                                # the kernel left the weight gradient alone
    # d(logits) against torch: (softmax - onehot) / rows, zero beyond the classes
    hh = out_h[:n_live].double()
    logp = torch.log_softmax(hh @ d(cw).double().t() + d(cb).double(), 1)
    y = d(labels)[d(node_of_row)[:n_live].long()]
    want_dlog = logp.exp()
    want_dlog[torch.arange(n_live, device=dev), y] -= 1.0
    want_dlog /= n_live
    assert rel(dlog[:n_live, :classes], want_dlog) <= 1e-5
    assert float(dlog[:n_live, classes:].abs().max() if classes < 64 else 0.0) == 0.0
    # the three problems in one launch
    tab1 = d(rng.standard_normal((30000, 100)).astype(np.float32))
    sidx1 = d(rng.integers(0, 30000, size=11264).astype(np.int32))
    agg1 = d(rng.standard_normal((11264, 100)).astype(np.float32))
    dz1 = d((rng.standard_normal((11264, H)) * (rng.random((11264, H)) > 0.5)).astype(np.float32))
    nr1 = torch.tensor([10900], dtype=torch.int32, device=dev)
    gw1, gw2 = torch.zeros((H, 200), device=dev), torch.zeros((H, 2 * H), device=dev)
    ops.sage_gemm_bwd_w_group([
        (tab1, sidx1, agg1, 100, dz1, None, H, nr1, 11264, gw1, 0, 0, 0),
        (table_d, d(self_idx), out_agg, H, out_dz, None, H, num_rows, rows, gw2, 0, 0, 0),
        (None, None, out_h, H, dlog, None, classes, num_rows, rows, gcw, 1, 0, (classes + 3) & ~3)], prec)
    torch.cuda.synchronize()
    X1 = torch.cat([tab1[sidx1[:10900].long()], agg1[:10900]], 1).double()
    assert rel(gw1, dz1[:10900].double().t() @ X1) <= TOL
    X2 = torch.cat([table_d[d(self_idx)[:n_live].long()], out_agg[:n_live]], 1).double()
    assert rel(gw2, out_dz[:n_live].double().t() @ X2) <= TOL
    assert rel(gcw, dlog[:n_live, :classes].double().t() @ hh) <= TOL
    assert torch.all(guard[classes * H:] == 7.0)


_TMA_GEMM_CHECK = r"""
import numpy as np, torch, sys
sys.path.insert(0, %r)
import graphsage_b200
from graphsage_b200 import native, ops
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
for (n_tab, rows, mx, dim, H, gcn, prec) in [(50000, 10500, 11264, 100, 128, False, native.PREC_TF32X3),
                                              (3000, 300, 384, 128, 64, False, native.PREC_TF32X3),
                                              (3000, 1000, 1000, 36, 128, True, native.PREC_TF32X3),
                                              (3000, 700, 1024, 100, 128, False, native.PREC_TF32)]:
    tab = torch.from_numpy(rng.standard_normal((n_tab, dim)).astype(np.float32)).to(dev)
    sidx = torch.from_numpy(rng.integers(0, n_tab, size=mx).astype(np.int32)).to(dev)
    agg = torch.from_numpy(rng.standard_normal((mx, dim)).astype(np.float32)).to(dev)
    k = dim if gcn else 2 * dim
    w = torch.from_numpy((rng.standard_normal((H, k)) * 0.1).astype(np.float32)).to(dev)
    nr = torch.tensor([rows], dtype=torch.int32, device=dev)
    X = (agg[:rows] if gcn else torch.cat([tab[sidx[:rows].long()], agg[:rows]], 1)).double()
    want = torch.relu(X @ w.double().t())
    for wl in (None, ops.split_lo(w)):
        out = torch.zeros((mx, H), device=dev)
        ops.sage_gemm_fwd(None if gcn else tab, sidx, agg, dim, w, H, gcn, nr, mx, relu=True, precision=prec, out=out, weight_lo=wl)
        torch.cuda.synchronize()
        err = float((out[:rows].double() - want).abs().max() / want.abs().max())
        assert err <= (1e-5 if prec == native.PREC_TF32X3 else 3e-3), (dim, H, gcn, err)
        assert rows == mx or float(out[rows:].abs().max()) == 0.0
print('TMA_GEMM_OK')
"""


@pytest.mark.parametrize('self_mode', ['threads', 'gather4'])
def test_forward_gemm_takes_the_tma_kernel_and_both_self_row_paths_agree_with_fp64(dev, self_mode):
    """The gathered-operand forward GEMM of csrc/sage_gemm_tma.cu must be the kernel that runs for 16-byte-aligned
    widths (GS_TMA_GATHER=2: a refusal is an error instead of a silent fall-back to the thread-staged kernel), with the
    self rows fetched by the epilogue warps (default) or by TMA gather4, W_lo supplied or split in the kernel, gcn,
    ragged tiles and a device-side row count.  The switches are read once per process, hence the subprocess."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GS_TMA_GATHER='2', GS_TMA_SELF=self_mode)
    r = subprocess.run([sys.executable, '-c', _TMA_GEMM_CHECK % root], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'TMA_GEMM_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_fused_sampler_unique_chain_equals_numpy(g, dev):
    """The 5-launch preparation chain: sample(+fetch from the device queue, +mark) -> bitmap scan -> emit/remap ->
    sample(+clear).  Unique ids / remap indices must equal numpy's on the drawn lists, the bitmap must be all-zero
    again afterwards, the queue cursor must advance by one per fetch (and wrap), and the draws must equal those of
    the plain sampler on the same (seed, offset) -- the extras never change what is drawn."""
    from graphsage_b200 import native, synth
    n = 20000
    rowptr_h, col_h = synth.powerlaw_graph(n, n * 12, seed=4)
    rowptr, col = _csr_dev(rowptr_h, col_h, dev)
    b_sz, k = 256, 10
    rng = np.random.default_rng(1)
    queue = torch.from_numpy(rng.integers(0, n, size=(3, b_sz)).astype(np.int32)).to(dev)
    desc = torch.tensor([queue.data_ptr(), 3, 0, 0], dtype=torch.int64, device=dev)
    ws = g.unique_bitmap_workspace(n, dev)
    words = (n + 31) // 32
    for step in range(4):                                   # 4 > 3 rows: the cursor wraps
        seeds = torch.full((b_sz,), -7, dtype=torch.int32, device=dev)
        nbr, cnt = g.sample_neighbors(rowptr, col, n, None, None, b_sz, k, k, native.SELF_DROP, 99, (step << 8) | 2,
                                      queue_desc=desc, fetch_dst=seeds, mark_bitmap=ws)
        want_seeds = queue[step % 3]
        assert torch.equal(seeds, want_seeds)
        assert desc.cpu().tolist()[2:] == [step + 1, 0]
        plain_nbr, plain_cnt = g.sample_neighbors(rowptr, col, n, want_seeds.contiguous(), None, b_sz, k, k, native.SELF_DROP,
                                                  99, (step << 8) | 2)
        assert torch.equal(nbr, plain_nbr) and torch.equal(cnt, plain_cnt)
        uniq, num_uniq, nbr_idx, self_idx = g.unique_remap_bitmap(seeds, None, b_sz, nbr, k, n, ws,
                                                                  flags=native.UNIQUE_MARKED | native.UNIQUE_LEAVE_MARKS)
        nbr_h, seeds_h = nbr.cpu().numpy(), seeds.cpu().numpy()
        want = np.unique(np.concatenate([seeds_h, nbr_h[nbr_h >= 0]]))
        nu = int(num_uniq.item())
        assert nu == len(want) and np.array_equal(uniq[:nu].cpu().numpy(), want)
        idx_h = nbr_idx.cpu().numpy()
        assert np.array_equal(want[idx_h[nbr_h >= 0]], nbr_h[nbr_h >= 0]) and (idx_h[nbr_h < 0] == -1).all()
        assert np.array_equal(want[self_idx.cpu().numpy()], seeds_h)
        # the next sampler (rows = the unique ids) clears the words it owns
        nbr1, cnt1 = g.sample_neighbors(rowptr, col, n, uniq, num_uniq, uniq.shape[0], k, k, native.SELF_DROP, 99,
                                        (step << 8) | 1, clear_bitmap=ws)
        assert int(ws[:words * 4].view(torch.int32).abs().sum().item()) == 0
        assert int((cnt1[:nu] > 0).sum().item()) == nu


@pytest.mark.parametrize('mode,gcn,dim', [(0, False, 100), (0, True, 100), (1, False, 64), (0, False, 500)])
def test_agg_fwd_x_writes_the_dense_input_rows(g, dev, mode, gcn, dim):
    """gs_agg_fwd_x: X[r] = [table[self] | mean/max of the row's list] and the low tf32 halves x - trunc(x);
    hi (= the raw word with its low 13 mantissa bits cleared) + lo must reproduce x exactly."""
    rng = np.random.default_rng(dim + mode)
    n, rows, stride = 5000, 700, 11 if gcn else 10
    table = torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)).to(dev)
    cnt = rng.integers(0 if mode == 0 else 1, stride + 1, size=rows).astype(np.int32)
    nbr = np.full((rows, stride), -1, dtype=np.int32)
    for r in range(rows):
        nbr[r, :cnt[r]] = np.sort(rng.choice(n, size=cnt[r], replace=False))
    nodes = rng.integers(0, n, size=rows).astype(np.int32)
    live = torch.tensor([rows - 13], dtype=torch.int32, device=dev)
    x, x_lo = g.agg_fwd_x(table, dim, torch.from_numpy(nbr).to(dev), stride, torch.from_numpy(cnt).to(dev),
                          None if gcn else torch.from_numpy(nodes).to(dev), live, rows, mode)
    want_agg = _agg_ref(table.cpu(), nbr[:rows - 13].astype(np.int64), cnt[:rows - 13], mode).numpy()
    got = x[:rows - 13].cpu().numpy()
    if gcn:
        assert got.shape[1] == dim and np.allclose(got, want_agg, rtol=1e-6, atol=1e-6, equal_nan=True)
    else:
        assert got.shape[1] == 2 * dim
        assert np.array_equal(got[:, :dim], table.cpu().numpy()[nodes[:rows - 13]])
        assert np.allclose(got[:, dim:], want_agg, rtol=1e-6, atol=1e-6, equal_nan=True)
    hi = (x[:rows - 13].view(torch.int32) & -8192).view(torch.float32)
    ok = torch.isfinite(x[:rows - 13])
    assert torch.equal((hi + x_lo[:rows - 13])[ok], x[:rows - 13][ok])


@pytest.mark.parametrize('precision,tol', [('tf32x3', 1e-5), ('tf32', 2e-3)])
@pytest.mark.parametrize('rows,live,dim,gcn', [(10900, 10000, 100, False), (300, None, 64, False), (1000, None, 128, True)])
def test_sage_gemm_dense_tma_path(g, dev, precision, tol, rows, live, dim, gcn):
    """The all-TMA forward kernel (dense [self | agg] rows + pre-split low halves) against fp64, with the zero fill
    of the gradient buffer, and against the gathered-operand kernel on the same numbers."""
    from graphsage_b200 import native
    prec = {'tf32x3': native.PREC_TF32X3, 'tf32': native.PREC_TF32}[precision]
    rng = np.random.default_rng(rows + dim)
    H, k = 128, dim if gcn else 2 * dim
    x = torch.from_numpy(rng.standard_normal((rows, k)).astype(np.float32)).to(dev)
    w = torch.from_numpy((rng.standard_normal((H, k)) * 0.1).astype(np.float32)).to(dev)
    x_lo, w_lo = g.split_lo(x), g.split_lo(w)
    nr = None if live is None else torch.tensor([live], dtype=torch.int32, device=dev)
    n_live = rows if live is None else live
    zero = torch.full((rows, H), 7.0, device=dev)
    self_t, agg = (None, x) if gcn else (x[:, :dim], x[:, dim:])
    native.launch_count_reset()
    out = g.sage_gemm_fwd(self_t, None, agg, dim, w, H, gcn, nr, rows, True, prec, zero_out=zero, x_lo=x_lo, weight_lo=w_lo)
    want = torch.relu(x[:n_live].double() @ w.double().t())
    assert rel(out[:n_live], want) <= tol
    assert float(zero[:n_live].abs().max()) == 0.0 and (live is None or float(zero[n_live:].min()) == 7.0)
    # the gathered-operand kernel on a copy whose halves are NOT adjacent (so it cannot take the dense path)
    agg2 = agg.clone()
    out2 = g.sage_gemm_fwd(self_t, None if gcn else torch.arange(rows, dtype=torch.int32, device=dev), agg2, dim, w, H, gcn, nr,
                           rows, True, prec)
    assert rel(out2[:n_live], want) <= tol


# ------------------------------------------------------------------------------------------------
# distributional tests of the native samplers (they cannot match random's stream)
# ------------------------------------------------------------------------------------------------
def test_sampler_distribution_on_a_hub_row(g, dev):
    """K1 on a hub (degree 5000, the cfg-3 case: Floyd's subset sampling over thousands of positions).  Every
    neighbour must be included with probability k/deg (chi-square over 5000 cells), no draw may repeat inside a row,
    and the draws of two rows / two offsets must be independent."""
    deg, k, trials = 5000, 10, 200_000
    rowptr = np.array([0, deg], dtype=np.int64)
    col = np.arange(1, deg + 1, dtype=np.int32)
    rp, cl = _csr_dev(rowptr, col, dev)
    nodes = torch.zeros((trials,), dtype=torch.int32, device=dev)
    nbr, cnt = g.sample_neighbors(rp, cl, deg + 1, nodes, None, trials, k, k, 0, 4242, 7)
    nbr = nbr.cpu().numpy()
    assert np.all(cnt.cpu().numpy() == k) and np.all(np.diff(nbr, axis=1) > 0)      # k distinct ids, ascending
    counts = np.bincount(nbr.ravel(), minlength=deg + 1)[1:]
    expect = trials * k / deg                                                        # 400 per cell
    chi2 = ((counts - expect) ** 2 / (expect * (1 - k / deg))).sum()
    dof = deg - 1
    assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof), chi2                              # +-6 sigma of chi2(4999)
    # the first and the last position of the row are drawn as often as any other (Floyd's "take j" branch)
    assert abs(counts[0] - expect) < 6 * np.sqrt(expect) and abs(counts[-1] - expect) < 6 * np.sqrt(expect)
    # positions are not correlated with the pick order: mean of the j-th smallest id follows the order statistics
    want_mean = (np.arange(1, k + 1) * (deg + 1)) / (k + 1)
    assert np.allclose(nbr.mean(axis=0), want_mean, rtol=0.01)
    # another offset gives another, equally uniform, draw; rows are independent of each other
    nbr2, _ = g.sample_neighbors(rp, cl, deg + 1, nodes, None, trials, k, k, 0, 4242, 8)
    nbr2 = nbr2.cpu().numpy()
    same = (nbr[:20000, None, :] == nbr2[:20000, :, None]).any(axis=1).sum(axis=1).mean()
    assert abs(same - k * k / deg) < 0.01, same                                      # E|A n B| = k^2/deg = 0.02


def test_random_walk_positives_are_uniform_over_the_neighbour_row(g, dev):
    """K5a, src/models.py:178: `random.choice(neighs)` -- each of the 6 one-step walks of a seed picks a neighbour
    uniformly; picks outside the train set or equal to the seed yield no pair (-1) but still consume a draw."""
    deg, walks, trials = 23, 6, 60_000
    n = deg + 1
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = deg                                       # node 0 has neighbours 1..deg, the others none
    col = np.arange(1, deg + 1, dtype=np.int32)
    rp, cl = _csr_dev(rowptr, col, dev)
    is_train = np.ones(n, dtype=np.uint8)
    is_train[[3, 7]] = 0                                   # two neighbours outside the train set: never a pair
    seeds = torch.zeros((trials,), dtype=torch.int32, device=dev)
    pos = g.random_walk_pos(rp, cl, n, seeds, walks, 1, torch.from_numpy(is_train).to(dev), 99, 3).cpu().numpy()
    assert pos.shape == (trials, walks)
    got = pos[pos >= 0]
    assert not np.isin(got, [0, 3, 7]).any()
    frac_pairs = (pos >= 0).mean()
    assert abs(frac_pairs - (deg - 2) / deg) < 0.005                                 # 2 of 23 choices are rejected
    counts = np.bincount(got, minlength=n)[1:]
    live = np.delete(counts, [2, 6])                                                 # neighbours 3 and 7
    expect = trials * walks / deg
    chi2 = ((live - expect) ** 2 / expect).sum()
    assert chi2 < 60, chi2                                                           # dof 20: p(>60) ~ 1e-6
    # the 6 walks of a seed are independent draws: P(two given walks agree) = 1/deg
    agree = (pos[:, 0] == pos[:, 1])[(pos[:, 0] >= 0) & (pos[:, 1] >= 0)].mean()
    assert abs(agree - 1 / (deg - 2)) < 0.01


def test_negative_sampler_is_uniform_over_the_far_train_nodes(g, dev):
    """K5b, src/models.py:163-164: `random.sample(far_nodes, num_neg)` -- a uniform num_neg-subset of the train nodes
    outside the seed's ball.  A path graph makes the ball known: with `hops` = 2 the ball of node 50 is 48..52."""
    n, hops, num_neg, trials = 101, 2, 6, 30_000
    src = np.arange(n - 1)
    from graphsage_b200 import synth
    rowptr, col = synth.edges_to_csr(n, src, src + 1)
    rp, cl = _csr_dev(rowptr, col, dev)
    train = np.arange(0, n, 2, dtype=np.int32)                                       # even nodes: 51 of them
    far = np.setdiff1d(train, np.arange(48, 53))                                     # minus 48, 50, 52 -> 48 far nodes
    seeds = torch.full((trials,), 50, dtype=torch.int32, device=dev)
    is_train = np.zeros(n, dtype=np.uint8)
    is_train[train] = 1
    for flags in (None, torch.from_numpy(is_train).to(dev)):      # the exact walk over the train list / rejection sampling
        neg, cnt = g.negative_sample(rp, cl, n, seeds, hops, num_neg, torch.from_numpy(train).to(dev), 7, 11, is_train=flags)
        neg = neg.cpu().numpy()
        assert np.all(cnt.cpu().numpy() == num_neg) and np.isin(neg, far).all()
        assert all(len(set(r)) == num_neg for r in neg[:2000])                       # without replacement
        counts = np.bincount(neg.ravel(), minlength=n)[far]
        expect = trials * num_neg / len(far)
        chi2 = ((counts - expect) ** 2 / (expect * (1 - num_neg / len(far)))).sum()
        assert chi2 < 110, (chi2, flags is None)                                     # dof 47: p(>110) ~ 1e-6
        pair = (neg[:, :1] < neg[:, 1:2]).mean() if flags is not None else 0.5       # acceptance order carries no bias
        assert abs(pair - 0.5) < 0.02
    # fewer far nodes than asked for: all of them, once each (:164)
    few, few_cnt = g.negative_sample(rp, cl, n, seeds[:4], 60, num_neg, torch.from_numpy(train).to(dev), 7, 12)
    assert np.all(few_cnt.cpu().numpy() == 0) and np.all(few.cpu().numpy() == -1)    # ball of 60 hops = the whole path


def test_sage_gemm_l2_normalize_epilogue(g, dev):
    """X1 (north_star item 3): the optional row-L2-normalise epilogue of K4, relu(X W^T) / max(||.||_2, 1e-12)
    (torch.nn.functional.normalize), forward only, off by default."""
    from graphsage_b200 import native
    rng = np.random.default_rng(12)
    rows, dim, H = 1000, 100, 128
    tab = torch.from_numpy(rng.standard_normal((3000, dim)).astype(np.float32)).to(dev)
    idx = torch.from_numpy(rng.integers(0, 3000, size=rows).astype(np.int32)).to(dev)
    agg = torch.from_numpy(rng.standard_normal((rows, dim)).astype(np.float32)).to(dev)
    w = torch.from_numpy((rng.standard_normal((H, 2 * dim)) * 0.1).astype(np.float32)).to(dev)
    w[:, :] = torch.where(torch.arange(H, device=dev)[:, None] == 5, torch.zeros_like(w), w)      # some dead columns
    out = g.sage_gemm_fwd(tab, idx, agg, dim, w, H, False, None, rows, True, native.PREC_TF32X3, l2_normalize=True)
    x = torch.cat([tab[idx.long()], agg], 1).double()
    want = torch.nn.functional.normalize(torch.relu(x @ w.double().t()), p=2, dim=1)
    assert rel(out, want) <= 1e-5
    assert torch.allclose(out.double().norm(dim=1), torch.ones(rows, dtype=torch.float64, device=dev), atol=1e-5)
    with pytest.raises(RuntimeError):
        g.sage_gemm_fwd(tab, idx, agg, dim, w, H, False, None, rows, True, native.PREC_FP32, l2_normalize=True)
