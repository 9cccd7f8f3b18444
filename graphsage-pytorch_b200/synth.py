"""Synthetic workloads of the shapes BASELINE.json names (data plumbing for tests and bench).

The reference ships edge lists only (its feature/label blobs are missing), so
every config here is "topology (shipped or generated) + synthetic features/labels of the
documented shape".  Everything is drawn from `numpy.random.default_rng(seed)` (PCG64, stream
stable across numpy versions) so that the GPU arm, the CPU oracle arm and the committed
golden fixtures all see identical inputs.

Graphs are returned as CSR (`rowptr` int64 [N+1], `col` int32 [nnz]); rows are sorted
ascending, have no duplicates (a dict-of-sets cannot hold any, src/dataCenter.py:33-41) and
both directions of every edge are present (src/dataCenter.py:40-41).
"""
from __future__ import annotations

import hashlib
import os
from typing import Tuple

import numpy as np


def edges_to_csr(n: int, src: np.ndarray, dst: np.ndarray, keep_self_loops: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Symmetrise + dedupe an edge list into CSR, i.e. what `adj[a].add(b); adj[b].add(a)`
    builds at src/dataCenter.py:40-41 / :84-85."""
    a = np.concatenate([src, dst]).astype(np.int64)
    b = np.concatenate([dst, src]).astype(np.int64)
    if len(a) < (1 << 20):
        if not keep_self_loops:
            keep = a != b
            a, b = a[keep], b[keep]
        key = np.unique(a * np.int64(n) + b)
        row = key // n
        col = (key - row * n).astype(np.int32)
        rowptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(row, minlength=n), out=rowptr[1:])
        return rowptr, col
    # large graphs: torch's multi-threaded sort is ~20x numpy's np.unique at 1e8 keys; integer
    # sort + unique is exact, so both routes give the same CSR
    import torch
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    key = ta * n + tb
    if not keep_self_loops:
        key = key[ta != tb]
    key = torch.unique_consecutive(torch.sort(key)[0])
    row = torch.div(key, n, rounding_mode='floor')
    col = (key - row * n).to(torch.int32)
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    return rowptr.numpy(), col.numpy()


def powerlaw_graph(n: int, num_edges: int, seed: int = 0, cache_dir: str | None = None) -> Tuple[np.ndarray, np.ndarray]:
    """Chung-Lu style generator of SURVEY.md §8(d): one endpoint ~ rank^-0.5, the other
    uniform, plus a ring so every node has degree >= 2.  `num_edges` counts undirected
    edges before de-duplication (ring included)."""
    tag = f"pl_{n}_{num_edges}_{seed}"
    if cache_dir:
        path = os.path.join(cache_dir, tag + ".npz")
        if os.path.exists(path):
            z = np.load(path)
            return z["rowptr"], z["col"]
    rng = np.random.default_rng(seed)
    m = max(0, num_edges - n)
    u = rng.random(m)
    src = np.minimum((u * u * n).astype(np.int64), n - 1)       # P(i) ~ (i+1)^-0.5
    dst = rng.integers(0, n, size=m, dtype=np.int64)
    ring = np.arange(n, dtype=np.int64)
    src = np.concatenate([src, ring])
    dst = np.concatenate([dst, (ring + 1) % n])
    rowptr, col = edges_to_csr(n, src, dst, keep_self_loops=False)
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, rowptr=rowptr, col=col)
        os.replace(tmp, path)
    return rowptr, col


def random_graph(n: int, avg_deg: float, seed: int = 0, isolated: int = 0, self_loops: int = 0):
    """Small Erdos-Renyi-ish graph for unit tests, with optional isolated nodes (the
    last `isolated` ids have no edges) and a few self-loops (Pubmed has 3)."""
    rng = np.random.default_rng(seed)
    live = n - isolated
    m = int(n * avg_deg / 2)
    src = rng.integers(0, live, size=m)
    dst = rng.integers(0, live, size=m)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    if self_loops:
        loops = rng.choice(live, size=self_loops, replace=False)
        src = np.concatenate([src, loops])
        dst = np.concatenate([dst, loops])
    return edges_to_csr(n, src, dst, keep_self_loops=True)


def features_binary(n: int, f: int, density: float, seed: int) -> np.ndarray:
    """Cora-like bag-of-words rows (cora/README:3-14): 0/1 with the given density."""
    rng = np.random.default_rng(seed)
    return (rng.random((n, f), dtype=np.float32) < density).astype(np.float32)


def features_sparse_float(n: int, f: int, density: float, seed: int) -> np.ndarray:
    """Pubmed-like TF-IDF rows (src/dataCenter.py:69-72): non-negative, mostly zero."""
    rng = np.random.default_rng(seed)
    val = rng.random((n, f), dtype=np.float32)
    keep = rng.random((n, f), dtype=np.float32) < density
    return (val * keep).astype(np.float32)


def features_normal(n: int, f: int, seed: int, dtype=np.float32, chunk: int = 1 << 18) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = np.empty((n, f), dtype=dtype)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        out[s:e] = rng.standard_normal((e - s, f), dtype=np.float32).astype(dtype)
    return out


def labels_uniform(n: int, num_classes: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, num_classes, size=n, dtype=np.int64)


def split_nodes(n: int, seed: int, test_split: int = 3, val_split: int = 6):
    """src/dataCenter.py:100-111 with an explicit Generator instead of numpy's global RNG."""
    perm = np.random.default_rng(seed).permutation(n).astype(np.int64)
    t, v = n // test_split, n // val_split
    return perm[:t], perm[t:t + v], perm[t + v:]


def xavier_uniform_np(rng: np.random.Generator, out_size: int, in_size: int) -> np.ndarray:
    bound = float(np.sqrt(6.0 / (in_size + out_size)))
    return rng.uniform(-bound, bound, size=(out_size, in_size)).astype(np.float32)


def digest(arr: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()[:16]


# The named configs of BASELINE.json (`configs[i]`), as shapes.
CONFIGS = {
    "cfg1_cora": dict(n=2708, feats=1433, classes=7, hidden=128, agg="MEAN", gcn=False, learn="sup", unsup_loss="normal", b_sz=20),
    "cfg2_pubmed": dict(n=19717, feats=500, classes=3, hidden=128, agg="MAX", gcn=False, learn="unsup", unsup_loss="normal", b_sz=20),
    "cfg3_products": dict(n=2_449_029, edges=61_859_140, feats=100, classes=47, hidden=128, agg="MEAN", gcn=False, learn="sup", b_sz=1024),
    "cfg4_reddit": dict(n=232_965, edges=57_307_946, feats=602, classes=41, hidden=128, agg="MEAN", gcn=True, learn="plus_unsup", unsup_loss="margin", b_sz=1024),
    "cfg5_100m": dict(n=100_000_000, edges=800_000_000, feats=128, classes=47, hidden=128, agg="MEAN", gcn=False, learn="sup", b_sz=8192),
}


def device_powerlaw_csr(n: int, mean_degree: float, device, seed: int = 0, max_degree: int = 10000):
    """Adjacency of BASELINE.json configs[4] generated directly in HBM (1.6B entries at 100M
    nodes never exist on the host): Pareto(alpha=2.5) out-degrees with the requested mean,
    uniform random targets.  torch's Philox generator makes the result identical on every rank
    that uses the same seed, so the replicas of a data-parallel job hold the same graph.
    Returns (rowptr int64 [n+1], col int32 [nnz]) on `device`."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    alpha = 2.5
    d_min = mean_degree * (alpha - 1.0) / alpha
    u = torch.rand((n,), generator=g, device=device).clamp_min_(1e-9)
    deg = torch.floor(d_min * u.pow_(-1.0 / alpha)).clamp_(1, max_degree).to(torch.int64)
    del u
    rowptr = torch.zeros((n + 1,), dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=rowptr[1:])
    del deg
    nnz = int(rowptr[-1].item())
    col = torch.empty((nnz,), dtype=torch.int32, device=device)
    step = 1 << 28
    for lo in range(0, nnz, step):                 # bounded temporaries
        hi = min(nnz, lo + step)
        col[lo:hi] = torch.randint(0, n, (hi - lo,), generator=g, device=device, dtype=torch.int32)
    return rowptr, col
