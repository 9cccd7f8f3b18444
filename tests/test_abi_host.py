"""CPU: the C-ABI library builds, loads and exports every symbol include/gsage_b200.h declares
(no compute calls without a GPU); host-side logic (adjacency conversion, synthetic graphs,
loud failure without CUDA); the product never imports the oracle."""
import os
import re
from collections import defaultdict

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'gsage_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gs_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import graphsage_b200  # noqa: F401
    from graphsage_b200 import native
    lib = native.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/gsage_b200.h but not exported'
    assert sorted(native.exported_symbols()) == names          # ctypes table mirrors the header one to one
    assert lib.gs_version() == native.ABI_VERSION
    assert b'bad argument' in lib.gs_error_string(-1)
    assert os.path.dirname(native.lib_path()) == os.path.join(ROOT, 'graphsage-pytorch_b200')   # in-tree .so


def test_bad_arguments_are_rejected_before_any_launch():
    from graphsage_b200 import native
    lib = native.load()
    assert lib.gs_sample_neighbors(None, None, 0, None, None, 4, 10, 10, 0, 1, 1, None, None, None, None) == -1
    assert lib.gs_agg_fwd(None, 0, 0, None, 0, None, None, 0, 0, None, 0, None, 0, None) == -1
    assert lib.gs_unique_workspace_bytes(1024, 11) == 16                      # single-CTA path needs no scratch
    assert lib.gs_unique_workspace_bytes(90112, 11) > 8 * 90112 * 12


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'graphsage-pytorch_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f


def test_adjacency_conversion_and_mapping_semantics():
    from graphsage_b200.graph import AdjCSR, adj_to_csr
    adj = defaultdict(set)
    edges = [(0, 3), (3, 1), (1, 1), (4, 0), (2, 4)]
    for a, b in edges:
        adj[a].add(b)
        adj[b].add(a)
    rowptr, col = adj_to_csr(adj, 6)                 # node 5 has no key: empty row (defaultdict semantics)
    assert rowptr.tolist() == [0, 2, 4, 5, 7, 9, 9]
    assert col.tolist() == [3, 4, 1, 3, 4, 0, 1, 0, 2]
    view = AdjCSR(rowptr, col)
    for v in range(6):
        assert view[v] == adj.get(v, set())
    assert view[17] == set() and len(view) == 6
    with pytest.raises(ValueError):
        adj_to_csr({0: {9}}, 3)


def test_synthetic_graph_is_a_valid_dict_of_sets_image():
    from graphsage_b200 import synth
    rowptr, col = synth.powerlaw_graph(5000, 60000, seed=0)
    n = len(rowptr) - 1
    row = np.repeat(np.arange(n), np.diff(rowptr))
    assert np.all(np.diff(rowptr) >= 2) and not np.any(row == col)
    key = row.astype(np.int64) * n + col
    assert np.all(np.diff(key) > 0)                                   # sorted rows, no duplicates
    rev = col.astype(np.int64) * n + row
    assert np.array_equal(np.sort(rev), key)                          # both directions present
    deg = np.diff(rowptr)
    assert deg.max() > 20 * np.median(deg)                            # heavy tail


def test_cpu_tensors_fail_loudly():
    from graphsage_b200 import models
    from graphsage_b200.graph import AdjCSR
    adj = AdjCSR(np.array([0, 1, 2]), np.array([1, 0]))
    m = models.GraphSage(2, 8, 4, torch.zeros(2, 8), adj, torch.device('cpu'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m([0, 1])
    layer = models.SageLayer(8, 4)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        layer(torch.zeros(2, 8), torch.zeros(2, 8))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        models.Classification(4, 3)(torch.zeros(2, 4))
    assert list(m.state_dict()) == ['sage_layer1.weight', 'sage_layer2.weight']
    assert m.sage_layer1.weight.shape == (4, 16) and m.sage_layer2.weight.shape == (4, 8)


def test_philox_host_matches_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors): the sampler's generator."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85

    def philox(ctr, key):
        c, k = list(ctr), list(key)
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
            k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
        return c

    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_micro_f1_is_sklearn_micro_f1_for_single_label_predictions():
    """inference.micro_f1 restates f1_score(average='micro') (src/utils.py:32,45) without leaving the device."""
    import numpy as np
    import torch
    from sklearn.metrics import f1_score
    from graphsage_b200 import inference
    rng = np.random.default_rng(3)
    for classes, n in ((7, 500), (3, 41), (47, 2000)):
        y = rng.integers(0, classes, size=n)
        p = np.where(rng.random(n) < 0.6, y, rng.integers(0, classes, size=n))
        assert abs(inference.micro_f1(torch.from_numpy(y), torch.from_numpy(p)) - f1_score(y, p, average='micro')) < 1e-12
    import pytest
    with pytest.raises(ValueError):
        inference.micro_f1(torch.zeros(3, dtype=torch.int64), torch.zeros(4, dtype=torch.int64))


def test_negative_radius_keeps_the_reference_ball_on_sparse_graphs_only():
    """UnsupervisedLoss.negative_hops: 5 hops (src/models.py:52,155-162) on the reference's data sets, fewer where
    that ball would be the whole graph (BASELINE configs[2], [3]: the reference's far set is empty there)."""
    from graphsage_b200.models import negative_radius
    assert negative_radius(10_556, 2_708, 5, 10_000) == 5              # Cora
    assert negative_radius(88_651, 19_717, 5, 10_000) == 5             # Pubmed
    assert negative_radius(123_714_814, 2_449_029, 5, 10_000) == 2     # cfg-3
    assert negative_radius(114_600_000, 232_965, 5, 10_000) == 1       # cfg-4
    assert negative_radius(0, 100, 5, 10_000) == 5 and negative_radius(10 ** 9, 10, 5, 10_000) == 1


def test_reference_arm_prints_the_contract_line(tmp_path):
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the same metric /
    unit / config keys as our arm, impl == "reference", a cpu_baseline describing the run and an e2e object with
    zero copy bytes.  Ranks other than 0 print nothing and exit 0."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--scale", "0.003", "--b_sz", "128"]
    env = dict(os.environ, GSAGE_CACHE=str(tmp_path), OMP_NUM_THREADS="1")       # what torchrun exports
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "seed_nodes_per_sec_fwd_bwd" and d["unit"] == "seed nodes/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    staged = os.path.isfile(os.path.join(root, "baseline", "_ref", "src", "models.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")      # the reference itself when it is staged
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)                   # every host thread, also under torchrun's OMP_NUM_THREADS=1
    assert d["cpu_baseline"]["value"] == d["value"] and d["config"]["b_sz_per_gpu"] == 128      # the batch is never shrunk
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("cfg3_products") and d["config"]["fanout"] == 10
    other = subprocess.run(cmd, capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=60)
    assert other.returncode == 0 and not [l for l in other.stdout.splitlines() if l.startswith("{")]


def test_ctypes_signatures_have_the_arity_the_header_declares():
    """Every entry of native._SIGNATURES must list exactly as many arguments as the declaration in include/gsage_b200.h
    (a missing entry shifts every later argument: ctypes would pass a pointer where the library reads a size)."""
    import os
    import re
    import graphsage_b200  # noqa: F401
    from graphsage_b200 import native
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r'/\*.*?\*/', '', open(os.path.join(root, 'include', 'gsage_b200.h')).read(), flags=re.S)
    bad = []
    for name, (_, args) in native._SIGNATURES.items():
        m = re.search(r'\b' + name + r'\s*\(([^;]*?)\)\s*;', hdr, flags=re.S)
        assert m, f'{name} is bound but not declared in the header'
        params = m.group(1).strip()
        n = 0 if params in ('', 'void') else len(params.split(','))
        if n != len(args):
            bad.append((name, n, len(args)))
    assert not bad, bad
