"""CPU: graphsage_b200.datacache against what the reference's own DataCenter produced on the
committed synthetic inputs (tests/golden/datacenter/, written by tests/golden/make_golden_datacenter.py).
Bar: bit-exact ids, labels, adjacency and splits; features equal to the reference's float64 values converted to
fp32 (the conversion main.py:52 does)."""
import os
import shutil

import numpy as np
import pytest

from graphsage_b200 import datacache
from graphsage_b200.graph import AdjCSR

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "datacenter")
SEED = 824


def config_for(root):
    return {"file_path.cora_content": os.path.join(root, "cora.content"),
            "file_path.cora_cite": os.path.join(root, "cora.cites"),
            "file_path.pubmed_paper": os.path.join(root, "Pubmed-Diabetes.NODE.paper.tab"),
            "file_path.pubmed_cites": os.path.join(root, "Pubmed-Diabetes.DIRECTED.cites.tab")}


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "reference_outputs.npz"))


@pytest.mark.parametrize("ds", ["cora", "pubmed"])
def test_parser_matches_the_reference_datacenter(ref, ds):
    np.random.seed(SEED)
    dc = datacache.DataCenter(config_for(GOLD))
    data = dc.load_dataSet(ds)
    feats, labels, adj = getattr(dc, ds + "_feats"), getattr(dc, ds + "_labels"), getattr(dc, ds + "_adj_lists")
    assert feats.dtype == np.float32 and np.array_equal(feats, ref[ds + "_feats64"].astype(np.float32))
    assert labels.dtype == np.int64 and np.array_equal(labels, ref[ds + "_labels"])
    assert isinstance(adj, AdjCSR)
    assert np.array_equal(adj.rowptr, ref[ds + "_rowptr"]) and np.array_equal(adj.col, ref[ds + "_col"])
    for part in ("test", "val", "train"):                       # same draw from numpy's global stream (:98-111)
        assert np.array_equal(getattr(dc, f"{ds}_{part}"), ref[f"{ds}_{part}"])
    n = data.num_nodes
    assert len(dc.__dict__[ds + "_test"]) == n // 3 and len(dc.__dict__[ds + "_val"]) == n // 6
    # dict-of-sets read semantics of the adjacency (src/models.py:279): adj[v] is the set the reference holds
    v = int(np.argmax(np.diff(adj.rowptr)))
    assert adj[v] == set(ref[ds + "_col"][ref[ds + "_rowptr"][v]:ref[ds + "_rowptr"][v + 1]].tolist())
    assert adj[n + 5] == set()


def test_self_citations_and_duplicates_follow_set_semantics(ref):
    data = datacache.parse_cora(*[config_for(GOLD)[k] for k in ("file_path.cora_content", "file_path.cora_cite")])
    assert (data.col[data.rowptr[5]:data.rowptr[6]] == 5).sum() == 1            # the a-a line gives one (a, a) entry
    for v in range(data.num_nodes):                                               # ascending, no duplicates
        row = data.col[data.rowptr[v]:data.rowptr[v + 1]]
        assert np.all(np.diff(row) > 0)
    rowptr, col = datacache.edges_to_csr(np.array([0, 1, 1, 2]), np.array([1, 0, 1, 0]), 4)
    assert rowptr.tolist() == [0, 2, 4, 5, 5] and col.tolist() == [1, 2, 0, 1, 0]
    with pytest.raises(ValueError):
        datacache.edges_to_csr(np.array([0]), np.array([4]), 4)


@pytest.mark.parametrize("ds", ["cora", "pubmed"])
def test_binary_cache_round_trip_and_reuse(tmp_path, ref, ds):
    src = tmp_path / "src"
    shutil.copytree(GOLD, src)
    cache = tmp_path / "cache"
    np.random.seed(SEED)
    first = datacache.DataCenter(config_for(str(src)), cache_dir=str(cache))
    first.load_dataSet(ds)
    assert os.path.isfile(cache / ds / "meta.json")
    # second load must come from the cache: make the text files unparsable, keep their mtime older than the cache
    for name in os.listdir(src):
        p = src / name
        mtime = os.path.getmtime(p)
        p.write_text("garbage\n")
        os.utime(p, (mtime, mtime))
    np.random.seed(SEED)
    again = datacache.DataCenter(config_for(str(src)), cache_dir=str(cache))
    data = again.load_dataSet(ds)
    assert isinstance(data.feats, np.memmap) and isinstance(data.col, np.memmap)
    for key in ("_feats", "_labels", "_test", "_val", "_train"):
        assert np.array_equal(getattr(first, ds + key), getattr(again, ds + key))
    assert np.array_equal(getattr(again, ds + "_adj_lists").col, ref[ds + "_col"])
    # a newer source file invalidates the cache (and the garbage then fails to parse)
    newer = os.path.getmtime(cache / ds / "meta.json") + 10
    for name in os.listdir(src):
        os.utime(src / name, (newer, newer))
    with pytest.raises(Exception):
        datacache.DataCenter(config_for(str(src)), cache_dir=str(cache)).load_dataSet(ds)
    # a damaged cache is detected, not trusted
    np.save(cache / ds / "col.npy", np.zeros(3, dtype=np.int32))
    with pytest.raises(ValueError):
        datacache.DataSet.load(str(cache / ds))


def test_isolated_node_is_rejected_like_the_reference(tmp_path):
    (tmp_path / "c.content").write_text("a 1 0 X\nb 0 1 Y\nc 1 1 X\n")
    (tmp_path / "c.cites").write_text("a b\n")
    with pytest.raises(AssertionError):                          # dataCenter.py:43: len(feat_data) == len(adj_lists)
        datacache.parse_cora(str(tmp_path / "c.content"), str(tmp_path / "c.cites"))
    (tmp_path / "c.cites").write_text("a b\nb c\n")
    d = datacache.parse_cora(str(tmp_path / "c.content"), str(tmp_path / "c.cites"))
    assert d.labels.tolist() == [0, 1, 0] and d.label_names == ["X", "Y"] and d.col.tolist() == [1, 0, 2, 1]
    with pytest.raises(ValueError):
        datacache.DataCenter({}).load_dataSet("reddit")
