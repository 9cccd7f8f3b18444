"""Data formats either side of the hot path: the reference's text inputs -> CSR +
fp32 features, parsed once and kept as a binary cache.

The reference's `DataCenter.load_dataSet` (src/dataCenter.py:13-96) parses `cora.content` /
`cora.cites` or the Pubmed `.tab` files into Python lists and a `defaultdict(set)` adjacency on every run, and
`main.py:52` then copies the features to the device.  The hot path wants the graph as CSR (graph.py) and
never needs the dict-of-sets.  Here:

* `parse_cora`, `parse_pubmed` read the same files with the same node numbering (order of first appearance
  in the content file, dataCenter.py:26,66), label numbering (order of first appearance, :27-29; Pubmed
  `label=k` -> k-1, :67) and feature columns (:25; Pubmed word map :63-72), and build the CSR directly
  from the edge list (symmetric, duplicates and multi-edges collapsed exactly as the sets do, self
  citations kept as the reference keeps them);
* `DataSet.save` / `DataSet.load` keep it as `.npy` files in one directory, loaded with `mmap_mode='r'`, so a
  large graph is paged in by the H2D copy instead of being re-parsed;
* `DataCenter` mirrors the attribute surface the reference's loops read (`<ds>_train/_val/_test`,
  `<ds>_feats`, `<ds>_labels`, `<ds>_adj_lists`), with the same `np.random.permutation` split
  (dataCenter.py:98-111), and hands out an `AdjCSR` where the reference has the dict: `GraphSage` and
  `UnsupervisedLoss` accept it as `adj_lists`.

Pure host code (numpy); nothing here touches the GPU.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Tuple

import numpy as np

from .graph import AdjCSR

__all__ = ["DataSet", "DataCenter", "parse_cora", "parse_pubmed", "edges_to_csr", "split_data"]

_FORMAT_VERSION = 1


def edges_to_csr(src: np.ndarray, dst: np.ndarray, num_nodes: int) -> Tuple[np.ndarray, np.ndarray]:
    """Undirected edge list -> CSR with the semantics of `adj[a].add(b); adj[b].add(a)` (dataCenter.py:40-41):
    both directions, duplicates collapsed, rows ascending.  A self citation a-a yields the entry (a, a) once."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    if src.shape != dst.shape:
        raise ValueError("edge endpoints differ in length")
    if len(src) and (min(src.min(), dst.min()) < 0 or max(src.max(), dst.max()) >= num_nodes):
        raise ValueError("edge endpoint outside 0..num_nodes-1")
    a = np.concatenate([src, dst])
    b = np.concatenate([dst, src])
    key = np.unique(a * np.int64(num_nodes) + b)            # sorted by (row, col), duplicates removed
    rows = key // num_nodes
    col = (key - rows * num_nodes).astype(np.int32)
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=num_nodes), out=rowptr[1:])
    return rowptr, col


def split_data(num_nodes: int, test_split: int = 3, val_split: int = 6):
    """src/dataCenter.py:98-111, same draw from numpy's global stream: (test, val, train) index arrays."""
    order = np.random.permutation(num_nodes)
    n_test, n_val = num_nodes // test_split, num_nodes // val_split
    return np.split(order, [n_test, n_test + n_val])


class DataSet:
    """One parsed data set: CSR adjacency, fp32 features (the dtype `main.py:52` converts to), int64 labels."""

    def __init__(self, rowptr, col, feats, labels, label_names=None, meta: Optional[Dict] = None):
        self.rowptr = np.asarray(rowptr, dtype=np.int64)
        self.col = np.asarray(col, dtype=np.int32)
        self.feats = np.asarray(feats, dtype=np.float32)
        self.labels = np.asarray(labels, dtype=np.int64)
        self.label_names = list(label_names) if label_names is not None else None
        self.meta = dict(meta or {})
        n = len(self.rowptr) - 1
        if self.feats.shape[0] != n or self.labels.shape[0] != n:
            raise ValueError("features / labels / adjacency disagree on the number of nodes")

    @property
    def num_nodes(self) -> int:
        return len(self.rowptr) - 1

    def adjacency(self) -> AdjCSR:
        return AdjCSR(self.rowptr, self.col)

    # ---- binary cache --------------------------------------------------------------------------
    def save(self, path: str) -> str:
        os.makedirs(path, exist_ok=True)
        for name in ("rowptr", "col", "feats", "labels"):
            np.save(os.path.join(path, name + ".npy"), np.ascontiguousarray(getattr(self, name)))
        with open(os.path.join(path, "meta.json"), "w") as fp:
            json.dump({"format": _FORMAT_VERSION, "num_nodes": self.num_nodes, "nnz": int(len(self.col)),
                       "num_feats": int(self.feats.shape[1]), "label_names": self.label_names, "meta": self.meta}, fp)
        return path

    @classmethod
    def load(cls, path: str, mmap: bool = True) -> "DataSet":
        with open(os.path.join(path, "meta.json")) as fp:
            info = json.load(fp)
        if info.get("format") != _FORMAT_VERSION:
            raise ValueError(f"{path}: cache format {info.get('format')} != {_FORMAT_VERSION}; parse again")
        mode = "r" if mmap else None
        arrs = {n: np.load(os.path.join(path, n + ".npy"), mmap_mode=mode) for n in ("rowptr", "col", "feats", "labels")}
        self = cls.__new__(cls)
        self.rowptr, self.col, self.feats, self.labels = arrs["rowptr"], arrs["col"], arrs["feats"], arrs["labels"]
        self.label_names, self.meta = info.get("label_names"), info.get("meta", {})
        if (len(self.rowptr) - 1 != info["num_nodes"] or len(self.col) != info["nnz"] or
                self.feats.shape != (info["num_nodes"], info["num_feats"]) or len(self.labels) != info["num_nodes"] or
                self.rowptr.dtype != np.int64 or self.col.dtype != np.int32 or self.feats.dtype != np.float32):
            raise ValueError(f"{path}: cache files do not match meta.json; parse again")
        return self


def _intern(table: Dict[str, int], key: str) -> int:
    """Dense id of `key` in order of first appearance."""
    idx = table.get(key)
    if idx is None:
        idx = table[key] = len(table)
    return idx


def parse_cora(content_file: str, cite_file: str) -> DataSet:
    """The Cora pair of files (read by src/dataCenter.py:14-52): whitespace-separated
    `<paper> <f0> ... <fF-1> <class name>` per paper, `<paper> <paper>` per citation.  Papers are numbered by line
    (:26), classes by first appearance (:27-29)."""
    paper_id: Dict[str, int] = {}
    class_id: Dict[str, int] = {}
    rows, classes = [], []
    with open(content_file) as fp:
        for line in fp:
            tokens = line.split()
            if not tokens:
                continue
            name, values, cls_name = tokens[0], tokens[1:-1], tokens[-1]
            paper_id[name] = len(rows)                       # a repeated paper keeps its LAST line number, as a dict does
            rows.append(np.array(values, dtype=np.float64))
            classes.append(_intern(class_id, cls_name))
    n = len(rows)
    ends = []
    with open(cite_file) as fp:
        for line in fp:
            pair = line.split()
            if len(pair) != 2:
                raise ValueError(f"{cite_file}: expected two paper ids per line")     # the reference asserts (:37)
            ends.append((paper_id[pair[0]], paper_id[pair[1]]))                       # KeyError for an unknown paper
    ends = np.asarray(ends, dtype=np.int64).reshape(-1, 2)
    rowptr, col = edges_to_csr(ends[:, 0], ends[:, 1], n)
    _require_no_isolated(rowptr, n)
    names = sorted(class_id, key=class_id.get)
    return DataSet(rowptr, col, np.stack(rows) if rows else np.zeros((0, 0)), classes, names, {"source": "cora"})


def parse_pubmed(paper_file: str, cites_file: str) -> DataSet:
    """The Pubmed-Diabetes pair of `.tab` files (read by src/dataCenter.py:54-96).  Paper file: a title line, a schema
    line of tab-separated `<type>:<name>:<default>` columns (first the label, last the summary), then per paper
    `<id>\tlabel=<k>\t<word>=<tfidf>...\tsummary=...`; the feature columns are the schema's words in schema order
    (:63,:69-72) and the class is k-1 (:67).  Cites file: two header lines, then `<n>\tpaper:<a>\t|\tpaper:<b>`."""
    paper_id: Dict[str, int] = {}
    rows, classes = [], []
    with open(paper_file) as fp:
        fp.readline()
        schema = [field.split(":")[1] for field in fp.readline().split("\t")]
        column = {word: k for k, word in enumerate(schema[1:-1])}          # words only: label first, summary last
        for line in fp:
            fields = line.split("\t")
            paper_id[fields[0]] = len(rows)
            classes.append(int(fields[1].partition("=")[2]) - 1)
            vec = np.zeros(len(column))
            for item in fields[2:-1]:                                      # the last field is the summary
                word, _, value = item.partition("=")
                vec[column[word]] = float(value)
            rows.append(vec)
    n = len(rows)
    ends = []
    with open(cites_file) as fp:
        fp.readline()
        fp.readline()
        for line in fp:
            fields = line.strip().split("\t")
            ends.append((paper_id[fields[1].partition(":")[2]], paper_id[fields[-1].partition(":")[2]]))   # :84-85
    ends = np.asarray(ends, dtype=np.int64).reshape(-1, 2)
    rowptr, col = edges_to_csr(ends[:, 0], ends[:, 1], n)
    _require_no_isolated(rowptr, n)
    return DataSet(rowptr, col, np.stack(rows) if rows else np.zeros((0, 0)), classes, None, {"source": "pubmed"})


def _require_no_isolated(rowptr: np.ndarray, n: int) -> None:
    # the reference asserts len(feat_data) == len(labels) == len(adj_lists) (:43,:88): every node has an edge
    if n and int((np.diff(rowptr) == 0).sum()):
        raise AssertionError("a node of the content file has no citation (the reference asserts the same, "
                             "src/dataCenter.py:43,88)")


class DataCenter:
    """Attribute-compatible stand-in for the reference's DataCenter (src/dataCenter.py:8-96) on top of a
    DataSet: `load_dataSet(ds)` parses (or loads the cache next to the files) and sets `<ds>_test`,
    `<ds>_val`, `<ds>_train` (np.int64 index arrays from `np.random.permutation`, :98-111), `<ds>_feats`
    (fp32 here; the reference holds float64 and converts at main.py:52), `<ds>_labels` and `<ds>_adj_lists`
    (an AdjCSR: read like the dict, and taken as is by the device path).

    `config` is any mapping with the reference's keys (`file_path.cora_content`, `file_path.cora_cite`,
    `file_path.pubmed_paper`, `file_path.pubmed_cites`; experiments.conf:1-8).  `cache_dir` (optional) is where
    the binary form lives; it is rebuilt when a source file is newer than the cache."""

    def __init__(self, config, cache_dir: Optional[str] = None):
        self.config = config
        self.cache_dir = cache_dir

    def _sources(self, dataSet: str):
        if dataSet == 'cora':
            return parse_cora, (self.config['file_path.cora_content'], self.config['file_path.cora_cite'])
        if dataSet == 'pubmed':
            return parse_pubmed, (self.config['file_path.pubmed_paper'], self.config['file_path.pubmed_cites'])
        raise ValueError(f"unknown data set {dataSet!r} (the reference knows 'cora' and 'pubmed')")

    def load_dataSet(self, dataSet: str = 'cora') -> DataSet:
        parse, files = self._sources(dataSet)
        data = None
        cache = os.path.join(self.cache_dir, dataSet) if self.cache_dir else None
        if cache and os.path.isfile(os.path.join(cache, "meta.json")):
            newest = max(os.path.getmtime(f) for f in files)
            if os.path.getmtime(os.path.join(cache, "meta.json")) >= newest:
                try:
                    data = DataSet.load(cache)
                except (ValueError, OSError):
                    data = None
        if data is None:
            data = parse(*files)
            if cache:
                data.save(cache)
        test, val, train = split_data(data.num_nodes)
        published = {"test": test, "val": val, "train": train, "feats": data.feats, "labels": data.labels,
                     "adj_lists": data.adjacency()}
        for suffix, value in published.items():          # the attribute names the reference's loops read
            setattr(self, f"{dataSet}_{suffix}", value)
        return data
