"""Harness that imports and drives the *actual* reference  --  TEST INFRASTRUCTURE ONLY.

Only usable in the build container, where `/root/reference` exists (it does not exist on
the GPU box, and nothing in `-m gpu` tests, `smoke()` or `bench.py` touches this file).
`tests/golden/make_golden.py` uses it to (1) pin `oracle/sage_oracle.py` against the
reference itself and (2) write the golden fixtures that travel with the repo.

Nothing from the reference is copied: its modules are imported from where they lie.
"""
from __future__ import annotations

import importlib
import os
import random
import sys
from collections import defaultdict
from typing import List, Tuple

import numpy as np

REFERENCE_ROOT = os.environ.get("GSAGE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models.py"))


def install_random_sample_shim() -> None:
    """`random.sample(set, k)` raises TypeError on Python >= 3.11, which the reference hits at
    src/models.py:282 and :164.  CPython <= 3.10 (the interpreter the reference targets,
    README.md:15) converted the set with tuple() internally; do the same."""
    if getattr(random.sample, "_gsage_shim", False):
        return
    original = random.sample

    def sample(population, k, **kw):
        if isinstance(population, (set, frozenset)):
            population = tuple(population)
        return original(population, k, **kw)

    sample._gsage_shim = True
    random.sample = sample


def load_reference_models():
    """Import the reference's `src.models` from /root/reference (read-only)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}; this harness only runs in the build container")
    install_random_sample_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("src.models")


def load_reference_utils():
    load_reference_models()
    return importlib.import_module("src.utils")


class SampleRecorder:
    """Record the (nodes, samp_neighs, unique_list) triple of every
    `GraphSage._get_unique_neighs_list` call (src/models.py:250) - the injection seam."""

    def __init__(self, model):
        self.model = model
        self.calls: List[Tuple[list, list, list]] = []
        self._orig = model._get_unique_neighs_list

    def __enter__(self):
        def wrapped(nodes, num_sample=10):
            samp, index_of, uniq = self._orig(nodes, num_sample)
            self.calls.append((list(nodes), [set(s) for s in samp], list(uniq)))
            return samp, index_of, uniq
        self.model._get_unique_neighs_list = wrapped
        return self

    def __exit__(self, *exc):
        self.model._get_unique_neighs_list = self._orig
        return False


def _first_appearance_ids(pairs):
    ids = {}
    for a, b in pairs:
        for x in (a, b):
            if x not in ids:
                ids[x] = len(ids)
    return ids


def cora_topology():
    """Parse cora/cora.cites as src/dataCenter.py:34-41 does; node ids by first appearance
    (the content file that fixes the reference's numbering is missing)."""
    pairs = []
    with open(os.path.join(REFERENCE_ROOT, "cora", "cora.cites")) as fp:
        for line in fp:
            info = line.strip().split()
            assert len(info) == 2
            pairs.append((info[0], info[1]))
    return _pairs_to_csr(pairs)


def pubmed_topology():
    """Parse Pubmed-Diabetes.DIRECTED.cites.tab as src/dataCenter.py:78-86 does."""
    pairs = []
    with open(os.path.join(REFERENCE_ROOT, "pubmed-data", "Pubmed-Diabetes.DIRECTED.cites.tab")) as fp:
        fp.readline()
        fp.readline()
        for line in fp:
            info = line.strip().split("\t")
            pairs.append((info[1].split(":")[1], info[-1].split(":")[1]))
    return _pairs_to_csr(pairs)


def _pairs_to_csr(pairs):
    ids = _first_appearance_ids(pairs)
    adj = defaultdict(set)
    for a, b in pairs:
        adj[ids[a]].add(ids[b])
        adj[ids[b]].add(ids[a])
    n = len(ids)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    cols = []
    for v in range(n):
        row = sorted(adj[v])
        cols.extend(row)
        rowptr[v + 1] = rowptr[v] + len(row)
    return rowptr, np.asarray(cols, dtype=np.int32)
