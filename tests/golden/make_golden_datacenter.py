"""Golden vectors for graphsage_b200.datacache  --  run in the BUILD container only.

Writes small synthetic inputs in the reference's two text formats under tests/golden/datacenter/ and runs the
REFERENCE's own DataCenter.load_dataSet (imported from /root/reference/src/dataCenter.py, nothing copied) on
them with numpy's global stream seeded like main.py:41; what it produced is stored in
tests/golden/datacenter/reference_outputs.npz.  tests/test_datacache.py compares our parser with it.

    python tests/golden/make_golden_datacenter.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "datacenter")
REF = os.environ.get("GSAGE_REFERENCE_ROOT", "/root/reference")
SEED = 824          # main.py:18,41


def write_cora(rng, n=60, feats=30, edges=170):
    ids = rng.choice(np.arange(10_000, 9_999_999), size=n, replace=False)
    label_names = ["Neural_Networks", "Theory", "Case_Based", "Genetic_Algorithms"]
    with open(os.path.join(OUT, "cora.content"), "w") as fp:
        for i in range(n):
            bits = (rng.random(feats) < 0.15).astype(int)
            fp.write("\t".join([str(ids[i])] + [str(b) for b in bits] + [label_names[rng.integers(0, 4)]]) + "\n")
    pairs = [(ids[i], ids[(i + 1) % n]) for i in range(n)]                    # ring: no isolated paper (:43)
    pairs += [(ids[a], ids[b]) for a, b in rng.integers(0, n, size=(edges, 2))]
    pairs += [pairs[3], pairs[7][::-1], (ids[5], ids[5])]                      # duplicate, reversed duplicate, self citation
    with open(os.path.join(OUT, "cora.cites"), "w") as fp:
        for a, b in pairs:
            fp.write(f"{a}\t{b}\n")


def write_pubmed(rng, n=80, words=25, edges=200):
    ids = rng.choice(np.arange(100_000, 99_999_999), size=n, replace=False)
    names = [f"w-{k:03d}" for k in range(words)]
    with open(os.path.join(OUT, "Pubmed-Diabetes.NODE.paper.tab"), "w") as fp:
        fp.write("NODE\tpaper\n")
        fp.write("\t".join(["cat=1,2,3:label"] + [f"numeric:{w}:0.0" for w in names] + ["string:summary"]) + "\n")
        for i in range(n):
            k = rng.integers(1, 7)
            chosen = rng.choice(words, size=k, replace=False)
            fields = [f"{names[w]}={rng.random() * 0.2:.6f}" for w in chosen]
            fp.write("\t".join([str(ids[i]), f"label={rng.integers(1, 4)}"] + fields +
                               ["summary=" + ",".join(names[w] for w in chosen)]) + "\n")
    pairs = [(ids[i], ids[(i + 1) % n]) for i in range(n)]
    pairs += [(ids[a], ids[b]) for a, b in rng.integers(0, n, size=(edges, 2))]
    pairs += [pairs[1], pairs[2][::-1], (ids[9], ids[9]), (ids[11], ids[11])]
    with open(os.path.join(OUT, "Pubmed-Diabetes.DIRECTED.cites.tab"), "w") as fp:
        fp.write("DIRECTED\tcites\n")
        fp.write("NO_FEATURES\n")
        for j, (a, b) in enumerate(pairs):
            fp.write(f"{j}\tpaper:{a}\t|\tpaper:{b}\n")


def csr_of(adj, n):
    rowptr = np.zeros(n + 1, dtype=np.int64)
    cols = []
    for v in range(n):
        row = sorted(int(x) for x in adj[v])
        cols.extend(row)
        rowptr[v + 1] = len(cols)
    return rowptr, np.asarray(cols, dtype=np.int32)


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(5)
    write_cora(rng)
    write_pubmed(rng)
    sys.path.insert(0, REF)
    from src.dataCenter import DataCenter          # the reference, where it lies
    config = {"file_path.cora_content": os.path.join(OUT, "cora.content"),
              "file_path.cora_cite": os.path.join(OUT, "cora.cites"),
              "file_path.pubmed_paper": os.path.join(OUT, "Pubmed-Diabetes.NODE.paper.tab"),
              "file_path.pubmed_cites": os.path.join(OUT, "Pubmed-Diabetes.DIRECTED.cites.tab")}
    out = {}
    for ds in ("cora", "pubmed"):
        np.random.seed(SEED)
        dc = DataCenter(config)
        dc.load_dataSet(ds)
        feats = getattr(dc, ds + "_feats")
        n = feats.shape[0]
        rowptr, col = csr_of(getattr(dc, ds + "_adj_lists"), n)
        out[ds + "_feats64"] = np.asarray(feats, dtype=np.float64)
        out[ds + "_labels"] = getattr(dc, ds + "_labels")
        out[ds + "_rowptr"], out[ds + "_col"] = rowptr, col
        for part in ("test", "val", "train"):
            out[f"{ds}_{part}"] = getattr(dc, f"{ds}_{part}")
        print(ds, "nodes", n, "entries", len(col), "feats", feats.shape[1], "labels", np.bincount(out[ds + "_labels"]))
    np.savez_compressed(os.path.join(OUT, "reference_outputs.npz"), **out)


if __name__ == "__main__":
    main()
