"""Diagnostics: the layer-1 aggregation kernel on 16 distinct cfg-3 frontiers (graph-replayed chain), for the library
given by GSAGE_LIB (variants built with -DGS_AGG_BATCH / -DGS_AGG_CTAS)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import graphsage_b200  # noqa
from graphsage_b200 import models, native, ops
from graphsage_b200.graph import AdjCSR
dev = torch.device('cuda:0')
cfg, rowptr, col, feats, labels, train = bench.build_workload(1.0)
model = models.GraphSage(2, cfg['feats'], cfg['hidden'], torch.from_numpy(feats).to(dev), AdjCSR(rowptr, col), dev, seed=824).to(dev)
batches = torch.from_numpy(bench.batches_for(train, 1024, 32, 0, 1).astype(np.int32)).to(dev)
csr, table, _ = model._state()
fronts, nbytes = [], []
for i in range(16):
    fr = model._run_agg1(model._run_sample(batches[i].contiguous(), None))[0]
    rows = int(fr.num_rows.item()); nnz = int(fr.cnt[:rows].sum().item())
    nbytes.append(nnz * 400 + rows * 400 + nnz * 4 + (rows + 1) * 4)
    fronts.append((fr, torch.empty_like(fr.agg)))
t = bench._chain_us(torch, dev, [(lambda f=f, o=o: ops.agg_fwd(table, 100, f.nbr, f.stride, f.cnt, f.num_rows, f.rows_max, 0, out=o)) for f, o in fronts], reps=8)
print(f"{os.environ.get('GSAGE_LIB', 'default'):40s} {t:6.2f} us  frac {np.mean(nbytes) / t / 1e3 / 6544:.3f}")
