"""CPU oracle for the GraphSAGE minibatch hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a restatement, in plain Python + torch-CPU fp32, of the algorithm that
`Lolash/graphSAGE-pytorch` runs in `src/models.py` (and the loss lines of `src/utils.py`).
It exists so that the CUDA path can be checked for parity and so that `bench.py` can time
the reference's CPU algorithm beside the GPU number.  It is NOT part of the product:
only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  Nothing under `graphsage-pytorch_b200/` imports it, and the
product path raises when its CUDA library is missing rather than falling back to this.

Parity pin: every function below is checked against the *imported reference itself*
(`oracle/ref_harness.py`, run in the build container where `/root/reference` exists) by
`tests/golden/make_golden.py`, which also writes the committed fixtures under
`tests/golden/`; `tests/test_oracle_golden.py` re-checks the oracle against those fixtures
on every run.  The reference ships no tests or golden vectors of its own, so
the pin is "reference executed on torch 2.11 CPU with the py>=3.11 `random.sample` shim".

Deliberate fidelity: sampling uses Python sets and `random.sample` in exactly the call
order of the reference so that, under the same `random.seed`, the oracle draws the same
neighbours as the reference; MEAN goes through the same dense row-normalised mask and
`mask.mm(embed)` so that the CPU baseline pays what the reference pays.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import random as _random
from typing import Dict, Iterable, List, Mapping, Optional, Sequence, Set, Tuple

import numpy as np
import torch
import torch.nn.functional as F

NUM_SAMPLE_DEFAULT = 10  # src/models.py:277 (default argument, never overridden by forward)


# --------------------------------------------------------------------------------------
# A1 / A2  sampling + unique/remap                     src/models.py:277-289
# --------------------------------------------------------------------------------------
def _sample_wo_replacement(rng, population: Set[int], k: int) -> List[int]:
    """`random.sample(set, k)` as CPython <= 3.10 evaluated it: the set is first turned
    into a tuple (set iteration order), then k distinct positions are drawn.  Python 3.11+
    rejects a set population, so the conversion is spelled out."""
    return rng.sample(tuple(population), k)


def sample_neighbors(adj: Mapping[int, Set[int]], nodes: Sequence[int], num_sample: Optional[int] = NUM_SAMPLE_DEFAULT,
                     rng=_random) -> List[Set[int]]:
    """src/models.py:279-285.  One set per destination node: all neighbours when the row
    has fewer than `num_sample` of them, else `num_sample` distinct uniform picks; the
    node itself is then always added."""
    rows = [adj[int(n)] for n in nodes]                                   # :279
    if num_sample is not None:                                            # :280
        picked = [set(_sample_wo_replacement(rng, row, num_sample)) if len(row) >= num_sample else row
                  for row in rows]                                        # :282
    else:
        picked = rows                                                     # :284
    return [p | {nodes[i]} for i, p in enumerate(picked)]                 # :285


def unique_and_index(samp_neighs: Sequence[Set[int]]) -> Tuple[Dict[int, int], List[int]]:
    """src/models.py:286-288.  U = list(union of all rows) in CPython set order and the
    node -> position dictionary."""
    uniq = list(set.union(*samp_neighs))                                  # :286
    return dict(zip(uniq, range(len(uniq)))), uniq                        # :287-288


def get_unique_neighs_list(adj, nodes, num_sample=NUM_SAMPLE_DEFAULT, rng=_random):
    """src/models.py:277-289, same return triple (samp_neighs, unique_nodes, unique_list)."""
    samp = sample_neighbors(adj, nodes, num_sample, rng)
    index_of, uniq = unique_and_index(samp)
    return samp, index_of, uniq


def canonical_unique_remap(nodes: Sequence[int], samp_neighs: Sequence[Set[int]], drop_self: bool,
                           width: Optional[int] = None):
    """Order-free statement of rows A2/A3/A6 for the bit-exact index contract.

    Returns (U_sorted int64[|U|], self_idx int64[R], cols int32[R, width] padded with -1,
    cnt int32[R]) where `cols[r, :cnt[r]]` are the positions in U_sorted of row r's
    neighbours in ascending node id, with the row's own node removed when `drop_self`
    (src/models.py:298) or kept (gcn, self was added at :285).  The device path produces
    exactly this form, so the comparison is integer-exact."""
    uniq = np.array(sorted(set.union(*[set(int(x) for x in s) for s in samp_neighs])), dtype=np.int64)
    rows = []
    for i, s in enumerate(samp_neighs):
        me = int(nodes[i])
        ids = sorted(int(x) for x in s if not (drop_self and int(x) == me))
        rows.append(ids)
    if width is None:
        width = max(1, max(len(r) for r in rows))
    cols = np.full((len(rows), width), -1, dtype=np.int32)
    cnt = np.zeros(len(rows), dtype=np.int32)
    for i, ids in enumerate(rows):
        cnt[i] = len(ids)
        cols[i, :len(ids)] = np.searchsorted(uniq, np.asarray(ids, dtype=np.int64))
    self_idx = np.searchsorted(uniq, np.asarray([int(n) for n in nodes], dtype=np.int64))
    return uniq, self_idx.astype(np.int64), cols, cnt


# --------------------------------------------------------------------------------------
# A3 / A4 / A5  aggregation                              src/models.py:291-330
# --------------------------------------------------------------------------------------
def aggregate(nodes: Sequence[int], pre_hidden: torch.Tensor, pre_neighs, gcn: bool, agg_func: str) -> torch.Tensor:
    """src/models.py:291-330.  `pre_neighs` = (unique_list, samp_neighs, unique_dict)."""
    uniq_list, samp_neighs, index_of = pre_neighs                         # :292
    assert len(nodes) == len(samp_neighs)                                 # :294
    assert all(nodes[i] in samp_neighs[i] for i in range(len(samp_neighs)))  # :295-296
    if not gcn:                                                           # :297-298
        samp_neighs = [samp_neighs[i] - {nodes[i]} for i in range(len(samp_neighs))]
    if len(pre_hidden) == len(index_of):                                  # :300-301
        embed = pre_hidden
    else:                                                                 # :303
        embed = pre_hidden[torch.LongTensor(uniq_list)]
    mask = torch.zeros(len(samp_neighs), len(index_of))                   # :305
    col_idx = [index_of[n] for row in samp_neighs for n in row]           # :306
    row_idx = [i for i in range(len(samp_neighs)) for _ in range(len(samp_neighs[i]))]  # :307
    mask[row_idx, col_idx] = 1                                            # :308
    if agg_func == 'MEAN':                                                # :311-314
        deg = mask.sum(1, keepdim=True)
        mask = mask.div(deg).to(embed.device)
        return mask.mm(embed)
    if agg_func == 'MAX':                                                 # :316-326
        out = []
        for row in (mask == 1):
            hit = row.nonzero()
            feat = embed[hit.squeeze()]
            if feat.dim() == 1:                                           # :322-323 single neighbour
                out.append(feat.view(1, -1))
            else:                                                         # :325 (raises on an empty row)
                out.append(torch.max(feat, 0)[0].view(1, -1))
        return torch.cat(out, 0)
    raise ValueError(agg_func)


# --------------------------------------------------------------------------------------
# A7  SageLayer                                          src/models.py:209-220
# --------------------------------------------------------------------------------------
def sage_layer(weight: torch.Tensor, self_feats: torch.Tensor, agg_feats: torch.Tensor, gcn: bool) -> torch.Tensor:
    """relu(W . cat[self, agg]^T)^T ; gcn consumes only the aggregate (src/models.py:215-220)."""
    combined = agg_feats if gcn else torch.cat([self_feats, agg_feats], dim=1)   # :215-218
    return F.relu(weight.mm(combined.t())).t()                            # :219


def xavier_uniform(rng: np.random.Generator, out_size: int, in_size: int) -> torch.Tensor:
    """Same distribution as nn.init.xavier_uniform_ (src/models.py:207,23) but drawn from a
    numpy Generator so fixtures are reproducible without torch's RNG."""
    bound = float(np.sqrt(6.0 / (in_size + out_size)))
    return torch.from_numpy(rng.uniform(-bound, bound, size=(out_size, in_size)).astype(np.float32))


# --------------------------------------------------------------------------------------
# GraphSage.forward                                      src/models.py:241-269
# --------------------------------------------------------------------------------------
def graphsage_forward(weights: Sequence[torch.Tensor], raw_features: torch.Tensor, adj, nodes_batch,
                      gcn: bool = False, agg_func: str = 'MEAN', rng=_random,
                      injected: Optional[Sequence[Sequence[Set[int]]]] = None, record: Optional[list] = None):
    """src/models.py:241-269.  `weights[l]` is sage_layer{l+1}.weight.

    `injected`, when given, is the per-call list of `(samp_neighs, unique_list)` in the order
    the reference makes the calls (the call for the batch itself comes first); it replaces
    `get_unique_neighs_list` so a recorded reference run can be replayed exactly.  The
    recorded `unique_list` carries the CPython set order the reference saw (rows of the
    next call are in that order); pass `None` for it to rebuild the order locally.
    `record`, when given, receives (nodes, samp_neighs, unique_list) per call."""
    num_layers = len(weights)
    lower = list(nodes_batch)                                             # :246
    layers = [(lower,)]                                                   # :247
    for call in range(num_layers):                                        # :249-251
        if injected is not None:
            rows, order = injected[call]
            samp = [set(s) for s in rows]
            if order is None:
                index_of, uniq = unique_and_index(samp)
            else:
                uniq = list(order)
                index_of = dict(zip(uniq, range(len(uniq))))
        else:
            samp, index_of, uniq = get_unique_neighs_list(adj, lower, rng=rng)
        if record is not None:
            record.append((list(lower), [set(s) for s in samp], list(uniq)))
        layers.insert(0, (uniq, samp, index_of))
        lower = uniq
    assert len(layers) == num_layers + 1                                  # :253
    h = raw_features                                                      # :255
    for index in range(1, num_layers + 1):                                # :256-267
        nb = layers[index][0]
        pre = layers[index - 1]
        agg = aggregate(nb, h, pre, gcn, agg_func)                        # :260
        if index > 1:                                                     # :262-263, :271-275
            nb = [pre[2][x] for x in nb]
        h = sage_layer(weights[index - 1], h[nb], agg, gcn)               # :265-266
    return h


# --------------------------------------------------------------------------------------
# A11 / A12  classifier + supervised loss               src/models.py:25-27, src/utils.py:161-163
# --------------------------------------------------------------------------------------
def classification(cls_weight: torch.Tensor, cls_bias: torch.Tensor, embeds: torch.Tensor) -> torch.Tensor:
    return torch.log_softmax(F.linear(embeds, cls_weight, cls_bias), 1)   # models.py:26


def supervised_loss(log_probs: torch.Tensor, labels_batch) -> torch.Tensor:
    loss = -torch.sum(log_probs[range(log_probs.size(0)), labels_batch], 0)   # utils.py:162
    return loss / log_probs.size(0)                                       # utils.py:163


# --------------------------------------------------------------------------------------
# A8  batch extension: positives / negatives            src/models.py:135-186
# --------------------------------------------------------------------------------------
class PairSampler:
    """State-for-state restatement of UnsupervisedLoss's sampling half (src/models.py:45-63,
    135-186).  Attribute names follow the reference (including its spelling) because the
    loss functions read them."""
    Q = 10             # :49
    N_WALKS = 6        # :50
    WALK_LEN = 1       # :51
    N_WALK_LEN = 5     # :52
    MARGIN = 3         # :53

    def __init__(self, adj, train_nodes, rng=_random):
        self.adj_lists = adj
        self.train_nodes = train_nodes
        self.rng = rng
        self.target_nodes = None
        self.positive_pairs: list = []
        self.negtive_pairs: list = []
        self.node_positive_pairs: dict = {}
        self.node_negtive_pairs: dict = {}
        self.unique_nodes_batch: list = []

    def extend_nodes(self, nodes, num_neg=6):                             # :135-148
        self.positive_pairs, self.node_positive_pairs = [], {}
        self.negtive_pairs, self.node_negtive_pairs = [], {}
        self.target_nodes = nodes
        self._random_walks(nodes)
        self._negatives(nodes, num_neg)
        self.unique_nodes_batch = list(set(i for p in self.positive_pairs for i in p)
                                       | set(i for p in self.negtive_pairs for i in p))   # :146
        assert set(self.target_nodes) < set(self.unique_nodes_batch)      # :147
        return self.unique_nodes_batch

    def _negatives(self, nodes, num_neg):                                 # :153-167
        for node in nodes:
            ball = {node}
            frontier = {node}
            for _ in range(self.N_WALK_LEN):                              # :157-162
                reached = set()
                for v in frontier:
                    reached |= self.adj_lists[int(v)]
                frontier = reached - ball
                ball |= reached
            far = set(self.train_nodes) - ball                            # :163
            picks = _sample_wo_replacement(self.rng, far, num_neg) if num_neg < len(far) else far   # :164
            pairs = [(node, n) for n in picks]
            self.negtive_pairs.extend(pairs)                              # :165
            self.node_negtive_pairs[node] = list(pairs)                   # :166

    def _random_walks(self, nodes):                                       # :169-186
        for node in nodes:
            if len(self.adj_lists[int(node)]) == 0:                       # :171-172
                continue
            mine = []
            for _ in range(self.N_WALKS):
                cur = node
                for _ in range(self.WALK_LEN):
                    nxt = self.rng.choice(list(self.adj_lists[int(cur)]))  # :177-178
                    if nxt != node and nxt in self.train_nodes:           # :180
                        self.positive_pairs.append((node, nxt))
                        mine.append((node, nxt))
                    cur = nxt
            self.node_positive_pairs[node] = mine                         # :185


# --------------------------------------------------------------------------------------
# A9 / A10  unsupervised losses                         src/models.py:65-132
# --------------------------------------------------------------------------------------
def _pair_cos(embeddings, pairs, node2index):
    left = [node2index[a] for a, _ in pairs]
    right = [node2index[b] for _, b in pairs]
    return F.cosine_similarity(embeddings[left], embeddings[right])      # :82,90,116,122


def loss_sage(sampler: PairSampler, embeddings: torch.Tensor, nodes) -> torch.Tensor:
    """src/models.py:65-98 ("normal" skip-gram loss)."""
    assert len(embeddings) == len(sampler.unique_nodes_batch)             # :66
    assert all(nodes[i] == sampler.unique_nodes_batch[i] for i in range(len(nodes)))   # :67
    node2index = {n: i for i, n in enumerate(sampler.unique_nodes_batch)}  # :68
    assert len(sampler.node_positive_pairs) == len(sampler.node_negtive_pairs)         # :71
    scores = []
    for node in sampler.node_positive_pairs:                              # :72
        pps, nps = sampler.node_positive_pairs[node], sampler.node_negtive_pairs[node]
        if len(pps) == 0 or len(nps) == 0:                                # :75-76
            continue
        neg = _pair_cos(embeddings, nps, node2index)
        neg = sampler.Q * torch.mean(torch.log(torch.sigmoid(-neg)), 0)   # :83
        pos = torch.log(torch.sigmoid(_pair_cos(embeddings, pps, node2index)))   # :91
        scores.append(torch.mean(-pos - neg).view(1, -1))                 # :94
    return torch.mean(torch.cat(scores, 0))                               # :96


def loss_margin(sampler: PairSampler, embeddings: torch.Tensor, nodes) -> torch.Tensor:
    """src/models.py:100-132 (max-margin loss; result has shape [1])."""
    assert len(embeddings) == len(sampler.unique_nodes_batch)             # :101
    assert all(nodes[i] == sampler.unique_nodes_batch[i] for i in range(len(nodes)))   # :102
    node2index = {n: i for i, n in enumerate(sampler.unique_nodes_batch)}
    assert len(sampler.node_positive_pairs) == len(sampler.node_negtive_pairs)         # :106
    scores = []
    for node in sampler.node_positive_pairs:
        pps, nps = sampler.node_positive_pairs[node], sampler.node_negtive_pairs[node]
        if len(pps) == 0 or len(nps) == 0:                                # :110-111
            continue
        pos, _ = torch.min(torch.log(torch.sigmoid(_pair_cos(embeddings, pps, node2index))), 0)   # :117
        neg, _ = torch.max(torch.log(torch.sigmoid(_pair_cos(embeddings, nps, node2index))), 0)   # :123
        scores.append(torch.max(torch.tensor(0.0), neg - pos + sampler.MARGIN).view(1, -1))      # :125
    return torch.mean(torch.cat(scores, 0), 0)                            # :128


# --------------------------------------------------------------------------------------
# one full fwd+bwd step, the unit bench.py's cpu_baseline times   src/utils.py:157-184
# --------------------------------------------------------------------------------------
def supervised_step(weights, cls_weight, cls_bias, raw_features, adj, nodes_batch, labels, gcn=False,
                    agg_func='MEAN', rng=_random, injected=None):
    """forward -> classifier -> NLL mean -> backward (src/utils.py:157-163,184).  Returns
    (loss, embeddings, log_probs); gradients are left on the leaf tensors."""
    embs = graphsage_forward(weights, raw_features, adj, nodes_batch, gcn, agg_func, rng, injected)
    logp = classification(cls_weight, cls_bias, embs)
    loss = supervised_loss(logp, labels[np.asarray(nodes_batch)])
    loss.backward()
    return loss, embs, logp


# --------------------------------------------------------------------------------------
# N4  classifier training on frozen embeddings           src/utils.py:80-111
# --------------------------------------------------------------------------------------
def train_classification(cls_weight: torch.Tensor, cls_bias: torch.Tensor, features: torch.Tensor, labels,
                         orders: Sequence[Sequence[int]], b_sz: int = 50, lr: float = 0.5, max_norm: float = 5.0):
    """The loop body of src/utils.py:90-107 for given per-epoch node orders (`orders[e]` is what
    `shuffle(train_nodes)` returned in epoch e, :91).  Plain SGD is stateless, so the optimizer of :82 is the
    in-place update below.  Returns the trained (weight, bias) and the last batch's loss."""
    w = cls_weight.clone().requires_grad_(True)
    b = cls_bias.clone().requires_grad_(True)
    labels = np.asarray(labels)
    loss = None
    for order in orders:
        order = np.asarray(order)
        for lo in range(0, len(order), b_sz):                             # :92-94
            nodes_batch = order[lo:lo + b_sz]
            logists = classification(w, b, features[nodes_batch])         # :97-99
            loss = -torch.sum(logists[range(logists.size(0)), labels[nodes_batch]], 0)    # :100
            loss = loss / len(nodes_batch)                                # :101
            loss.backward()                                               # :104
            torch.nn.utils.clip_grad_norm_([w, b], max_norm)              # :106
            with torch.no_grad():                                         # :107-108
                w -= lr * w.grad
                b -= lr * b.grad
                w.grad.zero_()
                b.grad.zero_()
    return w.detach(), b.detach(), (loss.detach() if loss is not None else None)


# --------------------------------------------------------------------------------------
# adjacency helpers shared by tests / bench (data plumbing, not reference behaviour)
# --------------------------------------------------------------------------------------
class LazySetAdjacency(Mapping):
    """dict-of-sets view over a CSR, materialising `set(neighbours)` on first touch.
    Lets the oracle run on graphs whose full dict-of-sets (~90 B per entry, SURVEY.md
    Appendix B) would take minutes to build; semantics for the reference algorithm are
    those of `defaultdict(set)` from src/dataCenter.py:33 (missing key -> empty set)."""

    def __init__(self, rowptr: np.ndarray, col: np.ndarray, cache: bool = True):
        self.rowptr, self.col, self._cache, self._on = rowptr, col, {}, cache

    def __getitem__(self, node):
        node = int(node)
        hit = self._cache.get(node)
        if hit is None:
            if 0 <= node < len(self.rowptr) - 1:
                hit = set(self.col[self.rowptr[node]:self.rowptr[node + 1]].tolist())
            else:
                hit = set()
            if self._on:
                self._cache[node] = hit
        return hit

    def __iter__(self):
        return iter(range(len(self.rowptr) - 1))

    def __len__(self):
        return len(self.rowptr) - 1


def csr_to_adj_dict(rowptr: np.ndarray, col: np.ndarray):
    from collections import defaultdict
    adj = defaultdict(set)
    for v in range(len(rowptr) - 1):
        adj[v] = set(col[rowptr[v]:rowptr[v + 1]].tolist())
    return adj
