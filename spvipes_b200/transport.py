"""`process_transport_plan`: cluster labels shared by the two groups, derived from a transport plan (setup-time, host side).

Restates reference model/spvipes.py:26-162: per group, cluster the cells at seven resolutions, score each resolution by the
negative mean entropy of the clusters' normalised transport distributions (:53-73), keep the best, name the clusters
`<group>_<id>`; take the MEDIAN transport value between every cluster pair (:104-118), match source and target clusters with
the Hungarian algorithm on the negated medians (:120-145), rename matched pairs `Cluster_<i>`, park every unmatched cluster
in one extra name (:139-143), and return an ordered categorical (:147-160).

The clustering itself is pluggable: the reference calls scanpy (normalize_total -> log1p -> pca -> neighbors -> leiden,
:88-101), which is not a dependency here.  `cluster_fn(X_group, resolution) -> labels` defaults to scanpy's Leiden when scanpy
is importable and otherwise to `knn_louvain` below (same preprocessing restated in numpy, a kNN graph and Louvain modularity
optimisation with a resolution parameter - the same objective Leiden refines, not the same partition).  Everything after the
clustering is deterministic and is pinned against the unmodified reference function (oracle/make_golden_transport.py ->
tests/golden_transport/, tests/test_transport.py).

The reference materialises an N1 x N2-row DataFrame for the pivot; here the median per cluster pair is taken on the sub-block
directly (same values, O(N1 N2) reads, no copy of the plan).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence

import numpy as np
import pandas as pd

RESOLUTIONS = (0.1, 0.3, 0.5, 0.7, 1.0, 1.5, 2.0)  # reference :53


def _entropy_rows(p: np.ndarray) -> np.ndarray:
    """scipy.stats.entropy(p, axis=1) for rows that already sum to one (natural log, 0 log 0 = 0)"""
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(p > 0, p * np.log(p), 0.0)
    return -t.sum(axis=1)


def resolution_score(labels: Sequence, group_plan: np.ndarray) -> float:
    """negative mean entropy of the per-cluster transport distributions (reference :57-71); clusters in order of appearance,
    as `Series.unique()` lists them"""
    labels = np.asarray(labels)
    uniq = pd.unique(labels)
    ct = np.zeros((len(uniq), group_plan.shape[1]))
    for i, c in enumerate(uniq):
        ct[i] = group_plan[labels == c].sum(axis=0)
    ct /= ct.sum(axis=1, keepdims=True)
    return float(-np.mean(_entropy_rows(ct)))


def match_clusters(plan: np.ndarray, clusters1: Sequence[str], clusters2: Sequence[str]) -> Dict[str, str]:
    """rename dictionary of reference :104-145: pivot of median transport values (rows / columns sorted, as pandas'
    pivot_table sorts them), Hungarian assignment on the negated pivot, unmatched clusters share one extra name"""
    from scipy.optimize import linear_sum_assignment
    c1, c2 = np.asarray(clusters1), np.asarray(clusters2)
    src, tgt = sorted(set(c1.tolist())), sorted(set(c2.tolist()))
    pivot = np.empty((len(src), len(tgt)))
    cols = [np.flatnonzero(c2 == t) for t in tgt]
    for i, s in enumerate(src):
        rows = plan[c1 == s]
        for j, cj in enumerate(cols):
            pivot[i, j] = np.median(rows[:, cj])
    row_ind, col_ind = linear_sum_assignment(-pivot)
    rename: Dict[str, str] = {}
    for i, (si, ti) in enumerate(zip(row_ind, col_ind)):
        rename[src[si]] = f"Cluster_{i}"
        rename[tgt[ti]] = f"Cluster_{i}"
    for c in (set(src) | set(tgt)) - set(rename):  # the reference re-evaluates len(rename_dict) // 2 per unmatched cluster (:141-143)
        rename[c] = f"Cluster_{len(rename) // 2}"
    return rename


# ------------------------------------------------------------------------------------------------ default clustering
def _preprocess(X: np.ndarray, n_comps: int = 50) -> np.ndarray:
    """normalize_total (target = median of the row totals) -> log1p -> PCA scores (scanpy defaults: 50 components, centred)"""
    X = np.asarray(X, dtype=np.float64)
    tot = X.sum(axis=1)
    target = np.median(tot[tot > 0]) if np.any(tot > 0) else 1.0
    X = np.log1p(X / np.maximum(tot, 1e-12)[:, None] * target)
    X = X - X.mean(axis=0, keepdims=True)
    k = int(min(n_comps, min(X.shape) - 1))
    if k < 1:
        return X
    u, s, _ = np.linalg.svd(X, full_matrices=False)
    return u[:, :k] * s[:k]


def _knn_graph(Z: np.ndarray, k: int = 15):
    """symmetric unweighted k-nearest-neighbour graph (scanpy's default n_neighbors = 15) as a scipy CSR matrix"""
    import scipy.sparse as sp
    n = Z.shape[0]
    k = int(min(k, n - 1))
    sq = (Z * Z).sum(1)
    rows, cols = [], []
    step = max(1, 2 ** 22 // max(n, 1))
    for lo in range(0, n, step):
        d = sq[lo:lo + step, None] - 2.0 * Z[lo:lo + step] @ Z.T + sq[None, :]
        d[np.arange(d.shape[0]), np.arange(lo, lo + d.shape[0])] = np.inf
        nn = np.argpartition(d, k - 1, axis=1)[:, :k] if k > 0 else np.empty((d.shape[0], 0), dtype=int)
        rows.append(np.repeat(np.arange(lo, lo + d.shape[0]), k))
        cols.append(nn.ravel())
    r, c = np.concatenate(rows), np.concatenate(cols)
    A = sp.csr_matrix((np.ones(len(r)), (r, c)), shape=(n, n))
    A = ((A + A.T) > 0).astype(np.float64)
    return A.tocsr()


def _louvain(A, resolution: float, max_levels: int = 10) -> np.ndarray:
    """Louvain modularity optimisation (Reichardt-Bornholdt resolution), deterministic node order; returns one label per node"""
    import scipy.sparse as sp
    n = A.shape[0]
    node_of = np.arange(n)          # community (at the current level's graph) of every original node
    W = A.tocsr().astype(np.float64)
    for _ in range(max_levels):
        m2 = W.sum()
        if m2 <= 0:
            break
        deg = np.asarray(W.sum(axis=1)).ravel()
        comm = np.arange(W.shape[0])
        tot = deg.copy()
        indptr, indices, data = W.indptr, W.indices, W.data
        moved_any = False
        for _sweep in range(50):
            moved = 0
            for i in range(W.shape[0]):
                ci = comm[i]
                nb, w = indices[indptr[i]:indptr[i + 1]], data[indptr[i]:indptr[i + 1]]
                if len(nb) == 0:
                    continue
                link: Dict[int, float] = {}
                for j, wij in zip(nb, w):
                    if j != i:
                        link[comm[j]] = link.get(comm[j], 0.0) + wij
                tot[ci] -= deg[i]
                best, best_gain = ci, link.get(ci, 0.0) - resolution * tot[ci] * deg[i] / m2
                for c, l in sorted(link.items()):
                    gain = l - resolution * tot[c] * deg[i] / m2
                    if gain > best_gain + 1e-12:
                        best, best_gain = c, gain
                tot[best] += deg[i]
                if best != ci:
                    comm[i] = best
                    moved += 1
            if moved == 0:
                break
            moved_any = True
        if not moved_any:
            break
        _, comm = np.unique(comm, return_inverse=True)
        node_of = comm[node_of]
        k = comm.max() + 1
        S = sp.csr_matrix((np.ones(len(comm)), (np.arange(len(comm)), comm)), shape=(len(comm), k))
        W = (S.T @ W @ S).tocsr()
    _, labels = np.unique(node_of, return_inverse=True)
    return labels


def knn_louvain(X: np.ndarray, resolution: float) -> np.ndarray:
    """default `cluster_fn` when scanpy is absent: the reference's preprocessing restated + kNN graph + Louvain"""
    return _louvain(_knn_graph(_preprocess(X)), resolution)


def _scanpy_leiden(X: np.ndarray, resolution: float) -> np.ndarray:
    import anndata as ad
    import scanpy as sc
    a = ad.AnnData(np.asarray(X, dtype=np.float32))
    sc.pp.normalize_total(a)
    sc.pp.log1p(a)
    sc.pp.pca(a)
    sc.pp.neighbors(a)
    sc.tl.leiden(a, resolution=resolution)
    return a.obs["leiden"].astype(str).to_numpy()


def default_cluster_fn() -> Callable[[np.ndarray, float], np.ndarray]:
    try:
        import anndata  # noqa: F401
        import scanpy  # noqa: F401
        return _scanpy_leiden
    except Exception:
        return knn_louvain


def process_transport_plan(transport_plan, adata, groups_key: str, cluster_fn: Optional[Callable] = None,
                           resolutions: Sequence[float] = RESOLUTIONS):
    """reference model/spvipes.py:26-162.  Sets adata.obs['group_cluster_labels'], adata.obs['processed_transport_labels'] and
    adata.uns['optimal_resolutions'] like the reference and returns the processed labels (ordered categorical values).
    cluster_fn(X_group [n_cells, n_group_genes], resolution) -> one label per cell."""
    plan = np.nan_to_num(np.asarray(transport_plan, dtype=np.float64), nan=0.0)
    cluster_fn = cluster_fn or default_cluster_fn()
    obs = adata.obs
    groups = pd.unique(obs[groups_key])
    if len(groups) != 2:
        raise ValueError("process_transport_plan expects exactly two groups")
    X = adata.X
    var_names = np.asarray(list(adata.var_names))
    cluster_labels = np.empty(len(obs), dtype=object)
    optimal: Dict = {}
    per_group = []
    for i, group in enumerate(groups):
        mask = (obs[groups_key] == group).to_numpy()
        gvars = adata.uns["groups_var_names"][group] if group in adata.uns["groups_var_names"] else adata.uns["groups_var_names"][i]
        vmask = np.isin(var_names, np.asarray(list(gvars)))
        Xg = X[mask][:, vmask]
        Xg = np.asarray(Xg.todense()) if hasattr(Xg, "todense") else np.asarray(Xg)
        gplan = plan if i == 0 else plan.T
        scores = [resolution_score(cluster_fn(Xg, res), gplan) for res in resolutions]
        best = resolutions[int(np.argmax(scores))]
        optimal[group] = best
        labels = np.asarray([f"{group}_{c}" for c in np.asarray(cluster_fn(Xg, best)).astype(str)], dtype=object)
        cluster_labels[mask] = labels
        per_group.append(labels)
    obs["group_cluster_labels"] = pd.Categorical(cluster_labels)
    rename = match_clusters(plan, per_group[0], per_group[1])
    categories = np.array(sorted(set(rename.values()), key=lambda x: int(x.split("_")[1])))
    obs["processed_transport_labels"] = pd.Categorical([rename[c] for c in cluster_labels], categories=categories, ordered=True)
    adata.uns["optimal_resolutions"] = optimal
    return obs["processed_transport_labels"].values
