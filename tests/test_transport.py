"""-m "not gpu": spvipes_b200.transport.process_transport_plan against fixtures produced by the UNMODIFIED reference function
(oracle/make_golden_transport.py; reference model/spvipes.py:26-162) with the clustering step injected, plus a smoke test of the
built-in clustering (kNN graph + Louvain) used when scanpy is absent."""
import glob
import os

import numpy as np
import pandas as pd
import pytest

from spvipes_b200.model import GroupedData
from spvipes_b200.transport import knn_louvain, match_clusters, process_transport_plan, resolution_score

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, "golden_transport", "*.npz")))


def _adata(z):
    groups = z["groups"].astype(str)
    var_names = [str(v) for v in z["var_names"]]
    uns = {"groups_var_names": {g: [v for v in var_names if v.startswith(g + "_")] for g in ("A", "B")}}
    return GroupedData(X=z["X"], obs=pd.DataFrame({"groups": groups}), var_names=var_names, uns=uns)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_matches_reference_function(path):
    z = np.load(path, allow_pickle=False)
    adata = _adata(z)
    n = {g: int((z["groups"].astype(str) == g).sum()) for g in ("A", "B")}
    assert n["A"] != n["B"]
    by_rows = {n[g]: g for g in n}

    def cluster_fn(Xg, res):  # the reference run used these tables in place of scanpy's leiden
        g = by_rows[Xg.shape[0]]
        assert Xg.shape[1] == len(adata.uns["groups_var_names"][g])  # only the group's own genes are handed to the clustering
        return z[f"leiden_{g}_{res}"].astype(str)

    labels = process_transport_plan(z["plan"].copy(), adata, "groups", cluster_fn=cluster_fn)
    assert adata.uns["optimal_resolutions"] == {"A": float(z["optimal_A"]), "B": float(z["optimal_B"])}
    assert np.array_equal(np.asarray(adata.obs["group_cluster_labels"]).astype(str), z["group_cluster_labels"].astype(str))
    assert list(labels.categories) == [str(c) for c in z["categories"]]
    assert labels.ordered
    assert np.array_equal(np.asarray(labels).astype(str), z["labels"].astype(str))


def test_resolution_score_prefers_pure_clusters():
    plan = np.kron(np.eye(3), np.ones((10, 10))) + 1e-3
    pure, mixed = np.repeat([0, 1, 2], 10), np.tile([0, 1, 2], 10)
    assert resolution_score(pure, plan) > resolution_score(mixed, plan)


def test_match_clusters_pairs_by_median_transport():
    plan = np.kron(np.array([[0.0, 1.0], [1.0, 0.0]]), np.ones((5, 7))) + 0.01
    c1 = np.array(["A_0"] * 5 + ["A_1"] * 5)
    c2 = np.array(["B_0"] * 7 + ["B_1"] * 7)
    r = match_clusters(plan, c1, c2)
    assert r["A_0"] == r["B_1"] and r["A_1"] == r["B_0"] and r["A_0"] != r["A_1"]


def test_builtin_clustering_separates_blobs():
    rs = np.random.RandomState(0)
    prof = rs.gamma(2.0, 1.0, (3, 40)) * np.array([[1.0], [6.0], [0.2]]) + 0.05
    prof[1, :20] *= 8
    prof[2, 20:] *= 8
    t = np.repeat([0, 1, 2], 60)
    X = rs.poisson(prof[t] * 3.0)
    lab = knn_louvain(X, 0.5)
    assert len(lab) == len(t)
    # every true type is dominated by one cluster and the three dominant clusters differ
    dom = [np.bincount(lab[t == k]).argmax() for k in range(3)]
    assert len(set(dom)) == 3
    for k in range(3):
        assert (lab[t == k] == dom[k]).mean() > 0.9
