"""TEST INFRASTRUCTURE.  Golden fixtures for `process_transport_plan` from the UNMODIFIED reference function
(/root/reference/src/spVIPES/model/spvipes.py:26-162), run in this container:

  * the function's source is cut out of the reference file with `ast` and executed as is;
  * scanpy (absent here) is replaced by a stand-in whose preprocessing calls are no-ops and whose `tl.leiden` assigns labels
    from a table prepared in advance (per group, per resolution) - the clustering is the pluggable part of
    spvipes_b200.transport, everything downstream of it (entropy scores, choice of resolution, pivot of medians, Hungarian
    matching, renaming, category order) is what these fixtures pin;
  * a minimal AnnData stand-in provides the slicing the function uses.

    python oracle/make_golden_transport.py        # writes tests/golden_transport/*.npz
"""
import ast
import os
import sys
import types

import numpy as np
import pandas as pd
from scipy.optimize import linear_sum_assignment
from scipy.stats import entropy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/spVIPES/model/spvipes.py"
OUT = os.path.join(ROOT, "tests", "golden_transport")


class FakeAnnData:
    """the part of AnnData the reference function touches: boolean row / column slicing, .copy(), .obs, .uns, .var_names, .shape"""

    def __init__(self, X, obs, var_names, uns, tag=None):
        self.X, self.obs, self.var_names, self.uns, self.tag = X, obs, pd.Index(var_names), uns, tag

    @property
    def shape(self):
        return self.X.shape

    def copy(self):
        return FakeAnnData(self.X.copy(), self.obs.copy(), self.var_names.copy(), self.uns, self.tag)

    def __getitem__(self, key):
        if isinstance(key, tuple):
            rows, cols = key
            cols = np.asarray(cols)
            assert rows == slice(None)
            return FakeAnnData(self.X[:, cols], self.obs, self.var_names[cols], self.uns, self.tag)
        rows = np.asarray(key)
        sub = FakeAnnData(self.X[rows], self.obs[rows].copy(), self.var_names, self.uns, self.tag)
        g = sub.obs["groups"].unique()
        if len(g) == 1:
            sub.tag = g[0]
        return sub


def reference_function(leiden_table):
    """process_transport_plan compiled from the reference file, with `sc` replaced by the stand-in"""
    src = open(REF).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "process_transport_plan")
    code = ast.get_source_segment(src, node)
    sc = types.SimpleNamespace(pp=types.SimpleNamespace(normalize_total=lambda a: None, log1p=lambda a: None, pca=lambda a: None,
                                                        neighbors=lambda a: None),
                               tl=types.SimpleNamespace())

    def leiden(adata, resolution=1.0, key_added="leiden"):
        adata.obs[key_added] = pd.Categorical(leiden_table[adata.tag][resolution].astype(str))
    sc.tl.leiden = leiden
    ns = {"np": np, "pd": pd, "sc": sc, "entropy": entropy, "linear_sum_assignment": linear_sum_assignment, "tqdm": lambda x, **k: x}
    exec(compile(code, REF, "exec"), ns)
    return ns["process_transport_plan"]


def make_case(seed, n1, n2, g1, g2, n_types, nan_frac=0.0, extra_b=0):
    rs = np.random.RandomState(seed)
    t1, t2 = rs.randint(0, n_types, n1), rs.randint(0, n_types, n2)
    emb1 = np.eye(n_types)[t1] * 3 + rs.randn(n1, n_types)
    emb2 = np.eye(n_types)[t2] * 3 + rs.randn(n2, n_types)
    plan = np.exp(-((emb1[:, None, :] - emb2[None, :, :]) ** 2).sum(-1) / 4.0)
    plan /= plan.sum()
    if nan_frac:
        plan[rs.rand(n1, n2) < nan_frac] = np.nan
    X = np.zeros((n1 + n2, g1 + g2), dtype=np.float32)
    X[:n1, :g1] = rs.poisson(2.0, (n1, g1))
    X[n1:, g1:] = rs.poisson(2.0, (n2, g2))
    groups = np.array(["A"] * n1 + ["B"] * n2)
    var_names = [f"A_g{i}" for i in range(g1)] + [f"B_g{i}" for i in range(g2)]
    # "clusterings": the true types coarsened / refined differently per resolution, with some label noise
    resolutions = [0.1, 0.3, 0.5, 0.7, 1.0, 1.5, 2.0]
    table = {}
    for name, t, n in (("A", t1, n1), ("B", t2, n2)):
        table[name] = {}
        for k, res in enumerate(resolutions):
            n_cl = max(2, min(n_types + 2, 2 + k)) + (extra_b if name == "B" else 0)
            lab = (t * 7 + k) % n_cl
            flip = rs.rand(n) < 0.05 * (k % 3)
            lab = np.where(flip, rs.randint(0, n_cl, n), lab)
            table[name][res] = lab
    return plan, X, groups, var_names, table, resolutions


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {"small": (3, 60, 50, 12, 9, 4, 0.0), "ragged_nan": (11, 83, 131, 7, 15, 6, 0.02), "many_types": (5, 150, 120, 10, 10, 9, 0.0),
             "unmatched": (23, 90, 110, 8, 8, 5, 0.0, 1)}
    for name, args in cases.items():
        plan, X, groups, var_names, table, resolutions = make_case(*args)
        obs = pd.DataFrame({"groups": groups})
        uns = {"groups_var_names": {"A": [v for v in var_names if v.startswith("A_")], "B": [v for v in var_names if v.startswith("B_")]}}
        adata = FakeAnnData(X, obs, var_names, uns)
        fn = reference_function(table)
        labels = fn(plan.copy(), adata, "groups")
        save = {"plan": plan, "X": X, "groups": groups, "var_names": np.array(var_names), "resolutions": np.array(resolutions),
                "labels": np.asarray(labels).astype(str), "categories": np.asarray(labels.categories).astype(str),
                "group_cluster_labels": np.asarray(adata.obs["group_cluster_labels"]).astype(str),
                "optimal_A": adata.uns["optimal_resolutions"]["A"], "optimal_B": adata.uns["optimal_resolutions"]["B"]}
        for g in ("A", "B"):
            for res in resolutions:
                save[f"leiden_{g}_{res}"] = table[g][res]
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **save)
        print(name, "optimal", adata.uns["optimal_resolutions"], "categories", list(labels.categories))


if __name__ == "__main__":
    sys.exit(main())
