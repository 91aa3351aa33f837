"""Synthetic multi-omics data + guidance graph in the shapes of BASELINE.json's configs."""
import networkx as nx
import numpy as np
import pandas as pd
import scipy.sparse

from glue_b200.anndata_lite import AnnData


def make_synthetic(n_cells=(10000, 10000), n_feat=(2000, 50000), n_pairs=50000, seed=0,
                   rates=(0.3, 0.05), rep_dim=100, paired=False, n_batches=0,
                   names=("rna", "atac")):
    rs = np.random.RandomState(seed)
    prefix = {"rna": "g", "atac": "p"}
    adatas, feats = {}, {}
    for k, n, f, rate in zip(names, n_cells, n_feat, rates):
        x = scipy.sparse.random(n, f, density=min(1.0, rate), format="csr", random_state=rs,
                                data_rvs=lambda size: rs.poisson(1.0, size) + 1.0
                                ).astype(np.float32)
        x[:, 0] = 1.0  # no empty library
        feats[k] = [f"{prefix.get(k, k)}{i}" for i in range(f)]
        cells = [f"cell{i}" for i in range(n)] if paired else [f"{k}_cell{i}" for i in range(n)]
        obs = pd.DataFrame(index=cells)
        if n_batches:
            obs["batch"] = rs.randint(0, n_batches, n).astype(str)
        adatas[k] = AnnData(X=x.tocsr(), obs=obs, var=pd.DataFrame(index=feats[k]),
                            obsm={"X_rep": rs.randn(n, rep_dim).astype(np.float32)})
    g = nx.Graph()
    a, b = names[0], names[1]
    g.add_nodes_from(feats[a] + feats[b])
    src = rs.randint(0, n_feat[0], n_pairs)
    dst = rs.randint(0, n_feat[1], n_pairs)
    wts = rs.uniform(0.05, 1.0, n_pairs)
    g.add_edges_from((feats[a][s], feats[b][t], {"weight": float(w), "sign": 1})
                     for s, t, w in zip(src, dst, wts))
    g.add_edges_from((v, v, {"weight": 1.0, "sign": 1}) for v in list(g.nodes))
    return adatas, g
