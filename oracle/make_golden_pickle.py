"""Write ``tests/golden/ref_model.dill`` + ``ref_model.npz``: a model object saved by the UNMODIFIED
reference (``Model.save``, scglue/models/base.py:364-383), with the reference imported under its real
package name so that the pickled class paths are the ones real model files carry
(``scglue.models.scglue.SCGLUEModel``, ``scglue.models.sc.NBDataDecoder``, ...), and what the reference's
own modules compute with it.  Run in the build container:

    python -m oracle.make_golden_pickle

TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np
import pandas as pd
import torch

from . import refstub

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    if "scglue" in sys.modules:
        sys.exit("run in a fresh interpreter: the alias package `scglue` must not be imported")
    refstub.ALIAS = "scglue"  # the reference under its own name
    ref = refstub.load()
    torch.manual_seed(5)
    torch.set_num_threads(1)
    from glue_b200.anndata_lite import AnnData
    rs = np.random.RandomState(0)
    genes, peaks = [f"g{i}" for i in range(9)], [f"p{i}" for i in range(14)]
    vertices = genes + peaks

    def adata(names, prob, rep_dim, batches):
        n = 7
        a = AnnData(X=rs.poisson(1.0, (n, len(names))).astype(np.float32),
                    obs=pd.DataFrame({"batch": rs.choice(batches, n)}, index=[f"c{i}" for i in range(n)]),
                    var=pd.DataFrame(index=names), obsm={"X_rep": rs.randn(n, rep_dim).astype(np.float32)})
        a.uns[ref.utils.config.ANNDATA_KEY] = {
            "prob_model": prob, "use_highly_variable": False, "features": list(names), "use_layer": None,
            "use_rep": "X_rep", "rep_dim": rep_dim, "use_batch": "batch", "batches": np.array(sorted(batches)),
            "use_cell_type": None, "cell_types": None, "use_dsc_weight": None, "use_obs_names": False}
        return a

    adatas = {"rna": adata(genes, "NB", 6, ["b0", "b1"]), "atac": adata(peaks, "ZINB", 5, ["b0", "b1", "b2"])}
    model = ref.models.scglue.SCGLUEModel(adatas, vertices, latent_dim=4, h_depth=2, h_dim=8, dropout=0.2,
                                          random_seed=3)
    with torch.no_grad():  # decoders / vertex table are zero-initialised
        model.net.g2v.vrepr.copy_(torch.randn(len(vertices), 4) * 0.3)
        for p in model.net.u2x.parameters():
            p.copy_(torch.randn_like(p) * 0.3)
    path = os.path.join(GOLDEN_DIR, "ref_model.dill")
    model.save(path)
    net = model.net
    net.eval()
    out = {f"state__{k}": v.detach().numpy() for k, v in net.state_dict().items()}
    out["vertices"] = np.array(vertices)
    for k, a in adatas.items():
        x, xrep = torch.as_tensor(a.X), torch.as_tensor(a.obsm["X_rep"])
        u, _ = net.x2u[k](x, xrep, lazy_normalizer=True)
        out[f"x__{k}"], out[f"xrep__{k}"], out[f"u_mean__{k}"] = a.X, a.obsm["X_rep"], u.mean.detach().numpy()
    # decoder means of the rna cells in the atac modality, given a feature embedding
    v = torch.randn(len(vertices), 4) * 0.5
    u_rna = torch.as_tensor(out["u_mean__rna"])
    b = torch.as_tensor(rs.randint(0, 3, u_rna.shape[0]))
    l = torch.full((u_rna.shape[0], 1), 50.0)
    dist = net.u2x["atac"](u_rna, v[getattr(net, "atac_idx")], b, l)
    out.update(dec_v=v.numpy(), dec_b=b.numpy(), dec_l=l.numpy(), dec_mean=dist.mean.detach().numpy())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ref_model.npz"), **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
