"""Host-side data logic: bit-exact negative sampling, pairing masks, loader semantics,
rank-strided batches (mirrors the reference's tests/models/test_data.py where applicable)."""
import networkx as nx
import numpy as np
import pandas as pd
import pytest
import scipy.sparse
import torch

from glue_b200 import models
from glue_b200.anndata_lite import AnnData
from glue_b200.dist import RankStridedBatches
from glue_b200.models.data import (AnnDataset, ArrayDataset, DataLoader, GraphDataset,
                                   ParallelDataLoader)
from glue_b200.utils import config


def _graph_from_golden(g):
    from tests.util import golden_graph
    graph, vertices, _, _ = golden_graph()
    assert list(vertices) == list(g["vertices"])
    return graph


def test_graph_dataset_matches_reference_bit_exact(golden):
    g = golden("graph")
    ds = GraphDataset(_graph_from_golden(g), pd.Index(g["vertices"]), neg_samples=3)
    assert np.array_equal(ds.eidx, g["eidx"]) and np.array_equal(ds.ewt, g["ewt"])
    assert np.array_equal(ds.esgn, g["esgn"])
    assert np.array_equal(ds.vprob, g["vprob"]) and np.array_equal(ds.eprob, g["eprob"])
    assert ds.size == int(g["size"])
    for seed in (0, 5):
        eidx, ewt, esgn = ds.propose_shuffle(seed)
        assert np.array_equal(eidx, g[f"samp{seed}_eidx"])
        assert np.array_equal(ewt, g[f"samp{seed}_ewt"])
        assert np.array_equal(esgn, g[f"samp{seed}_esgn"])
    # shuffle n uses seed random_seed + n; background look-ahead does not change the stream
    for workers in (0, 1):
        ds.prepare_shuffle(num_workers=workers, random_seed=5)
        ds.shuffle()
        assert np.array_equal(ds.samp_eidx.numpy(), g["samp5_eidx"])
        ds.clean()


def test_negative_samples_are_never_true_edges(golden):
    g = golden("graph")
    ds = GraphDataset(_graph_from_golden(g), pd.Index(g["vertices"]), neg_samples=5)
    eidx, ewt, esgn = ds.propose_shuffle(11)
    true = {(i, j, s) for (i, j), s in zip(ds.eidx.T, ds.esgn)}
    neg = ewt == 0
    assert neg.sum() == 5 * (~neg).sum()
    assert not any((i, j, s) in true for (i, j), s in zip(eidx[:, neg].T, esgn[neg]))
    assert all((i, j, s) in true for (i, j), s in zip(eidx[:, ~neg].T, esgn[~neg]))


def test_graph_dataset_validation():
    vertices = pd.Index(["a", "b"])
    for attrs, msg in (({"weight": 1.5, "sign": 1}, "weight"), ({"weight": 0.5, "sign": 2}, "sign")):
        graph = nx.Graph()
        graph.add_edge("a", "b", **attrs)
        with pytest.raises(ValueError, match=msg):
            GraphDataset(graph, vertices)
    graph = nx.Graph()
    graph.add_edge("a", "c", weight=0.5, sign=1)
    with pytest.raises(ValueError, match="Missing vertices"):
        GraphDataset(graph, vertices)


def test_array_dataset_split_and_shuffle():
    a, b = np.arange(100).reshape(50, 2), scipy.sparse.random(30, 4, 0.5, format="csr")
    ds = ArrayDataset(a, b, getitem_size=7)
    assert len(ds) == 8
    item = ds[7]
    assert item[0].shape == (1, 2) and item[1].shape == (1, 4)
    train, val = ds.random_split([0.8, 0.2], random_state=0)
    assert train.sizes == [40, 24] and val.sizes == [10, 6]
    assert set(train.view_idx[0]).isdisjoint(val.view_idx[0])
    train.prepare_shuffle(num_workers=0, random_seed=3)
    train.shuffle()
    assert sorted(train.shuffle_idx[0]) == sorted(train.view_idx[0])
    with pytest.raises(ValueError):
        ds.random_split([0.5, 0.6])
    with pytest.raises(ValueError):
        ArrayDataset(np.empty((0, 3)))


def _paired_adatas():
    rs = np.random.RandomState(0)
    shared = [f"s{i}" for i in range(100)]
    rna_names = shared + [f"r{i}" for i in range(50)]
    atac_names = shared[::-1] + [f"a{i}" for i in range(20)]
    rna = AnnData(X=rs.poisson(2, (150, 5)).astype(np.float32), obs=pd.DataFrame(index=rna_names),
                  var=pd.DataFrame(index=[f"g{i}" for i in range(5)]),
                  layers={"arange": np.tile(np.arange(150)[:, None], (1, 5)).astype(np.float32)})
    atac = AnnData(X=scipy.sparse.csr_matrix(rs.poisson(1, (120, 6)).astype(np.float32)),
                   obs=pd.DataFrame(index=atac_names),
                   var=pd.DataFrame(index=[f"p{i}" for i in range(6)]),
                   layers={"arange": scipy.sparse.csr_matrix(
                       np.tile(np.array([shared.index(n) if n in shared else -1
                                         for n in atac_names])[:, None] + 0.0, (1, 6)).astype(np.float32))})
    return rna, atac, shared


def test_anndataset_pairing_mask():
    """Rows that are marked as paired must carry the same cell in every modality
    (property checked by the reference's tests/models/test_data.py:40-96)."""
    rna, atac, shared = _paired_adatas()
    for a in (rna, atac):
        models.configure_dataset(a, "NB", use_highly_variable=False, use_layer="arange",
                                 use_obs_names=True)
    ds = AnnDataset([rna, atac], [rna.uns[config.ANNDATA_KEY], atac.uns[config.ANNDATA_KEY]],
                    mode="train", getitem_size=16)
    assert ds.size == 170  # 100 shared + 50 + 20
    ds.prepare_shuffle(num_workers=0, random_seed=1)
    ds.shuffle()
    for i in range(len(ds)):
        *arrays, pmsk = ds[i]
        x_rna, x_atac = arrays[0], arrays[1]
        both = pmsk[:, 0] & pmsk[:, 1]
        # rna layer holds its own row number (shared cell i is row i); atac holds the shared index
        assert torch.equal(x_rna[both, 0], x_atac[both, 0])
        assert pmsk.any(dim=1).all()
    train, val = ds.random_split([0.9, 0.1], random_state=0)
    assert train.size + val.size == ds.size


def test_anndataset_errors():
    rna, atac, _ = _paired_adatas()
    with pytest.raises(ValueError, match="highly variable"):
        models.configure_dataset(rna, "NB")
    with pytest.raises(ValueError, match="use_layer"):
        models.configure_dataset(rna, "NB", use_highly_variable=False, use_layer="nope")
    with pytest.raises(ValueError, match="use_rep"):
        models.configure_dataset(rna, "NB", use_highly_variable=False, use_rep="nope")
    with pytest.raises(ValueError, match="use_batch"):
        models.configure_dataset(rna, "NB", use_highly_variable=False, use_batch="nope")
    models.configure_dataset(rna, "NB", use_highly_variable=False)
    with pytest.raises(ValueError, match="match the number"):
        AnnDataset([rna, atac], [rna.uns[config.ANNDATA_KEY]])
    with pytest.raises(ValueError, match="mode"):
        AnnDataset([rna], [rna.uns[config.ANNDATA_KEY]], mode="xxx")


def test_nan_sparse_densify():
    x = scipy.sparse.csr_matrix(np.array([[0, 2.0, 0], [1.0, 0, 0]], dtype=np.float32))
    dense = AnnDataset._index_array(x, np.array([1, 0]), nan_sparse=True)
    assert np.isnan(dense).sum() == 4 and dense[0, 0] == 1.0 and dense[1, 1] == 2.0


def test_parallel_loader_cycles_the_short_loader():
    ds1, ds2 = ArrayDataset(np.arange(40).reshape(20, 2)), ArrayDataset(np.arange(6).reshape(3, 2))
    for ds in (ds1, ds2):
        ds.prepare_shuffle(num_workers=0, random_seed=0)
    pl = ParallelDataLoader(DataLoader(ds1, batch_size=4, shuffle=True),
                            DataLoader(ds2, batch_size=2, shuffle=True), cycle_flags=[False, True])
    batches = list(pl)
    assert len(batches) == 5 and all(len(b) == 2 for b in batches)
    with pytest.raises(ValueError):
        ParallelDataLoader(DataLoader(ds1, batch_size=4), cycle_flags=[True, False])


def test_explicit_sampler_equals_shuffle_true():
    """GLUETrainer._loaders builds RandomSampler/BatchSampler by hand; the batch stream must be
    the one DataLoader(shuffle=True, generator=G) yields (what the reference constructs)."""
    def stream(explicit):
        ds = ArrayDataset(np.arange(90).reshape(45, 2), getitem_size=2)
        ds.prepare_shuffle(num_workers=0, random_seed=4)
        gen = torch.Generator().manual_seed(4)
        if explicit:
            bs = torch.utils.data.BatchSampler(
                torch.utils.data.RandomSampler(ds, generator=gen), 4, True)
            dl = DataLoader(ds, reshuffle=True, batch_sampler=bs, generator=gen)
        else:
            dl = DataLoader(ds, batch_size=4, shuffle=True, drop_last=True, generator=gen)
        return [b[0].clone() for _ in range(3) for b in dl]
    a, b = stream(True), stream(False)
    assert len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a, b))


def test_rank_strided_batches_partition_the_stream():
    base = [[i] for i in range(11)]
    parts = [list(RankStridedBatches(base, r, 4)) for r in range(4)]
    assert all(len(p) == 2 for p in parts)  # 11 // 4, tail dropped so ranks stay in lock-step
    flat = sorted(b[0] for p in parts for b in p)
    assert flat == list(range(8))


def test_estimate_balancing_weight_matches_populations():
    """Two datasets with the same three populations in different proportions: after balancing,
    the weighted population masses agree (scglue/data.py:705-836 semantics), weights have mean 1."""
    import pandas as pd
    from glue_b200.anndata_lite import AnnData
    from glue_b200.data import estimate_balancing_weight
    rs = np.random.RandomState(0)
    centers = np.eye(3, 8) * 5.0
    def make(counts):
        lab = np.repeat(np.arange(3), counts)
        rep = centers[lab] + rs.randn(lab.size, 8) * 0.05
        a = AnnData(X=np.zeros((lab.size, 1), dtype=np.float32), obs=pd.DataFrame(index=[f"c{i}" for i in range(lab.size)]),
                    var=pd.DataFrame(index=["f0"]), obsm={"X_emb": rep.astype(np.float32)})
        return a, lab
    a, la = make([300, 100, 50])
    b, lb = make([60, 200, 200])
    estimate_balancing_weight(a, b, use_rep="X_emb", resolution=0.2)
    wa, wb = np.asarray(a.obs["balancing_weight"]), np.asarray(b.obs["balancing_weight"])
    assert abs(wa.mean() - 1) < 1e-6 and abs(wb.mean() - 1) < 1e-6
    ma = np.array([wa[la == k].sum() for k in range(3)]) / wa.sum()
    mb = np.array([wb[lb == k].sum() for k in range(3)]) / wb.sum()
    assert np.abs(ma - mb).max() < 0.05, (ma, mb)


def test_device_prefetcher_cpu_passthrough_and_order():
    from glue_b200.models.data import DevicePrefetcher
    batches = [[torch.full((2,), float(i)), torch.arange(3) + i] for i in range(7)]
    out = [[t.clone() for t in b] for b in DevicePrefetcher(iter(batches), torch.device("cpu"), depth=3)]
    assert [int(b[0][0]) for b in out] == list(range(7))
    assert all(torch.equal(o[1], b[1]) for o, b in zip(out, batches))


def test_flat_rmsprop_refuses_cpu_parameters():
    import pytest
    from glue_b200.optim import FlatRMSprop
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FlatRMSprop([torch.nn.Parameter(torch.zeros(3))], lr=1e-3)


def test_loss_terms_as_matrix_vector_product():
    """SCGLUETrainer._combine: every reported term is a fixed linear combination of scalar components;
    values and gradients must equal the term-by-term arithmetic of scglue.py:349-376."""
    import types
    from glue_b200.models.scglue import SCGLUETrainer
    stub = types.SimpleNamespace(net=types.SimpleNamespace(device=torch.device("cpu")))
    gen = torch.Generator().manual_seed(0)
    comps = {n: torch.randn((), generator=gen).requires_grad_(True)
             for n in ("dsc", "g_nll", "g_kl", "sup", "x_rna_nll", "x_rna_kl", "x_atac_nll", "x_atac_kl", "cos")}
    lam_kl, lam_data, lam_graph, lam_align, lam_cos, mw = 0.04, 1.0, 0.02, 0.05, 0.02, {"rna": 1.0, "atac": 0.7}
    rows = {"dsc_loss": {"dsc": 1.0}, "g_elbo": {"g_nll": 1.0, "g_kl": lam_kl},
            "x_rna_elbo": {"x_rna_nll": 1.0, "x_rna_kl": lam_kl}, "cos_loss": {"cos": 1.0}}
    vae = {"g_nll": lam_graph * 2, "g_kl": lam_graph * 2 * lam_kl, "cos": lam_cos}
    for k in ("rna", "atac"):
        vae[f"x_{k}_nll"], vae[f"x_{k}_kl"] = lam_data * mw[k], lam_data * mw[k] * lam_kl
    rows["vae_loss"], rows["gen_loss"] = vae, dict(vae, dsc=-lam_align)
    out = SCGLUETrainer._combine(stub, comps, rows)
    c = comps
    x_elbo = {k: c[f"x_{k}_nll"] + lam_kl * c[f"x_{k}_kl"] for k in ("rna", "atac")}
    g_elbo = c["g_nll"] + lam_kl * c["g_kl"]
    vae_ref = lam_data * sum(mw[k] * x_elbo[k] for k in mw) + lam_graph * 2 * g_elbo + lam_cos * c["cos"]
    gen_ref = vae_ref - lam_align * c["dsc"]
    assert torch.allclose(out["g_elbo"], g_elbo) and torch.allclose(out["x_rna_elbo"], x_elbo["rna"])
    assert torch.allclose(out["vae_loss"], vae_ref, atol=1e-6) and torch.allclose(out["gen_loss"], gen_ref, atol=1e-6)
    g = torch.autograd.grad(out["gen_loss"], list(comps.values()), allow_unused=True)
    g_ref = torch.autograd.grad(gen_ref, list(comps.values()), allow_unused=True)
    for name, a, b in zip(comps, g, g_ref):
        a = torch.zeros(()) if a is None else a
        b = torch.zeros(()) if b is None else b
        assert torch.allclose(a, b, atol=1e-7), name
    # the coefficient matrix is cached per set of weights
    assert len(stub._combine_cache) == 1
    SCGLUETrainer._combine(stub, comps, rows)
    assert len(stub._combine_cache) == 1


def test_paired_mean_and_cosine_reference_path_on_cpu():
    """PairedSCGLUETrainer._umean without the fused kernel (CPU tensors): the formulas of
    scglue.py:618-622 / 654-660, kept as the specification the CUDA kernel is tested against."""
    import types
    import torch.nn.functional as F
    from glue_b200.models.scglue import PairedSCGLUETrainer
    gen = torch.Generator().manual_seed(1)
    usamp = {"rna": torch.randn(6, 4, generator=gen), "atac": torch.randn(6, 4, generator=gen)}
    pmsk = torch.tensor([[1, 1], [1, 0], [0, 1], [1, 1], [1, 1], [0, 1]], dtype=torch.bool)
    stub = types.SimpleNamespace(net=types.SimpleNamespace(keys=["rna", "atac"]), _pmsk=pmsk, normalize_u=False,
                                 lam_cos=0.02)
    ustack, umean, pf, cos = PairedSCGLUETrainer._umean(stub, usamp)
    assert torch.allclose(umean[1], usamp["rna"][1]) and torch.allclose(umean[2], usamp["atac"][2])
    assert torch.allclose(umean[0], 0.5 * (usamp["rna"][0] + usamp["atac"][0]))
    want = sum(1 - (F.cosine_similarity(usamp[k], umean) * pf[i]).sum() / pf[i].sum()
               for i, k in enumerate(["rna", "atac"]))
    assert torch.allclose(cos, want)


@pytest.mark.parametrize("n,size,dtype", [(52000, 200000, np.float32), (152000, 104321, np.float32),
                                          (7, 5000, np.float64), (1000, 4096, np.float64), (300, 100, np.float32)])
def test_cdf_sampler_reproduces_randomstate_choice(n, size, dtype):
    """_CdfSampler.choice must return what RandomState.choice(n, size, p=p) returns, draw for draw,
    and leave the generator in the same state (the bucket table only short-cuts the binary search)."""
    from glue_b200.models.data import _CdfSampler
    w = np.random.RandomState(n).rand(n).astype(dtype)
    w[:: max(2, n // 11)] = 0          # zero-probability entries
    w[1] = w.sum() * 3                 # one dominant entry (many draws in one cdf step)
    p = (w / w.sum()).astype(dtype)
    sampler = _CdfSampler(p)
    for seed in (0, 1):
        a, b = np.random.RandomState(seed), np.random.RandomState(seed)
        want = a.choice(n, size, replace=True, p=p)
        got = sampler.choice(b, size)
        assert np.array_equal(got, want)
        assert np.array_equal(a.random_sample(3), b.random_sample(3)), "generator state must match"


def test_key_filter_has_no_false_negatives():
    from glue_b200.models.data import _KeyFilter
    rs = np.random.RandomState(3)
    keys = np.unique(rs.randint(0, 2 ** 40, 150000).astype(np.int64))
    filt = _KeyFilter(keys)
    assert filt.maybe(keys).all()
    probe = rs.randint(0, 2 ** 40, 500000).astype(np.int64)
    maybe = filt.maybe(probe)
    assert maybe[np.isin(probe, keys)].all()
    assert maybe.mean() < 0.05, "the filter should reject almost every non-member"


def test_negative_sampler_large_graph_matches_set_based_reference():
    """The vectorised sampler against a literal restatement of scglue/models/data.py:794-820
    (Python set of (i, j, s) tuples) on a graph large enough to take the table / filter paths."""
    import networkx as nx
    from glue_b200.models.data import GraphDataset
    rs0 = np.random.RandomState(11)
    n_v, n_e = 3000, 9000
    g = nx.Graph()
    names = [f"v{i}" for i in range(n_v)]
    g.add_nodes_from(names)
    for s, t, w in zip(rs0.randint(0, n_v, n_e), rs0.randint(0, n_v, n_e), rs0.uniform(0.05, 1, n_e)):
        g.add_edge(names[s], names[t], weight=float(w), sign=1 if (s + t) % 5 else -1)
    g.add_edges_from((v, v, {"weight": 1.0, "sign": 1}) for v in names)
    ds = GraphDataset(g, names, neg_samples=10, weighted_sampling=True, deemphasize_loops=True)
    assert ds.effective_enum * 10 >= 4096, "large enough for the bucket-table path"
    got_idx, got_wt, got_sgn = ds.propose_shuffle(5)

    rs = np.random.RandomState(5)
    eset = {(i, j, s) for (i, j), s in zip(ds.eidx.T.tolist(), ds.esgn.tolist())}
    pi, pj, pw, ps = ds.eidx[0], ds.eidx[1], ds.ewt, ds.esgn
    psamp = rs.choice(ds.ewt.size, ds.effective_enum, replace=True, p=ds.eprob)
    pi_, pj_, pw_, ps_ = pi[psamp], pj[psamp], pw[psamp], ps[psamp]
    pw_ = np.ones_like(pw_)
    ni_ = np.tile(pi_, ds.neg_samples)
    nw_ = np.zeros(pw_.size * ds.neg_samples, dtype=pw_.dtype)
    ns_ = np.tile(ps_, ds.neg_samples)
    nj_ = rs.choice(ds.vnum, pj_.size * ds.neg_samples, replace=True, p=ds.vprob)
    remain = np.where([item in eset for item in zip(ni_.tolist(), nj_.tolist(), ns_.tolist())])[0]
    while remain.size:
        newnj = rs.choice(ds.vnum, remain.size, replace=True, p=ds.vprob)
        nj_[remain] = newnj
        remain = remain[[item in eset for item in zip(ni_[remain].tolist(), newnj.tolist(), ns_[remain].tolist())]]
    idx = np.stack([np.concatenate([pi_, ni_]), np.concatenate([pj_, nj_])])
    w, s = np.concatenate([pw_, nw_]), np.concatenate([ps_, ns_])
    perm = rs.permutation(idx.shape[1])
    assert np.array_equal(got_idx, idx[:, perm])
    assert np.array_equal(got_wt, w[perm]) and np.array_equal(got_sgn, s[perm])


def test_shuffle_lookahead_depth_does_not_change_the_stream():
    """Shuffle n of a dataset uses seed random_seed + n whatever the number of look-ahead threads."""
    from glue_b200.models.data import ArrayDataset
    base = np.arange(1000)
    streams = []
    for workers in (0, 1, 3):
        ds = ArrayDataset(base, getitem_size=10)
        ds.prepare_shuffle(num_workers=workers, random_seed=7)
        seq = []
        for _ in range(5):
            ds.shuffle()
            seq.append(ds.shuffle_idx[0].copy())
        assert len(ds._lookahead) == workers
        ds.clean()
        assert not ds.has_workers
        streams.append(seq)
    for seq in streams[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(seq, streams[0]))
    assert np.array_equal(streams[0][2], np.random.RandomState(9).permutation(base))


@pytest.mark.parametrize("nan_sparse", [False, True])
def test_device_csr_rows_equal_scipy_rows(nan_sparse):
    """DeviceCSR.index_select(0, rows) == csr[rows] densified (NaN for structural zeros of
    nan_sparse modalities), including empty rows, repeated rows and a stored explicit zero."""
    import scipy.sparse
    from glue_b200.models.data import AnnDataset, DeviceCSR
    rs = np.random.RandomState(0)
    dense = rs.poisson(0.2, (50, 37)).astype(np.float32)
    dense[7] = 0          # empty rows
    dense[49] = 0
    dense[3] = rs.poisson(3, 37) + 1  # a full row (defines max_row_nnz)
    csr = scipy.sparse.csr_matrix(dense)
    csr.data[0] = 0.0     # an explicitly stored zero stays a value, not a structural zero
    dev = DeviceCSR(csr, torch.device("cpu"), fill=float("nan") if nan_sparse else 0.0)
    rows = np.array([3, 7, 0, 49, 3, 12, 48, 7])
    got = dev.index_select(0, torch.as_tensor(rows)).numpy()
    want = AnnDataset._index_array(csr, rows, nan_sparse)
    assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True)
    assert dev.index_select(0, torch.as_tensor(np.array([], dtype=np.int64))).shape == (0, 37)
    empty = DeviceCSR(scipy.sparse.csr_matrix((5, 9), dtype=np.float32), torch.device("cpu"))
    assert np.array_equal(empty.index_select(0, torch.tensor([1, 4])).numpy(), np.zeros((2, 9), np.float32))


def test_anndataset_keeps_csr_on_device_when_dense_does_not_fit():
    """to_device falls back from dense to CSR residency under a tight budget; fetched minibatches
    are the same either way."""
    from tests.synth import make_synthetic
    from glue_b200 import models
    from glue_b200.models.data import AnnDataset, DeviceCSR
    from glue_b200.utils import config
    adatas, graph = make_synthetic(n_cells=(120, 120), n_feat=(40, 300), n_pairs=100, seed=3, rep_dim=8)
    for a in adatas.values():
        models.configure_dataset(a, "NB", use_highly_variable=False, use_rep="X_rep")
    cfgs = [a.uns[config.ANNDATA_KEY] for a in adatas.values()]
    dense_ds = AnnDataset(list(adatas.values()), cfgs, mode="train").to_device("cpu")
    old = config.DEVICE_DATA_BUDGET_BYTES
    try:
        config.DEVICE_DATA_BUDGET_BYTES = 120 * 340 * 4 // 2   # half of the dense count matrices
        sparse_ds = AnnDataset(list(adatas.values()), cfgs, mode="train").to_device("cpu")
    finally:
        config.DEVICE_DATA_BUDGET_BYTES = old
    assert any(isinstance(d, DeviceCSR) for group in sparse_ds.device_data for d in group)
    rows = torch.as_tensor(np.random.RandomState(1).randint(0, 120, (16, 2)))
    pmsk = torch.ones(16, 2, dtype=torch.bool)
    for a, b in zip(dense_ds.fetch_device(rows, pmsk), sparse_ds.fetch_device(rows, pmsk)):
        assert torch.equal(a, b)


def test_engine_averages_loss_vectors_like_dicts():
    """Engine.run: steps that return a LossVector (one device vector) give the same epoch metrics as
    steps that return a dict of scalars."""
    from glue_b200.models.base import Engine, LossVector
    names = ["dsc_loss", "vae_loss", "gen_loss", "g_nll"]
    vals = [torch.tensor([1.0, 2.0, 3.0, 4.0]) * (i + 1) for i in range(5)]
    as_vec = Engine(lambda eng, batch: LossVector(names, batch), tracked=["vae_loss", "g_nll"])
    as_dict = Engine(lambda eng, batch: {n: batch[i] for i, n in enumerate(names)}, tracked=["vae_loss", "g_nll"])
    m_vec = as_vec.run(vals, max_epochs=1).metrics
    m_dict = as_dict.run(vals, max_epochs=1).metrics
    assert m_vec == m_dict == {"vae_loss": 6.0, "g_nll": 12.0}
    lv = LossVector(names, vals[0].clone().requires_grad_(True))
    assert set(lv) == set(names) and len(lv) == 4 and float(lv["gen_loss"]) == 3.0
    assert not lv.detach()["gen_loss"].requires_grad and dict(lv.items()).keys() == set(names)


def test_split_datasets_own_their_shuffle_state():
    """AnnDataset.random_split -> prepare_shuffle(num_workers=2) -> shuffle -> clean on the subsets
    (the sequence GLUETrainer.fit runs, scglue/models/glue.py:561-602): a split carries its own
    look-ahead table and seed, and cleaning one split leaves the others alone."""
    rna, atac, _ = _paired_adatas()
    for a in (rna, atac):
        models.configure_dataset(a, "NB", use_highly_variable=False, use_obs_names=True)
    ds = AnnDataset([rna, atac], [rna.uns[config.ANNDATA_KEY], atac.uns[config.ANNDATA_KEY]],
                    mode="train", getitem_size=8)
    ds.prepare_shuffle(num_workers=1, random_seed=0)      # the parent has threads in flight
    train, val = ds.random_split([0.9, 0.1], random_state=0)
    assert train._lookahead == {} and val._lookahead == {} and train._lookahead is not val._lookahead
    assert train.shuffle_seed is None and not train.has_workers
    with pytest.raises(RuntimeError, match="prepare_shuffle"):
        train.shuffle()
    for sub in (train, val):
        sub.prepare_shuffle(num_workers=2, random_seed=3)
        assert len(sub._lookahead) == 2
    rows = []
    for _ in range(3):
        train.shuffle()
        rows.append(train.shuffle_idx.copy())
        assert sorted(train.shuffle_idx[:, 0][train.shuffle_pmsk[:, 0]]) == \
            sorted(train._get_idx_pmsk(train.view_idx)[0][:, 0][train._get_idx_pmsk(train.view_idx)[1][:, 0]])
    assert not np.array_equal(rows[0], rows[1])
    val.shuffle()
    train.clean()
    assert not train.has_workers and val.has_workers and ds.has_workers
    val.clean(), ds.clean()
    # the same seeds without look-ahead give the same stream
    train2, _ = ds.random_split([0.9, 0.1], random_state=0)
    train2.prepare_shuffle(num_workers=0, random_seed=3)
    for want in rows:
        train2.shuffle()
        assert np.array_equal(train2.shuffle_idx, want)
    # a dataset whose table was lost (older pickles) still cleans
    train2._lookahead = None
    train2.clean()
    assert train2._lookahead == {}


def _stub_fit_model(monkeypatch, paired=True, n_batches=0):
    """A (Paired)SCGLUEModel on the CPU whose trainer returns constant losses instead of launching
    kernels: everything around the step -- split, shuffles, loaders, engine, plugins -- is real."""
    from tests.synth import make_synthetic
    from glue_b200.models import scglue as gs
    from glue_b200.models.base import LossVector
    monkeypatch.setattr(config, "CPU_ONLY", True)
    adatas, graph = make_synthetic(n_cells=(96, 80), n_feat=(12, 30), n_pairs=40, seed=1, rep_dim=6,
                                   paired=paired, n_batches=n_batches)
    for a in adatas.values():
        models.configure_dataset(a, "NB", use_highly_variable=False, use_rep="X_rep",
                                 use_obs_names=paired, use_batch="batch" if n_batches else None)
    base = gs.PairedSCGLUETrainer if paired else gs.SCGLUETrainer

    class StubTrainer(base):
        seen = []

        def _out(self, engine, data, kind):
            n_mod = len(self.net.keys)
            x = data[:n_mod]
            eidx = data[-3]
            StubTrainer.seen.append((kind, engine.state.epoch, tuple(t.shape[0] for t in x), eidx.shape[1]))
            assert all(t.shape[1] == f for t, f in zip(x, (12, 30)))
            vals = torch.arange(len(self.required_losses), dtype=torch.float32) + 1.0 / engine.state.epoch
            return LossVector(self.required_losses, vals)

        def train_step(self, engine, data):
            return self._out(engine, data, "train")

        def val_step(self, engine, data):
            return self._out(engine, data, "val")

    cls = gs.PairedSCGLUEModel if paired else gs.SCGLUEModel
    monkeypatch.setattr(cls, "TRAINER_TYPE", StubTrainer)
    vertices = sorted(graph.nodes)
    model = cls(adatas, vertices, latent_dim=4, h_dim=8, random_seed=0)
    model.compile()
    return model, adatas, graph, StubTrainer


@pytest.mark.parametrize("paired", [True, False])
def test_fit_plumbing_runs_on_cpu_with_a_stub_step(monkeypatch, tmp_path, paired):
    """GLUETrainer.fit / Trainer.fit without a GPU: 90/10 split, block-shuffled loaders, the cycling
    graph loader, per-epoch validation, LR scheduling / early stopping hooks, checkpoint directory,
    clean-up of the shuffle threads (scglue/models/glue.py:486-667, base.py:117-214)."""
    model, adatas, graph, stub = _stub_fit_model(monkeypatch, paired=paired)
    stub.seen.clear()
    model.fit(adatas, graph, data_batch_size=16, max_epochs=3, patience=2, reduce_lr_patience=1,
              directory=tmp_path)
    kinds = [k for k, *_ in stub.seen]
    assert kinds.count("train") > 0 and kinds.count("val") > 0
    epochs = sorted({e for k, e, *_ in stub.seen if k == "train"})
    assert epochs == [1, 2, 3]
    n_train = sum(1 for k, e, *_ in stub.seen if k == "train" and e == 1)
    n_view = 96 if paired else 96 + 80
    assert n_train == (round(n_view * 0.9) + 3) // 4 // 4   # blocks of 4 rows, 4 blocks per minibatch, drop_last
    # (the one short block of the view can land in any minibatch)
    assert all(b[0] == b[1] and 12 < b[0] <= 16 for k, e, b, _ in stub.seen if k == "train")
    assert model.trainer.eidx is None and model.trainer.align_burnin is None   # cleared after fit
    # a second fit on the same model works (datasets are rebuilt, threads were joined)
    model.fit(adatas, graph, data_batch_size=16, max_epochs=1, patience=None, reduce_lr_patience=None,
              directory=tmp_path)
    # get_losses runs the validation step over one pass
    losses = model.get_losses(adatas, graph, data_batch_size=32)
    assert set(losses) == set(model.trainer.required_losses)


def test_packed_sparse_host_batches_equal_dense_batches():
    """Host-resident datasets hand sparse count rows to the device as (position, value) pairs
    (``SparseRows``); unpacked they must be the dense minibatch of ``AnnDataset.__getitem__``
    (scglue/models/data.py:397-436) bit for bit, including all-zero batches and padded buckets."""
    from tests.synth import make_synthetic
    from glue_b200.models.data import SparseRows, is_packed_sparse, unpack_sparse_rows
    adatas, graph = make_synthetic(n_cells=(150, 150), n_feat=(40, 700), n_pairs=100, seed=3, rep_dim=8)
    for a in adatas.values():
        models.configure_dataset(a, "NB", use_highly_variable=False, use_rep="X_rep")
    cfgs = [a.uns[config.ANNDATA_KEY] for a in adatas.values()]

    def batches(sparse):
        ds = AnnDataset(list(adatas.values()), cfgs, mode="train", getitem_size=8)
        ds.sparse_items = sparse
        ds.prepare_shuffle(num_workers=0, random_seed=2)
        return list(DataLoader(ds, batch_size=4, shuffle=True, generator=torch.Generator().manual_seed(2)))

    dense, packed = batches(False), batches(True)
    assert len(dense) == len(packed)
    seen_packed = 0
    for d, p in zip(dense, packed):
        for i, (td, tp) in enumerate(zip(d, p)):
            if is_packed_sparse(tp) and i < 2:
                seen_packed += 1
                assert tp.shape[1] % SparseRows.BUCKET == 0
                got = unpack_sparse_rows(tp, td.shape[0], td.shape[1])
                assert got.dtype == td.dtype and torch.equal(got, td)
            else:
                assert torch.equal(td, tp)
    assert seen_packed == 2 * len(dense)
    # an all-zero block
    zero = SparseRows(scipy.sparse.csr_matrix((5, 9), dtype=np.float32))
    assert torch.equal(unpack_sparse_rows(SparseRows.pack([zero, zero]), 10, 9), torch.zeros(10, 9))
    # direct indexing keeps the reference's contract: dense tensors
    ds = AnnDataset(list(adatas.values()), cfgs, mode="train", getitem_size=8)
    assert all(isinstance(t, torch.Tensor) and not is_packed_sparse(t) for t in ds[0])


def test_balancing_arithmetic_matches_reference_restatement_given_clusters():
    """The cluster-matching half of estimate_balancing_weight (scglue/data.py:793-836) against
    oracle/balancing.py, for two and for three datasets, with the clusters given (``use_clusters``: what a
    scanpy user would get from leiden); the clustering half is a documented divergence (glue_b200/data.py)."""
    import pandas as pd
    from glue_b200.anndata_lite import AnnData
    from glue_b200.data import estimate_balancing_weight
    from oracle.balancing import balancing_weights
    rs = np.random.RandomState(1)
    centers = rs.randn(5, 10) * 3
    def make(counts, seed):
        r = np.random.RandomState(seed)
        lab = np.repeat(np.arange(len(counts)), counts)
        rep = centers[lab % 5] + r.randn(lab.size, 10) * 0.4
        order = r.permutation(lab.size)
        lab, rep = lab[order], rep[order]
        a = AnnData(X=np.zeros((lab.size, 1), dtype=np.float32), obs=pd.DataFrame({"cl": lab.astype(str)},
                    index=[f"c{i}" for i in range(lab.size)]), var=pd.DataFrame(index=["f0"]),
                    obsm={"X_emb": rep.astype(np.float32)})
        return a, lab
    sets = [make([40, 25, 60, 10], 2), make([15, 50, 30, 30, 20], 3), make([33, 33, 33], 4)]
    for group in (sets[:2], sets):
        adatas = [a for a, _ in group]
        estimate_balancing_weight(*adatas, use_rep="X_emb", use_clusters="cl", cutoff=0.5, power=4.0)
        # pd.factorize(sort=True) orders the string labels "0", "1", ... like the integers 0 .. 4
        want = balancing_weights([np.asarray(a.obsm["X_emb"], dtype=float) for a in adatas], [lab for _, lab in group])
        for a, w in zip(adatas, want):
            got = np.asarray(a.obs["balancing_weight"], dtype=float)
            assert np.allclose(got, w, rtol=1e-6, atol=1e-9)
            assert abs(got.mean() - 1) < 1e-9
    with pytest.raises(ValueError, match="cannot be found"):
        estimate_balancing_weight(*[a for a, _ in sets[:2]], use_rep="X_emb", use_clusters="nope")


def test_fast_loader_iteration_equals_torch_iterator(monkeypatch):
    """config.FAST_LOADER iterates the (torch) samplers directly; the batch stream over several epochs --
    including the draw torch's iterator takes from the generator for its base seed -- is the one torch's
    DataLoader iterator yields, for the constructor forms the trainer and the tests use."""
    from glue_b200.utils import config

    def stream(fast, explicit):
        monkeypatch.setattr(config, "FAST_LOADER", fast)
        ds = ArrayDataset(np.arange(94).reshape(47, 2), getitem_size=2)
        ds.prepare_shuffle(num_workers=0, random_seed=7)
        gen = torch.Generator().manual_seed(7)
        if explicit:
            bs = torch.utils.data.BatchSampler(torch.utils.data.RandomSampler(ds, generator=gen), 4, False)
            dl = DataLoader(ds, reshuffle=True, batch_sampler=bs, generator=gen)
        else:
            dl = DataLoader(ds, batch_size=4, shuffle=True, drop_last=True, generator=gen)
        out = [b[0].clone() for _ in range(3) for b in dl]
        return out, gen.get_state()

    for explicit in (True, False):
        (a, sa), (b, sb) = stream(True, explicit), stream(False, explicit)
        assert len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a, b))
        assert torch.equal(sa, sb), "the generator must end in the same state"


@pytest.mark.parametrize("n_v,n_e,neg,weighted", [(40, 500, 10, True), (40, 500, 1, False), (3000, 9000, 10, True),
                                                  (25, 200, 0, True)])
def test_native_negative_sampler_equals_numpy_restatement(n_v, n_e, neg, weighted):
    """``glue_sample_graph_epoch`` (one native call per graph epoch, the default) against the numpy restatement of
    scglue/models/data.py:794-822, bit for bit: dense little graphs (a third of all vertex pairs are edges, so
    the redraw loop runs several rounds), a graph large enough for the bucket tables, no negatives at all; the
    same arrays from concurrent look-ahead threads."""
    import threading
    import networkx as nx
    from glue_b200.utils import config
    rs0 = np.random.RandomState(n_v + neg)
    g = nx.Graph()
    names = [f"v{i}" for i in range(n_v)]
    g.add_nodes_from(names)
    for s, t, w in zip(rs0.randint(0, n_v, n_e), rs0.randint(0, n_v, n_e), rs0.uniform(0.05, 1, n_e)):
        g.add_edge(names[s], names[t], weight=float(w), sign=1 if (s + t) % 3 else -1)
    g.add_edges_from((v, v, {"weight": 1.0, "sign": 1}) for v in names)
    ds = GraphDataset(g, names, neg_samples=neg, weighted_sampling=weighted)
    assert config.NATIVE_SAMPLER
    seeds = (0, 1, 77, 2 ** 32 - 1)
    want = {seed: ds._propose_shuffle_numpy(seed) for seed in seeds}
    for seed in seeds:
        got = ds.propose_shuffle(seed)
        for a, b in zip(got, want[seed]):
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
    got = {}
    threads = [threading.Thread(target=lambda s=seed: got.__setitem__(s, ds.propose_shuffle(s))) for seed in seeds]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert all(np.array_equal(a, b) for seed in seeds for a, b in zip(got[seed], want[seed]))
    # the look-ahead machinery on top of it
    ds.prepare_shuffle(num_workers=2, random_seed=76)
    ds.shuffle(), ds.shuffle()
    assert np.array_equal(ds.samp_eidx.numpy(), want[77][0]) and np.array_equal(ds.samp_ewt.numpy(), want[77][1])
    ds.clean()


def test_native_sampler_argument_errors():
    from glue_b200 import _lib
    lib = _lib.lib()
    a = np.zeros(4, dtype=np.int64)
    assert lib.glue_sample_graph_epoch(0, None, a.ctypes.data, a.ctypes.data, 4, a.ctypes.data, a.ctypes.data, 10,
                                       a.ctypes.data, a.ctypes.data, 10, 4, a.ctypes.data, 4, a.ctypes.data, 12, 1, 1,
                                       a.ctypes.data, a.ctypes.data, a.ctypes.data, a.ctypes.data) == -1
    assert lib.glue_sample_graph_epoch(0, a.ctypes.data, a.ctypes.data, a.ctypes.data, 4, a.ctypes.data,
                                       a.ctypes.data, 10, a.ctypes.data, a.ctypes.data, 10, 4, a.ctypes.data, 4,
                                       a.ctypes.data, 12, 2 ** 31, 3, a.ctypes.data, a.ctypes.data, a.ctypes.data,
                                       a.ctypes.data) == -2


@pytest.mark.parametrize("paired", [True, False])
def test_anndataset_shuffle_with_cached_indexers_equals_the_literal_form(paired):
    """``propose_shuffle`` permutes cached row numbers instead of looking the permuted cell ids up again; the
    literal form of scglue/models/data.py:617-620 (permute the ids, ``get_indexer``, random fill on the same
    RandomState) gives the same rows and masks, on the full view and on a split."""
    from tests.synth import make_synthetic
    from glue_b200.utils import get_rs
    adatas, _ = make_synthetic(n_cells=(96, 80), n_feat=(12, 30), n_pairs=40, seed=1, rep_dim=6, paired=paired)
    for a in adatas.values():
        models.configure_dataset(a, "NB", use_highly_variable=False, use_rep="X_rep", use_obs_names=paired)
    ds = AnnDataset(list(adatas.values()), [a.uns[config.ANNDATA_KEY] for a in adatas.values()], mode="train",
                    getitem_size=8)
    train, val = ds.random_split([0.8, 0.2], random_state=3)
    for sub in (ds, train, val):
        for seed in (0, 9):
            rs = get_rs(seed)
            want = sub._get_idx_pmsk(rs.permutation(sub.view_idx), random_fill=True, random_state=rs)
            got = sub.propose_shuffle(seed)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert train._view_indexers()[0].size == train.size and ds._view_indexers()[0].size == ds.size


@pytest.mark.gpu
def test_device_graph_dataset_streams_epochs_through_pooled_pinned_slabs():
    """A device-resident GraphDataset: the native sampler writes each graph epoch into a pinned slab on a
    look-ahead thread, ``accept_shuffle`` uploads it asynchronously and the slab returns to the process-wide
    pool once the copy has run.  The edges on the device are the numpy restatement's, epoch after epoch, and
    slabs are recycled instead of pinned anew."""
    import networkx as nx
    from glue_b200.models.data import _PinnedSlabs
    rs0 = np.random.RandomState(4)
    n_v, n_e = 3000, 9000
    g = nx.Graph()
    names = [f"v{i}" for i in range(n_v)]
    g.add_nodes_from(names)
    for s, t, w in zip(rs0.randint(0, n_v, n_e), rs0.randint(0, n_v, n_e), rs0.uniform(0.05, 1, n_e)):
        g.add_edge(names[s], names[t], weight=float(w), sign=1 if (s + t) % 5 else -1)
    g.add_edges_from((v, v, {"weight": 1.0, "sign": 1}) for v in names)
    host = GraphDataset(g, names, neg_samples=10)
    ds = GraphDataset(g, names, neg_samples=10).to_device(torch.device("cuda:0"))
    pinned_before = _PinnedSlabs.allocated
    for fit in range(2):  # two fits' worth: the second one finds its slabs in the pool
        ds.prepare_shuffle(num_workers=2, random_seed=3)
        for n in range(6):
            ds.shuffle()
            want = host._propose_shuffle_numpy(3 + n)
            assert ds.samp_eidx.is_cuda and ds.samp_eidx.dtype == torch.int64
            assert np.array_equal(ds.samp_eidx.cpu().numpy(), want[0])
            assert np.array_equal(ds.samp_ewt.cpu().numpy(), want[1])
            assert np.array_equal(ds.samp_esgn.cpu().numpy(), want[2])
        ds.clean()
        assert not ds.__dict__.get("_slabs") and not ds.__dict__.get("_uploads")
        if fit == 0:
            pinned_first = _PinnedSlabs.allocated - pinned_before
            assert 0 < pinned_first <= 2 + 3, "look-ahead depth plus the uploads in flight"
    assert _PinnedSlabs.allocated - pinned_before == pinned_first, "the second fit pinned nothing"


def test_deferred_metrics_resolve_once_and_validation_is_issued_before_they_are_read():
    """Engine.run keeps the epoch means of LossVector steps on the device until somebody reads them
    (DeferredMetrics: a Mapping that turns into plain floats at the first access), and Trainer.fit registers the
    validation pass ahead of every handler that reads the training metrics."""
    from glue_b200.models.base import DeferredMetrics, Engine, LossVector, Trainer
    names = ["dsc_loss", "vae_loss", "gen_loss"]
    vals = [torch.tensor([1.0, 2.0, 3.0]) * (i + 1) for i in range(4)]
    eng = Engine(lambda e, batch: LossVector(names, batch), tracked=["vae_loss", "gen_loss"])
    metrics = eng.run(vals, max_epochs=1).metrics
    assert isinstance(metrics, DeferredMetrics) and metrics._host is None        # nothing read yet
    assert list(metrics) == ["vae_loss", "gen_loss"] and len(metrics) == 2 and metrics._host is None
    assert metrics["vae_loss"] == 5.0 and metrics._means is None                 # first read: resolved for good
    assert dict(metrics) == {"vae_loss": 5.0, "gen_loss": 7.5} and metrics == {"vae_loss": 5.0, "gen_loss": 7.5}

    order = []

    class Stub(Trainer):
        def __init__(self):
            super().__init__(torch.nn.Linear(1, 1))
            self.required_losses = ["vae_loss"]

        def train_step(self, engine, data):
            return LossVector(["vae_loss"], data)

        def val_step(self, engine, data):
            order.append("val_step")
            return LossVector(["vae_loss"], data)

        def report_metrics(self, train_state, val_state):
            order.append(("report", type(train_state.metrics).__name__, train_state.metrics._host is not None))

    Stub().fit([torch.tensor([2.0])], val_loader=[torch.tensor([4.0])], max_epochs=1)
    # the validation step ran first; by the time of the report the non-finite guard had read the metrics
    assert order == ["val_step", ("report", "DeferredMetrics", True)]


def test_boolean_config_switches_from_the_environment(monkeypatch):
    from glue_b200 import utils
    monkeypatch.setenv("GLUE_B200_FLAGS", "EARLY_ALIGN_BACKWARD=0, FAST_LOADER=1")
    cfg = type(utils.config)()
    assert cfg.EARLY_ALIGN_BACKWARD is False and cfg.FAST_LOADER is True and cfg.NATIVE_SAMPLER is True
    monkeypatch.setenv("GLUE_B200_FLAGS", "TMP_PREFIX=1")
    with pytest.raises(AttributeError, match="not a boolean"):
        type(utils.config)()
