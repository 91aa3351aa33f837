"""Step-level GPU parity: one full train_step (discriminator update + generator update,
RMSprop) of the new trainers against what the UNMODIFIED reference produced for the same
parameters, batch, sampled edges and injected noise (tests/golden/step_*.npz)."""
import numpy as np
import pytest
import torch

from tests.util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


class GoldenNoise:
    """Replays the reference's standard-normal draws (phase 0 = dsc pass, 1 = gen pass)."""

    def __init__(self, g, first_key):
        self.g, self.first_key, self.phase = g, first_key, -1

    def __call__(self, like, tag):
        if tag == f"u_{self.first_key}":
            self.phase += 1
        name = ("dsc", "gen")[self.phase]
        key = {"burnin": "burnin", "v": "v"}.get(tag, tag)
        arr = self.g[f"noise_{name}__{key}"]
        return torch.as_tensor(arr).to(like.device).reshape(like.shape)


def build(golden_npz, paired):
    from glue_b200.models import sc
    from glue_b200.models.scglue import SCGLUE, PairedSCGLUETrainer, SCGLUETrainer
    from glue_b200.utils import config

    g = golden_npz
    init = {k[len("init__"):]: torch.as_tensor(v) for k, v in g.items() if k.startswith("init__")}
    V, K = init["g2v.vrepr"].shape
    if "keys" in g:  # fixtures of the BASELINE-config shapes describe themselves
        keys = [str(k) for k in g["keys"]]
        prob = {k: str(g[f"prob__{k}"]) for k in keys}
        mw = {k: float(g[f"mw__{k}"]) for k in keys}
    else:
        keys = ["rna", "atac"]
        prob = {"rna": "NB", "atac": "ZINB" if paired else "NB"}
        mw = {"rna": 1.0, "atac": 2.0}
    dec = {"NB": sc.NBDataDecoder, "ZINB": sc.ZINBDataDecoder, "ZILN": sc.ZILNDataDecoder,
           "Normal": sc.NormalDataDecoder}
    x2u, u2x, idx = {}, {}, {}
    for k in keys:
        in_dim = init[f"x2u.{k}.linear_0.weight"].shape[1]
        h_dim = init[f"x2u.{k}.linear_0.weight"].shape[0]
        enc = sc.NBDataEncoder if prob[k] in ("NB", "ZINB") else sc.VanillaDataEncoder
        x2u[k] = enc(in_dim, K, h_depth=2, h_dim=h_dim, dropout=0.0)
        nb, F = init[f"u2x.{k}.scale_lin"].shape
        u2x[k] = dec[prob[k]](F, n_batches=nb)
        idx[k] = init[f"{k}_idx"]
    du = sc.Discriminator(K, len(keys), n_batches=0, h_depth=2, h_dim=h_dim, dropout=0.0)
    net = SCGLUE(sc.GraphEncoder(V, K), sc.GraphDecoder(), x2u, u2x, idx, du, sc.Prior())
    net.load_state_dict(init)
    common = dict(lam_data=1.0, lam_kl=1.0, lam_graph=0.02, lam_align=0.05, lam_sup=0.02,
                  dsc_steps=1, normalize_u=False, modality_weight=mw,
                  optim="RMSprop", lr=2e-3)
    if paired:
        trainer = PairedSCGLUETrainer(net, lam_joint_cross=0.02, lam_real_cross=0.02,
                                      lam_cos=0.02, **common)
    else:
        trainer = SCGLUETrainer(net, **common)
    return net, trainer, keys


@pytest.mark.parametrize("paired", [False, True, "cfg3", "cfg4"])
def test_train_step_matches_reference(native_lib, golden, paired):
    """``paired``: False / True = the round-1 fixtures (SCGLUE / PairedSCGLUE, two modalities); "cfg3" = three
    modalities with a ZILN decoder and sign -1 guidance edges, "cfg4" = ZINB decoders with 20 batches."""
    tag = paired if isinstance(paired, str) else None
    paired = paired is True
    g = golden(f"step_{tag}" if tag else ("step_paired" if paired else "step_scglue"))
    graph = golden("graph")
    net, trainer, keys = build(g, paired)
    dev = net.device
    assert dev.type == "cuda"
    trainer.eidx = torch.as_tensor(graph["eidx"]).to(dev)
    trainer.enorm = torch.as_tensor(graph["enorm_keepvar"]).to(dev)
    trainer.esgn = torch.as_tensor(graph["esgn"]).to(dev)
    trainer.align_burnin = int(g["align_burnin"])
    trainer.noise = GoldenNoise(g, keys[0])

    B = g["x__rna"].shape[0]
    t = torch.as_tensor
    data = ([t(g[f"x__{k}"]) for k in keys] + [t(g[f"xrep__{k}"]) for k in keys]
            + [t(g[f"xbch__{k}"]) for k in keys]
            + [torch.full((B,), -1, dtype=torch.int64) for _ in keys]
            + [t(g[f"xdwt__{k}"]) for k in keys]
            + [t(g["pmsk"]), t(g["eidx"]), t(g["ewt"]), t(g["esgn"])])

    class Eng:
        class state:
            epoch = int(g["epoch"])

    losses = trainer.train_step(Eng, data)
    for name, val in losses.items():
        assert_close(val, g[f"loss__{name}"], TOL, f"loss {name}")
    for name, p in net.named_parameters():
        key = f"gengrad__{name}"
        if key in g:
            assert p.grad is not None, name
            assert_close(p.grad, g[key], TOL, f"gradient {name}")
    from tests.util import assert_post_step_close
    for name, val in net.state_dict().items():
        want = g[f"final__{name}"]
        if not val.dtype.is_floating_point:
            assert np.array_equal(val.cpu().numpy(), want), name
            continue
        gref = g.get(f"dscgrad__{name}") if name.startswith("du.") else g.get(f"gengrad__{name}")
        assert_post_step_close(val.detach().cpu().numpy(), want, gref, TOL, f"post-step {name}")


def test_short_fit_runs_and_learns(native_lib):
    """End-to-end: configure_dataset -> SCGLUEModel.fit (device-resident pipeline) ->
    encode_data / encode_graph / decode_data / save+load round trip."""
    import tempfile
    from tests.synth import make_synthetic
    from glue_b200 import models

    adatas, graph = make_synthetic(n_cells=(300, 260), n_feat=(40, 90), n_pairs=150, seed=3,
                                   rep_dim=12)
    for k, a in adatas.items():
        models.configure_dataset(a, "NB", use_highly_variable=False, use_rep="X_rep")
    model = models.SCGLUEModel(adatas, sorted(graph.nodes), latent_dim=16, h_dim=32, random_seed=1)
    model.compile()
    with tempfile.TemporaryDirectory() as tmp:
        model.fit(adatas, graph, data_batch_size=32, graph_batch_size=64, max_epochs=4,
                  align_burnin=2, patience=3, reduce_lr_patience=2, directory=tmp)
        emb = model.encode_data("rna", adatas["rna"])
        vemb = model.encode_graph(graph)
        assert emb.shape == (300, 16) and np.isfinite(emb).all()
        assert vemb.shape == (len(graph.nodes), 16) and np.isfinite(vemb).all()
        dec = model.decode_data("rna", "atac", adatas["rna"], graph)
        assert dec.shape == (300, 90) and np.isfinite(dec).all()
        losses = model.get_losses(adatas, graph)
        assert np.isfinite(list(losses.values())).all()
        model.save(f"{tmp}/m.dill")
        loaded = models.load_model(f"{tmp}/m.dill")
        assert np.allclose(loaded.encode_data("rna", adatas["rna"]), emb, atol=1e-5)


@pytest.mark.parametrize("paired", [False, True])
def test_graph_replayed_steps_equal_eager_steps(native_lib, golden, paired):
    """Six train steps from the same state and inputs, once launched eagerly and once through
    StepGraphs (2 eager warm-up steps, capture, replays): identical kernels on identical data,
    so losses and final parameters must agree bit for bit."""
    from glue_b200.utils import config

    g = golden("step_paired" if paired else "step_scglue")
    graph = golden("graph")
    keys = ["rna", "atac"]
    B = g["x__rna"].shape[0]
    t = lambda a: torch.as_tensor(a).to(DEV)
    data = ([t(g[f"x__{k}"]) for k in keys] + [t(g[f"xrep__{k}"]) for k in keys]
            + [t(g[f"xbch__{k}"]) for k in keys]
            + [torch.full((B,), -1, dtype=torch.int64, device=DEV) for _ in keys]
            + [t(g[f"xdwt__{k}"]) for k in keys]
            + [t(g["pmsk"]), t(g["eidx"]), t(g["ewt"]), t(g["esgn"])])

    class Eng:
        class state:
            epoch = int(g["epoch"])

    def run(use_graphs):
        config.USE_CUDA_GRAPHS = use_graphs
        net, trainer, _ = build(g, paired)
        trainer.eidx, trainer.enorm, trainer.esgn = (t(graph["eidx"]), t(graph["enorm_keepvar"]),
                                                      t(graph["esgn"]))
        trainer.align_burnin = int(g["align_burnin"])
        cache = {}

        def noise(like, tag):  # fixed device-resident draws: the same in every step and run
            key = (tag, tuple(like.shape))
            if key not in cache:
                gen = torch.Generator().manual_seed(abs(hash(key)) % (2 ** 31))
                cache[key] = torch.randn(like.shape, generator=gen).to(like.device)
            return cache[key]

        with torch.random.fork_rng(devices=[0]):
            torch.manual_seed(0)
            trainer.noise = noise
            out = None
            for _ in range(6):
                out = trainer.train_step(Eng, data)
            out = {k: v.detach().clone() for k, v in out.items()}
        torch.cuda.synchronize()
        return net, trainer, out

    try:
        net_e, _, loss_e = run(False)
        net_g, tr_g, loss_g = run(True)
    finally:
        config.USE_CUDA_GRAPHS = True
    info = tr_g.graphs.info()
    assert info["graphs"] == 1 and info["replays"] == 4, info
    assert info["kernels_per_step"] and info["kernels_per_step"] > 10, info
    for name in loss_e:
        assert torch.equal(loss_e[name], loss_g[name]), name
    for (name, a), (_, b) in zip(net_e.state_dict().items(), net_g.state_dict().items()):
        assert torch.equal(a, b), name


def test_fit_scglue_two_stage_runs(native_lib):
    """`fit_SCGLUE` (scglue/models/__init__.py:165-273): pretrain, balancing weights, fine-tune
    from the adopted pretrained model; steps replay from captured graphs in both stages."""
    import tempfile
    from tests.synth import make_synthetic
    from glue_b200 import models

    adatas, graph = make_synthetic(n_cells=(400, 380), n_feat=(60, 150), n_pairs=300, seed=5,
                                   rep_dim=16)
    for a in adatas.values():
        models.configure_dataset(a, "NB", use_highly_variable=False, use_rep="X_rep")
    with tempfile.TemporaryDirectory() as tmp:
        glue = models.fit_SCGLUE(adatas, graph, init_kws=dict(latent_dim=16, h_dim=32, random_seed=2),
                                 fit_kws=dict(data_batch_size=64, graph_batch_size=128, max_epochs=5,
                                              align_burnin=2, patience=4, reduce_lr_patience=2,
                                              directory=tmp))
        assert "balancing_weight" in adatas["rna"].obs.columns
        w = np.asarray(adatas["rna"].obs["balancing_weight"])
        assert np.isfinite(w).all() and abs(w.mean() - 1) < 1e-5
        emb = {k: glue.encode_data(k, a) for k, a in adatas.items()}
        assert all(np.isfinite(e).all() and e.shape[1] == 16 for e in emb.values())
        info = glue.trainer.graphs.info()
        assert info["disabled"] is None and info["replays"] > 0, info


def test_partially_paired_fit_runs(native_lib):
    """PairedSCGLUEModel on partially paired cells (obs_names shared by a subset): the masked
    row-weight formulation of the paired losses trains, with static shapes and graph replay."""
    import tempfile
    from tests.synth import make_synthetic
    from glue_b200 import models

    adatas, graph = make_synthetic(n_cells=(300, 300), n_feat=(40, 90), n_pairs=200, seed=8,
                                   rep_dim=12, paired=True)
    # only the first 120 cells are shared between the modalities
    atac = adatas["atac"]
    atac.obs.index = [f"cell{i}" if i < 120 else f"atac_only{i}" for i in range(300)]
    for a in adatas.values():
        models.configure_dataset(a, "NB", use_highly_variable=False, use_rep="X_rep",
                                 use_obs_names=True)
    model = models.PairedSCGLUEModel(adatas, sorted(graph.nodes), latent_dim=16, h_dim=32,
                                     random_seed=3)
    model.compile()
    with tempfile.TemporaryDirectory() as tmp:
        model.fit(adatas, graph, data_batch_size=64, graph_batch_size=96, max_epochs=3,
                  align_burnin=1, patience=3, reduce_lr_patience=2, directory=tmp)
        losses = model.get_losses(adatas, graph)
    assert np.isfinite(list(losses.values())).all()
    for key in ("joint_cross_loss", "real_cross_loss", "cos_loss"):
        assert key in losses
    assert model.trainer.graphs.info()["disabled"] is None


def test_training_continues_correctly_after_save(native_lib):
    """Model.save moves the network to the CPU and back, which re-allocates every parameter:
    captured steps must be dropped and the fused optimizer must re-adopt the new storage, so
    that a second fit keeps training the network's own parameters."""
    import tempfile
    from tests.synth import make_synthetic
    from glue_b200 import models

    adatas, graph = make_synthetic(n_cells=(260, 240), n_feat=(40, 70), n_pairs=120, seed=11, rep_dim=12)
    for a in adatas.values():
        models.configure_dataset(a, "NB", use_highly_variable=False, use_rep="X_rep")
    model = models.SCGLUEModel(adatas, sorted(graph.nodes), latent_dim=16, h_dim=32, random_seed=4)
    model.compile()
    kw = dict(data_batch_size=32, graph_batch_size=64, max_epochs=3, align_burnin=1, patience=None,
              reduce_lr_patience=None)
    with tempfile.TemporaryDirectory() as tmp:
        model.fit(adatas, graph, directory=tmp, **kw)
        assert model.trainer.graphs.info()["replays"] > 0
        model.save(f"{tmp}/m.dill")
        before = {k: v.detach().clone() for k, v in model.net.state_dict().items()}
        replays = model.trainer.graphs.replays
        model.fit(adatas, graph, directory=tmp, **kw)
    after = model.net.state_dict()
    moved = [k for k in before if before[k].dtype.is_floating_point and not torch.equal(before[k], after[k])]
    assert "g2v.vrepr" in moved and "du.pred.weight" in moved, "second fit must update the parameters"
    assert model.trainer.graphs.replays > replays
    opt = model.trainer.vae_optim
    assert all(p.data_ptr() == v.data_ptr() for p, v in zip(opt._plist, opt._pviews))


def test_repeated_fits_with_step_graphs_equal_eager_fits(native_lib):
    """Round-2 regression: a minibatch signature seen exactly once in the first ``fit`` (the short last
    block) is captured at its first occurrence in the second ``fit``; the CSR of the guidance graph used to
    be built -- and cached process-wide -- inside that capture, i.e. in the graph's private memory pool, and
    the next eager consumer (validation) read garbage.  Two fits with captured steps must give what two
    fits with eager steps give (dropout 0: same arithmetic, same order), and stay finite."""
    from tests.synth import make_synthetic
    from glue_b200 import models
    from glue_b200.utils import config

    def run(use_graphs):
        adatas, graph = make_synthetic(n_cells=(330, 290), n_feat=(40, 900), n_pairs=400, seed=5, rep_dim=12)
        for a in adatas.values():
            models.configure_dataset(a, "NB", use_highly_variable=False, use_rep="X_rep")
        old = config.USE_CUDA_GRAPHS
        config.USE_CUDA_GRAPHS = use_graphs
        try:
            model = models.SCGLUEModel(adatas, sorted(graph.nodes), latent_dim=16, h_dim=32, dropout=0.0,
                                       random_seed=2)
            model.compile()
            kw = dict(data_batch_size=32, max_epochs=1, patience=None, reduce_lr_patience=None, align_burnin=50)
            for _ in range(3):
                model.fit(adatas, graph, **kw)
            losses = model.get_losses(adatas, graph)
            n_graphs = model.trainer.graphs.info()["graphs"]
            return model.encode_graph(graph), model.encode_data("rna", adatas["rna"]), losses, n_graphs
        finally:
            config.USE_CUDA_GRAPHS = old

    v_g, u_g, l_g, n_graphs = run(True)
    v_e, u_e, l_e, _ = run(False)
    assert n_graphs >= 1, "the captured path must have been exercised"
    assert np.isfinite(v_g).all() and np.isfinite(u_g).all() and np.isfinite(list(l_g.values())).all()
    assert_close(v_g, v_e, 1e-4, "vertex embeddings after three fits (graphs vs eager)")
    assert_close(u_g, u_e, 1e-4, "cell embeddings after three fits (graphs vs eager)")
    for k in l_e:
        assert abs(l_g[k] - l_e[k]) <= 1e-4 * max(1.0, abs(l_e[k])), k


@pytest.mark.parametrize("keep_sparse", [False, True])
@pytest.mark.parametrize("nan_sparse", [False, True])
def test_device_resident_minibatches_equal_host_minibatches(native_lib, keep_sparse, nan_sparse):
    """Row f1 of the scope table: a minibatch gathered on the device (dense ``index_select`` or
    ``DeviceCSR`` densify, ``AnnDataset.to_device``) is, bit for bit, the minibatch the host path of the
    reference builds (``AnnDataset.__getitem__``, scglue/models/data.py:397-436), including NaN for the
    structural zeros of ``nan_sparse`` modalities; and so is the packed-sparse host feed after unpacking."""
    from tests.synth import make_synthetic
    from glue_b200 import models
    from glue_b200.models.data import AnnDataset, DataLoader, DeviceCSR, is_packed_sparse, unpack_sparse_rows
    from glue_b200.utils import config
    adatas, graph = make_synthetic(n_cells=(300, 260), n_feat=(70, 900), n_pairs=200, seed=4, rep_dim=8)
    for k, a in adatas.items():
        models.configure_dataset(a, "NB", use_highly_variable=False, use_rep="X_rep",
                                 nan_sparse=nan_sparse and k == "atac")
    cfgs = [a.uns[config.ANNDATA_KEY] for a in adatas.values()]

    def batches(device=None, sparse_items=False):
        ds = AnnDataset(list(adatas.values()), cfgs, mode="train", getitem_size=8)
        if device is not None:
            old = config.DEVICE_DATA_BUDGET_BYTES
            if keep_sparse:
                config.DEVICE_DATA_BUDGET_BYTES = 300 * 970 * 4 // 2
            try:
                ds.to_device(device)
            finally:
                config.DEVICE_DATA_BUDGET_BYTES = old
            if keep_sparse:
                assert any(isinstance(t, DeviceCSR) for grp in ds.device_data for t in grp)
        ds.sparse_items = sparse_items
        ds.prepare_shuffle(num_workers=0, random_seed=6)
        return list(DataLoader(ds, batch_size=4, shuffle=True, generator=torch.Generator().manual_seed(6)))

    host = batches()
    dev = batches(torch.device(DEV))
    packed = batches(sparse_items=True)
    assert len(host) == len(dev) == len(packed) > 5
    for hb, db, pb in zip(host, dev, packed):
        for i, (h, d_, p) in enumerate(zip(hb, db, pb)):
            assert d_.is_cuda and d_.dtype == h.dtype
            assert torch.equal(torch.nan_to_num(d_.cpu(), nan=-7.0), torch.nan_to_num(h, nan=-7.0)), i
            assert torch.equal(torch.isnan(d_.cpu()), torch.isnan(h)), i
            if is_packed_sparse(p) and i < 2:
                got = unpack_sparse_rows(p.to(DEV), h.shape[0], h.shape[1]).cpu()
                assert torch.equal(got, h), i
            else:
                assert torch.equal(torch.nan_to_num(p, nan=-7.0), torch.nan_to_num(h, nan=-7.0)), i
