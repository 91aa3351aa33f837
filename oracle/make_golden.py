"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference.

Run in the build container (the only place ``/root/reference`` exists):

    python -m oracle.make_golden

Every fixture stores the inputs and what the reference's own modules returned
for them (forward values and autograd gradients).  While generating, the CPU
oracle (``oracle/scglue_cpu.py``, ``oracle/sampling.py``) is run on the same
inputs and must agree, so a fixture is never written from an unpinned oracle.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import networkx as nx
import numpy as np
import pandas as pd
import torch

from tests.util import make_graph

from . import refstub, sampling
from . import scglue_cpu as oc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _close(a, b, what, rtol=1e-5, atol=1e-6):
    a, b = _np(a), _np(b)
    if not np.allclose(a, b, rtol=rtol, atol=atol, equal_nan=True):
        raise AssertionError(f"oracle disagrees with reference on {what}: "
                             f"max abs diff {np.nanmax(np.abs(a - b))}")


def _exact(a, b, what):
    if not np.array_equal(_np(a), _np(b)):
        raise AssertionError(f"oracle not bit-exact with reference on {what}")


def golden_graph(ref):
    from tests.util import golden_graph as build
    g, vertices, rna, atac = build()
    ds = ref.models.data.GraphDataset(g, pd.Index(vertices), neg_samples=3,
                                      weighted_sampling=True, deemphasize_loops=True)
    eidx, ewt, esgn = ds.eidx, ds.ewt, ds.esgn
    out = dict(
        edge_src=np.array([e[0] for e in g.edges]), edge_dst=np.array([e[1] for e in g.edges]),
        edge_weight=np.array([d["weight"] for _, _, d in g.edges(data=True)]),
        edge_sign=np.array([d["sign"] for _, _, d in g.edges(data=True)]),
        vertices=np.array(vertices), eidx=eidx, ewt=ewt, esgn=esgn,
        vprob=ds.vprob, eprob=ds.eprob, effective_enum=ds.effective_enum, size=ds.size,
    )
    mine = sampling.GraphSampler(eidx, ewt, esgn, neg_samples=3)
    _exact(mine.vprob, ds.vprob, "vprob")
    _exact(mine.eprob, ds.eprob, "eprob")
    for method in ("in", "out", "sym", "keepvar"):
        out[f"enorm_{method}"] = ref.num.normalize_edges(eidx, ewt, method=method)
        _exact(sampling.normalize_edges(eidx, ewt, method), out[f"enorm_{method}"], method)
    for direction in ("in", "out", "both"):
        out[f"degree_{direction}"] = ref.num.vertex_degrees(eidx, ewt, direction=direction)
        _exact(sampling.vertex_degrees(eidx, ewt, direction=direction),
               out[f"degree_{direction}"], direction)
    for seed in (0, 5):
        s_eidx, s_ewt, s_esgn = ds.propose_shuffle(seed)
        m = mine.propose_shuffle(seed)
        _exact(m[0], s_eidx, "sampled eidx"), _exact(m[1], s_ewt, "sampled ewt")
        _exact(m[2], s_esgn, "sampled esgn")
        out[f"samp{seed}_eidx"], out[f"samp{seed}_ewt"], out[f"samp{seed}_esgn"] = (
            s_eidx, s_ewt, s_esgn)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "graph.npz"), **out)
    return out


def golden_modules(ref, graph):
    torch.manual_seed(11)
    sc, nn = ref.models.sc, ref.models.nn
    K = 50
    eidx = torch.as_tensor(graph["eidx"])
    enorm = torch.as_tensor(graph["enorm_keepvar"])
    esgn = torch.as_tensor(graph["esgn"])
    V = int(eidx.max()) + 1
    out = {}

    # --- GraphConv ---
    inp = torch.randn(V, K, requires_grad=True)
    gout = torch.randn(V, K)
    res = nn.GraphConv()(inp, eidx, enorm, esgn)
    (gin,) = torch.autograd.grad(res, inp, gout)
    _close(oc.graph_conv(inp, eidx, enorm, esgn), res, "graph_conv")
    out.update(gc_in=_np(inp), gc_gout=_np(gout), gc_out=_np(res), gc_gin=_np(gin))

    # --- GraphEncoder ---
    enc = sc.GraphEncoder(V, K)
    with torch.no_grad():
        enc.vrepr.copy_(torch.randn(V, K) * 0.5)
    dist = enc(eidx, enorm, esgn)
    st = {"g2v." + k: v for k, v in enc.state_dict().items()}
    loc, std = oc.graph_encoder(st, eidx, enorm, esgn)
    _close(loc, dist.mean, "g2v loc"), _close(std, dist.stddev, "g2v std")
    out.update({"ge_" + k.replace(".", "_"): _np(v) for k, v in enc.state_dict().items()})
    out.update(ge_loc=_np(dist.mean), ge_std=_np(dist.stddev))

    # --- GraphDecoder + BCE on a sampled minibatch ---
    v = (torch.randn(V, K) * 0.3).requires_grad_(True)
    s_eidx = torch.as_tensor(graph["samp0_eidx"])
    s_ewt = torch.as_tensor(graph["samp0_ewt"])
    s_esgn = torch.as_tensor(graph["samp0_esgn"])
    bern = sc.GraphDecoder()(v, s_eidx, s_esgn)
    logp = bern.log_prob(s_ewt)
    g_nll = -logp
    pos = (s_ewt != 0).to(torch.int64)
    n_pos = pos.sum().item()
    n_neg = pos.numel() - n_pos
    bins = torch.zeros(2).scatter_add_(0, pos, g_nll)
    nll = (bins[0] / max(n_neg, 1) + bins[1] / max(n_pos, 1)) / ((n_pos > 0) + (n_neg > 0))
    (gv,) = torch.autograd.grad(nll, v)
    _close(oc.graph_nll(v, s_eidx, s_esgn, s_ewt), nll, "graph nll")
    out.update(gd_v=_np(v), gd_logits=_np(bern.logits), gd_logp=_np(logp), gd_nll=_np(nll),
               gd_gv=_np(gv))

    # --- data decoders ---
    B, Fn, nb = 9, 23, 3
    u = (torch.randn(B, K) * 0.5).requires_grad_(True)
    vf = (torch.randn(Fn, K) * 0.5).requires_grad_(True)
    b = torch.randint(0, nb, (B,))
    counts = torch.poisson(torch.rand(B, Fn) * 3.0)
    counts[0, :3] = torch.tensor([40.0, 0.0, 7.0])
    l = counts.sum(dim=1, keepdim=True)
    cont = torch.randn(B, Fn) * 2.0
    pos_cont = torch.exp(torch.randn(B, Fn)) * (torch.rand(B, Fn) > 0.4)
    out.update(dd_u=_np(u), dd_v=_np(vf), dd_b=_np(b), dd_l=_np(l), dd_counts=_np(counts),
               dd_cont=_np(cont), dd_poscont=_np(pos_cont))
    for name, cls, xval, uses_l in (
        ("NB", sc.NBDataDecoder, counts, True), ("ZINB", sc.ZINBDataDecoder, counts, True),
        ("Normal", sc.NormalDataDecoder, cont, False), ("ZILN", sc.ZILNDataDecoder, pos_cont, False),
    ):
        dec = cls(Fn, n_batches=nb)
        with torch.no_grad():
            for p in dec.parameters():
                p.copy_(torch.randn_like(p) * 0.7)
        dist = dec(u, vf, b, l if uses_l else None)
        lp = dist.log_prob(xval)
        loss = -lp.mean()
        params = dict(dec.named_parameters())
        grads = torch.autograd.grad(loss, [u, vf] + list(params.values()))
        st = {f"u2x.m.{k}": p for k, p in params.items()}
        _close(oc.DECODER_LOG_PROB[name](st, "u2x.m.", u, vf, b, l if uses_l else None, xval),
               lp, f"{name} log_prob", rtol=1e-4, atol=1e-5)
        out[f"dd_{name}_logp"] = _np(lp)
        out[f"dd_{name}_mean"] = _np(dist.mean)
        out[f"dd_{name}_gu"], out[f"dd_{name}_gv"] = _np(grads[0]), _np(grads[1])
        for (k, p), g in zip(params.items(), grads[2:]):
            out[f"dd_{name}_p_{k}"] = _np(p)
            out[f"dd_{name}_g_{k}"] = _np(g)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "modules.npz"), **out)


def _build_reference_net(ref, V, K, feat_idx, prob, rep_dim, h_dim, n_batches):
    sc = ref.models.sc
    enc_cls = {"NB": sc.NBDataEncoder, "ZINB": sc.NBDataEncoder,
               "Normal": sc.VanillaDataEncoder, "ZILN": sc.VanillaDataEncoder}
    dec_cls = {"NB": sc.NBDataDecoder, "ZINB": sc.ZINBDataDecoder,
               "Normal": sc.NormalDataDecoder, "ZILN": sc.ZILNDataDecoder}
    x2u = {k: enc_cls[prob[k]](rep_dim[k] or len(feat_idx[k]), K, h_depth=2, h_dim=h_dim,
                               dropout=0.0) for k in feat_idx}
    u2x = {k: dec_cls[prob[k]](len(feat_idx[k]), n_batches=n_batches[k]) for k in feat_idx}
    idx = {k: torch.as_tensor(np.asarray(v, dtype=np.int64)) for k, v in feat_idx.items()}
    du = sc.Discriminator(K, len(feat_idx), n_batches=0, h_depth=2, h_dim=h_dim, dropout=0.0)
    return ref.models.scglue.SCGLUE(sc.GraphEncoder(V, K), sc.GraphDecoder(), x2u, u2x, idx, du,
                                    sc.Prior())


def golden_step(ref, graph, paired: bool):
    """One full train_step of the reference trainer (dsc update + gen update)."""
    tag = "paired" if paired else "scglue"
    torch.manual_seed(23 + paired)
    K, h_dim, B = 50, 32, 16
    vertices = list(graph["vertices"])
    V = len(vertices)
    feat_idx = {"rna": [i for i, v in enumerate(vertices) if v.startswith("g")],
                "atac": [i for i, v in enumerate(vertices) if v.startswith("p")]}
    prob = {"rna": "NB", "atac": "ZINB" if paired else "NB"}
    rep_dim = {"rna": 0, "atac": 10}
    n_batches = {"rna": 2, "atac": 1}
    net = _build_reference_net(ref, V, K, feat_idx, prob, rep_dim, h_dim, n_batches)
    with torch.no_grad():  # decoders / vrepr are zero-initialised: perturb so every term is live
        net.g2v.vrepr.copy_(torch.randn(V, K) * 0.3)
        for p in net.u2x.parameters():
            p.copy_(torch.randn_like(p) * 0.3)
    common = dict(lam_data=1.0, lam_kl=1.0, lam_graph=0.02, lam_align=0.05, lam_sup=0.02,
                  dsc_steps=1, normalize_u=False, modality_weight={"rna": 1.0, "atac": 2.0},
                  optim="RMSprop", lr=2e-3)
    if paired:
        trainer = ref.models.scglue.PairedSCGLUETrainer(
            net, lam_joint_cross=0.02, lam_real_cross=0.02, lam_cos=0.02, **common)
    else:
        trainer = ref.models.scglue.SCGLUETrainer(net, **common)
    trainer.eidx = torch.as_tensor(graph["eidx"])
    trainer.enorm = torch.as_tensor(graph["enorm_keepvar"])
    trainer.esgn = torch.as_tensor(graph["esgn"])
    trainer.align_burnin = 4
    epoch = 2
    init_state = {k: v.clone() for k, v in net.state_dict().items()}

    x = {"rna": torch.poisson(torch.rand(B, len(feat_idx["rna"])) * 4.0 + 0.5),
         "atac": torch.poisson(torch.rand(B, len(feat_idx["atac"])) * 0.6)}
    x["rna"][:, 0] += 1.0  # no empty libraries
    x["atac"][:, 0] += 1.0
    xrep = {"rna": torch.empty(B, 0), "atac": torch.randn(B, 10)}
    xbch = {"rna": torch.randint(0, 2, (B,)), "atac": torch.zeros(B, dtype=torch.int64)}
    xlbl = {k: torch.full((B,), -1, dtype=torch.int64) for k in x}
    xdwt = {k: torch.rand(B) + 0.5 for k in x}
    xflag = {k: torch.full((B,), i, dtype=torch.int64) for i, k in enumerate(x)}
    pmsk = torch.rand(B, 2) > 0.3
    pmsk[:, 0] |= ~pmsk[:, 1]  # every row observed in at least one modality
    pmsk[0] = True
    mb = slice(0, 200)
    s_eidx = torch.as_tensor(graph["samp5_eidx"][:, mb])
    s_ewt = torch.as_tensor(graph["samp5_ewt"][mb])
    s_esgn = torch.as_tensor(graph["samp5_esgn"][mb])
    data = ((x, xrep, xbch, xlbl, xdwt, xflag, pmsk, s_eidx, s_ewt, s_esgn) if paired
            else (x, xrep, xbch, xlbl, xdwt, xflag, s_eidx, s_ewt, s_esgn))

    def draw_noise(seed, full):
        torch.manual_seed(seed)
        u_eps = {k: torch.randn(B, K) for k in x}
        burn = torch.randn(2 * B, K)
        v_eps = torch.randn(V, K) if full else None
        return oc.Noise(u_eps=u_eps, v_eps=v_eps, burnin_eps=burn)

    net.train()
    out = {}
    # discriminator half-step
    torch.manual_seed(101)
    losses = trainer.compute_losses(data, epoch, dsc_only=True)
    net.zero_grad(set_to_none=True)
    losses["dsc_loss"].backward()
    out["dsc_only_loss"] = _np(losses["dsc_loss"])
    for n, p in net.du.named_parameters():
        out[f"dscgrad__du.{n}"] = _np(p.grad)
    trainer.dsc_optim.step()
    # generator half-step
    torch.manual_seed(202)
    losses = trainer.compute_losses(data, epoch)
    net.zero_grad(set_to_none=True)
    losses["gen_loss"].backward()
    for k, v in losses.items():
        out[f"loss__{k}"] = _np(v)
    for n, p in net.named_parameters():
        if p.grad is not None:
            out[f"gengrad__{n}"] = _np(p.grad)
    trainer.vae_optim.step()
    for k, v in net.state_dict().items():
        out[f"final__{k}"] = _np(v)

    # same step through the oracle (explicit noise drawn in the reference's RNG order)
    cfg = oc.StepConfig(["rna", "atac"], prob, {"rna": True, "atac": True}, n_batches,
                        modality_weight={"rna": 1.0, "atac": 2.0}, paired=paired,
                        lam_joint_cross=0.02 if paired else 0.0,
                        lam_real_cross=0.02 if paired else 0.0, lam_cos=0.02 if paired else 0.0,
                        align_burnin=4)
    state = {k: v.clone() for k, v in init_state.items()}
    batch = dict(x=x, xrep=xrep, xbch=xbch, xdwt=xdwt, pmsk=pmsk, eidx=s_eidx, ewt=s_ewt,
                 esgn=s_esgn)
    cpu = oc.CpuTrainer(cfg, state, (trainer.eidx, trainer.enorm, trainer.esgn), lr=2e-3)
    o_losses = cpu.train_step(batch, epoch, noise_dsc=draw_noise(101, False),
                              noise_gen=draw_noise(202, True))
    for k, v in losses.items():
        _close(o_losses[k], v, f"{tag} loss {k}", rtol=1e-5, atol=1e-6)
    for k, v in net.state_dict().items():
        _close(state[k], v, f"{tag} post-step {k}", rtol=1e-4, atol=1e-6)

    for k, v in init_state.items():
        out[f"init__{k}"] = _np(v)
    for k in x:
        out[f"x__{k}"], out[f"xrep__{k}"], out[f"xbch__{k}"], out[f"xdwt__{k}"] = (
            _np(x[k]), _np(xrep[k]), _np(xbch[k]), _np(xdwt[k]))
    for seed, full, name in ((101, False, "dsc"), (202, True, "gen")):
        nz = draw_noise(seed, full)
        for k in x:
            out[f"noise_{name}__u_{k}"] = _np(nz.u_eps[k])
        out[f"noise_{name}__burnin"] = _np(nz.burnin_eps)
        if full:
            out[f"noise_{name}__v"] = _np(nz.v_eps)
    out.update(pmsk=_np(pmsk), eidx=_np(s_eidx), ewt=_np(s_ewt), esgn=_np(s_esgn),
               epoch=epoch, align_burnin=4)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"step_{tag}.npz"), **out)


def golden_step_config(ref, graph, tag, keys, prefix, prob, rep_dim, n_batches, B=16):
    """One full train_step of the reference ``SCGLUETrainer`` for a BASELINE-config shaped model:
    ``cfg3`` = three modalities, one of them ZILN, guidance edges of both signs;
    ``cfg4`` = ZINB decoders with 20 batches (``use_batch``).  Same layout as ``step_scglue.npz``
    plus ``keys`` / ``prob__<k>`` so that the test builds the model from the file alone."""
    torch.manual_seed(31 + len(keys) + max(n_batches.values()))
    K, h_dim = 50, 32
    vertices = list(graph["vertices"])
    V = len(vertices)
    feat_idx = {k: [i for i, v in enumerate(vertices) if v.startswith(prefix[k])] for k in keys}
    net = _build_reference_net(ref, V, K, feat_idx, prob, rep_dim, h_dim, n_batches)
    with torch.no_grad():
        net.g2v.vrepr.copy_(torch.randn(V, K) * 0.3)
        for p in net.u2x.parameters():
            p.copy_(torch.randn_like(p) * 0.3)
    mw = {k: 1.0 + 0.5 * i for i, k in enumerate(keys)}
    trainer = ref.models.scglue.SCGLUETrainer(
        net, lam_data=1.0, lam_kl=1.0, lam_graph=0.02, lam_align=0.05, lam_sup=0.02, dsc_steps=1,
        normalize_u=False, modality_weight=mw, optim="RMSprop", lr=2e-3)
    trainer.eidx = torch.as_tensor(graph["eidx"])
    trainer.enorm = torch.as_tensor(graph["enorm_keepvar"])
    trainer.esgn = torch.as_tensor(graph["esgn"])
    assert (graph["esgn"] < 0).any(), "the golden graph must carry sign -1 edges"
    trainer.align_burnin = 4
    epoch = 2
    init_state = {k: v.clone() for k, v in net.state_dict().items()}

    x, xrep = {}, {}
    for k in keys:
        F = len(feat_idx[k])
        if prob[k] in ("NB", "ZINB"):
            x[k] = torch.poisson(torch.rand(B, F) * 2.0 + 0.2)
            x[k][:, 0] += 1.0  # no empty libraries
        else:  # ZILN: positive continuous values with exact zeros
            x[k] = torch.exp(torch.randn(B, F)) * (torch.rand(B, F) > 0.45)
        xrep[k] = torch.randn(B, rep_dim[k]) if rep_dim[k] else torch.empty(B, 0)
    xbch = {k: torch.randint(0, n_batches[k], (B,)) for k in keys}
    xlbl = {k: torch.full((B,), -1, dtype=torch.int64) for k in keys}
    xdwt = {k: torch.rand(B) + 0.5 for k in keys}
    xflag = {k: torch.full((B,), i, dtype=torch.int64) for i, k in enumerate(keys)}
    mb = slice(0, 200)
    s_eidx = torch.as_tensor(graph["samp5_eidx"][:, mb])
    s_ewt = torch.as_tensor(graph["samp5_ewt"][mb])
    s_esgn = torch.as_tensor(graph["samp5_esgn"][mb])
    data = (x, xrep, xbch, xlbl, xdwt, xflag, s_eidx, s_ewt, s_esgn)

    def draw_noise(seed, full):
        torch.manual_seed(seed)
        u_eps = {k: torch.randn(B, K) for k in keys}
        burn = torch.randn(len(keys) * B, K)
        v_eps = torch.randn(V, K) if full else None
        return oc.Noise(u_eps=u_eps, v_eps=v_eps, burnin_eps=burn)

    net.train()
    out = {}
    torch.manual_seed(101)
    losses = trainer.compute_losses(data, epoch, dsc_only=True)
    net.zero_grad(set_to_none=True)
    losses["dsc_loss"].backward()
    out["dsc_only_loss"] = _np(losses["dsc_loss"])
    for n, p in net.du.named_parameters():
        out[f"dscgrad__du.{n}"] = _np(p.grad)
    trainer.dsc_optim.step()
    torch.manual_seed(202)
    losses = trainer.compute_losses(data, epoch)
    net.zero_grad(set_to_none=True)
    losses["gen_loss"].backward()
    for k, v in losses.items():
        out[f"loss__{k}"] = _np(v)
    for n, p in net.named_parameters():
        if p.grad is not None:
            out[f"gengrad__{n}"] = _np(p.grad)
    trainer.vae_optim.step()
    for k, v in net.state_dict().items():
        out[f"final__{k}"] = _np(v)

    cfg = oc.StepConfig(keys, prob, {k: prob[k] in ("NB", "ZINB") for k in keys}, n_batches,
                        modality_weight=mw, paired=False, align_burnin=4)
    state = {k: v.clone() for k, v in init_state.items()}
    batch = dict(x=x, xrep=xrep, xbch=xbch, xdwt=xdwt, pmsk=None, eidx=s_eidx, ewt=s_ewt, esgn=s_esgn)
    cpu = oc.CpuTrainer(cfg, state, (trainer.eidx, trainer.enorm, trainer.esgn), lr=2e-3)
    o_losses = cpu.train_step(batch, epoch, noise_dsc=draw_noise(101, False), noise_gen=draw_noise(202, True))
    for k, v in losses.items():
        _close(o_losses[k], v, f"{tag} loss {k}", rtol=1e-5, atol=1e-6)
    for k, v in net.state_dict().items():
        _close(state[k], v, f"{tag} post-step {k}", rtol=1e-4, atol=1e-6)

    for k, v in init_state.items():
        out[f"init__{k}"] = _np(v)
    for k in keys:
        out[f"x__{k}"], out[f"xrep__{k}"], out[f"xbch__{k}"], out[f"xdwt__{k}"] = (
            _np(x[k]), _np(xrep[k]), _np(xbch[k]), _np(xdwt[k]))
        out[f"prob__{k}"] = np.array(prob[k])
        out[f"mw__{k}"] = np.array(mw[k])
    for seed, full, name in ((101, False, "dsc"), (202, True, "gen")):
        nz = draw_noise(seed, full)
        for k in keys:
            out[f"noise_{name}__u_{k}"] = _np(nz.u_eps[k])
        out[f"noise_{name}__burnin"] = _np(nz.burnin_eps)
        if full:
            out[f"noise_{name}__v"] = _np(nz.v_eps)
    out.update(keys=np.array(keys), pmsk=np.ones((B, len(keys)), dtype=bool), eidx=_np(s_eidx), ewt=_np(s_ewt),
               esgn=_np(s_esgn), epoch=epoch, align_burnin=4)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"step_{tag}.npz"), **out)


def golden_new_configs(ref, graph):
    golden_step_config(ref, graph, "cfg3", ["rna", "atac", "met"], {"rna": "g", "atac": "p", "met": "z"},
                       {"rna": "NB", "atac": "NB", "met": "ZILN"}, {"rna": 0, "atac": 10, "met": 0},
                       {"rna": 1, "atac": 1, "met": 1})
    golden_step_config(ref, graph, "cfg4", ["rna", "atac"], {"rna": "g", "atac": "p"},
                       {"rna": "ZINB", "atac": "ZINB"}, {"rna": 0, "atac": 10}, {"rna": 20, "atac": 20})


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "new-configs":
        # (adds step_cfg3.npz / step_cfg4.npz without rewriting the round-1 fixtures)
        ref = refstub.load()
        torch.set_num_threads(1)
        golden_new_configs(ref, dict(np.load(os.path.join(GOLDEN_DIR, "graph.npz"), allow_pickle=False)))
        return
    if not refstub.available():
        sys.exit("reference checkout not found; golden vectors can only be regenerated "
                 "in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = refstub.load()
    torch.set_num_threads(1)
    graph = golden_graph(ref)
    golden_modules(ref, graph)
    golden_step(ref, graph, paired=False)
    golden_step(ref, graph, paired=True)
    golden_new_configs(ref, graph)
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)))


if __name__ == "__main__":
    main()
