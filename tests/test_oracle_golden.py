"""The CPU oracle against (a) the reference's own known-answer test for edge normalisation
(reference tests/test_num.py:76-113 with the fixture of tests/conftest.py:277-289) and
(b) golden vectors produced by the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import sampling
from oracle import scglue_cpu as oc

EIDX = np.array([[0, 1, 2, 1], [0, 1, 2, 2]])
EWT = np.array([1.0, 0.4, 0.7, 0.1])


@pytest.mark.parametrize("mod", ["oracle", "product"])
def test_reference_known_answers(mod):
    if mod == "oracle":
        vd, ne = sampling.vertex_degrees, sampling.normalize_edges
    else:
        from glue_b200 import num
        vd, ne = num.vertex_degrees, num.normalize_edges
    assert np.allclose(vd(EIDX, EWT, direction="in"), [1.0, 0.4, 0.8])
    assert np.allclose(vd(EIDX, EWT, direction="out"), [1.0, 0.5, 0.7])
    assert np.allclose(vd(EIDX, EWT, direction="both"), [1.0, 0.5, 0.8])
    assert np.allclose(ne(EIDX, EWT, method="in"), [1.0, 1.0, 0.7 / 0.8, 0.1 / 0.8])
    assert np.allclose(ne(EIDX, EWT, method="out"), [1.0, 0.4 / 0.5, 1.0, 0.1 / 0.5])
    assert np.allclose(ne(EIDX, EWT, method="sym"),
                       [1.0, 0.4 / np.sqrt(0.4 * 0.5), 0.7 / np.sqrt(0.8 * 0.7),
                        0.1 / np.sqrt(0.8 * 0.5)])
    assert np.allclose(ne(EIDX, EWT, method="keepvar"),
                       [1.0, 0.4 / np.sqrt(0.4), 0.7 / np.sqrt(0.8), 0.1 / np.sqrt(0.8)])
    with pytest.raises(ValueError):
        ne(EIDX, EWT, method="xxx")
    with pytest.raises(ValueError):
        vd(EIDX, EWT, direction="xxx")


def test_graph_golden_bit_exact(golden):
    g = golden("graph")
    eidx, ewt, esgn = g["eidx"], g["ewt"], g["esgn"]
    for m in ("in", "out", "sym", "keepvar"):
        assert np.array_equal(sampling.normalize_edges(eidx, ewt, m), g[f"enorm_{m}"])
    for d in ("in", "out", "both"):
        assert np.array_equal(sampling.vertex_degrees(eidx, ewt, direction=d), g[f"degree_{d}"])
    s = sampling.GraphSampler(eidx, ewt, esgn, neg_samples=3)
    assert np.array_equal(s.vprob, g["vprob"]) and np.array_equal(s.eprob, g["eprob"])
    assert s.effective_enum == int(g["effective_enum"]) and s.size == int(g["size"])
    for seed in (0, 5):
        got = s.propose_shuffle(seed)
        assert np.array_equal(got[0], g[f"samp{seed}_eidx"])
        assert np.array_equal(got[1], g[f"samp{seed}_ewt"])
        assert np.array_equal(got[2], g[f"samp{seed}_esgn"])


def test_modules_golden(golden):
    g, m = golden("graph"), golden("modules")
    t = torch.as_tensor
    eidx, enorm, esgn = t(g["eidx"]), t(g["enorm_keepvar"]), t(g["esgn"])
    inp = t(m["gc_in"]).requires_grad_(True)
    out = oc.graph_conv(inp, eidx, enorm, esgn)
    (gin,) = torch.autograd.grad(out, inp, t(m["gc_gout"]))
    assert np.allclose(out.detach(), m["gc_out"], atol=1e-6)
    assert np.allclose(gin, m["gc_gin"], atol=1e-6)
    st = {"g2v." + k[3:].replace("_weight", ".weight").replace("_bias", ".bias"): t(v)
          for k, v in m.items() if k.startswith("ge_") and k not in ("ge_loc", "ge_std")}
    st = {k.replace("g2v.loc.", "g2v.loc.").replace("g2v.std_lin.", "g2v.std_lin."): v
          for k, v in st.items()}
    loc, std = oc.graph_encoder(st, eidx, enorm, esgn)
    assert np.allclose(loc, m["ge_loc"], atol=1e-6) and np.allclose(std, m["ge_std"], atol=1e-6)
    v = t(m["gd_v"]).requires_grad_(True)
    nll = oc.graph_nll(v, t(g["samp0_eidx"]), t(g["samp0_esgn"]), t(g["samp0_ewt"]))
    (gv,) = torch.autograd.grad(nll, v)
    assert np.allclose(nll.detach(), m["gd_nll"], rtol=1e-6)
    assert np.allclose(gv, m["gd_gv"], atol=1e-7)
    names = {"NB": ("scale_lin", "bias", "log_theta"),
             "ZINB": ("scale_lin", "bias", "log_theta", "zi_logits"),
             "Normal": ("scale_lin", "bias", "std_lin"),
             "ZILN": ("scale_lin", "bias", "std_lin", "zi_logits")}
    xkey = {"NB": "dd_counts", "ZINB": "dd_counts", "Normal": "dd_cont", "ZILN": "dd_poscont"}
    for fam, ps in names.items():
        u, vf = t(m["dd_u"]).requires_grad_(True), t(m["dd_v"]).requires_grad_(True)
        st = {f"u2x.m.{n}": t(m[f"dd_{fam}_p_{n}"]).requires_grad_(True) for n in ps}
        l = t(m["dd_l"]) if fam in ("NB", "ZINB") else None
        lp = oc.DECODER_LOG_PROB[fam](st, "u2x.m.", u, vf, t(m["dd_b"]), l, t(m[xkey[fam]]))
        assert np.allclose(lp.detach(), m[f"dd_{fam}_logp"], rtol=1e-4, atol=1e-5), fam
        grads = torch.autograd.grad(-lp.mean(), [u, vf] + list(st.values()))
        assert np.allclose(grads[0], m[f"dd_{fam}_gu"], rtol=1e-4, atol=1e-6), fam
        assert np.allclose(grads[1], m[f"dd_{fam}_gv"], rtol=1e-4, atol=1e-6), fam


@pytest.mark.parametrize("paired", [False, True])
def test_full_step_golden(golden, paired):
    g = golden("step_paired" if paired else "step_scglue")
    graph = golden("graph")
    t = torch.as_tensor
    keys = ["rna", "atac"]
    prob = {"rna": "NB", "atac": "ZINB" if paired else "NB"}
    state = {k[len("init__"):]: t(v).clone() for k, v in g.items() if k.startswith("init__")}
    cfg = oc.StepConfig(keys, prob, {k: True for k in keys}, {"rna": 2, "atac": 1},
                        modality_weight={"rna": 1.0, "atac": 2.0}, paired=paired,
                        lam_joint_cross=0.02 if paired else 0.0,
                        lam_real_cross=0.02 if paired else 0.0, lam_cos=0.02 if paired else 0.0,
                        align_burnin=int(g["align_burnin"]))
    batch = dict(x={k: t(g[f"x__{k}"]) for k in keys}, xrep={k: t(g[f"xrep__{k}"]) for k in keys},
                 xbch={k: t(g[f"xbch__{k}"]) for k in keys},
                 xdwt={k: t(g[f"xdwt__{k}"]) for k in keys}, pmsk=t(g["pmsk"]),
                 eidx=t(g["eidx"]), ewt=t(g["ewt"]), esgn=t(g["esgn"]))
    noise = {name: oc.Noise(u_eps={k: t(g[f"noise_{name}__u_{k}"]) for k in keys},
                            burnin_eps=t(g[f"noise_{name}__burnin"]),
                            v_eps=t(g["noise_gen__v"]) if name == "gen" else None)
             for name in ("dsc", "gen")}
    full = (t(graph["eidx"]), t(graph["enorm_keepvar"]), t(graph["esgn"]))
    trainer = oc.CpuTrainer(cfg, state, full, lr=2e-3)
    losses = trainer.train_step(batch, int(g["epoch"]), noise_dsc=noise["dsc"],
                                noise_gen=noise["gen"])
    for k, v in losses.items():
        assert np.allclose(v.detach(), g[f"loss__{k}"], rtol=1e-5, atol=1e-6), k
    from tests.util import assert_post_step_close
    for k, v in state.items():
        if not v.dtype.is_floating_point:
            assert np.array_equal(v.numpy(), g[f"final__{k}"]), k
            continue
        gref = g.get(f"dscgrad__{k}") if k.startswith("du.") else g.get(f"gengrad__{k}")
        assert_post_step_close(v.detach().numpy(), g[f"final__{k}"], gref, 1e-4, k)


@pytest.mark.parametrize("tag", ["cfg3", "cfg4"])
def test_full_step_golden_new_configs(golden, tag):
    """The oracle against the reference-generated steps at the shapes of BASELINE configs 3 / 4: three
    modalities with a ZILN decoder and guidance edges of both signs; ZINB decoders with 20 batches."""
    g = golden(f"step_{tag}")
    graph = golden("graph")
    t = torch.as_tensor
    keys = [str(k) for k in g["keys"]]
    prob = {k: str(g[f"prob__{k}"]) for k in keys}
    n_batches = {k: int(g[f"init__u2x.{k}.scale_lin"].shape[0]) for k in keys}
    state = {k[len("init__"):]: t(v).clone() for k, v in g.items() if k.startswith("init__")}
    cfg = oc.StepConfig(keys, prob, {k: prob[k] in ("NB", "ZINB") for k in keys}, n_batches,
                        modality_weight={k: float(g[f"mw__{k}"]) for k in keys}, paired=False,
                        align_burnin=int(g["align_burnin"]))
    batch = dict(x={k: t(g[f"x__{k}"]) for k in keys}, xrep={k: t(g[f"xrep__{k}"]) for k in keys},
                 xbch={k: t(g[f"xbch__{k}"]) for k in keys}, xdwt={k: t(g[f"xdwt__{k}"]) for k in keys},
                 pmsk=None, eidx=t(g["eidx"]), ewt=t(g["ewt"]), esgn=t(g["esgn"]))
    noise = {name: oc.Noise(u_eps={k: t(g[f"noise_{name}__u_{k}"]) for k in keys},
                            burnin_eps=t(g[f"noise_{name}__burnin"]),
                            v_eps=t(g["noise_gen__v"]) if name == "gen" else None)
             for name in ("dsc", "gen")}
    full = (t(graph["eidx"]), t(graph["enorm_keepvar"]), t(graph["esgn"]))
    assert (graph["esgn"] < 0).any()
    trainer = oc.CpuTrainer(cfg, state, full, lr=2e-3)
    losses = trainer.train_step(batch, int(g["epoch"]), noise_dsc=noise["dsc"], noise_gen=noise["gen"])
    for k, v in losses.items():
        assert np.allclose(v.detach(), g[f"loss__{k}"], rtol=1e-5, atol=1e-6), k
    from tests.util import assert_post_step_close
    for k, v in state.items():
        if not v.dtype.is_floating_point:
            assert np.array_equal(v.numpy(), g[f"final__{k}"]), k
            continue
        gref = g.get(f"dscgrad__{k}") if k.startswith("du.") else g.get(f"gengrad__{k}")
        assert_post_step_close(v.detach().numpy(), g[f"final__{k}"], gref, 1e-4, k)
