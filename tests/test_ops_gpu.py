"""GPU parity tests of the native kernels (through the C ABI) against the CPU oracle and
the golden vectors generated from the reference.  Tolerance for floating point: 1e-4
max-norm relative (BASELINE.json north_star); integer / index outputs bit-exact."""
import numpy as np
import pytest
import torch

from oracle import sampling
from oracle import scglue_cpu as oc
from tests.util import assert_close, random_graph, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops(native_lib):
    from glue_b200 import ops as _ops
    return _ops


def _cuda(x):
    return torch.as_tensor(x).to(DEV)


# --------------------------------------------------------------------------- CSR
@pytest.mark.parametrize("V,E", [(37, 200), (5000, 40000), (52000, 100000)])
def test_csr_build_bit_exact(ops, V, E):
    rs = np.random.RandomState(V)
    eidx, ewt, esgn = random_graph(rs, V, E, loops=(V != 37))
    eidx[:, : E // 10] = eidx[:, E // 10: 2 * (E // 10)]  # multi-edges
    enorm = sampling.normalize_edges(eidx, ewt)
    csr = ops.GraphCSR(_cuda(eidx), _cuda(enorm), _cuda(esgn), V)
    for got, want in ((csr.fwd, sampling.csr_forward(eidx, enorm, esgn, V)),
                      (csr.bwd, sampling.csr_transposed(eidx, enorm, esgn, V))):
        for g, w in zip(got, want):
            assert g.dtype == torch.as_tensor(w).dtype
            assert np.array_equal(g.cpu().numpy(), w)


def test_csr_golden(ops, golden):
    g = golden("graph")
    V = len(g["vertices"])
    csr = ops.GraphCSR(_cuda(g["eidx"]), _cuda(g["enorm_keepvar"]), _cuda(g["esgn"]), V)
    want = sampling.csr_forward(g["eidx"], g["enorm_keepvar"], g["esgn"], V)
    for got, w in zip(csr.fwd, want):
        assert np.array_equal(got.cpu().numpy(), w)


# --------------------------------------------------------------------- GraphConv
@pytest.mark.parametrize("V,E,K", [(37, 200, 50), (5000, 40000, 50), (3000, 20000, 16),
                                   (1000, 5000, 7), (52000, 100000, 50), (800, 3000, 128)])
def test_graphconv_vs_oracle(ops, V, E, K):
    rs = np.random.RandomState(K + V)
    eidx, ewt, esgn = random_graph(rs, V, E)
    enorm = sampling.normalize_edges(eidx, ewt)
    inp = torch.randn(V, K, generator=torch.Generator().manual_seed(1))
    gout = torch.randn(V, K, generator=torch.Generator().manual_seed(2))
    ref_in = inp.clone().requires_grad_(True)
    ref = oc.graph_conv(ref_in, torch.as_tensor(eidx), torch.as_tensor(enorm), torch.as_tensor(esgn))
    (ref_gin,) = torch.autograd.grad(ref, ref_in, gout)

    csr = ops.GraphCSR(_cuda(eidx), _cuda(enorm), _cuda(esgn), V)
    x = inp.to(DEV).requires_grad_(True)
    out = ops.graph_conv(x, csr)
    (gin,) = torch.autograd.grad(out, x, gout.to(DEV))
    assert_close(out, ref, 1e-6, "graphconv fwd")
    assert_close(gin, ref_gin, 1e-6, "graphconv bwd")
    # atomic-free => run-to-run bit identical
    assert torch.equal(out, ops.graph_conv(x, csr))


def test_graphconv_golden(ops, golden):
    g, m = golden("graph"), golden("modules")
    V = m["gc_in"].shape[0]
    csr = ops.GraphCSR(_cuda(g["eidx"]), _cuda(g["enorm_keepvar"]), _cuda(g["esgn"]), V)
    x = _cuda(m["gc_in"]).requires_grad_(True)
    out = ops.graph_conv(x, csr)
    (gin,) = torch.autograd.grad(out, x, _cuda(m["gc_gout"]))
    assert_close(out, m["gc_out"], 1e-6, "graphconv fwd vs reference")
    assert_close(gin, m["gc_gin"], 1e-6, "graphconv bwd vs reference")


def test_graphconv_isolated_vertices(ops):
    V, K = 64, 50
    eidx = np.array([[0, 1, 2], [5, 5, 7]], dtype=np.int64)
    ewt = np.array([0.3, 0.6, 1.0], dtype=np.float32)
    esgn = np.array([1.0, -1.0, 1.0], dtype=np.float32)
    enorm = sampling.normalize_edges(eidx, ewt)
    csr = ops.GraphCSR(_cuda(eidx), _cuda(enorm), _cuda(esgn), V)
    x = torch.randn(V, K)
    out = ops.graph_conv(x.to(DEV), csr)
    ref = oc.graph_conv(x, torch.as_tensor(eidx), torch.as_tensor(enorm), torch.as_tensor(esgn))
    assert torch.equal(out.cpu() == 0, ref == 0)
    assert_close(out, ref, 1e-6)


# ----------------------------------------------------------------- graph decoder
@pytest.mark.parametrize("V,E,K,pos_frac", [(60, 333, 50, 0.1), (5000, 36000, 50, 1 / 11),
                                            (400, 1000, 16, 0.0), (400, 1000, 7, 1.0)])
def test_graph_nll_vs_oracle(ops, V, E, K, pos_frac):
    rs = np.random.RandomState(E)
    eidx = rs.randint(0, V, (2, E)).astype(np.int64)
    eidx[1, :5] = eidx[0, :5]  # self loops
    esgn = np.where(rs.rand(E) < 0.3, -1.0, 1.0).astype(np.float32)
    ewt = (rs.rand(E) < pos_frac).astype(np.float32)
    v = torch.randn(V, K, generator=torch.Generator().manual_seed(3)) * 0.4
    ref_v = v.clone().requires_grad_(True)
    ref = oc.graph_nll(ref_v, torch.as_tensor(eidx), torch.as_tensor(esgn), torch.as_tensor(ewt))
    (ref_g,) = torch.autograd.grad(ref, ref_v)

    dv = v.to(DEV).requires_grad_(True)
    loss = ops.graph_nll(dv, _cuda(eidx), _cuda(esgn), _cuda(ewt))
    (g,) = torch.autograd.grad(3.0 * loss, dv)
    assert_close(loss, ref, TOL, "graph nll")
    assert_close(g, 3.0 * ref_g, TOL, "graph nll grad")
    logits = ops.graph_logits(dv, _cuda(eidx), _cuda(esgn))
    assert_close(logits, oc.graph_decoder_logits(v, torch.as_tensor(eidx), torch.as_tensor(esgn)),
                 1e-5, "graph logits")
    loss2 = ops.graph_nll(dv, _cuda(eidx), _cuda(esgn), _cuda(ewt))
    (g2,) = torch.autograd.grad(3.0 * loss2, dv)
    assert torch.equal(g, g2), "segmented-reduction backward must be deterministic"


@pytest.mark.parametrize("V,E,K", [(3000, 6000, 50), (3000, 6000, 64), (500, 3000, 7), (400, 2000, 72)])
def test_graph_nll_hub_segments_and_untouched_rows(ops, V, E, K):
    """A hub vertex with thousands of incidences (segments longer than one chunk of the walk), long runs of
    untouched vertices in the middle and at both ends of the table: every row of the gradient is written
    (zeros where no sampled edge ends), sums match the oracle and are bit-reproducible."""
    rs = np.random.RandomState(V + K)
    lo, hi = V // 10, V // 2  # only vertices in [lo, hi) occur, except the hub
    eidx = rs.randint(lo, hi, (2, E)).astype(np.int64)
    hub = lo + 3
    eidx[0, rs.rand(E) < 0.3] = hub
    eidx[1, rs.rand(E) < 0.1] = hub
    eidx[:, (eidx[0] >= lo + 100) & (eidx[0] < lo + 400)] = lo + 50  # a hole of 300 untouched vertices
    eidx[1, (eidx[1] >= lo + 100) & (eidx[1] < lo + 400)] = lo + 51
    esgn = np.where(rs.rand(E) < 0.3, -1.0, 1.0).astype(np.float32)
    ewt = (rs.rand(E) < 1 / 11).astype(np.float32)
    v = torch.randn(V, K, generator=torch.Generator().manual_seed(5)) * 0.3
    ref_v = v.clone().requires_grad_(True)
    ref = oc.graph_nll(ref_v, torch.as_tensor(eidx), torch.as_tensor(esgn), torch.as_tensor(ewt))
    (ref_g,) = torch.autograd.grad(ref, ref_v)
    dv = v.to(DEV).requires_grad_(True)
    grads = []
    for _ in range(2):
        loss = ops.graph_nll(dv, _cuda(eidx), _cuda(esgn), _cuda(ewt))
        junk = torch.full((V, K), float("nan"), device=DEV)  # the gradient buffer reuses this block
        del junk
        (g,) = torch.autograd.grad(loss, dv)
        grads.append(g)
    assert_close(loss, ref, TOL, "graph nll")
    assert_close(grads[0], ref_g, TOL, "graph nll grad")
    untouched = np.setdiff1d(np.arange(V), np.unique(eidx))
    assert len(untouched) > V // 2
    assert torch.equal(grads[0][_cuda(untouched)], torch.zeros(len(untouched), K, device=DEV))
    assert torch.equal(grads[0], grads[1])


def test_graph_decoder_golden(ops, golden):
    g, m = golden("graph"), golden("modules")
    v = _cuda(m["gd_v"]).requires_grad_(True)
    eidx, ewt, esgn = _cuda(g["samp0_eidx"]), _cuda(g["samp0_ewt"]), _cuda(g["samp0_esgn"])
    loss = ops.graph_nll(v, eidx, esgn, ewt)
    (gv,) = torch.autograd.grad(loss, v)
    assert_close(loss, m["gd_nll"], TOL, "graph nll vs reference")
    assert_close(gv, m["gd_gv"], TOL, "graph nll grad vs reference")
    logits = ops.graph_logits(v, eidx, esgn)
    assert_close(logits, m["gd_logits"], 1e-5)
    logp = torch.distributions.Bernoulli(logits=logits).log_prob(ewt)
    (gv2,) = torch.autograd.grad(logp.sum(), v)
    ref_v = torch.as_tensor(m["gd_v"]).requires_grad_(True)
    ref_lp = torch.distributions.Bernoulli(logits=oc.graph_decoder_logits(
        ref_v, torch.as_tensor(g["samp0_eidx"]), torch.as_tensor(g["samp0_esgn"]))).log_prob(
        torch.as_tensor(g["samp0_ewt"]))
    (ref_g2,) = torch.autograd.grad(ref_lp.sum(), ref_v)
    assert_close(logp, m["gd_logp"], 1e-5)
    assert_close(gv2, ref_g2, TOL, "logits backward")


# ------------------------------------------------------------------ data decoder
PARAM_NAMES = {"NB": ("scale_lin", "bias", "log_theta"),
               "ZINB": ("scale_lin", "bias", "log_theta", "zi_logits"),
               "Normal": ("scale_lin", "bias", "std_lin"),
               "ZILN": ("scale_lin", "bias", "std_lin", "zi_logits")}


def _decoder_case(family, B, F, K, nb, seed, nan_frac=0.0, extra_rows=0):
    gen = torch.Generator().manual_seed(seed)
    u = torch.randn(B, K, generator=gen) * 0.5
    v = torch.randn(F + extra_rows, K, generator=gen) * 0.5
    b = torch.randint(0, nb, (B,), generator=gen)
    params = {n: torch.randn(nb, F, generator=gen) * 0.5 for n in PARAM_NAMES[family]}
    if family in ("NB", "ZINB"):
        x = torch.poisson(torch.rand(B, F, generator=gen) * 2.0, generator=gen)
        x[0, : min(4, F)] = torch.tensor([40.0, 13.0, 0.0, 200.0])[: min(4, F)]
        l = x.sum(dim=1, keepdim=True) + 1.0
    elif family == "Normal":
        x, l = torch.randn(B, F, generator=gen) * 2.0, None
    else:
        x = torch.exp(torch.randn(B, F, generator=gen)) * (torch.rand(B, F, generator=gen) > 0.4)
        l = None
    if nan_frac:
        x[torch.rand(B, F, generator=gen) < nan_frac] = float("nan")
    vidx = torch.randperm(F + extra_rows, generator=gen)[:F] if extra_rows else None
    return u, v, b, l, x, params, vidx


def _oracle_nll(family, u, v, b, l, x, params, vidx, reduce):
    u, v = u.clone().requires_grad_(True), v.clone().requires_grad_(True)
    st = {f"u2x.m.{k}": p.clone().requires_grad_(True) for k, p in params.items()}
    vv = v if vidx is None else v[vidx]
    lp = oc.DECODER_LOG_PROB[family](st, "u2x.m.", u, vv, b, l, torch.nan_to_num(x, nan=1.0))
    if reduce == "nanmean":
        mask = ~torch.isnan(x)
        loss = -(lp * mask).sum() / mask.sum()
    else:
        loss = -lp.mean()
    grads = torch.autograd.grad(loss, [u, v] + list(st.values()))
    return loss, lp, grads


def _device_nll(ops, family, u, v, b, l, x, params, vidx, reduce, scale=1.0):
    d = lambda t: None if t is None else t.to(DEV)
    du, dv = d(u).requires_grad_(True), d(v).requires_grad_(True)
    dp = [d(params[n]).requires_grad_(True) for n in PARAM_NAMES[family]]
    zi = dp[3] if len(dp) == 4 else None
    loss = ops.decoder_nll(family, du, dv, dp[0], dp[1], dp[2], zi, b=d(b), l=d(l), x=d(x),
                           vidx=d(vidx), reduce=reduce)
    grads = torch.autograd.grad(scale * loss, [du, dv] + dp)
    return loss, grads


CASES = [
    # family, B, F, K, n_batches, nan_frac, extra_rows, reduce
    ("NB", 9, 23, 50, 3, 0.0, 0, "mean"),
    ("NB", 128, 2000, 50, 1, 0.0, 0, "nanmean"),
    ("NB", 128, 700, 50, 1, 0.0, 300, "nanmean"),
    ("NB", 200, 1500, 50, 4, 0.0, 0, "mean"),
    ("NB", 33, 129, 10, 2, 0.0, 0, "nanmean"),
    ("NB", 64, 400, 64, 1, 0.0, 0, "nanmean"),
    ("ZINB", 128, 2000, 50, 20, 0.0, 0, "nanmean"),
    ("ZINB", 17, 95, 50, 1, 0.0, 7, "mean"),
    ("Normal", 128, 1000, 50, 3, 0.05, 0, "nanmean"),
    ("Normal", 40, 333, 16, 1, 0.0, 0, "mean"),
    ("ZILN", 128, 1000, 50, 2, 0.05, 0, "nanmean"),
    ("ZILN", 300, 260, 50, 1, 0.0, 40, "mean"),
    ("NB", 1, 5, 50, 1, 0.0, 0, "nanmean"),
]


@pytest.mark.parametrize("family,B,F,K,nb,nan_frac,extra,reduce", CASES)
def test_decoder_nll_vs_oracle(ops, family, B, F, K, nb, nan_frac, extra, reduce):
    case = _decoder_case(family, B, F, K, nb, seed=B * 1000 + F, nan_frac=nan_frac, extra_rows=extra)
    ref_loss, _, ref_grads = _oracle_nll(family, *case, reduce)
    loss, grads = _device_nll(ops, family, *case, reduce, scale=2.5)
    assert_close(loss, ref_loss, TOL, f"{family} nll")
    names = ["u", "v"] + list(PARAM_NAMES[family])
    for n, g, rg in zip(names, grads, ref_grads):
        assert_close(g, 2.5 * rg, TOL, f"{family} d/d{n}")
    loss2, grads2 = _device_nll(ops, family, *case, reduce, scale=2.5)
    assert torch.equal(loss, loss2) and all(torch.equal(a, b) for a, b in zip(grads, grads2)), \
        "decoder must be run-to-run deterministic"


@pytest.mark.parametrize("family", ["NB", "ZINB", "Normal", "ZILN"])
def test_decoder_golden(ops, golden, family):
    m = golden("modules")
    xkey = {"NB": "dd_counts", "ZINB": "dd_counts", "Normal": "dd_cont", "ZILN": "dd_poscont"}[family]
    u, v = _cuda(m["dd_u"]).requires_grad_(True), _cuda(m["dd_v"]).requires_grad_(True)
    b, l, x = _cuda(m["dd_b"]), _cuda(m["dd_l"]), _cuda(m[xkey])
    ps = [_cuda(m[f"dd_{family}_p_{n}"]).requires_grad_(True) for n in PARAM_NAMES[family]]
    zi = ps[3] if len(ps) == 4 else None
    use_l = l if family in ("NB", "ZINB") else None
    # fused NLL + gradients vs what the reference's autograd produced
    loss = ops.decoder_nll(family, u, v, ps[0], ps[1], ps[2], zi, b=b, l=use_l, x=x, reduce="mean")
    grads = torch.autograd.grad(loss, [u, v] + ps)
    assert_close(loss, -m[f"dd_{family}_logp"].mean(), TOL, "nll vs reference")
    assert_close(grads[0], m[f"dd_{family}_gu"], TOL, "d/du vs reference")
    assert_close(grads[1], m[f"dd_{family}_gv"], TOL, "d/dv vs reference")
    for n, g in zip(PARAM_NAMES[family], grads[2:]):
        assert_close(g, m[f"dd_{family}_g_{n}"], TOL, f"d/d{n} vs reference")
    # materialised log_prob / mean
    logp = ops.decoder_log_prob(family, u, v, ps[0], ps[1], ps[2], zi, b=b, l=use_l, x=x)
    ref_lp = torch.as_tensor(m[f"dd_{family}_logp"])
    assert (logp.cpu() - ref_lp).abs().max().item() <= 1e-4 * max(1.0, ref_lp.abs().max().item())
    mean = ops.decoder_mean(family, u, v, ps[0], ps[1], ps[2], zi, b=b, l=use_l)
    assert_close(mean, m[f"dd_{family}_mean"], TOL, "mean vs reference")
    # autograd through the materialised log_prob (per-element upstream weights)
    w = torch.rand(logp.shape, generator=torch.Generator().manual_seed(5)).to(DEV)
    g2 = torch.autograd.grad((logp * w).sum(), [u, v] + ps)
    cu, cv = torch.as_tensor(m["dd_u"]).requires_grad_(True), torch.as_tensor(m["dd_v"]).requires_grad_(True)
    st = {f"u2x.m.{n}": torch.as_tensor(m[f"dd_{family}_p_{n}"]).requires_grad_(True)
          for n in PARAM_NAMES[family]}
    lp_cpu = oc.DECODER_LOG_PROB[family](st, "u2x.m.", cu, cv, torch.as_tensor(m["dd_b"]),
                                        torch.as_tensor(m["dd_l"]) if use_l is not None else None,
                                        torch.as_tensor(m[xkey]))
    ref_g2 = torch.autograd.grad((lp_cpu * w.cpu()).sum(), [cu, cv] + list(st.values()))
    for g, rg in zip(g2, ref_g2):
        assert_close(g, rg, TOL, "log_prob backward")


@pytest.mark.parametrize("family,B,F,nb,nan_frac", [
    ("NB", 128, 50000, 1, 0.0), ("ZINB", 128, 50000, 1, 0.0), ("Normal", 128, 48000, 1, 0.0),
    ("ZINB", 128, 50000, 20, 0.0),    # BASELINE config 4: 20 batches at the full ATAC width
    ("ZILN", 128, 48000, 1, 0.05),    # config 3: methylation modality, 5 % NaN (nan_sparse / nanmean)
    ("NB", 512, 50000, 1, 0.0),       # config 5: atlas batch (four cell blocks, feature-side accumulation)
])
def test_decoder_full_size_properties(ops, family, B, F, nb, nan_frac):
    """BASELINE-size checks: fused loss == (nan)mean of the materialised log_prob; every gradient agrees
    with the CPU oracle; results are deterministic."""
    K = 50
    case = _decoder_case(family, B, F, K, nb, seed=99, nan_frac=nan_frac)
    u, v, b, l, x, params, vidx = case
    loss, grads = _device_nll(ops, family, *case, "nanmean")
    d = lambda t: None if t is None else t.to(DEV)
    dp = [d(params[n]) for n in PARAM_NAMES[family]]
    zi = dp[3] if len(dp) == 4 else None
    if not nan_frac:
        logp = ops.decoder_log_prob(family, d(u), d(v), dp[0], dp[1], dp[2], zi, b=d(b), l=d(l), x=d(x))
        assert_close(loss, -logp.double().mean().float(), TOL, "fused loss vs materialised mean")
    ref_loss, _, ref_grads = _oracle_nll(family, *case, "nanmean")
    assert_close(loss, ref_loss, TOL)
    for n, g, rg in zip(["u", "v"] + list(PARAM_NAMES[family]), grads, ref_grads):
        assert_close(g, rg, TOL, f"{family} full-size gradient d/d{n}")
    loss2, grads2 = _device_nll(ops, family, *case, "nanmean")
    assert torch.equal(loss, loss2) and all(torch.equal(a, c) for a, c in zip(grads, grads2))


MULTI_CASES = [
    # family, B, F, K, n_batches, n_terms, extra_rows
    ("NB", 128, 2000, 50, 1, 3, 0),
    ("NB", 128, 700, 50, 1, 3, 300),
    ("NB", 77, 391, 50, 1, 2, 0),
    ("NB", 128, 1300, 50, 4, 3, 0),
    ("ZINB", 128, 2000, 50, 1, 3, 0),
    ("ZINB", 100, 600, 50, 20, 3, 0),
    ("Normal", 128, 1000, 50, 3, 3, 0),
    ("ZILN", 128, 1000, 50, 1, 3, 11),
    ("NB", 128, 640, 64, 1, 3, 0),
    ("NB", 128, 2000, 50, 1, 5, 0),   # more terms than one launch sequence takes: two groups
]


@pytest.mark.parametrize("family,B,F,K,nb,T,extra", MULTI_CASES)
def test_decoder_multi_terms_vs_oracle(ops, family, B, F, K, nb, T, extra):
    """Several loss terms of one modality in one launch sequence (reconstruction / joint-cross /
    real-cross of PairedSCGLUETrainer, scglue.py:603-652): every loss, and the gradients of
    sum_t c_t loss_t, against the CPU oracle evaluated term by term."""
    u, v, b, l, x, params, vidx = _decoder_case(family, B, F, K, nb, seed=B * 1000 + F + T, extra_rows=extra)
    gen = torch.Generator().manual_seed(T)
    us = [u] + [torch.randn(B, K, generator=gen) * 0.5 for _ in range(T - 1)]
    masks = [torch.ones(B)] + [(torch.rand(B, generator=gen) < 0.6).float() for _ in range(T - 1)]
    masks[-1][:] = masks[-1] if masks[-1].sum() > 0 else 1.0
    rws = [None] + [m / (m.sum().clamp(min=1.0) * F) for m in masks[1:]]
    coefs = [1.0, 0.02, 0.02 * 0.7, 0.3, 1.7][:T]
    # oracle
    ous = [t.clone().requires_grad_(True) for t in us]
    ov = v.clone().requires_grad_(True)
    st = {f"u2x.m.{k}": p.clone().requires_grad_(True) for k, p in params.items()}
    vv = ov if vidx is None else ov[vidx]
    ref_losses = []
    for t in range(T):
        lp = oc.DECODER_LOG_PROB[family](st, "u2x.m.", ous[t], vv, b, l, x)
        ref_losses.append(-lp.mean() if rws[t] is None else -(lp * rws[t].unsqueeze(1)).sum())
    ref_grads = torch.autograd.grad(sum(c * ls for c, ls in zip(coefs, ref_losses)), ous + [ov] + list(st.values()))
    # device
    d = lambda t: None if t is None else t.to(DEV)
    dus = [d(t).requires_grad_(True) for t in us]
    dv = d(v).requires_grad_(True)
    dp = [d(params[n]).requires_grad_(True) for n in PARAM_NAMES[family]]
    zi = dp[3] if len(dp) == 4 else None

    def run():
        losses = ops.decoder_nll_multi(family, dus, dv, dp[0], dp[1], dp[2], zi, b=d(b), l=d(l), x=d(x), vidx=d(vidx),
                                       reduce="mean", row_weights=[d(w) for w in rws], upstreams=coefs)
        grads = torch.autograd.grad(sum(c * ls for c, ls in zip(coefs, losses)), dus + [dv] + dp)
        return losses, grads

    losses, grads = run()
    for t in range(T):
        assert_close(losses[t], ref_losses[t], TOL, f"{family} loss of term {t}")
    names = [f"u{t}" for t in range(T)] + ["v"] + list(PARAM_NAMES[family])
    for n, g, rg in zip(names, grads, ref_grads):
        assert_close(g, rg, TOL, f"{family} multi d/d{n}")
    losses2, grads2 = run()
    assert all(torch.equal(a, b_) for a, b_ in zip(losses, losses2))
    assert all(torch.equal(a, b_) for a, b_ in zip(grads, grads2)), "run-to-run deterministic"
    # loss-only path (validation under no_grad)
    with torch.no_grad():
        val = ops.decoder_nll_multi(family, dus, dv, dp[0], dp[1], dp[2], zi, b=d(b), l=d(l), x=d(x), vidx=d(vidx),
                                    reduce="mean", row_weights=[d(w) for w in rws], upstreams=coefs)
    for t in range(T):
        assert_close(val[t], ref_losses[t], TOL, f"{family} no-grad loss of term {t}")


def test_decoder_multi_full_size(ops):
    """Three NB terms at the ATAC shape of BASELINE configs 1 / 2 (B = 128, F = 50 000) against three
    single-term calls of the same kernels and against the oracle."""
    family, B, F, K = "NB", 128, 50000, 50
    u, v, b, l, x, params, vidx = _decoder_case(family, B, F, K, 1, seed=7)
    gen = torch.Generator().manual_seed(3)
    us = [u, torch.randn(B, K, generator=gen) * 0.5, torch.randn(B, K, generator=gen) * 0.5]
    masks = [(torch.rand(B, generator=gen) < 0.7).float() for _ in range(3)]
    rws = [m / (m.sum() * F) for m in masks]
    coefs = [1.0, 0.02, 0.02]
    d = lambda t: None if t is None else t.to(DEV)
    dus = [d(t).requires_grad_(True) for t in us]
    dv = d(v).requires_grad_(True)
    dp = [d(params[n]).requires_grad_(True) for n in PARAM_NAMES[family]]
    losses = ops.decoder_nll_multi(family, dus, dv, dp[0], dp[1], dp[2], None, b=d(b), l=d(l), x=d(x),
                                   reduce="mean", row_weights=[d(w) for w in rws], upstreams=coefs)
    grads = torch.autograd.grad(sum(c * ls for c, ls in zip(coefs, losses)), dus + [dv] + dp)
    singles = [ops.decoder_nll(family, dus[t], dv, dp[0], dp[1], dp[2], None, b=d(b), l=d(l), x=d(x), reduce="mean",
                               row_weight=d(rws[t])) for t in range(3)]
    sgrads = torch.autograd.grad(sum(c * ls for c, ls in zip(coefs, singles)), dus + [dv] + dp)
    for t in range(3):
        assert_close(losses[t], singles[t], 1e-5, f"loss of term {t} vs single-term call")
    for g, sg in zip(grads, sgrads):
        assert_close(g, sg, 2e-5, "fused vs term-by-term gradients")
    ous = [t.clone().requires_grad_(True) for t in us]
    ov = v.clone().requires_grad_(True)
    st = {f"u2x.m.{k}": p.clone().requires_grad_(True) for k, p in params.items()}
    ref = []
    for t in range(3):
        lp = oc.DECODER_LOG_PROB[family](st, "u2x.m.", ous[t], ov, b, l, x)
        ref.append(-(lp * rws[t].unsqueeze(1)).sum())
    ref_grads = torch.autograd.grad(sum(c * ls for c, ls in zip(coefs, ref)), ous + [ov] + list(st.values()))
    for t in range(3):
        assert_close(losses[t], ref[t], TOL, f"loss of term {t} vs oracle")
    for g, rg in zip(grads, ref_grads):
        assert_close(g, rg, TOL, "full-size multi-term gradient vs oracle")


def test_decoder_rejects_cpu_tensors(ops):
    u, v = torch.randn(4, 50), torch.randn(6, 50)
    p = torch.zeros(1, 6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.decoder_nll("NB", u, v, p, p, p, l=torch.ones(4, 1), x=torch.ones(4, 6))


# ------------------------------------------------------------------ graph encoder head
@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("V,K,with_noise", [(37, 50, True), (5000, 50, True), (52000, 50, True),
                                            (1000, 16, False), (777, 64, True), (300, 2, True)])
def test_graph_encoder_head_vs_torch(ops, V, K, with_noise, impl, monkeypatch):
    """Fused Gaussian head (Linear x2 + softplus + sample + KL) against the ATen formulation of
    scglue/models/sc.py:44-46 and scglue/models/scglue.py:329-340, values and every gradient;
    ``impl``: backward pass with the FFMA kernel (1) / on warp-level tensor cores (2, the default)."""
    monkeypatch.setattr(ops, "GRAPHENC_IMPL", impl)
    import torch.distributions as D
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(V + K)
    ptr_ = torch.randn(V, K, generator=gen) * 0.7
    wl, ws = torch.randn(K, K, generator=gen) * 0.2, torch.randn(K, K, generator=gen) * 0.2
    bl, bs = torch.randn(K, generator=gen) * 0.1, torch.randn(K, generator=gen) * 0.1
    noise = torch.randn(V, K, generator=gen) if with_noise else None
    gv = torch.randn(V, K, generator=gen)
    ckl = 0.37

    def run(dev, fused):
        args = [t.to(dev).requires_grad_(True) for t in (ptr_, wl, bl, ws, bs)]
        nz = None if noise is None else noise.to(dev)
        if fused:
            vsamp, kl, mu, sigma = ops.graph_encoder_head(*args, nz)
        else:
            p, a, b, c, d = args
            q = D.Normal(F.linear(p, a, b), F.softplus(F.linear(p, c, d)) + 1e-7, validate_args=False)
            mu, sigma = q.mean, q.stddev
            vsamp = mu if nz is None else mu + nz * sigma
            kl = D.kl_divergence(q, D.Normal(torch.zeros((), device=dev), torch.ones((), device=dev))).sum()
        loss = (vsamp * gv.to(dev)).sum() + ckl * kl
        grads = torch.autograd.grad(loss, args)
        return [vsamp, kl, mu, sigma] + list(grads)

    got, want = run(DEV, True), run("cpu", False)
    for name, g, w in zip(["vsamp", "kl", "mu", "sigma", "d/dptr", "d/dW_loc", "d/db_loc", "d/dW_std",
                           "d/db_std"], got, want):
        assert_close(g, w, 2e-5, name)
    again = run(DEV, True)
    assert all(torch.equal(a, b) for a, b in zip(got, again)), "graph encoder head must be deterministic"


# ------------------------------------------------------------------ fused optimizer
def test_flat_rmsprop_matches_torch_rmsprop(native_lib):
    """FlatRMSprop (one fused kernel over flat buffers) against torch.optim.RMSprop over several
    steps: a parameter without gradient is skipped, the learning rate changes mid-way, and the
    state survives a state_dict round trip."""
    from glue_b200.optim import FlatRMSprop
    gen = torch.Generator().manual_seed(3)
    shapes = [(52000, 50), (50, 50), (50,), (3, 7), (1, 50001), (5,)]
    init = [torch.randn(s, generator=gen) for s in shapes]
    ref_p = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    our_p = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    ref = torch.optim.RMSprop(ref_p, lr=2e-3)
    ours = FlatRMSprop(our_p, lr=2e-3)
    for step in range(6):
        grads = [torch.randn(s, generator=gen).to(DEV) * (10.0 ** (step - 3)) for s in shapes]
        for ps in (ref_p, our_p):
            for i, (p, g) in enumerate(zip(ps, grads)):
                p.grad = None if (i == 3 and step % 2) else g.clone()
        if step == 3:
            for opt in (ref, ours):
                opt.param_groups[0]["lr"] = 5e-4
        if step == 4:  # checkpoint round trip (EarlyStopping restores optimizer state this way)
            ours.load_state_dict(ours.state_dict())
        ref.step()
        ours.step()
        ours.zero_grad()
        for a, b in zip(our_p, ref_p):
            assert_close(a, b, 1e-6, f"parameters after step {step}")
    for a, b in zip(our_p, ref_p):
        assert_close(ours.state[a]["square_avg"], ref.state[b]["square_avg"], 1e-6, "square_avg")
    # parameters are views of one flat buffer, and survive being re-allocated by .to()
    assert all(p.data_ptr() == v.data_ptr() for p, v in zip(our_p, ours._pviews))
    our_p[1].data = our_p[1].data.clone()
    our_p[1].grad = torch.ones_like(our_p[1])
    ours.step()
    assert our_p[1].data_ptr() == ours._pviews[1].data_ptr()


# ------------------------------------------------------------------ KL of the data encoders
@pytest.mark.parametrize("B,K", [(128, 50), (1, 2), (37, 64), (1000, 50)])
def test_kl_unit_normal_vs_torch_distributions(ops, B, K):
    import torch.distributions as D
    gen = torch.Generator().manual_seed(B + K)
    mean = torch.randn(B, K, generator=gen).requires_grad_(True)
    std = (torch.rand(B, K, generator=gen) * 2 + 1e-3).requires_grad_(True)
    n_features = 2000
    ref = D.kl_divergence(D.Normal(mean, std), D.Normal(torch.tensor(0.0), torch.tensor(1.0))) \
        .sum(dim=1).mean() / n_features
    rg = torch.autograd.grad(3.0 * ref, [mean, std])
    dm, ds = mean.detach().to(DEV).requires_grad_(True), std.detach().to(DEV).requires_grad_(True)
    out = ops.kl_unit_normal(dm, ds, 1.0 / (B * n_features))
    g = torch.autograd.grad(3.0 * out, [dm, ds])
    assert_close(out, ref, TOL, "kl")
    assert_close(g[0], rg[0], TOL, "d kl / d mean")
    assert_close(g[1], rg[1], TOL, "d kl / d std")
    out2 = ops.kl_unit_normal(dm, ds, 1.0 / (B * n_features))
    assert torch.equal(out, out2), "fixed summation order"


# ------------------------------------------------------------------ paired cells: mean + cosine term
@pytest.mark.parametrize("M,B,K,empty_mod", [(2, 128, 50, False), (3, 37, 64, False), (2, 16, 8, True),
                                             (2, 1000, 50, False)])
def test_paired_mean_cos_vs_torch(ops, M, B, K, empty_mod):
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(M * 1000 + B)
    ustack = torch.randn(M, B, K, generator=gen).requires_grad_(True)
    pf = (torch.rand(M, B, generator=gen) < 0.7).float()
    pf[0, pf.sum(0) == 0] = 1.0  # every cell is observed somewhere
    if empty_mod:
        pf[1] = 0.0              # a modality without cells in this minibatch: its term is skipped
        pf[0] = 1.0
    w = torch.randn(B, K, generator=gen)  # a consumer of umean (the joint-cross decoder call)

    def reference(us):
        pstack = pf.unsqueeze(2).expand_as(us)
        umean = (us * pstack).sum(dim=0) / pstack.sum(dim=0)
        cos = torch.zeros(())
        for i in range(M):
            n = pf[i].sum()
            sim = (F.cosine_similarity(us[i], umean) * pf[i]).sum() / n.clamp(min=1.0)
            cos = cos + torch.where(n > 0, 1 - sim, torch.zeros(()))
        return umean, cos

    umean_r, cos_r = reference(ustack)
    (g_r,) = torch.autograd.grad((umean_r * w).sum() + 0.7 * cos_r, [ustack])
    du = ustack.detach().to(DEV).requires_grad_(True)
    umean, cos = ops.paired_mean_cos(du, pf.to(DEV))
    (g,) = torch.autograd.grad((umean * w.to(DEV)).sum() + 0.7 * cos, [du])
    assert_close(umean, umean_r, TOL, "umean")
    assert_close(cos, cos_r, TOL, "cos_loss")
    assert_close(g, g_r, TOL, "d / d ustack")
    # only one of the two outputs used downstream
    (g_only_cos,) = torch.autograd.grad(ops.paired_mean_cos(du, pf.to(DEV))[1], [du])
    (g_only_cos_r,) = torch.autograd.grad(reference(ustack)[1], [ustack])
    if empty_mod:  # umean == ustack[0]: the cosine is 1 and its gradient vanishes (rounding noise only)
        assert g_only_cos.abs().max().item() < 1e-5 and g_only_cos_r.abs().max().item() < 1e-5
    else:
        assert_close(g_only_cos, g_only_cos_r, TOL, "d cos / d ustack")


# ------------------------------------------------------------------ hidden block of the data encoders
@pytest.mark.parametrize("B,H", [(128, 256), (37, 70), (2, 32), (300, 256)])
def test_act_bn_dropout_without_dropout_equals_torch_modules(ops, B, H):
    gen = torch.Generator().manual_seed(B + H)
    z = torch.randn(B, H, generator=gen) * 2
    ref_bn, bn = torch.nn.BatchNorm1d(H), torch.nn.BatchNorm1d(H).to(DEV)
    with torch.no_grad():
        for m in (ref_bn, bn):
            m.weight.copy_(torch.rand(H, generator=torch.Generator().manual_seed(1)) + 0.5)
            m.bias.copy_(torch.randn(H, generator=torch.Generator().manual_seed(2)))
            m.running_mean.copy_(torch.randn(H, generator=torch.Generator().manual_seed(3)))
    w = torch.randn(B, H, generator=gen)
    zr = z.clone().requires_grad_(True)
    yr = ref_bn(torch.nn.functional.leaky_relu(zr, 0.2))
    gr = torch.autograd.grad((yr * w).sum(), [zr, ref_bn.weight, ref_bn.bias])
    zd = z.to(DEV).requires_grad_(True)
    y = ops.act_bn_dropout(zd, bn, 0.0, None)
    g = torch.autograd.grad((y * w.to(DEV)).sum(), [zd, bn.weight, bn.bias])
    assert_close(y, yr, 1e-5, "block output")
    for name, a, b in zip(["z", "gamma", "beta"], g, gr):
        assert_close(a, b, 1e-4, f"d / d {name}")
    assert_close(bn.running_mean, ref_bn.running_mean, 1e-5, "running_mean")
    assert_close(bn.running_var, ref_bn.running_var, 1e-5, "running_var")
    assert int(bn.num_batches_tracked) == int(ref_bn.num_batches_tracked) == 1


def test_act_bn_dropout_masks(ops):
    B, H, p = 256, 256, 0.2
    z = torch.randn(B, H, generator=torch.Generator().manual_seed(0)).to(DEV).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(H).to(DEV)
    state = ops.new_dropout_state(torch.device(DEV), generator=torch.Generator().manual_seed(5))
    out, mean, rstd, keep = ops.act_bn_dropout(z, bn, p, state, return_aux=True)
    assert state.tolist()[1:] == [1, 0], "call counter advanced, ticket back at zero"
    frac = keep.float().mean().item()
    assert abs(frac - (1 - p)) < 0.01, frac
    # kept entries are the un-dropped block output scaled by 1 / (1 - p), dropped ones are zero
    bn2 = torch.nn.BatchNorm1d(H).to(DEV)
    full = bn2(torch.nn.functional.leaky_relu(z.detach(), 0.2))
    assert_close(out, full * keep / (1 - p), 1e-5, "masked output")
    # backward uses the same mask
    w = torch.randn(B, H, generator=torch.Generator().manual_seed(1)).to(DEV)
    (g,) = torch.autograd.grad((out * w).sum(), [z])
    z2 = z.detach().clone().requires_grad_(True)
    bn3 = torch.nn.BatchNorm1d(H).to(DEV)
    (g_ref,) = torch.autograd.grad((bn3(torch.nn.functional.leaky_relu(z2, 0.2)) * keep / (1 - p) * w).sum(), [z2])
    assert_close(g, g_ref, 1e-4, "d / d z under the mask")
    # a second call draws a different mask; the same state replays the same stream
    _, _, _, keep_b = ops.act_bn_dropout(z, bn, p, state, return_aux=True)
    assert not torch.equal(keep, keep_b)
    state2 = ops.new_dropout_state(torch.device(DEV), generator=torch.Generator().manual_seed(5))
    _, _, _, keep_c = ops.act_bn_dropout(z, bn, p, state2, return_aux=True)
    assert torch.equal(keep, keep_c)
    # columns and rows are not correlated in an obvious way
    assert abs(keep.float().mean(0).std().item() - (p * (1 - p) / B) ** 0.5) < 0.01


# ------------------------------------------------------------------ whole data-encoder MLP
def _make_encoders(in_dim, h_dim, depth, p, n):
    """``n`` data encoders with identical weights (non-trivial BatchNorm affine and running statistics)."""
    from glue_b200.models import sc
    torch.manual_seed(in_dim + h_dim + depth)
    encs = [sc.VanillaDataEncoder(in_dim, 50, h_depth=depth, h_dim=h_dim, dropout=p)]
    with torch.no_grad():
        for i in range(depth):
            bn = getattr(encs[0], f"bn_{i}")
            bn.weight.copy_(torch.rand(h_dim) + 0.5), bn.bias.copy_(torch.randn(h_dim) * 0.3)
            bn.running_mean.copy_(torch.randn(h_dim) * 0.1)
    for _ in range(n - 1):
        other = sc.VanillaDataEncoder(in_dim, 50, h_depth=depth, h_dim=h_dim, dropout=p)
        other.load_state_dict(encs[0].state_dict())
        encs.append(other)
    return encs


def _encoder_run(enc, x, wl, ws):
    u, _ = enc(x, x)
    loss = (u.mean * wl).sum() + (u.stddev * ws).sum()
    names = [n for n, _ in enc.named_parameters()]
    grads = torch.autograd.grad(loss, [p for _, p in enc.named_parameters()])
    return u.mean, u.stddev, dict(zip(names, grads))


def _assert_encoder_grads(got, want, tol=1e-4):
    for name in want:
        if name.startswith("linear_") and name.endswith(".bias"):
            # a bias in front of BatchNorm has gradient zero up to rounding (the batch mean removes it):
            # compared on the scale of the layer's weight gradient
            scale = want[name.replace(".bias", ".weight")].abs().max().item()
            assert (got[name].cpu() - want[name].cpu()).abs().max().item() <= tol * scale, f"d / d {name}"
        else:
            assert_close(got[name], want[name], tol, f"d / d {name}")


@pytest.mark.parametrize("B,in_dim,h_dim,depth", [(128, 100, 256, 2), (77, 100, 256, 2), (128, 52, 64, 1),
                                                  (128, 256, 256, 3), (8, 100, 32, 2)])
def test_fused_encoder_mlp_vs_torch_modules(ops, B, in_dim, h_dim, depth):
    """Whole-MLP kernels (3 launches each way) against the reference formulation with torch modules on the
    CPU (scglue/models/sc.py:136-178), dropout off: outputs, every parameter gradient, BatchNorm state."""
    from glue_b200.utils import config
    ref, enc = _make_encoders(in_dim, h_dim, depth, 0.0, 2)
    enc = enc.to(DEV)
    gen = torch.Generator().manual_seed(B)
    x = torch.randn(B, in_dim, generator=gen)
    wl, ws = torch.randn(B, 50, generator=gen), torch.randn(B, 50, generator=gen)
    want = _encoder_run(ref, x, wl, ws)
    launches = ops.kernel_launches()
    assert config.USE_FUSED_ENCODER
    got = _encoder_run(enc, x.to(DEV), wl.to(DEV), ws.to(DEV))
    assert ops.kernel_launches() - launches == 2 * (depth + 1), "h_depth + 1 launches each way"
    assert_close(got[0], want[0], 1e-5, "loc")
    assert_close(got[1], want[1], 1e-5, "std")
    _assert_encoder_grads(got[2], want[2])
    for i in range(depth):
        a, b = getattr(enc, f"bn_{i}"), getattr(ref, f"bn_{i}")
        assert_close(a.running_mean, b.running_mean, 1e-5, "running_mean")
        assert_close(a.running_var, b.running_var, 1e-5, "running_var")
        assert int(a.num_batches_tracked) == int(b.num_batches_tracked) == 1
    again = _encoder_run(_make_encoders(in_dim, h_dim, depth, 0.0, 1)[0].to(DEV), x.to(DEV), wl.to(DEV), ws.to(DEV))
    assert torch.equal(again[0], got[0]) and all(torch.equal(again[2][n], got[2][n]) for n in got[2]), \
        "fixed-order reductions: bit-reproducible"


def _encoder_fp64(enc, x, keeps, p, wl, ws):
    """float64 restatement of the encoder forward under given keep masks; returns loc, std, parameter gradients."""
    import torch.nn.functional as F
    params = {n: q.detach().double().requires_grad_(True) for n, q in enc.named_parameters()}
    h = x.double()
    for i, keep in enumerate(keeps):
        bn = getattr(enc, f"bn_{i}")
        a = F.leaky_relu(h @ params[f"linear_{i}.weight"].T + params[f"linear_{i}.bias"], 0.2)
        y = (a - a.mean(0)) / torch.sqrt(a.var(0, unbiased=False) + bn.eps)
        h = (y * params[f"bn_{i}.weight"] + params[f"bn_{i}.bias"]) * keep.double() / (1 - p)
    loc = h @ params["loc.weight"].T + params["loc.bias"]
    std = F.softplus(h @ params["std_lin.weight"].T + params["std_lin.bias"]) + 1e-7
    grads = torch.autograd.grad((loc * wl.double()).sum() + (std * ws.double()).sum(), list(params.values()))
    return loc, std, dict(zip(params, grads))


@pytest.mark.parametrize("B", [128, 100])
def test_fused_encoder_mlp_with_dropout(ops, B, monkeypatch):
    """With dropout on, the fused kernels and the per-op path (torch Linear + the hidden-block kernel) draw the
    same Philox masks from the same state.  Under those masks the fused outputs and gradients match a float64
    restatement to 1e-4 (the per-op path, cuBLAS products, is checked against the same restatement), the call
    counters advance alike; under no_grad (discriminator pass) the fused path gives the same values."""
    from glue_b200.utils import config
    p = 0.2
    fused, per_op = (e.to(DEV) for e in _make_encoders(100, 256, 2, p, 2))
    dev = torch.device(DEV)
    for i in range(2):  # same seeds
        per_op._dropout_state(i, dev).copy_(fused._dropout_state(i, dev))
    gen = torch.Generator().manual_seed(B)
    x = torch.randn(B, 100, generator=gen).to(DEV)
    wl, ws = torch.randn(B, 50, generator=gen).to(DEV), torch.randn(B, 50, generator=gen).to(DEV)
    for step in range(2):
        keeps = []
        for i in range(2):  # the masks of this call: a function of (seed, call, row, column) only
            state = fused._dropout_state(i, dev).clone()
            keeps.append(ops.act_bn_dropout(torch.zeros(B, 256, device=DEV), torch.nn.BatchNorm1d(256).to(DEV), p,
                                            state, return_aux=True)[3])
        ref = _encoder_fp64(fused, x, keeps, p, wl, ws)
        got = _encoder_run(fused, x, wl, ws)
        monkeypatch.setattr(config, "USE_FUSED_ENCODER", False)
        other = _encoder_run(per_op, x, wl, ws)
        monkeypatch.setattr(config, "USE_FUSED_ENCODER", True)
        ref_g = {n: g.float() for n, g in ref[2].items()}
        assert_close(got[0], ref[0].float(), 1e-5, "loc")
        assert_close(got[1], ref[1].float(), 1e-5, "std")
        _assert_encoder_grads(got[2], ref_g)
        # (the per-op path: same masks, hence the same values; its cuBLAS weight gradients are only good to a
        # few 1e-4 of the float64 result at B = 100, the fused kernels to 1e-6)
        assert_close(other[0], ref[0].float(), 1e-5, "per-op loc")
        _assert_encoder_grads(other[2], ref_g, tol=3e-3)
        for i in range(2):
            assert fused._dropout_state(i, dev).tolist() == per_op._dropout_state(i, dev).tolist()
    with torch.no_grad():
        u_f, _ = fused(x, x)
        monkeypatch.setattr(config, "USE_FUSED_ENCODER", False)
        u_p, _ = per_op(x, x)
    assert_close(u_f.mean, u_p.mean, 1e-5, "loc under no_grad")
    assert_close(u_f.stddev, u_p.stddev, 1e-5, "std under no_grad")
    fused.eval()  # BatchNorm on its running statistics: the torch modules take over
    monkeypatch.setattr(config, "USE_FUSED_ENCODER", True)
    launches = ops.kernel_launches()
    fused(x, x)
    assert ops.kernel_launches() == launches


# ------------------------------------------------------------------ discriminator MLP
@pytest.mark.parametrize("rows,in_dim,n_batches,n_out,p", [(256, 50, 0, 2, 0.0), (384, 50, 0, 3, 0.0),
                                                           (256, 50, 20, 2, 0.0), (100, 50, 0, 2, 0.0),
                                                           (256, 50, 0, 2, 0.2), (300, 50, 20, 3, 0.2)])
def test_fused_discriminator_vs_torch_modules(ops, rows, in_dim, n_batches, n_out, p, monkeypatch):
    """Discriminator blocks + `pred` as h_depth + 1 launches each way (row tiles of 128, the 50-wide input staged
    without bulk copies) against a float64 restatement of scglue/models/sc.py:576-624 under the same dropout masks:
    logits, input gradient (the alignment term back-propagates into the encoders) and parameter gradients."""
    import torch.nn.functional as F
    from glue_b200.models import sc
    from glue_b200.utils import config
    torch.manual_seed(rows + n_out)
    dsc = sc.Discriminator(in_dim, n_out, n_batches=n_batches, h_depth=2, h_dim=256, dropout=p).to(DEV)
    dev = torch.device(DEV)
    gen = torch.Generator().manual_seed(rows)
    x = (torch.randn(rows, in_dim, generator=gen)).to(DEV).requires_grad_(True)
    b = torch.randint(0, max(n_batches, 1), (rows,), generator=gen).to(DEV)
    wout = torch.randn(rows, n_out, generator=gen).to(DEV)
    width = in_dim + n_batches
    keeps = []
    for i in range(2):
        if p > 0:
            state = dsc._dropout_state(i, dev).clone()
            keeps.append(ops.act_bn_dropout(torch.zeros(rows, 256, device=DEV), torch.nn.BatchNorm1d(256).to(DEV), p,
                                            state, return_aux=True)[3].double())
        else:
            keeps.append(torch.ones(rows, 256, device=DEV, dtype=torch.float64))
    params = {n: q.detach().double().requires_grad_(True) for n, q in dsc.named_parameters()}
    x64 = x.detach().double().requires_grad_(True)
    h = torch.cat([x64, F.one_hot(b, num_classes=n_batches).double()], dim=1) if n_batches else x64
    assert h.shape[1] == width
    for i in range(2):
        h = F.leaky_relu(h @ params[f"linear_{i}.weight"].T + params[f"linear_{i}.bias"], 0.2) * keeps[i] / (1 - p)
    ref = h @ params["pred.weight"].T + params["pred.bias"]
    ref_g = torch.autograd.grad((ref * wout.double()).sum(), [x64] + list(params.values()))
    launches = ops.kernel_launches()
    out = dsc(x, b)
    got_g = torch.autograd.grad((out * wout).sum(), [x] + list(dsc.parameters()))
    assert ops.kernel_launches() - launches == 6 + (2 if p > 0 else 0) * 0, "three launches each way"
    assert_close(out, ref.float(), 1e-5, "logits")
    for name, a, r in zip(["x"] + list(params), got_g, ref_g):
        assert_close(a, r.float(), 1e-4, f"d / d {name}")
    if p > 0:
        assert dsc._dropout_state(0, dev).tolist()[1:] == [1, 0]
    # the per-op path (torch modules) gives the same values when nothing is dropped
    if p == 0:
        monkeypatch.setattr(config, "USE_FUSED_ENCODER", False)
        assert_close(dsc(x, b), out, 1e-5, "per-op logits")


def test_multi_copy(ops):
    """One launch moves a whole minibatch into the static input buffers of a captured step: mixed dtypes, sizes
    from 1 byte to tens of MB, misaligned views, more than 16 tensors."""
    gen = torch.Generator().manual_seed(0)
    big = torch.randn(128, 50000, generator=gen).to(DEV)
    srcs = [big, torch.randint(0, 9, (2, 35861), generator=gen).to(DEV), torch.rand(128, generator=gen).to(DEV) > 0.5,
            torch.randn(1001, generator=gen).to(DEV)[1:], torch.randn(3, generator=gen).to(DEV)[1:2],
            torch.randint(0, 255, (7,), generator=gen).to(torch.uint8).to(DEV)]
    srcs += [torch.randn(37 + i, generator=gen).to(DEV) for i in range(14)]  # 20 tensors: two launches
    dsts = [torch.zeros_like(s).contiguous() for s in srcs]
    srcs = [s.contiguous() if i != 3 else s for i, s in enumerate(srcs)]  # (index 3 stays a 4-byte-offset view)
    launches = ops.kernel_launches()
    ops.multi_copy(dsts, srcs)
    assert ops.kernel_launches() - launches == 2
    for d, s in zip(dsts, srcs):
        assert torch.equal(d, s)


def test_flat_rmsprop_partial_exchange(native_lib):
    """The graph branch's gradients are exchanged ahead of the rest (all_reduce_subset on a contiguous run of
    parameters, all_reduce_mean(skip=...) for the others): the flat buffer ends up holding every gradient, the
    update equals the one of a plain step (world size 1: the exchanges themselves are no-ops)."""
    from glue_b200.optim import FlatRMSprop
    gen = torch.Generator().manual_seed(11)
    shapes = [(500, 50), (50, 50), (50,), (7, 3), (9,)]
    init = [torch.randn(s, generator=gen) for s in shapes]
    grads = [torch.randn(s, generator=gen).to(DEV) for s in shapes]
    outs = []
    for split in (False, True):
        ps = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
        opt = FlatRMSprop(ps, lr=2e-3)
        for p, g in zip(ps[:3], grads[:3]):
            p.grad = g.clone()
        skip = opt.all_reduce_subset(ps[:3]) if split else None
        if split:
            assert skip == (0, opt._offsets[3])
            assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(ps[:3], opt._gviews[:3]))
            assert opt.all_reduce_subset([ps[0], ps[2]]) is None  # not contiguous: nothing done
        for p, g in zip(ps[3:], grads[3:]):
            p.grad = g.clone()
        opt.all_reduce_mean(skip=skip)
        for v, g in zip(opt._gviews, grads):
            assert torch.equal(v, g)
        opt.step()
        outs.append([p.detach().clone() for p in ps])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n,c", [(256, 2), (100, 3), (7, 32)])
def test_weighted_cross_entropy_matches_torch(ops, n, c):
    """Discriminator loss (scglue/models/scglue.py:326-329) and its gradient from one launch: against
    F.cross_entropy * weight, summed, over the number of cells; with the coefficient folded the stored gradient
    is that multiple of torch's and ignores the incoming one; under no_grad nothing is stored."""
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(n + c)
    logits = (4 * torch.randn(n, c, generator=gen)).to(DEV)
    target = torch.randint(0, c, (n,), generator=gen).to(DEV)
    weight = torch.rand(n, generator=gen).to(DEV)
    ref_in = logits.double().requires_grad_(True)
    ref = (F.cross_entropy(ref_in, target, reduction="none") * weight.double()).sum() / n
    ref.backward()
    x = logits.clone().requires_grad_(True)
    got = ops.weighted_cross_entropy(x, target, weight)
    (got * 3.0).backward()
    assert_close(got.detach(), ref.detach().float(), 1e-6, "loss")
    assert_close(x.grad, 3.0 * ref_in.grad.float(), 1e-5, "d / d logits")
    y = logits.clone().requires_grad_(True)
    folded = ops.weighted_cross_entropy(y, target, weight, upstream=-0.05)
    assert torch.equal(folded, got.detach())
    folded.backward()  # (the caller back-propagates -0.05 * folded; what arrives here is ignored)
    assert_close(y.grad, -0.05 * ref_in.grad.float(), 1e-5, "folded d / d logits")
    with torch.no_grad():
        assert torch.equal(ops.weighted_cross_entropy(logits, target, weight, upstream=1.0), got.detach())
    with pytest.raises(RuntimeError, match="UNSUPPORTED|unsupported|-2"):
        ops.weighted_cross_entropy(torch.zeros(4, 33, device=DEV), target[:4] * 0, weight[:4])
