"""CPU oracle for the SCGLUE per-step training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``glue_b200/`` may import this file:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline.

It restates, as plain functions over a flat ``state`` dict of fp32 CPU tensors
(the keys are the reference's ``state_dict`` names), the arithmetic that the
reference spreads over ``torch.nn.Module`` classes.  Every function names the
reference lines it follows (paths relative to the reference checkout).  The
op *sequence* is kept ATen-for-ATen (gather + ``scatter_add_``, materialised
``[B, F]`` temporaries, ``torch.distributions`` log-probs) so that timing this
file on host cores is a fair stand-in for timing the reference's own CPU path.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified
reference (``oracle/refstub.py``) in the build container, runs both on the same
seeded inputs and stores the reference's outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those fixtures and
against the reference's own known-answer test for edge normalisation
(``tests/test_num.py:76-113``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
import torch.distributions as D
import torch.nn.functional as F

EPS = 1e-7  # scglue/num.py:12

Tensor = torch.Tensor
State = Mapping[str, Tensor]


# --------------------------------------------------------------------------
# a1  GraphConv                                      scglue/models/nn.py:26-57
# --------------------------------------------------------------------------
def graph_conv(inp: Tensor, eidx: Tensor, enorm: Tensor, esgn: Tensor) -> Tensor:
    src, dst = eidx[0], eidx[1]
    msg = inp[src] * (esgn * enorm).unsqueeze(1)
    out = torch.zeros_like(inp)
    out.scatter_add_(0, dst.unsqueeze(1).expand_as(msg), msg)
    return out


# --------------------------------------------------------------------------
# a2  GraphEncoder                                   scglue/models/sc.py:21-46
# --------------------------------------------------------------------------
def graph_encoder(state: State, eidx: Tensor, enorm: Tensor, esgn: Tensor,
                  prefix: str = "g2v.") -> Tuple[Tensor, Tensor]:
    """Returns (loc, std) of the vertex posterior."""
    ptr = graph_conv(state[prefix + "vrepr"], eidx, enorm, esgn)
    loc = F.linear(ptr, state[prefix + "loc.weight"], state[prefix + "loc.bias"])
    std = F.softplus(
        F.linear(ptr, state[prefix + "std_lin.weight"], state[prefix + "std_lin.bias"])
    ) + EPS
    return loc, std


# --------------------------------------------------------------------------
# a3  GraphDecoder + Bernoulli log-prob + pos/neg balanced mean
#     scglue/models/sc.py:54-59, scglue/models/scglue.py:331-338
# --------------------------------------------------------------------------
def graph_decoder_logits(v: Tensor, eidx: Tensor, esgn: Tensor) -> Tensor:
    return esgn * (v[eidx[0]] * v[eidx[1]]).sum(dim=1)


def graph_nll(v: Tensor, eidx: Tensor, esgn: Tensor, ewt: Tensor) -> Tensor:
    nll = -D.Bernoulli(logits=graph_decoder_logits(v, eidx, esgn)).log_prob(ewt)
    pos = (ewt != 0).to(torch.int64)
    n_pos = int(pos.sum().item())
    n_neg = pos.numel() - n_pos
    bins = torch.zeros(2, dtype=nll.dtype, device=nll.device)
    bins.scatter_add_(0, pos, nll)
    avgc = (n_pos > 0) + (n_neg > 0)
    return (bins[0] / max(n_neg, 1) + bins[1] / max(n_pos, 1)) / avgc


def kl_normal_std(loc: Tensor, std: Tensor) -> Tensor:
    """KL(N(loc,std) || N(0,1)) elementwise (torch kl._kl_normal_normal)."""
    return D.kl_divergence(D.Normal(loc, std), D.Normal(torch.zeros(()), torch.ones(())))


# --------------------------------------------------------------------------
# a4-a6  data decoders                    scglue/models/sc.py:280-478, prob.py
# --------------------------------------------------------------------------
def _affine_logits(state: State, prefix: str, u: Tensor, v: Tensor, b: Tensor,
                   bias_key: str = "bias") -> Tensor:
    scale = F.softplus(state[prefix + "scale_lin"][b])
    return scale * (u @ v.t()) + state[prefix + bias_key][b]


def nb_params(state: State, prefix: str, u, v, b, l) -> Tuple[Tensor, Tensor]:
    """sc.py:390-397 -> (total_count, nb_logits)."""
    mu = F.softmax(_affine_logits(state, prefix, u, v, b), dim=1) * l
    log_theta = state[prefix + "log_theta"][b]
    return log_theta.exp(), (mu + EPS).log() - log_theta


def nb_log_prob(state: State, prefix: str, u, v, b, l, x) -> Tensor:
    total_count, logits = nb_params(state, prefix, u, v, b, l)
    return D.NegativeBinomial(total_count, logits=logits, validate_args=False).log_prob(x)


def nb_mean(state: State, prefix: str, u, v, b, l) -> Tensor:
    total_count, logits = nb_params(state, prefix, u, v, b, l)
    return total_count * torch.exp(logits)


def zinb_log_prob(state: State, prefix: str, u, v, b, l, x) -> Tensor:
    """sc.py:465-478 + prob.py:144-153 (boolean-mask formulation kept)."""
    total_count, logits = nb_params(state, prefix, u, v, b, l)
    zi = state[prefix + "zi_logits"][b].expand_as(logits)
    raw = D.NegativeBinomial(total_count, logits=logits, validate_args=False).log_prob(x)
    out = torch.empty_like(raw)
    zmask = x.abs() < EPS
    z_zi, nz_zi = zi[zmask], zi[~zmask]
    out[zmask] = (raw[zmask].exp() + z_zi.exp() + EPS).log() - F.softplus(z_zi)
    out[~zmask] = raw[~zmask] - F.softplus(nz_zi)
    return out


def normal_params(state: State, prefix: str, u, v, b) -> Tuple[Tensor, Tensor]:
    """sc.py:298-308 / 359-369 -> (loc, std)."""
    loc = _affine_logits(state, prefix, u, v, b)
    std = F.softplus(state[prefix + "std_lin"][b]) + EPS
    return loc, std


def normal_log_prob(state: State, prefix: str, u, v, b, l, x) -> Tensor:
    loc, std = normal_params(state, prefix, u, v, b)
    return D.Normal(loc, std, validate_args=False).log_prob(x)


def ziln_log_prob(state: State, prefix: str, u, v, b, l, x) -> Tensor:
    """sc.py:359-369 + prob.py:110-118."""
    loc, std = normal_params(state, prefix, u, v, b)
    std = std.expand_as(loc)
    zi = state[prefix + "zi_logits"][b].expand_as(loc)
    out = torch.empty_like(x)
    zmask = x.abs() < EPS
    z_zi, nz_zi = zi[zmask], zi[~zmask]
    out[zmask] = z_zi - F.softplus(z_zi)
    out[~zmask] = D.LogNormal(loc[~zmask], std[~zmask], validate_args=False).log_prob(
        x[~zmask]) - F.softplus(nz_zi)
    return out


DECODER_LOG_PROB = {
    "NB": nb_log_prob,
    "ZINB": zinb_log_prob,
    "Normal": normal_log_prob,
    "ZILN": ziln_log_prob,
}


# --------------------------------------------------------------------------
# a7  data encoders                              scglue/models/sc.py:136-230
# --------------------------------------------------------------------------
def data_encoder(state: State, prefix: str, x: Tensor, xrep: Tensor, *, count_input: bool,
                 lazy_normalizer: bool, h_depth: int = 2, training: bool = True,
                 dropout_masks: Optional[Sequence[Tensor]] = None,
                 bn_momentum: float = 0.1, bn_eps: float = 1e-5,
                 ) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
    """Returns (loc, std, l).  ``count_input`` selects NBDataEncoder (library size
    + log1p CPM/100 normalisation, sc.py:222-230) over VanillaDataEncoder.
    ``dropout_masks`` are pre-scaled keep masks (one per hidden layer) or None
    for p = 0; batch-norm running stats in ``state`` are updated in place when
    ``training`` (as ``torch.nn.BatchNorm1d`` does)."""
    if xrep.numel():
        l = None if (lazy_normalizer or not count_input) else x.sum(dim=1, keepdim=True)
        ptr = xrep
    else:
        l = x.sum(dim=1, keepdim=True) if count_input else None
        ptr = (x * (1e4 / l)).log1p() if count_input else x
    for layer in range(h_depth):
        ptr = F.linear(ptr, state[f"{prefix}linear_{layer}.weight"],
                       state[f"{prefix}linear_{layer}.bias"])
        ptr = F.leaky_relu(ptr, negative_slope=0.2)
        ptr = F.batch_norm(
            ptr, state[f"{prefix}bn_{layer}.running_mean"],
            state[f"{prefix}bn_{layer}.running_var"],
            state[f"{prefix}bn_{layer}.weight"], state[f"{prefix}bn_{layer}.bias"],
            training=training, momentum=bn_momentum, eps=bn_eps)
        if training and f"{prefix}bn_{layer}.num_batches_tracked" in state:
            state[f"{prefix}bn_{layer}.num_batches_tracked"] += 1
        if dropout_masks is not None:
            ptr = ptr * dropout_masks[layer]
    loc = F.linear(ptr, state[prefix + "loc.weight"], state[prefix + "loc.bias"])
    std = F.softplus(
        F.linear(ptr, state[prefix + "std_lin.weight"], state[prefix + "std_lin.bias"])) + EPS
    return loc, std, l


# --------------------------------------------------------------------------
# a8  discriminator                              scglue/models/sc.py:576-624
# --------------------------------------------------------------------------
def discriminator(state: State, x: Tensor, b: Tensor, *, n_batches: int = 0, h_depth: int = 2,
                  dropout_masks: Optional[Sequence[Tensor]] = None, prefix: str = "du.") -> Tensor:
    if n_batches:
        x = torch.cat([x, F.one_hot(b, num_classes=n_batches)], dim=1)
    for layer in range(h_depth):
        x = F.leaky_relu(F.linear(x, state[f"{prefix}linear_{layer}.weight"],
                                  state[f"{prefix}linear_{layer}.bias"]), negative_slope=0.2)
        if dropout_masks is not None:
            x = x * dropout_masks[layer]
    return F.linear(x, state[prefix + "pred.weight"], state[prefix + "pred.bias"])


# --------------------------------------------------------------------------
# a9/a10  loss algebra of one training step    scglue/models/scglue.py:292-376
#                                               scglue/models/scglue.py:550-701
# --------------------------------------------------------------------------
class StepConfig:
    """Static description of a model (what the reference keeps on module objects)."""

    def __init__(self, keys: Sequence[str], prob_models: Mapping[str, str],
                 count_encoder: Mapping[str, bool], n_batches: Mapping[str, int],
                 du_n_batches: int = 0, h_depth: int = 2,
                 lam_data: float = 1.0, lam_kl: float = 1.0, lam_graph: float = 0.02,
                 lam_align: float = 0.05, lam_sup: float = 0.02, normalize_u: bool = False,
                 modality_weight: Optional[Mapping[str, float]] = None,
                 lam_joint_cross: float = 0.0, lam_real_cross: float = 0.0, lam_cos: float = 0.0,
                 paired: bool = False, align_burnin: Optional[float] = None,
                 burnin_noise_exag: float = 1.5):
        self.keys = list(keys)
        self.prob_models = dict(prob_models)
        self.count_encoder = dict(count_encoder)
        self.n_batches = dict(n_batches)
        self.du_n_batches = du_n_batches
        self.h_depth = h_depth
        self.lam_data, self.lam_kl, self.lam_graph = lam_data, lam_kl, lam_graph
        self.lam_align, self.lam_sup, self.normalize_u = lam_align, lam_sup, normalize_u
        mw = dict(modality_weight or {k: 1.0 for k in keys})
        norm = sum(mw.values()) / len(mw)  # glue.py:325-326
        self.modality_weight = {k: w / norm for k, w in mw.items()}
        self.lam_joint_cross, self.lam_real_cross, self.lam_cos = (
            lam_joint_cross, lam_real_cross, lam_cos)
        self.paired = paired
        self.align_burnin = align_burnin
        self.burnin_noise_exag = burnin_noise_exag


class Noise:
    """Explicit randomness of one ``compute_losses`` call.  ``None`` entries are
    drawn from the global torch RNG in the reference's order (usamp per key,
    burn-in noise, vsamp)."""

    def __init__(self, u_eps: Optional[Mapping[str, Tensor]] = None,
                 v_eps: Optional[Tensor] = None, burnin_eps: Optional[Tensor] = None,
                 enc_dropout: Optional[Mapping[str, Sequence[Tensor]]] = None,
                 du_dropout: Optional[Sequence[Tensor]] = None):
        self.u_eps, self.v_eps, self.burnin_eps = u_eps, v_eps, burnin_eps
        self.enc_dropout, self.du_dropout = enc_dropout, du_dropout


def _decoder_nll(cfg: StepConfig, state: State, k: str, u, v, b, l, x, nan_aware: bool) -> Tensor:
    lp = DECODER_LOG_PROB[cfg.prob_models[k]](state, f"u2x.{k}.", u, v, b, l, x)
    return -(lp.nanmean() if nan_aware else lp.mean())


def compute_losses(cfg: StepConfig, state: Dict[str, Tensor], batch: Mapping[str, object],
                   graph: Tuple[Tensor, Tensor, Tensor], epoch: int, *, dsc_only: bool = False,
                   noise: Optional[Noise] = None, training: bool = True) -> Dict[str, Tensor]:
    """``batch`` holds dicts ``x, xrep, xbch, xdwt`` keyed by modality, optional ``pmsk``
    ``[B, n_modalities]`` bool, and ``eidx, ewt, esgn`` of the sampled edge minibatch.
    ``graph`` = ``(eidx, enorm, esgn)`` of the full guidance graph."""
    noise = noise or Noise()
    keys = cfg.keys
    x, xrep, xbch, xdwt = batch["x"], batch["xrep"], batch["xbch"], batch["xdwt"]
    u_loc, u_std, l = {}, {}, {}
    for k in keys:
        u_loc[k], u_std[k], l[k] = data_encoder(
            state, f"x2u.{k}.", x[k], xrep[k], count_input=cfg.count_encoder[k],
            lazy_normalizer=dsc_only, h_depth=cfg.h_depth, training=training,
            dropout_masks=None if noise.enc_dropout is None else noise.enc_dropout[k])
    usamp = {}
    for k in keys:
        eps = torch.randn_like(u_loc[k]) if noise.u_eps is None else noise.u_eps[k]
        usamp[k] = u_loc[k] + eps * u_std[k]
    if cfg.normalize_u:
        usamp = {k: F.normalize(usamp[k], dim=1) for k in keys}

    u_cat = torch.cat([u_loc[k] for k in keys])
    xbch_cat = torch.cat([xbch[k] for k in keys])
    xdwt_cat = torch.cat([xdwt[k] for k in keys])
    xflag_cat = torch.cat([torch.full((x[k].shape[0],), i, dtype=torch.int64)
                           for i, k in enumerate(keys)])
    anneal = max(1 - (epoch - 1) / cfg.align_burnin, 0) if cfg.align_burnin else 0
    if anneal:
        eps = torch.randn_like(u_cat) if noise.burnin_eps is None else noise.burnin_eps
        # D.Normal(0, std).sample() runs under no_grad: eps * std + 0, detached
        u_cat = u_cat + (anneal * cfg.burnin_noise_exag) * (eps * u_cat.std(dim=0)).detach()
    dsc_logits = discriminator(state, u_cat, xbch_cat, n_batches=cfg.du_n_batches,
                               h_depth=cfg.h_depth, dropout_masks=noise.du_dropout)
    dsc_loss = F.cross_entropy(dsc_logits, xflag_cat, reduction="none")
    dsc_loss = (dsc_loss * xdwt_cat).sum() / xdwt_cat.numel()
    if dsc_only:
        return {"dsc_loss": cfg.lam_align * dsc_loss}

    v_loc, v_std = graph_encoder(state, *graph)
    v_eps = torch.randn_like(v_loc) if noise.v_eps is None else noise.v_eps
    vsamp = v_loc + v_eps * v_std

    g_nll = graph_nll(vsamp, batch["eidx"], batch["esgn"], batch["ewt"])
    g_kl = kl_normal_std(v_loc, v_std).sum(dim=1).mean() / vsamp.shape[0]
    g_elbo = g_nll + cfg.lam_kl * g_kl

    nan_aware = not cfg.paired  # scglue.py:345 (.nanmean) vs scglue.py:603 (.mean)
    x_nll = {k: _decoder_nll(cfg, state, k, usamp[k], vsamp[state[f"{k}_idx"]],
                             xbch[k], l[k], x[k], nan_aware) for k in keys}
    x_kl = {k: kl_normal_std(u_loc[k], u_std[k]).sum(dim=1).mean() / x[k].shape[1] for k in keys}
    x_elbo = {k: x_nll[k] + cfg.lam_kl * x_kl[k] for k in keys}
    x_elbo_sum = sum(cfg.modality_weight[k] * x_elbo[k] for k in keys)
    sup_loss = torch.tensor(0.0)

    vae_loss = (cfg.lam_data * x_elbo_sum + cfg.lam_graph * len(keys) * g_elbo
                + cfg.lam_sup * sup_loss)
    losses = {"dsc_loss": dsc_loss, "g_nll": g_nll, "g_kl": g_kl, "g_elbo": g_elbo}
    if cfg.paired:
        pmsk = batch["pmsk"].T  # [n_modalities, B]
        ustack = torch.stack([usamp[k] for k in keys])
        pstack = pmsk.unsqueeze(2).expand_as(ustack)
        umean = (ustack * pstack).sum(dim=0) / pstack.sum(dim=0)
        if cfg.normalize_u:
            umean = F.normalize(umean, dim=1)
        zero = torch.as_tensor(0.0)
        joint = zero
        if cfg.lam_joint_cross:
            joint = sum(
                cfg.modality_weight[k] * _decoder_nll(
                    cfg, state, k, umean[m], vsamp[state[f"{k}_idx"]], xbch[k][m],
                    None if l[k] is None else l[k][m], x[k][m], False)
                for k, m in zip(keys, pmsk) if m.sum())
        real = zero
        if cfg.lam_real_cross:
            real = 0.0
            for kt, mt in zip(keys, pmsk):
                acc = torch.as_tensor(0.0)
                for ks, ms in zip(keys, pmsk):
                    if ks == kt:
                        continue
                    m = ms & mt
                    if m.sum():
                        acc = acc + _decoder_nll(
                            cfg, state, kt, usamp[ks][m], vsamp[state[f"{kt}_idx"]], xbch[kt][m],
                            None if l[kt] is None else l[kt][m], x[kt][m], False)
                real = real + cfg.modality_weight[kt] * acc
        cos = zero
        if cfg.lam_cos:
            cos = sum(1 - F.cosine_similarity(ustack[i, m], umean[m]).mean()
                      for i, m in enumerate(pmsk) if m.sum())
        vae_loss = (vae_loss + cfg.lam_joint_cross * joint + cfg.lam_real_cross * real
                    + cfg.lam_cos * cos)
        losses.update(joint_cross_loss=joint, real_cross_loss=real, cos_loss=cos)
    losses["vae_loss"] = vae_loss
    losses["gen_loss"] = vae_loss - cfg.lam_align * dsc_loss
    for k in keys:
        losses.update({f"x_{k}_nll": x_nll[k], f"x_{k}_kl": x_kl[k], f"x_{k}_elbo": x_elbo[k]})
    return losses


# parameter groups of the two optimizers (glue.py:329-341)
def vae_param_names(state: State) -> List[str]:
    return [k for k in state if k.split(".")[0] in ("g2v", "v2g", "x2u", "u2x")
            and not _is_buffer(k)]


def dsc_param_names(state: State) -> List[str]:
    return [k for k in state if k.startswith("du.") and not _is_buffer(k)]


def _is_buffer(name: str) -> bool:
    return name.endswith(("running_mean", "running_var", "num_batches_tracked", "_idx"))


class CpuTrainer:
    """The body of ``SCGLUETrainer.train_step`` (scglue.py:378-401): one
    discriminator update then one generator update, both RMSprop (scglue.py:939)."""

    def __init__(self, cfg: StepConfig, state: Dict[str, Tensor],
                 graph: Tuple[Tensor, Tensor, Tensor], lr: float = 2e-3):
        self.cfg, self.state, self.graph = cfg, state, graph
        for name in vae_param_names(state) + dsc_param_names(state):
            state[name].requires_grad_(True)
        self.vae_optim = torch.optim.RMSprop([state[n] for n in vae_param_names(state)], lr=lr)
        self.dsc_optim = torch.optim.RMSprop([state[n] for n in dsc_param_names(state)], lr=lr)

    def _zero(self) -> None:
        for t in self.state.values():
            t.grad = None

    def train_step(self, batch, epoch: int, noise_dsc: Optional[Noise] = None,
                   noise_gen: Optional[Noise] = None) -> Dict[str, Tensor]:
        losses = compute_losses(self.cfg, self.state, batch, self.graph, epoch,
                                dsc_only=True, noise=noise_dsc)
        self._zero()
        losses["dsc_loss"].backward()
        self.dsc_optim.step()
        losses = compute_losses(self.cfg, self.state, batch, self.graph, epoch, noise=noise_gen)
        self._zero()
        losses["gen_loss"].backward()
        self.vae_optim.step()
        return losses
