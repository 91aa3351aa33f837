r"""
SCGLUE component modules (API of ``scglue/models/sc.py``).

Class names, constructor signatures, ``forward`` signatures and parameter names match
the reference so that ``state_dict`` s interchange and ``register_prob_model`` works the
same way.  What differs is what ``forward`` launches: the graph encoder's propagation,
the graph decoder and the NB / ZINB / Normal / ZILN data decoders run on the sm_100a
kernels of ``libglue_b200.so``; encoders, discriminator, classifier and prior are small
dense layers and stay on cuBLAS through PyTorch.
"""

import collections
from typing import Optional, Tuple

import torch
import torch.distributions as D
import torch.nn.functional as F

from .. import ops
from ..num import EPS
from ..utils import config
from . import glue
from .nn import GraphConv
from .prob import ZIN, Bernoulli, EdgeBernoulli, FusedDataDistribution

# ------------------------------------------------------------------ graph modules


class GraphEncoder(glue.GraphEncoder):
    r"""
    Graph encoder (``scglue/models/sc.py:21-46``): one propagation step over the
    normalised signed guidance graph, then a Gaussian posterior per vertex.

    Parameters ``vrepr`` (zero-initialised), ``loc.*``, ``std_lin.*`` as in the reference.
    """

    def __init__(self, vnum: int, out_features: int) -> None:
        super().__init__()
        self.vrepr = torch.nn.Parameter(torch.zeros(vnum, out_features))
        self.conv = GraphConv()
        self.loc = torch.nn.Linear(out_features, out_features)
        self.std_lin = torch.nn.Linear(out_features, out_features)

    def forward(self, eidx: torch.Tensor, enorm: torch.Tensor, esgn: torch.Tensor) -> D.Normal:
        ptr = self.conv(self.vrepr, eidx, enorm, esgn)
        return D.Normal(self.loc(ptr), F.softplus(self.std_lin(ptr)) + EPS, validate_args=False)

    def sample_with_kl(self, eidx: torch.Tensor, enorm: torch.Tensor, esgn: torch.Tensor,
                       noise: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        r"""Training-path entry point: ``(v.mean + noise * v.stddev, kl(v || N(0, 1)).sum())`` for
        ``v = self(eidx, enorm, esgn)``, computed by the fused head kernel right after the
        propagation (same values as the two-step formulation of ``scglue.py:328-340``)."""
        ptr = self.conv(self.vrepr, eidx, enorm, esgn)
        vsamp, kl_sum, _, _ = ops.graph_encoder_head(ptr, self.loc.weight, self.loc.bias,
                                                     self.std_lin.weight, self.std_lin.bias, noise)
        return vsamp, kl_sum


class GraphDecoder(glue.GraphDecoder):
    r"""
    Graph decoder (``scglue/models/sc.py:49-59``).  Returns a lazy
    :class:`~glue_b200.models.prob.EdgeBernoulli`; ``.log_prob(ewt)`` gives per-edge values
    as in the reference, ``.balanced_nll(ewt)`` the fused training loss.
    """

    def forward(self, v: torch.Tensor, eidx: torch.Tensor, esgn: torch.Tensor) -> EdgeBernoulli:
        return EdgeBernoulli(v, eidx, esgn)


# ------------------------------------------------------------------ data encoders


class DataEncoder(glue.DataEncoder):
    r"""
    MLP data encoder (``scglue/models/sc.py:62-178``): ``h_depth`` blocks of
    Linear -> LeakyReLU(0.2) -> BatchNorm1d -> Dropout, then ``loc`` / ``std_lin`` heads.
    """

    def __init__(self, in_features: int, out_features: int, h_depth: int = 2, h_dim: int = 256,
                 dropout: float = 0.2) -> None:
        super().__init__()
        self.h_depth = h_depth
        width = in_features
        for layer in range(h_depth):
            setattr(self, f"linear_{layer}", torch.nn.Linear(width, h_dim))
            setattr(self, f"act_{layer}", torch.nn.LeakyReLU(negative_slope=0.2))
            setattr(self, f"bn_{layer}", torch.nn.BatchNorm1d(h_dim))
            setattr(self, f"dropout_{layer}", torch.nn.Dropout(p=dropout))
            width = h_dim
        self.loc = torch.nn.Linear(width, out_features)
        self.std_lin = torch.nn.Linear(width, out_features)

    def _dropout_state(self, layer: int, device: torch.device) -> torch.Tensor:
        r"""Philox state of hidden block ``layer`` for the fused kernel (seed drawn from torch's
        generator at first use, call counter advanced on the device); not part of ``state_dict``."""
        name = f"_dropout_state_{layer}"
        state = getattr(self, name, None)
        if state is None or state.device != device:
            state = ops.new_dropout_state(device)
            self.register_buffer(name, state, persistent=False)
        return state

    def _fused_forward(self, ptr: torch.Tensor):
        r"""The whole MLP in ``h_depth + 1`` launches each way (``ops.encoder_mlp``) when the shapes
        allow it (a training-mode minibatch of <= 128 cells, widths <= 256); ``None`` otherwise."""
        if not (config.USE_FUSED_MLP and config.USE_FUSED_ENCODER and self.h_depth >= 1 and ptr.is_cuda):
            return None
        lins = [getattr(self, f"linear_{i}") for i in range(self.h_depth)]
        bns = [getattr(self, f"bn_{i}") for i in range(self.h_depth)]
        acts = [getattr(self, f"act_{i}") for i in range(self.h_depth)]
        drops = [getattr(self, f"dropout_{i}") for i in range(self.h_depth)]
        if any(not bn.training or bn.momentum is None or not bn.affine for bn in bns):
            return None
        if any(type(a) is not torch.nn.LeakyReLU for a in acts) or len({a.negative_slope for a in acts}) != 1:
            return None
        ps = {d.p if d.training else 0.0 for d in drops}
        if len(ps) != 1 or ptr.requires_grad:
            return None
        if not ops.encoder_mlp_supported(ptr, [lin.weight for lin in lins], self.loc.out_features):
            return None
        p = ps.pop()
        rngs = [self._dropout_state(i, ptr.device) if p > 0 else None for i in range(self.h_depth)]
        return ops.encoder_mlp(ptr, lins, bns, (self.loc, self.std_lin), p, rngs,
                               negative_slope=acts[0].negative_slope)

    def compute_l(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        raise NotImplementedError  # pragma: no cover

    def normalize(self, x: torch.Tensor, l: Optional[torch.Tensor]) -> torch.Tensor:
        raise NotImplementedError  # pragma: no cover

    def forward(self, x: torch.Tensor, xrep: torch.Tensor, lazy_normalizer: bool = True,
                l: Optional[torch.Tensor] = None) -> Tuple[D.Normal, Optional[torch.Tensor]]:
        r"""
        ``xrep`` (if non-empty) replaces the normalised ``x`` as network input; the
        normaliser ``l`` is still computed from ``x`` unless ``lazy_normalizer``.
        ``l`` may be passed in when the caller keeps per-cell library sizes precomputed
        (device-resident pipeline) -- it then is not recomputed from ``x``.
        """
        if xrep.numel():
            if l is None and not lazy_normalizer:
                l = self.compute_l(x)
            ptr = xrep
        else:
            if l is None:
                l = self.compute_l(x)
            ptr = self.normalize(x, l)
        fused = self._fused_forward(ptr)
        if fused is not None:
            return D.Normal(fused[0], fused[1], validate_args=False), l
        for layer in range(self.h_depth):
            ptr = getattr(self, f"linear_{layer}")(ptr)
            act, bn = getattr(self, f"act_{layer}"), getattr(self, f"bn_{layer}")
            dropout = getattr(self, f"dropout_{layer}")
            if (config.USE_FUSED_MLP and ptr.is_cuda and ptr.dtype == torch.float32 and bn.training
                    and bn.momentum is not None and type(act) is torch.nn.LeakyReLU):
                # activation + batch normalisation + dropout in one launch each way
                p = dropout.p if dropout.training else 0.0
                ptr = ops.act_bn_dropout(ptr, bn, p, self._dropout_state(layer, ptr.device) if p > 0 else None,
                                         negative_slope=act.negative_slope)
            else:
                ptr = dropout(bn(act(ptr)))
        loc = self.loc(ptr)
        std = F.softplus(self.std_lin(ptr)) + EPS
        return D.Normal(loc, std, validate_args=False), l


class VanillaDataEncoder(DataEncoder):
    r"""Encoder without input normalisation (``sc.py:181-203``)."""

    def compute_l(self, x):
        return None

    def normalize(self, x, l):
        return x


class NBDataEncoder(DataEncoder):
    r"""Encoder for count data: library size + ``log1p(x * 1e4 / l)`` (``sc.py:206-230``)."""

    TOTAL_COUNT = 1e4

    def compute_l(self, x):
        return x.sum(dim=1, keepdim=True)

    def normalize(self, x, l):
        return (x * (self.TOTAL_COUNT / l)).log1p()


# ------------------------------------------------------------------ data decoders


class DataDecoder(glue.DataDecoder):
    r"""Abstract data decoder ``forward(u, v, b, l) -> distribution`` (``sc.py:233-277``)."""

    def __init__(self, out_features: int, n_batches: int = 1) -> None:
        super().__init__()


class _FusedDecoder(DataDecoder):
    r"""
    Shared implementation of the four kernel-backed decoders.  Per-feature parameter
    tables have shape ``[n_batches, out_features]`` and are zero-initialised exactly like
    the reference.
    """

    FAMILY: str = ""
    DISPERSION: str = ""  # name of the third parameter table
    ZERO_INFLATED: bool = False

    def __init__(self, out_features: int, n_batches: int = 1) -> None:
        super().__init__(out_features, n_batches=n_batches)
        self.scale_lin = torch.nn.Parameter(torch.zeros(n_batches, out_features))
        self.bias = torch.nn.Parameter(torch.zeros(n_batches, out_features))
        setattr(self, self.DISPERSION, torch.nn.Parameter(torch.zeros(n_batches, out_features)))
        if self.ZERO_INFLATED:
            self.zi_logits = torch.nn.Parameter(torch.zeros(n_batches, out_features))

    def forward(self, u: torch.Tensor, v: torch.Tensor, b: torch.Tensor,
                l: Optional[torch.Tensor], vidx: Optional[torch.Tensor] = None
                ) -> FusedDataDistribution:
        r"""
        ``u`` sample latent ``[B, K]``; ``v`` feature latent ``[F, K]``; ``b`` batch index
        ``[B]``; ``l`` normaliser ``[B, 1]`` (count models).  With ``vidx`` given, ``v`` is
        the whole vertex table and ``v[vidx]`` is gathered inside the kernel.
        """
        return FusedDataDistribution(
            self.FAMILY, u, v, b, l, self.scale_lin, self.bias, getattr(self, self.DISPERSION),
            self.zi_logits if self.ZERO_INFLATED else None, vidx=vidx)


class NormalDataDecoder(_FusedDecoder):
    r"""Normal decoder (``sc.py:280-308``): ``loc = softplus(scale_lin[b]) u v^T + bias[b]``,
    ``std = softplus(std_lin[b]) + EPS``."""
    FAMILY, DISPERSION = "Normal", "std_lin"


class ZILNDataDecoder(_FusedDecoder):
    r"""Zero-inflated log-normal decoder (``sc.py:340-369`` + ``prob.py:110-118``)."""
    FAMILY, DISPERSION, ZERO_INFLATED = "ZILN", "std_lin", True

    def __init__(self, out_features: int, n_batches: int = 1) -> None:
        # parameter registration order of the reference: scale_lin, bias, zi_logits, std_lin
        DataDecoder.__init__(self, out_features, n_batches=n_batches)
        for name in ("scale_lin", "bias", "zi_logits", "std_lin"):
            setattr(self, name, torch.nn.Parameter(torch.zeros(n_batches, out_features)))


class NBDataDecoder(_FusedDecoder):
    r"""Negative binomial decoder (``sc.py:372-397``): ``mu = softmax(logit) * l``,
    ``NB(exp(log_theta[b]), logits=log(mu + EPS) - log_theta[b])``."""
    FAMILY, DISPERSION = "NB", "log_theta"


class ZINBDataDecoder(_FusedDecoder):
    r"""Zero-inflated negative binomial decoder (``sc.py:447-478`` + ``prob.py:144-153``)."""
    FAMILY, DISPERSION, ZERO_INFLATED = "ZINB", "log_theta", True


class ZINDataDecoder(DataDecoder):
    r"""Zero-inflated normal decoder (``sc.py:311-337``); no config on the hot path uses it,
    so it stays a plain PyTorch module."""

    def __init__(self, out_features: int, n_batches: int = 1) -> None:
        super().__init__(out_features, n_batches=n_batches)
        for name in ("scale_lin", "bias", "std_lin", "zi_logits"):
            setattr(self, name, torch.nn.Parameter(torch.zeros(n_batches, out_features)))

    def forward(self, u, v, b, l) -> ZIN:
        loc = F.softplus(self.scale_lin[b]) * (u @ v.t()) + self.bias[b]
        std = F.softplus(self.std_lin[b]) + EPS
        return ZIN(self.zi_logits[b].expand_as(loc), loc, std)


class BernoulliDataDecoder(DataDecoder):
    r"""Bernoulli decoder (``sc.py:547-573``), plain PyTorch."""

    def __init__(self, out_features: int, n_batches: int = 1) -> None:
        super().__init__(out_features, n_batches=n_batches)
        self.scale_lin = torch.nn.Parameter(torch.zeros(n_batches, out_features))
        self.bias = torch.nn.Parameter(torch.zeros(n_batches, out_features))

    def forward(self, u, v, b, l) -> Bernoulli:
        return Bernoulli(F.softplus(self.scale_lin[b]) * (u @ v.t()) + self.bias[b])


# ----------------------------------------------------- discriminator / classifier / prior


class Discriminator(torch.nn.Sequential, glue.Discriminator):
    r"""
    Modality discriminator MLP (``scglue/models/sc.py:576-624``); optional one-hot batch
    covariate and input standardisation.
    """

    def __init__(self, in_features: int, out_features: int, n_batches: int = 0, h_depth: int = 2,
                 h_dim: Optional[int] = 256, dropout: float = 0.2, norm: bool = False) -> None:
        self.n_batches = n_batches
        self.norm = norm
        layers = collections.OrderedDict()
        width = in_features + n_batches
        for layer in range(h_depth):
            layers[f"linear_{layer}"] = torch.nn.Linear(width, h_dim)
            layers[f"act_{layer}"] = torch.nn.LeakyReLU(negative_slope=0.2)
            layers[f"dropout_{layer}"] = torch.nn.Dropout(p=dropout)
            width = h_dim
        layers["pred"] = torch.nn.Linear(width, out_features)
        super().__init__(layers)

    def _dropout_state(self, layer: int, device: torch.device) -> torch.Tensor:
        r"""Philox state of block ``layer``'s dropout for the fused kernels (as ``DataEncoder``)."""
        name = f"_dropout_state_{layer}"
        state = getattr(self, name, None)
        if state is None or state.device != device:
            state = ops.new_dropout_state(device)
            self.register_buffer(name, state, persistent=False)
        return state

    def _fused_forward(self, x: torch.Tensor):
        r"""All blocks + ``pred`` in ``h_depth + 1`` launches each way (``ops.discriminator_mlp``) when the
        shapes allow it; ``None`` otherwise."""
        if not (config.USE_FUSED_MLP and config.USE_FUSED_ENCODER and x.is_cuda and x.dtype == torch.float32):
            return None
        names = [n for n, _ in self.named_children()]
        depth = sum(1 for n in names if n.startswith("linear_"))
        if depth < 1 or names != [f"{kind}_{i}" for i in range(depth) for kind in ("linear", "act", "dropout")] + ["pred"]:
            return None
        lins = [getattr(self, f"linear_{i}") for i in range(depth)]
        acts = [getattr(self, f"act_{i}") for i in range(depth)]
        drops = [getattr(self, f"dropout_{i}") for i in range(depth)]
        if any(type(a) is not torch.nn.LeakyReLU for a in acts) or len({a.negative_slope for a in acts}) != 1:
            return None
        ps = {d.p if d.training else 0.0 for d in drops}
        if len(ps) != 1:
            return None
        x = x.contiguous()
        if not ops.discriminator_mlp_supported(x, [lin.weight for lin in lins], self.pred.out_features):
            return None
        p = ps.pop()
        rngs = [self._dropout_state(i, x.device) if p > 0 else None for i in range(depth)]
        return ops.discriminator_mlp(x, lins, self.pred, p, rngs, negative_slope=acts[0].negative_slope)

    def forward(self, x: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        if self.norm:
            x = (x - x.mean(dim=0)) / (x.std(dim=0) + EPS)
        if self.n_batches:
            x = torch.cat([x, F.one_hot(b, num_classes=self.n_batches)], dim=1)
        fused = self._fused_forward(x)
        return fused if fused is not None else super().forward(x)


class Classifier(torch.nn.Linear):
    r"""Linear label classifier (``sc.py:627-637``)."""


class Prior(glue.Prior):
    r"""Standard-normal latent prior held as buffers (``sc.py:640-660``)."""

    def __init__(self, loc: float = 0.0, std: float = 1.0) -> None:
        super().__init__()
        self.register_buffer("loc", torch.as_tensor(loc, dtype=torch.get_default_dtype()))
        self.register_buffer("std", torch.as_tensor(std, dtype=torch.get_default_dtype()))

    def forward(self) -> D.Normal:
        return D.Normal(self.loc, self.std, validate_args=False)
