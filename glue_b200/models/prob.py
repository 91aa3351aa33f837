r"""
Probability distributions (API of ``scglue/models/prob.py``) plus the lazy, kernel-backed
distributions returned by the fused decoders.

``NB``/``ZINB``/``ZILN``/``ZIN`` are plain ``torch.distributions`` objects for users who
build distributions by hand; they are written with ``torch.where`` instead of boolean-mask
indexing, so they have static shapes and no host synchronisation.
"""

import torch
import torch.distributions as D
import torch.nn.functional as F

from .. import ops
from ..num import EPS


def _zero_nan_grad(grad: torch.Tensor) -> torch.Tensor:
    return grad.nan_to_num()


class NB(D.NegativeBinomial):
    r"""Negative binomial -- the reference uses ``torch.distributions.NegativeBinomial``
    directly (``scglue/models/sc.py:397``); this alias completes the ``prob`` namespace."""


class ZINB(D.NegativeBinomial):
    r"""
    Zero-inflated negative binomial (``scglue/models/prob.py:121-153``).

    ``zi_logits``: zero-inflation logits; ``total_count`` / ``logits``: NB parameters.
    """

    def __init__(self, zi_logits, total_count, logits=None, validate_args=None):
        super().__init__(total_count, logits=logits, validate_args=validate_args)
        self.zi_logits = zi_logits

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        raw = super().log_prob(value)
        zi = self.zi_logits
        zero_case = torch.log(raw.exp() + zi.exp() + EPS)
        return torch.where(value.abs() < EPS, zero_case, raw) - F.softplus(zi)


class ZILN(D.LogNormal):
    r"""Zero-inflated log-normal (``scglue/models/prob.py:90-118``)."""

    def __init__(self, zi_logits, loc, scale, validate_args=None):
        super().__init__(loc, scale, validate_args=validate_args)
        self.zi_logits = zi_logits

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        is_zero = value.abs() < EPS
        safe = torch.where(is_zero, torch.ones_like(value), value)
        ln = D.LogNormal(self.loc, self.scale, validate_args=False).log_prob(safe)
        return torch.where(is_zero, self.zi_logits, ln) - F.softplus(self.zi_logits)


class ZIN(D.Normal):
    r"""Zero-inflated normal (``scglue/models/prob.py:52-87``)."""

    def __init__(self, zi_logits, loc, scale):
        # as the reference (``prob.py:70-75``): NaN gradients (NaN observations, or the unselected
        # branch of the ``where`` below overflowing) are zeroed on their way into the parameters
        for t in (zi_logits, loc, scale):
            if t.requires_grad:
                t.register_hook(_zero_nan_grad)
        super().__init__(loc, scale, validate_args=False)
        self.zi_logits = zi_logits

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        raw = super().log_prob(value)
        zi = self.zi_logits
        zero_case = torch.log(raw.exp() + zi.exp() + EPS)
        return torch.where(value.abs() < EPS, zero_case, raw) - F.softplus(zi)


class Bernoulli(D.Bernoulli):
    r"""Bernoulli without argument validation (``scglue/models/prob.py:202-207``)."""

    def __init__(self, logits):
        super().__init__(logits=logits, validate_args=False)


# ----------------------------------------------------------- kernel-backed objects


class FusedDataDistribution:
    r"""
    Lazy reconstruction distribution returned by the fused data decoders.

    It behaves like the ``torch.distributions`` object the reference decoder returns as
    far as the model code uses it -- ``log_prob(x) -> [B, F]`` and ``.mean`` -- and adds
    :meth:`nll`, the reduced entry point the trainers call: ``-log_prob(x).(nan)mean()``
    with all gradients in one fused launch sequence, never materialising ``[B, F]``.
    """

    def __init__(self, family, u, v, b, l, scale_lin, bias, disp, zi_logits=None, vidx=None):
        self.family = family
        self.u, self.v, self.b, self.l, self.vidx = u, v, b, l, vidx
        self.scale_lin, self.bias, self.disp, self.zi_logits = scale_lin, bias, disp, zi_logits

    def _params(self):
        return self.scale_lin, self.bias, self.disp, self.zi_logits

    def nll(self, value: torch.Tensor, reduce: str = "nanmean", unit_upstream: bool = False,
            row_weight: torch.Tensor = None, upstream: float = None) -> torch.Tensor:
        return ops.decoder_nll(self.family, self.u, self.v, *self._params(), b=self.b, l=self.l,
                               x=value, vidx=self.vidx, reduce=reduce,
                               unit_upstream=unit_upstream, row_weight=row_weight,
                               upstream=upstream)

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        return ops.decoder_log_prob(self.family, self.u, self.v, *self._params(), b=self.b,
                                    l=self.l, x=value, vidx=self.vidx)

    @property
    def mean(self) -> torch.Tensor:
        return ops.decoder_mean(self.family, self.u, self.v, *self._params(), b=self.b, l=self.l,
                                vidx=self.vidx)


class EdgeBernoulli:
    r"""
    Lazy edge distribution returned by the fused :class:`GraphDecoder`:
    ``logits = esgn * <v[sidx], v[tidx]>`` (``scglue/models/sc.py:57-59``).
    """

    def __init__(self, v, eidx, esgn):
        self.v, self.eidx, self.esgn = v, eidx, esgn

    @property
    def logits(self) -> torch.Tensor:
        return ops.graph_logits(self.v, self.eidx, self.esgn)

    @property
    def probs(self) -> torch.Tensor:
        return torch.sigmoid(self.logits)

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        return -F.binary_cross_entropy_with_logits(self.logits, value, reduction="none")

    def balanced_nll(self, value: torch.Tensor, upstream: float = None) -> torch.Tensor:
        r"""Mean NLL of positive and of negative edges, averaged
        (``scglue/models/scglue.py:331-338``), fused with the gather-dot and free of
        the reference's ``.item()`` host synchronisation.  ``upstream``: see ``ops.graph_nll``."""
        return ops.graph_nll(self.v, self.eidx, self.esgn, value, upstream=upstream)
