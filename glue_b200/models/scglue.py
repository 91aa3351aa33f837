r"""
SCGLUE network, trainers and public model classes (API of ``scglue/models/scglue.py``).

``SCGLUETrainer.compute_losses`` evaluates the same objective as the reference
(``scglue.py:292-376``; paired variant ``scglue.py:550-701``) but routes the three hot
sub-computations through the sm_100a kernels:

* graph encoder propagation  -> CSR GraphConv (forward + transposed-CSR backward);
* graph reconstruction loss  -> gather-dot + fused BCE + pos/neg balancing, no host sync;
* data reconstruction losses -> fused decoder NLL, gathering ``v[idx_k]`` inside the kernel
  and producing every gradient in the forward launch sequence.
"""

import copy
from itertools import chain
from math import ceil
from typing import Callable, List, Mapping, Optional, Tuple, Union

import networkx as nx
import numpy as np
import pandas as pd
import torch
import torch.distributions as D
import torch.nn.functional as F

from .. import dist as gdist
from ..num import normalize_edges
from ..utils import AUTO, config, logged
from . import sc
from .base import LossVector, Model
from .data import AnnDataset, DataLoader, GraphDataset
from .glue import GLUE, GLUETrainer
from .graphs import StepGraphs
from .nn import freeze_running_stats

# ---------------------------------------------------------- prob-model registry

_ENCODER_MAP: Mapping[str, type] = {}
_DECODER_MAP: Mapping[str, type] = {}


def register_prob_model(prob_model: str, encoder: type, decoder: type) -> None:
    r"""
    Register a data encoder / decoder pair under a ``prob_model`` name
    (extension point of the reference, ``scglue.py:36-61`` / ``docs/dev.rst``).
    """
    _ENCODER_MAP[prob_model] = encoder
    _DECODER_MAP[prob_model] = decoder


register_prob_model("Normal", sc.VanillaDataEncoder, sc.NormalDataDecoder)
register_prob_model("ZIN", sc.VanillaDataEncoder, sc.ZINDataDecoder)
register_prob_model("ZILN", sc.VanillaDataEncoder, sc.ZILNDataDecoder)
register_prob_model("NB", sc.NBDataEncoder, sc.NBDataDecoder)
register_prob_model("ZINB", sc.NBDataEncoder, sc.ZINBDataDecoder)
register_prob_model("Bernoulli", sc.VanillaDataEncoder, sc.BernoulliDataDecoder)


# ---------------------------------------------------------------------- network


class SCGLUE(GLUE):
    r"""GLUE network for single-cell multi-omics integration, with an optional cell-type
    classifier ``u2c`` (``scglue.py:67-103``)."""

    def __init__(self, g2v, v2g, x2u, u2x, idx, du, prior, u2c=None) -> None:
        super().__init__(g2v, v2g, x2u, u2x, idx, du, prior)
        self.u2c = u2c.to(self.device) if u2c else None


def _kl_to_prior(q: D.Normal, prior: D.Normal) -> torch.Tensor:
    return D.kl_divergence(q, prior).sum(dim=1).mean()


# --------------------------------------------------------------------- trainers


class _Posterior:
    r"""What :meth:`SCGLUETrainer._alignment` reads of an encoder's posterior: its mean."""

    def __init__(self, mean: torch.Tensor) -> None:
        self.mean = mean


class _InjectGrad(torch.autograd.Function):
    r"""``forward(streams, leaves, *xs) = 0`` (a scalar that enters the objective with coefficient 1);
    ``backward`` hands ``leaves[i].grad`` to ``xs[i]``.  ``leaves[i]`` is the detached copy of ``xs[i]`` that a
    part of the objective was computed from and back-propagated ahead of the rest (the alignment term, see
    ``compute_losses``): its gradient joins the others that reach ``xs[i]`` when the rest of the objective is
    back-propagated."""

    @staticmethod
    def forward(ctx, streams, leaves, *xs):
        ctx.leaves, ctx.streams = leaves, streams
        return xs[0].new_zeros(())

    @staticmethod
    def backward(ctx, g):
        grads = [leaf.grad for leaf in ctx.leaves]
        ctx.leaves = None
        if any(grad is None for grad in grads):
            raise RuntimeError("early alignment backward: the gradient was not produced")
        for grad in grads:  # (allocated on this stream, consumed on the encoders')
            for stream in ctx.streams:
                grad.record_stream(stream)
        return (None, None, *grads)


@logged
class SCGLUETrainer(GLUETrainer):
    r"""
    Trainer for :class:`SCGLUE` (``scglue.py:159-420``).  One ``train_step`` = one
    discriminator update followed by one generator (VAE) update.
    """

    BURNIN_NOISE_EXAG: float = 1.5  # burn-in noise exaggeration

    def __init__(self, net: SCGLUE, lam_data: float = None, lam_kl: float = None,
                 lam_graph: float = None, lam_align: float = None, lam_sup: float = None,
                 dsc_steps: int = None, normalize_u: bool = None,
                 modality_weight: Mapping[str, float] = None, optim: str = None,
                 lr: float = None, **kwargs) -> None:
        super().__init__(net, lam_data=lam_data, lam_kl=lam_kl, lam_graph=lam_graph,
                         lam_align=lam_align, dsc_steps=dsc_steps,
                         modality_weight=modality_weight, optim=optim, lr=lr, **kwargs)
        for name, val in dict(lam_sup=lam_sup, normalize_u=normalize_u).items():
            if val is None:
                raise ValueError(f"`{name}` must be specified!")
        self.lam_sup, self.normalize_u = lam_sup, normalize_u
        self.freeze_u = False
        if net.u2c:
            self.required_losses.append("sup_loss")
            self.vae_optim = self._make_optim(self._vae_modules() + [net.u2c])
        # `noise(shape_like, tag)` supplies standard-normal draws; tests inject host noise
        self.noise: Callable[[torch.Tensor, str], torch.Tensor] = \
            lambda like, tag: torch.randn_like(like)
        # whole-step CUDA graphs (see models/graphs.py)
        self.graphs = StepGraphs(self)
        self._side_streams: List[torch.cuda.Stream] = []
        self._dsc_stream: Optional[torch.cuda.Stream] = None
        self._nll_stream: Optional[torch.cuda.Stream] = None
        self._nll_pending = False
        self._early_stream: Optional[torch.cuda.Stream] = None
        self._early_pending = False
        self._early_align_done = False
        self._early_reduced = None
        self._in_generator_step = False
        self.capturing = False
        self.anneal_buffer = torch.zeros((), dtype=torch.float32, device=net.device)

    @property
    def freeze_u(self) -> bool:
        r"""Whether cell embeddings (encoders + discriminator) are frozen."""
        return self._freeze_u

    @freeze_u.setter
    def freeze_u(self, freeze_u: bool) -> None:
        self._freeze_u = freeze_u
        for p in chain(self.net.x2u.parameters(), self.net.du.parameters()):
            p.requires_grad_(not freeze_u)

    # -- data formatting -------------------------------------------------------------------
    PAIRED = False

    def format_data(self, data: List[torch.Tensor]):
        r"""
        Split the flat loader output ``[x.., xrep.., xbch.., xlbl.., xdwt.., (pmsk,) eidx, ewt,
        esgn]`` (``scglue.py:255-290``) into per-modality dicts on the compute device.
        Host tensors are copied (pinned + ``non_blocking``); device-resident batches pass
        through untouched.
        """
        device, keys = self.net.device, self.net.keys
        K = len(keys)
        data = self.unpack_data(data)
        move = lambda t: t.to(device, non_blocking=True)
        groups = [dict(zip(keys, map(move, data[i * K:(i + 1) * K]))) for i in range(5)]
        x, xrep, xbch, xlbl, xdwt = groups
        rest = data[5 * K:]
        pmsk = move(rest[0])
        eidx, ewt, esgn = (move(t) for t in rest[1:])
        xflag = {k: self._modality_flag(x[k].shape[0], i, device) for i, k in enumerate(keys)}
        if self.PAIRED:
            return x, xrep, xbch, xlbl, xdwt, xflag, pmsk, eidx, ewt, esgn
        return x, xrep, xbch, xlbl, xdwt, xflag, eidx, ewt, esgn

    def _modality_flag(self, n: int, index: int, device) -> torch.Tensor:
        r"""``[n]`` int64 tensor filled with the modality index (the discriminator's target); constant, so built
        once per (n, index) -- outside any capture: the first steps of a signature run eagerly."""
        cache = self.__dict__.setdefault("_xflag_cache", {})
        key = (n, index, str(device))
        flag = cache.get(key)
        if flag is None:
            if self.capturing:
                return torch.full((n,), index, dtype=torch.int64, device=device)
            if len(cache) > 32:
                cache.clear()
            flag = cache[key] = torch.full((n,), index, dtype=torch.int64, device=device)
        return flag

    def prepare_capture(self) -> None:
        r"""Build, outside the capture, what the step looks up in process-wide caches: the CSR of the
        full guidance graph (``ops.csr_for``)."""
        g2v = self.net.g2v
        if self.eidx is not None and hasattr(g2v, "vrepr") and g2v.vrepr.is_cuda:
            from .. import ops
            ops.csr_for(self.eidx, self.enorm, self.esgn, g2v.vrepr.shape[0])

    def unpack_data(self, data: List[torch.Tensor]) -> List[torch.Tensor]:
        r"""Count matrices that arrive as packed sparse rows (``data.SparseRows``: host-resident
        datasets) are moved to the device and densified there; everything else passes through."""
        from .data import is_packed_sparse, unpack_sparse_rows
        K = len(self.net.keys)
        if not any(is_packed_sparse(t) for t in data[:K]):
            return data
        device = self.net.device
        data = list(data)
        for i, k in enumerate(self.net.keys):
            if is_packed_sparse(data[i]):
                n_rows = data[2 * K + i].shape[0]  # (one batch index per cell)
                n_cols = getattr(self.net, f"{k}_idx").shape[0]
                data[i] = unpack_sparse_rows(data[i].to(device, non_blocking=True), n_rows, n_cols)
        return data

    # -- shared pieces of the objective ----------------------------------------------------
    def _parallel(self, thunks, first_join: Optional[Tuple[int, Callable]] = None):
        r"""Run independent branches of the step on forked CUDA streams (branch 0 stays on the
        current stream) and join them.  A step is hundreds of tiny kernels; inside the captured
        graph the forks become parallel branches, so the per-kernel launch latencies of e.g.
        the two data encoders and the graph encoder overlap instead of adding up.  Autograd
        replays every backward node on its forward stream, so the backward pass forks the
        same way.  Pure scheduling: the arithmetic of each branch is unchanged.

        ``first_join = (n, fn)``: the current stream joins the first ``n`` branches, runs
        ``fn(results[:n])`` and only then joins the rest -- work that needs the short branches only does
        not wait for the long one."""
        device = self.net.device
        n_first, fn = first_join if first_join is not None else (len(thunks), None)
        if device.type != "cuda" or not config.MULTI_STREAM or len(thunks) < 2:
            results = [f() for f in thunks]
            if fn is not None:
                fn(results[:n_first])
            return results
        main = torch.cuda.current_stream(device)
        while len(self._side_streams) < len(thunks) - 1:
            self._side_streams.append(torch.cuda.Stream(device))
        sides = self._side_streams[:len(thunks) - 1]
        for side in sides:
            side.wait_stream(main)
        results = [thunks[0]()]
        for side, f in zip(sides, thunks[1:]):
            with torch.cuda.stream(side):
                results.append(f())
        for side in sides[:max(n_first - 1, 0)]:
            main.wait_stream(side)
        if fn is not None:
            fn(results[:n_first])
        for side in sides[max(n_first - 1, 0):]:
            main.wait_stream(side)
        return results

    def _encode_one(self, k: str, x, xrep, dsc_only: bool):
        u, l = self.net.x2u[k](x[k], xrep[k], lazy_normalizer=dsc_only)
        usamp = u.mean + self.noise(u.mean, f"u_{k}") * u.stddev
        if self.normalize_u:
            usamp = F.normalize(usamp, dim=1)
        return u, l, usamp

    def _encode(self, x, xrep, dsc_only: bool, extra=None, before_join=None, after_encoders=None):
        r"""Per-modality encoders (and optionally one more independent branch, ``extra``) side
        by side; returns ``(u, l, usamp[, extra_result])``.  ``after_encoders(u, l, usamp)`` is issued on the
        current stream once the encoders have joined, before the ``extra`` branches are waited for."""
        keys = self.net.keys
        thunks = [lambda k=k: self._encode_one(k, x, xrep, dsc_only) for k in keys]
        extras = [] if extra is None else (list(extra) if isinstance(extra, (list, tuple)) else [extra])
        # (`before_join`: work for the forking stream while the branches run: it is issued behind branch 0,
        # which shares that stream, and ahead of the join)
        self._encode_aux = None
        if before_join is not None:
            first = thunks[0]

            def first_then_aux():
                res = first()
                self._encode_aux = before_join()
                return res
            thunks = [first_then_aux] + thunks[1:]
        first_join = None
        if after_encoders is not None:
            first_join = (len(keys), lambda res: after_encoders(*({k: o[i] for k, o in zip(keys, res)}
                                                                  for i in range(3))))
        out = self._parallel(thunks + extras, first_join)
        u = {k: o[0] for k, o in zip(keys, out)}
        l = {k: o[1] for k, o in zip(keys, out)}
        usamp = {k: o[2] for k, o in zip(keys, out)}
        if extra is not None:
            return u, l, usamp, out[len(keys)]  # (further branches are run for their side effects)
        return u, l, usamp

    def burnin_anneal(self, epoch: int) -> float:
        r"""Burn-in noise coefficient of the alignment term (``scglue.py:318-324``)."""
        return max(1 - (epoch - 1) / self.align_burnin, 0) if self.align_burnin else 0

    def _alignment_inputs(self, xbch, xdwt, xflag):
        r"""Per-cell covariates of the discriminator, concatenated over the modalities (the same
        for the discriminator pass and the generator pass of a step: built once)."""
        keys = self.net.keys
        return (torch.cat([xbch[k] for k in keys]), torch.cat([xdwt[k] for k in keys]),
                torch.cat([xflag[k] for k in keys]))

    def _alignment(self, u, xbch, xdwt, xflag, xlbl, epoch: int, cat_inputs=None, upstream=None):
        r"""Discriminator loss on the (burn-in jittered) posterior means (``scglue.py:318-329``).  ``upstream``:
        the coefficient with which the caller back-propagates the result (``ops.weighted_cross_entropy`` folds it
        into the gradient it computes with the loss; ``None``: plain autograd semantics)."""
        net = self.net
        u_cat = torch.cat([u[k].mean for k in net.keys])
        xbch_cat, xdwt_cat, xflag_cat = cat_inputs or self._alignment_inputs(xbch, xdwt, xflag)
        anneal = self.burnin_anneal(epoch)
        if anneal:
            with torch.no_grad():  # D.Normal(0, std).sample(): not differentiated through
                jitter = self.noise(u_cat, "burnin") * u_cat.std(dim=0)
            # inside a captured step the coefficient is read from a device scalar
            # (fp32 product in both cases, so replayed and eager steps agree bit for bit)
            if self.capturing or self.graphs.usable():
                if not self.capturing:  # eager steps launch what a capture launches (see StepGraphs.step)
                    self.anneal_buffer.fill_(anneal)
                coeff = self.anneal_buffer * self.BURNIN_NOISE_EXAG
            else:
                coeff = float(np.float32(anneal) * np.float32(self.BURNIN_NOISE_EXAG))
            u_cat = u_cat + coeff * jitter
        logits = net.du(u_cat, xbch_cat)
        if config.USE_FUSED_OBJECTIVE and logits.is_cuda:
            from .. import ops
            if ops.weighted_cross_entropy_supported(logits) and xdwt_cat.dtype == torch.float32:
                return u_cat, ops.weighted_cross_entropy(logits, xflag_cat, xdwt_cat, upstream=upstream)
        dsc_loss = F.cross_entropy(logits, xflag_cat, reduction="none")
        dsc_loss = (dsc_loss * xdwt_cat).sum() / xdwt_cat.numel()
        return u_cat, dsc_loss

    def _supervision(self, u_cat, xlbl):
        net = self.net
        if not net.u2c:
            return torch.zeros((), device=net.device)
        xlbl_cat = torch.cat([xlbl[k] for k in net.keys])
        labelled = xlbl_cat >= 0  # masked mean without boolean indexing (static shapes)
        ce = F.cross_entropy(net.u2c(u_cat), xlbl_cat.clamp(min=0), reduction="none")
        return (ce * labelled).sum() / labelled.sum().clamp(min=1)

    def _input_branches(self, data):
        r"""Thunks that depend on the minibatch only; they run beside the encoders."""
        return []

    def _graph_terms(self, eidx, ewt, esgn, prior):
        net = self.net
        g2v = net.g2v
        latent = g2v.vrepr.shape[1] if hasattr(g2v, "vrepr") else 0
        if getattr(self, "_unit_prior", None) is None:  # one host read, at the first step
            self._unit_prior = (float(net.prior.loc) == 0.0 and float(net.prior.std) == 1.0
                                if hasattr(net.prior, "loc") else False)
        if (hasattr(g2v, "sample_with_kl") and g2v.vrepr.is_cuda and latent <= 64
                and latent % 2 == 0 and self._unit_prior):
            # fused Gaussian head: sample and KL in one launch (prior is N(0, 1), sc.py:640-660)
            vsamp, kl_sum = g2v.sample_with_kl(self.eidx, self.enorm, self.esgn,
                                               self.noise(g2v.vrepr, "v"))
            g_kl = kl_sum / (vsamp.shape[0] * vsamp.shape[0])
            from .. import ops
            if ops.V_GRAD_SINK is not None:
                # early backward pass (see compute_losses): the consumers of the vertex embeddings see a fresh
                # leaf, so that no autograd edge leads from the objective back into the graph branch -- its
                # gradient arrives through ``ops.V_GRAD_SINK`` instead
                self._graph_branch = (vsamp, kl_sum)
                vsamp = vsamp.detach().requires_grad_(True)
                g_kl = g_kl.detach()
        else:
            v = g2v(self.eidx, self.enorm, self.esgn)
            vsamp = v.mean + self.noise(v.mean, "v") * v.stddev
            g_kl = _kl_to_prior(v, prior) / vsamp.shape[0]
        g_nll = self._graph_nll(vsamp, eidx, esgn, ewt)
        return vsamp, g_nll, g_kl

    def _graph_nll(self, vsamp, eidx, esgn, ewt):
        r"""Edge reconstruction term (``scglue.py:331-338``).  Only the objective's last line reads it,
        while the data decoders wait for ``vsamp``: it is issued on a stream of its own, forked where
        ``vsamp`` is complete, so that the branch that produced ``vsamp`` can be joined -- and the
        decoders can start -- without waiting for the edge kernels and the incidence sort of their
        backward pass.  ``_join_graph_nll`` is called before the value is read."""
        net = self.net
        device = net.device
        # inside a training step the term's coefficient in gen_loss is known: its gradient is computed with
        # the forward kernels (see ``ops.graph_nll``), off the critical path of the backward pass
        up = self.lam_graph * len(net.keys) if device.type == "cuda" else None
        if device.type != "cuda" or not config.MULTI_STREAM:
            return net.v2g(vsamp, eidx, esgn).balanced_nll(ewt, upstream=up)
        if self._nll_stream is None:
            self._nll_stream = torch.cuda.Stream(device)
        side = self._nll_stream
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            vsamp.record_stream(side)
            g_nll = net.v2g(vsamp, eidx, esgn).balanced_nll(ewt, upstream=up)
        self._nll_pending = True
        return g_nll

    def _early_backward_ok(self) -> bool:
        r"""Whether the graph branch can be back-propagated right behind the decoders: a training step's
        generator pass on CUDA, the fused Gaussian head, and every decoder one of the kernel-backed ones
        with a non-zero coefficient -- then every gradient with respect to ``vsamp`` is computed in the
        forward pass with its coefficient folded in (``ops.V_GRAD_SINK``), none flows back through autograd."""
        net = self.net
        if not (self._in_generator_step and torch.is_grad_enabled() and net.device.type == "cuda"
                and config.MULTI_STREAM and config.EARLY_GRAPH_BACKWARD):
            return False
        g2v = net.g2v
        if not (hasattr(g2v, "sample_with_kl") and hasattr(g2v, "vrepr") and g2v.vrepr.shape[1] <= 64
                and g2v.vrepr.shape[1] % 2 == 0 and getattr(self, "_unit_prior", None)):
            return False
        if not self.lam_graph or not self.lam_data:
            return False
        return all(isinstance(net.u2x[k], sc._FusedDecoder) and self.modality_weight[k] for k in net.keys)

    def _early_alignment_ok(self) -> bool:
        r"""The generator's alignment term can be back-propagated as soon as it is computed: training step,
        coefficient folded into the fused cross-entropy, nothing else reads the concatenated means."""
        net = self.net
        return (self._in_generator_step and torch.is_grad_enabled() and net.device.type == "cuda"
                and config.MULTI_STREAM and config.EARLY_ALIGN_BACKWARD and config.USE_FUSED_OBJECTIVE
                and not net.u2c and not self.freeze_u and len(net.keys) <= 32)

    def _early_graph_backward(self, v_grads, vnum: int) -> None:
        r"""``vsamp.backward(sum of the handed-out gradients)`` and the KL term of the vertex posterior, on a
        stream of its own: autograd replays the graph branch's backward kernels (Gaussian head, weight-gradient
        reduce, transposed GraphConv) on that branch's stream; the stream this is called from only forks."""
        from .. import ops
        net = self.net
        branch = self._graph_branch
        self._graph_branch = None
        if not v_grads or branch is None:
            raise RuntimeError("early graph backward: no gradient was handed out")
        device = net.device
        main = torch.cuda.current_stream(device)
        if self._early_stream is None:
            self._early_stream = torch.cuda.Stream(device)
        side = self._early_stream
        side.wait_stream(main)
        vsamp, kl_sum = branch
        coeff = self.lam_graph * len(net.keys) * self.lam_kl / (vnum * vnum)  # d gen_loss / d kl_sum
        with torch.cuda.stream(side):
            total = v_grads[0]
            total.record_stream(side)
            for g in v_grads[1:]:
                g.record_stream(side)
                total.add_(g)
            torch.autograd.backward([vsamp, kl_sum], [total, ops._upstream_scalar(coeff, device).reshape(kl_sum.shape)])
            # data parallel: the vertex embeddings are 10 of the 12 MB of generator gradients -- exchanged
            # here, beside the discriminator's and the encoders' backward passes, instead of behind them
            # (one GPU: copying them into the flat gradient buffer here instead of with the rest in front of the
            # optimizer was tried and cost 13 us per step)
            self._early_reduced = None
            if gdist.is_distributed() and hasattr(self.vae_optim, "all_reduce_subset"):
                self._early_reduced = self.vae_optim.all_reduce_subset(list(net.g2v.parameters()))
        self._early_pending = True

    def _join_early_backward(self) -> None:
        if self._early_pending:
            torch.cuda.current_stream(self.net.device).wait_stream(self._early_stream)
            self._early_pending = False

    def _join_graph_nll(self, g_nll):
        if self._nll_pending:
            cur = torch.cuda.current_stream(self.net.device)
            cur.wait_stream(self._nll_stream)
            g_nll.record_stream(cur)
            self._nll_pending = False

    def _data_nll(self, k: str, usamp, vsamp, xbch, l, x, reduce: str,
                  row_weight: Optional[torch.Tensor] = None,
                  upstream: Optional[float] = None) -> torch.Tensor:
        r"""``-u2x[k](usamp, vsamp[idx_k], xbch, l).log_prob(x).<reduce>()``; kernel-backed
        decoders take the whole vertex table and gather ``idx_k`` rows themselves.  With
        ``row_weight`` ``[B]`` the result is ``-sum_i row_weight[i] sum_j log_prob[i, j]``."""
        net = self.net
        decoder, idx = net.u2x[k], getattr(net, f"{k}_idx")
        if isinstance(decoder, sc._FusedDecoder):
            # `upstream`: the coefficient of this term in the objective that is differentiated
            # (gen_loss), folded into the kernel's weights instead of a pass over the gradients
            return decoder(usamp, vsamp, xbch, l, vidx=idx).nll(x, reduce=reduce,
                                                                row_weight=row_weight,
                                                                upstream=upstream)
        log_prob = decoder(usamp, vsamp[idx], xbch, l).log_prob(x)
        if row_weight is not None:
            return -(log_prob * row_weight.unsqueeze(1)).sum()
        return -(log_prob.nanmean() if reduce == "nanmean" else log_prob.mean())

    DATA_REDUCE = "nanmean"  # scglue.py:345

    def compute_losses(self, data, epoch: int, dsc_only: bool = False,
                       dsc_update: Optional[Callable[[torch.Tensor], None]] = None
                       ) -> Mapping[str, torch.Tensor]:
        r"""Objective of one pass (``scglue.py:292-376``).  With ``dsc_update`` given, this call
        is a whole training step: the discriminator pass (encoders forward, discriminator loss,
        ``dsc_update(dsc_loss)`` = backward + optimizer step) runs first *in program order* but
        on its own stream, beside the part of the generator pass that does not read the
        discriminator (data encoders, graph branch, decoders); the generator's alignment term
        waits for the updated discriminator."""
        net = self.net
        x, xrep, xbch, xlbl, xdwt, xflag, eidx, ewt, esgn = data
        prior = net.prior()
        dsc_stream = None
        cat_inputs = None
        if dsc_update is not None:
            # encoders, discriminator pass first: BatchNorm statistics and dropout streams advance
            # in the reference's order because each modality keeps its stream for both passes
            with torch.no_grad():
                u_d, _, _ = self._encode(x, xrep, True, before_join=lambda: self._alignment_inputs(xbch, xdwt, xflag))
            cat_inputs = self._encode_aux
            main = torch.cuda.current_stream(net.device)
            if self._dsc_stream is None:
                self._dsc_stream = torch.cuda.Stream(net.device)
            dsc_stream = self._dsc_stream
            dsc_stream.wait_stream(main)
            with torch.cuda.stream(dsc_stream):
                for k in net.keys:  # allocated on other streams, consumed here
                    u_d[k].mean.record_stream(dsc_stream)
                for t in cat_inputs:
                    t.record_stream(dsc_stream)
                _, dsc_loss_d = self._alignment(u_d, xbch, xdwt, xflag, xlbl, epoch, cat_inputs,
                                                upstream=self.lam_align)
                dsc_update(self.lam_align * dsc_loss_d)
            del u_d, dsc_loss_d
        if cat_inputs is None:
            cat_inputs = self._alignment_inputs(xbch, xdwt, xflag)
        if dsc_only:
            # only the discriminator is updated in this pass: the encoders run forward (BatchNorm
            # statistics and dropout streams advance as in the reference) but are not
            # differentiated -- the reference computes and then discards those gradients
            # (``scglue.py:386-390``: zero_grad before the generator pass)
            with torch.no_grad():
                u, l, usamp = self._encode(x, xrep, dsc_only)
            u_cat, dsc_loss = self._alignment(u, xbch, xdwt, xflag, xlbl, epoch, cat_inputs, upstream=self.lam_align)
            return {"dsc_loss": self.lam_align * dsc_loss}
        # the graph branch (propagation, Gaussian head, edge reconstruction) is independent of
        # the data encoders: it runs beside them
        from .. import ops
        early = self._early_backward_ok()
        self._graph_branch = None
        ops.V_GRAD_SINK = [] if early else None
        # what reads the encoders' outputs only -- the paired mean embedding + cosine term, the KL of the
        # posteriors -- is issued when the encoders have joined: the graph branch is the long one (Gaussian
        # head), and these two dozen small launches used to sit between its end and the decoders' start
        # (the mean embedding) and on the objective's tail behind the decoders (the KL)
        x_kl = {}

        def after_encoders(u, l, usamp):
            self._prepare_extra(usamp)
            x_kl.update({k: self._data_kl(u[k], prior, x[k].shape[1]) for k in net.keys})
        u, l, usamp, (vsamp, g_nll, g_kl) = self._encode(
            x, xrep, dsc_only, extra=[lambda: self._graph_terms(eidx, ewt, esgn, prior)] + self._input_branches(data),
            after_encoders=after_encoders)
        # one stream per modality: the reconstruction term and, for paired data, the joint-cross /
        # real-cross terms of a modality share its observations, feature embeddings and decoder
        # parameters and run as ONE fused launch sequence (the feature table is staged once, the
        # observations are read from HBM once); the modalities run side by side, so the small one
        # (16 CTAs for 2 000 genes) fills the SMs that the tail of the large one leaves idle
        # the alignment term needs the encoders' posteriors and the updated discriminator, not the
        # decoders: issued on the discriminator stream (behind the update), it runs beside them
        main = torch.cuda.current_stream(net.device) if net.device.type == "cuda" else None
        if dsc_stream is not None:
            dsc_stream.wait_stream(main)
            with torch.cuda.stream(dsc_stream):
                for k in net.keys:
                    u[k].mean.record_stream(dsc_stream)
                # (vi-b) the alignment term's gradient is computed with its forward pass, coefficient folded
                # (`ops.weighted_cross_entropy`): the discriminator is back-propagated right here, down to
                # detached copies of the posterior means, while the current stream goes on to the decoders
                # and the objective's tail; the gradients join the rest at the means (`_InjectGrad`)
                early_align = self._early_alignment_ok()
                u_align = {k: _Posterior(u[k].mean.detach().requires_grad_(True)) for k in net.keys} \
                    if early_align else u
                u_cat, dsc_loss = self._alignment(u_align, xbch, xdwt, xflag, xlbl, epoch, cat_inputs,
                                                  upstream=-self.lam_align)
                align_inject = None
                if early_align:
                    align_inject = _InjectGrad.apply([main] + list(self._side_streams),
                                                     [u_align[k].mean for k in net.keys], *[u[k].mean for k in net.keys])
                align_ready = torch.cuda.Event()
                align_ready.record(dsc_stream)
                if early_align:
                    for p in net.du.parameters():  # (this pass's discriminator gradients are not used: no
                        p.grad = None              # accumulation into the update's)
                    torch.autograd.backward([dsc_loss])
                    u_cat, dsc_loss = u_cat.detach(), dsc_loss.detach()
                    self._early_align_done = True
        try:
            results = self._parallel([lambda k=k: self._modality_nlls(k, data, usamp, vsamp, l) for k in net.keys])
        finally:
            v_grads, ops.V_GRAD_SINK = ops.V_GRAD_SINK, None
        if early:
            # every gradient that reaches `vsamp` exists by now: the graph branch's backward pass starts here,
            # beside the alignment term, the objective tail and the discriminator / encoder backward passes
            self._join_graph_nll(g_nll)
            self._early_graph_backward(v_grads, vsamp.shape[0])
        x_nll, self._extra_parts = {}, {k: {} for k in net.keys}
        for k, terms in zip(net.keys, results):
            for kind, res in terms:  # (several real-cross terms add up in term order)
                if kind == "nll":
                    x_nll[k] = res
                else:
                    parts = self._extra_parts[k]
                    parts[kind] = res if kind not in parts else parts[kind] + res
        if dsc_stream is not None:
            # (the loss value is there once the forward pass is; the early backward pass behind it is joined by
            # autograd, where its gradients are consumed)
            main.wait_event(align_ready) if align_inject is not None else main.wait_stream(dsc_stream)
            u_cat.record_stream(main), dsc_loss.record_stream(main)
        else:
            align_inject = None
            u_cat, dsc_loss = self._alignment(u, xbch, xdwt, xflag, xlbl, epoch, cat_inputs,
                                              upstream=-self.lam_align)
        self._join_graph_nll(g_nll)
        sup_loss = self._supervision(u_cat, xlbl)
        extra_comps, extra_terms = self._extra_losses(data, usamp, vsamp, l)
        # nothing that carries an autograd graph may outlive the step on `self`: it would keep
        # this step's AccumulateGrad nodes (bound to the current streams) alive into a capture
        self._extra_parts = self._shared_extra = self._pair_w = None

        # The objective and every reported term are fixed linear combinations of a dozen scalar
        # components (``scglue.py:349-376``; paired terms ``scglue.py:654-672``).  Evaluated one
        # scalar op at a time that is ~60 tiny launches forward + backward on the serial tail of
        # the step; here: one stack, one matrix-vector product each way.
        comps = {"dsc": dsc_loss, "g_nll": g_nll, "g_kl": g_kl, "sup": sup_loss}
        if align_inject is not None:
            align_inject.record_stream(main)
            comps["align_inject"] = align_inject
        for k in net.keys:
            comps[f"x_{k}_nll"], comps[f"x_{k}_kl"] = x_nll[k], x_kl[k]
        comps.update(extra_comps)
        n_mod = len(net.keys)
        rows = {"dsc_loss": {"dsc": 1.0}, "vae_loss": None, "gen_loss": None,  # (filled in below)
                "g_nll": {"g_nll": 1.0}, "g_kl": {"g_kl": 1.0},
                "g_elbo": {"g_nll": 1.0, "g_kl": self.lam_kl}}
        vae = {"g_nll": self.lam_graph * n_mod, "g_kl": self.lam_graph * n_mod * self.lam_kl,
               "sup": self.lam_sup}
        for k in net.keys:
            w = self.lam_data * self.modality_weight[k]
            rows[f"x_{k}_nll"] = {f"x_{k}_nll": 1.0}
            rows[f"x_{k}_kl"] = {f"x_{k}_kl": 1.0}
            rows[f"x_{k}_elbo"] = {f"x_{k}_nll": 1.0, f"x_{k}_kl": self.lam_kl}
            vae[f"x_{k}_nll"], vae[f"x_{k}_kl"] = w, w * self.lam_kl
        for name, (weight, row) in extra_terms.items():
            rows[name] = row
            for comp, coeff in row.items():
                vae[comp] = vae.get(comp, 0.0) + weight * coeff
        rows["vae_loss"] = vae
        rows["gen_loss"] = dict(vae, dsc=-self.lam_align)
        if align_inject is not None:
            rows["gen_loss"]["align_inject"] = 1.0
        if net.u2c:
            rows["sup_loss"] = {"sup": 1.0}
        return self._combine(comps, rows)

    def _data_kl(self, u: D.Normal, prior: D.Normal, n_features: int) -> torch.Tensor:
        r"""``kl_divergence(u, prior).sum(dim=1).mean() / n_features`` (``scglue.py:349-353``);
        one fused launch each way for the standard-normal prior."""
        if getattr(self, "_unit_prior", None) and u.mean.is_cuda and config.USE_FUSED_OBJECTIVE:
            from .. import ops
            return ops.kl_unit_normal(u.mean, u.stddev, 1.0 / (u.mean.shape[0] * n_features))
        return _kl_to_prior(u, prior) / n_features

    def _combine(self, comps: Mapping[str, torch.Tensor],
                 rows: Mapping[str, Mapping[str, float]]) -> Mapping[str, torch.Tensor]:
        r"""``{name: sum_c rows[name][c] * comps[c]}`` as one matrix-vector product; the coefficient
        matrix lives on the device, cached per set of loss weights (a change of weights is a new
        step-graph key, so the matrix is always built outside a capture)."""
        names, keys = list(rows), list(comps)
        flat = tuple(tuple(float(rows[n].get(c, 0.0)) for c in keys) for n in names)
        cache = self.__dict__.setdefault("_combine_cache", {})
        cache_key = (tuple(names), tuple(keys), flat)
        mat = cache.get(cache_key)
        if mat is None:
            if len(cache) > 8:
                cache.clear()
            mat = cache[cache_key] = torch.tensor(flat, dtype=torch.float32, device=self.net.device)
        return LossVector(names, mat @ torch.stack([comps[c].reshape(()) for c in keys]))

    def _modality_nlls(self, k: str, data, usamp, vsamp, l):
        r"""Every decoder-based loss term whose *target* modality is ``k`` as ``[(kind, value), ...]``:
        the reconstruction term ``"nll"`` (``scglue.py:342-347``) followed by the extra terms of
        :meth:`_extra_terms_for`.  Kernel-backed decoders evaluate all of them in one fused launch
        sequence (``ops.decoder_nll_multi``); other decoders term by term."""
        net = self.net
        x, xrep, xbch, *_ = data
        specs = [("nll", usamp[k], None, self.lam_data * self.modality_weight[k])]
        specs += self._extra_terms_for(k, data, usamp, vsamp, l)
        decoder, idx = net.u2x[k], getattr(net, f"{k}_idx")
        fused = (isinstance(decoder, sc._FusedDecoder) and config.USE_FUSED_TERMS and len(specs) > 1
                 and all(c for *_, c in specs))
        if fused:
            from .. import ops
            rws = [w for _, _, w, _ in specs]  # (None: the plain ``DATA_REDUCE`` of the reconstruction term)
            values = ops.decoder_nll_multi(
                decoder.FAMILY, [u for _, u, _, _ in specs], vsamp, decoder.scale_lin, decoder.bias,
                getattr(decoder, decoder.DISPERSION), decoder.zi_logits if decoder.ZERO_INFLATED else None,
                b=xbch[k], l=l[k], x=x[k], vidx=idx, reduce=self.DATA_REDUCE, row_weights=rws,
                upstreams=[c for *_, c in specs])
            return [(kind, val) for (kind, *_), val in zip(specs, values)]
        return [(kind, self._data_nll(k, u, vsamp, xbch[k], l[k], x[k],
                                      self.DATA_REDUCE if w is None else "mean", row_weight=w, upstream=c))
                for kind, u, w, c in specs]

    def _prepare_extra(self, usamp) -> None:
        r"""Quantities shared by the per-modality extra terms (computed before the fork)."""

    def _extra_terms_for(self, k: str, data, usamp, vsamp, l):
        r"""Extra decoder-based loss terms whose *target* modality is ``k``, as
        ``[(kind, u [B, K], row_weight [B], coefficient in gen_loss), ...]``; combined by
        :meth:`_extra_losses`."""
        return []

    def _extra_losses(self, data, usamp, vsamp, l):
        r"""``(components, terms)``: extra scalar components ``{name: tensor}`` and the reported /
        weighted terms built from them ``{term: (weight in vae_loss, {component: coefficient})}``."""
        return {}, {}

    # -- steps -------------------------------------------------------------------------------
    def train_step(self, engine, data: List[torch.Tensor]) -> Mapping[str, torch.Tensor]:
        r"""One discriminator + one generator update.  On CUDA, steps whose input signature
        has been seen before are replayed from a captured graph (:class:`StepGraphs`)."""
        data = self.unpack_data(data)
        if self.graphs.usable():
            return self.graphs.step(engine, data)
        return self.eager_train_step(engine, data)

    def eager_train_step(self, engine, data: List[torch.Tensor]) -> Mapping[str, torch.Tensor]:
        self.net.train()
        data = self.format_data(data)
        epoch = engine.state.epoch
        dsc_update = None
        self.net.zero_grad(set_to_none=True)  # (the generator pass may start accumulating before its backward())
        if self.freeze_u:
            self.net.x2u.apply(freeze_running_stats)
            self.net.du.apply(freeze_running_stats)
        elif (self.dsc_steps == 1 and config.MULTI_STREAM and config.OVERLAP_DISCRIMINATOR
              and self.net.device.type == "cuda"):
            def dsc_update(scaled_dsc_loss):  # runs on the discriminator stream
                self.net.zero_grad(set_to_none=True)
                scaled_dsc_loss.backward()
                self._sync_grads("dsc")
                self.dsc_optim.step()
        else:  # discriminator step
            for _ in range(self.dsc_steps):
                losses = self.compute_losses(data, epoch, dsc_only=True)
                self.net.zero_grad(set_to_none=True)
                losses["dsc_loss"].backward()  # already scaled by lam_align
                self._sync_grads("dsc")
                self.dsc_optim.step()
        self._in_generator_step = True
        try:
            losses = self.compute_losses(data, epoch, dsc_update=dsc_update)  # generator step
        finally:
            self._in_generator_step = False
            from .. import ops
            ops.V_GRAD_SINK = None
            self._graph_branch = None
        if self._early_align_done:
            # the discriminator holds the generator pass's gradients already (cleared before that early
            # backward pass); nothing else has a gradient but the graph branch, if it went early too
            self._early_align_done = False
        elif self._early_pending:
            # the graph branch's gradients are already accumulating: only the discriminator's (left over
            # from its own update, not part of the generator's optimizer) are cleared
            for p in self.net.du.parameters():
                p.grad = None
        else:
            self.net.zero_grad(set_to_none=True)
        losses["gen_loss"].backward()
        self._join_early_backward()
        self._sync_grads("vae", skip=self._early_reduced)
        self._early_reduced = None
        self.vae_optim.step()
        # detached: the returned values feed metrics only, and a live autograd graph would keep
        # this step's AccumulateGrad nodes (bound to this stream) alive into a later capture
        if isinstance(losses, LossVector):
            return losses.detach()
        return {name: value.detach() for name, value in losses.items()}

    def val_step(self, engine, data: List[torch.Tensor]) -> Mapping[str, torch.Tensor]:
        r"""Validation step (``scglue.py:403-408``); replayed from a captured graph like the training
        step once its input signature has been seen."""
        data = self.unpack_data(data)
        if self.graphs.usable():
            return self.graphs.step(engine, data, mode="val")
        return self.eager_val_step(engine, data)

    @torch.no_grad()
    def eager_val_step(self, engine, data: List[torch.Tensor]) -> Mapping[str, torch.Tensor]:
        self.net.eval()
        losses = self.compute_losses(self.format_data(data), engine.state.epoch)
        if isinstance(losses, LossVector):
            return losses.detach()
        return {name: value.detach() for name, value in losses.items()}

    def __repr__(self) -> str:
        indent = lambda o: repr(o).replace("    ", "  ").replace("\n", "\n  ")
        return (f"{type(self).__name__}(\n  lam_graph: {self.lam_graph}\n"
                f"  lam_align: {self.lam_align}\n  vae_optim: {indent(self.vae_optim)}\n"
                f"  dsc_optim: {indent(self.dsc_optim)}\n  freeze_u: {self.freeze_u}\n)")


@logged
class PairedSCGLUETrainer(SCGLUETrainer):
    r"""
    Trainer for partially paired data (``scglue.py:423-701``): adds joint-cross,
    real-cross and cosine losses over the cells observed in several modalities.  The
    reference takes the subsets with boolean-mask indexing (dynamic shapes, an implicit host
    synchronisation per subset); here they are per-cell weights of the fused decoder, so the
    step has static shapes and can be captured in a CUDA graph.
    """

    PAIRED = True
    DATA_REDUCE = "mean"  # scglue.py:603

    def __init__(self, net: SCGLUE, lam_data: float = None, lam_kl: float = None,
                 lam_graph: float = None, lam_align: float = None, lam_sup: float = None,
                 lam_joint_cross: float = None, lam_real_cross: float = None,
                 lam_cos: float = None, dsc_steps: int = None, normalize_u: bool = None,
                 modality_weight: Mapping[str, float] = None, optim: str = None,
                 lr: float = None, **kwargs) -> None:
        super().__init__(net, lam_data=lam_data, lam_kl=lam_kl, lam_graph=lam_graph,
                         lam_align=lam_align, lam_sup=lam_sup, dsc_steps=dsc_steps,
                         normalize_u=normalize_u, modality_weight=modality_weight, optim=optim,
                         lr=lr, **kwargs)
        for name, val in dict(lam_joint_cross=lam_joint_cross, lam_real_cross=lam_real_cross,
                              lam_cos=lam_cos).items():
            if val is None:
                raise ValueError(f"`{name}` must be specified!")
        self.lam_joint_cross, self.lam_real_cross, self.lam_cos = (
            lam_joint_cross, lam_real_cross, lam_cos)
        self.required_losses += ["joint_cross_loss", "real_cross_loss", "cos_loss"]

    def compute_losses(self, data, epoch: int, dsc_only: bool = False, dsc_update=None):
        x, xrep, xbch, xlbl, xdwt, xflag, pmsk, eidx, ewt, esgn = data
        self._pmsk = pmsk
        return super().compute_losses((x, xrep, xbch, xlbl, xdwt, xflag, eidx, ewt, esgn),
                                      epoch, dsc_only=dsc_only, dsc_update=dsc_update)

    def _input_branches(self, data):
        return [lambda: self._pair_weights(data[0])]

    def _pair_weights(self, x) -> None:
        r"""Per-cell weights ``mask / (n_masked * n_features)`` of the joint-cross and real-cross terms
        (see :meth:`_extra_terms_for`).  They depend on the pairing mask only -- a dozen tiny launches
        that used to sit between the encoders and the decoders: computed beside the encoders instead."""
        keys = self.net.keys
        pf = self._pmsk.T.to(torch.float32)
        weight = lambda maskf, k: maskf / (maskf.sum().clamp(min=1.0) * x[k].shape[1])
        table = {}
        for i, k in enumerate(keys):
            if self.lam_joint_cross:
                table[(i,)] = weight(pf[i], k)
            if self.lam_real_cross:
                for j in range(len(keys)):
                    if i != j:
                        table[(i, j)] = weight(pf[i] * pf[j], k)
        self._pair_w = table

    def _prepare_extra(self, usamp) -> None:
        self._shared_extra = self._umean(usamp)

    def _umean(self, usamp):
        r"""Mean embedding of each cell over the modalities it is observed in
        (``scglue.py:618-622``) and the cosine term between every modality's embedding and that
        mean (``scglue.py:654-660``); computed once per step, on the main stream.  On CUDA both come
        out of one fused launch (and one for their backward)."""
        keys = self.net.keys
        pmsk = self._pmsk.T  # [n_modalities, B]
        pf = pmsk.to(torch.float32)
        ustack = torch.stack([usamp[k] for k in keys])
        if (ustack.is_cuda and not self.normalize_u and len(keys) <= 8 and ustack.shape[2] <= 64
                and config.USE_FUSED_OBJECTIVE):
            from .. import ops
            umean, cos = ops.paired_mean_cos(ustack, pf.contiguous())
            return ustack, umean, pf, cos
        pstack = pmsk.unsqueeze(2).expand_as(ustack)
        umean = (ustack * pstack).sum(dim=0) / pstack.sum(dim=0)
        if self.normalize_u:
            umean = F.normalize(umean, dim=1)
        zero = torch.zeros((), device=ustack.device)
        cos = zero
        if self.lam_cos:
            for i in range(len(keys)):
                n = pf[i].sum()
                sim = (F.cosine_similarity(ustack[i], umean) * pf[i]).sum() / n.clamp(min=1.0)
                cos = cos + torch.where(n > 0, 1 - sim, zero)
        return ustack, umean, pf, cos

    def _extra_terms_for(self, k: str, data, usamp, vsamp, l):
        r"""Joint-cross and real-cross reconstruction of modality ``k`` (``scglue.py:624-652``) as
        ``[("joint" | "real", u, row_weight, coefficient), ...]``.  Row subsets (observed in ``k``;
        observed in both modalities of an ordered pair) enter as per-cell weights
        ``mask / (n_masked * n_features)``: the mean over the subset without compacting rows --
        static shapes, no host read of the subset sizes."""
        net = self.net
        x = data[0]
        keys = net.keys
        i = keys.index(k)
        _, umean, pf, _ = self._shared_extra
        table = getattr(self, "_pair_w", None) or {}
        weight = lambda maskf, key: table[key] if key in table else \
            maskf / (maskf.sum().clamp(min=1.0) * x[k].shape[1])
        terms = []
        if self.lam_joint_cross:
            terms.append(("joint", umean, weight(pf[i], (i,)), self.lam_joint_cross * self.modality_weight[k]))
        if self.lam_real_cross:
            for j, ks in enumerate(keys):
                if i != j:
                    terms.append(("real", usamp[ks], weight(pf[i] * pf[j], (i, j)),
                                  self.lam_real_cross * self.modality_weight[k]))
        return terms

    def _extra_losses(self, data, usamp, vsamp, l):
        keys = self.net.keys
        ustack, umean, pf, cos = self._shared_extra
        comps, joint_row, real_row = {}, {}, {}
        for k in keys:
            parts = self._extra_parts[k]
            if "joint" in parts:
                comps[f"joint_{k}"], joint_row[f"joint_{k}"] = parts["joint"], self.modality_weight[k]
            if "real" in parts:
                comps[f"real_{k}"], real_row[f"real_{k}"] = parts["real"], self.modality_weight[k]
        comps["cos"] = cos
        return comps, {"joint_cross_loss": (self.lam_joint_cross, joint_row),
                       "real_cross_loss": (self.lam_real_cross, real_row),
                       "cos_loss": (self.lam_cos, {"cos": 1.0})}


# ------------------------------------------------------------------- public API


@logged
class SCGLUEModel(Model):
    r"""
    GLUE model for single-cell multi-omics data integration (``scglue.py:707-1362``).

    Parameters
    ----------
    adatas
        Configured datasets indexed by modality name (see ``configure_dataset``)
    vertices
        Guidance-graph vertices; must cover the features of every modality
    latent_dim, h_depth, h_dim, dropout
        Latent size (<= 64 for the fused decoders) and encoder / discriminator MLP shape
    disc_norm
        Standardise discriminator inputs
    shared_batches
        Batches are shared across modalities (discriminator sees the batch one-hot)
    random_seed
        Seed for initialisation and all sampling
    """

    NET_TYPE = SCGLUE
    TRAINER_TYPE = SCGLUETrainer

    GRAPH_BATCHES: int = 32  # graph minibatches per graph epoch
    # optimisation progress (learning rate * iterations) behind the AUTO rules
    ALIGN_BURNIN_PRG: float = 8.0
    MAX_EPOCHS_PRG: float = 48.0
    PATIENCE_PRG: float = 4.0
    REDUCE_LR_PATIENCE_PRG: float = 2.0

    def __init__(self, adatas: Mapping[str, "AnnData"], vertices: List[str], latent_dim: int = 50,
                 h_depth: int = 2, h_dim: int = 256, dropout: float = 0.2,
                 disc_norm: bool = False, shared_batches: bool = False,
                 random_seed: int = 0) -> None:
        self.vertices = pd.Index(vertices)
        self.random_seed = random_seed
        torch.manual_seed(self.random_seed)

        # construction order fixes the RNG stream of the default initialisers:
        # g2v, v2g, then per modality encoder + decoder, then du, prior, u2c (scglue.py:767-841)
        g2v = sc.GraphEncoder(self.vertices.size, latent_dim)
        v2g = sc.GraphDecoder()
        self.modalities, idx, x2u, u2x, all_ct = {}, {}, {}, {}, set()
        for k, adata in adatas.items():
            if config.ANNDATA_KEY not in adata.uns:
                raise ValueError(f"The '{k}' dataset has not been configured. "
                                 f"Please call `configure_dataset` first!")
            cfg = copy.deepcopy(adata.uns[config.ANNDATA_KEY])
            if cfg["rep_dim"] and cfg["rep_dim"] < latent_dim:
                self.logger.warning("It is recommended that `use_rep` dimensionality "
                                    "be equal or larger than `latent_dim`.")
            feature_idx = self.vertices.get_indexer(cfg["features"]).astype(np.int64)
            if feature_idx.min() < 0:
                raise ValueError("Not all modality features exist in the graph!")
            idx[k] = torch.as_tensor(feature_idx)
            x2u[k] = _ENCODER_MAP[cfg["prob_model"]](
                cfg["rep_dim"] or len(cfg["features"]), latent_dim, h_depth=h_depth, h_dim=h_dim,
                dropout=dropout)
            cfg["batches"] = pd.Index([]) if cfg["batches"] is None else pd.Index(cfg["batches"])
            u2x[k] = _DECODER_MAP[cfg["prob_model"]](
                len(cfg["features"]), n_batches=max(cfg["batches"].size, 1))
            all_ct = all_ct.union(set() if cfg["cell_types"] is None else cfg["cell_types"])
            self.modalities[k] = cfg
        all_ct = pd.Index(all_ct).sort_values()
        for cfg in self.modalities.values():
            cfg["cell_types"] = all_ct
        if shared_batches:
            batch_sets = [cfg["batches"] for cfg in self.modalities.values()]
            for batches in batch_sets:
                if not np.array_equal(batches, batch_sets[0]):
                    raise RuntimeError("Batches must match when using `shared_batches`!")
            du_n_batches = batch_sets[0].size
        else:
            du_n_batches = 0
        du = sc.Discriminator(latent_dim, len(self.modalities), n_batches=du_n_batches,
                              h_depth=h_depth, h_dim=h_dim, dropout=dropout, norm=disc_norm)
        prior = sc.Prior()
        super().__init__(g2v, v2g, x2u, u2x, idx, du, prior,
                         u2c=None if all_ct.empty else sc.Classifier(latent_dim, all_ct.size))

    # -- cell embedding freezing / model adoption -------------------------------------------
    def freeze_cells(self) -> None:
        self.trainer.freeze_u = True

    def unfreeze_cells(self) -> None:
        self.trainer.freeze_u = False

    def adopt_pretrained_model(self, source: "SCGLUEModel", submodule: Optional[str] = None
                               ) -> None:
        r"""Copy every parameter / buffer that exists in ``source`` with the same name and
        shape (``scglue.py:855-887``)."""
        def descend(obj, dotted):
            for part in dotted.split("."):
                obj = getattr(obj, part)
            return obj
        src, dst = source.net, self.net
        if submodule:
            src, dst = descend(src, submodule), descend(dst, submodule)
        for name, target in chain(dst.named_parameters(), dst.named_buffers()):
            try:
                value = descend(src, name)
            except AttributeError:
                self.logger.warning("Missing: %s", name)
                continue
            if value.shape != target.shape:
                self.logger.warning("Shape mismatch: %s", name)
                continue
            with torch.no_grad():
                target.data.copy_(value.data.to(device=target.device, dtype=target.dtype))

    # -- compile / fit ----------------------------------------------------------------------
    def compile(self, lam_data: float = 1.0, lam_kl: float = 1.0, lam_graph: float = 0.02,
                lam_align: float = 0.05, lam_sup: float = 0.02, dsc_steps: int = 1,
                normalize_u: bool = False, modality_weight: Optional[Mapping[str, float]] = None,
                lr: float = 2e-3, **kwargs) -> None:
        r"""Create the trainer (RMSprop on both parameter groups, ``scglue.py:889-942``)."""
        if modality_weight is None:
            modality_weight = {k: 1.0 for k in self.net.keys}
        super().compile(lam_data=lam_data, lam_kl=lam_kl, lam_graph=lam_graph,
                        lam_align=lam_align, lam_sup=lam_sup, dsc_steps=dsc_steps,
                        normalize_u=normalize_u, modality_weight=modality_weight,
                        optim="RMSprop", lr=lr, **kwargs)

    def _auto(self, name: str, value, progress: float, batch_per_epoch: float):
        if value != AUTO:
            return value
        value = max(ceil(progress / self.trainer.lr / batch_per_epoch), ceil(progress))
        self.logger.info("Setting `%s` = %d", name, value)
        return value

    def fit(self, adatas: Mapping[str, "AnnData"], graph: nx.Graph, neg_samples: int = 10,
            val_split: float = 0.1, data_batch_size: int = 128,
            graph_batch_size: Union[int, str] = AUTO, align_burnin: Union[int, str] = AUTO,
            safe_burnin: bool = True, max_epochs: Union[int, str] = AUTO,
            patience: Optional[Union[int, str]] = AUTO,
            reduce_lr_patience: Optional[Union[int, str]] = AUTO,
            wait_n_lrs: int = 1, directory=None, device_resident: bool = True) -> None:
        r"""
        Fit the model (``scglue.py:944-1055``).  ``device_resident`` keeps count matrices and
        the sampled edge stream in HBM (falls back to host-side fetching when they exceed
        ``config.DEVICE_DATA_BUDGET_BYTES``); the minibatch index streams are unchanged.
        """
        data = AnnDataset([adatas[k] for k in self.net.keys],
                          [self.modalities[k] for k in self.net.keys], mode="train")
        from ..graph import check_graph
        check_graph(graph, adatas.values(), cov="ignore", attr="error", loop="warn", sym="warn")
        graph = GraphDataset(graph, self.vertices, neg_samples=neg_samples,
                             weighted_sampling=True, deemphasize_loops=True)
        if device_resident and self.net.device.type == "cuda":
            try:
                data.to_device(self.net.device)
                graph.to_device(self.net.device)
            except MemoryError as e:
                self.logger.warning("%s", e)
        batch_per_epoch = data.size * (1 - val_split) / data_batch_size
        if graph_batch_size == AUTO:
            graph_batch_size = ceil(graph.size / self.GRAPH_BATCHES)
            self.logger.info("Setting `graph_batch_size` = %d", graph_batch_size)
        align_burnin = self._auto("align_burnin", align_burnin, self.ALIGN_BURNIN_PRG,
                                  batch_per_epoch)
        max_epochs = self._auto("max_epochs", max_epochs, self.MAX_EPOCHS_PRG, batch_per_epoch)
        patience = self._auto("patience", patience, self.PATIENCE_PRG, batch_per_epoch)
        reduce_lr_patience = self._auto("reduce_lr_patience", reduce_lr_patience,
                                        self.REDUCE_LR_PATIENCE_PRG, batch_per_epoch)
        if self.trainer.freeze_u:
            self.logger.info("Cell embeddings are frozen")
        super().fit(data, graph, val_split=val_split, data_batch_size=data_batch_size,
                    graph_batch_size=graph_batch_size, align_burnin=align_burnin,
                    safe_burnin=safe_burnin, max_epochs=max_epochs, patience=patience,
                    reduce_lr_patience=reduce_lr_patience, wait_n_lrs=wait_n_lrs,
                    random_seed=self.random_seed, directory=directory)

    @torch.no_grad()
    def get_losses(self, adatas: Mapping[str, "AnnData"], graph: nx.Graph, neg_samples: int = 10,
                   data_batch_size: int = 128, graph_batch_size: Union[int, str] = AUTO
                   ) -> Mapping[str, float]:
        r"""Loss values on the given data (``scglue.py:1057-1108``)."""
        data = AnnDataset([adatas[k] for k in self.net.keys],
                          [self.modalities[k] for k in self.net.keys], mode="train")
        graph = GraphDataset(graph, self.vertices, neg_samples=neg_samples,
                             weighted_sampling=True, deemphasize_loops=True)
        if graph_batch_size == AUTO:
            graph_batch_size = ceil(graph.size / self.GRAPH_BATCHES)
        return super().get_losses(data, graph, data_batch_size=data_batch_size,
                                  graph_batch_size=graph_batch_size,
                                  random_seed=self.random_seed)

    # -- inference ----------------------------------------------------------------------------
    @torch.no_grad()
    def encode_graph(self, graph: nx.Graph, n_sample: Optional[int] = None) -> np.ndarray:
        r"""Vertex (feature) embeddings ``[V, K]``, or ``[V, n_sample, K]`` samples
        (``scglue.py:1110-1149``).  One GraphConv + two 50x50 GEMMs."""
        self.net.eval()
        graph = GraphDataset(graph, self.vertices)
        device = self.net.device
        enorm = torch.as_tensor(normalize_edges(graph.eidx, graph.ewt), device=device)
        esgn = torch.as_tensor(graph.esgn, device=device)
        eidx = torch.as_tensor(graph.eidx, device=device)
        v = self.net.g2v(eidx, enorm, esgn)
        if n_sample:
            return torch.cat([v.sample((1,)).cpu() for _ in range(n_sample)]
                             ).permute(1, 0, 2).numpy()
        return v.mean.detach().cpu().numpy()

    def _eval_loader(self, key: str, adata, batch_size: int) -> DataLoader:
        data = AnnDataset([adata], [self.modalities[key]], mode="eval", getitem_size=batch_size)
        pin = config.DATALOADER_PIN_MEMORY and not config.CPU_ONLY and torch.cuda.is_available()
        return DataLoader(data, batch_size=1, shuffle=False,
                          num_workers=config.DATALOADER_NUM_WORKERS, pin_memory=pin,
                          drop_last=False, persistent_workers=False)

    @torch.no_grad()
    def encode_data(self, key: str, adata, batch_size: int = 128,
                    n_sample: Optional[int] = None) -> np.ndarray:
        r"""Cell embeddings ``[N, K]`` (posterior means) or ``[N, n_sample, K]`` samples
        (``scglue.py:1151-1208``)."""
        self.net.eval()
        encoder, device = self.net.x2u[key], self.net.device
        result = []
        for x, xrep, *_ in self._eval_loader(key, adata, batch_size):
            u = encoder(x.to(device, non_blocking=True), xrep.to(device, non_blocking=True),
                        lazy_normalizer=True)[0]
            if n_sample:
                result.append(u.sample((n_sample,)).cpu().permute(1, 0, 2))
            else:
                result.append(u.mean.detach().cpu())
        return torch.cat(result).numpy()

    @torch.no_grad()
    def classify_data(self, key: str, adata, batch_size: int = 128) -> pd.DataFrame:
        r"""Cell-type probabilities from the linear classifier (``scglue.py:1210-1262``)."""
        self.net.eval()
        encoder, classifier, device = self.net.x2u[key], self.net.u2c, self.net.device
        if classifier is None:
            raise RuntimeError("The model was built without cell type supervision!")
        result = []
        for x, xrep, *_ in self._eval_loader(key, adata, batch_size):
            u = encoder(x.to(device, non_blocking=True), xrep.to(device, non_blocking=True),
                        lazy_normalizer=True)[0]
            result.append(classifier(u.mean).softmax(dim=-1).cpu())
        return pd.DataFrame(torch.cat(result).numpy(), index=adata.obs_names,
                            columns=self.modalities[key]["cell_types"])

    @torch.no_grad()
    def decode_data(self, source_key: str, target_key: str, adata, graph: nx.Graph,
                    target_libsize: Optional[Union[float, np.ndarray]] = None,
                    target_batch: Optional[np.ndarray] = None, batch_size: int = 128
                    ) -> np.ndarray:
        r"""Cross-modality imputation: encode with ``source_key``, decode with
        ``target_key`` (mean of the decoded distribution, ``scglue.py:1264-1356``)."""
        l = target_libsize or 1.0
        if not isinstance(l, np.ndarray):
            l = np.asarray(l)
        l = l.squeeze()
        if l.ndim == 0:
            l = l[np.newaxis]
        elif l.ndim > 1:
            raise ValueError("`target_libsize` cannot be >1 dimensional")
        if l.size == 1:
            l = np.repeat(l, adata.shape[0])
        if l.size != adata.shape[0]:
            raise ValueError("`target_libsize` must have the same size as `adata`!")
        l = l.reshape((-1, 1)).astype(np.float32)

        use_batch = self.modalities[target_key]["use_batch"]
        batches = self.modalities[target_key]["batches"]
        if use_batch and target_batch is not None:
            target_batch = np.asarray(target_batch)
            if target_batch.size != adata.shape[0]:
                raise ValueError("`target_batch` must have the same size as `adata`!")
            b = batches.get_indexer(target_batch)
        else:
            b = np.zeros(adata.shape[0], dtype=int)

        net, device = self.net, self.net.device
        net.eval()
        u = self.encode_data(source_key, adata, batch_size=batch_size)
        v = torch.as_tensor(self.encode_graph(graph), device=device)
        v = v[getattr(net, f"{target_key}_idx")].contiguous()
        decoder = net.u2x[target_key]
        result = []
        for start in range(0, adata.shape[0], batch_size):
            s = slice(start, start + batch_size)
            u_ = torch.as_tensor(u[s], device=device)
            b_ = torch.as_tensor(b[s], dtype=torch.int64, device=device)
            l_ = torch.as_tensor(l[s], device=device)
            result.append(decoder(u_, v, b_, l_).mean.detach().cpu())
        return torch.cat(result).numpy()

    def upgrade(self) -> None:
        if hasattr(self, "domains"):
            self.logger.warning("Upgrading model generated by older versions...")
            self.modalities = getattr(self, "domains")
            delattr(self, "domains")

    def __repr__(self) -> str:
        return (f"SCGLUE model with the following network and trainer:\n\n"
                f"{repr(self.net)}\n\n{repr(self.trainer)}\n")


@logged
class PairedSCGLUEModel(SCGLUEModel):
    r"""GLUE model for partially paired data (``scglue.py:1365-1459``); cells are paired
    through ``obs_names`` when datasets are configured with ``use_obs_names=True``."""

    TRAINER_TYPE = PairedSCGLUETrainer

    def compile(self, lam_data: float = 1.0, lam_kl: float = 1.0, lam_graph: float = 0.02,
                lam_align: float = 0.05, lam_sup: float = 0.02, lam_joint_cross: float = 0.02,
                lam_real_cross: float = 0.02, lam_cos: float = 0.02, dsc_steps: int = 1,
                normalize_u: bool = False, modality_weight: Optional[Mapping[str, float]] = None,
                lr: float = 2e-3, **kwargs) -> None:
        super().compile(lam_data=lam_data, lam_kl=lam_kl, lam_graph=lam_graph,
                        lam_align=lam_align, lam_sup=lam_sup, lam_joint_cross=lam_joint_cross,
                        lam_real_cross=lam_real_cross, lam_cos=lam_cos, dsc_steps=dsc_steps,
                        normalize_u=normalize_u, modality_weight=modality_weight, lr=lr, **kwargs)
