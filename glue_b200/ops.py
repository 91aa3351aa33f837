r"""
``torch.autograd.Function`` wrappers around the native kernels.

Each op states which reference computation it replaces; parity tests live in
``tests/test_ops_gpu.py``.  PyTorch is used here for device memory, streams and
autograd bookkeeping only -- all arithmetic happens in ``libglue_b200.so``.
"""

import ctypes
from collections import OrderedDict
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import FAMILY, DecoderArgs, check, current_stream, ptr, require_cuda

# ------------------------------------------------------------------- profiling


class KernelTimer:
    r"""Optional CUDA-event bracketing of native calls (used by ``bench.py`` to time the
    dominant kernel sequence *inside* the timed training steps, on the launching stream)."""

    def __init__(self) -> None:
        self.enabled = False
        self.records = []  # (name, start_event, end_event)

    def start(self, name: str, device):
        if not self.enabled:
            return None
        ev = (name, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[1].record(torch.cuda.current_stream(device))
        return ev

    def stop(self, ev, device) -> None:
        if ev is not None:
            ev[2].record(torch.cuda.current_stream(device))
            self.records.append(ev)

    def summary(self):
        r"""``{name: (n_calls, total_ms)}``; call after ``torch.cuda.synchronize()``."""
        out = {}
        for name, a, b in self.records:
            n, t = out.get(name, (0, 0.0))
            out[name] = (n + 1, t + a.elapsed_time(b))
        return out

    def reset(self) -> None:
        self.records = []


TIMER = KernelTimer()


def kernel_launches() -> int:
    r"""Kernels launched by ``libglue_b200.so`` in this process so far."""
    return int(_lib.lib().glue_kernel_launches())


# ------------------------------------------------------------------ workspace


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"glue_b200: `{name}` must be float32, got {t.dtype}")
    return t.contiguous()


def _i64(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.int64:
        raise TypeError(f"glue_b200: `{name}` must be int64, got {t.dtype}")
    return t.contiguous()


def multi_copy(dsts, srcs) -> None:
    r"""``dst.copy_(src)`` for up to 16 pairs of contiguous device tensors of equal byte size in one
    launch (``glue_multi_copy``): the minibatch of a captured step going into its static input buffers."""
    n = len(dsts)
    if n == 0:
        return
    if n > 16:
        multi_copy(dsts[:16], srcs[:16])
        return multi_copy(dsts[16:], srcs[16:])
    dev = require_cuda(*dsts, *srcs)
    arr = ctypes.c_void_p * n
    nbytes = (ctypes.c_int64 * n)(*[d.numel() * d.element_size() for d in dsts])
    check(_lib.lib().glue_multi_copy(n, arr(*[s.data_ptr() for s in srcs]), arr(*[d.data_ptr() for d in dsts]),
                                     nbytes, current_stream(dev)), "multi_copy")


# ------------------------------------------------------------------------ CSR


class GraphCSR:
    r"""
    Guidance graph in the two CSR layouts the GraphConv kernel consumes:
    rows = edge targets (forward) and rows = edge sources (backward).

    Layout definition (bit-exact with the numpy recipe of ``oracle/sampling.py``):
    stable sort of the edges by row key, ``indptr = cumsum(bincount(key))``,
    ``col`` int32, ``val = esgn * enorm`` fp32 exactly as ``scglue/models/nn.py:53``,
    multi-edges kept.
    """

    def __init__(self, eidx: torch.Tensor, enorm: torch.Tensor, esgn: torch.Tensor, vnum: int):
        device = require_cuda(eidx, enorm, esgn)
        eidx, enorm, esgn = _i64(eidx, "eidx"), _f32(enorm, "enorm"), _f32(esgn, "esgn")
        if eidx.dim() != 2 or eidx.shape[0] != 2:
            raise ValueError("`eidx` must have shape [2, n_edges]")
        n_edges = eidx.shape[1]
        if enorm.shape != (n_edges,) or esgn.shape != (n_edges,):
            raise ValueError("`enorm` / `esgn` must have shape [n_edges]")
        self.vnum, self.n_edges, self.device = int(vnum), n_edges, device
        val = esgn * enorm  # same rounding as the reference (nn.py:53)
        self.fwd = self._build(eidx[1], eidx[0], val)
        self.bwd = self._build(eidx[0], eidx[1], val)

    def _build(self, key, other, val) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        L = _lib.lib()
        dev = self.device
        indptr = torch.empty(self.vnum + 1, dtype=torch.int32, device=dev)
        col = torch.empty(self.n_edges, dtype=torch.int32, device=dev)
        out_val = torch.empty(self.n_edges, dtype=torch.float32, device=dev)
        nbytes = L.glue_csr_build_workspace_bytes(self.n_edges, self.vnum)
        ws = _workspace(nbytes, dev)
        key, other = key.contiguous(), other.contiguous()
        check(L.glue_csr_build(ptr(key), ptr(other), ptr(val), self.n_edges, self.vnum, ptr(indptr),
                               ptr(col), ptr(out_val), ptr(ws), ws.numel(), current_stream(dev)),
              "csr_build")
        return indptr, col, out_val


_CSR_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_CSR_CACHE_SIZE = 8


def csr_for(eidx: torch.Tensor, enorm: torch.Tensor, esgn: torch.Tensor, vnum: int) -> GraphCSR:
    r"""Cached :class:`GraphCSR` for a ``(eidx, enorm, esgn)`` triple.  The graph encoder
    is called with the same full-graph tensors every step (``scglue.py:328``); the key
    holds storage address + version counter, and the cache keeps the tensors alive so an
    address cannot be recycled under it."""
    key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in (eidx, enorm, esgn)) + (vnum,)
    hit = _CSR_CACHE.get(key)
    if hit is not None:
        _CSR_CACHE.move_to_end(key)
        return hit[0]
    csr = GraphCSR(eidx, enorm, esgn, vnum)
    if torch.cuda.is_current_stream_capturing():
        # arrays allocated during a CUDA-graph capture live in the graph's private pool and hold
        # their values only while that graph replays: usable by this capture, never by anyone else
        return csr
    _CSR_CACHE[key] = (csr, eidx, enorm, esgn)
    while len(_CSR_CACHE) > _CSR_CACHE_SIZE:
        _CSR_CACHE.popitem(last=False)
    return csr


# ------------------------------------------------------------------ GraphConv


def _spmm(csr_part, inp: torch.Tensor) -> torch.Tensor:
    indptr, col, val = csr_part
    out = torch.empty_like(inp)
    ev = TIMER.start("graphconv", inp.device)
    check(_lib.lib().glue_graphconv_csr(ptr(indptr), ptr(col), ptr(val), ptr(inp), ptr(out),
                                        inp.shape[0], inp.shape[1], current_stream(inp.device)),
          "graphconv_csr")
    TIMER.stop(ev, inp.device)
    return out


class _GraphConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp: torch.Tensor, csr: GraphCSR) -> torch.Tensor:
        ctx.csr = csr
        return _spmm(csr.fwd, inp)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        return _spmm(ctx.csr.bwd, _f32(grad_out, "grad_out")), None


def graph_conv(inp: torch.Tensor, csr: GraphCSR) -> torch.Tensor:
    r"""``res[t] = sum_{e: tidx_e = t} esgn_e enorm_e inp[sidx_e]`` -- replaces the
    gather + ``scatter_add_`` of ``scglue/models/nn.py:52-57`` (and its backward)."""
    require_cuda(inp)
    inp = _f32(inp, "input")
    if inp.dim() != 2 or inp.shape[0] != csr.vnum:
        raise ValueError("`input` must have shape [n_vertices, n_features]")
    return _GraphConvFn.apply(inp, csr)


# --------------------------------------------------------------- graph encoder head


# Diagnostics: pin the kernel of the graph-encoder head's backward pass (0 = the library picks,
# 1 = FFMA, 2 = warp-level tensor cores); tests and A/B benches set it.
GRAPHENC_IMPL = 0


class _GraphEncHeadFn(torch.autograd.Function):
    r"""``(vsamp, kl_sum, mu, sigma)`` of the graph encoder's Gaussian head in one launch; the
    backward pass produces every gradient from ``d L / d vsamp`` and ``d L / d kl_sum``."""

    @staticmethod
    def forward(ctx, ptr_, w_loc, b_loc, w_std, b_std, noise):
        dev = ptr_.device
        V, K = ptr_.shape
        L = _lib.lib()
        mu, sigma, vsamp = (torch.empty_like(ptr_) for _ in range(3))
        kl = torch.empty((), dtype=torch.float32, device=dev)
        ws = _workspace(L.glue_graphenc_workspace_bytes(V, K), dev)
        ev = TIMER.start("graphenc_head_fwd", dev)
        check(L.glue_graphenc_head_fwd_impl(ptr(ptr_), ptr(w_loc), ptr(b_loc), ptr(w_std), ptr(b_std),
                                            ptr(noise), V, K, ptr(mu), ptr(sigma), ptr(vsamp), ptr(kl),
                                            ptr(ws), ws.numel(), GRAPHENC_IMPL, current_stream(dev)),
              "graphenc_head_fwd")
        TIMER.stop(ev, dev)
        ctx.save_for_backward(ptr_, w_loc, w_std, noise, mu, sigma)
        ctx.mark_non_differentiable(mu, sigma)
        return vsamp, kl, mu, sigma

    @staticmethod
    def backward(ctx, g_vsamp, g_kl, _g_mu, _g_sigma):
        ptr_, w_loc, w_std, noise, mu, sigma = ctx.saved_tensors
        dev = ptr_.device
        V, K = ptr_.shape
        L = _lib.lib()
        g_ptr = torch.empty_like(ptr_)
        g_wl, g_ws = torch.empty_like(w_loc), torch.empty_like(w_std)
        g_bl = torch.empty(K, dtype=torch.float32, device=dev)
        g_bs = torch.empty(K, dtype=torch.float32, device=dev)
        ws = _workspace(L.glue_graphenc_workspace_bytes(V, K), dev)
        g_vsamp = None if g_vsamp is None else _f32(g_vsamp, "grad_vsamp")
        g_kl = None if g_kl is None else _f32(g_kl, "grad_kl")
        check(L.glue_graphenc_head_bwd_impl(ptr(ptr_), ptr(w_loc), ptr(w_std), ptr(noise), ptr(mu),
                                            ptr(sigma), ptr(g_vsamp), ptr(g_kl), V, K, ptr(g_ptr),
                                            ptr(g_wl), ptr(g_bl), ptr(g_ws), ptr(g_bs), ptr(ws), ws.numel(),
                                            GRAPHENC_IMPL, current_stream(dev)), "graphenc_head_bwd")
        return g_ptr, g_wl, g_bl, g_ws, g_bs, None


def graph_encoder_head(ptr_, w_loc, b_loc, w_std, b_std, noise=None):
    r"""Gaussian head of the graph encoder (``scglue/models/sc.py:44-46``) fused with the
    reparameterised sample and the KL term against ``N(0, 1)``:
    returns ``(vsamp, kl_sum, mu, sigma)`` with ``vsamp = mu + noise * sigma`` and
    ``kl_sum = kl_divergence(N(mu, sigma), N(0, 1)).sum()``; ``mu`` / ``sigma`` are detached."""
    require_cuda(ptr_, w_loc, b_loc, w_std, b_std, noise)
    args = [_f32(t, n) for t, n in ((ptr_, "ptr"), (w_loc, "w_loc"), (b_loc, "b_loc"),
                                    (w_std, "w_std"), (b_std, "b_std"))]
    V, K = args[0].shape
    if K > 64 or K % 2:
        raise ValueError("glue_b200 graph encoder head supports even latent_dim <= 64")
    if noise is not None:
        noise = _f32(noise, "noise")
        if noise.shape != args[0].shape:
            raise ValueError("`noise` must have the shape of the vertex table")
    return _GraphEncHeadFn.apply(*args, noise)


# ------------------------------------------------------------- KL of the data encoders


class _KlUnitNormalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, std, scale):
        dev = mean.device
        out = torch.empty((), dtype=torch.float32, device=dev)
        check(_lib.lib().glue_kl_unit_normal_fwd(ptr(mean), ptr(std), mean.numel(), scale, ptr(out),
                                                 current_stream(dev)), "kl_unit_normal_fwd")
        ctx.save_for_backward(mean, std)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, g):
        mean, std = ctx.saved_tensors
        dev = mean.device
        g = _f32(g, "grad_kl")
        g_mean, g_std = torch.empty_like(mean), torch.empty_like(std)
        check(_lib.lib().glue_kl_unit_normal_bwd(ptr(mean), ptr(std), mean.numel(), ctx.scale, ptr(g),
                                                 ptr(g_mean), ptr(g_std), current_stream(dev)),
              "kl_unit_normal_bwd")
        return g_mean, g_std, None


def kl_unit_normal(mean: torch.Tensor, std: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    r"""``scale * kl_divergence(Normal(mean, std), Normal(0, 1)).sum()`` in one launch each way
    (``scglue/models/scglue.py:349-353`` with the standard-normal prior of ``sc.py:640-660``;
    ``scale = 1 / (batch * n_features)`` turns it into ``.sum(dim=1).mean() / n_features``)."""
    require_cuda(mean, std)
    mean, std = _f32(mean, "mean"), _f32(std, "std")
    if mean.shape != std.shape:
        raise ValueError("`mean` and `std` must have the same shape")
    return _KlUnitNormalFn.apply(mean, std, float(scale))


class _WeightedCeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, weight, scale, upstream, grad_enabled):
        dev = logits.device
        n, c = logits.shape
        out = torch.empty((), dtype=torch.float32, device=dev)
        need_grad = bool(grad_enabled and ctx.needs_input_grad[0])
        grad = torch.empty_like(logits) if need_grad else None
        check(_lib.lib().glue_weighted_ce_fwd(ptr(logits), ptr(target), ptr(weight), n, c, scale,
                                              1.0 if upstream is None else float(upstream), ptr(out), ptr(grad),
                                              current_stream(dev)), "weighted_ce_fwd")
        ctx.grad, ctx.folded = grad, upstream is not None
        return out

    @staticmethod
    def backward(ctx, g):
        grad, ctx.grad = ctx.grad, None
        if grad is None:
            raise RuntimeError("weighted_cross_entropy: no gradient was prepared (called under no_grad)")
        return (grad if ctx.folded else grad * g), None, None, None, None, None


def weighted_cross_entropy_supported(logits: torch.Tensor) -> bool:
    return logits.is_cuda and logits.dim() == 2 and logits.shape[1] <= 32 and logits.dtype == torch.float32


def weighted_cross_entropy(logits: torch.Tensor, target: torch.Tensor, weight: torch.Tensor,
                           upstream: Optional[float] = None) -> torch.Tensor:
    r"""``(F.cross_entropy(logits, target, reduction="none") * weight).sum() / weight.numel()`` -- the
    discriminator loss of ``scglue/models/scglue.py:326-329`` -- and its gradient in ONE launch.

    ``upstream``: the coefficient with which the caller's objective takes the term (``+lam_align`` in the
    discriminator update, ``-lam_align`` in the generator's loss).  It is folded into the stored gradient, which
    autograd then hands on as it is: the caller must back-propagate exactly that multiple of the result, once.
    Without it the stored gradient is scaled by the incoming one (one more launch)."""
    require_cuda(logits, target, weight)
    logits = _f32(logits, "logits")
    if logits.dim() != 2 or target.shape != (logits.shape[0],) or weight.shape != (logits.shape[0],):
        raise ValueError("`logits` [n, classes], `target` [n], `weight` [n] expected")
    if target.dtype != torch.int64:
        raise TypeError("`target` must be int64")
    return _WeightedCeFn.apply(logits, target.contiguous(), _f32(weight, "weight"), 1.0 / logits.shape[0],
                               upstream, torch.is_grad_enabled())


class _PairedMeanCosFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ustack, pmask):
        dev = ustack.device
        M, B, K = ustack.shape
        umean = torch.empty(B, K, dtype=torch.float32, device=dev)
        cos = torch.empty((), dtype=torch.float32, device=dev)
        n_k = torch.empty(M, dtype=torch.float32, device=dev)
        psim = torch.empty(M, B, dtype=torch.float32, device=dev)
        check(_lib.lib().glue_paired_mean_cos_fwd(ptr(ustack), ptr(pmask), M, B, K, ptr(umean), ptr(cos),
                                                  ptr(n_k), ptr(psim), current_stream(dev)), "paired_mean_cos_fwd")
        ctx.save_for_backward(ustack, pmask, n_k)
        return umean, cos

    @staticmethod
    def backward(ctx, g_umean, g_cos):
        ustack, pmask, n_k = ctx.saved_tensors
        dev = ustack.device
        M, B, K = ustack.shape
        g_umean = None if g_umean is None else _f32(g_umean, "grad_umean")
        g_cos = None if g_cos is None else _f32(g_cos, "grad_cos")
        g_u = torch.empty_like(ustack)
        check(_lib.lib().glue_paired_mean_cos_bwd(ptr(ustack), ptr(pmask), ptr(n_k), M, B, K, ptr(g_umean),
                                                  ptr(g_cos), ptr(g_u), current_stream(dev)),
              "paired_mean_cos_bwd")
        return g_u, None


def paired_mean_cos(ustack: torch.Tensor, pmask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    r"""``(umean, cos_loss)`` of paired cells (``scglue/models/scglue.py:618-622, 654-660``):
    ``umean = (ustack * pmask).sum(0) / pmask.sum(0)`` and
    ``cos_loss = sum_k [n_k > 0] (1 - (cosine_similarity(ustack[k], umean) * pmask[k]).sum() / n_k)``
    in one launch each way.  ``ustack``: ``[M, B, K]``, ``pmask``: ``[M, B]`` fp32 0 / 1."""
    require_cuda(ustack, pmask)
    ustack, pmask = _f32(ustack, "ustack"), _f32(pmask, "pmask")
    if ustack.dim() != 3 or pmask.shape != ustack.shape[:2]:
        raise ValueError("`ustack` must be [M, B, K] and `pmask` [M, B]")
    if ustack.shape[0] > 8 or ustack.shape[2] > 64:
        raise ValueError("glue_b200 paired mean supports <= 8 modalities and latent_dim <= 64")
    return _PairedMeanCosFn.apply(ustack, pmask)


# ------------------------------------------------------------- hidden block of the data encoders


class _ActBnDropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, gamma, beta, running_mean, running_var, num_batches_tracked, rng_state,
                slope, eps, momentum, p):
        dev = z.device
        B, H = z.shape
        out = torch.empty_like(z)
        mean = torch.empty(H, dtype=torch.float32, device=dev)
        rstd = torch.empty(H, dtype=torch.float32, device=dev)
        keep = torch.empty(B, H, dtype=torch.uint8, device=dev) if p > 0 else None
        check(_lib.lib().glue_act_bn_dropout_fwd(
            ptr(z), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(num_batches_tracked),
            ptr(rng_state), B, H, slope, eps, momentum, p, ptr(out), ptr(mean), ptr(rstd), ptr(keep),
            current_stream(dev)), "act_bn_dropout_fwd")
        ctx.save_for_backward(z, gamma, mean, rstd, keep)
        ctx.slope, ctx.p = slope, p
        ctx.mark_non_differentiable(mean, rstd)
        return out, mean, rstd, keep

    @staticmethod
    def backward(ctx, g_out, _g_mean, _g_rstd, _g_keep):
        z, gamma, mean, rstd, keep = ctx.saved_tensors
        dev = z.device
        B, H = z.shape
        g_out = _f32(g_out, "grad_out")
        g_z = torch.empty_like(z)
        g_gamma = torch.empty(H, dtype=torch.float32, device=dev)
        g_beta = torch.empty(H, dtype=torch.float32, device=dev)
        check(_lib.lib().glue_act_bn_dropout_bwd(
            ptr(z), ptr(gamma), ptr(mean), ptr(rstd), ptr(keep), ptr(g_out), B, H, ctx.slope, ctx.p,
            ptr(g_z), ptr(g_gamma), ptr(g_beta), current_stream(dev)), "act_bn_dropout_bwd")
        return g_z, g_gamma, g_beta, None, None, None, None, None, None, None, None


def act_bn_dropout(z: torch.Tensor, bn: torch.nn.BatchNorm1d, p: float, rng_state: Optional[torch.Tensor],
                   negative_slope: float = 0.2, return_aux: bool = False):
    r"""``dropout(bn(leaky_relu(z)))`` of a training-mode hidden block (``scglue/models/sc.py:62-178``)
    in one launch each way.  ``bn``'s running statistics and batch counter are updated as
    ``torch.nn.BatchNorm1d`` does; the keep mask is drawn from Philox4x32-10 keyed by
    ``rng_state`` (``uint64 [3]``: seed, call counter, ticket; advanced on the device).
    ``return_aux`` also returns ``(batch mean, 1 / sqrt(var + eps), keep mask)``."""
    require_cuda(z, bn.weight, bn.bias)
    if not bn.training or not bn.affine or bn.momentum is None:
        raise ValueError("act_bn_dropout implements training-mode affine BatchNorm1d with a fixed momentum")
    if p > 0 and rng_state is None:
        raise ValueError("`rng_state` is required when p > 0")
    z = _f32(z, "z")
    if z.dim() != 2 or z.shape[1] != bn.num_features:
        raise ValueError("`z` must be [batch, num_features]")
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    nbt = bn.num_batches_tracked if bn.track_running_stats else None
    out, mean, rstd, keep = _ActBnDropoutFn.apply(z, bn.weight, bn.bias, rm, rv, nbt, rng_state,
                                                  float(negative_slope), float(bn.eps), float(bn.momentum),
                                                  float(p))
    return (out, mean, rstd, keep) if return_aux else out


# ------------------------------------------------------------- whole data-encoder MLP


def _mlp_widths_ok(x: torch.Tensor, weights, n_head_cols: int, max_rows: int) -> bool:
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or not (2 <= x.shape[0] <= max_rows):
        return False
    widths = [x.shape[1]] + [w.shape[0] for w in weights] + [n_head_cols]
    if any(w % 2 or w > 256 or w <= 0 for w in widths):
        return False
    return x.stride(1) == 1 and x.data_ptr() % 8 == 0


def encoder_mlp_supported(x: torch.Tensor, weights, n_out: int) -> bool:
    r"""Shapes the fused encoder kernels take (``include/glue_b200.h``): at most 128 rows (BatchNorm needs the
    whole batch in one CTA), every operand width even and at most 256, fp32."""
    return _mlp_widths_ok(x, weights, 2 * n_out, 128)


def discriminator_mlp_supported(x: torch.Tensor, weights, n_out: int) -> bool:
    r"""As :func:`encoder_mlp_supported` for blocks without BatchNorm: up to 1024 rows (row tiles); any
    number of outputs."""
    return 0 < n_out <= 256 and _mlp_widths_ok(x, weights, 2, 1024)


class _MlpFn(torch.autograd.Function):
    r"""``depth`` hidden blocks + head(s).  Tensors: x, then per block (W, b, gamma, beta) -- gamma / beta
    ``None`` for blocks without BatchNorm --, then (W0, b0, W1, b1) of the heads (W1 / b1 ``None``: a single
    plain head).  ``blocks`` carries the per-block BatchNorm buffers / dropout states.  Returns
    ``(loc, std)``; ``std`` is ``None`` for a single head."""

    @staticmethod
    def forward(ctx, grad_enabled, blocks, slope, p, x, *params):
        L = _lib.lib()
        dev, st = x.device, current_stream(x.device)
        depth = len(blocks)
        B = x.shape[0]
        wloc, bloc, wstd, bstd = params[4 * depth:]
        two_heads = wstd is not None
        need_grad = bool(grad_enabled and (x.requires_grad or any(t is not None and t.requires_grad for t in params)))
        saved, keeps = [], []
        cur = x
        for i, blk in enumerate(blocks):
            W, b, gamma, beta = params[4 * i:4 * i + 4]
            H, Din = W.shape
            bn = gamma is not None
            out = torch.empty(B, H, dtype=torch.float32, device=dev)
            z = torch.empty(B, H, dtype=torch.float32, device=dev) if need_grad else None
            mean = torch.empty(H, dtype=torch.float32, device=dev) if (need_grad and bn) else None
            rstd = torch.empty(H, dtype=torch.float32, device=dev) if (need_grad and bn) else None
            keep = torch.empty(B, H, dtype=torch.uint8, device=dev) if (need_grad and p > 0) else None
            check(L.glue_mlp_hidden_fwd(
                ptr(cur), cur.stride(0), ptr(W), ptr(b), ptr(gamma), ptr(beta), ptr(blk.get("running_mean")),
                ptr(blk.get("running_var")), ptr(blk.get("num_batches_tracked")), ptr(blk["rng"]) if p > 0 else None,
                B, Din, H, slope, blk.get("eps", 0.0), blk.get("momentum", 0.0), p, ptr(z), ptr(out), ptr(mean),
                ptr(rstd), ptr(keep), st), "mlp_hidden_fwd")
            saved += [cur, z, mean, rstd]
            keeps.append(keep)
            cur = out
        O, H = wloc.shape
        loc = torch.empty(B, O, dtype=torch.float32, device=dev)
        std = torch.empty(B, O, dtype=torch.float32, device=dev) if two_heads else None
        spre = torch.empty(B, O, dtype=torch.float32, device=dev) if (need_grad and two_heads) else None
        check(L.glue_mlp_heads_fwd(ptr(cur), cur.stride(0), ptr(wloc), ptr(bloc), ptr(wstd), ptr(bstd), B, H, O,
                                   ptr(loc), ptr(std), ptr(spre), st), "mlp_heads_fwd")
        ctx.need_grad = need_grad
        if need_grad:
            # (None entries are fine for save_for_backward)
            ctx.save_for_backward(cur, spre, *saved, *keeps, *params)
            ctx.depth, ctx.slope, ctx.p = depth, slope, p
            ctx.x_grad = x.requires_grad
        if not two_heads:
            ctx.mark_non_differentiable()
        return loc, std

    @staticmethod
    def backward(ctx, g_loc, g_std):
        if not ctx.need_grad:
            raise RuntimeError("the fused MLP was run without gradient buffers")
        L = _lib.lib()
        t = ctx.saved_tensors
        depth, slope, p = ctx.depth, ctx.slope, ctx.p
        h, spre = t[0], t[1]
        saved = t[2:2 + 4 * depth]
        keeps = t[2 + 4 * depth:2 + 5 * depth]
        params = t[2 + 5 * depth:]
        wloc, bloc, wstd, bstd = params[4 * depth:]
        two_heads = wstd is not None
        dev, st = h.device, current_stream(h.device)
        B, (O, H) = h.shape[0], wloc.shape
        g_loc = torch.zeros(B, O, dtype=torch.float32, device=dev) if g_loc is None else _f32(g_loc, "grad_loc")
        if two_heads:
            g_std = torch.zeros(B, O, dtype=torch.float32, device=dev) if g_std is None else _f32(g_std, "grad_std")
        n_cols = 2 * O if two_heads else O
        G = torch.empty(B, n_cols, dtype=torch.float32, device=dev)
        g_wloc, g_bloc = torch.empty_like(wloc), torch.empty_like(bloc)
        g_wstd = torch.empty_like(wstd) if two_heads else None
        g_bstd = torch.empty_like(bstd) if two_heads else None
        check(L.glue_mlp_heads_bwd(ptr(h), h.stride(0), ptr(g_loc), ptr(g_std) if two_heads else None, ptr(spre), B, H, O,
                                   ptr(G), ptr(g_wloc), ptr(g_bloc), ptr(g_wstd), ptr(g_bstd), st), "mlp_heads_bwd")
        grads = [None] * (4 * depth)
        above, k_above, wa0, wa1, n0 = G, n_cols, wloc, wstd, O
        g_x = None
        for i in reversed(range(depth)):
            W, b, gamma, beta = params[4 * i:4 * i + 4]
            inp, z, mean, rstd = saved[4 * i:4 * i + 4]
            Hi, Din = W.shape
            want_dz = i > 0 or ctx.x_grad
            dz = torch.empty(B, Hi, dtype=torch.float32, device=dev) if want_dz else None
            gW, gb = torch.empty_like(W), torch.empty_like(b)
            gg = torch.empty_like(gamma) if gamma is not None else None
            gbeta = torch.empty_like(beta) if beta is not None else None
            check(L.glue_mlp_hidden_bwd(
                ptr(above), k_above, ptr(wa0), ptr(wa1), n0, ptr(z), ptr(gamma), ptr(mean), ptr(rstd),
                ptr(keeps[i]) if p > 0 else None, ptr(inp), inp.stride(0), Din, B, Hi, slope, p, ptr(dz), ptr(gW),
                ptr(gb), ptr(gg), ptr(gbeta), st), "mlp_hidden_bwd")
            grads[4 * i:4 * i + 4] = [gW, gb, gg, gbeta]
            if i == 0 and ctx.x_grad:
                g_x = dz @ W  # (the input gradient of the first Linear: one small library product)
            above, k_above, wa0, wa1, n0 = dz, Hi, W, None, Hi
        return (None, None, None, None, g_x, *grads, g_wloc, g_bloc, g_wstd, g_bstd)


def encoder_mlp(x: torch.Tensor, linears, bns, heads, p: float, rng_states, negative_slope: float = 0.2):
    r"""``(loc, std)`` of a data encoder (``scglue/models/sc.py:136-178``) in ``h_depth + 1`` launches
    each way: ``linears`` / ``bns`` are the ``Linear`` / training-mode ``BatchNorm1d`` modules of the
    hidden blocks, ``heads = (loc, std_lin)``, ``rng_states[i]`` the Philox state of block ``i``'s
    dropout (as :func:`act_bn_dropout`).  The input takes no gradient."""
    require_cuda(x)
    if x.requires_grad:
        raise ValueError("encoder_mlp does not differentiate its input")
    blocks, params = [], []
    for lin, bn, rng in zip(linears, bns, rng_states):
        if not bn.training or not bn.affine or bn.momentum is None:
            raise ValueError("encoder_mlp implements training-mode affine BatchNorm1d with a fixed momentum")
        track = bn.track_running_stats
        blocks.append({"running_mean": bn.running_mean if track else None,
                       "running_var": bn.running_var if track else None,
                       "num_batches_tracked": bn.num_batches_tracked if track else None,
                       "rng": rng, "eps": float(bn.eps), "momentum": float(bn.momentum)})
        params += [lin.weight, lin.bias, bn.weight, bn.bias]
    params += [heads[0].weight, heads[0].bias, heads[1].weight, heads[1].bias]
    return _MlpFn.apply(torch.is_grad_enabled(), blocks, float(negative_slope), float(p), x, *params)


def discriminator_mlp(x: torch.Tensor, linears, pred, p: float, rng_states, negative_slope: float = 0.2):
    r"""Logits of the modality discriminator (``scglue/models/sc.py:576-624``: ``h_depth`` x [Linear ->
    LeakyReLU -> Dropout], then ``pred``) in ``h_depth + 1`` launches each way; differentiable in ``x``
    (the generator's alignment term back-propagates into the encoders)."""
    require_cuda(x)
    blocks, params = [], []
    for lin, rng in zip(linears, rng_states):
        blocks.append({"rng": rng})
        params += [lin.weight, lin.bias, None, None]
    params += [pred.weight, pred.bias, None, None]
    return _MlpFn.apply(torch.is_grad_enabled(), blocks, float(negative_slope), float(p), x, *params)[0]


def new_dropout_state(device, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    r"""``uint64 [3]`` Philox state for :func:`act_bn_dropout` (stored as int64): a seed drawn from
    torch's (CPU) generator, call counter 0, ticket 0."""
    seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, generator=generator)
    return torch.cat([seed, torch.zeros(2, dtype=torch.int64)]).to(device)


# ------------------------------------------------- edge-partitioned GraphConv (atlas scale)


class PartitionedGraphCSR:
    r"""
    Row-block partition of the forward (target-sorted) CSR over the ranks of the default
    process group: rank ``r`` owns target rows ``[r * rows_per, (r + 1) * rows_per)`` and
    therefore the edges that point into them (BASELINE.json config 5: "the graph's edges are
    additionally partitioned").  Because the partition is by *target* row, the partial
    aggregates of different ranks are disjoint row blocks, so the exchange is an all-gather
    of ``[V / N, K]`` blocks rather than an all-reduce of ``[V, K]`` partial sums.
    The transposed CSR stays whole on every rank for the backward pass.
    """

    def __init__(self, csr: GraphCSR, rank: int, world: int) -> None:
        self.full, self.rank, self.world = csr, rank, world
        self.rows_per = -(-csr.vnum // world)
        r0 = min(csr.vnum, rank * self.rows_per)
        r1 = min(csr.vnum, r0 + self.rows_per)
        indptr, col, val = csr.fwd
        e0, e1 = int(indptr[r0]), int(indptr[r1])  # one host read at construction
        self.row0, self.n_rows = r0, r1 - r0
        self.block = ((indptr[r0:r1 + 1] - e0).contiguous(), col[e0:e1].contiguous(),
                      val[e0:e1].contiguous())


def _spmm_rows(csr_part, inp: torch.Tensor, n_rows: int, out: torch.Tensor) -> torch.Tensor:
    r"""``out[:n_rows] = A_block @ inp`` for a CSR block whose columns index rows of ``inp``."""
    indptr, col, val = csr_part
    if n_rows:
        check(_lib.lib().glue_graphconv_csr(ptr(indptr), ptr(col), ptr(val), ptr(inp), ptr(out),
                                            n_rows, inp.shape[1], current_stream(inp.device)),
              "graphconv_csr")
    return out


class _PartitionedGraphConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp: torch.Tensor, part: PartitionedGraphCSR) -> torch.Tensor:
        import torch.distributed as td
        ctx.part = part
        K = inp.shape[1]
        mine = torch.zeros(part.rows_per, K, dtype=inp.dtype, device=inp.device)
        _spmm_rows(part.block, inp, part.n_rows, mine)
        gathered = torch.empty(part.world * part.rows_per, K, dtype=inp.dtype, device=inp.device)
        td.all_gather_into_tensor(gathered, mine)
        return gathered[:part.full.vnum]

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        # d L_r / d out differs per rank (each rank scores its own cells), so every rank applies
        # the whole transposed graph to its own upstream gradient; the per-rank results meet in
        # the ordinary data-parallel gradient all-reduce -- no extra collective here.
        return _spmm(ctx.part.full.bwd, _f32(grad_out, "grad_out")), None


def graph_conv_partitioned(inp: torch.Tensor, part: PartitionedGraphCSR) -> torch.Tensor:
    r"""GraphConv with the forward aggregation split over ranks by target-row blocks and
    re-assembled with one all-gather; numerically identical to :func:`graph_conv` (each row
    is summed by exactly one rank, in the same edge order)."""
    inp = _f32(inp, "input")
    if inp.dim() != 2 or inp.shape[0] != part.full.vnum:
        raise ValueError("`input` must have shape [n_vertices, n_features]")
    return _PartitionedGraphConvFn.apply(inp, part)


_PART_CACHE: "OrderedDict[int, PartitionedGraphCSR]" = OrderedDict()


def partition_for(csr: GraphCSR, rank: int, world: int) -> PartitionedGraphCSR:
    key = id(csr)
    hit = _PART_CACHE.get(key)
    if hit is None or hit.full is not csr or (hit.rank, hit.world) != (rank, world):
        hit = PartitionedGraphCSR(csr, rank, world)
        _PART_CACHE[key] = hit
        while len(_PART_CACHE) > _CSR_CACHE_SIZE:
            _PART_CACHE.popitem(last=False)
    return hit


# -------------------------------------------------------------- graph decoder


def _graphdec_ws(n_edges: int, vnum: int, device) -> torch.Tensor:
    return _workspace(_lib.lib().glue_graphdec_workspace_bytes(n_edges, vnum), device)


# While this is a list, the loss functions whose gradients are computed in their forward pass with a folded
# coefficient (decoder_nll / decoder_nll_multi / graph_nll with ``upstream``) append their gradient with
# respect to the vertex table ``v`` to it instead of returning it from ``backward``: the caller (the training
# step) adds those pieces up and back-propagates them through the graph branch right away, beside the rest of
# the forward pass, instead of waiting for the objective's backward pass.
V_GRAD_SINK: Optional[list] = None

_UPSTREAM_SCALARS: dict = {}


def _upstream_scalar(value: float, device) -> torch.Tensor:
    """Device scalar holding a loss coefficient (cached: inside a captured step no tensor is built)."""
    key = (float(value), str(device))
    hit = _UPSTREAM_SCALARS.get(key)
    if hit is None:
        if len(_UPSTREAM_SCALARS) > 64:
            _UPSTREAM_SCALARS.clear()
        hit = _UPSTREAM_SCALARS[key] = torch.full((), float(value), dtype=torch.float32, device=device)
    return hit


class _GraphNllFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, eidx, esgn, ewt, grad_enabled, upstream):
        dev = v.device
        n_edges, (vnum, k) = eidx.shape[1], v.shape
        loss = torch.empty((), dtype=torch.float32, device=dev)
        # (grad mode is always off inside Function.forward: the caller's mode comes in as a flag, so that
        # validation under no_grad does not sort the incidence list)
        need_grad = bool(grad_enabled and ctx.needs_input_grad[0])
        coef = torch.empty(n_edges, dtype=torch.float32, device=dev) if need_grad else None
        ws = _graphdec_ws(n_edges, vnum, dev)
        # one launch: the last CTA publishes the loss and the class weights; the vertex-sorted
        # incidence list of the backward pass depends on `eidx` only and is sorted right behind it,
        # beside the rest of the forward pass, instead of on the tail of the backward pass
        check(_lib.lib().glue_graph_nll_fwd(
            ptr(v), ptr(eidx), ptr(esgn), ptr(ewt), n_edges, vnum, k, ptr(loss),
            ptr(coef) if need_grad else None, 1 if need_grad else 0, ptr(ws), ws.numel(),
            current_stream(dev)), "graph_nll_fwd")
        ctx.eager = None
        if need_grad and upstream is not None:
            # the coefficient of this term in the differentiated objective is known: its gradient is
            # computed here, beside the forward pass of everything else, and handed over as it is
            grad_v = torch.empty_like(v)
            check(_lib.lib().glue_graph_nll_bwd(
                ptr(v), ptr(eidx), ptr(esgn), ptr(ewt), ptr(coef), ptr(upstream), n_edges, vnum, k, ptr(grad_v),
                ptr(ws), ws.numel(), current_stream(dev)), "graph_nll_bwd")
            if V_GRAD_SINK is not None:
                V_GRAD_SINK.append(grad_v)
                ctx.eager = False  # (handed out: nothing flows to `v` from here)
            else:
                ctx.eager = grad_v
        elif need_grad:
            ctx.save_for_backward(v, eidx, esgn, ewt, coef, ws)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.eager is False:
            return None, None, None, None, None, None
        if ctx.eager is not None:
            grad_v, ctx.eager = ctx.eager, None
            return grad_v, None, None, None, None, None
        v, eidx, esgn, ewt, coef, ws = ctx.saved_tensors
        dev = v.device
        n_edges, (vnum, k) = eidx.shape[1], v.shape
        grad_v = torch.empty_like(v)
        up = _f32(grad_out, "grad_out")
        check(_lib.lib().glue_graph_nll_bwd(
            ptr(v), ptr(eidx), ptr(esgn), ptr(ewt), ptr(coef), ptr(up), n_edges, vnum, k, ptr(grad_v),
            ptr(ws), ws.numel(), current_stream(dev)), "graph_nll_bwd")
        return grad_v, None, None, None, None, None


class _GraphLogitsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, eidx, esgn, ewt_dummy):
        dev = v.device
        n_edges, (vnum, k) = eidx.shape[1], v.shape
        logits = torch.empty(n_edges, dtype=torch.float32, device=dev)
        ws = _graphdec_ws(n_edges, vnum, dev)
        check(_lib.lib().glue_graphdec_bce_fwd(
            ptr(v), ptr(eidx), ptr(esgn), ptr(ewt_dummy), n_edges, vnum, k, None, ptr(logits), None,
            None, ptr(ws), ws.numel(), current_stream(dev)), "graphdec_bce_fwd")
        ctx.save_for_backward(v, eidx, esgn)
        return logits

    @staticmethod
    def backward(ctx, grad_logits):
        v, eidx, esgn = ctx.saved_tensors
        dev = v.device
        n_edges, (vnum, k) = eidx.shape[1], v.shape
        grad_v = torch.empty_like(v)
        ws = _graphdec_ws(n_edges, vnum, dev)
        coef = _f32(grad_logits, "grad_logits")
        check(_lib.lib().glue_graphdec_bwd(
            ptr(v), ptr(eidx), ptr(esgn), ptr(coef), n_edges, vnum, k, ptr(grad_v), ptr(ws),
            ws.numel(), current_stream(dev)), "graphdec_bwd")
        return grad_v, None, None, None


def _graph_args(v, eidx, esgn, ewt=None):
    require_cuda(v, eidx, esgn, ewt)
    v, eidx, esgn = _f32(v, "v"), _i64(eidx, "eidx"), _f32(esgn, "esgn")
    if eidx.dim() != 2 or eidx.shape[0] != 2 or esgn.shape != (eidx.shape[1],):
        raise ValueError("`eidx` must be [2, n_edges] and `esgn` [n_edges]")
    if eidx.shape[1] == 0:
        raise ValueError("empty edge minibatch")
    if ewt is not None:
        ewt = _f32(ewt, "ewt")
        if ewt.shape != esgn.shape:
            raise ValueError("`ewt` must be [n_edges]")
    return v, eidx, esgn, ewt


def graph_nll(v, eidx, esgn, ewt, upstream: Optional[float] = None) -> torch.Tensor:
    r"""Balanced Bernoulli NLL of the sampled edges -- replaces ``GraphDecoder.forward``
    + ``Bernoulli.log_prob`` + the pos/neg averaging of ``scglue/models/scglue.py:331-338``
    (without its ``.item()`` host sync).

    ``upstream`` (as for the data decoders): the coefficient with which the caller adds this term to the
    objective it differentiates.  The gradient ``upstream * d nll / d v`` is then computed right behind the
    forward kernels and the backward pass hands it over unchanged -- only valid when exactly that objective
    is back-propagated, once (``SCGLUETrainer.train_step``: ``gen_loss``)."""
    args = _graph_args(v, eidx, esgn, ewt)
    up = None if upstream is None else _upstream_scalar(upstream, args[0].device)
    return _GraphNllFn.apply(*args, torch.is_grad_enabled(), up)


def graph_logits(v, eidx, esgn) -> torch.Tensor:
    r"""``esgn * <v[sidx], v[tidx]>`` (``scglue/models/sc.py:57-58``), differentiable in ``v``."""
    v, eidx, esgn, _ = _graph_args(v, eidx, esgn)
    return _GraphLogitsFn.apply(v, eidx, esgn, esgn)


# --------------------------------------------------------------- data decoder


# Diagnostics: pin a decoder implementation for the calls that follow (``glue_decoder_args.impl``:
# low byte 1 / 2 / 3 = CUDA-core / per-row-staged tcgen05 / TMA-staged tcgen05 kernels, bit 8 = without
# the streamlined NB pass B).  0 = the library chooses.  Tests and A/B benches set it.
DECODER_IMPL = 0


class DecoderCall:
    r"""Argument marshalling for one ``glue_decoder_*`` call (keeps tensors alive)."""

    def __init__(self, family: str, u, v, vidx, b, l, x, scale_lin, bias, disp, zi_logits,
                 reduce: str = "nanmean", row_weight=None, grad_scale: float = 0.0, n_terms: int = 0):
        dev = require_cuda(u, v, vidx, b, l, x, scale_lin, bias, disp, zi_logits, row_weight)
        if family not in FAMILY:
            raise ValueError(f"unknown decoder family {family!r}")
        self.family, self.device = family, dev
        self.u, self.v = _f32(u, "u"), _f32(v, "v")
        self.vidx = None if vidx is None else _i64(vidx, "vidx")
        self.b = None if b is None else _i64(b, "b")
        self.l = None if l is None else _f32(l, "l").reshape(-1)
        self.x = None if x is None else _f32(x, "x")
        self.scale_lin, self.bias, self.disp = (
            _f32(scale_lin, "scale_lin"), _f32(bias, "bias"), _f32(disp, "disp"))
        self.zi = None if zi_logits is None else _f32(zi_logits, "zi_logits")
        self.B, self.K = self.u.shape
        self.n_batches, self.F = self.scale_lin.shape
        self.reduce = {"nanmean": 0, "mean": 1}[reduce]
        self.grad_scale = float(grad_scale or 0.0)
        self.row_weight = None if row_weight is None else _f32(row_weight, "row_weight").reshape(-1)
        if self.row_weight is not None and self.row_weight.numel() != self.B:
            raise ValueError("`row_weight` must have one entry per cell")
        if self.v.dim() != 2 or self.v.shape[1] != self.K:
            raise ValueError("`u` and `v` must share the latent dimension")
        if self.K > 64:
            raise ValueError("glue_b200 decoders support latent_dim <= 64")
        if self.vidx is None and self.v.shape[0] != self.F:
            raise ValueError("`v` must have one row per feature")
        if self.vidx is not None and self.vidx.shape != (self.F,):
            raise ValueError("`vidx` must have one entry per feature")
        if self.x is not None and self.x.shape != (self.B, self.F):
            raise ValueError("`x` must have shape [n_cells, n_features]")
        if family in ("NB", "ZINB") and (self.l is None or self.l.numel() != self.B):
            raise ValueError("NB / ZINB decoders need the library size `l` [n_cells, 1]")
        if family in ("ZINB", "ZILN") and self.zi is None:
            raise ValueError("zero-inflated decoders need `zi_logits`")
        for name, t in (("bias", self.bias), ("disp", self.disp), ("zi_logits", self.zi)):
            if t is not None and t.shape != self.scale_lin.shape:
                raise ValueError(f"`{name}` must have shape [n_batches, n_features]")
        # cells sorted by batch index so that every CTA sees each parameter row as one run
        self.row_order = None
        if self.n_batches > 1 and self.b is not None:
            self.row_order = torch.sort(self.b, stable=True).indices.to(torch.int32)
        if n_terms:
            nbytes = _lib.lib().glue_decoder_multi_workspace_bytes(
                FAMILY[family], self.B, self.F, self.K, self.n_batches, n_terms)
            if nbytes == 0:
                raise ValueError("glue_b200: shape outside the fused several-terms decoder path")
        else:
            nbytes = _lib.lib().glue_decoder_workspace_bytes(
                FAMILY[family], self.B, self.F, self.K, self.n_batches)
        self.ws = _workspace(nbytes, dev)

    def args(self, **out) -> DecoderArgs:
        a = DecoderArgs()
        a.family, a.B, a.F, a.K = FAMILY[self.family], self.B, self.F, self.K
        a.n_batches, a.reduce, a.impl = self.n_batches, self.reduce, DECODER_IMPL
        a.grad_scale = self.grad_scale
        a.u, a.v, a.vidx, a.b = ptr(self.u), ptr(self.v), ptr(self.vidx), ptr(self.b)
        a.row_order, a.l, a.x = ptr(self.row_order), ptr(self.l), ptr(self.x)
        a.scale_lin, a.bias, a.disp = ptr(self.scale_lin), ptr(self.bias), ptr(self.disp)
        a.zi_logits = ptr(self.zi)
        a.row_weight = ptr(self.row_weight)
        a.workspace, a.workspace_bytes = ptr(self.ws), self.ws.numel()
        for key, t in out.items():
            setattr(a, key, ptr(t))
        return a

    def grad_buffers(self):
        dev = self.device
        gu = torch.empty_like(self.u)
        # rows of the latent table that are not features of this modality get zero gradient
        gv = torch.zeros_like(self.v) if self.vidx is not None else torch.empty_like(self.v)
        gs, gb, gd = (torch.empty_like(self.scale_lin) for _ in range(3))
        gz = None if self.zi is None else torch.empty_like(self.zi)
        return gu, gv, gs, gb, gd, gz

    def stream(self) -> int:
        return current_stream(self.device)


class _DecoderNllFn(torch.autograd.Function):
    r"""``-log_prob(x).(nan)mean()`` with every gradient produced by the same launch
    sequence (the ``[B, F]`` logits never reach HBM)."""

    @staticmethod
    def forward(ctx, u, v, scale_lin, bias, disp, zi_logits, vidx, b, l, x, family, reduce,
                unit_upstream, row_weight, upstream, grad_enabled):
        call = DecoderCall(family, u, v, vidx, b, l, x, scale_lin, bias, disp, zi_logits, reduce,
                           row_weight, grad_scale=upstream)
        unit_upstream = unit_upstream or bool(upstream)
        loss = torch.empty((), dtype=torch.float32, device=call.device)
        # (grad mode is always off inside Function.forward: the caller's mode comes in as a flag, so
        # that validation / get_losses under no_grad take the loss-only kernels)
        need_grad = grad_enabled and any(t is not None and t.requires_grad
                                         for t in (u, v, scale_lin, bias, disp, zi_logits))
        L = _lib.lib()
        if need_grad:
            gu, gv, gs, gb, gd, gz = call.grad_buffers()
            a = call.args(loss=loss, grad_u=gu, grad_v=gv, grad_scale_lin=gs, grad_bias=gb,
                          grad_disp=gd, grad_zi_logits=gz)
            ev = TIMER.start(f"decoder_fwd_bwd_{family}_F{call.F}_B{call.B}", call.device)
            check(L.glue_decoder_nll_fwd_bwd(ctypes.byref(a), call.stream()), "decoder_nll_fwd_bwd")
            TIMER.stop(ev, call.device)
            if V_GRAD_SINK is not None and unit_upstream and gv is not None:
                V_GRAD_SINK.append(gv)
                gv = None
            ctx.grads = (gu, gv, gs, gb, gd, gz)
        else:
            a = call.args(loss=loss)
            check(L.glue_decoder_nll_fwd(ctypes.byref(a), call.stream()), "decoder_nll_fwd")
            ctx.grads = None
        ctx.unit_upstream = unit_upstream
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.grads is None:
            raise RuntimeError("decoder_nll: no input required grad at forward time")
        grads = ctx.grads
        ctx.grads = None
        if not ctx.unit_upstream:
            live = [g for g in grads if g is not None]
            torch._foreach_mul_(live, grad_out)
        return (*grads, None, None, None, None, None, None, None, None, None, None)


def decoder_nll(family: str, u, v, scale_lin, bias, disp, zi_logits=None, *, b=None, l=None,
                x=None, vidx=None, reduce: str = "nanmean", unit_upstream: bool = False,
                row_weight=None, upstream: float = None):
    r"""Fused data decoder + negative log-likelihood.

    Replaces ``-u2x(u, v, b, l).log_prob(x).nanmean()`` (``scglue/models/scglue.py:342-347``;
    ``reduce="mean"`` is the ``PairedSCGLUETrainer`` variant, ``scglue.py:603``).  With
    ``vidx`` given, ``v`` is the full vertex table and ``v[vidx]`` is gathered inside the
    kernel (saves the ``[F, K]`` gather and its ``index_put_`` backward).
    ``unit_upstream=True`` promises that this loss enters the final objective with
    coefficient exactly 1, so backward hands the gradients over without rescaling;
    ``upstream=c`` promises coefficient exactly ``c`` (a non-zero Python number): the kernels
    fold it into their weights, the returned loss stays unscaled.  **Contract** of both: the
    incoming autograd gradient is not read, so whatever is differentiated must contain this loss
    with exactly that coefficient (``SCGLUETrainer.train_step`` differentiates ``gen_loss`` only;
    anything else -- scaled losses, gradient accumulation with a factor, differentiating a single
    term -- has to leave both arguments unset).
    ``row_weight`` (``[B]``, not differentiated) replaces the mean by
    ``-sum_i row_weight[i] sum_j log_prob[i, j]``: weighted / masked row subsets with static
    shapes (the paired-cell losses of ``scglue.py:603-640`` use ``mask / (n_masked * F)``)."""
    if x is None:
        raise ValueError("`x` is required")
    return _DecoderNllFn.apply(u, v, scale_lin, bias, disp, zi_logits, vidx, b, l, x, family,
                               reduce, unit_upstream, row_weight, upstream, torch.is_grad_enabled())


# ------------------------------------------------- several loss terms of one modality per launch


MAX_TERMS = _lib.MAX_TERMS


class _DecoderNllMultiFn(torch.autograd.Function):
    r"""``T`` loss terms ``-sum_i w_ti sum_j log_prob_t[i, j]`` that share ``x``, ``v``, ``b``, ``l`` and
    the parameter tables, in one launch sequence; the gradients that come back are those of
    ``sum_t upstream[t] * loss[t]`` (folded into the kernels' weights)."""

    @staticmethod
    def forward(ctx, v, scale_lin, bias, disp, zi_logits, vidx, b, l, x, family, reduce, upstreams,
                grad_enabled, n_terms, *us_rws):
        us, rws = us_rws[:n_terms], us_rws[n_terms:]
        call = DecoderCall(family, us[0], v, vidx, b, l, x, scale_lin, bias, disp, zi_logits, reduce,
                           rws[0], n_terms=n_terms)
        dev = call.device
        us = [_f32(u, "u") for u in us]
        rws = [None if w is None else _f32(w, "row_weight").reshape(-1) for w in rws]
        for u, w in zip(us, rws):
            if u.shape != (call.B, call.K):
                raise ValueError("every term's `u` must have shape [n_cells, latent_dim]")
            if w is not None and w.numel() != call.B:
                raise ValueError("`row_weight` must have one entry per cell")
        losses = torch.empty(n_terms, dtype=torch.float32, device=dev)
        need_grad = grad_enabled and any(t is not None and t.requires_grad
                                         for t in (*us, v, scale_lin, bias, disp, zi_logits))
        m = _lib.DecoderMultiArgs()
        L = _lib.lib()
        if need_grad:
            _, gv, gs, gb, gd, gz = call.grad_buffers()
            gus = [torch.empty_like(u) for u in us]
            m.shared = call.args(grad_v=gv, grad_scale_lin=gs, grad_bias=gb, grad_disp=gd,
                                 grad_zi_logits=gz)
        else:
            m.shared = call.args()
        m.n_terms = n_terms
        for t in range(n_terms):
            m.u[t], m.row_weight[t] = ptr(us[t]), ptr(rws[t])
            m.grad_scale[t] = float(upstreams[t])
            m.loss[t] = losses.data_ptr() + 4 * t
            m.grad_u[t] = ptr(gus[t]) if need_grad else None
        ev = TIMER.start(f"decoder_multi{n_terms}_{family}_F{call.F}_B{call.B}", dev)
        fn = L.glue_decoder_nll_multi_fwd_bwd if need_grad else L.glue_decoder_nll_multi_fwd
        check(fn(ctypes.byref(m), call.stream()), "decoder_nll_multi")
        TIMER.stop(ev, dev)
        if need_grad and V_GRAD_SINK is not None and gv is not None:
            V_GRAD_SINK.append(gv)
            gv = None
        ctx.grads = (gv, gs, gb, gd, gz, gus) if need_grad else None
        ctx.n_terms = n_terms
        return losses

    @staticmethod
    def backward(ctx, _grad_losses):
        if ctx.grads is None:
            raise RuntimeError("decoder_nll_multi: no input required grad at forward time")
        gv, gs, gb, gd, gz, gus = ctx.grads
        ctx.grads = None
        return (gv, gs, gb, gd, gz, None, None, None, None, None, None, None, None, None,
                *gus, *([None] * ctx.n_terms))


def decoder_multi_supported(family: str, B: int, F: int, K: int, n_batches: int, n_terms: int,
                            reduce: str, row_weights) -> bool:
    r"""Shapes the fused several-terms path takes (mirror of ``tc3_eligible``, decoder_tc3.cu)."""
    if family not in FAMILY or not 1 <= n_terms <= MAX_TERMS:
        return False
    if B > 128 or K > 64 or K % 2 or n_batches > 24:
        return False
    if n_terms > 1 and reduce == "nanmean" and any(w is None for w in row_weights):
        return False
    return True


def decoder_nll_multi(family: str, us, v, scale_lin, bias, disp, zi_logits=None, *, b=None, l=None,
                      x=None, vidx=None, reduce: str = "mean", row_weights=None, upstreams=None):
    r"""Several decoder loss terms of one modality that differ only in the cell embeddings ``us[t]``,
    the per-cell weights ``row_weights[t]`` (``None`` = the plain ``reduce``) and their coefficient
    ``upstreams[t]`` in the objective -- the reconstruction, joint-cross and real-cross terms of
    ``PairedSCGLUETrainer`` (``scglue/models/scglue.py:603-652``).  Returns the tuple of (unscaled)
    losses.  **Contract**: the objective that is differentiated must contain term ``t`` with
    coefficient exactly ``upstreams[t]`` (the kernels fold it into their weights; the incoming
    autograd gradients are not read).  Up to three terms run as one launch sequence that stages the
    feature table once and reads ``x`` from HBM once; anything the fused path does not take is
    evaluated term by term with :func:`decoder_nll`."""
    n = len(us)
    row_weights = list(row_weights) if row_weights is not None else [None] * n
    if upstreams is None or any(not c for c in upstreams):
        raise ValueError("`upstreams` (non-zero coefficient of every term in the objective) is required")
    if x is None:
        raise ValueError("`x` is required")
    B, K = us[0].shape
    nb, F = scale_lin.shape
    out = []
    for lo in range(0, n, MAX_TERMS):
        grp = slice(lo, min(n, lo + MAX_TERMS))
        g_us, g_rw, g_up = list(us[grp]), row_weights[grp], [float(c) for c in upstreams[grp]]
        if decoder_multi_supported(family, B, F, K, nb, len(g_us), reduce, g_rw):
            losses = _DecoderNllMultiFn.apply(v, scale_lin, bias, disp, zi_logits, vidx, b, l, x, family,
                                              reduce, tuple(g_up), torch.is_grad_enabled(), len(g_us),
                                              *g_us, *g_rw)
            out.extend(losses.unbind(0))
        else:
            out.extend(decoder_nll(family, u, v, scale_lin, bias, disp, zi_logits, b=b, l=l, x=x, vidx=vidx,
                                   reduce=reduce, row_weight=w, upstream=c)
                       for u, w, c in zip(g_us, g_rw, g_up))
    return tuple(out)


class _DecoderLogProbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, v, scale_lin, bias, disp, zi_logits, vidx, b, l, x, family):
        call = DecoderCall(family, u, v, vidx, b, l, x, scale_lin, bias, disp, zi_logits)
        out = torch.empty(call.B, call.F, dtype=torch.float32, device=call.device)
        a = call.args()
        check(_lib.lib().glue_decoder_log_prob(ctypes.byref(a), ptr(out), call.stream()),
              "decoder_log_prob")
        ctx.call = call
        return out

    @staticmethod
    def backward(ctx, grad_out):
        call = ctx.call
        wmat = _f32(grad_out, "grad_out")
        gu, gv, gs, gb, gd, gz = call.grad_buffers()
        a = call.args(wmat=wmat, grad_u=gu, grad_v=gv, grad_scale_lin=gs, grad_bias=gb,
                      grad_disp=gd, grad_zi_logits=gz)
        check(_lib.lib().glue_decoder_nll_fwd_bwd(ctypes.byref(a), call.stream()),
              "decoder_nll_fwd_bwd")
        return gu, gv, gs, gb, gd, gz, None, None, None, None, None


def decoder_log_prob(family: str, u, v, scale_lin, bias, disp, zi_logits=None, *, b=None, l=None,
                     x=None, vidx=None) -> torch.Tensor:
    r"""Materialised ``u2x(u, v, b, l).log_prob(x)`` ``[B, F]`` for API users; its backward
    runs the fused kernels with the incoming gradient as per-element weights."""
    return _DecoderLogProbFn.apply(u, v, scale_lin, bias, disp, zi_logits, vidx, b, l, x, family)


@torch.no_grad()
def decoder_mean(family: str, u, v, scale_lin, bias, disp, zi_logits=None, *, b=None, l=None,
                 vidx=None) -> torch.Tensor:
    r"""Mean of the decoded distribution ``[B, F]`` (``decode_data``, ``scglue.py:1355``)."""
    call = DecoderCall(family, u, v, vidx, b, l, None, scale_lin, bias, disp, zi_logits)
    out = torch.empty(call.B, call.F, dtype=torch.float32, device=call.device)
    a = call.args()
    check(_lib.lib().glue_decoder_mean(ctypes.byref(a), ptr(out), call.stream()), "decoder_mean")
    return out
