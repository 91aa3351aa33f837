"""GPU tests of the tcgen05 plumbing behind the fused decoder (v2):

* ``glue_tc_selftest`` pins the UMMA operand conventions (K-major / MN-major descriptors over the
  128-byte-swizzled [rows][64 bf16] tiles) against a plain matrix product;
* the tensor-core decoder is compared with the CUDA-core decoder (``impl`` 1) on the
  same inputs -- both are also compared with the oracle in tests/test_ops_gpu.py.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from tests.util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def bf16_bits(x: torch.Tensor) -> np.ndarray:
    return x.to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)


def swz_tile(bits: np.ndarray) -> np.ndarray:
    """[rows, 64] bf16 bit patterns -> physical [rows, 64] with the 16-byte chunks of every row
    XOR-swizzled by (row & 7) (UMMA SWIZZLE_128B)."""
    rows = bits.shape[0]
    assert bits.shape == (rows, 64)
    out = np.zeros_like(bits)
    for r in range(rows):
        for ch in range(8):
            p = ch ^ (r & 7)
            out[r, p * 8:(p + 1) * 8] = bits[r, ch * 8:(ch + 1) * 8]
    return out


def desc_base(lbo, sbo):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | 2 << 61


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run_selftest(lib, a_img, b_img, a_desc, b_desc, idsc, a_off, b_off, N):
    a = torch.from_numpy(a_img.view(np.uint8).reshape(-1).copy()).to(DEV)
    b = torch.from_numpy(b_img.view(np.uint8).reshape(-1).copy()).to(DEV)
    out = torch.zeros(128, N, device=DEV)
    nk = len(a_off)
    ao = (ctypes.c_uint32 * nk)(*a_off)
    bo = (ctypes.c_uint32 * nk)(*b_off)
    rc = lib.glue_tc_selftest(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), a_desc, b_desc, idsc, nk,
                              ao, bo, N, out.data_ptr(), None)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return out.cpu()


def _rand_bf16(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(torch.bfloat16).float()


def test_umma_k_major_logits_tile(native_lib):
    A, B = _rand_bf16((128, 64), 1), _rand_bf16((128, 64), 2)
    out = run_selftest(native_lib, swz_tile(bf16_bits(A)), swz_tile(bf16_bits(B)), desc_base(16, 1024),
                       desc_base(16, 1024), idesc(128, 128, 0, 0), [32 * k for k in range(4)],
                       [32 * k for k in range(4)], 128)
    assert_close(out, A @ B.T, 1e-5, "K-major x K-major")


def _x_image(X):
    bits = bf16_bits(X)  # [128 features, 128 cells]
    return np.concatenate([swz_tile(bits[:, :64]), swz_tile(bits[:, 64:])], axis=0)


def test_umma_d1_k_major_times_mn_major(native_lib):
    X, U = _rand_bf16((128, 128), 3), _rand_bf16((128, 64), 4)  # X[feature, cell], U[cell, k]
    a_off = [(k >> 2) * 16384 + (k & 3) * 32 for k in range(8)]
    b_off = [2048 * k for k in range(8)]
    out = run_selftest(native_lib, _x_image(X), swz_tile(bf16_bits(U)), desc_base(16, 1024),
                       desc_base(1024, 1024), idesc(128, 64, 0, 1), a_off, b_off, 64)
    assert_close(out, X @ U, 1e-5, "d/dv tile")


def test_umma_d2_mn_major_times_mn_major(native_lib):
    X, V = _rand_bf16((128, 128), 5), _rand_bf16((128, 64), 6)  # X[feature, cell], V[feature, k]
    off = [2048 * k for k in range(8)]
    out = run_selftest(native_lib, _x_image(X), swz_tile(bf16_bits(V)), desc_base(16384, 1024),
                       desc_base(1024, 1024), idesc(128, 64, 1, 1), off, off, 64)
    assert_close(out, X.T @ V, 1e-5, "d/du partial")


# ------------------------------------------------------------------ v2 / v3 against v1
import contextlib


@contextlib.contextmanager
def _impl(ops, impl):
    """Pin a decoder implementation (``glue_decoder_args.impl``) for the calls inside the block."""
    old = ops.DECODER_IMPL
    ops.DECODER_IMPL = impl
    try:
        yield
    finally:
        ops.DECODER_IMPL = old


def _run_decoder(ops, family, case, reduce):
    from tests.test_ops_gpu import _device_nll
    return _device_nll(ops, family, *case, reduce)


@pytest.mark.parametrize("family,B,F,K,nb", [
    ("NB", 128, 2000, 50, 1), ("NB", 128, 50000, 50, 1), ("NB", 77, 1300, 50, 1), ("NB", 20, 140, 8, 1),
    ("ZINB", 128, 3000, 50, 20), ("NB", 100, 5000, 50, 5), ("Normal", 128, 4000, 50, 3),
    ("ZILN", 96, 2500, 50, 1), ("NB", 128, 150000, 50, 1),
    # more than 128 cells: cell blocks of 128, feature-side outputs accumulate over the blocks
    ("NB", 512, 4000, 50, 1), ("NB", 300, 1000, 50, 3), ("ZINB", 257, 700, 50, 1), ("Normal", 130, 900, 16, 1),
])
@pytest.mark.parametrize("ver", [2, 3])
def test_tensor_core_decoder_matches_cuda_core_decoder(native_lib, family, B, F, K, nb, ver):
    """``ver``: 2 = per-row-staged tcgen05 kernels, 3 = TMA-staged ones (<= 128 cells; larger calls
    fall through to 2), each against the CUDA-core kernels (``impl`` 1)."""
    from glue_b200 import ops
    from tests.test_ops_gpu import PARAM_NAMES, _decoder_case
    case = _decoder_case(family, B, F, K, nb, seed=B + F)
    with _impl(ops, 1):
        l1, g1 = _run_decoder(ops, family, case, "nanmean")
    with _impl(ops, ver):
        l2, g2 = _run_decoder(ops, family, case, "nanmean")
        l3, g3 = _run_decoder(ops, family, case, "nanmean")
    assert_close(l2, l1, 1e-4, "loss")
    for n, a, b in zip(["u", "v"] + list(PARAM_NAMES[family]), g2, g1):
        assert_close(a, b, 1e-4, f"{family} d/d{n} (v{ver} vs v1)")
    assert torch.equal(l2, l3) and all(torch.equal(a, b) for a, b in zip(g2, g3)), "must be deterministic"


# ------------------------------------------------------------------ kernel variants of the main pass
# variant 1 = streamlined NB pass B (32-bit unpredicated x loads, quadratic gamma terms for the counts
# 0 / 1 / 2, c_i from the tensor cores; the default for NB with one parameter row).  It must agree
# with the base kernel (variant 0 = ``impl`` bit 8), which the tests above / test_ops_gpu.py tie to the
# CUDA-core decoder and to the oracle.  Both tcgen05 implementations (``ver`` 2 / 3) have the pair.
def _variant_case(B, F, K, seed, sparse, nan_frac, extra_rows):
    from tests.test_ops_gpu import _decoder_case
    u, v, b, l, x, params, vidx = _decoder_case("NB", B, F, K, 1, seed=seed, extra_rows=extra_rows)
    gen = torch.Generator().manual_seed(seed + 1)
    if sparse:  # ATAC-like: mostly 0, some 1 / 2, a few larger counts
        x = torch.poisson(torch.full((B, F), 0.08), generator=gen)
        x[torch.rand(B, F, generator=gen) < 0.002] = 5.0
        l = x.sum(dim=1, keepdim=True) + 1.0
    if nan_frac:
        x[torch.rand(B, F, generator=gen) < nan_frac] = float("nan")
    return u, v, b, l, x, params, vidx


def _run_variant(ops, variant, case, reduce, row_weight=None, ver=2):
    u, v, b, l, x, params, vidx = case
    d = lambda t: None if t is None else t.to(DEV)
    du, dv = d(u).requires_grad_(True), d(v).requires_grad_(True)
    dp = [d(params[n]).requires_grad_(True) for n in ("scale_lin", "bias", "log_theta")]
    with _impl(ops, ver | (0 if variant else 0x100)):
        loss = ops.decoder_nll("NB", du, dv, dp[0], dp[1], dp[2], None, b=d(b), l=d(l), x=d(x), vidx=d(vidx),
                               reduce=reduce, row_weight=d(row_weight))
        grads = torch.autograd.grad(loss, [du, dv] + dp)
        torch.cuda.synchronize()
    return loss, grads


VARIANT_CASES = [
    # B, F, K, sparse, nan_frac, extra_rows, reduce, weighted
    (128, 2000, 50, False, 0.0, 0, "nanmean", False),
    (128, 50000, 50, True, 0.0, 0, "nanmean", False),
    (128, 6000, 50, True, 0.01, 0, "nanmean", False),
    (128, 3000, 50, False, 0.01, 0, "nanmean", False),
    (77, 1300, 50, True, 0.0, 0, "mean", False),
    (100, 5000, 50, True, 0.0, 300, "nanmean", False),
    (20, 140, 8, False, 0.0, 0, "nanmean", False),
    (5, 130, 62, True, 0.0, 0, "nanmean", False),
    (64, 400, 64, True, 0.0, 0, "nanmean", False),     # K = 64: no spare operand column, bit 0 is ignored
    (1, 5, 50, False, 0.0, 0, "nanmean", False),
    (512, 4000, 50, True, 0.0, 0, "nanmean", False),   # four cell blocks
    (300, 1000, 50, False, 0.0, 40, "mean", False),
    (128, 2500, 50, True, 0.0, 0, "mean", True),       # per-cell weights (paired losses)
    (200, 1500, 50, True, 0.01, 0, "mean", True),
]


@pytest.mark.parametrize("ver", [2, 3])
@pytest.mark.parametrize("variant", [1])
@pytest.mark.parametrize("B,F,K,sparse,nan_frac,extra,reduce,weighted", VARIANT_CASES)
def test_main_pass_variants_match_base_kernel(native_lib, variant, B, F, K, sparse, nan_frac, extra, reduce, weighted,
                                              ver):
    from glue_b200 import ops
    case = _variant_case(B, F, K, seed=7 * B + F, sparse=sparse, nan_frac=nan_frac, extra_rows=extra)
    if nan_frac and reduce == "mean":
        case[4].nan_to_num_(nan=3.0)  # (NaN only has a meaning under nanmean)
    rw = None
    if weighted:
        rw = torch.rand(B, generator=torch.Generator().manual_seed(B)) * (torch.arange(B) % 3 != 0)
        if nan_frac:
            case[4].nan_to_num_(nan=1.0)
    l0, g0 = _run_variant(ops, 0, case, reduce, rw, ver)
    l1, g1 = _run_variant(ops, variant, case, reduce, rw, ver)
    assert torch.isfinite(l0).all() and torch.isfinite(l1).all()
    assert_close(l1, l0, 2e-5, f"loss (variant {variant})")
    for n, a, b in zip(["u", "v", "scale_lin", "bias", "log_theta"], g1, g0):
        assert_close(a, b, 1e-4, f"d/d{n} (variant {variant} vs base)")
    l2, g2 = _run_variant(ops, variant, case, reduce, rw, ver)
    assert torch.equal(l1, l2) and all(torch.equal(a, b) for a, b in zip(g1, g2)), "variants must be deterministic"


@pytest.mark.parametrize("ver", [2, 3])
@pytest.mark.parametrize("variant", [0, 1])
def test_main_pass_variants_vs_oracle(native_lib, variant, ver):
    from glue_b200 import ops
    from tests.test_ops_gpu import TOL, _oracle_nll
    case = _variant_case(128, 1500, 50, seed=11, sparse=True, nan_frac=0.005, extra_rows=0)
    ref_loss, _, ref_grads = _oracle_nll("NB", *case, "nanmean")
    loss, grads = _run_variant(ops, variant, case, "nanmean", ver=ver)
    assert_close(loss, ref_loss, TOL, "nll vs oracle")
    for n, g, rg in zip(["u", "v", "scale_lin", "bias", "log_theta"], grads, ref_grads):
        assert_close(g, rg, TOL, f"d/d{n} vs oracle (variant {variant})")
