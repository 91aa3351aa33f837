"""GPU tests of the tcgen05 plumbing behind the fused decoder (v2):

* ``glue_tc_selftest`` pins the UMMA operand conventions (K-major / MN-major descriptors over the
  128-byte-swizzled [rows][64 bf16] tiles) against a plain matrix product;
* the tensor-core decoder is compared with the CUDA-core decoder (``GLUE_B200_DECODER=v1``) on the
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


# ------------------------------------------------------------------ v2 against v1
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
def test_tensor_core_decoder_matches_cuda_core_decoder(native_lib, family, B, F, K, nb):
    from glue_b200 import ops
    from tests.test_ops_gpu import PARAM_NAMES, _decoder_case
    case = _decoder_case(family, B, F, K, nb, seed=B + F)
    os.environ["GLUE_B200_DECODER"] = "v1"
    try:
        l1, g1 = _run_decoder(ops, family, case, "nanmean")
    finally:
        os.environ.pop("GLUE_B200_DECODER")
    l2, g2 = _run_decoder(ops, family, case, "nanmean")
    assert_close(l2, l1, 1e-4, "loss")
    for n, a, b in zip(["u", "v"] + list(PARAM_NAMES[family]), g2, g1):
        assert_close(a, b, 1e-4, f"{family} d/d{n} (v2 vs v1)")
    l3, g3 = _run_decoder(ops, family, case, "nanmean")
    assert torch.equal(l2, l3) and all(torch.equal(a, b) for a, b in zip(g2, g3)), "v2 must be deterministic"
