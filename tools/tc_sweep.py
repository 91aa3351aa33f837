"""Exploratory sweep of UMMA descriptor hypotheses through glue_tc_selftest (run on a B200).
Prints the max-norm relative error of every variant against the intended matrix product."""
import itertools
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glue_b200 import _lib
from tests.test_tc_gpu import (_rand_bf16, _x_image, bf16_bits, desc_base, idesc, run_selftest, swz_tile)
from tests.util import rel_err

lib = _lib.lib()
A, B = _rand_bf16((128, 64), 1), _rand_bf16((128, 64), 2)
for lbo, sbo in [(16, 1024), (0, 1024), (1024, 1024)]:
    out = run_selftest(lib, swz_tile(bf16_bits(A)), swz_tile(bf16_bits(B)), desc_base(lbo, sbo), desc_base(lbo, sbo),
                       idesc(128, 128, 0, 0), [32 * k for k in range(4)], [32 * k for k in range(4)], 128)
    print(f"fwd  K/K  lbo={lbo} sbo={sbo}: err={rel_err(out, A @ B.T):.3e}", flush=True)

X, U = _rand_bf16((128, 128), 3), _rand_bf16((128, 64), 4)
a_off = [(k >> 2) * 16384 + (k & 3) * 32 for k in range(8)]
for (lbo, sbo), step in itertools.product([(1024, 1024), (0, 1024), (1024, 0), (16, 1024), (128, 1024), (1024, 128)],
                                          [2048]):
    out = run_selftest(lib, _x_image(X), swz_tile(bf16_bits(U)), desc_base(16, 1024), desc_base(lbo, sbo),
                       idesc(128, 64, 0, 1), a_off, [step * k for k in range(8)], 64)
    print(f"D1   K/MN b: lbo={lbo} sbo={sbo} step={step}: err={rel_err(out, X @ U):.3e}", flush=True)

V = _rand_bf16((128, 64), 6)
for (lbo, sbo) in [(16384, 1024), (1024, 16384), (16384, 128), (128, 16384)]:
    off = [2048 * k for k in range(8)]
    out = run_selftest(lib, _x_image(X), swz_tile(bf16_bits(V)), desc_base(lbo, sbo), desc_base(1024, 1024),
                       idesc(128, 64, 1, 1), off, off, 64)
    print(f"D2   MN/MN a: lbo={lbo} sbo={sbo}: err={rel_err(out, X.T @ V):.3e}", flush=True)
