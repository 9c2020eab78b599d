"""GPU parity at BASELINE.json's operating points: the CUDA path (through the C ABI) against the float64
oracle's committed outputs (tests/golden/large_*.npz, generator tests/golden/make_golden_large.py):

  large_nmf_euclid_16384_k256    configs[1]: nmf.m euclidean, 16384 x 16384, K = 256, 200 iterations
  large_nmf_euclid_4096_k256     the same K and iteration count at a quarter of the side
  large_nmf_kl_8192_k128         configs[2], one GPU's column shard: nmf.m KL, 8192 x 8192, K = 128, 50 iterations
  large_cnmf_1025x20000_k64_t8   configs[3]: cnmf.m euclidean, 1025 x 20000, K = 64, T = 8, 30 iterations
  large_nmfsc_4096_k128_h07      configs[4]: nmfsc.m, 4096 x 4096, K = 128, H_sparsity = 0.7, 30 iterations
  large_nmf_is_8192_k128         nmf.m Itakura-Saito, 8192 x 8192, K = 128, 30 iterations (ab_fused.cuh)
  large_nmf_ab_8192_k128         nmf.m alpha-beta (0.5, 0.5) with H sparsity, same shape, 20 iterations

Stated tolerances (BASELINE.json north_star): cost within 1e-4 relative of the reference at EVERY
iteration, W*H within 1e-3 relative (Frobenius norm, on a 256-row x 256-column window of V_hat whose
oracle factors are stored)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_large import LARGE, large_inputs, window_recon  # noqa: E402

pytestmark = pytest.mark.gpu

COST_TOL = 1e-4
RECON_TOL = 1e-3


@pytest.mark.parametrize("name", sorted(LARGE))
def test_large_parity_vs_oracle_golden(name):
    from nmf_toolbox_b200 import api

    path = os.path.join(HERE, "golden", name + ".npz")
    assert os.path.exists(path), "golden fixture missing: run tests/golden/make_golden_large.py " + name
    g = np.load(path)
    alg, Vt, K, T, cfg = large_inputs(name)
    h = api.Handle(0)
    try:
        h.set_V(Vt.T)  # column-major m x n view of the [n][m] array: no copy on the host
        if alg == "nmf":
            W, H, c = h.nmf(K, cfg)
        elif alg == "cnmf":
            W, H, c = h.cnmf(K, T, cfg)
        else:
            W, H, c = h.nmfsc(K, cfg)
            halv_H, halv_W = h.last_halvings()
    finally:
        h.close()
    co = g["cost"]
    assert len(c) == len(co), (len(c), len(co))
    rel = np.abs(c - co) / np.abs(co)
    assert float(rel.max()) < COST_TOL, (int(rel.argmax()), float(rel.max()))
    rows, c0, c1 = g["rows"], int(g["c0"]), int(g["c1"])
    R = window_recon(W[rows], H[:, c0 - (T - 1): c1], T)
    Ro = window_recon(g["W_rows"], g["H_win"], T)
    err = float(np.linalg.norm(R - Ro) / np.linalg.norm(Ro))
    assert err < RECON_TOL, err
    if alg == "nmf":  # nmf.m:169
        np.testing.assert_allclose((W.astype(np.float64) ** 2).sum(0), 1.0, rtol=1e-5)
    if alg == "nmfsc" and "halvings_H" in g.files:  # the device-side line search halves exactly where the reference does
        assert halv_H.tolist() == g["halvings_H"].tolist(), (halv_H.tolist(), g["halvings_H"].tolist())
        assert not halv_W.any()
    if alg == "nmfsc":  # rows of H keep the requested sparseness and unit L2 norm (nmfsc.m:106-109,154-157)
        n = H.shape[1]
        Hd = H.astype(np.float64)
        l1, l2 = np.abs(Hd).sum(1), np.sqrt((Hd ** 2).sum(1))
        np.testing.assert_allclose((np.sqrt(n) - l1 / l2) / (np.sqrt(n) - 1), 0.7, atol=1e-4)
        np.testing.assert_allclose(l2, 1.0, atol=1e-4)
