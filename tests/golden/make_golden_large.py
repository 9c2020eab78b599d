"""Generates tests/golden/large_*.npz: oracle cost curves at BASELINE.json's operating points.

The float64 oracle (oracle/nmf_oracle.py, the literal restatement of nmf.m / cnmf.m / nmfsc.m)
needs seconds per iteration at these sizes (16384 x 16384, K = 256: ~4 s per iteration on 8
cores), so its outputs are computed once here and committed; the GPU tests regenerate the
*inputs* from the recorded seeds and compare

  * the whole cost curve (every iteration), and
  * the reconstruction W*H on a window of rows x columns, from the stored rows of W and the
    stored columns of H of the oracle's final factors.

Inputs are float32 values (so the device sees exactly the numbers the oracle saw), generated
with numpy.random.default_rng(seed) (PCG64: platform independent).

    python tests/golden/make_golden_large.py [case ...]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

LARGE = {
    # name: (algorithm, m, n, K, T, iterations, config extras)   -- BASELINE.json configs[1..4]
    "large_nmf_euclid_16384_k256": ("nmf", 16384, 16384, 256, 1, 200, dict(divergence="euclidean")),
    "large_nmf_euclid_4096_k256": ("nmf", 4096, 4096, 256, 1, 200, dict(divergence="euclidean")),
    "large_nmf_kl_8192_k128": ("nmf", 8192, 8192, 128, 1, 50, dict(divergence="kl")),
    "large_cnmf_1025x20000_k64_t8": ("cnmf", 1025, 20000, 64, 8, 30, dict(divergence="euclidean")),
    "large_nmfsc_4096_k128_h07": ("nmfsc", 4096, 4096, 128, 1, 30, dict(H_sparsity=0.7)),
    # the widened divergences at the shape their fused kernel (ab_fused.cuh) is timed on
    "large_nmf_is_8192_k128": ("nmf", 8192, 8192, 128, 1, 30, dict(divergence="is")),
    "large_nmf_ab_8192_k128": ("nmf", 8192, 8192, 128, 1, 20, dict(divergence="ab", alpha=0.5, beta=0.5, H_sparsity=0.05)),
}
WIN_ROWS = 256
WIN_COLS = 256


def large_inputs(name):
    """(alg, Vt, K, T, cfg): Vt is [n][m] float32 = V in column-major order (V = Vt.T)."""
    alg, m, n, K, T, iters, extra = LARGE[name]
    rng = np.random.default_rng(1000 + sum(map(ord, name)))
    Vt = rng.random((n, m), dtype=np.float32)
    np.maximum(Vt, np.float32(2.0 ** -24), out=Vt)
    if alg == "cnmf":
        W0 = rng.random((m, K, T), dtype=np.float32) + np.float32(1e-3)
    else:
        W0 = rng.random((m, K), dtype=np.float32) + np.float32(1e-3)
    H0 = rng.random((K, n), dtype=np.float32) + np.float32(1e-3)
    if alg == "nmfsc":  # nmfsc.m:79-80: unit-L2 rows (the reference's default initialisation)
        H0 = (H0 / np.sqrt((H0.astype(np.float64) ** 2).sum(1, keepdims=True))).astype(np.float32)
    cfg = dict(extra, W_init=W0, H_init=H0, maxiter=iters, tolerance=1e-300)
    return alg, Vt, K, T, cfg


def window(m, n, T):
    """Rows (strided) and a contiguous column window [c0, c1) used for the reconstruction check."""
    rows = np.unique(np.linspace(0, m - 1, WIN_ROWS).astype(np.int64))
    c0 = max(T - 1, n // 2 - WIN_COLS // 2)
    return rows, c0, min(n, c0 + WIN_COLS)


def window_recon(W_rows, H_win, T):
    """V_hat[rows, c0:c1] from W[rows] (R x K or R x K x T) and H[:, c0-(T-1):c1]
    (ReconstructFromDecomposition.m:31,36-38)."""
    W_rows = np.asarray(W_rows, np.float64)
    H_win = np.asarray(H_win, np.float64)
    if W_rows.ndim == 2:
        return W_rows @ H_win
    ncol = H_win.shape[1] - (T - 1)
    out = np.zeros((W_rows.shape[0], ncol))
    for t in range(T):  # column j of the window reads H(:, j - t)
        out += W_rows[:, :, t] @ H_win[:, T - 1 - t: T - 1 - t + ncol]
    return out


def run(name):
    from oracle import nmf_oracle as O

    alg, Vt, K, T, cfg = large_inputs(name)
    V = Vt.T.astype(np.float64)
    cfg = dict(cfg, W_init=cfg["W_init"].astype(np.float64), H_init=cfg["H_init"].astype(np.float64))
    if alg == "nmf":
        W, H, cost = O.nmf(V, K, cfg)
    elif alg == "cnmf":
        W, H, cost = O.cnmf(V, K, T, cfg)
    else:
        info = {}
        W, H, cost = O.nmfsc(V, K, cfg, info=info)
        extra = dict(halvings_H=np.asarray(info["halvings_H"], np.int32), halvings_W=np.asarray(info["halvings_W"], np.int32))
        return W, H, cost, extra
    return W, H, cost, {}


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, (alg, m, n, K, T, iters, extra) in LARGE.items():
        if only and name not in only:
            continue
        t0 = time.time()
        W, H, cost, extra = run(name)
        rows, c0, c1 = window(m, n, T)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), cost=cost, W_rows=W[rows].astype(np.float32),
                            H_win=H[:, c0 - (T - 1): c1].astype(np.float32), rows=rows, c0=c0, c1=c1,
                            seconds=time.time() - t0, **extra)
        print(name, len(cost), cost[0], cost[-1], "%.0f s" % (time.time() - t0), flush=True)
