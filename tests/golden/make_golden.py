"""Generates tests/golden/*.npz.

The reference (colinvaz/nmf-toolbox) is MATLAB and ships no golden vectors, and no
MATLAB/Octave exists in the build image, so these fixtures are outputs of the
literal NumPy restatement (oracle/nmf_oracle.py) at the commit that created
them.  They pin the oracle against silent drift and give the GPU tests fixed
targets that do not depend on re-running the oracle.  Inputs are regenerated
from the recorded seeds (numpy.random.default_rng, PCG64) - only the outputs are
stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import nmf_oracle as O  # noqa: E402

CASES = {
    # name: (algorithm, m, n, K, T, iters, config extras)
    "nmf_euclid_512": ("nmf", 512, 512, 16, 1, 50, dict(divergence="euclidean")),  # BASELINE.json configs[0]
    "nmf_euclid_sparse": ("nmf", 200, 333, 24, 1, 40, dict(divergence="euclidean", W_sparsity=0.1, H_sparsity=0.2)),
    "nmf_kl_300": ("nmf", 300, 257, 20, 1, 40, dict(divergence="kl")),
    "nmf_kl_sparse": ("nmf", 129, 400, 8, 1, 30, dict(divergence="kl_divergence", W_sparsity=0.05, H_sparsity=0.1)),
    "nmf_is_200": ("nmf", 200, 260, 12, 1, 40, dict(divergence="is")),
    "nmf_ab_half_half": ("nmf", 257, 300, 10, 1, 40, dict(divergence="ab", alpha=0.5, beta=0.5)),
    "nmf_ab_2_1_sparse": ("nmf", 150, 400, 8, 1, 30, dict(divergence="ab_divergence", alpha=2, beta=1, W_sparsity=0.05,
                                                           H_sparsity=0.1)),
    "cnmf_euclid": ("cnmf", 129, 700, 8, 4, 40, dict(divergence="euclidean")),
    "cnmf_frobenius_sparse": ("cnmf", 100, 300, 6, 3, 20, dict(divergence="frobenius", W_sparsity=0.05, H_sparsity=0.1)),
    "cnmf_kl": ("cnmf", 120, 400, 6, 4, 30, dict(divergence="kl")),
    "cnmf_ab_half_1": ("cnmf", 100, 350, 5, 3, 25, dict(divergence="ab", alpha=0.5, beta=1.0, H_sparsity=0.05)),
    "lnmf_200": ("lnmf", 200, 300, 12, 1, 40, dict()),
    "cnmfsc_h06": ("cnmfsc", 129, 500, 6, 4, 30, dict(H_sparsity=0.6)),
    "cnmfsc_plain": ("cnmfsc", 100, 360, 5, 3, 30, dict()),
    "nmfsc_h07": ("nmfsc", 512, 512, 16, 1, 60, dict(H_sparsity=0.7)),
    "nmfsc_plain": ("nmfsc", 200, 300, 8, 1, 40, dict()),
}


def inputs(name):
    alg, m, n, K, T, iters, extra = CASES[name]
    seed = sum(map(ord, name))
    rng = np.random.default_rng(seed)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    if alg == "cnmf":
        W0 = rng.random((m, K, T))
    elif alg == "cnmfsc":
        W0 = 0.3 * rng.random((m, K, T))
    elif alg == "nmfsc":
        W0 = rng.random((m, K))
    else:
        W0 = np.maximum(rng.random((m, K)), O.EPS)
    H0 = np.maximum(rng.random((K, n)), O.EPS)
    if alg in ("nmfsc", "cnmfsc"):
        V = V * 3.0
        H0 = H0 / np.sqrt((H0 ** 2).sum(1, keepdims=True))
    cfg = dict(extra, W_init=W0, H_init=H0, maxiter=iters, tolerance=1e-300)
    return alg, V, K, T, cfg


def run(name):
    alg, V, K, T, cfg = inputs(name)
    if alg == "nmf":
        return O.nmf(V, K, cfg)
    if alg == "lnmf":
        return O.lnmf(V, K, cfg)
    if alg == "cnmfsc":
        return O.cnmfsc(V, K, T, cfg)
    if alg == "cnmf":
        return O.cnmf(V, K, T, cfg)
    return O.nmfsc(V, K, cfg)


if __name__ == "__main__":
    only = sys.argv[1:]  # optional: regenerate just these cases
    for name in CASES:
        if only and name not in only:
            continue
        W, H, cost = run(name)
        Vhat = O.reconstruct_from_decomposition(W, H)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), cost=cost, W=W.astype(np.float32), H=H.astype(np.float32),
                            vhat_norm=np.linalg.norm(Vhat), vhat_sample=Vhat[:8, :8])
        print(name, cost[0], cost[-1], len(cost))
    # projfunc known answers
    rng = np.random.default_rng(99)
    S = rng.random((6, 500))
    outs, its = [], []
    for i, sp in enumerate([0.1, 0.3, 0.5, 0.7, 0.9, 0.95]):
        k1 = np.sqrt(500) - (np.sqrt(500) - 1) * sp
        v, it = O.projfunc(S[i], k1, 1.0, 1)
        outs.append(v)
        its.append(it)
    np.savez_compressed(os.path.join(HERE, "projfunc.npz"), v=np.array(outs), iters=np.array(its))
    print("projfunc iters", its)
