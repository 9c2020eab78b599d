"""CPU ORACLE (second form) - TEST INFRASTRUCTURE ONLY.  Not part of the product.

The same iterations as ``oracle/nmf_oracle.py`` re-derived in the Gram / trace /
stacked form that the CUDA path implements, still in
NumPy float64.  It exists to pin the literal restatement (both must agree to
~1e-12) and to validate, on the CPU, the algebra the kernels rely on:

* ``diag(H*X'*W)_k = sum_i W_ik (X*H')_ik``  (no m x n temporaries, nmf.m:149-153)
* Euclidean cost by the trace trick ``0.5(|V|^2 - 2<N,H> + <W'W, HH'>)`` (nmf.m:208)
* CNMF as ONE nmf-style update on the stacked (Wc, Hs) pair + a fold (cnmf.m:187-231)
* column sharding over ``shards`` virtual ranks with one packed reduction per
  iteration - only summation order differs from 1 rank.

Only ``tests/`` may import this module.
"""
from __future__ import annotations

import numpy as np

EPS = 2.0 ** -52


def _col_blocks(n, shards):
    edges = [(n * r) // shards for r in range(shards + 1)]
    return [(edges[r], edges[r + 1]) for r in range(shards)]


def nmf_gram(V, K, config, shards=1):
    """nmf.m (Euclidean / KL, single source) in Gram form.  Returns W, H, cost."""
    V = np.asarray(V, dtype=np.float64)
    m, n = V.shape
    div = config.get("divergence", "euclidean")
    lamW = max(float(config.get("W_sparsity", 0) or 0), 0.0)
    lamH = max(float(config.get("H_sparsity", 0) or 0), 0.0)
    W_fixed = bool(config.get("W_fixed", False))
    H_fixed = bool(config.get("H_fixed", False))
    maxiter = int(config.get("maxiter", 100) or 100)
    if maxiter <= 0:
        maxiter = 100
    tol = config.get("tolerance", 1e-3)
    if tol is None or tol <= 0:
        tol = 1e-3
    W = np.array(config["W_init"], dtype=np.float64)
    H = np.array(config["H_init"], dtype=np.float64)
    W = W / np.sqrt(np.sum(W ** 2, axis=0))  # nmf.m:133
    blocks = _col_blocks(n, shards)
    vsq = sum(np.sum(V[:, a:b] ** 2) for a, b in blocks)
    cost = np.zeros(maxiter)
    euclid = div == "euclidean"
    if not euclid and div not in ("kl", "kl_divergence"):
        raise ValueError("nmf_gram covers the euclidean and kl divergences only")

    for it in range(maxiter):
        if not W_fixed:
            if euclid:
                # packed all-reduce payload: A (m x K), G_H (K x K)
                A = sum(V[:, a:b] @ H[:, a:b].T for a, b in blocks)
                GH = sum(H[:, a:b] @ H[:, a:b].T for a, b in blocks)
                B = W @ GH
                av = np.sum(W * A, axis=0)  # diag(H V' W)
                bv = np.sum(W * B, axis=0)  # diag(H V_hat' W)
                neg = A + W * bv
                pos = B + W * av
            else:
                # payload: R (m x K), hs (K)
                R = sum((V[:, a:b] / (W @ H[:, a:b])) @ H[:, a:b].T for a, b in blocks)
                hs = sum(H[:, a:b].sum(axis=1) for a, b in blocks)
                ws = W.sum(axis=0)
                cv = np.sum(W * R, axis=0)  # diag(H (V./V_hat)' W)
                neg = R + W * (hs * ws)     # diag(H ones(n,m) W)_k = hs_k * ws_k
                pos = hs[None, :] + W * cv  # ones(m,n) H' = 1 hs'
            W = W * (neg / np.fmax(pos + lamW, EPS))
            W = W / np.sqrt(np.sum(W ** 2, axis=0))
        if not H_fixed:
            if euclid:
                GW = W.T @ W
                for a, b in blocks:  # purely local per shard
                    N = W.T @ V[:, a:b]
                    D = GW @ H[:, a:b]
                    H[:, a:b] = H[:, a:b] * (N / np.fmax(D + lamH, EPS))
            else:
                ws = W.sum(axis=0)
                for a, b in blocks:
                    N = W.T @ (V[:, a:b] / (W @ H[:, a:b]))
                    H[:, a:b] = H[:, a:b] * (N / np.fmax(ws[:, None] + lamH, EPS))
        if euclid:
            GW = W.T @ W
            nh = sum(np.sum((W.T @ V[:, a:b]) * H[:, a:b]) for a, b in blocks)
            GH = sum(H[:, a:b] @ H[:, a:b].T for a, b in blocks)
            c = 0.5 * (vsq - 2.0 * nh + np.sum(GW * GH))
        else:
            c = 0.0
            for a, b in blocks:
                S = W @ H[:, a:b]
                Vb = V[:, a:b]
                c += np.sum(Vb * np.log(Vb / S) - Vb + S)
        c += lamW * np.sum(np.abs(W)) + lamH * np.sum(np.abs(H))
        cost[it] = c
        if it > 0 and cost[it] < cost[it - 1] and cost[it - 1] - cost[it] < tol:
            cost = cost[: it + 1]
            break
    return W, H, cost


def _stack_shift(H, T):
    """Hs = [H_1; ...; H_T], H_t = H shifted right by t-1 with zero fill (cnmf.m:188)."""
    K, n = H.shape
    Hs = np.zeros((K * T, n))
    for t in range(T):
        Hs[t * K:(t + 1) * K, t:] = H[:, : n - t]
    return Hs


def _fold(P, K, T):
    """fold(P)[:, j] = sum_t P_t[:, j + t - 1]  (zero beyond n)  (cnmf.m:218-227)."""
    n = P.shape[1]
    out = np.zeros((K, n))
    for t in range(T):
        out[:, : n - t] += P[t * K:(t + 1) * K, t:]
    return out


def cnmf_stacked(V, K, T, config):
    """cnmf.m, Euclidean, single source, in stacked form.  W is m x K x T."""
    V = np.asarray(V, dtype=np.float64)
    m, n = V.shape
    lamW = max(float(config.get("W_sparsity", 0) or 0), 0.0)
    lamH = max(float(config.get("H_sparsity", 0) or 0), 0.0)
    maxiter = int(config.get("maxiter", 100) or 100)
    tol = config.get("tolerance", 1e-3)
    if tol is None or tol <= 0:
        tol = 1e-3
    W3 = np.array(config["W_init"], dtype=np.float64)
    H = np.array(config["H_init"], dtype=np.float64)
    # cnmf.m:161-165: per-basis tensor norm / T, H compensated (only here)
    wn = np.sqrt(np.sum(W3 ** 2, axis=(0, 2))) / T
    W3 = W3 / wn[None, :, None]
    H = H * wn[:, None]
    # Wc = [W_1 ... W_T]  (m x KT), column index = t*K + k
    Wc = np.concatenate([W3[:, :, t] for t in range(T)], axis=1)
    vsq = np.sum(V ** 2)
    cost = np.zeros(maxiter)
    for it in range(maxiter):
        Hs = _stack_shift(H, T)
        A = V @ Hs.T
        G = Hs @ Hs.T
        B = Wc @ G
        av = np.sum(Wc * A, axis=0)
        bv = np.sum(Wc * B, axis=0)
        Wc = Wc * ((A + Wc * bv) / np.fmax(B + Wc * av + lamW, EPS))
        ss = np.sum(Wc ** 2, axis=0).reshape(T, K).sum(axis=0)  # per basis k over (i, t)
        wn = np.sqrt(ss) / T
        Wc = Wc / np.tile(wn, T)[None, :]
        GW = Wc.T @ Wc
        Pn = Wc.T @ V
        Pp = GW @ Hs
        H = H * (_fold(Pn, K, T) / np.fmax(_fold(Pp, K, T) + lamH, EPS))
        Hs = _stack_shift(H, T)
        c = 0.5 * (vsq - 2.0 * np.sum(Pn * Hs) + np.sum(GW * (Hs @ Hs.T)))
        c += lamW * np.sum(np.abs(Wc)) + lamH * np.sum(np.abs(H))
        cost[it] = c
        if it > 0 and cost[it] < cost[it - 1] and cost[it - 1] - cost[it] < tol:
            cost = cost[: it + 1]
            break
    W3 = np.stack([Wc[:, t * K:(t + 1) * K] for t in range(T)], axis=2)
    return W3, H, cost
