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
* IS / AB (and cnmf's KL / IS / AB) in "two-weight" form: A = Qn H', B = Qp H' and the
  Euclidean-shaped W step; per-basis lambda / fixed vectors == the per-source cell loops
* cnmfsc's W gradients through the Gram matrix of the shifted stack

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


def cnmf_stacked_sharded(V, K, T, config, shards):
    """cnmf.m, Euclidean, as cnmf_driver.cu runs it on `shards` column-sharded ranks (simulated in one
    process): every rank owns consecutive columns [lo, hi) of V and H, keeps the T-1 columns of H before and
    after them and the T-1 columns of V after them (halos), forms Hs, P = Wc'V and D = (Wc'Wc)Hs on own + halo
    columns and folds for its own columns; A = V Hs', Hs Hs' and the scalar sums are summed over the ranks,
    the W step is replicated.  Must agree with the literal cnmf to rounding for any `shards`."""
    V = np.asarray(V, dtype=np.float64)
    m, n = V.shape
    lamW = max(float(config.get("W_sparsity", 0) or 0), 0.0)
    lamH = max(float(config.get("H_sparsity", 0) or 0), 0.0)
    maxiter = int(config.get("maxiter", 100) or 100)
    h = T - 1
    W3 = np.array(config["W_init"], dtype=np.float64)
    H = np.array(config["H_init"], dtype=np.float64)
    wn = np.sqrt(np.sum(W3 ** 2, axis=(0, 2))) / T
    W3 = W3 / wn[None, :, None]
    H = H * wn[:, None]
    Wc = np.concatenate([W3[:, :, t] for t in range(T)], axis=1)
    bounds = _col_blocks(n, shards)
    assert all(hi - lo >= h for lo, hi in bounds), "every shard needs at least T-1 columns"
    Hown = [H[:, lo:hi].copy() for lo, hi in bounds]  # each rank's own columns
    vsq = np.sum(V ** 2)
    cost = np.zeros(maxiter)

    def local_views(r):
        """(hL, hR, H with halos, Hs over own + right-halo columns) of rank r after a halo exchange."""
        lo, hi = bounds[r]
        hL = h if r > 0 else 0
        hR = h if r + 1 < shards else 0
        left = Hown[r - 1][:, Hown[r - 1].shape[1] - hL:] if hL else np.zeros((K, 0))
        right = Hown[r + 1][:, :hR] if hR else np.zeros((K, 0))
        Hx = np.concatenate([left, Hown[r], right], axis=1)
        nx = (hi - lo) + hR
        Hs = np.zeros((K * T, nx))
        for t in range(T):  # hstack_kernel: Hs_t[:, j] = Hx[:, hL + j - t] when j + hL >= t
            for j in range(nx):
                if j + hL >= t:
                    Hs[t * K:(t + 1) * K, j] = Hx[:, hL + j - t]
        return hL, hR, Hs

    for it in range(maxiter):
        A = np.zeros((m, K * T))
        G = np.zeros((K * T, K * T))
        loc = []
        for r, (lo, hi) in enumerate(bounds):
            hL, hR, Hs = local_views(r)
            nown = hi - lo
            A += V[:, lo:hi] @ Hs[:, :nown].T       # own columns only
            G += Hs[:, :nown] @ Hs[:, :nown].T
            loc.append((hR, Hs))
        B = Wc @ G
        av = np.sum(Wc * A, axis=0)
        bv = np.sum(Wc * B, axis=0)
        Wc = Wc * ((A + Wc * bv) / np.fmax(B + Wc * av + lamW, EPS))
        ss = np.sum(Wc ** 2, axis=0).reshape(T, K).sum(axis=0)
        Wc = Wc / np.tile(np.sqrt(ss) / T, T)[None, :]
        GW = Wc.T @ Wc
        dot_nh = 0.0
        for r, (lo, hi) in enumerate(bounds):
            hR, Hs = loc[r]
            nown = hi - lo
            P = Wc.T @ V[:, lo:hi + hR]              # own + right-halo columns of V
            D = GW @ Hs
            neg = np.zeros((K, nown))
            pos = np.zeros((K, nown))
            for t in range(T):                       # fold_update_kernel: j + t < n_src
                w = min(nown, nown + hR - t)
                neg[:, :w] += P[t * K:(t + 1) * K, t:t + w]
                pos[:, :w] += D[t * K:(t + 1) * K, t:t + w]
            Hown[r] = Hown[r] * (neg / np.fmax(pos + lamH, EPS))
            dot_nh += np.sum(neg * Hown[r])          # <fold(P), H_new> = <P, Hs_new> summed over the ranks
        G2 = np.zeros((K * T, K * T))
        for r, (lo, hi) in enumerate(bounds):
            _, _, Hs = local_views(r)
            G2 += Hs[:, :hi - lo] @ Hs[:, :hi - lo].T
        c = 0.5 * (vsq - 2.0 * dot_nh + np.sum(GW * G2))
        c += lamW * np.sum(np.abs(Wc)) + lamH * sum(np.sum(np.abs(x)) for x in Hown)
        cost[it] = c
    W3 = np.stack([Wc[:, t * K:(t + 1) * K] for t in range(T)], axis=2)
    return W3, np.concatenate(Hown, axis=1), cost


# --------------------------------------------------------------------------
# IS / AB divergences ("two-weight" form of nmf_driver.cu::plan_two_weight and
# cnmf_driver.cu): both gradients are contractions with element-wise weights.
# --------------------------------------------------------------------------
def _weights(V, V_hat, div, alpha, beta):
    """Qn, Qp, outer exponent (nmf.m:154-164, 185-195)."""
    if div in ("is_divergence", "is"):
        return V / V_hat ** 2, 1.0 / V_hat, 1.0
    if div in ("kl_divergence", "kl"):  # cnmf.m:141-143 reaches KL as alpha = 1, beta = 0
        alpha, beta = 1.0, 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        if alpha == 0:  # dual (nmf.m:124-128)
            return V ** (alpha - 1) * V_hat ** beta, V ** (alpha + beta - 1), 1.0 / beta
        return V ** alpha * V_hat ** (beta - 1), V_hat ** (alpha + beta - 1), 1.0 / alpha


def nmf_two_weight(V, K, config, lam_w=None, lam_h=None, fix_w=None, fix_h=None):
    """nmf.m for 'is' / 'ab' (single source, or concatenated sources with per-basis vectors) the way
    the device computes it: A = Qn H', B = Qp H', neg = A + W diag(<W_k, B_k>), pos = B + W diag(<W_k, A_k>)
    (the Euclidean-shaped W step), N = W' Qn, D = W' Qp.  Returns W, H, cost."""
    V = np.asarray(V, dtype=np.float64)
    div = config["divergence"]
    alpha, beta = float(config.get("alpha", 1)), float(config.get("beta", 1))
    maxiter = int(config.get("maxiter", 100))
    W = np.array(config["W_init"], dtype=np.float64)
    H = np.array(config["H_init"], dtype=np.float64)
    lam_w = np.full(K, max(float(config.get("W_sparsity", 0) or 0), 0.0)) if lam_w is None else np.asarray(lam_w, float)
    lam_h = np.full(K, max(float(config.get("H_sparsity", 0) or 0), 0.0)) if lam_h is None else np.asarray(lam_h, float)
    fix_w = np.zeros(K, bool) if fix_w is None else np.asarray(fix_w, bool)
    fix_h = np.zeros(K, bool) if fix_h is None else np.asarray(fix_h, bool)
    W = W / np.sqrt((W ** 2).sum(0))  # nmf.m:132
    cost = np.zeros(maxiter)
    for it in range(maxiter):
        Qn, Qp, expo = _weights(V, W @ H, div, alpha, beta)
        A, B = Qn @ H.T, Qp @ H.T
        a, b = (W * A).sum(0), (W * B).sum(0)
        neg, pos = (A + W * b) ** expo, (B + W * a) ** expo
        Wn = W * (neg / np.fmax(pos + lam_w, EPS))
        Wn = Wn / np.sqrt((Wn ** 2).sum(0))
        W = np.where(fix_w, W, Wn)
        Qn, Qp, expo = _weights(V, W @ H, div, alpha, beta)
        N, D = (W.T @ Qn) ** expo, (W.T @ Qp) ** expo
        Hn = H * (N / np.fmax(D + lam_h[:, None], EPS))
        H = np.where(fix_h[:, None], H, Hn)
        V_hat = W @ H
        if div in ("is_divergence", "is"):
            c = np.sum(np.log(V_hat / V) + V / V_hat - 1)
        else:
            c = (np.float64(-1.0) / np.float64(alpha * beta)) * np.sum(
                V ** alpha * V_hat ** beta
                - (alpha * V ** (alpha + beta) + beta * V_hat ** (alpha + beta) + beta) / np.float64(alpha + beta))
        cost[it] = c + np.sum(lam_w * np.abs(W).sum(0)) + np.sum(lam_h * np.abs(H).sum(1))
    return W, H, cost


def cnmf_two_weight(V, K, T, config):
    """cnmf.m for 'kl' / 'is' / 'ab' in stacked two-weight form (cnmf_driver.cu): A = Qn Hs', B = Qp Hs',
    per-column Euclidean-shaped W step, per-basis normalisation, fold(Wc' Qn), fold(Wc' Qp) with the
    unshifted V_pos of the KL branch (cnmf.m:221-222)."""
    V = np.asarray(V, dtype=np.float64)
    m, n = V.shape
    div = config["divergence"]
    alpha, beta = float(config.get("alpha", 1)), float(config.get("beta", 1))
    lamW = max(float(config.get("W_sparsity", 0) or 0), 0.0)
    lamH = max(float(config.get("H_sparsity", 0) or 0), 0.0)
    maxiter = int(config.get("maxiter", 100))
    W3 = np.array(config["W_init"], dtype=np.float64).reshape(m, K, T)
    H = np.array(config["H_init"], dtype=np.float64)
    nrm = np.sqrt((W3 ** 2).sum(axis=(0, 2))) / T  # cnmf.m:157-166
    W3 = W3 / nrm[None, :, None]
    H = H * nrm[:, None]
    Wc = np.concatenate([W3[:, :, t] for t in range(T)], axis=1)  # column k + K*t
    kl = div in ("kl_divergence", "kl")
    cost = np.zeros(maxiter)
    for it in range(maxiter):
        Hs = _stack_shift(H, T)
        Qn, Qp, expo = _weights(V, Wc @ Hs, div, alpha, beta)
        A, B = Qn @ Hs.T, Qp @ Hs.T
        a, b = (Wc * A).sum(0), (Wc * B).sum(0)
        Wc = Wc * ((A + Wc * b) ** expo / np.fmax((B + Wc * a) ** expo + lamW, EPS))
        sq = (Wc ** 2).sum(0).reshape(T, K).sum(0)  # per basis over all frames
        Wc = Wc / np.tile(np.sqrt(sq) / T, T)
        Qn, Qp, expo = _weights(V, Wc @ Hs, div, alpha, beta)
        neg = _fold(Wc.T @ Qn, K, T)
        P = Wc.T @ (Qp if not np.isscalar(Qp) else np.full_like(V, Qp))
        pos = P.reshape(T, K, n).sum(0) if kl else _fold(P, K, T)
        H = H * (neg ** expo / np.fmax(pos ** expo + lamH, EPS))
        V_hat = Wc @ _stack_shift(H, T)
        if kl:
            c = np.sum(V * np.log(V / V_hat) - V + V_hat)
        elif div in ("is_divergence", "is"):
            c = np.sum(np.log(V_hat / V) + V / V_hat - 1)
        else:
            c = (np.float64(-1.0) / np.float64(alpha * beta)) * np.sum(
                V ** alpha * V_hat ** beta
                - (alpha * V ** (alpha + beta) + beta * V_hat ** (alpha + beta) + beta) / np.float64(alpha + beta))
        cost[it] = c + lamW * np.abs(Wc).sum() + lamH * np.abs(H).sum()
    W3 = np.stack([Wc[:, t * K:(t + 1) * K] for t in range(T)], axis=2)
    return W3, H, cost


def cnmfsc_w_gradients(V, W0c, Wprev, H, K, T, t):
    """The two W gradients of cnmfsc.m the way cnmfsc_driver.cu forms them, for frame t (0-based):
    multiplicative branch (lines 257-263)  pos_t = Wc_current (Hs Hs')[:, frame t];
    sparse branch, t >= 1 (lines 218-224 after the trial of line 235)  pos_t = Wprev (H Hs_t') taken from the
    rows of frame t and the first K columns of the symmetric Gram matrix."""
    Hs = _stack_shift(H, T)
    G = Hs @ Hs.T
    neg = (V @ Hs.T)[:, t * K:(t + 1) * K]
    pos_mu = W0c @ G[:, t * K:(t + 1) * K]
    pos_sparse = Wprev @ G[t * K:(t + 1) * K, :K].T if Wprev is not None else None
    return neg, pos_mu, pos_sparse
