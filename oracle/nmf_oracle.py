"""CPU ORACLE - TEST INFRASTRUCTURE ONLY.  Not part of the product.

Literal, line-by-line NumPy float64 restatement of the reference's MATLAB hot
path (colinvaz/nmf-toolbox): ``nmf.m``, ``cnmf.m``, ``nmfsc.m``, ``projfunc.m``
and ``ReconstructFromDecomposition.m``, plus ``cnmfsc.m``, ``lnmf.m`` and ``constrainednmf.m``.  It deliberately keeps the reference's
exact operation sequence - the dense ``ones(n, m)`` products, the
``diag(diag(...))`` terms, the per-frame loops, the redundant GEMMs - so that it
is a readable statement of WHAT the reference computes, not a fast one.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module, and only as the checker or
the timed CPU baseline.  The product (``nmf_toolbox_b200`` / ``libnmfb200.so``)
never imports, links or executes anything under ``oracle/``.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 and no MATLAB/Octave exists in this environment, so
this restatement cannot be checked against outputs of the reference itself.
It is pinned only by (i) invariants derivable from the reference code
(tests/test_oracle.py), (ii) an independently written Gram/trace-form
re-derivation (oracle/restructured.py) agreeing to 1e-12, and (iii) sklearn's
``NMF(solver="mu")`` for the one textbook Lee-Seung update in scope
(nmfsc.m:182,232).

MATLAB semantics honoured: column-major shapes, ``eps`` = 2**-52, ``x.^0 == 1``,
``max(a, b)`` ignoring NaN (``np.fmax``), ``real(sqrt(negative)) == 0``
(projfunc.m:37), ``find(v <= 0)`` (projfunc.m:49), 1-based shifts.
"""
from __future__ import annotations

import numpy as np

EPS = 2.0 ** -52  # MATLAB eps

__all__ = [
    "EPS",
    "reconstruct_from_decomposition",
    "projfunc",
    "nmf",
    "cnmf",
    "nmfsc",
    "cnmfsc",
    "lnmf",
    "constrainednmf",
    "constrained_label_matrix",
]


class ReferenceError_(ValueError):
    """Stands in for a MATLAB ``error(...)`` raised by the reference."""


def _diagdiag(M):
    """MATLAB ``diag(diag(M))``."""
    return np.diag(np.diag(M))


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


# --------------------------------------------------------------------------
# ReconstructFromDecomposition.m:23-39
# --------------------------------------------------------------------------
def reconstruct_from_decomposition(W, H):
    """V_hat = W*H (matrix W) or sum_t W(:,:,t) * [zeros(K,t-1) H(:,1:n-t+1)] (3-D W).

    Follows ReconstructFromDecomposition.m:23-39.  Lists play the role of cell
    arrays: W cells are concatenated horizontally, H cells vertically
    (``cell2mat`` of a 1xS / Sx1 cell, lines 23-28).
    """
    if isinstance(W, (list, tuple)):
        W = np.concatenate(list(W), axis=1)  # RFD.m:24
    if isinstance(H, (list, tuple)):
        H = np.concatenate(list(H), axis=0)  # RFD.m:27
    W = np.asarray(W, dtype=np.float64)
    H = np.asarray(H, dtype=np.float64)
    if W.ndim == 2:  # RFD.m:30-31
        return W @ H
    if W.ndim == 3:  # RFD.m:32-38
        m, K, T = W.shape
        n = H.shape[1]
        V_hat = np.zeros((m, n))
        for t in range(1, T + 1):
            H_shifted = np.concatenate([np.zeros((K, t - 1)), H[:, : n - t + 1]], axis=1)
            V_hat = V_hat + W[:, :, t - 1] @ H_shifted
        return V_hat
    raise ReferenceError_("W must be a matrix or a 3-D tensor")


# --------------------------------------------------------------------------
# projfunc.m:13-65
# --------------------------------------------------------------------------
def projfunc(s, k1, k2, nn=1):
    """Hoyer's L1/L2 projection, projfunc.m:13-65.  Returns ``(v, usediters)``."""
    s = np.asarray(s, dtype=np.float64).reshape(-1).copy()
    N = s.size  # projfunc.m:13
    isneg = None
    if not nn:  # projfunc.m:16-19
        isneg = s < 0
        s = np.abs(s)
    v = s + (k1 - s.sum()) / N  # projfunc.m:22
    zerocoeff = np.zeros(0, dtype=np.int64)  # projfunc.m:25
    j = 0
    while True:
        midpoint = np.ones(N) * k1 / (N - zerocoeff.size)  # projfunc.m:31
        midpoint[zerocoeff] = 0  # projfunc.m:32
        w = v - midpoint  # projfunc.m:33
        a = np.sum(w ** 2)  # projfunc.m:34
        b = 2 * (w @ v)  # projfunc.m:35
        c = np.sum(v ** 2) - k2  # projfunc.m:36
        disc = b * b - 4 * a * c
        root = np.sqrt(disc) if disc >= 0 else 0.0  # real(sqrt(.)), projfunc.m:37
        with np.errstate(divide="ignore", invalid="ignore"):
            alphap = (-b + root) / (2 * a)
        v = alphap * w + v  # projfunc.m:38
        if np.all(v >= 0):  # projfunc.m:40-44
            usediters = j + 1
            break
        j += 1  # projfunc.m:46
        zerocoeff = np.nonzero(v <= 0)[0]  # projfunc.m:49
        v[zerocoeff] = 0  # projfunc.m:50
        tempsum = v.sum()  # projfunc.m:51
        v = v + (k1 - tempsum) / (N - zerocoeff.size)  # projfunc.m:52
        v[zerocoeff] = 0  # projfunc.m:53
        if not np.all(np.isfinite(v)):
            # MATLAB would loop forever on NaN (all(NaN>=0) is false); bail out instead.
            raise ReferenceError_("projfunc diverged (non-finite values)")
    if not nn:  # projfunc.m:58-60
        v = (-2.0 * isneg + 1.0) * v
    # projfunc.m:63-65: imaginary-part check is vacuous here (real arithmetic throughout).
    return v, usediters


# --------------------------------------------------------------------------
# shared config defaulting (nmf.m:238-413, cnmf.m:271-449)
# --------------------------------------------------------------------------
def _per_source(value, num_sources, what, clamp_nonneg):
    """nmf.m:312-401 - scalar or 1-element cell is broadcast, S-element cell is kept."""
    if value is None or (isinstance(value, (list, tuple)) and len(value) == 0):
        return [0 if clamp_nonneg else False] * num_sources
    if isinstance(value, (list, tuple)) and len(value) > 1:
        if len(value) != num_sources:
            raise ReferenceError_(
                f"Requested {num_sources} sources. Given {len(value)} {what}."
            )
        return [max(v, 0) if clamp_nonneg else v for v in value]
    v = value[0] if isinstance(value, (list, tuple)) else value
    if clamp_nonneg:
        v = max(v, 0)
    return [v] * num_sources


def _validate(V, num_basis_elems, config, context_len=None, rng=None):
    """Private ValidateParameters of nmf.m (238-413) / cnmf.m (271-449)."""
    cfg = dict(config or {})
    m, n = V.shape
    S = len(num_basis_elems)
    rng = rng or np.random.default_rng()
    cfg.setdefault("divergence", "euclidean")  # nmf.m:250-252
    is_ab = cfg["divergence"] in ("ab_divergence", "ab")
    if "alpha" not in cfg or not is_ab:  # nmf.m:255-259
        cfg["alpha"] = 1
    if "beta" not in cfg or not is_ab:  # nmf.m:262-266
        cfg["beta"] = 1

    H_init = cfg.get("H_init")
    if H_init is None or (isinstance(H_init, (list, tuple)) and len(H_init) == 0):  # nmf.m:269-278
        is_H_cell = S != 1
        H_init = [np.fmax(rng.random((num_basis_elems[s], n)), EPS) for s in range(S)]
    elif isinstance(H_init, (list, tuple)):
        if len(H_init) != S:  # nmf.m:279-280
            raise ReferenceError_(
                f"Requested {S} sources. Given {len(H_init)} initial encoding matrices."
            )
        is_H_cell = True
        H_init = [np.array(h, dtype=np.float64) for h in H_init]
    else:  # nmf.m:281-283
        is_H_cell = False
        H_init = [np.array(H_init, dtype=np.float64)]
    cfg["H_init"] = H_init

    W_init = cfg.get("W_init")
    if W_init is None or (isinstance(W_init, (list, tuple)) and len(W_init) == 0):
        is_W_cell = S != 1
        W_init = []
        for s in range(S):
            if context_len is None:  # nmf.m:298-299
                w = np.fmax(rng.random((m, num_basis_elems[s])), EPS)
                w = w @ np.diag(1.0 / np.sqrt(np.sum(w ** 2, axis=0)))
            else:  # cnmf.m:331-335
                w = rng.random((m, num_basis_elems[s], context_len))
                for k in range(num_basis_elems[s]):
                    w_norm = np.linalg.norm(w[:, k, :], "fro") / context_len
                    w[:, k, :] = w[:, k, :] / w_norm
            W_init.append(w)
    elif isinstance(W_init, (list, tuple)):
        if len(W_init) != S:  # nmf.m:301-302
            raise ReferenceError_(
                f"Requested {S} sources. Given {len(W_init)} initial basis matrices."
            )
        is_W_cell = True
        W_init = [np.array(w, dtype=np.float64) for w in W_init]
    else:
        is_W_cell = False
        W_init = [np.array(W_init, dtype=np.float64)]
    cfg["W_init"] = W_init

    cfg["W_sparsity"] = _per_source(cfg.get("W_sparsity"), S, "sparsity levels", True)
    cfg["H_sparsity"] = _per_source(cfg.get("H_sparsity"), S, "sparsity levels", True)
    cfg["W_fixed"] = _per_source(cfg.get("W_fixed"), S, "update switches", False)
    cfg["H_fixed"] = _per_source(cfg.get("H_fixed"), S, "update switches", False)
    if "maxiter" not in cfg or cfg["maxiter"] is None or cfg["maxiter"] <= 0:  # nmf.m:404-406
        cfg["maxiter"] = 100
    if "tolerance" not in cfg or cfg["tolerance"] is None or cfg["tolerance"] <= 0:  # nmf.m:409-411
        cfg["tolerance"] = 1e-3
    return cfg, is_W_cell, is_H_cell


def _cost(divergence, V, V_hat, alpha, beta):
    """nmf.m:206-215 / cnmf.m:239-248 (no 'frobenius' case: cost stays 0, a reference quirk)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        if divergence == "euclidean":
            return 0.5 * np.sum(np.sum((V - V_hat) ** 2))
        if divergence in ("kl_divergence", "kl"):
            return np.sum(np.sum(V * np.log(V / V_hat) - V + V_hat))
        if divergence in ("is_divergence", "is"):
            return np.sum(np.sum(np.log(V_hat / V) + (V / V_hat) - 1))
        if divergence in ("ab_divergence", "ab"):
            # IEEE division as in MATLAB: alpha*beta == 0 or alpha+beta == 0 give Inf / NaN, no error
            return (np.float64(-1.0) / np.float64(alpha * beta)) * np.sum(
                np.sum(
                    V ** alpha * V_hat ** beta
                    - (alpha * V ** (alpha + beta) + beta * V_hat ** (alpha + beta) + beta)
                    / np.float64(alpha + beta)
                )
            )
    return 0.0


# --------------------------------------------------------------------------
# nmf.m:109-236
# --------------------------------------------------------------------------
def nmf(V, num_basis_elems, config=None, rng=None):
    """``[W, H, cost] = nmf(V, num_basis_elems, config)`` - nmf.m:1, loop 143-225."""
    V = np.asarray(V, dtype=np.float64)
    m, n = V.shape  # nmf.m:113
    num_basis_elems = _as_list(num_basis_elems)  # nmf.m:114-116
    S = len(num_basis_elems)
    cfg, is_W_cell, is_H_cell = _validate(V, num_basis_elems, config, None, rng)  # nmf.m:118
    div = cfg["divergence"]
    alpha, beta = cfg["alpha"], cfg["beta"]
    if div in ("ab_divergence", "ab") and alpha == 0 and beta == 0:  # nmf.m:120-122
        raise ReferenceError_("alpha = 0 and beta = 0 is not supported at this time.")
    use_dual = alpha == 0  # nmf.m:124-128

    W = [w.copy() for w in cfg["W_init"]]  # nmf.m:130
    H = [h.copy() for h in cfg["H_init"]]  # nmf.m:131
    for s in range(S):  # nmf.m:132-134
        W[s] = W[s] @ np.diag(1.0 / np.sqrt(np.sum(W[s] ** 2, axis=0)))
    W_all = np.concatenate(W, axis=1)  # nmf.m:136
    H_all = np.concatenate(H, axis=0)  # nmf.m:137
    V_hat = reconstruct_from_decomposition(W_all, H_all)  # nmf.m:139
    maxiter = int(cfg["maxiter"])
    cost = np.zeros(maxiter)  # nmf.m:141
    # the dense ones(n, m) / ones(m, n) of nmf.m:152-156,184,187 (only the KL / IS branches use them)
    needs_ones = div in ("kl_divergence", "kl", "is_divergence", "is")
    ones_nm = np.ones((n, m)) if needs_ones else None
    ones_mn = np.ones((m, n)) if needs_ones else None

    with np.errstate(divide="ignore", invalid="ignore"):
        for it in range(1, maxiter + 1):  # nmf.m:143
            for s in range(S):  # nmf.m:145
                if not cfg["W_fixed"][s]:
                    Ws, Hs = W[s], H[s]
                    if div == "euclidean":  # nmf.m:149-150
                        neg = V @ Hs.T + Ws @ _diagdiag(Hs @ V_hat.T @ Ws)
                        pos = V_hat @ Hs.T + Ws @ _diagdiag(Hs @ V.T @ Ws)
                    elif div in ("kl_divergence", "kl"):  # nmf.m:152-153
                        neg = (V / V_hat) @ Hs.T + Ws @ _diagdiag(Hs @ ones_nm @ Ws)
                        pos = ones_mn @ Hs.T + Ws @ _diagdiag(Hs @ (V.T / V_hat.T) @ Ws)
                    elif div in ("is_divergence", "is"):  # nmf.m:155-156
                        neg = (V / V_hat ** 2) @ Hs.T + Ws @ _diagdiag(Hs @ (ones_nm / V_hat.T) @ Ws)
                        pos = (ones_mn / V_hat) @ Hs.T + Ws @ _diagdiag(Hs @ (V.T / V_hat.T ** 2) @ Ws)
                    elif div in ("ab_divergence", "ab"):  # nmf.m:158-164
                        if use_dual:
                            neg = ((V ** (alpha - 1) * V_hat ** beta) @ Hs.T
                                   + Ws @ _diagdiag(Hs @ V.T ** (alpha + beta - 1) @ Ws)) ** (1.0 / beta)
                            pos = (V ** (alpha + beta - 1) @ Hs.T
                                   + Ws @ _diagdiag(Hs @ (V ** (alpha - 1) * V_hat ** beta).T @ Ws)) ** (1.0 / beta)
                        else:
                            neg = ((V ** alpha * V_hat ** (beta - 1)) @ Hs.T
                                   + Ws @ _diagdiag(Hs @ V_hat.T ** (alpha + beta - 1) @ Ws)) ** (1.0 / alpha)
                            pos = (V_hat ** (alpha + beta - 1) @ Hs.T
                                   + Ws @ _diagdiag(Hs @ (V ** alpha * V_hat ** (beta - 1)).T @ Ws)) ** (1.0 / alpha)
                    else:  # nmf.m:165-166
                        raise ReferenceError_(
                            "No update equations defined for cost function with divergence type " + str(div)
                        )
                    Ws = Ws * (neg / np.fmax(pos + cfg["W_sparsity"][s], EPS))  # nmf.m:168
                    Ws = Ws @ np.diag(1.0 / np.sqrt(np.sum(Ws ** 2, axis=0)))  # nmf.m:169
                    W[s] = Ws
            W_all = np.concatenate(W, axis=1)  # nmf.m:172
            V_hat = reconstruct_from_decomposition(W_all, H_all)  # nmf.m:173

            for s in range(S):  # nmf.m:176
                if not cfg["H_fixed"][s]:
                    Ws, Hs = W[s], H[s]
                    if div == "euclidean":  # nmf.m:180-181
                        neg = Ws.T @ V
                        pos = Ws.T @ V_hat
                    elif div in ("kl_divergence", "kl"):  # nmf.m:183-184
                        neg = Ws.T @ (V / V_hat)
                        pos = Ws.T @ ones_mn
                    elif div in ("is_divergence", "is"):  # nmf.m:186-187
                        neg = Ws.T @ (V / V_hat ** 2)
                        pos = Ws.T @ (ones_mn / V_hat)
                    elif div in ("ab_divergence", "ab"):  # nmf.m:189-195
                        if use_dual:
                            neg = (Ws.T @ (V ** (alpha - 1) * V_hat ** beta)) ** (1.0 / beta)
                            pos = (Ws.T @ V ** (alpha + beta - 1)) ** (1.0 / beta)
                        else:
                            neg = (Ws.T @ (V ** alpha * V_hat ** (beta - 1))) ** (1.0 / alpha)
                            pos = (Ws.T @ V_hat ** (alpha + beta - 1)) ** (1.0 / alpha)
                    else:  # nmf.m:196-197
                        raise ReferenceError_(
                            "No update equations defined for cost function with divergence type " + str(div)
                        )
                    H[s] = Hs * (neg / np.fmax(pos + cfg["H_sparsity"][s], EPS))  # nmf.m:199
            H_all = np.concatenate(H, axis=0)  # nmf.m:202
            V_hat = reconstruct_from_decomposition(W_all, H_all)  # nmf.m:203

            c = _cost(div, V, V_hat, alpha, beta)  # nmf.m:206-215
            for s in range(S):  # nmf.m:216-218
                c = c + cfg["W_sparsity"][s] * np.sum(np.abs(W[s])) + cfg["H_sparsity"][s] * np.sum(np.abs(H[s]))
            cost[it - 1] = c
            # nmf.m:221-224
            if it > 1 and cost[it - 1] < cost[it - 2] and cost[it - 2] - cost[it - 1] < cfg["tolerance"]:
                cost = cost[:it]
                break

    W_out = W if is_W_cell else W[0]  # nmf.m:228-230
    H_out = H if is_H_cell else H[0]  # nmf.m:232-234
    return W_out, H_out, cost


# --------------------------------------------------------------------------
# cnmf.m:122-269
# --------------------------------------------------------------------------
def cnmf(V, num_basis_elems, context_len, config=None, rng=None):
    """``[W, H, cost] = cnmf(V, num_basis_elems, context_len, config)`` - cnmf.m:1, loop 175-258."""
    V = np.asarray(V, dtype=np.float64)
    m, n = V.shape  # cnmf.m:126
    num_basis_elems = _as_list(num_basis_elems)
    S = len(num_basis_elems)
    T = int(context_len)
    cfg, is_W_cell, is_H_cell = _validate(V, num_basis_elems, config, T, rng)  # cnmf.m:131
    div = cfg["divergence"]
    alpha, beta = cfg["alpha"], cfg["beta"]
    if div in ("ab_divergence", "ab") and alpha == 0 and beta == 0:  # cnmf.m:133-135
        raise ReferenceError_("alpha = 0 and beta = 0 is not supported at this time.")
    if div in ("euclidean", "frobenius"):  # cnmf.m:137-147
        alpha, beta = 1, 1
    elif div in ("kl_divergence", "kl"):
        alpha, beta = 1, 0
    elif div in ("is_divergence", "is"):
        alpha, beta = 1, -1
    use_dual = alpha == 0  # cnmf.m:149-153
    is_kl = div in ("kl_divergence", "kl")

    W = [w.copy() for w in cfg["W_init"]]  # cnmf.m:155
    H = [h.copy() for h in cfg["H_init"]]  # cnmf.m:156
    for s in range(S):  # cnmf.m:157-166
        for k in range(num_basis_elems[s]):
            w_norm = np.linalg.norm(W[s][:, k, :], "fro") / T
            W[s][:, k, :] = W[s][:, k, :] / w_norm
            H[s][k, :] = w_norm * H[s][k, :]
    W_all = np.concatenate(W, axis=1)  # cnmf.m:168
    H_all = np.concatenate(H, axis=0)  # cnmf.m:169
    V_hat = reconstruct_from_decomposition(W_all, H_all)  # cnmf.m:171
    maxiter = int(cfg["maxiter"])
    cost = np.zeros(maxiter)  # cnmf.m:173

    with np.errstate(divide="ignore", invalid="ignore"):
        for it in range(1, maxiter + 1):  # cnmf.m:175
            for s in range(S):  # cnmf.m:177
                if not cfg["W_fixed"][s]:
                    Ks = num_basis_elems[s]
                    for t in range(1, T + 1):  # cnmf.m:180 / 187
                        H_shifted = np.concatenate([np.zeros((Ks, t - 1)), H[s][:, : n - t + 1]], axis=1)
                        Wt = W[s][:, :, t - 1]
                        if use_dual:  # cnmf.m:182-184
                            g_neg = ((V ** (alpha - 1) * V_hat ** beta) @ H_shifted.T
                                     + Wt @ _diagdiag(H_shifted @ V.T ** (alpha + beta - 1) @ Wt)) ** (1.0 / beta)
                            g_pos = (V ** (alpha + beta - 1) @ H_shifted.T
                                     + Wt @ _diagdiag(H_shifted @ (V ** (alpha - 1) * V_hat ** beta).T @ Wt)) ** (1.0 / beta)
                        else:  # cnmf.m:191-193
                            g_neg = ((V ** alpha * V_hat ** (beta - 1)) @ H_shifted.T
                                     + Wt @ _diagdiag(H_shifted @ V_hat.T ** (alpha + beta - 1) @ Wt)) ** (1.0 / alpha)
                            g_pos = (V_hat ** (alpha + beta - 1) @ H_shifted.T
                                     + Wt @ _diagdiag(H_shifted @ (V ** alpha * V_hat ** (beta - 1)).T @ Wt)) ** (1.0 / alpha)
                        W[s][:, :, t - 1] = Wt * (g_neg / np.fmax(g_pos + cfg["W_sparsity"][s], EPS))
                    for k in range(Ks):  # cnmf.m:196-199
                        w_norm = np.linalg.norm(W[s][:, k, :], "fro") / T
                        W[s][:, k, :] = W[s][:, k, :] / w_norm
            W_all = np.concatenate(W, axis=1)  # cnmf.m:202
            H_all = np.concatenate(H, axis=0)  # cnmf.m:203
            V_hat = reconstruct_from_decomposition(W_all, H_all)  # cnmf.m:204

            for s in range(S):  # cnmf.m:207
                if not cfg["H_fixed"][s]:
                    Ks = num_basis_elems[s]
                    if use_dual:  # cnmf.m:210-211
                        V_neg = V ** (alpha - 1) * V_hat ** beta
                        V_pos = V ** (alpha + beta - 1)
                    else:  # cnmf.m:213-214
                        V_neg = V ** alpha * V_hat ** (beta - 1)
                        V_pos = V_hat ** (alpha + beta - 1)
                    g_neg = np.zeros((Ks, n))  # cnmf.m:216
                    g_pos = np.zeros((Ks, n))  # cnmf.m:217
                    for t in range(1, T + 1):  # cnmf.m:218
                        V_neg_shifted = np.concatenate([V_neg[:, t - 1:], np.zeros((m, t - 1))], axis=1)
                        if is_kl:  # cnmf.m:220-221
                            V_pos_shifted = V_pos
                        else:  # cnmf.m:223
                            V_pos_shifted = np.concatenate([V_pos[:, t - 1:], np.zeros((m, t - 1))], axis=1)
                        g_neg = g_neg + W[s][:, :, t - 1].T @ V_neg_shifted  # cnmf.m:225
                        g_pos = g_pos + W[s][:, :, t - 1].T @ V_pos_shifted  # cnmf.m:226
                    p = (1.0 / beta) if use_dual else (1.0 / alpha)  # cnmf.m:228-232
                    H[s] = H[s] * (g_neg ** p / np.fmax(g_pos ** p + cfg["H_sparsity"][s], EPS))
            H_all = np.concatenate(H, axis=0)  # cnmf.m:235
            V_hat = reconstruct_from_decomposition(W_all, H_all)  # cnmf.m:236

            c = _cost(div, V, V_hat, alpha, beta)  # cnmf.m:239-248 ('frobenius' -> 0)
            for s in range(S):  # cnmf.m:249-251
                c = c + cfg["W_sparsity"][s] * np.sum(np.abs(W[s])) + cfg["H_sparsity"][s] * np.sum(np.abs(H[s]))
            cost[it - 1] = c
            # cnmf.m:254-257
            if it > 1 and cost[it - 1] < cost[it - 2] and cost[it - 2] - cost[it - 1] < cfg["tolerance"]:
                cost = cost[:it]
                break

    W_out = W if is_W_cell else W[0]
    H_out = H if is_H_cell else H[0]
    return W_out, H_out, cost


# --------------------------------------------------------------------------
# lnmf.m:47-93
# --------------------------------------------------------------------------
def lnmf(V, num_basis_elems, config=None, rng=None):
    """[W, H, cost] = lnmf(V, num_basis_elems, config): local NMF, lnmf.m:47-93.

    Quirks kept: W columns have unit SUM (lines 63, 75); the loop is left without trimming `cost`
    (lines 88-90), so the returned vector always has maxiter entries (zeros after the stop); the
    stop test uses <= twice."""
    V = np.asarray(V, dtype=np.float64)
    m, n = V.shape
    cfg = dict(config or {})
    rng = rng or np.random.default_rng()
    K = int(num_basis_elems)
    H = cfg.get("H_init")
    H = np.fmax(rng.random((K, n)), EPS) if H is None or np.size(H) == 0 else np.array(H, dtype=np.float64)  # lnmf.m:103-105
    W = cfg.get("W_init")
    if W is None or np.size(W) == 0:  # lnmf.m:107-110
        W = np.fmax(rng.random((m, K)), EPS)
    W = np.array(W, dtype=np.float64)
    W_fixed = bool(cfg.get("W_fixed") or False)
    H_fixed = bool(cfg.get("H_fixed") or False)
    maxiter = cfg.get("maxiter")
    maxiter = 100 if maxiter is None or maxiter <= 0 else int(maxiter)  # lnmf.m:120-122
    tol = cfg.get("tolerance")
    tol = 1e-3 if tol is None or tol <= 0 else float(tol)  # lnmf.m:124-126
    W = W @ np.diag(1.0 / np.sum(W, axis=0))  # lnmf.m:63
    V_hat = reconstruct_from_decomposition(W, H)  # lnmf.m:66
    cost = np.zeros(maxiter)
    ones_mn = np.ones((m, n))
    for it in range(maxiter):
        if not W_fixed:  # lnmf.m:72-77
            W = W * (((V / V_hat) @ H.T) / np.fmax(ones_mn @ H.T, EPS))
            W = W @ np.diag(1.0 / np.sum(W, axis=0))
            V_hat = reconstruct_from_decomposition(W, H)
        if not H_fixed:  # lnmf.m:80-83
            H = np.sqrt(H * (W.T @ (V / V_hat)))
            V_hat = reconstruct_from_decomposition(W, H)
        cost[it] = np.sum(np.sum(V * np.log(V / V_hat) - V + V_hat))  # lnmf.m:86
        if it > 0 and cost[it] <= cost[it - 1] and cost[it - 1] - cost[it] <= tol:  # lnmf.m:88-90 (no trim)
            break
    return W, H, cost


# --------------------------------------------------------------------------
# nmfsc.m:57-245
# --------------------------------------------------------------------------
def nmfsc(V, num_basis_elems, config=None, rng=None, info=None):
    """``[W, H, cost] = nmfsc(V, num_basis_elems, config)`` - nmfsc.m:1, loop 141-245.

    ``info`` (optional dict) receives bookkeeping that the reference does not
    return (number of step halvings per iteration) - used by tests only.
    """
    V = np.asarray(V, dtype=np.float64)
    if V.min() < 0:  # nmfsc.m:57-59
        raise ReferenceError_("Negative values in data!")
    V = V / V.max()  # nmfsc.m:62
    m, n = V.shape  # nmfsc.m:65
    K = int(num_basis_elems)
    cfg = dict(config or {})
    rng = rng or np.random.default_rng()
    if cfg.get("W_init") is None:  # nmfsc.m:73-75
        cfg["W_init"] = rng.random((m, K))
    if cfg.get("H_init") is None:  # nmfsc.m:78-81
        h = rng.random((K, n))
        cfg["H_init"] = np.diag(1.0 / np.sqrt(np.sum(h ** 2, axis=1))) @ h
    W = np.array(cfg["W_init"], dtype=np.float64)  # nmfsc.m:83
    H = np.array(cfg["H_init"], dtype=np.float64)  # nmfsc.m:84

    L1a = L1s = None
    if cfg.get("W_sparsity") is None:  # nmfsc.m:87-88
        cfg["W_sparsity"] = 0
    elif cfg["W_sparsity"] > 0:  # nmfsc.m:89-97
        if cfg["W_sparsity"] > 1:
            cfg["W_sparsity"] = 1
        L1a = np.sqrt(m) - (np.sqrt(m) - 1) * cfg["W_sparsity"]
        for k in range(K):
            W[:, k] = projfunc(W[:, k], L1a, 1, 1)[0]
    if cfg.get("H_sparsity") is None:  # nmfsc.m:100-101
        cfg["H_sparsity"] = 0
    elif cfg["H_sparsity"] > 0:  # nmfsc.m:102-110
        if cfg["H_sparsity"] > 1:
            cfg["H_sparsity"] = 1
        L1s = np.sqrt(n) - (np.sqrt(n) - 1) * cfg["H_sparsity"]
        for k in range(K):
            H[k, :] = projfunc(H[k, :], L1s, 1, 1)[0]
    W_fixed = bool(cfg.get("W_fixed") or False)  # nmfsc.m:113-115
    H_fixed = bool(cfg.get("H_fixed") or False)  # nmfsc.m:118-120
    maxiter = cfg.get("maxiter")
    if maxiter is None or maxiter <= 0:  # nmfsc.m:123-125
        maxiter = 100
    maxiter = int(maxiter)
    tolerance = cfg.get("tolerance")
    if tolerance is None or tolerance <= 0:  # nmfsc.m:128-130
        tolerance = 1e-3

    stepsizeW = 1.0  # nmfsc.m:133
    stepsizeH = 1.0  # nmfsc.m:134
    cost = np.zeros(maxiter + 1)  # nmfsc.m:137
    V_hat = reconstruct_from_decomposition(W, H)  # nmfsc.m:138
    cost[0] = 0.5 * np.sum(np.sum((V - V_hat) ** 2))  # nmfsc.m:139
    halvings_H, halvings_W = [], []

    with np.errstate(divide="ignore", invalid="ignore"):
        for it in range(1, maxiter + 1):  # nmfsc.m:141
            if not H_fixed:  # nmfsc.m:143
                neg = W.T @ V  # nmfsc.m:144
                pos = W.T @ V_hat  # nmfsc.m:145
                if cfg["H_sparsity"] > 0:  # nmfsc.m:146
                    dH = pos - neg  # nmfsc.m:148
                    begobj = cost[it - 1]  # nmfsc.m:149
                    nh = 0
                    while True:  # nmfsc.m:152
                        Hnew = H - stepsizeH * dH  # nmfsc.m:154
                        for k in range(K):  # nmfsc.m:155-157
                            Hnew[k, :] = projfunc(Hnew[k, :], L1s, 1, 1)[0]
                        V_hat = reconstruct_from_decomposition(W, Hnew)  # nmfsc.m:160
                        newobj = 0.5 * np.sum(np.sum((V - V_hat) ** 2))  # nmfsc.m:161
                        if newobj <= begobj:  # nmfsc.m:164-166
                            break
                        stepsizeH = stepsizeH / 2  # nmfsc.m:169
                        nh += 1
                        if stepsizeH < 1e-200:  # nmfsc.m:170-174
                            cost = cost[:it]
                            if info is not None:
                                info.update(halvings_H=halvings_H, halvings_W=halvings_W, converged_early=True)
                            return W, H, cost
                    halvings_H.append(nh)
                    stepsizeH = 1.2 * stepsizeH  # nmfsc.m:178
                    H = Hnew  # nmfsc.m:179
                else:  # nmfsc.m:181-188
                    H = H * (neg / np.fmax(pos, EPS))
                    norms = np.sqrt(np.sum(H ** 2, axis=1))
                    H = np.diag(1.0 / norms) @ H
                    W = W @ np.diag(norms)
            if not W_fixed:  # nmfsc.m:192
                V_hat = reconstruct_from_decomposition(W, H)  # nmfsc.m:193
                neg = V @ H.T  # nmfsc.m:194
                pos = V_hat @ H.T  # nmfsc.m:195
                if cfg["W_sparsity"] > 0:  # nmfsc.m:196
                    begobj = 0.5 * np.sum(np.sum((V - V_hat) ** 2))  # nmfsc.m:197
                    dW = pos - neg  # nmfsc.m:200
                    nh = 0
                    while True:  # nmfsc.m:203
                        Wnew = W - stepsizeW * dW  # nmfsc.m:205
                        for k in range(K):  # nmfsc.m:206-208
                            Wnew[:, k] = projfunc(Wnew[:, k], L1a, 1, 1)[0]
                        V_hat = reconstruct_from_decomposition(Wnew, H)  # nmfsc.m:211
                        newobj = 0.5 * np.sum(np.sum((V - V_hat) ** 2))  # nmfsc.m:212
                        if newobj <= begobj:  # nmfsc.m:215-217
                            break
                        stepsizeW = stepsizeW / 2  # nmfsc.m:220
                        nh += 1
                        if stepsizeW < 1e-200:  # nmfsc.m:221-225
                            cost = cost[:it]
                            if info is not None:
                                info.update(halvings_H=halvings_H, halvings_W=halvings_W, converged_early=True)
                            return W, H, cost
                    halvings_W.append(nh)
                    stepsizeW = 1.2 * stepsizeW  # nmfsc.m:228
                    W = Wnew  # nmfsc.m:229
                else:  # nmfsc.m:232
                    W = W * (neg / np.fmax(pos, EPS))
            V_hat = reconstruct_from_decomposition(W, H)  # nmfsc.m:237
            cost[it] = 0.5 * np.sum(np.sum((V - V_hat) ** 2))  # nmfsc.m:238
            # nmfsc.m:241-244
            if it > 1 and cost[it] < cost[it - 1] and cost[it - 1] - cost[it] < tolerance:
                cost = cost[: it + 1]
                break
    if info is not None:
        info.update(halvings_H=halvings_H, halvings_W=halvings_W, converged_early=False)
    return W, H, cost


# --------------------------------------------------------------------------
# cnmfsc.m:66-277
# --------------------------------------------------------------------------
def cnmfsc(V, num_basis_elems, context_len, config=None, rng=None, info=None):
    """``[W, H, cost] = cnmfsc(V, num_basis_elems, context_len, config)`` - cnmfsc.m:1, loop 152-274.

    Literal, including the reference's quirks: the initial sparseness projection is applied to W
    while the iterations start from the unprojected W0 (lines 88-111); inside the W line search the
    trial reconstruction is ``ReconstructFromDecomposition(Wnew, H)`` with the 2-D frame ``Wnew``
    (line 227: a plain product with the unshifted H), and that V_hat is what the next frame sees;
    the multiplicative W branch updates V_hat incrementally with a clamp at 0 (line 262)."""
    V = np.asarray(V, dtype=np.float64)
    if V.min() < 0:  # cnmfsc.m:67-69
        raise ReferenceError_("Negative values in data!")
    V = V / V.max()  # cnmfsc.m:72
    m, n = V.shape
    K, T = int(num_basis_elems), int(context_len)
    cfg = dict(config or {})
    rng = rng or np.random.default_rng()
    if cfg.get("W_init") is None:  # cnmfsc.m:83-85
        cfg["W_init"] = rng.random((m, K, T))
    if cfg.get("H_init") is None:  # cnmfsc.m:88-91
        h = rng.random((K, n))
        cfg["H_init"] = np.diag(1.0 / np.sqrt(np.sum(h ** 2, axis=1))) @ h
    W0 = np.array(cfg["W_init"], dtype=np.float64).reshape(m, K, T)  # cnmfsc.m:93
    W = W0.copy()  # cnmfsc.m:94
    H = np.array(cfg["H_init"], dtype=np.float64)  # cnmfsc.m:95
    L1a = L1s = None
    if cfg.get("W_sparsity") is None:  # cnmfsc.m:98-99
        cfg["W_sparsity"] = 0
    elif cfg["W_sparsity"] > 0:  # cnmfsc.m:100-110 (projects W, not W0)
        if cfg["W_sparsity"] > 1:
            cfg["W_sparsity"] = 1
        L1a = np.sqrt(m) - (np.sqrt(m) - 1) * cfg["W_sparsity"]
        for t in range(T):
            for k in range(K):
                W[:, k, t] = projfunc(W[:, k, t], L1a, 1, 1)[0]
    if cfg.get("H_sparsity") is None:  # cnmfsc.m:114-115
        cfg["H_sparsity"] = 0
    elif cfg["H_sparsity"] > 0:  # cnmfsc.m:116-124
        if cfg["H_sparsity"] > 1:
            cfg["H_sparsity"] = 1
        L1s = np.sqrt(n) - (np.sqrt(n) - 1) * cfg["H_sparsity"]
        for k in range(K):
            H[k, :] = projfunc(H[k, :], L1s, 1, 1)[0]
    W_fixed = bool(cfg.get("W_fixed") or False)  # cnmfsc.m:127-129
    H_fixed = bool(cfg.get("H_fixed") or False)  # cnmfsc.m:132-134
    maxiter = cfg.get("maxiter")
    maxiter = 100 if maxiter is None or maxiter <= 0 else int(maxiter)  # cnmfsc.m:137-139
    tolerance = cfg.get("tolerance")
    tolerance = 1e-3 if tolerance is None or tolerance <= 0 else float(tolerance)  # cnmfsc.m:142-144

    def shift_left(X, t):  # [X(:, t:n) zeros(m, t-1)] with t 1-based
        out = np.zeros_like(X)
        out[:, : n - (t - 1)] = X[:, t - 1:]
        return out

    def shift_right(X, t):  # [zeros(K, t-1) X(:, 1:n-t+1)]
        out = np.zeros_like(X)
        out[:, t - 1:] = X[:, : n - (t - 1)]
        return out

    stepsizeW = np.ones(T)  # cnmfsc.m:147
    stepsizeH = 1.0  # cnmfsc.m:148
    cost = np.zeros(maxiter + 1)  # cnmfsc.m:151
    V_hat = reconstruct_from_decomposition(W, H)  # cnmfsc.m:152 (the PROJECTED W, while the loop starts from W0)
    cost[0] = 0.5 * np.sum(np.sum((V - V_hat) ** 2))  # cnmfsc.m:153
    trials_H, trials_W = [], []
    with np.errstate(divide="ignore", invalid="ignore"):
        for it in range(1, maxiter + 1):  # cnmfsc.m:155
            if not H_fixed:  # cnmfsc.m:157
                neg = np.zeros((K, n))
                pos = np.zeros((K, n))
                for t in range(1, T + 1):  # cnmfsc.m:160-165
                    neg = neg + W0[:, :, t - 1].T @ shift_left(V, t)
                    pos = pos + W0[:, :, t - 1].T @ shift_left(V_hat, t)
                if cfg["H_sparsity"] > 0:  # cnmfsc.m:166
                    dH = pos - neg  # cnmfsc.m:168
                    begobj = cost[it - 1]  # cnmfsc.m:169
                    nt = 0
                    while True:  # cnmfsc.m:172
                        Hnew = H - stepsizeH * dH  # cnmfsc.m:174
                        for k in range(K):  # cnmfsc.m:175-177
                            Hnew[k, :] = projfunc(Hnew[k, :], L1s, 1, 1)[0]
                        V_hat = reconstruct_from_decomposition(W0, Hnew)  # cnmfsc.m:180
                        newobj = 0.5 * np.sum(np.sum((V - V_hat) ** 2))  # cnmfsc.m:181
                        nt += 1
                        if newobj <= begobj:  # cnmfsc.m:184-186
                            break
                        stepsizeH = stepsizeH / 2  # cnmfsc.m:189
                        if stepsizeH < 1e-200:  # cnmfsc.m:190-194
                            return W, H, cost[:it]
                    trials_H.append(nt)
                    stepsizeH = 1.2 * stepsizeH  # cnmfsc.m:198
                    H = Hnew  # cnmfsc.m:199
                else:
                    H = H * (neg / (pos + EPS))  # cnmfsc.m:202
                    norms = np.sqrt(np.sum(H ** 2, axis=1))  # cnmfsc.m:205
                    H = np.diag(1.0 / norms) @ H  # cnmfsc.m:206
                    for t in range(T):  # cnmfsc.m:207-209
                        W0[:, :, t] = W0[:, :, t] @ np.diag(norms)
            if not W_fixed:  # cnmfsc.m:214
                V_hat = reconstruct_from_decomposition(W0, H)  # cnmfsc.m:215
                if cfg["W_sparsity"] > 0:  # cnmfsc.m:216
                    for t in range(1, T + 1):  # cnmfsc.m:217
                        begobj = 0.5 * np.sum(np.sum((V - V_hat) ** 2))  # cnmfsc.m:218
                        H_shifted = shift_right(H, t)  # cnmfsc.m:221
                        dW = V_hat @ H_shifted.T - V @ H_shifted.T  # cnmfsc.m:222-224
                        nt = 0
                        while True:  # cnmfsc.m:227
                            Wnew = W0[:, :, t - 1] - stepsizeW[t - 1] * dW  # cnmfsc.m:229
                            for k in range(K):  # cnmfsc.m:230-232
                                Wnew[:, k] = projfunc(Wnew[:, k], L1a, 1, 1)[0]
                            V_hat = reconstruct_from_decomposition(Wnew, H)  # cnmfsc.m:235 (2-D Wnew: plain product)
                            newobj = 0.5 * np.sum(np.sum((V - V_hat) ** 2))  # cnmfsc.m:236
                            nt += 1
                            if newobj <= begobj:  # cnmfsc.m:239-241
                                break
                            stepsizeW[t - 1] = stepsizeW[t - 1] / 2  # cnmfsc.m:244
                            if stepsizeW[t - 1] < 1e-200:  # cnmfsc.m:245-249
                                return W, H, cost[:it]
                        trials_W.append(nt)
                        stepsizeW[t - 1] = 1.2 * stepsizeW[t - 1]  # cnmfsc.m:252
                        W[:, :, t - 1] = Wnew  # cnmfsc.m:253
                else:
                    for t in range(1, T + 1):  # cnmfsc.m:257-263
                        H_shifted = shift_right(H, t)
                        neg = V @ H_shifted.T
                        pos = V_hat @ H_shifted.T
                        W[:, :, t - 1] = W0[:, :, t - 1] * (neg / np.fmax(pos, EPS))
                        V_hat = np.fmax(V_hat + (W[:, :, t - 1] - W0[:, :, t - 1]) @ H_shifted, 0)
            W0 = W.copy()  # cnmfsc.m:266
            V_hat = reconstruct_from_decomposition(W0, H)  # cnmfsc.m:269
            cost[it] = 0.5 * np.sum(np.sum((V - V_hat) ** 2))  # cnmfsc.m:270
            if it > 1 and cost[it] < cost[it - 1] and cost[it - 1] - cost[it] < tolerance:  # cnmfsc.m:273-276
                cost = cost[: it + 1]
                break
    if info is not None:
        info["trials_H"] = trials_H
        info["trials_W"] = trials_W
    return W, H, cost


# --------------------------------------------------------------------------
# constrainednmf.m:90-267
# --------------------------------------------------------------------------
def constrained_label_matrix(labels):
    """constrainednmf.m:147-170: class labels -> (sorted_idx, A).

    Samples are reordered so that unlabeled ones (label -1) come first and samples of one class are
    contiguous; ``A`` is the (n_unlabeled + num_classes) x n indicator matrix of that ORDERED
    arrangement: identity on the unlabeled samples, one row per class below it."""
    labels = np.asarray(labels).ravel()
    n = labels.size
    num_labeled = int(np.sum(labels > -1))  # constrainednmf.m:149
    uniq, inv = np.unique(labels, return_inverse=True)  # constrainednmf.m:151 / 156 (1-based in MATLAB)
    proc = inv + 1
    if num_labeled < n:  # some unlabeled samples (they got processed label 1)
        proc = proc - 1  # constrainednmf.m:152
        proc[proc == 0] = -1  # constrainednmf.m:153
        num_classes = len(uniq) - 1  # constrainednmf.m:154
    else:
        num_classes = len(uniq)  # constrainednmf.m:170
    sorted_idx = np.argsort(proc, kind="stable")  # constrainednmf.m:163 (MATLAB's sort is stable)
    sorted_labels = proc[sorted_idx]
    n_unl = n - num_labeled
    C = np.zeros((num_classes, num_labeled))  # constrainednmf.m:166-169
    for samp in range(n_unl, n):
        C[sorted_labels[samp] - 1, samp - n_unl] = 1
    A = np.block([[np.eye(n_unl), np.zeros((n_unl, num_labeled))],  # constrainednmf.m:170
                  [np.zeros((num_classes, n_unl)), C]])
    return sorted_idx, A


def constrainednmf(V, labels, num_basis_elems, config=None, rng=None):
    """``[W, H, Z, A, cost] = constrainednmf(V, labels, num_basis_elems, config)`` - constrainednmf.m:1.

    Extension for reproducibility (the reference draws ``Z = rand(...)`` unconditionally at line 174):
    ``config['Z_init']``, a num_basis_elems x (n_unlabeled + num_classes) matrix in the ORDERED
    arrangement (unlabeled samples first, then one column per class)."""
    V = np.asarray(V, dtype=np.float64)
    m, n = V.shape  # constrainednmf.m:96
    labels = np.asarray(labels).ravel()
    if labels.size != n:  # constrainednmf.m:98
        raise ReferenceError_("Length of the label vector not equal to number of samples.")
    cfg = dict(config or {})
    rng = rng or np.random.default_rng(0)
    K = int(num_basis_elems)
    W = np.array(cfg["W_init"], dtype=np.float64) if cfg.get("W_init") is not None else rng.random((m, K))  # 100-102
    lam_w = cfg.get("W_sparsity") or 0  # 103-105
    lam_z = cfg.get("Z_sparsity") or 0  # 106-108
    W_fixed = bool(cfg.get("W_fixed") or False)  # 109-111
    Z_fixed = bool(cfg.get("Z_fixed") or False)  # 112-114
    div = cfg.get("divergence", "euclidean")  # 115-117
    is_ab = div in ("ab_divergence", "ab")
    alpha = cfg.get("alpha", 1) if is_ab else 1  # 118-122
    beta = cfg.get("beta", 1) if is_ab else 1  # 123-127
    use_dual = alpha == 0  # 128-132
    maxiter = cfg.get("maxiter")
    if maxiter is None or maxiter <= 0:  # 133-135
        maxiter = 100
    tol = cfg.get("tolerance")
    if tol is None or tol <= 0:  # 136-138
        tol = 1e-3
    if is_ab and alpha == 0 and beta == 0:  # 140-142
        raise ReferenceError_("alpha = 0 and beta = 0 is not supported at this time.")
    W = W @ np.diag(1.0 / np.sqrt(np.sum(W ** 2, axis=0)))  # 144-145

    sorted_idx, A = constrained_label_matrix(labels)  # 147-170
    V = V[:, sorted_idx]  # 164
    nz = A.shape[0]
    Z = np.array(cfg["Z_init"], dtype=np.float64) if cfg.get("Z_init") is not None else rng.random((K, nz))  # 174
    H = Z @ A  # 177
    V_hat = reconstruct_from_decomposition(W, H)  # 179
    cost = np.zeros(int(maxiter))  # 181
    ones_nm, ones_mn = np.ones((n, m)), np.ones((m, n))

    with np.errstate(divide="ignore", invalid="ignore"):
        for it in range(1, int(maxiter) + 1):  # 183
            if not W_fixed:  # 185
                if div == "euclidean":  # 187-189
                    neg = V @ H.T + W @ _diagdiag(H @ V_hat.T @ W)
                    pos = V_hat @ H.T + W @ _diagdiag(H @ V.T @ W)
                elif div in ("kl_divergence", "kl"):  # 190-192
                    neg = (V / V_hat) @ H.T + W @ _diagdiag(H @ ones_nm @ W)
                    pos = ones_mn @ H.T + W @ _diagdiag(H @ (V.T / V_hat.T) @ W)
                elif div in ("is_divergence", "is"):  # 193-195
                    neg = (V / V_hat ** 2) @ H.T + W @ _diagdiag(H @ (ones_nm / V_hat.T) @ W)
                    pos = (ones_mn / V_hat) @ H.T + W @ _diagdiag(H @ (V.T / V_hat.T ** 2) @ W)
                elif is_ab:  # 196-203
                    if use_dual:
                        neg = ((V ** (alpha - 1) * V_hat ** beta) @ H.T
                               + W @ _diagdiag(H @ V.T ** (alpha + beta - 1) @ W)) ** (1.0 / beta)
                        pos = (V ** (alpha + beta - 1) @ H.T
                               + W @ _diagdiag(H @ (V ** (alpha - 1) * V_hat ** beta).T @ W)) ** (1.0 / beta)
                    else:
                        neg = ((V ** alpha * V_hat ** (beta - 1)) @ H.T
                               + W @ _diagdiag(H @ V_hat.T ** (alpha + beta - 1) @ W)) ** (1.0 / alpha)
                        pos = (V_hat ** (alpha + beta - 1) @ H.T
                               + W @ _diagdiag(H @ (V ** alpha * V_hat ** (beta - 1)).T @ W)) ** (1.0 / alpha)
                else:  # 204-205
                    raise ReferenceError_(
                        "No update equations defined for cost function with divergence type " + str(div))
                W = W * (neg / np.fmax(pos + lam_w, EPS))  # 207
                W = W @ np.diag(1.0 / np.sqrt(np.sum(W ** 2, axis=0)))  # 208
            V_hat = reconstruct_from_decomposition(W, H)  # 210

            if not Z_fixed:  # 213
                if div == "euclidean":  # 215-217
                    neg = W.T @ V @ A.T
                    pos = W.T @ V_hat @ A.T
                elif div in ("kl_divergence", "kl"):  # 218-220
                    neg = W.T @ (V / V_hat) @ A.T
                    pos = W.T @ ones_mn @ A.T
                elif div in ("is_divergence", "is"):  # 221-223
                    neg = W.T @ (V / V_hat ** 2) @ A.T
                    pos = W.T @ (ones_mn / (W @ H)) @ A.T
                elif is_ab:  # 224-231
                    if use_dual:
                        neg = (W.T @ (V ** (alpha - 1) * V_hat ** beta) @ A.T) ** (1.0 / beta)
                        pos = (W.T @ V ** (alpha + beta - 1) @ A.T) ** (1.0 / beta)
                    else:
                        # constrainednmf.m:229 reads  W' * V.^alpha .* V_hat.^(beta-1) * A'  - MATLAB
                        # evaluates * and .* left to right, so a K x n matrix meets an m x n one:
                        # "Matrix dimensions must agree" unless m == K (a defect of the reference)
                        if m != K:
                            raise ReferenceError_("Matrix dimensions must agree.")
                        neg = ((W.T @ V ** alpha) * V_hat ** (beta - 1) @ A.T) ** (1.0 / alpha)
                        pos = (W.T @ V_hat ** (alpha + beta - 1) @ A.T) ** (1.0 / alpha)
                else:  # 232-233
                    raise ReferenceError_(
                        "No update equations defined for cost function with divergence type " + str(div))
                Z = Z * (neg / np.fmax(pos + lam_z, EPS))  # 235
            H = Z @ A  # 237
            V_hat = reconstruct_from_decomposition(W, H)  # 238

            c = _cost(div, V, V_hat, alpha, beta)  # 241-250
            c = c + lam_w * np.sum(np.abs(W)) + lam_z * np.sum(np.abs(Z))  # 251
            cost[it - 1] = c
            if it > 1 and cost[it - 1] < cost[it - 2] and cost[it - 2] - cost[it - 1] < tol:  # 254-257
                cost = cost[:it]
                break

    # constrainednmf.m:260-267: A (and with it H) back in the original sample order
    A_out = np.zeros_like(A)
    A_out[:, sorted_idx] = A
    H = Z @ A_out
    return W, H, Z, A_out, cost
