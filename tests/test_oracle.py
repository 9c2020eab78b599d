"""CPU tests of the oracle itself: golden fixtures, invariants derivable from the
reference code, and the Gram / trace / stacked re-derivation
that the CUDA path implements (oracle/restructured.py) against the literal form.

PARITY UNPINNED: the reference has no tests or golden vectors and cannot run here
(MATLAB); these are the pins we can add.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CASES, inputs, run  # noqa: E402
from oracle import nmf_oracle as O  # noqa: E402
from oracle import restructured as R  # noqa: E402


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    W, H, cost = run(name)
    assert len(cost) == len(g["cost"])
    np.testing.assert_allclose(cost, g["cost"], rtol=1e-9)
    Vhat = O.reconstruct_from_decomposition(W, H)
    np.testing.assert_allclose(np.linalg.norm(Vhat), float(g["vhat_norm"]), rtol=1e-9)
    np.testing.assert_allclose(Vhat[:8, :8], g["vhat_sample"], rtol=1e-7)


def test_projfunc_golden_and_constraints():
    g = np.load(os.path.join(HERE, "golden", "projfunc.npz"))
    rng = np.random.default_rng(99)
    S = rng.random((6, 500))
    for i, sp in enumerate([0.1, 0.3, 0.5, 0.7, 0.9, 0.95]):
        k1 = np.sqrt(500) - (np.sqrt(500) - 1) * sp
        v, it = O.projfunc(S[i], k1, 1.0, 1)
        np.testing.assert_allclose(v, g["v"][i], atol=1e-12)
        assert it == g["iters"][i]
        # projfunc.m:3-7: sum(abs(v)) = k1, sum(v.^2) = k2, v >= 0
        assert abs(v.sum() - k1) < 1e-9 and abs((v ** 2).sum() - 1.0) < 1e-9 and v.min() >= 0
        v2, _ = O.projfunc(v, k1, 1.0, 1)  # idempotent
        np.testing.assert_allclose(v2, v, atol=1e-9)


def test_projfunc_signed():
    rng = np.random.default_rng(5)
    s = rng.standard_normal(300)
    v, _ = O.projfunc(s, 8.0, 1.5, 0)
    assert abs(np.abs(v).sum() - 8.0) < 1e-9 and abs((v ** 2).sum() - 1.5) < 1e-9
    assert np.all(np.sign(v[v != 0]) == np.sign(s[v != 0]))


@pytest.mark.parametrize("div", ["euclidean", "kl"])
def test_exact_fixed_point(div):
    """V = W0*H0 with unit-L2 W0 columns: neg == pos, nothing moves, cost 0 (nmf.m:149-153,180-184)."""
    rng = np.random.default_rng(1)
    W0 = rng.random((40, 5)) + 0.1
    W0 /= np.sqrt((W0 ** 2).sum(0))
    H0 = rng.random((5, 60)) + 0.1
    W, H, cost = O.nmf(W0 @ H0, 5, dict(divergence=div, W_init=W0, H_init=H0, maxiter=5, tolerance=1e-300))
    np.testing.assert_allclose(W, W0, rtol=1e-10)
    np.testing.assert_allclose(H, H0, rtol=1e-10)
    assert np.all(np.abs(cost) < 1e-12)  # rounding noise of sum(V.*log(V./V_hat) - V + V_hat) only


@pytest.mark.parametrize("div", ["euclidean", "kl"])
def test_cnmf_T1_equals_nmf(div):
    rng = np.random.default_rng(2)
    V = rng.random((50, 80)) + 1e-3
    W0 = rng.random((50, 6)) + 1e-3
    W0 /= np.sqrt((W0 ** 2).sum(0))
    H0 = rng.random((6, 80)) + 1e-3
    cfg = dict(divergence=div, W_init=W0, H_init=H0, maxiter=15, tolerance=1e-300)
    W1, H1, c1 = O.nmf(V, 6, cfg)
    W2, H2, c2 = O.cnmf(V, 6, 1, dict(cfg, W_init=W0[:, :, None]))
    np.testing.assert_allclose(c1, c2, rtol=1e-10)
    np.testing.assert_allclose(W1, W2[:, :, 0], rtol=1e-8)
    np.testing.assert_allclose(H1, H2, rtol=1e-8)


def test_nmf_invariants():
    alg, V, K, T, cfg = inputs("nmf_euclid_sparse")
    W, H, cost = O.nmf(V, K, cfg)
    np.testing.assert_allclose((W ** 2).sum(0), 1.0, rtol=1e-12)  # nmf.m:169
    alg, V, K, T, cfg = inputs("nmf_euclid_512")
    _, _, cost = O.nmf(V, K, cfg)
    assert np.all(np.diff(cost) <= 0)  # monotone on dense random V


def test_stop_rule_and_trim():
    """nmf.m:221-224: stop when 0 < cost(i-1) - cost(i) < tolerance; cost trimmed to executed iterations."""
    alg, V, K, T, cfg = inputs("nmf_euclid_512")
    _, _, full = O.nmf(V, K, cfg)
    d = -np.diff(full)
    tol = float(np.sort(d)[len(d) // 2])
    first = int(np.argmax((d > 0) & (d < tol))) + 2  # 1-based iteration that triggers the break
    _, _, c = O.nmf(V, K, dict(cfg, tolerance=tol))
    assert len(c) == first
    np.testing.assert_allclose(c, full[:first], rtol=1e-12)
    # nmf.m:404-411: non-positive maxiter / tolerance are reset to 100 / 1e-3
    _, _, c = O.nmf(V[:40, :40], 3, dict(maxiter=0, tolerance=-1), rng=np.random.default_rng(0))
    assert 2 <= len(c) <= 100


def test_nmfsc_rows_keep_sparseness():
    alg, V, K, T, cfg = inputs("nmfsc_h07")
    W, H, cost = O.nmfsc(V, K, dict(cfg, maxiter=15))
    n = H.shape[1]
    l1 = np.abs(H).sum(1)
    l2 = np.sqrt((H ** 2).sum(1))
    sp = (np.sqrt(n) - l1 / l2) / (np.sqrt(n) - 1)
    np.testing.assert_allclose(sp, 0.7, atol=1e-9)
    np.testing.assert_allclose(l2, 1.0, atol=1e-9)
    assert len(cost) == 16 and np.all(np.diff(cost) <= 1e-12)


def test_errors():
    V = np.ones((4, 4))
    with pytest.raises(ValueError):
        O.nmf(V, 2, dict(divergence="frobenius", maxiter=2), rng=np.random.default_rng(0))  # nmf.m:165-166
    with pytest.raises(ValueError):
        O.nmf(V, 2, dict(divergence="ab", alpha=0, beta=0), rng=np.random.default_rng(0))  # nmf.m:120-122
    with pytest.raises(ValueError):
        O.nmfsc(-V, 2, dict(maxiter=2), rng=np.random.default_rng(0))  # nmfsc.m:57-59


# ---------------------------------------------------------------- restructured forms
@pytest.mark.parametrize("div,lw,lh", [("euclidean", 0, 0), ("euclidean", 0.1, 0.2), ("kl", 0, 0), ("kl", 0.05, 0.1)])
@pytest.mark.parametrize("shards", [1, 3])
def test_gram_form_matches_literal(div, lw, lh, shards):
    rng = np.random.default_rng(3)
    V = rng.random((70, 95)) + 1e-3
    cfg = dict(divergence=div, W_init=rng.random((70, 7)) + 1e-3, H_init=rng.random((7, 95)) + 1e-3,
               W_sparsity=lw, H_sparsity=lh, maxiter=25, tolerance=1e-300)
    W1, H1, c1 = O.nmf(V, 7, cfg)
    W2, H2, c2 = R.nmf_gram(V, 7, cfg, shards=shards)
    np.testing.assert_allclose(c2, c1, rtol=1e-10)
    np.testing.assert_allclose(W2 @ H2, W1 @ H1, rtol=1e-9)


def test_stacked_cnmf_matches_literal():
    rng = np.random.default_rng(4)
    V = rng.random((60, 150)) + 1e-3
    cfg = dict(divergence="euclidean", W_init=rng.random((60, 5, 4)), H_init=rng.random((5, 150)) + 1e-3,
               W_sparsity=0.02, H_sparsity=0.03, maxiter=20, tolerance=1e-300)
    W1, H1, c1 = O.cnmf(V, 5, 4, cfg)
    W2, H2, c2 = R.cnmf_stacked(V, 5, 4, cfg)
    np.testing.assert_allclose(c2, c1, rtol=1e-10)
    np.testing.assert_allclose(W2, W1, rtol=1e-8)
    np.testing.assert_allclose(H2, H1, rtol=1e-8)


def test_sklearn_cross_check_plain_mu():
    """Independent pin for the one textbook update in scope (nmfsc.m:182,232 = Lee-Seung MU)."""
    sk = pytest.importorskip("sklearn.decomposition")
    rng = np.random.default_rng(6)
    V = rng.random((30, 40))
    W0 = rng.random((30, 4)) + 0.1
    H0 = rng.random((4, 40)) + 0.1
    Vn = V / V.max()
    # two plain MU sweeps in nmfsc order (H then W), no sparsity
    W, H, cost = O.nmfsc(V, 4, dict(W_init=W0, H_init=H0, maxiter=3, tolerance=1e-300))
    model = sk.NMF(n_components=4, init="custom", solver="mu", beta_loss="frobenius", max_iter=3, tol=0)
    Wsk = model.fit_transform(Vn, W=W0.copy(), H=H0.copy())
    # sklearn updates W first and does not renormalise; compare the objective only (same order of magnitude
    # and both decreasing) - the two are different schedules of the same multiplicative rule.
    obj_sk = 0.5 * np.linalg.norm(Vn - Wsk @ model.components_) ** 2
    assert cost[-1] < cost[0] and obj_sk < cost[0]
    assert abs(np.log(obj_sk / cost[-1])) < 0.5


def test_lnmf_oracle_invariants():
    """lnmf.m: unit-sum bases (63, 75), KL cost non-increasing on this data, untrimmed cost (88-90)."""
    rng = np.random.default_rng(4)
    V = np.maximum(rng.random((40, 60)), 2.0 ** -24)
    cfg = dict(W_init=rng.random((40, 5)) + 1e-3, H_init=rng.random((5, 60)) + 1e-3, maxiter=50, tolerance=1e-300)
    W, H, c = O.lnmf(V, 5, cfg)
    np.testing.assert_allclose(W.sum(0), 1.0, rtol=1e-12)
    assert len(c) == 50 and np.all(np.diff(c) <= 1e-9 * c[:-1])
    Vh = W @ H
    assert abs(c[-1] - np.sum(V * np.log(V / Vh) - V + Vh)) < 1e-9 * c[-1]
    W2, H2, c2 = O.lnmf(V, 5, dict(cfg, maxiter=400, tolerance=1e-2))
    assert len(c2) == 400 and 1 < np.count_nonzero(c2) < 400


def test_cnmfsc_oracle_behaviour():
    """cnmfsc.m: T == 1 with no sparseness is nmfsc's multiplicative branch; the H line search decreases the
    cost; with W_sparsity > 0 the W line search compares the full model (line 218) with a single-frame trial
    (line 235) and the function returns by step-size underflow with the cost trimmed (lines 245-249)."""
    rng = np.random.default_rng(8)
    m, n, K, T = 30, 90, 3, 3
    V = rng.random((m, n))
    W0, H0 = 0.2 * rng.random((m, K, T)), 0.2 * rng.random((K, n))
    W1, H1, c1 = O.cnmfsc(V, K, 1, dict(W_init=W0[:, :, :1], H_init=H0, maxiter=8, tolerance=1e-300))
    W2, H2, c2 = O.nmfsc(V, K, dict(W_init=W0[:, :, 0], H_init=H0, maxiter=8, tolerance=1e-300))
    np.testing.assert_allclose(c1, c2, rtol=1e-9)
    np.testing.assert_allclose(W1[:, :, 0], W2, rtol=1e-8)
    info = {}
    W, H, c = O.cnmfsc(V, K, T, dict(W_init=W0, H_init=H0, H_sparsity=0.5, maxiter=12, tolerance=1e-300), info=info)
    assert len(c) == 13 and np.all(np.diff(c) <= 0) and len(info["trials_H"]) == 12
    k1 = np.sqrt(n) - (np.sqrt(n) - 1) * 0.5
    np.testing.assert_allclose(np.abs(H).sum(1), k1, rtol=1e-9)  # rows keep the requested L1 at unit L2
    np.testing.assert_allclose((H ** 2).sum(1), 1.0, rtol=1e-9)
    W, H, c = O.cnmfsc(V, K, T, dict(W_init=W0, H_init=H0, W_sparsity=0.5, maxiter=12, tolerance=1e-300))
    assert len(c) <= 2  # "Algorithm converged" (step size below 1e-200) in the first or second iteration
    # with W held fixed the projected W replaces the unprojected W0 at the end of the first iteration (line 266)
    W, H, c = O.cnmfsc(V, K, T, dict(W_init=W0, H_init=H0, W_sparsity=0.5, W_fixed=True, maxiter=6, tolerance=1e-300))
    assert len(c) == 7 and abs((W ** 2).sum(0) - 1).max() < 1e-9


@pytest.mark.parametrize("div,alpha,beta", [("is", 1, 1), ("ab", 0.5, 0.5), ("ab", 2, 1), ("ab", 1.5, -0.5), ("ab", 0, 1)])
def test_two_weight_form_matches_literal_nmf(div, alpha, beta):
    """The device algebra of the IS / AB path (A = Qn H', B = Qp H', Euclidean-shaped W step) is the literal
    nmf.m:154-164,185-195 - checked here on the CPU, incl. the dual updates and sparsity."""
    rng = np.random.default_rng(12)
    m, n, K = 30, 44, 4
    V = 0.5 + rng.random((m, n))
    cfg = dict(divergence=div, alpha=alpha, beta=beta, W_init=rng.random((m, K)) + 0.1, H_init=rng.random((K, n)) + 0.1,
               maxiter=2 if alpha == 0 else 6, tolerance=1e-300, W_sparsity=0.05, H_sparsity=0.02)
    with np.errstate(all="ignore"):
        W, H, c = O.nmf(V, K, cfg)
        W2, H2, c2 = R.nmf_two_weight(V, K, cfg)
    np.testing.assert_allclose(W2, W, rtol=1e-9)
    np.testing.assert_allclose(H2, H, rtol=1e-9)
    fin = np.isfinite(c)
    assert np.array_equal(fin, np.isfinite(c2))
    np.testing.assert_allclose(c2[fin], c[fin], rtol=1e-10)


def test_per_basis_vectors_equal_per_source_loops():
    """Concatenated sources with per-basis lambda / fixed flags == the reference's per-source cell loops."""
    rng = np.random.default_rng(13)
    m, n, sizes = 24, 40, [2, 3, 2]
    V = 0.2 + rng.random((m, n))
    W0 = [rng.random((m, k)) + 0.1 for k in sizes]
    H0 = [rng.random((k, n)) + 0.1 for k in sizes]
    lw, lh = [0.0, 0.2, 0.05], [0.3, 0.0, 0.1]
    cfg = dict(divergence="is", W_init=W0, H_init=H0, maxiter=5, tolerance=1e-300, W_sparsity=lw, H_sparsity=lh,
               W_fixed=[True, False, False], H_fixed=[False, False, True])
    W, H, c = O.nmf(V, sizes, cfg)
    cat = dict(cfg, W_init=np.concatenate(W0, 1), H_init=np.concatenate(H0, 0))
    W2, H2, c2 = R.nmf_two_weight(V, sum(sizes), cat, lam_w=np.repeat(lw, sizes), lam_h=np.repeat(lh, sizes),
                                  fix_w=np.repeat([True, False, False], sizes), fix_h=np.repeat([False, False, True], sizes))
    np.testing.assert_allclose(W2, np.concatenate(W, 1), rtol=1e-9)
    np.testing.assert_allclose(H2, np.concatenate(H, 0), rtol=1e-9)
    np.testing.assert_allclose(c2, c, rtol=1e-10)


@pytest.mark.parametrize("div,alpha,beta", [("kl", 1, 1), ("is", 1, 1), ("ab", 0.5, 1.0), ("ab", 2, 1)])
def test_stacked_two_weight_cnmf_matches_literal(div, alpha, beta):
    rng = np.random.default_rng(14)
    m, n, K, T = 20, 60, 3, 3
    V = 0.2 + rng.random((m, n))
    cfg = dict(divergence=div, alpha=alpha, beta=beta, W_init=rng.random((m, K, T)) + 0.1, H_init=rng.random((K, n)) + 0.1,
               maxiter=5, tolerance=1e-300, W_sparsity=0.03, H_sparsity=0.05)
    W, H, c = O.cnmf(V, K, T, cfg)
    W2, H2, c2 = R.cnmf_two_weight(V, K, T, cfg)
    np.testing.assert_allclose(W2, W, rtol=1e-9)
    np.testing.assert_allclose(H2, H, rtol=1e-9)
    np.testing.assert_allclose(c2, c, rtol=1e-10)


def test_cnmfsc_gram_forms_of_the_w_gradients():
    """pos_t of cnmfsc.m:222,260 through the Gram matrix of the shifted stack, as the device forms it."""
    rng = np.random.default_rng(15)
    m, n, K, T = 18, 50, 3, 4
    V, H = rng.random((m, n)), rng.random((K, n))
    W3 = rng.random((m, K, T))
    Wc = np.concatenate([W3[:, :, t] for t in range(T)], axis=1)
    V_hat = O.reconstruct_from_decomposition(W3, H)
    Wprev = rng.random((m, K))
    for t in range(T):
        Hsh = np.zeros_like(H)
        Hsh[:, t:] = H[:, : n - t]
        neg, pos_mu, pos_sparse = R.cnmfsc_w_gradients(V, Wc, Wprev, H, K, T, t)
        np.testing.assert_allclose(neg, V @ Hsh.T, rtol=1e-12)
        np.testing.assert_allclose(pos_mu, V_hat @ Hsh.T, rtol=1e-11)
        np.testing.assert_allclose(pos_sparse, (Wprev @ H) @ Hsh.T, rtol=1e-11)  # V_hat = RFD(Wnew, H), line 235


# ---------------------------------------------------------------- constrainednmf.m
def test_constrainednmf_all_unlabeled_is_nmf():
    """With every sample unlabeled A = I (constrainednmf.m:170), so V ~ W*Z*A is plain nmf with H = Z: the two
    restatements (written from different files of the reference) must agree to rounding."""
    rng = np.random.default_rng(3)
    m, n, K = 30, 44, 4
    V = rng.random((m, n)) + 0.01
    W0, Z0 = rng.random((m, K)) + 0.1, rng.random((K, n)) + 0.1
    for div in ("euclidean", "kl", "is"):
        cfg = dict(divergence=div, W_init=W0, maxiter=25, tolerance=1e-300, W_sparsity=0.05)
        W, H, Z, A, c = O.constrainednmf(V, -np.ones(n, int), K, dict(cfg, Z_init=Z0, Z_sparsity=0.1))
        Wn, Hn, cn = O.nmf(V, K, dict(cfg, H_init=Z0, H_sparsity=0.1))
        assert np.array_equal(A, np.eye(n))
        np.testing.assert_allclose(c, cn, rtol=1e-12)
        np.testing.assert_allclose(W, Wn, rtol=1e-10)
        np.testing.assert_allclose(H, Hn, rtol=1e-10)
        np.testing.assert_allclose(Z, Hn, rtol=1e-10)


def test_constrainednmf_labels_tie_columns_and_cost_decreases():
    rng = np.random.default_rng(4)
    m, n, K = 25, 60, 3
    V = rng.random((m, n)) + 0.01
    labels = rng.integers(-1, 4, size=n) * 5  # classes 0, 5, 10, 15 and unlabeled (-5 -> treated as a class!)
    labels[labels < 0] = -1
    W, H, Z, A, c = O.constrainednmf(V, labels, K, dict(maxiter=40, tolerance=1e-300, W_init=rng.random((m, K)),
                                                        Z_init=rng.random((K, int((labels < 0).sum()) + 4))))
    assert A.shape == ((labels < 0).sum() + 4, n) and np.all(A.sum(0) == 1)
    np.testing.assert_allclose(H, Z @ A)
    for cls in (0, 5, 10, 15):  # samples of one class share their encoding (constrainednmf.m:237)
        cols = np.flatnonzero(labels == cls)
        assert np.all(H[:, cols] == H[:, cols[:1]])
    assert np.all(np.diff(c) <= 1e-9 * c[:-1])
    np.testing.assert_allclose((W ** 2).sum(0), 1.0, rtol=1e-12)  # constrainednmf.m:208


def test_constrainednmf_reference_defects_reproduced():
    rng = np.random.default_rng(5)
    V = rng.random((12, 20)) + 0.01
    lab = rng.integers(0, 3, size=20)
    with pytest.raises(O.ReferenceError_):  # constrainednmf.m:229: K x n .* m x n
        O.constrainednmf(V, lab, 4, dict(divergence="ab", alpha=0.5, beta=1.0, maxiter=2))
    with pytest.raises(O.ReferenceError_):  # constrainednmf.m:140-142
        O.constrainednmf(V, lab, 4, dict(divergence="ab", alpha=0, beta=0, maxiter=2))
    with pytest.raises(O.ReferenceError_):  # constrainednmf.m:98
        O.constrainednmf(V, lab[:-1], 4, dict(maxiter=2))
    W, H, Z, A, c = O.constrainednmf(V, lab, 4, dict(divergence="ab", alpha=0, beta=1.0, maxiter=5, tolerance=1e-300))
    assert len(c) == 5  # the dual branch (alpha = 0) is well defined; its cost is Inf/NaN (-1/(alpha*beta))


@pytest.mark.parametrize("shards,T,n", [(2, 4, 61), (3, 5, 50), (4, 2, 33), (8, 8, 120), (2, 1, 20)])
def test_cnmf_column_shards_with_halos_match_literal(shards, T, n):
    """The halo scheme of cnmf_driver.cu for several GPUs (T-1 columns of H on both sides, of V on the right;
    own-column partial sums of A = V Hs' and Hs Hs') simulated rank by rank equals the literal cnmf.m loop."""
    from oracle import restructured as R

    rng = np.random.default_rng(shards * 10 + T)
    m, K = 23, 3
    V = rng.random((m, n)) + 0.01
    cfg = dict(divergence="euclidean", W_init=rng.random((m, K, T)) + 0.1, H_init=rng.random((K, n)) + 0.1, maxiter=12,
               tolerance=1e-300, W_sparsity=0.03, H_sparsity=0.07)
    Wo, Ho, co = O.cnmf(V, K, T, cfg)
    Ws, Hs, cs = R.cnmf_stacked_sharded(V, K, T, cfg, shards)
    np.testing.assert_allclose(cs, co, rtol=1e-10)
    np.testing.assert_allclose(Ws, Wo, rtol=1e-9)
    np.testing.assert_allclose(Hs, Ho, rtol=1e-9)
