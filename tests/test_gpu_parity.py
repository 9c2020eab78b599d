"""GPU parity tests: the CUDA path (through the C ABI, via ctypes) against the CPU
oracle and the committed golden fixtures, on the same seeded inputs.

Stated tolerances (BASELINE.json north_star): cost curve within 1e-4 relative of
the reference over the run, W*H reconstruction within 1e-3 relative (Frobenius).
The kernels multiply tf32-rounded operands with fp32 accumulation; the oracle is
float64."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CASES, inputs  # noqa: E402
from oracle import nmf_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

COST_TOL = 1e-4
RECON_TOL = 1e-3


@pytest.fixture(scope="module")
def api():
    from nmf_toolbox_b200 import api as A

    return A


@pytest.fixture(scope="module")
def handle(api):
    h = api.Handle(0)
    yield h
    h.close()


def recon_err(W, H, Wo, Ho):
    R = O.reconstruct_from_decomposition(np.asarray(W, np.float64), np.asarray(H, np.float64))
    Ro = O.reconstruct_from_decomposition(Wo, Ho)
    return np.linalg.norm(R - Ro) / np.linalg.norm(Ro)


def cost_err(c, co):
    assert len(c) == len(co), (len(c), len(co))
    return float(np.max(np.abs(c - co) / np.maximum(np.abs(co), 1e-300)))


# ---------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", sorted(CASES))
def test_against_golden(api, handle, name):
    alg, V, K, T, cfg = inputs(name)
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    if alg == "nmf":
        W, H, c = api.nmf(V, K, cfg, handle=handle)
    elif alg == "lnmf":
        W, H, c = api.lnmf(V, K, cfg, handle=handle)
    elif alg == "cnmf":
        W, H, c = api.cnmf(V, K, T, cfg, handle=handle)
    elif alg == "cnmfsc":
        W, H, c = api.cnmfsc(V, K, T, cfg, handle=handle)
    else:
        W, H, c = api.nmfsc(V, K, cfg, handle=handle)
    assert cost_err(c, g["cost"]) < COST_TOL
    assert recon_err(W, H, g["W"].astype(np.float64), g["H"].astype(np.float64)) < RECON_TOL


# ---------------------------------------------------------------- nmf
@pytest.mark.parametrize("div", ["euclidean", "kl"])
@pytest.mark.parametrize("m,n,K,iters", [(1024, 768, 32, 200), (257, 1030, 40, 60), (100, 64, 3, 30), (2048, 2048, 64, 100)])
def test_nmf_vs_oracle(api, handle, div, m, n, K, iters):
    rng = np.random.default_rng(m + n)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(divergence=div, W_init=np.maximum(rng.random((m, K)), O.EPS), H_init=np.maximum(rng.random((K, n)), O.EPS),
               maxiter=iters, tolerance=1e-300)
    W, H, c = api.nmf(V, K, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, K, cfg)
    assert cost_err(c, co) < COST_TOL
    assert recon_err(W, H, Wo, Ho) < RECON_TOL
    np.testing.assert_allclose((W.astype(np.float64) ** 2).sum(0), 1.0, rtol=1e-5)  # nmf.m:169


@pytest.mark.parametrize("m,n,K,iters", [(700, 900, 128, 40), (1000, 333, 100, 30), (513, 1025, 96, 30)])
def test_nmf_kl_fused_kernel_shapes(api, handle, m, n, K, iters):
    """KL with K up to 128 runs on the fused kernel (kl_fused.cuh): ragged sizes, several column splits."""
    rng = np.random.default_rng(K + m)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(divergence="kl_divergence", W_init=rng.random((m, K)) + 1e-3, H_init=rng.random((K, n)) + 1e-3,
               W_sparsity=0.05, H_sparsity=0.02, maxiter=iters, tolerance=1e-300)
    W, H, c = api.nmf(V, K, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, K, cfg)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL


def test_nmf_kl_large_K_fallback(api, handle):
    """K > 128 exceeds the fused kernel's tensor-memory plan: the unfused path must give the same answers."""
    rng = np.random.default_rng(77)
    m, n, K = 400, 500, 150
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(divergence="kl", W_init=rng.random((m, K)) + 1e-3, H_init=rng.random((K, n)) + 1e-3, maxiter=25,
               tolerance=1e-300)
    W, H, c = api.nmf(V, K, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, K, cfg)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL


@pytest.mark.parametrize("h_split", ["0", "1"])
def test_nmf_euclid_both_h_paths(api, handle, h_split, monkeypatch):
    """The H step is either fused into the W'V contraction (many sample tiles) or a split-K
    contraction + element-wise update (few tiles, e.g. small column shards): same answers."""
    monkeypatch.setenv("NMFB_H_SPLIT", h_split)
    rng = np.random.default_rng(31)
    m, n, K = 600, 1500, 48
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(divergence="euclidean", W_init=rng.random((m, K)) + 1e-3, H_init=rng.random((K, n)) + 1e-3,
               W_sparsity=0.02, H_sparsity=0.05, maxiter=40, tolerance=1e-300)
    W, H, c = api.nmf(V, K, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, K, cfg)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL


def test_nmf_euclid_tail_helpers(api, handle, monkeypatch):
    """38 / 39 pair tiles per contraction: too many for split-K, too few to fill 148 SMs.  Helper CTA pairs
    contract the last third of every row tile's k-blocks and hand the partial sums to the primaries
    (panel_gemm.cuh, GemmArgs::sk_*); same parity bar, and the run without helpers agrees to rounding."""
    m, n, K, iters = 9600, 9750, 40, 10
    rng = np.random.default_rng(5)
    V = np.maximum(rng.random((m, n), dtype=np.float32), 2.0 ** -24).astype(np.float64)
    cfg = dict(divergence="euclidean", W_init=np.maximum(rng.random((m, K)), O.EPS),
               H_init=np.maximum(rng.random((K, n)), O.EPS), maxiter=iters, tolerance=1e-300)
    W, H, c = api.nmf(V, K, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, K, cfg)
    assert cost_err(c, co) < COST_TOL
    assert recon_err(W, H, Wo, Ho) < RECON_TOL
    monkeypatch.setenv("NMFB_TAIL_HELPERS", "0")
    W1, H1, c1 = api.nmf(V, K, cfg, handle=handle)
    assert cost_err(c, c1) < 1e-5
    assert recon_err(W, H, W1.astype(np.float64), H1.astype(np.float64)) < 1e-4


@pytest.mark.parametrize("m,n,K,h_split", [(700, 900, 300, "1"), (700, 900, 300, "0"), (900, 700, 512, "1"), (64, 50, 70, "1")])
def test_nmf_euclid_more_bases_than_one_column_chunk(api, handle, m, n, K, h_split, monkeypatch):
    """K > 256 spans several 256-wide accumulator chunks (grid.y > 1); K > m, n is legal too."""
    monkeypatch.setenv("NMFB_H_SPLIT", h_split)
    rng = np.random.default_rng(K)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(divergence="euclidean", W_init=rng.random((m, K)) + 1e-3, H_init=rng.random((K, n)) + 1e-3, maxiter=20,
               tolerance=1e-300)
    W, H, c = api.nmf(V, K, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, K, cfg)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL


def test_nmf_direct_cost_mode(api, handle):
    alg, V, K, T, cfg = inputs("nmf_euclid_512")
    Wo, Ho, co = O.nmf(V, K, cfg)
    W, H, c = api.nmf(V, K, dict(cfg, cost_mode=api.COST_DIRECT), handle=handle)
    assert cost_err(c, co) < 2e-5


@pytest.mark.parametrize("div", ["euclidean", "kl"])
def test_nmf_sparsity_terms(api, handle, div):
    rng = np.random.default_rng(8)
    V = np.maximum(rng.random((300, 500)), 2.0 ** -24)
    cfg = dict(divergence=div, W_init=rng.random((300, 24)) + 1e-3, H_init=rng.random((24, 500)) + 1e-3,
               W_sparsity=0.3, H_sparsity=0.7, maxiter=40, tolerance=1e-300)
    W, H, c = api.nmf(V, 24, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, 24, cfg)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL


@pytest.mark.parametrize("div", ["euclidean", "kl"])
@pytest.mark.parametrize("which", ["W_fixed", "H_fixed"])
def test_nmf_fixed_factor(api, handle, div, which):
    """nmf.m:51-60,146,177: a fixed factor keeps its (normalised) initial value."""
    rng = np.random.default_rng(9)
    V = np.maximum(rng.random((200, 260)), 2.0 ** -24)
    W0, H0 = rng.random((200, 12)) + 1e-3, rng.random((12, 260)) + 1e-3
    cfg = dict(divergence=div, W_init=W0, H_init=H0, maxiter=25, tolerance=1e-300)
    cfg[which] = True
    W, H, c = api.nmf(V, 12, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, 12, cfg)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL
    if which == "W_fixed":
        np.testing.assert_allclose(W, W0 / np.sqrt((W0 ** 2).sum(0)), rtol=1e-5)  # nmf.m:133 still applies
    else:
        np.testing.assert_allclose(H, H0, rtol=1e-6)


def test_nmf_stop_rule_and_trim(api, handle):
    """nmf.m:221-224: the device-side stop test ends the loop at the same iteration as the oracle."""
    alg, V, K, T, cfg = inputs("nmf_euclid_512")
    _, _, full = O.nmf(V, K, cfg)
    d = -np.diff(full)
    tol = float(np.sort(d)[len(d) // 2]) * 1.0001
    _, _, co = O.nmf(V, K, dict(cfg, tolerance=tol))
    W, H, c = api.nmf(V, K, dict(cfg, tolerance=tol, cost_mode=api.COST_DIRECT), handle=handle)
    assert 2 <= len(co) < 50 and len(c) == len(co)
    assert cost_err(c, co) < COST_TOL
    # defaults: maxiter <= 0 -> 100, tolerance <= 0 -> 1e-3 (nmf.m:404-411)
    W, H, c = api.nmf(V[:64, :64], 4, dict(maxiter=0, tolerance=-1, seed=3), handle=handle)
    assert 2 <= len(c) <= 100 and np.all(np.isfinite(c))


def test_nmf_stop_rule_default_cost_mode(api, handle):
    """The default Euclidean cost is the Gram / trace identity evaluated from tf32 products (relative accuracy
    ~1e-5, include/nmfb200.h): the stop test of nmf.m:221-224 fires within one iteration of the reference's as
    long as the tolerance is resolvable at that accuracy, and the cost entries agree up to there."""
    alg, V, K, T, cfg = inputs("nmf_euclid_512")
    _, _, full = O.nmf(V, K, cfg)
    d = -np.diff(full)
    tol = float(np.sort(d)[len(d) // 2]) * 1.0001
    assert tol > 1e-5 * full[-1]
    _, _, co = O.nmf(V, K, dict(cfg, tolerance=tol))
    W, H, c = api.nmf(V, K, dict(cfg, tolerance=tol), handle=handle)
    assert 2 <= len(co) < 50 and abs(len(c) - len(co)) <= 1
    k = min(len(c), len(co))
    assert cost_err(c[:k], co[:k]) < COST_TOL


def test_nmf_per_source_settings_many_columns(api, handle):
    """Per-source settings force the split-K H step; with many column tiles its grid exceeds the SM count and
    must NOT wait at the device-side gate for gram(W) (which would need a free SM: deadlock)."""
    m, n, sizes = 512, 40000, [8, 8]
    rng = np.random.default_rng(12)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(divergence="euclidean", W_init=[rng.random((m, k)) + 1e-3 for k in sizes],
               H_init=[rng.random((k, n)) + 1e-3 for k in sizes], W_sparsity=[0.0, 0.1], H_sparsity=[0.2, 0.0],
               H_fixed=[False, True], maxiter=6, tolerance=1e-300)
    W, H, c = api.nmf(V, sizes, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, sizes, cfg)
    assert cost_err(c, co) < COST_TOL
    assert recon_err(np.concatenate(W, 1), np.concatenate(H, 0), np.concatenate(Wo, 1), np.concatenate(Ho, 0)) < RECON_TOL


@pytest.mark.parametrize("div", ["euclidean", "kl"])
def test_nmf_exact_fixed_point(api, handle, div):
    """V = W0*H0 with unit-L2 columns: factors do not move (to tf32 rounding), cost ~ 0."""
    rng = np.random.default_rng(1)
    W0 = rng.random((256, 8)) + 0.1
    W0 /= np.sqrt((W0 ** 2).sum(0))
    H0 = rng.random((8, 320)) + 0.1
    V = W0 @ H0
    W, H, c = api.nmf(V, 8, dict(divergence=div, W_init=W0, H_init=H0, maxiter=5, tolerance=1e-300,
                                 cost_mode=api.COST_DIRECT), handle=handle)
    np.testing.assert_allclose(W, W0, rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(H, H0, rtol=2e-3, atol=1e-6)
    assert np.all(np.abs(c) < 1e-6 * 0.5 * (V ** 2).sum())


def test_nmf_multi_source_cells(api, handle):
    """nmf.m:11-16: cell-array inputs with a common sparsity level == one factorisation of the concatenation."""
    rng = np.random.default_rng(12)
    V = np.maximum(rng.random((150, 200)), 2.0 ** -24)
    W0 = [rng.random((150, 5)) + 1e-3, rng.random((150, 7)) + 1e-3]
    H0 = [rng.random((5, 200)) + 1e-3, rng.random((7, 200)) + 1e-3]
    cfg = dict(W_init=W0, H_init=H0, W_sparsity=[0.1, 0.1], maxiter=20, tolerance=1e-300)
    W, H, c = api.nmf(V, [5, 7], cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, [5, 7], cfg)
    assert isinstance(W, list) and W[0].shape == (150, 5) and H[1].shape == (7, 200)
    assert cost_err(c, co) < COST_TOL
    assert recon_err(np.concatenate(W, 1), np.concatenate(H, 0), np.concatenate(Wo, 1), np.concatenate(Ho, 0)) < RECON_TOL


def test_errors(api, handle):
    V = np.ones((8, 8))
    with pytest.raises(api.NmfbError) as e:  # nmf.m:165-166: nmf has no 'frobenius'
        api.nmf(V, 2, dict(divergence="frobenius"), handle=handle)
    assert e.value.code == 4
    with pytest.raises(api.NmfbError) as e:
        api.nmf(V, 2, dict(divergence="itakura"), handle=handle)
    assert e.value.code == 4
    with pytest.raises(api.NmfbError) as e:  # nmf.m:120-122
        api.nmf(V, 2, dict(divergence="ab", alpha=0, beta=0), handle=handle)
    assert e.value.code == 5
    with pytest.raises(api.NmfbError) as e:  # nmfsc.m:57-59
        api.nmfsc(-V, 2, dict(maxiter=2), handle=handle)
    assert e.value.code == 6 and "Negative values in data!" in str(e.value)
    with pytest.raises(api.NmfbError) as e:  # cnmf.m:133-135
        api.cnmf(V, 2, 2, dict(divergence="ab", alpha=0, beta=0), handle=handle)
    assert e.value.code == 5


# ---------------------------------------------------------------- nmf, IS and AB divergences
@pytest.mark.parametrize("div,alpha,beta,lw,lh", [
    ("is", 1, 1, 0, 0), ("is", 1, 1, 0.05, 0.1),            # nmf.m:154-156,185-187,211-212
    ("ab", 1, 1, 0, 0), ("ab", 0.5, 0.5, 0, 0), ("ab", 2, 1, 0.05, 0.1),  # nmf.m:161-163,192-194,213-214
    ("ab", 0.5, 1.0, 0, 0), ("ab", 1.5, -0.5, 0.02, 0),
])
@pytest.mark.parametrize("m,n,K,iters", [(257, 330, 8, 40), (1024, 768, 32, 60)])
def test_nmf_is_ab_vs_oracle(api, handle, div, alpha, beta, lw, lh, m, n, K, iters):
    rng = np.random.default_rng(m + n)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(divergence=div, alpha=alpha, beta=beta, W_init=np.maximum(rng.random((m, K)), O.EPS),
               H_init=np.maximum(rng.random((K, n)), O.EPS), maxiter=iters, tolerance=1e-300,
               W_sparsity=lw, H_sparsity=lh)
    W, H, c = api.nmf(V, K, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, K, cfg)
    assert cost_err(c, co) < COST_TOL
    assert recon_err(W, H, Wo, Ho) < RECON_TOL
    np.testing.assert_allclose((W.astype(np.float64) ** 2).sum(0), 1.0, rtol=1e-5)  # nmf.m:169


@pytest.mark.parametrize("div,alpha,beta", [("is", 1, 1), ("ab", 0.5, 0.5), ("ab", 2, 1)])
def test_nmf_is_ab_fused_kernel_shapes(api, handle, div, alpha, beta):
    """ab_fused.cuh at a size with several row blocks, several column splits per half and ragged last tiles
    (2100 x 4170, K = 100 -> 128 padded basis columns): V_hat and both weight matrices stay on chip."""
    m, n, K, iters = 2100, 4170, 100, 12
    rng = np.random.default_rng(m + n)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(divergence=div, alpha=alpha, beta=beta, W_init=np.maximum(rng.random((m, K)), O.EPS),
               H_init=np.maximum(rng.random((K, n)), O.EPS), maxiter=iters, tolerance=1e-300, H_sparsity=0.05)
    W, H, c = api.nmf(V, K, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, K, cfg)
    assert cost_err(c, co) < COST_TOL
    assert recon_err(W, H, Wo, Ho) < RECON_TOL


@pytest.mark.parametrize("K,env", [(24, "1"), (160, None)])
def test_nmf_is_ab_unfused_path(api, handle, K, env, monkeypatch):
    """Beyond 128 basis columns (or with NMFB_AB_UNFUSED=1) the two weight matrices are written by the EPI_ABQ
    epilogue and contracted by four panel GEMMs, as in round 1."""
    if env:
        monkeypatch.setenv("NMFB_AB_UNFUSED", env)
    m, n, iters = 700, 900, 20
    rng = np.random.default_rng(K)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    for div, alpha, beta in [("is", 1, 1), ("ab", 0.5, 1.0)]:
        cfg = dict(divergence=div, alpha=alpha, beta=beta, W_init=np.maximum(rng.random((m, K)), O.EPS),
                   H_init=np.maximum(rng.random((K, n)), O.EPS), maxiter=iters, tolerance=1e-300)
        W, H, c = api.nmf(V, K, cfg, handle=handle)
        Wo, Ho, co = O.nmf(V, K, cfg)
        assert cost_err(c, co) < COST_TOL
        assert recon_err(W, H, Wo, Ho) < RECON_TOL


def test_nmf_ab_nonfinite_cost(api, handle):
    """alpha + beta == 0 and alpha * beta == 0 divide by zero in the reference's cost (nmf.m:214): the
    cost entries are Inf / NaN there and here, the factors are still the reference's."""
    m, n, K = 300, 200, 6
    rng = np.random.default_rng(5)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    for alpha, beta in [(1, -1), (1, 0)]:
        cfg = dict(divergence="ab", alpha=alpha, beta=beta, W_init=rng.random((m, K)) + 1e-3,
                   H_init=rng.random((K, n)) + 1e-3, maxiter=20, tolerance=1e-300)
        W, H, c = api.nmf(V, K, cfg, handle=handle)
        with np.errstate(all="ignore"):
            Wo, Ho, co = O.nmf(V, K, cfg)
        assert len(c) == len(co) and not np.isfinite(co).any() and not np.isfinite(c).any()
        assert np.array_equal(np.isnan(c), np.isnan(co))
        assert recon_err(W, H, Wo, Ho) < RECON_TOL


def test_nmf_ab_dual_updates(api, handle):
    """alpha == 0 selects the dual update equations (nmf.m:124-128,159-160,190-191).  On the
    reference they collapse H within a few iterations, so the factors are compared after two."""
    m, n, K = 300, 260, 6
    rng = np.random.default_rng(9)
    V = 0.5 + rng.random((m, n))
    for beta in (1.0, 2.0):
        cfg = dict(divergence="ab", alpha=0, beta=beta, W_init=rng.random((m, K)) + 1e-3,
                   H_init=rng.random((K, n)) + 1e-3, maxiter=2, tolerance=1e-300)
        W, H, c = api.nmf(V, K, cfg, handle=handle)
        with np.errstate(all="ignore"):
            Wo, Ho, co = O.nmf(V, K, cfg)
        assert np.isfinite(Wo).all() and np.isfinite(Ho).all()
        np.testing.assert_allclose(W, Wo, rtol=5e-3, atol=1e-6)
        np.testing.assert_allclose(H, Ho, rtol=5e-3, atol=1e-12)
        assert np.array_equal(np.isfinite(c), np.isfinite(co))


@pytest.mark.parametrize("fixed", ["W", "H"])
def test_nmf_is_fixed_factor(api, handle, fixed):
    m, n, K = 200, 300, 5
    rng = np.random.default_rng(17)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(divergence="is", W_init=rng.random((m, K)) + 1e-3, H_init=rng.random((K, n)) + 1e-3, maxiter=25,
               tolerance=1e-300, W_fixed=fixed == "W", H_fixed=fixed == "H")
    W, H, c = api.nmf(V, K, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, K, cfg)
    assert cost_err(c, co) < COST_TOL
    assert recon_err(W, H, Wo, Ho) < RECON_TOL


# ---------------------------------------------------------------- cnmfsc
@pytest.mark.parametrize("m,n,K,T,sH,iters", [(129, 500, 6, 4, 0.6, 30), (200, 700, 16, 8, 0.7, 20), (120, 400, 5, 3, 0.0, 40),
                                               (300, 640, 40, 5, 0.5, 12), (64, 256, 4, 1, 0.4, 25)])
def test_cnmfsc_vs_oracle(api, handle, m, n, K, T, sH, iters):
    """cnmfsc.m:155-276 with W_sparsity = 0: projected-gradient / multiplicative H step, frame-by-frame W step."""
    rng = np.random.default_rng(m + T)
    V = rng.random((m, n)) * 3.0
    H0 = rng.random((K, n))
    H0 /= np.sqrt((H0 ** 2).sum(1, keepdims=True))
    cfg = dict(W_init=0.3 * rng.random((m, K, T)), H_init=H0, H_sparsity=sH, maxiter=iters, tolerance=1e-300)
    W, H, c = api.cnmfsc(V, K, T, cfg, handle=handle)
    Wo, Ho, co = O.cnmfsc(V, K, T, cfg)
    assert W.shape == (m, K, T) and len(c) == len(co) == iters + 1
    assert cost_err(c, co) < COST_TOL
    assert recon_err(W, H, Wo, Ho) < RECON_TOL
    if sH > 0:
        k1 = np.sqrt(n) - (np.sqrt(n) - 1) * sH
        np.testing.assert_allclose(np.abs(H.astype(np.float64)).sum(1), k1, rtol=1e-4)


def test_cnmfsc_fixed_factors_and_errors(api, handle):
    rng = np.random.default_rng(77)
    m, n, K, T = 100, 300, 5, 3
    V = rng.random((m, n))
    base = dict(W_init=0.3 * rng.random((m, K, T)), H_init=0.3 * rng.random((K, n)), maxiter=12, tolerance=1e-300)
    for extra in (dict(W_fixed=True), dict(W_fixed=True, H_sparsity=0.5), dict(H_fixed=True), dict(W_fixed=True, H_fixed=True)):
        cfg = dict(base, **extra)
        W, H, c = api.cnmfsc(V, K, T, cfg, handle=handle)
        Wo, Ho, co = O.cnmfsc(V, K, T, cfg)
        assert len(c) == len(co), extra
        assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL, extra
    with pytest.raises(api.NmfbError) as e:  # cnmfsc.m:67-69
        api.cnmfsc(-V, K, T, dict(maxiter=2), handle=handle)
    assert e.value.code == 6


@pytest.mark.parametrize("extra", [dict(W_sparsity=0.5), dict(W_sparsity=0.5, H_sparsity=0.6), dict(W_sparsity=0.4, W_fixed=True),
                                   dict(W_sparsity=0.4, W_fixed=True, H_sparsity=0.5), dict(W_sparsity=0.6, H_fixed=True)])
@pytest.mark.parametrize("scale", [1.0, 0.2])
def test_cnmfsc_w_sparsity_literal(api, handle, extra, scale):
    """cnmfsc.m:100-110, 216-254 with W_sparsity > 0, quirks included: W is projected but the loop starts from the
    unprojected W0; the W line search reconstructs its trial from one frame alone (line 235) and usually ends by
    step-size underflow with the cost trimmed (lines 245-249).  Same stopping point, same cost entries, same factors."""
    rng = np.random.default_rng(5)
    m, n, K, T = 60, 150, 4, 3
    V = rng.random((m, n))
    cfg = dict(W_init=scale * rng.random((m, K, T)), H_init=scale * rng.random((K, n)), maxiter=8, tolerance=1e-300, **extra)
    W, H, c = api.cnmfsc(V, K, T, cfg, handle=handle)
    Wo, Ho, co = O.cnmfsc(V, K, T, cfg)
    assert len(c) == len(co), (len(c), len(co))
    assert cost_err(c, co) < COST_TOL
    assert np.linalg.norm(W - Wo) / np.linalg.norm(Wo) < 2e-3
    assert np.linalg.norm(H - Ho) / np.linalg.norm(Ho) < 2e-3


# ---------------------------------------------------------------- lnmf
@pytest.mark.parametrize("m,n,K,iters", [(300, 420, 10, 60), (1024, 768, 32, 40), (129, 1000, 128, 20)])
def test_lnmf_vs_oracle(api, handle, m, n, K, iters):
    """lnmf.m:63-90: unit-sum bases, square-root H step, KL cost."""
    rng = np.random.default_rng(m + K)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(W_init=rng.random((m, K)) + 1e-3, H_init=rng.random((K, n)) + 1e-3, maxiter=iters, tolerance=1e-300)
    W, H, c = api.lnmf(V, K, cfg, handle=handle)
    Wo, Ho, co = O.lnmf(V, K, cfg)
    assert cost_err(c, co) < COST_TOL
    assert recon_err(W, H, Wo, Ho) < RECON_TOL
    np.testing.assert_allclose(W.astype(np.float64).sum(0), 1.0, rtol=1e-5)  # lnmf.m:75


def test_lnmf_stop_leaves_cost_untrimmed(api, handle):
    """lnmf.m:88-90 breaks without trimming: maxiter entries, zeros after the stopping iteration."""
    rng = np.random.default_rng(3)
    V = np.maximum(rng.random((120, 150)), 2.0 ** -24)
    cfg = dict(W_init=rng.random((120, 4)) + 1e-3, H_init=rng.random((4, 150)) + 1e-3, maxiter=300, tolerance=2.0)
    W, H, c = api.lnmf(V, 4, cfg, handle=handle)
    Wo, Ho, co = O.lnmf(V, 4, cfg)
    assert len(c) == len(co) == 300
    assert np.count_nonzero(c) == np.count_nonzero(co) < 300
    k = np.count_nonzero(co)
    assert cost_err(c[:k], co[:k]) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL
    for fixed in ("W_fixed", "H_fixed"):
        cfg2 = dict(cfg, maxiter=15, tolerance=1e-300, **{fixed: True})
        W, H, c = api.lnmf(V, 4, cfg2, handle=handle)
        Wo, Ho, co = O.lnmf(V, 4, cfg2)
        assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL


def test_kl_variants_beyond_the_fused_kernel(api, handle):
    """K > 128 does not fit the fused KL kernel's tensor-memory plan: lnmf (lnmf.m:71-92), per-source settings
    (nmf.m:51-60) and constrainednmf's Z step then run on the unfused contractions with N = W'(V./V_hat) stored."""
    rng = np.random.default_rng(44)
    m, n, K = 300, 420, 136
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(W_init=rng.random((m, K)) + 1e-3, H_init=rng.random((K, n)) + 1e-3, maxiter=12, tolerance=1e-300)
    W, H, c = api.lnmf(V, K, cfg, handle=handle)
    Wo, Ho, co = O.lnmf(V, K, cfg)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL
    sizes = [100, 36]
    cfg2 = dict(divergence="kl", W_init=[cfg["W_init"][:, :100], cfg["W_init"][:, 100:]],
                H_init=[cfg["H_init"][:100], cfg["H_init"][100:]], W_sparsity=[0.0, 0.05], H_sparsity=[0.1, 0.0],
                H_fixed=[False, True], maxiter=12, tolerance=1e-300)
    W, H, c = api.nmf(V, sizes, cfg2, handle=handle)
    Wo, Ho, co = O.nmf(V, sizes, cfg2)
    assert cost_err(c, co) < COST_TOL
    assert recon_err(np.concatenate(W, 1), np.concatenate(H, 0), np.concatenate(Wo, 1), np.concatenate(Ho, 0)) < RECON_TOL
    labels = rng.integers(-1, 6, size=n)
    nz = int((labels < 0).sum()) + 6
    cfg3 = dict(divergence="kl", W_init=cfg["W_init"], Z_init=rng.random((K, nz)) + 1e-3, maxiter=12, tolerance=1e-300)
    W, H, Z, A, c = api.constrainednmf(V, labels, K, cfg3, handle=handle)
    Wo, Ho, Zo, Ao, co = O.constrainednmf(V, labels, K, cfg3)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL


# ---------------------------------------------------------------- constrainednmf
@pytest.mark.parametrize("div,extra", [("euclidean", {}), ("kl", {}), ("is", {}),
                                       ("euclidean", dict(W_sparsity=0.05, Z_sparsity=0.1)),
                                       ("kl", dict(W_sparsity=0.02, Z_sparsity=0.05)),
                                       ("euclidean", dict(Z_fixed=True)), ("kl", dict(W_fixed=True))])
@pytest.mark.parametrize("m,n,K,classes,unl", [(200, 330, 8, 5, 0.3), (129, 1000, 24, 12, 0.0), (300, 257, 6, 3, 0.9)])
def test_constrainednmf_vs_oracle(api, handle, div, extra, m, n, K, classes, unl):
    """constrainednmf.m:183-257 on the nmf kernels: W step as nmf.m, Z step on the class-summed gradients,
    H = Z*A; labels with gaps and unlabeled samples (-1), outputs in the original sample order."""
    rng = np.random.default_rng(m + K)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    labels = rng.integers(0, classes, size=n) * 3 + 1
    labels[rng.random(n) < unl] = -1
    nz = int((labels < 0).sum()) + len(np.unique(labels[labels >= 0]))
    cfg = dict(divergence=div, W_init=rng.random((m, K)) + 1e-3, Z_init=rng.random((K, nz)) + 1e-3, maxiter=40,
               tolerance=1e-300, **extra)
    W, H, Z, A, c = api.constrainednmf(V, labels, K, cfg, handle=handle)
    Wo, Ho, Zo, Ao, co = O.constrainednmf(V, labels, K, cfg)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL
    assert np.array_equal(A, Ao)
    np.testing.assert_allclose(H, Z.astype(np.float64) @ A, rtol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(Z - Zo) / np.linalg.norm(Zo), 0, atol=5e-3)
    if extra.get("Z_fixed"):
        np.testing.assert_allclose(Z, cfg["Z_init"], rtol=2e-6)


def test_constrainednmf_ab_dual_and_errors(api, handle):
    rng = np.random.default_rng(9)
    m, n, K = 120, 150, 5
    V = 0.5 + rng.random((m, n))
    labels = rng.integers(-1, 3, size=n)
    nz = int((labels < 0).sum()) + 3
    # dual updates (constrainednmf.m:128-132); as in nmf.m they collapse the encoding within a few iterations on
    # the reference, so the factors are compared after two
    cfg = dict(divergence="ab", alpha=0, beta=1.0, W_init=rng.random((m, K)) + 1e-3, Z_init=rng.random((K, nz)) + 1e-3,
               maxiter=2, tolerance=1e-300)
    W, H, Z, A, c = api.constrainednmf(V, labels, K, cfg, handle=handle)
    with np.errstate(all="ignore"):
        Wo, Ho, Zo, Ao, co = O.constrainednmf(V, labels, K, cfg)
    assert np.isfinite(Wo).all() and np.isfinite(Zo).all()
    np.testing.assert_allclose(W, Wo, rtol=5e-3, atol=1e-6)
    np.testing.assert_allclose(Z, Zo, rtol=5e-3, atol=1e-12)
    assert np.array_equal(np.isfinite(c), np.isfinite(co))  # -1/(alpha*beta) with alpha = 0 (constrainednmf.m:249)
    with pytest.raises(api.NmfbError) as e:  # constrainednmf.m:229 cannot be evaluated unless m == K
        api.constrainednmf(V, labels, K, dict(cfg, alpha=0.5), handle=handle)
    assert e.value.code == 3
    with pytest.raises(api.NmfbError) as e:  # constrainednmf.m:140-142
        api.constrainednmf(V, labels, K, dict(cfg, alpha=0, beta=0), handle=handle)
    assert e.value.code == 5
    with pytest.raises(api.NmfbError):  # constrainednmf.m:98
        api.constrainednmf(V, labels[:-1], K, cfg, handle=handle)


# ---------------------------------------------------------------- nmf, multi-source cell arrays
@pytest.mark.parametrize("div", ["euclidean", "kl", "is"])
@pytest.mark.parametrize("case", ["sparsity", "w_fixed", "h_fixed", "mixed"])
def test_nmf_multi_source_per_source_settings(api, handle, div, case):
    """nmf.m:11-16,51-60,144-201: sources with different sparsity levels / held fixed (the
    semi-supervised use: a pre-trained dictionary stays fixed while another one is learned)."""
    m, n, sizes = 300, 420, [6, 10, 4]
    rng = np.random.default_rng(31)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    W0 = [rng.random((m, k)) + 1e-3 for k in sizes]
    H0 = [rng.random((k, n)) + 1e-3 for k in sizes]
    cfg = dict(divergence=div, W_init=W0, H_init=H0, maxiter=30, tolerance=1e-300)
    if case in ("sparsity", "mixed"):
        cfg.update(W_sparsity=[0.0, 0.2, 0.05], H_sparsity=[0.3, 0.0, 0.1])
    if case in ("w_fixed", "mixed"):
        cfg.update(W_fixed=[True, False, False])
    if case in ("h_fixed", "mixed"):
        cfg.update(H_fixed=[False, False, True])
    W, H, c = api.nmf(V, sizes, cfg, handle=handle)
    Wo, Ho, co = O.nmf(V, sizes, cfg)
    assert isinstance(W, list) and [w.shape[1] for w in W] == sizes and [x.shape[0] for x in H] == sizes
    assert cost_err(c, co) < COST_TOL
    assert recon_err(np.concatenate(W, 1), np.concatenate(H, 0), np.concatenate(Wo, 1), np.concatenate(Ho, 0)) < RECON_TOL
    if case in ("w_fixed", "mixed"):  # the fixed dictionary is returned as normalised at the start (nmf.m:132)
        Wn = W0[0] / np.sqrt((W0[0] ** 2).sum(0))
        np.testing.assert_allclose(W[0], Wn, rtol=2e-6)
    if case in ("h_fixed", "mixed"):
        np.testing.assert_allclose(H[2], H0[2], rtol=2e-6)


# ---------------------------------------------------------------- cnmf
@pytest.mark.parametrize("m,n,K,T,iters,lw,lh", [(129, 700, 8, 4, 60, 0, 0), (200, 1000, 16, 5, 40, 0.05, 0.1),
                                                  (1025, 2000, 64, 8, 30, 0, 0), (64, 90, 3, 7, 25, 0, 0)])
def test_cnmf_vs_oracle(api, handle, m, n, K, T, iters, lw, lh):
    rng = np.random.default_rng(m)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(divergence="euclidean", W_init=rng.random((m, K, T)), H_init=np.maximum(rng.random((K, n)), O.EPS),
               W_sparsity=lw, H_sparsity=lh, maxiter=iters, tolerance=1e-300)
    W, H, c = api.cnmf(V, K, T, cfg, handle=handle)
    Wo, Ho, co = O.cnmf(V, K, T, cfg)
    assert W.shape == (m, K, T)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL


@pytest.mark.parametrize("div,alpha,beta", [("kl", 1, 1), ("is", 1, 1), ("ab", 0.5, 0.5), ("ab", 2, 1), ("ab", 1, 1),
                                            ("kl_divergence", 1, 1)])
@pytest.mark.parametrize("m,n,K,T,iters,lw,lh", [(129, 700, 8, 4, 40, 0, 0), (200, 500, 6, 5, 30, 0.05, 0.1),
                                                  (513, 1000, 64, 8, 15, 0, 0)])
def test_cnmf_kl_is_ab_vs_oracle(api, handle, div, alpha, beta, m, n, K, T, iters, lw, lh):
    """cnmf.m:137-147,177-233: KL / IS / AB as one alpha-beta family (with the unshifted V_pos of the
    KL branch, cnmf.m:221-222)."""
    rng = np.random.default_rng(m + T)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    cfg = dict(divergence=div, alpha=alpha, beta=beta, W_init=rng.random((m, K, T)) + 1e-3,
               H_init=np.maximum(rng.random((K, n)), O.EPS), W_sparsity=lw, H_sparsity=lh, maxiter=iters,
               tolerance=1e-300)
    W, H, c = api.cnmf(V, K, T, cfg, handle=handle)
    Wo, Ho, co = O.cnmf(V, K, T, cfg)
    assert W.shape == (m, K, T)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL


def test_cnmf_T1_equals_nmf(api, handle):
    rng = np.random.default_rng(2)
    V = np.maximum(rng.random((300, 400)), 2.0 ** -24)
    W0 = rng.random((300, 10)) + 1e-3
    W0 /= np.sqrt((W0 ** 2).sum(0))
    H0 = rng.random((10, 400)) + 1e-3
    cfg = dict(divergence="euclidean", W_init=W0, H_init=H0, maxiter=20, tolerance=1e-300)
    W1, H1, c1 = api.nmf(V, 10, cfg, handle=handle)
    W2, H2, c2 = api.cnmf(V, 10, 1, dict(cfg, W_init=W0[:, :, None]), handle=handle)
    assert cost_err(c2, c1) < 2e-5
    np.testing.assert_allclose(W2[:, :, 0], W1, rtol=2e-3, atol=1e-6)


# ---------------------------------------------------------------- nmfsc / projfunc
@pytest.mark.parametrize("sW,sH,iters", [(None, 0.7, 100), (None, None, 40), (0.5, 0.5, 40), (0.6, None, 10)])
def test_nmfsc_vs_oracle(api, handle, sW, sH, iters):
    rng = np.random.default_rng(2)
    m, n, K = 512, 512, 16
    V = rng.random((m, n)) * 3.0
    H0 = rng.random((K, n))
    H0 /= np.sqrt((H0 ** 2).sum(1, keepdims=True))
    cfg = dict(W_init=rng.random((m, K)), H_init=H0, W_sparsity=sW, H_sparsity=sH, maxiter=iters, tolerance=1e-300)
    W, H, c = api.nmfsc(V, K, cfg, handle=handle)
    info = {}
    Wo, Ho, co = O.nmfsc(V, K, cfg, info=info)
    assert cost_err(c, co) < COST_TOL
    # the line searches run on the device (no host decision): same accept / halve sequence as the reference
    hH, hW = handle.last_halvings()
    if sH:
        assert hH.tolist() == list(info["halvings_H"])
    if sW:
        assert hW.tolist() == list(info["halvings_W"])
    if sH:
        Hd = H.astype(np.float64)
        l1, l2 = np.abs(Hd).sum(1), np.sqrt((Hd ** 2).sum(1))
        np.testing.assert_allclose((np.sqrt(n) - l1 / l2) / (np.sqrt(n) - 1), sH, atol=1e-4)  # rows keep sparseness
        np.testing.assert_allclose(l2, 1.0, atol=1e-4)


@pytest.mark.parametrize("slots", ["1", "3"])
def test_nmfsc_line_search_slots(api, handle, slots, monkeypatch):
    """The number of trial slots queued per iteration only changes how far a long search spills into the
    following kernel patterns, never the result (first iterations need up to ~10 halvings)."""
    monkeypatch.setenv("NMFB_LS_SLOTS", slots)
    rng = np.random.default_rng(5)
    m, n, K = 300, 400, 12
    V = rng.random((m, n)) * 2.0
    H0 = rng.random((K, n))
    H0 /= np.sqrt((H0 ** 2).sum(1, keepdims=True))
    cfg = dict(W_init=rng.random((m, K)), H_init=H0, W_sparsity=0.4, H_sparsity=0.6, maxiter=25, tolerance=1e-300)
    W, H, c = api.nmfsc(V, K, cfg, handle=handle)
    info = {}
    Wo, Ho, co = O.nmfsc(V, K, cfg, info=info)
    assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL
    hH, hW = handle.last_halvings()
    assert hH.tolist() == list(info["halvings_H"]) and hW.tolist() == list(info["halvings_W"])


def test_nmfsc_stop_rule_and_fixed_factors(api, handle):
    """nmfsc.m:241-244 on the device: the loop ends at the reference's iteration, cost keeps iter+1 entries;
    W_fixed / H_fixed skip the corresponding half (nmfsc.m:143,192)."""
    rng = np.random.default_rng(8)
    m, n, K = 256, 320, 8
    V = rng.random((m, n)) * 2.0
    H0 = rng.random((K, n))
    H0 /= np.sqrt((H0 ** 2).sum(1, keepdims=True))
    cfg = dict(W_init=rng.random((m, K)), H_init=H0, H_sparsity=0.5, maxiter=60, tolerance=1e-300)
    _, _, full = O.nmfsc(V, K, cfg)
    d = -np.diff(full)[1:]
    tol = float(np.sort(d)[len(d) // 2]) * 1.0001
    Wo, Ho, co = O.nmfsc(V, K, dict(cfg, tolerance=tol))
    W, H, c = api.nmfsc(V, K, dict(cfg, tolerance=tol), handle=handle)
    assert 3 <= len(co) < 61 and abs(len(c) - len(co)) <= 1
    k = min(len(c), len(co))
    assert cost_err(c[:k], co[:k]) < COST_TOL
    for fixed in ("W_fixed", "H_fixed"):
        cfg2 = dict(cfg, maxiter=12, **{fixed: True})
        W, H, c = api.nmfsc(V, K, cfg2, handle=handle)
        Wo, Ho, co = O.nmfsc(V, K, cfg2)
        assert cost_err(c, co) < COST_TOL and recon_err(W, H, Wo, Ho) < RECON_TOL


@pytest.mark.parametrize("N,sp", [(100, 0.7), (4096, 0.7), (1000, 0.3), (5000, 0.95), (20000, 0.5), (33, 0.9)])
def test_projfunc_vs_oracle(api, handle, N, sp):
    rng = np.random.default_rng(N)
    s = rng.random(N)
    k1 = np.sqrt(N) - (np.sqrt(N) - 1) * sp
    v, it = api.projfunc(s, k1, 1.0, 1, handle=handle)
    vo, ito = O.projfunc(s, k1, 1.0, 1)
    assert it == ito
    np.testing.assert_allclose(v, vo, atol=2e-5)
    assert abs(v.astype(np.float64).sum() - k1) < 1e-3 * k1 and abs((v.astype(np.float64) ** 2).sum() - 1) < 1e-4
    assert v.min() >= 0


def test_projfunc_signed_and_batched(api, handle):
    g = np.load(os.path.join(HERE, "golden", "projfunc.npz"))
    rng = np.random.default_rng(99)
    S = rng.random((6, 500))
    for i, sp in enumerate([0.1, 0.3, 0.5, 0.7, 0.9, 0.95]):
        k1 = np.sqrt(500) - (np.sqrt(500) - 1) * sp
        v, it = api.projfunc(S[i], k1, 1.0, 1, handle=handle)
        assert it == g["iters"][i]
        np.testing.assert_allclose(v, g["v"][i], atol=2e-5)
    s = np.random.default_rng(5).standard_normal((4, 777))
    V, its = api.projfunc(s, 10.0, 2.0, 0, handle=handle)
    for i in range(4):
        vo, ito = O.projfunc(s[i], 10.0, 2.0, 0)
        assert its[i] == ito
        np.testing.assert_allclose(V[i], vo, atol=2e-5)


# ---------------------------------------------------------------- ReconstructFromDecomposition
def test_reconstruct(api, handle):
    rng = np.random.default_rng(3)
    W, H = rng.random((300, 24)), rng.random((24, 501))
    R = api.ReconstructFromDecomposition(W, H, handle=handle)
    np.testing.assert_allclose(R, W @ H, rtol=5e-6)
    W3, H3 = rng.random((129, 8, 5)), rng.random((8, 333))
    R = api.ReconstructFromDecomposition(W3, H3, handle=handle)
    np.testing.assert_allclose(R, O.reconstruct_from_decomposition(W3, H3), rtol=5e-6, atol=1e-6)
    R = api.ReconstructFromDecomposition([W[:, :10], W[:, 10:]], [H[:10], H[10:]], handle=handle)  # RFD.m:23-28
    np.testing.assert_allclose(R, W @ H, rtol=5e-6)
