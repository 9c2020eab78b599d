"""Kernel-level checks of panel_gemm (tcgen05 TF32 + TMA) against a float64 matmul
of the same tf32-representable operands, through the library's C debug hooks.

Run as a script (`python tests/test_gpu_panel_gemm.py`) for a verbose report
including a timing of the north-star shape.
"""
import ctypes
import sys
import os

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[1, 2, 4, "2+tail"],
                ids=["cta1", "cta_pair", "two_pairs_multicast", "cta_pair_tail_helpers"])
def cta_group(request):
    """Every kernel-level test runs on the single-CTA kernel, on the CTA-pair (cta_group::2) kernel, on the
    cluster of two pairs that share the Y slab by TMA multicast (shapes it cannot take fall back to pairs) and
    on the pair kernel with tail helpers (two extra CTA pairs that contract the last third / half of every row
    tile's k-blocks and hand the partial sums to the primaries; launches with several column chunks or split-K
    run without them)."""
    tail = request.param == "2+tail"
    os.environ["NMFB_DEBUG_CG"] = "2" if tail else str(request.param)
    if tail:
        os.environ["NMFB_DEBUG_TAIL"] = "2"
    yield request.param
    os.environ.pop("NMFB_DEBUG_CG", None)
    os.environ.pop("NMFB_DEBUG_TAIL", None)


class DebugMat(ctypes.Structure):
    _fields_ = [
        ("base", ctypes.c_void_p),
        ("inner", ctypes.c_longlong),
        ("outer", ctypes.c_longlong),
        ("pitch", ctypes.c_longlong),
        ("mn_major", ctypes.c_int),
    ]


_TEST_LIB = None


def _lib():
    """tests/libnmfb200_test.so: panel_gemm + the kernel-level hooks (built by __graft_entry__.build())."""
    global _TEST_LIB
    if _TEST_LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnmfb200_test.so")
        if not os.path.exists(path):
            raise RuntimeError(path + " not found: run `python __graft_entry__.py` (build())")
        _TEST_LIB = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    return _TEST_LIB


def tf32_round(x):
    """round-to-nearest (ties away) to a 10-bit mantissa, like cvt.rna.tf32.f32."""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def mat(t, mn_major=False):
    """t: 2-D torch tensor whose memory is [outer][pitch] with `inner` valid elements."""
    assert t.stride(1) == 1
    return DebugMat(t.data_ptr(), t.shape[1], t.shape[0], t.stride(0), 1 if mn_major else 0)


def run_store(X0, Y0, kdim0, rows, ncols, X1=None, Y1=None, kdim1=0, splits=1, x0_mn=False, x1_mn=False):
    lib = _lib()
    dev = X0.device
    ldo = (rows + 3) // 4 * 4
    max_splits = 256
    out0 = torch.full((max_splits if splits != 1 else 1, ncols, ldo), float("nan"), device=dev)
    out1 = torch.full((ncols, ldo), float("nan"), device=dev)
    err = ctypes.create_string_buffer(512)
    used = ctypes.c_int(0)
    mx0, my0 = mat(X0, x0_mn), mat(Y0)
    if X1 is not None:
        mx1, my1 = mat(X1, x1_mn), mat(Y1)
        px1, py1 = ctypes.byref(mx1), ctypes.byref(my1)
    else:
        px1 = py1 = None
    rc = lib.nmfb_debug_gemm_store(
        ctypes.byref(mx0), ctypes.byref(my0), ctypes.c_longlong(kdim0), px1, py1, ctypes.c_longlong(kdim1),
        rows, ncols, splits, ctypes.c_void_p(out0.data_ptr()), ctypes.c_void_p(out1.data_ptr()),
        ctypes.c_longlong(ldo), ctypes.c_longlong(ncols * ldo), ctypes.byref(used), err, 512)
    assert rc == 0, err.value.decode()
    o0 = out0[: used.value].sum(0)[:, :rows].T  # rows x ncols
    o1 = out1[:, :rows].T
    return o0, o1, used.value


# fp32 accumulation with promotion every 32 MMA steps: bias <= ~2e-6 (see panel_gemm.cuh)
TOL = 5e-6


def relerr(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def rand(shape, dev, seed):
    """tf32-representable uniform matrix whose row pitch is padded to a multiple of 4 floats (TMA)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    pitch = (shape[1] + 3) // 4 * 4
    buf = torch.zeros((shape[0], pitch))
    buf[:, : shape[1]] = tf32_round(torch.rand(shape, generator=g))
    return buf.to(dev)[:, : shape[1]]


@pytest.mark.parametrize("rows,kdim,ncols", [(128, 32, 32), (128, 256, 256), (300, 1000, 64), (1025, 777, 256), (257, 4100, 512)])
def test_kmajor_single_phase(rows, kdim, ncols):
    dev = torch.device("cuda:0")
    X = rand((rows, kdim), dev, 1)
    Y = rand((ncols, kdim), dev, 2)
    o0, _, _ = run_store(X, Y, kdim, rows, ncols)
    ref = X.double() @ Y.double().T
    e = relerr(o0, ref)
    assert e < TOL, e


@pytest.mark.parametrize("rows,kdim,ncols,kdim1", [(384, 512, 256, 256), (200, 333, 96, 96)])
def test_two_phases(rows, kdim, ncols, kdim1):
    dev = torch.device("cuda:0")
    X0 = rand((rows, kdim), dev, 1)
    Y0 = rand((ncols, kdim), dev, 2)
    X1 = rand((rows, kdim1), dev, 3)
    Y1 = rand((ncols, kdim1), dev, 4)
    o0, o1, _ = run_store(X0, Y0, kdim, rows, ncols, X1, Y1, kdim1)
    e0 = relerr(o0, X0.double() @ Y0.double().T)
    e1 = relerr(o1, X1.double() @ Y1.double().T)
    assert e0 < TOL and e1 < TOL, (e0, e1)


@pytest.mark.parametrize("rows,kdim,ncols", [(128, 32, 32), (300, 1000, 64), (1025, 2000, 256)])
def test_mn_major_x(rows, kdim, ncols):
    """X supplied with the output-row index contiguous (e.g. column-major V for A = V*H')."""
    dev = torch.device("cuda:0")
    X = rand((rows, kdim), dev, 1)
    ldr = (rows + 3) // 4 * 4
    Xt = torch.zeros((kdim, ldr), device=dev)
    Xt[:, :rows] = X.T
    Y = rand((ncols, kdim), dev, 2)
    o0, _, _ = run_store(Xt[:, :rows], Y, kdim, rows, ncols, x0_mn=True)
    e = relerr(o0, X.double() @ Y.double().T)
    assert e < TOL, e


def test_mn_major_second_phase():
    dev = torch.device("cuda:0")
    rows, kdim, ncols, kdim1 = 384, 512, 128, 128
    X0 = rand((rows, kdim), dev, 1)
    Y0 = rand((ncols, kdim), dev, 2)
    X1 = rand((rows, kdim1), dev, 3)
    Y1 = rand((ncols, kdim1), dev, 4)
    X1t = X1.T.contiguous()
    o0, o1, _ = run_store(X0, Y0, kdim, rows, ncols, X1t, Y1, kdim1, x1_mn=True)
    e0 = relerr(o0, X0.double() @ Y0.double().T)
    e1 = relerr(o1, X1.double() @ Y1.double().T)
    assert e0 < TOL and e1 < TOL, (e0, e1)


@pytest.mark.parametrize("rows,kdim,ncols,kdim1,x_mn", [(1024, 2048, 256, 256, True), (2048, 1000, 128, 128, False),
                                                        (512, 4096 + 64, 512, 0, True)])
def test_shapes_of_the_two_pair_multicast_kernel(rows, kdim, ncols, kdim1, x_mn):
    """Even numbers of 256-row tiles and 128 / 256-column slabs: the shapes that run on clusters of two CTA pairs
    sharing the Y slab by TMA multicast (A = V H' with column-major V, the H-step contraction, a second phase)."""
    dev = torch.device("cuda:0")
    X = rand((rows, kdim), dev, 1)
    Y0 = rand((ncols, kdim), dev, 2)
    if x_mn:
        Xt = X.T.contiguous()
        Xarg = Xt
    else:
        Xarg = X
    X1 = rand((rows, kdim1), dev, 3) if kdim1 else None
    Y1 = rand((ncols, kdim1), dev, 4) if kdim1 else None
    o0, o1, _ = run_store(Xarg, Y0, kdim, rows, ncols, X1, Y1, kdim1, x0_mn=x_mn)
    assert relerr(o0, X.double() @ Y0.double().T) < TOL
    if kdim1:
        assert relerr(o1, X1.double() @ Y1.double().T) < TOL


@pytest.mark.parametrize("splits", [0, 3, 7])
def test_split_k(splits):
    dev = torch.device("cuda:0")
    rows, kdim, ncols = 256, 4096 + 40, 256
    X = rand((rows, kdim), dev, 1)
    Y = rand((ncols, kdim), dev, 2)
    o0, _, used = run_store(X, Y, kdim, rows, ncols, splits=splits)
    assert used >= 1
    e = relerr(o0, X.double() @ Y.double().T)
    assert e < TOL, e


def test_fused_h_update():
    lib = _lib()
    dev = torch.device("cuda:0")
    n, m, K = 300, 520, 64  # rows of the panel = samples
    Vt = rand((n, m), dev, 1)          # V column-major: [n][m]
    Wc = rand((K, m), dev, 2)          # W column-major: [K][m]
    G = Wc.double() @ Wc.double().T    # K x K
    G32 = tf32_round(G.float())
    ldh = 304
    Hm = torch.zeros((K, ldh), device=dev)
    Hm[:, :n] = torch.rand((K, n), generator=torch.Generator().manual_seed(5)).to(dev)
    Hr32 = tf32_round(Hm.clone())
    Hc32 = Hr32[:, :n].T.contiguous()   # [n][K]
    H0 = Hm.clone()
    Hr_in = Hr32.clone()
    tiles = (n + 127) // 128
    partials = torch.zeros((tiles, 2), dtype=torch.float64, device=dev)  # kernel accumulates into row 0
    Hc_out = torch.zeros((n, K), device=dev)
    err = ctypes.create_string_buffer(512)
    mx0, my0 = mat(Vt), mat(Wc)
    mx1, my1 = mat(Hr_in[:, :n], True), mat(G32)
    lam = 0.25
    rc = lib.nmfb_debug_gemm_hupdate(
        ctypes.byref(mx0), ctypes.byref(my0), ctypes.c_longlong(m), ctypes.byref(mx1), ctypes.byref(my1),
        ctypes.c_longlong(K), n, K, ctypes.c_void_p(Hm.data_ptr()), ctypes.c_void_p(Hr32.data_ptr()),
        ctypes.c_void_p(Hc_out.data_ptr()), ctypes.c_longlong(ldh), ctypes.c_longlong(K), ctypes.c_float(lam),
        ctypes.c_void_p(partials.data_ptr()), err, 512)
    assert rc == 0, err.value.decode()
    N = (Vt.double() @ Wc.double().T).T              # K x n
    D = G32.double() @ Hr_in[:, :n].double()         # K x n
    eps = 2.0 ** -52
    Href = H0[:, :n].double() * (N / torch.clamp(D + lam, min=eps))
    e = relerr(Hm[:, :n], Href)
    assert e < 1e-5, e
    assert relerr(Hr32[:, :n], Href) < 1e-3
    assert torch.equal(Hr32[:, :n], tf32_round(Hm[:, :n].clone()))
    assert torch.equal(Hc_out, Hr32[:, :n].T)
    p = partials.sum(0)
    assert abs(p[0].item() - (N * Href).sum().item()) / (N * Href).sum().item() < 1e-5
    assert abs(p[1].item() - Href.sum().item()) / Href.sum().item() < 1e-5
    _ = Hc32


def _bench_north_star():
    dev = torch.device("cuda:0")
    m = n = 16384
    K = 256
    g = torch.Generator(device=dev).manual_seed(0)
    V = tf32_round(torch.rand((m, n), device=dev, generator=g))   # row-major m x n
    H = tf32_round(torch.rand((K, n), device=dev, generator=g))
    W = tf32_round(torch.rand((m, K), device=dev, generator=g))   # row-major m x K
    Gm = tf32_round(torch.rand((K, K), device=dev, generator=g))
    for label, kw in [("K-major V", dict()), ("MN-major V", dict(x0_mn=True))]:
        X0 = V if not kw else V  # same memory; for MN-major it is read as V^T (square, so shapes agree)
        for _ in range(2):
            run_store(X0, H, n, m, K, W, Gm, K, **kw)
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        reps = 5
        for _ in range(reps):
            run_store(X0, H, n, m, K, W, Gm, K, **kw)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / reps
        fl = 2.0 * m * n * K + 2.0 * m * K * K
        print(f"[north-star A-GEMM, {label}] {ms:.3f} ms/launch (incl. alloc+sync)  {fl / ms / 1e9:.1f} TFLOP/s  "
              f"{m * n * 4 / ms / 1e6:.0f} GB/s of V")
        o0, o1, _ = run_store(X0, H, n, m, K, W, Gm, K, **kw)
        Xd = V.T if kw else V
        ref = (Xd[:512].double() @ H.double().T)
        print("   relerr rows[:512]:", relerr(o0[:512], ref), " acc1:", relerr(o1[:512], W[:512].double() @ Gm.double().T))


if __name__ == "__main__":
    import traceback

    tests = [
        ("kmajor", lambda: [test_kmajor_single_phase(*a) for a in [(128, 32, 32), (128, 256, 256), (300, 1000, 64), (1025, 777, 256), (257, 4100, 512)]]),
        ("two_phases", lambda: [test_two_phases(*a) for a in [(384, 512, 256, 256), (200, 333, 96, 96)]]),
        ("mn_major", lambda: [test_mn_major_x(*a) for a in [(128, 32, 32), (300, 1000, 64), (1025, 2000, 256)]]),
        ("mn_major_phase1", test_mn_major_second_phase),
        ("split_k", lambda: [test_split_k(s) for s in (0, 3, 7)]),
        ("fused_h_update", test_fused_h_update),
        ("bench", _bench_north_star),
    ]
    for name, fn in tests:
        try:
            fn()
            print(f"PASS {name}", flush=True)
        except Exception:
            print(f"FAIL {name}", flush=True)
            traceback.print_exc()
