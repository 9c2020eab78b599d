"""Host-side mirror of the reference's function-call interface, on top of the C ABI.

The functions keep the reference's names, argument order and meaning:

    W, H, cost = nmf(V, num_basis_elems, config)                 # nmf.m:1
    W, H, cost = cnmf(V, num_basis_elems, context_len, config)   # cnmf.m:1
    W, H, cost = nmfsc(V, num_basis_elems, config)               # nmfsc.m:1
    V_hat      = ReconstructFromDecomposition(W, H)              # ReconstructFromDecomposition.m:1
    v, iters   = projfunc(s, k1, k2, nn)                         # projfunc.m:1

``config`` is a dict with the reference's struct fields (``divergence``,
``W_init``, ``H_init``, ``W_sparsity``, ``H_sparsity``, ``W_fixed``, ``H_fixed``,
``maxiter``, ``tolerance``; nmf.m:17-65).  Arrays are NumPy; MATLAB ``error()``
sites raise :class:`NmfbError`.  All arithmetic happens in ``libnmfb200.so`` on
the GPU - this module only marshals arguments; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib

__all__ = [
    "NmfbError", "Handle", "nmf", "cnmf", "nmfsc", "cnmfsc", "lnmf", "constrainednmf", "ReconstructFromDecomposition",
    "projfunc",
    "default_handle",
]

NMFB_OK = 0
ERR_NAMES = {
    1: "INVALID_ARGUMENT", 2: "CUDA", 3: "UNSUPPORTED", 4: "DIVERGENCE", 5: "AB_ZERO",
    6: "NEGATIVE_DATA", 7: "NO_DATA", 8: "PROJFUNC",
}
DIV_EUCLIDEAN, DIV_KL, DIV_FROBENIUS, DIV_IS, DIV_AB = 0, 1, 2, 3, 4
_DIV_CODES = {
    "euclidean": DIV_EUCLIDEAN,
    "kl_divergence": DIV_KL, "kl": DIV_KL,
    "frobenius": DIV_FROBENIUS,
    "is_divergence": DIV_IS, "is": DIV_IS,
    "ab_divergence": DIV_AB, "ab": DIV_AB,
}
COST_AUTO, COST_DIRECT = 0, 1


class NmfbError(RuntimeError):
    """A MATLAB ``error(...)`` of the reference, or a CUDA failure."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{ERR_NAMES.get(code, code)}] {message}")
        self.code = code
        self.message = message


class _Config(ctypes.Structure):
    _fields_ = [
        ("divergence", ctypes.c_int),
        ("alpha", ctypes.c_double),
        ("beta", ctypes.c_double),
        ("W_init", ctypes.c_void_p),
        ("H_init", ctypes.c_void_p),
        ("W_sparsity", ctypes.c_double),
        ("H_sparsity", ctypes.c_double),
        ("W_fixed", ctypes.c_int),
        ("H_fixed", ctypes.c_int),
        ("maxiter", ctypes.c_int),
        ("tolerance", ctypes.c_double),
        ("seed", ctypes.c_ulonglong),
        ("cost_mode", ctypes.c_int),
        ("W_sparsity_k", ctypes.c_void_p),
        ("H_sparsity_k", ctypes.c_void_p),
        ("W_fixed_k", ctypes.c_void_p),
        ("H_fixed_k", ctypes.c_void_p),
    ]


def _bind(lib):
    if getattr(lib, "_nmfb_bound", False):
        return lib
    P, I, D, LL = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_longlong
    PI = ctypes.POINTER(ctypes.c_int)
    sig = {
        "nmfb_create": ([ctypes.POINTER(P), I], I),
        "nmfb_destroy": ([P], None),
        "nmfb_trim": ([P], ctypes.c_int),
        "nmfb_last_error": ([P], ctypes.c_char_p),
        "nmfb_set_V": ([P, P, I, I], I),
        "nmfb_set_V_device": ([P, P, I, I, LL], I),
        "nmfb_nmf": ([P, I, ctypes.POINTER(_Config), P, P, P, PI], I),
        "nmfb_lnmf": ([P, I, ctypes.POINTER(_Config), P, P, P, PI], I),
        "nmfb_cnmfsc": ([P, I, I, ctypes.POINTER(_Config), P, P, P, PI], I),
        "nmfb_constrainednmf": ([P, I, ctypes.POINTER(_Config), P, I, P, P, P, P, P, PI], I),
        "nmfb_cnmf": ([P, I, I, ctypes.POINTER(_Config), P, P, P, PI], I),
        "nmfb_nmfsc": ([P, I, ctypes.POINTER(_Config), P, P, P, PI], I),
        "nmfb_reconstruct": ([P, P, P, I, I, I, I, P], I),
        "nmfb_projfunc": ([P, P, I, I, D, D, I, P, P], I),
        "nmfb_nmf_begin": ([P, I, ctypes.POINTER(_Config)], I),
        "nmfb_nmf_step": ([P, I], I),
        "nmfb_nmf_sync": ([P, PI, ctypes.POINTER(D)], I),
        "nmfb_nmf_end": ([P, P, P, P, PI], I),
        "nmfb_launch_count": ([P], LL),
        "nmfb_malloc_count": ([P], LL),
        "nmfb_last_loop": ([P, PI, ctypes.POINTER(D)], I),
        "nmfb_last_halvings": ([P, PI, I], I),
        "nmfb_profile_enable": ([P, I], I),
        "nmfb_profile_get": ([P, ctypes.POINTER(D), ctypes.POINTER(D), PI], I),
        "nmfb_profile_get_all": ([P, ctypes.POINTER(D)], I),
        "nmfb_comm_unique_id": ([ctypes.c_char_p], I),
        "nmfb_comm_init": ([P, ctypes.c_char_p, I, I], I),
        "nmfb_version": ([], ctypes.c_char_p),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    lib._nmfb_bound = True
    return lib


def _f32_colmajor(a, shape=None) -> np.ndarray:
    a = np.asarray(a)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise NmfbError(1, f"expected an array of shape {tuple(shape)}, got {tuple(a.shape)}")
    return np.asfortranarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Handle:
    """Owns the device buffers of one GPU (``nmfb_create`` / ``nmfb_destroy``)."""

    def __init__(self, device: int = 0):
        self.lib = _bind(_lib.load())
        h = ctypes.c_void_p()
        rc = self.lib.nmfb_create(ctypes.byref(h), device)
        if rc != NMFB_OK:
            raise NmfbError(rc, self.lib.nmfb_last_error(None).decode())
        self._h = h
        self.device = device
        self.shape = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.nmfb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != NMFB_OK:
            raise NmfbError(rc, self.lib.nmfb_last_error(self._h).decode())

    # -- data
    def trim(self):
        """Release the device blocks cached from earlier calls (``nmfb_trim``)."""
        self._check(self.lib.nmfb_trim(self._h))

    def set_V(self, V):
        V = np.asarray(V)
        if V.ndim != 2:
            raise NmfbError(1, "V must be a matrix")
        Vf = _f32_colmajor(V)
        self._check(self.lib.nmfb_set_V(self._h, _ptr(Vf), V.shape[0], V.shape[1]))
        self.shape = tuple(V.shape)

    def set_V_device(self, ptr: int, m: int, n: int, ld: int):
        self._check(self.lib.nmfb_set_V_device(self._h, ctypes.c_void_p(ptr), m, n, ld))
        self.shape = (m, n)

    # -- multi-GPU
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        self._check(self.lib.nmfb_comm_init(self._h, unique_id, rank, nranks))

    def unique_id(self) -> bytes:
        buf = ctypes.create_string_buffer(128)
        rc = self.lib.nmfb_comm_unique_id(buf)
        if rc != NMFB_OK:
            raise NmfbError(rc, "ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
        return buf.raw

    def launch_count(self) -> int:
        return int(self.lib.nmfb_launch_count(self._h))

    def last_loop(self):
        """(iterations executed, device ms) of the iteration loop of the last one-call algorithm run."""
        it, ms = ctypes.c_int(0), ctypes.c_double(0)
        self._check(self.lib.nmfb_last_loop(self._h, ctypes.byref(it), ctypes.byref(ms)))
        return it.value, ms.value

    def last_halvings(self):
        """nmfsc: (halvings of the H search, halvings of the W search) per iteration of the last call."""
        n = self.lib.nmfb_last_halvings(self._h, None, 0)
        buf = (ctypes.c_int * max(n, 1))()
        self.lib.nmfb_last_halvings(self._h, buf, n)
        a = np.asarray(buf[:n], dtype=np.int32).reshape(-1, 2)
        return a[:, 0].copy(), a[:, 1].copy()

    def malloc_count(self) -> int:
        """cudaMalloc calls made by the handle so far (steady-state calls of one shape add none)."""
        return int(self.lib.nmfb_malloc_count(self._h))

    def profile_enable(self, on: bool = True):
        self._check(self.lib.nmfb_profile_enable(self._h, int(on)))

    def profile_get(self):
        """Average device ms of the W-step and H-step contractions, and how many were timed."""
        a, b, c = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int(0)
        self._check(self.lib.nmfb_profile_get(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def profile_get_all(self):
        """Average device ms of: W-step GEMM, H-step GEMM, gram(H)+cost, element-wise W step, gram(W)."""
        arr = (ctypes.c_double * 5)()
        self._check(self.lib.nmfb_profile_get_all(self._h, arr))
        return dict(zip(["w_gemm", "h_gemm", "gram_h_cost", "w_elementwise", "gram_w"], [float(x) for x in arr]))

    # -- config marshalling
    def _config(self, config, m, n, K, T=None, for_nmfsc=False):
        cfg = dict(config or {})
        c = _Config()
        keep = []
        if not for_nmfsc:
            div = cfg.get("divergence", "euclidean")  # nmf.m:250-252
            if div not in _DIV_CODES:  # nmf.m:165-166
                raise NmfbError(4, "No update equations defined for cost function with divergence type " + str(div))
            c.divergence = _DIV_CODES[div]
            c.alpha = float(cfg.get("alpha", 1))
            c.beta = float(cfg.get("beta", 1))
        W0 = cfg.get("W_init")
        if W0 is not None and np.size(W0) > 0:
            shape = (m, K) if T is None else (m, K, T)
            W0 = _f32_colmajor(W0, shape)
            keep.append(W0)
            c.W_init = W0.ctypes.data
        H0 = cfg.get("H_init")
        if H0 is not None and np.size(H0) > 0:
            H0 = _f32_colmajor(H0, (K, n))
            keep.append(H0)
            c.H_init = H0.ctypes.data
        c.W_sparsity = float(cfg.get("W_sparsity") or 0)
        c.H_sparsity = float(cfg.get("H_sparsity") or 0)
        c.W_fixed = int(bool(cfg.get("W_fixed") or False))
        c.H_fixed = int(bool(cfg.get("H_fixed") or False))
        # per-basis overrides (what a multi-source call with different per-source settings becomes)
        for key, dt in (("W_sparsity_k", np.float64), ("H_sparsity_k", np.float64), ("W_fixed_k", np.int32),
                        ("H_fixed_k", np.int32)):
            v = cfg.get(key)
            if v is not None:
                arr = np.ascontiguousarray(np.asarray(v).astype(dt).ravel())
                if arr.size != K:
                    raise NmfbError(1, f"{key} needs one value per basis ({K}), got {arr.size}")
                keep.append(arr)
                setattr(c, key, arr.ctypes.data)
        mi = cfg.get("maxiter")
        c.maxiter = int(mi) if mi is not None else 0
        tol = cfg.get("tolerance")
        c.tolerance = float(tol) if tol is not None else 0.0
        c.seed = int(cfg.get("seed", 0))
        c.cost_mode = int(cfg.get("cost_mode", COST_AUTO))
        maxiter = c.maxiter if c.maxiter > 0 else 100
        return c, keep, maxiter

    # -- the reference's functions on the V currently held
    def nmf(self, K: int, config=None):
        m, n = self.shape
        c, keep, maxiter = self._config(config, m, n, K)
        W = np.empty((m, K), dtype=np.float32, order="F")
        H = np.empty((K, n), dtype=np.float32, order="F")
        cost = np.zeros(maxiter, dtype=np.float64)
        nc = ctypes.c_int(0)
        self._check(self.lib.nmfb_nmf(self._h, K, ctypes.byref(c), _ptr(W), _ptr(H), _ptr(cost), ctypes.byref(nc)))
        del keep
        return W, H, cost[: nc.value].copy()

    def lnmf(self, K: int, config=None):
        m, n = self.shape
        c, keep, maxiter = self._config(dict(config or {}, divergence="kl"), m, n, K)
        W = np.empty((m, K), dtype=np.float32, order="F")
        H = np.empty((K, n), dtype=np.float32, order="F")
        cost = np.zeros(maxiter, dtype=np.float64)
        nc = ctypes.c_int(0)
        self._check(self.lib.nmfb_lnmf(self._h, K, ctypes.byref(c), _ptr(W), _ptr(H), _ptr(cost), ctypes.byref(nc)))
        del keep
        return W, H, cost[: nc.value].copy()

    def constrainednmf(self, K: int, col2z, nz: int, config=None):
        """nmfb_constrainednmf on the V currently held (columns already in the ordered arrangement).
        Returns W, H (ordered), Z, cost."""
        m, n = self.shape
        cfg = dict(config or {})
        cfg["H_sparsity"] = cfg.pop("Z_sparsity", None)  # the C struct carries them in the H fields
        cfg["H_fixed"] = cfg.pop("Z_fixed", None)
        Z0 = cfg.pop("Z_init", None)
        cfg.pop("H_init", None)
        c, keep, maxiter = self._config(cfg, m, n, K)
        col2z = np.ascontiguousarray(np.asarray(col2z, dtype=np.int32))
        if col2z.size != n:
            raise NmfbError(1, "the column map needs one entry per sample")
        if Z0 is not None and np.size(Z0) > 0:
            Z0 = _f32_colmajor(Z0, (K, nz))
        else:
            Z0 = None
        W = np.empty((m, K), dtype=np.float32, order="F")
        H = np.empty((K, n), dtype=np.float32, order="F")
        Z = np.empty((K, nz), dtype=np.float32, order="F")
        cost = np.zeros(maxiter, dtype=np.float64)
        nc = ctypes.c_int(0)
        self._check(self.lib.nmfb_constrainednmf(self._h, K, ctypes.byref(c), _ptr(col2z), int(nz), _ptr(Z0), _ptr(W),
                                                 _ptr(H), _ptr(Z), _ptr(cost), ctypes.byref(nc)))
        del keep
        return W, H, Z, cost[: nc.value].copy()

    def cnmf(self, K: int, T: int, config=None):
        m, n = self.shape
        c, keep, maxiter = self._config(config, m, n, K, T)
        W = np.empty((m, K, T), dtype=np.float32, order="F")
        H = np.empty((K, n), dtype=np.float32, order="F")
        cost = np.zeros(maxiter, dtype=np.float64)
        nc = ctypes.c_int(0)
        self._check(self.lib.nmfb_cnmf(self._h, K, T, ctypes.byref(c), _ptr(W), _ptr(H), _ptr(cost), ctypes.byref(nc)))
        del keep
        return W, H, cost[: nc.value].copy()

    def nmfsc(self, K: int, config=None):
        m, n = self.shape
        c, keep, maxiter = self._config(config, m, n, K, for_nmfsc=True)
        W = np.empty((m, K), dtype=np.float32, order="F")
        H = np.empty((K, n), dtype=np.float32, order="F")
        cost = np.zeros(maxiter + 1, dtype=np.float64)
        nc = ctypes.c_int(0)
        self._check(self.lib.nmfb_nmfsc(self._h, K, ctypes.byref(c), _ptr(W), _ptr(H), _ptr(cost), ctypes.byref(nc)))
        del keep
        return W, H, cost[: nc.value].copy()

    def cnmfsc(self, K: int, T: int, config=None):
        m, n = self.shape
        c, keep, maxiter = self._config(config, m, n, K, T, for_nmfsc=True)
        W = np.empty((m, K, T), dtype=np.float32, order="F")
        H = np.empty((K, n), dtype=np.float32, order="F")
        cost = np.zeros(maxiter + 1, dtype=np.float64)
        nc = ctypes.c_int(0)
        self._check(self.lib.nmfb_cnmfsc(self._h, K, T, ctypes.byref(c), _ptr(W), _ptr(H), _ptr(cost), ctypes.byref(nc)))
        del keep
        return W, H, cost[: nc.value].copy()

    # -- stepping interface (bench.py)
    def nmf_begin(self, K: int, config=None):
        m, n = self.shape
        c, keep, maxiter = self._config(config, m, n, K)
        self._check(self.lib.nmfb_nmf_begin(self._h, K, ctypes.byref(c)))
        self._session = (m, n, K, maxiter)

    def nmf_step(self, iters: int):
        self._check(self.lib.nmfb_nmf_step(self._h, iters))

    def nmf_sync(self):
        done = ctypes.c_int(0)
        ms = ctypes.c_double(0)
        self._check(self.lib.nmfb_nmf_sync(self._h, ctypes.byref(done), ctypes.byref(ms)))
        return done.value, ms.value

    def nmf_end(self, want_factors=True):
        m, n, K, maxiter = self._session
        W = np.empty((m, K), dtype=np.float32, order="F") if want_factors else None
        H = np.empty((K, n), dtype=np.float32, order="F") if want_factors else None
        cost = np.zeros(maxiter, dtype=np.float64)
        nc = ctypes.c_int(0)
        self._check(self.lib.nmfb_nmf_end(self._h, _ptr(W), _ptr(H), _ptr(cost), ctypes.byref(nc)))
        return W, H, cost[: nc.value].copy()

    def reconstruct(self, W, H):
        W = np.asarray(W)
        H = np.asarray(H)
        if W.ndim == 2:
            m, K = W.shape
            T = 1
        elif W.ndim == 3:
            m, K, T = W.shape
        else:
            raise NmfbError(1, "W must be a matrix or a 3-D tensor")
        if H.ndim != 2 or H.shape[0] != K:
            raise NmfbError(1, "H must be num_basis_elems-by-n")
        n = H.shape[1]
        Wf = _f32_colmajor(W)
        Hf = _f32_colmajor(H)
        out = np.empty((m, n), dtype=np.float32, order="F")
        self._check(self.lib.nmfb_reconstruct(self._h, _ptr(Wf), _ptr(Hf), m, K, T, n, _ptr(out)))
        return out

    def projfunc(self, s, k1, k2, nn=1):
        s = np.asarray(s, dtype=np.float32)
        single = s.ndim == 1
        S = np.ascontiguousarray(s.reshape(1, -1) if single else s)
        count, N = S.shape
        out = np.empty_like(S)
        iters = np.zeros(count, dtype=np.int32)
        self._check(self.lib.nmfb_projfunc(self._h, _ptr(S), N, count, float(k1), float(k2), int(bool(nn)),
                                            _ptr(out), _ptr(iters)))
        if single:
            return out[0], int(iters[0])
        return out, iters


_default: Optional[Handle] = None


def default_handle() -> Handle:
    global _default
    if _default is None:
        _default = Handle(0)
    return _default


def _cat_sources(x, axis):
    return np.concatenate([np.asarray(a) for a in x], axis=axis)


def _split(a, sizes, axis):
    idx = np.cumsum(sizes)[:-1]
    return [np.asfortranarray(p) for p in np.split(a, idx, axis=axis)]


def _multi_source(config, sizes, per_basis=False):
    """nmf.m:11-16, 284-400: cell-array inputs.  The per-source loops of the
    reference (nmf.m:144-171, 175-201) never refresh V_hat between sources, so S
    sources are exactly one factorisation with the bases concatenated; settings
    that differ between sources (sparsity levels, fixed sources: the
    semi-supervised use of nmf.m:51-60) become per-basis vectors
    (``nmfb_config::*_k``) when the engine supports them (``per_basis``: nmf)."""
    cfg = dict(config or {})
    S = len(sizes)
    for key in ("W_sparsity", "H_sparsity", "W_fixed", "H_fixed"):
        v = cfg.get(key)
        if isinstance(v, (list, tuple)):
            if len(v) not in (0, 1, S):  # nmf.m:317-318 etc.
                raise NmfbError(1, f"Requested {S} sources. Given {len(v)} values for {key}.")
            if len(v) == 0:
                cfg[key] = None
            elif len(set(v)) == 1:
                cfg[key] = v[0]
            elif per_basis:
                vals = [max(float(x), 0.0) for x in v] if key.endswith("sparsity") else [int(bool(x)) for x in v]
                cfg[key + "_k"] = np.repeat(np.asarray(vals), sizes)
                cfg[key] = None
            else:
                raise NmfbError(3, f"per-source {key} values are not supported by this function's accelerated path")
    for key, axis in (("W_init", 1), ("H_init", 0)):
        v = cfg.get(key)
        if isinstance(v, (list, tuple)) and len(v) > 0:
            if len(v) != S:  # nmf.m:279-280, 301-302
                raise NmfbError(1, f"Requested {S} sources. Given {len(v)} initial matrices for {key}.")
            cfg[key] = _cat_sources(v, axis)
    return cfg


def nmf(V, num_basis_elems, config=None, handle: Optional[Handle] = None):
    """``[W, H, cost] = nmf(V, num_basis_elems, config)`` (nmf.m:1)."""
    h = handle or default_handle()
    h.set_V(V)
    if isinstance(num_basis_elems, (list, tuple)):
        sizes = [int(k) for k in num_basis_elems]
        cfg = _multi_source(config, sizes, per_basis=True)
        W, H, cost = h.nmf(sum(sizes), cfg)
        if len(sizes) == 1 and not isinstance((config or {}).get("W_init"), (list, tuple)):
            return W, H, cost
        return _split(W, sizes, 1), _split(H, sizes, 0), cost
    return h.nmf(int(num_basis_elems), config)


def lnmf(V, num_basis_elems, config=None, handle: Optional[Handle] = None):
    """``[W, H, cost] = lnmf(V, num_basis_elems, config)`` (lnmf.m:1)."""
    h = handle or default_handle()
    h.set_V(V)
    return h.lnmf(int(num_basis_elems), config)


def cnmf(V, num_basis_elems, context_len, config=None, handle: Optional[Handle] = None):
    """``[W, H, cost] = cnmf(V, num_basis_elems, context_len, config)`` (cnmf.m:1)."""
    h = handle or default_handle()
    h.set_V(V)
    if isinstance(num_basis_elems, (list, tuple)):
        sizes = [int(k) for k in num_basis_elems]
        cfg = _multi_source(config, sizes)
        W, H, cost = h.cnmf(sum(sizes), int(context_len), cfg)
        return _split(W, sizes, 1), _split(H, sizes, 0), cost
    return h.cnmf(int(num_basis_elems), int(context_len), config)


def _label_arrangement(labels):
    """constrainednmf.m:147-170 (host work): labels -> (sorted_idx, col2z of the ordered samples, nz, A ordered).
    Unlabeled samples (label -1) come first and keep a column of Z each; every class shares one."""
    labels = np.asarray(labels).ravel()
    n = labels.size
    num_labeled = int(np.sum(labels > -1))                 # line 149
    uniq, inv = np.unique(labels, return_inverse=True)     # lines 151 / 156
    proc = inv + 1
    if num_labeled < n:
        proc = proc - 1                                    # lines 152-153
        proc[proc == 0] = -1
        num_classes = len(uniq) - 1
    else:
        num_classes = len(uniq)
    sorted_idx = np.argsort(proc, kind="stable")           # line 163
    sorted_labels = proc[sorted_idx]
    n_unl = n - num_labeled
    col2z = np.where(sorted_labels < 0, np.arange(n), n_unl + sorted_labels - 1).astype(np.int32)
    nz = n_unl + num_classes
    A = np.zeros((nz, n), dtype=np.float32)                # lines 166-170
    A[col2z, np.arange(n)] = 1
    return sorted_idx, col2z, nz, A


def constrainednmf(V, labels, num_basis_elems, config=None, handle: Optional[Handle] = None):
    """``[W, H, Z, A, cost] = constrainednmf(V, labels, num_basis_elems, config)`` (constrainednmf.m:1).
    ``config`` fields as in the reference (``W_init``, ``W_sparsity``, ``Z_sparsity``, ``W_fixed``, ``Z_fixed``,
    ``divergence``, ``alpha``, ``beta``, ``maxiter``, ``tolerance``) plus the extension ``Z_init``
    (num_basis_elems x (n_unlabeled + num_classes), ordered arrangement; the reference always draws rand).
    H and A are returned in the original sample order (constrainednmf.m:260-267)."""
    h = handle or default_handle()
    V = np.asarray(V)
    labels = np.asarray(labels).ravel()
    if V.ndim != 2 or labels.size != V.shape[1]:  # constrainednmf.m:98
        raise NmfbError(1, "Length of the label vector not equal to number of samples.")
    sorted_idx, col2z, nz, A_sorted = _label_arrangement(labels)
    h.set_V(V[:, sorted_idx])                     # constrainednmf.m:164
    W, H_sorted, Z, cost = h.constrainednmf(int(num_basis_elems), col2z, nz, config)
    H = np.empty_like(H_sorted)
    H[:, sorted_idx] = H_sorted
    A = np.zeros_like(A_sorted)
    A[:, sorted_idx] = A_sorted
    return W, H, Z, A, cost


def nmfsc(V, num_basis_elems, config=None, handle: Optional[Handle] = None):
    """``[W, H, cost] = nmfsc(V, num_basis_elems, config)`` (nmfsc.m:1)."""
    h = handle or default_handle()
    h.set_V(V)
    return h.nmfsc(int(num_basis_elems), config)


def cnmfsc(V, num_basis_elems, context_len, config=None, handle: Optional[Handle] = None):
    """``[W, H, cost] = cnmfsc(V, num_basis_elems, context_len, config)`` (cnmfsc.m:1)."""
    h = handle or default_handle()
    h.set_V(V)
    return h.cnmfsc(int(num_basis_elems), int(context_len), config)


def ReconstructFromDecomposition(W, H, handle: Optional[Handle] = None):
    """``V_hat = ReconstructFromDecomposition(W, H)`` (ReconstructFromDecomposition.m:1).
    Lists play the role of cell arrays (lines 23-28)."""
    if isinstance(W, (list, tuple)):
        W = _cat_sources(W, 1)
    if isinstance(H, (list, tuple)):
        H = _cat_sources(H, 0)
    return (handle or default_handle()).reconstruct(W, H)


def projfunc(s, k1, k2, nn=1, handle: Optional[Handle] = None):
    """``[v, usediters] = projfunc(s, k1, k2, nn)`` (projfunc.m:1)."""
    return (handle or default_handle()).projfunc(s, k1, k2, nn)
