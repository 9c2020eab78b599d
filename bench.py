#!/usr/bin/env python
"""bench.py - NMF update iterations/sec on BASELINE.json's workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one multiplicative-update iteration (W update, H update, cost) of the reference
function the config names, on synthetic dense V = max(U(0,1), 2^-24):

  --config 2 (default, the north star)  nmf.m  euclidean     V 16384 x 16384, K = 256          (nmf.m:143-225)
  --config 3                            nmf.m  KL divergence V 8192 x 65536,  K = 128          (nmf.m:152-153,183-184)
  --config 4                            cnmf.m euclidean     V 1025 x 20000,  K = 64, T = 8    (cnmf.m:175-258)
  --config 5                            nmfsc.m H_sparsity .7 V 4096 x 4096,  K = 128          (nmfsc.m:141-245)

Configs 2 and 3 shard the columns of V and H over the ranks (strong scaling, W replicated, one packed
exchange per iteration); every rank generates its columns of the SAME V (seeded per global column
block), so the cost curve must not depend on N - it is compared with the committed one-GPU curve.
Configs 4 and 5 are single-GPU algorithms here: with N > 1 every rank runs a replica ("weak").

`value`   : iterations/s with V, W, H resident in HBM, K iterations timed with CUDA events on the
            engine's stream, max over ranks.
`e2e`     : the same metric through the reference-facing call (nmf / cnmf / nmfsc on the C ABI) with
            HOST buffers: upload of V / W_init / H_init, K iterations, download of W, H and the cost
            trace inside the timed region.  One untimed full-size call first, then `--e2e-calls`
            timed calls; the median is reported, every call is listed.
`roofline`: the dominant kernel of the config, timed per launch with CUDA events inside the timed
            region, against MEASURED_PEAKS.json (tf32 = half the bf16 tensor rate; a cuBLAS tf32
            GEMM is also timed in this run and reported beside it).
`cpu_baseline` / --impl reference: oracle/nmf_oracle.py (the literal float64 restatement of the .m
            file; the reference is MATLAB and cannot run here) on this box's host cores, BLAS
            threads set explicitly.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _usable_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use the host cores it
    # reports, so the BLAS thread count is fixed here, before NumPy loads its BLAS.
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(_usable_cores())

METRIC = "nmf_update_iterations_per_sec"
UNIT = "iterations/s"

CONFIGS = {
    2: dict(alg="nmf", divergence="euclidean", m=16384, n=16384, K=256, T=1, sharded=True,
            workload="nmf.m euclidean MU, V=16384x16384 dense uniform, K=256, fp32 storage / tf32 tensor-core operands"),
    3: dict(alg="nmf", divergence="kl", m=8192, n=65536, K=128, T=1, sharded=True,
            workload="nmf.m KL-divergence MU, V=8192x65536 dense uniform, K=128, fp32 storage / tf32 tensor-core operands"),
    4: dict(alg="cnmf", divergence="euclidean", m=1025, n=20000, K=64, T=8, sharded=False,
            workload="cnmf.m euclidean convolutive MU, V=1025x20000 dense uniform, K=64, T=8"),
    5: dict(alg="nmfsc", divergence=None, m=4096, n=4096, K=128, T=1, sharded=False, H_sparsity=0.7,
            workload="nmfsc.m projected-gradient H (H_sparsity=0.7) + multiplicative W, V=4096x4096 dense uniform, K=128"),
    # not BASELINE configs: the widened divergences of SURVEY section 8(f), at the shape VERDICT r1 quotes for them
    6: dict(alg="nmf", divergence="is", m=8192, n=8192, K=128, T=1, sharded=True,
            workload="nmf.m Itakura-Saito MU (nmf.m:154-156,185-187), V=8192x8192 dense uniform, K=128"),
    7: dict(alg="nmf", divergence="ab", alpha=0.5, beta=0.5, m=8192, n=8192, K=128, T=1, sharded=True,
            workload="nmf.m alpha-beta MU (alpha=beta=0.5, nmf.m:161-163,192-194), V=8192x8192 dense uniform, K=128"),
}
V_SEED = 1234
V_BLOCK = 256  # V is generated in blocks of 256 global columns, each with its own seed


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def f_alg(c, m, n, K, T):
    """Algorithmic flops per iteration, SURVEY.md section 8(d)."""
    if c == 3:
        return 8.0 * m * n * K
    if c in (6, 7):  # per half: V_hat and two weighted products
        return 12.0 * m * n * K
    KT = K * T
    return 4.0 * m * n * KT + 4.0 * (m + n) * KT * KT


def measured_peaks(burst: bool):
    """(tf32 TFLOP/s, hbm GB/s, source).  MEASURED_PEAKS.json is driver-written (bf16 burst and
    sustained figures, copy bandwidth); kind::tf32 runs at half the bf16 tensor rate."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    kind = "burst" if burst else "sustained"
    try:
        with open(path) as f:
            p = json.load(f)
        bf = float(p["bf16_tflops"] if burst or "bf16_tflops_sustained" not in p else p["bf16_tflops_sustained"])
        hbm = float(p["hbm_gbs"])
        return bf / 2.0, hbm, f"MEASURED_PEAKS.json bf16_tflops ({kind}) / 2 and hbm_gbs, of measured"
    except Exception:
        return 1590.0 / 2.0, 6650.0, "fallback of B200_PROFILING.md: 1.59 PFLOP/s bf16 / 2, 6650 GB/s"


class ClockSampler:
    """SM clock / power / throttle reasons sampled (NVML, every ~2 ms) while the timed region runs."""

    def __init__(self, index: int):
        self.samples = []
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self._nvml = None

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetCurrentClocksEventReasons(self._h),
                                     nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0))
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self._nvml is None:
            return
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)
        if self._nvml is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self._nvml
        sm = sorted(s[0] for s in self.samples)
        bits = 0
        for s in self.samples:
            bits |= int(s[1])
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = [k for k, v in names.items() if bits & int(v)]
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        pw = sorted(s[2] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": mx, "reasons": reasons,
                "power_w_median": pw[len(pw) // 2], "samples": len(self.samples)}


# --------------------------------------------------------------------------- host inputs
def host_factors(cfg, m, n, K, T):
    """W_init / H_init (float32, the values both arms start from)."""
    import numpy as np

    rng = np.random.default_rng(7)
    if cfg["alg"] == "cnmf":
        W0 = np.asfortranarray(rng.random((m, K, T), dtype=np.float32) + np.float32(1e-3))
    else:
        W0 = np.asfortranarray(np.maximum(rng.random((m, K), dtype=np.float32), 1e-7))
    H0 = np.maximum(rng.random((K, n), dtype=np.float32), 1e-7)
    if cfg["alg"] == "nmfsc":  # nmfsc.m:79-80
        H0 = (H0 / np.sqrt((H0.astype(np.float64) ** 2).sum(1, keepdims=True))).astype(np.float32)
    return W0, H0


def oracle_call(cfg, V, K, T, c):
    from oracle import nmf_oracle as O

    if cfg["alg"] == "nmf":
        return O.nmf(V, K, c)
    if cfg["alg"] == "cnmf":
        return O.cnmf(V, K, T, c)
    return O.nmfsc(V, K, c)


def cpu_reference(config_id, cfg, m, n, K, T, steps, warmup, budget_s):
    """The oracle (literal float64 restatement of the config's .m file) on the host cores.

    One untimed iteration measures the cost of a step; `steps` timed iterations then run at full size
    when they fit `budget_s`, otherwise on the first n/f columns of the workload (per-iteration work is
    linear in n for fixed m, K) with the rate scaled by 1/f.  Returns (value, info dict)."""
    import numpy as np
    from threadpoolctl import threadpool_info, threadpool_limits

    cores = _usable_cores()
    shrink = 8 if config_id == 3 else 1  # the 8192 x 65536 KL oracle needs ~30 GB of float64 temporaries: one shard
    rng = np.random.default_rng(0)
    with threadpool_limits(limits=cores):
        threads = max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
        W0, H0 = host_factors(cfg, m, n, K, T)

        def inputs(f):
            nn = n // f
            V = np.maximum(rng.random((m, nn), dtype=np.float32), np.float32(2.0 ** -24)).astype(np.float64)
            c = dict(W_init=W0.astype(np.float64), H_init=H0[:, :nn].astype(np.float64), tolerance=1e-300)
            if cfg["divergence"]:
                c["divergence"] = cfg["divergence"]
            for key in ("alpha", "beta"):
                if key in cfg:
                    c[key] = cfg[key]
            if cfg.get("H_sparsity"):
                c["H_sparsity"] = cfg["H_sparsity"]
            return V, c

        V, c = inputs(shrink)
        t0 = time.time()
        oracle_call(cfg, V, K, T, dict(c, maxiter=1))
        t1 = time.time() - t0
        log(f"[cpu] one untimed iteration at n/{shrink}: {t1:.2f} s on {threads} BLAS threads")
        f = shrink
        while steps * t1 * shrink / f > budget_s and n // (2 * f) >= 4 * K:
            f *= 2
        if f != shrink:
            V, c = inputs(f)
        t0 = time.time()
        _, _, cost = oracle_call(cfg, V, K, T, dict(c, maxiter=steps))
        dt = time.time() - t0
    done = len(cost) - (1 if cfg["alg"] == "nmfsc" else 0)
    value = done / dt / f
    sample = (f"{done} iterations of oracle.nmf_oracle.{cfg['alg']} (literal float64 {cfg['alg']}.m) on "
              + (f"the full {m}x{n} workload" if f == 1 else
                 f"the first n/{f} = {n // f} columns of the {m}x{n} workload (work per iteration is linear in n); rate scaled by 1/{f}")
              + f"; 1 untimed iteration before, {dt:.1f} s timed")
    log(f"[cpu] {done} iterations in {dt:.2f} s (column fraction 1/{f}) -> {value:.4f} it/s")
    return value, {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}, done


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    m, n, K, T = args.m or cfg["m"], args.n or cfg["n"], args.k or cfg["K"], cfg["T"]
    v, info, done = cpu_reference(args.config, cfg, m, n, K, T, args.steps, args.warmup, 150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": 1, "ms_per_step": 1000.0 / v, "higher_is_better": True,
        "scaling": "strong" if cfg["sharded"] else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "m": m, "n": n, "K": K},
        "cpu_baseline": info,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is MATLAB (no MATLAB/Octave in this image): this arm times the NumPy restatement of the "
                ".m file on the host cores; steps/warmup are the iterations actually run",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- device inputs
def gen_V_columns(torch, dev, m, lo, hi):
    """Columns [lo, hi) of the global synthetic V as a [hi-lo][ld] tensor (= column-major m x (hi-lo) with
    leading dimension ld = m rounded up to 4 floats, the 16-byte pitch the engine's TMA maps need).
    Each block of V_BLOCK global columns has its own Philox seed, so any sharding sees the same V."""
    ld = (m + 3) // 4 * 4
    full = torch.zeros((hi - lo, ld), device=dev, dtype=torch.float32)
    out = full[:, :m]
    g = torch.Generator(device=dev)
    for b in range(lo // V_BLOCK, (hi + V_BLOCK - 1) // V_BLOCK):
        g.manual_seed(V_SEED * 1000003 + b)
        blk = torch.rand((V_BLOCK, m), device=dev, generator=g, dtype=torch.float32)
        s, e = max(lo, b * V_BLOCK), min(hi, (b + 1) * V_BLOCK)
        out[s - lo:e - lo] = blk[s - b * V_BLOCK:e - b * V_BLOCK]
    out.clamp_(min=2.0 ** -24)
    return full, ld


def cublas_tf32_peak(torch, dev):
    """cuBLAS tf32 GEMM 8192^3 on this GPU, in this run: best of 10 (burst) and back to back for ~1.5 s."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        N = 8192
        a = torch.randn((N, N), device=dev)
        b = torch.randn((N, N), device=dev)
        c = torch.empty((N, N), device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(1500.0 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        fl = 2.0 * N ** 3
        return {"burst_tflops": fl / (best * 1e-3) / 1e12, "sustained_tflops": fl * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
                "how": "torch.matmul fp32 with allow_tf32 (cuBLAS), 8192^3, best of 10 / back to back for ~1.5 s, after the timed region"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-calls", type=int, default=5)
    ap.add_argument("--m", type=int, default=0)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--save-cost", default="", help="write the cost trace of this run (JSON) to this path")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from nmf_toolbox_b200 import api
    from nmf_toolbox_b200.distributed import init_comm, shard_bounds

    m, n, K, T = args.m or cfg["m"], args.n or cfg["n"], args.k or cfg["K"], cfg["T"]
    alg = cfg["alg"]
    sharded = cfg["sharded"]
    steps = args.steps
    warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    h = api.Handle(local_rank)
    if sharded:
        init_comm(h, dist, rank, world)
        lo, hi = shard_bounds(n, world, rank)
    else:
        lo, hi = 0, n  # replicas: every rank factors the whole V
    nl = hi - lo

    Vd, ldv = gen_V_columns(torch, dev, m, lo, hi)
    W0, H0full = host_factors(cfg, m, n, K, T)
    H0 = np.asfortranarray(H0full[:, lo:hi])
    total = warmup + steps
    base = dict(W_init=W0, H_init=H0, tolerance=1e-300)
    if cfg["divergence"]:
        base["divergence"] = cfg["divergence"]
    for key in ("alpha", "beta"):
        if key in cfg:
            base[key] = cfg[key]
    if cfg.get("H_sparsity"):
        base["H_sparsity"] = cfg["H_sparsity"]

    def one_call(maxiter):
        c = dict(base, maxiter=maxiter)
        if alg == "nmf":
            return h.nmf(K, c)
        if alg == "cnmf":
            return h.cnmf(K, T, c)
        return h.nmfsc(K, c)

    # ------------------------------------------------------------ resident timing
    h.set_V_device(Vd.data_ptr(), m, nl, ldv)
    sampler = ClockSampler(local_rank)
    if alg == "nmf":
        h.nmf_begin(K, dict(base, maxiter=total + 1))
        h.nmf_step(warmup)
        h.nmf_sync()
        h.profile_enable(True)
        barrier()
        if rank == 0:
            sampler.start()
        launches0 = h.launch_count()
        t0 = time.perf_counter()
        h.nmf_step(steps)
        done, dev_ms = h.nmf_sync()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        launches = h.launch_count() - launches0
        breakdown = h.profile_get_all()
        h.profile_enable(False)
        _, _, cost = h.nmf_end(want_factors=False)
        assert len(cost) == total and np.all(np.isfinite(cost)), "cost trace incomplete"
        timed_iters = steps
    else:
        # one-call algorithms: an untimed call of `warmup` iterations, then ONE call of `steps` iterations whose
        # iteration loop is bracketed by CUDA events on the engine's stream (nmfb_last_loop)
        one_call(warmup)
        barrier()
        if rank == 0:
            sampler.start()
        launches0 = h.launch_count()
        t0 = time.perf_counter()
        _, _, cost = one_call(steps)
        wall_ms = (time.perf_counter() - t0) * 1e3
        timed_iters, dev_ms = h.last_loop()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        launches = h.launch_count() - launches0
        # per-kernel times come from a separate, shorter call with event pairs around the launch groups
        # (the event pairs force direct launches where the timed call replays a CUDA graph)
        h.profile_enable(True)
        one_call(min(steps, 20))
        breakdown = h.profile_get_all()
        h.profile_enable(False)
        assert np.all(np.isfinite(cost)) and timed_iters == steps, ("iteration loop ended early", timed_iters, len(cost))
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    for k2 in list(breakdown):
        breakdown[k2] = max_over_ranks(breakdown[k2])
    replicas = 1 if sharded else world
    value = replicas * timed_iters / (dev_ms / 1e3)
    monotone = bool(np.all(np.diff(cost) <= 1e-6 * np.abs(cost[:-1])))

    # the cost curve must not depend on how many GPUs share the columns: compare with the committed
    # one-GPU curve of the same workload (tests/golden/bench_cost_cfgC_n1.json, written by --save-cost)
    cross = None
    gpath = os.path.join(ROOT, "tests", "golden", f"bench_cost_cfg{args.config}_n1.json")
    if alg == "nmf" and os.path.exists(gpath):
        with open(gpath) as f:
            g = json.load(f)
        if (g["m"], g["n"], g["K"]) == (m, n, K):
            k = min(len(g["cost"]), len(cost))
            gc = np.asarray(g["cost"][:k])
            rel = float(np.max(np.abs(cost[:k] - gc) / np.abs(gc)))
            # The tensor core truncates when it adds into its accumulator (a bias that grows with the chain length) and
            # N shards split the contractions into shorter chains; the Euclidean trace-form cost amplifies that ~7x:
            # measured 1.1e-6 / 1.8e-6 / 5.6e-6 at N = 2 / 4 / 8 (same with the plain all-reduce path), KL 1e-7.
            # The bound is therefore the level at which the one-GPU curve matches the float64 oracle, 1e-5.
            tol = 1e-5 if cfg["divergence"] == "euclidean" else 2e-6
            cross = {"iterations_compared": k, "max_rel_diff_vs_1gpu": rel, "tolerance": tol, "ok": bool(rel < tol)}
    if args.save_cost and rank == 0:
        with open(args.save_cost, "w") as f:
            json.dump({"config": args.config, "m": m, "n": n, "K": K, "n_gpus": world, "cost": [float(x) for x in cost]}, f)

    # ------------------------------------------------------------ end to end (host buffers)
    e2e = None
    if not args.no_e2e:
        Vh = torch.empty((nl, m), dtype=torch.float32, pin_memory=True)
        Vh.copy_(Vd[:, :m])
        torch.cuda.synchronize()
        Vnp = Vh.numpy().T  # m x nl, column-major view of the pinned buffer
        del Vd
        torch.cuda.empty_cache()
        mallocs = []
        runs = []
        for i in range(1 + max(1, args.e2e_calls)):  # call 0 is the untimed warm-up (same shape: fills the block cache)
            barrier()
            t0 = time.perf_counter()
            h.set_V(Vnp)
            W, H, c2 = one_call(steps)
            torch.cuda.synchronize()
            el = time.perf_counter() - t0
            barrier()
            mallocs.append(h.malloc_count())
            if i > 0:
                runs.append(max_over_ranks(el))
        el = sorted(runs)[len(runs) // 2]
        h2d = (Vnp.size + W0.size + H0.size) * 4
        d2h = (W.size + H.size) * 4 + c2.size * 8
        e2e = {"value": replicas * steps / el, "unit": UNIT, "h2d_bytes_per_step": h2d / steps,
               "d2h_bytes_per_step": d2h / steps, "seconds": el, "runs_seconds": runs,
               "spread": (max(runs) - min(runs)) / el, "fraction_of_resident": (replicas * steps / el) / value,
               "cudaMalloc_calls_during_timed_calls": mallocs[-1] - mallocs[0],
               "call": f"Handle.set_V(V_host) + Handle.{alg}(...) == nmfb_set_V + nmfb_{alg} (C ABI), host buffers in and out; "
                       "1 untimed call, then the listed calls, median reported"}
    else:
        del Vd

    # ------------------------------------------------------------ report
    tf32_cublas = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            tf32_cublas = cublas_tf32_peak(torch, dev)
        except Exception as ex:  # pragma: no cover
            tf32_cublas = {"error": str(ex)}
    if rank == 0:
        Kp = (K * T + 31) // 32 * 32
        timed_s = dev_ms / 1e3
        capped = clocks and ("sw_power_cap" in (clocks.get("reasons") or []))
        slow = clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] < 0.97 * clocks["sm_max_mhz"]
        burst = timed_s < 1.0 and not capped and not slow
        peak_tf, hbm_gbs, src = measured_peaks(burst)
        fa = f_alg(args.config, m, n, K, T)
        roof = {"peak_source": src, "peak_kind": "burst" if burst else "sustained",
                "why": "timed region %.3f s, SM clock median %s MHz, power cap %s" % (timed_s, clocks and clocks.get("sm_mhz"), bool(capped)),
                "tf32_cublas_measured_in_run": tf32_cublas, "hbm_peak_gbs": hbm_gbs, "launches_timed": steps}
        if args.config == 2:
            ms = breakdown["h_gemm"]
            fl = 2.0 * m * nl * Kp + 2.0 * nl * Kp * Kp
            roof.update(bound="tensor", kernel="panel_gemm_kernel<EPI_HUPDATE,2> (N = W'V and D = (W'W)H fused with the H update, nmf.m:180-199)",
                        achieved=fl / (ms * 1e-3) / 1e12, peak=peak_tf, unit="TFLOP/s", ms_per_launch=ms,
                        algorithmic_flops_per_launch=fl, algorithmic_bytes_per_launch=4.0 * m * nl,
                        hbm_gbs_of_V_stream=4.0 * m * nl / (ms * 1e-3) / 1e9, w_step_gemm_ms_per_launch=breakdown["w_gemm"])
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                    t = json.load(f)
                if world == 1 and (t.get("m"), t.get("n"), t.get("K")) in ((m, n, K), (None, None, None)):
                    traffic = t.get("h_step_gemm_dram_bytes_per_launch")
                    roof["traffic_source"] = "static: " + str(t.get("source", "profiles/ncu_traffic.json")) + " (an ncu --set full capture of this kernel at this shape, not re-measured in this run)"
            except Exception:
                pass
            roof["traffic"] = traffic
        elif args.config == 3:
            ms_w, ms_h = breakdown["w_gemm"], breakdown["h_gemm"]
            ms = max(ms_w, ms_h)
            fl = 4.0 * m * nl * Kp
            roof.update(bound="tensor", kernel="kl_fused_kernel (S = F G' -> Q = V./S -> OUT += Q G on chip; W half and H half, nmf.m:152-153,183-184)",
                        achieved=fl / (ms * 1e-3) / 1e12, peak=peak_tf, unit="TFLOP/s", ms_per_launch=ms,
                        ms_per_launch_w_half=ms_w, ms_per_launch_h_half=ms_h, algorithmic_flops_per_launch=fl,
                        algorithmic_bytes_per_launch=4.0 * m * nl, hbm_gbs_of_V_stream=4.0 * m * nl / (ms * 1e-3) / 1e9, traffic=None)
        elif args.config in (6, 7):
            ms_w, ms_h = breakdown["w_gemm"], breakdown["h_gemm"]
            ms = max(ms_w, ms_h)
            fl = 6.0 * m * nl * Kp
            roof.update(bound="tensor", kernel="ab_fused_kernel (S = F G' -> Qn, Qp in TMEM -> OUTn += Qn G, OUTp += Qp G; W half and H half, nmf.m:154-164,185-195)",
                        achieved=fl / (ms * 1e-3) / 1e12, peak=peak_tf, unit="TFLOP/s", ms_per_launch=ms,
                        ms_per_launch_w_half=ms_w, ms_per_launch_h_half=ms_h, algorithmic_flops_per_launch=fl,
                        algorithmic_bytes_per_launch=4.0 * m * nl, hbm_gbs_of_V_stream=4.0 * m * nl / (ms * 1e-3) / 1e9, traffic=None)
        elif args.config == 4:
            ms = breakdown["h_gemm"]
            fl = 2.0 * m * n * Kp + 2.0 * n * Kp * Kp
            roof.update(bound="tensor", kernel="panel_gemm_kernel<EPI_STORE,2> split-K + slab sum (P = Wc'V, D = (Wc'Wc)Hs, cnmf.m:216-227)",
                        achieved=fl / (ms * 1e-3) / 1e12, peak=peak_tf, unit="TFLOP/s", ms_per_launch=ms,
                        algorithmic_flops_per_launch=fl, w_step_gemm_ms_per_launch=breakdown["w_gemm"],
                        fold_update_ms_per_launch=breakdown["gram_h_cost"], traffic=None)
        else:
            ms = breakdown["h_gemm"]
            by = 4.0 * m * n
            roof.update(bound="hbm", kernel="resid_fused_kernel (objective 0.5|V - WH|^2 of a line-search trial, split-tf32, nmfsc.m:160-161,237-238)",
                        achieved=by / (ms * 1e-3) / 1e9 if ms > 0 else None, peak=hbm_gbs, unit="GB/s", ms_per_launch=ms,
                        algorithmic_bytes_per_launch=by, algorithmic_flops_per_launch=2.0 * m * n * K, traffic=None)
        if roof.get("traffic") is None and world == 1:  # static ncu figure of the same kernel at the same shape, if recorded
            try:
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                    tc = (json.load(f).get("configs") or {}).get(str(args.config))
                if tc and (tc.get("m"), tc.get("n"), tc.get("K")) == (m, n, K):
                    roof["traffic"] = tc.get("dram_bytes_per_launch")
                    roof["traffic_source"] = "static: %s, %s (not re-measured in this run)" % (tc.get("kernel"), tc.get("source"))
            except Exception:
                pass
        roof["frac"] = (roof["achieved"] / roof["peak"]) if roof.get("achieved") else None
        if tf32_cublas and roof.get("unit") == "TFLOP/s" and "burst_tflops" in tf32_cublas:
            roof["frac_of_cublas_tf32_in_run"] = roof["achieved"] / tf32_cublas["burst_tflops" if burst else "sustained_tflops"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": dev_ms / timed_iters, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "m": m, "n": n, "K": K},
            "config_detail": {"baseline_config": args.config, "T": T, "columns_per_gpu": nl,
                              "l2": "inputs_exceed_l2 (V shard %.0f MiB streamed every iteration)" % (m * nl * 4 / 2 ** 20)
                              if m * nl * 4 > 126 * 2 ** 20 else "V shard %.0f MiB fits the 126 MB L2; all factor / scratch buffers are rewritten every iteration" % (m * nl * 4 / 2 ** 20),
                              "parallelism": ("columns of V and H sharded over %d GPU(s), W replicated" % world) if sharded
                              else ("%d independent replica(s) (this algorithm runs on one GPU)" % world)},
            "wall_ms_per_step": wall_ms / timed_iters,
            "algorithmic_tflops": fa * value / 1e12,
            "tensor_frac_of_step": (fa * value / 1e12) / (peak_tf * world),
            "roofline": roof,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "kernel_ms": breakdown,
            "clocks": clocks,
            "cost_first_last": [float(cost[0]), float(cost[-1])],
            "cost_monotone": monotone,
            "cost_vs_1gpu": cross,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                _, info, _ = cpu_reference(args.config, cfg, m, n, K, T, 5 if args.config in (2, 3, 6, 7) else 10, 1, 25.0)
                line["cpu_baseline"] = info
            except MemoryError as e:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": _usable_cores(), "kind": "port",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()
    if cross is not None and not cross["ok"]:
        log("cost curve differs from the one-GPU curve: %r" % (cross,))
        sys.exit(3)


if __name__ == "__main__":
    main()
