"""world_size-2 gloo test of the column-sharded algorithm, on CPU.

Each rank holds a column shard of V and H and the full W and runs the Gram-form
iteration with ONE all-reduce per iteration of the packed payload
[A_r = V_r H_r' | G_H,r = H_r H_r' | hs_r | cost partials] - the same exchange the
CUDA engine performs through NCCL.  The result must match the single-rank oracle
up to summation order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EPS = 2.0 ** -52


def _sharded_nmf(rank, world, V, W0, H0, div, iters, lw, lh):
    """NumPy statement of what one rank of the engine does (csrc/nmf_driver.cu)."""
    from nmf_toolbox_b200.distributed import shard_bounds

    lo, hi = shard_bounds(V.shape[1], world, rank)
    Vr, H = V[:, lo:hi].copy(), H0[:, lo:hi].copy()
    W = W0 / np.sqrt((W0 ** 2).sum(0))
    m, K = W.shape

    def allreduce(*arrs):
        flat = torch.from_numpy(np.concatenate([a.ravel() for a in arrs]))
        dist.all_reduce(flat)
        out, o = [], 0
        for a in arrs:
            out.append(flat[o:o + a.size].numpy().reshape(a.shape).copy())
            o += a.size
        return out

    (vsq, vsum, vlogv) = allreduce(np.array([(Vr ** 2).sum()]), np.array([Vr.sum()]), np.array([(Vr * np.log(Vr)).sum()]))
    cost = []
    for _ in range(iters):
        if div == "euclidean":
            A, GH = allreduce(Vr @ H.T, H @ H.T)
            B = W @ GH
            a, b = (W * A).sum(0), (W * B).sum(0)
            W = W * ((A + W * b) / np.maximum(B + W * a + lw, EPS))
            W = W / np.sqrt((W ** 2).sum(0))
            GW = W.T @ W
            N = W.T @ Vr
            H = H * (N / np.maximum(GW @ H + lh, EPS))
            nh, sh, GH2 = allreduce(np.array([(N * H).sum()]), np.array([H.sum()]), H @ H.T)
            cost.append(0.5 * (vsq[0] - 2 * nh[0] + (GW * GH2).sum()) + lw * W.sum() + lh * sh[0])
        else:
            Q = Vr / (W @ H)
            R, hs = allreduce(Q @ H.T, H.sum(1))
            ws = W.sum(0)
            c = (W * R).sum(0)
            W = W * ((R + W * (hs * ws)) / np.maximum(hs + W * c + lw, EPS))
            W = W / np.sqrt((W ** 2).sum(0))
            ws = W.sum(0)
            Q = Vr / (W @ H)
            H = H * ((W.T @ Q) / np.maximum(ws[:, None] + lh, EPS))
            S = W @ H
            vls, ss, sh = allreduce(np.array([(Vr * np.log(S)).sum()]), np.array([S.sum()]), np.array([H.sum()]))
            cost.append(vlogv[0] - vls[0] - vsum[0] + ss[0] + lw * W.sum() + lh * sh[0])
    return W, H, np.array(cost), (lo, hi)


def _worker(rank, world, port, div, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    V = rng.random((48, 70)) + 1e-3
    W0 = rng.random((48, 6)) + 1e-3
    H0 = rng.random((6, 70)) + 1e-3
    W, H, cost, (lo, hi) = _sharded_nmf(rank, world, V, W0, H0, div, 12, 0.05, 0.1)
    np.savez(os.path.join(tmp, f"r{rank}.npz"), W=W, H=H, cost=cost, lo=lo, hi=hi)
    dist.destroy_process_group()


@pytest.mark.parametrize("div", ["euclidean", "kl"])
def test_two_rank_sharded_matches_oracle(div, tmp_path):
    from oracle import nmf_oracle as O

    world = 2
    port = 29500 + (os.getpid() % 500) + (0 if div == "euclidean" else 1)
    mp.spawn(_worker, args=(world, port, div, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(11)
    V = rng.random((48, 70)) + 1e-3
    W0 = rng.random((48, 6)) + 1e-3
    H0 = rng.random((6, 70)) + 1e-3
    Wo, Ho, co = O.nmf(V, 6, dict(divergence=div, W_init=W0, H_init=H0, maxiter=12, tolerance=1e-300,
                                  W_sparsity=0.05, H_sparsity=0.1))
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for p in parts:
        np.testing.assert_allclose(p["cost"], co, rtol=1e-9)   # global cost trace on every rank
        np.testing.assert_allclose(p["W"], Wo, rtol=1e-8)       # W replicated
        np.testing.assert_allclose(p["H"], Ho[:, int(p["lo"]):int(p["hi"])], rtol=1e-8)  # H sharded
    assert parts[0]["hi"] == parts[1]["lo"] and parts[1]["hi"] == 70
