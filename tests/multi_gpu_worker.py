"""Worker for tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU):
column-sharded nmf through the engine's NCCL all-reduce; rank 0 gathers the H shards and
writes W, H, cost to an .npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def variant_config(variant, K):
    """Extra config fields of a test variant (shared with tests/test_gpu_multi.py)."""
    if variant == "w_fixed":  # nmf.m:146
        return dict(W_fixed=True)
    if variant == "per_source":  # first third of the bases fixed in W, different sparsity levels (nmf.m:51-60)
        k1 = K // 3
        return dict(W_sparsity=None, H_sparsity=None,
                    W_fixed_k=[1] * k1 + [0] * (K - k1), W_sparsity_k=[0.0] * k1 + [0.1] * (K - k1),
                    H_sparsity_k=[0.2] * k1 + [0.0] * (K - k1))
    return {}


def main():
    import torch
    import torch.distributed as dist
    from nmf_toolbox_b200 import api
    from nmf_toolbox_b200.distributed import nmf_sharded, shard_bounds

    out, div, m, n, K, iters = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    variant = sys.argv[7] if len(sys.argv) > 7 else "plain"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(21)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    W0 = rng.random((m, K)) + 1e-3
    H0 = rng.random((K, n)) + 1e-3
    lo, hi = shard_bounds(n, world, rank)
    h = api.Handle(local)
    cfg = dict(divergence=div, W_init=W0, H_init=H0[:, lo:hi], maxiter=iters, tolerance=1e-300,
               W_sparsity=0.05, H_sparsity=0.1)
    cfg.update(variant_config(variant, K))
    if variant.startswith("cnmf"):  # convolutive: column shards with (T-1)-column halos
        from nmf_toolbox_b200.distributed import cnmf_sharded
        T = int(variant[4:])
        W0 = rng.random((m, K, T)) + 1e-3
        cfg = dict(divergence=div, W_init=W0, H_init=H0[:, lo:hi], maxiter=iters, tolerance=1e-300,
                   W_sparsity=0.05, H_sparsity=0.1)
        W, H, cost = cnmf_sharded(h, dist, V[:, lo:hi], K, T, cfg, rank, world)
    else:
        W, H, cost = nmf_sharded(h, dist, V[:, lo:hi], K, cfg, rank, world)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, H))
    if rank == 0:
        Hfull = np.zeros((K, n), dtype=np.float32)
        for a, b, Hp in parts:
            Hfull[:, a:b] = Hp
        np.savez(out, W=W, H=Hfull, cost=cost)
    Wt = torch.from_numpy(np.ascontiguousarray(W)).cuda()
    Wref = Wt.clone()
    dist.broadcast(Wref, src=0)
    assert torch.equal(Wt, Wref), "W must be bit-identical on every rank"
    h.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
