"""Column-sharded nmf over several GPUs, one process per GPU.

V and H are partitioned by columns, W is replicated.  Rank r owns columns
``shard_bounds(n, world, r)``.  The H update is purely local; the W update needs
the sum over ranks of the m x K numerator partial V_r*H_r', the K x K Gram
matrix H_r*H_r' and a few scalar sums, which the engine all-reduces once per
iteration over NCCL (``nmfb_comm_init``).  ``torch.distributed`` (NCCL or gloo)
is only plumbing here: rendezvous, the broadcast of the NCCL unique id and the
gather of the H shards.
"""
from __future__ import annotations

import numpy as np

__all__ = ["shard_bounds", "shard_columns", "init_comm", "nmf_sharded", "cnmf_sharded", "pack_layout"]


def shard_bounds(n: int, world: int, rank: int):
    """Half-open column range [lo, hi) of `rank` when n columns are split over `world` ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_columns(X, world: int, rank: int):
    """Column shard of a matrix (V: m x n, H: K x n)."""
    lo, hi = shard_bounds(X.shape[1], world, rank)
    return np.asfortranarray(X[:, lo:hi])


def pack_layout(m: int, K: int):
    """Element offsets of the packed fp32 all-reduce buffer: [A (Kp x ldw) | G_H (Kp x Kp)].
    Mirrors NmfSession::packed in csrc/nmf_driver.cu."""
    Kp = (K + 31) // 32 * 32
    ldw = (m + 3) // 4 * 4
    return {"A": (0, Kp * ldw), "G_H": (Kp * ldw, Kp * ldw + Kp * Kp), "total": Kp * ldw + Kp * Kp}


def init_comm(handle, dist, rank: int, world: int):
    """Create the engine's NCCL communicator; the unique id travels through `dist`
    (an initialised torch.distributed module, any backend)."""
    if world == 1:
        return
    box = [handle.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    handle.comm_init(box[0], rank, world)


def nmf_sharded(handle, dist, V_shard, num_basis_elems, config, rank: int, world: int):
    """``nmf`` on this rank's column shard.  ``config['H_init']`` must be the rank's
    shard of H_init, ``config['W_init']`` the full W_init (identical on all ranks).
    Returns (W replicated, H shard, global cost trace)."""
    if config is None or config.get("W_init") is None or config.get("H_init") is None:
        raise ValueError("sharded runs need explicit W_init / H_init (every rank must start from the same W)")
    init_comm(handle, dist, rank, world)
    handle.set_V(V_shard)
    return handle.nmf(int(num_basis_elems), config)


def cnmf_sharded(handle, dist, V_shard, num_basis_elems, context_len, config, rank: int, world: int):
    """``cnmf`` ('euclidean' / 'frobenius') on this rank's block of CONSECUTIVE columns (time frames) of V.
    ``config['H_init']`` is the rank's shard of H_init, ``config['W_init']`` the full m x K x T tensor.  The
    engine exchanges the (context_len - 1)-column halos of H (every iteration) and of V (once) with the
    neighbouring ranks itself; every shard must hold at least context_len - 1 columns.
    Returns (W replicated, H shard, global cost trace)."""
    if config is None or config.get("W_init") is None or config.get("H_init") is None:
        raise ValueError("sharded runs need explicit W_init / H_init (every rank must start from the same W)")
    init_comm(handle, dist, rank, world)
    handle.set_V(V_shard)
    return handle.cnmf(int(num_basis_elems), int(context_len), config)
