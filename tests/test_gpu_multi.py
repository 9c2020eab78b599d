"""Column-sharded nmf on 2 GPUs (one process per GPU, NCCL): the cost curve must match the
single-GPU run to summation order (1e-6) and the oracle to the parity tolerance."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("div", ["euclidean", "kl", "is"])
def test_two_gpu_matches_single_and_oracle(div, tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from nmf_toolbox_b200 import api
    from oracle import nmf_oracle as O

    m, n, K, iters = 384, 1000, 24, 30
    out = str(tmp_path / "multi.npz")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), out, div, str(m), str(n), str(K), str(iters)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = np.load(out)
    rng = np.random.default_rng(21)
    V = np.maximum(rng.random((m, n)), 2.0 ** -24)
    W0 = rng.random((m, K)) + 1e-3
    H0 = rng.random((K, n)) + 1e-3
    cfg = dict(divergence=div, W_init=W0, H_init=H0, maxiter=iters, tolerance=1e-300, W_sparsity=0.05, H_sparsity=0.1)
    h = api.Handle(0)
    W1, H1, c1 = api.nmf(V, K, cfg, handle=h)
    h.close()
    Wo, Ho, co = O.nmf(V, K, cfg)
    assert len(got["cost"]) == iters
    np.testing.assert_allclose(got["cost"], c1, rtol=1e-6)   # only the summation order differs
    np.testing.assert_allclose(got["cost"], co, rtol=1e-4)
    R, Ro = got["W"].astype(np.float64) @ got["H"].astype(np.float64), Wo @ Ho
    assert np.linalg.norm(R - Ro) / np.linalg.norm(Ro) < 1e-3
