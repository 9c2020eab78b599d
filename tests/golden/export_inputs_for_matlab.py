"""Step 1 of pinning the oracle against the REAL toolbox (for anyone who has MATLAB or Octave; the build
image has neither): write the seeded inputs of every golden case as .mat files.

    python tests/golden/export_inputs_for_matlab.py [--large]      # -> tests/golden/matlab_io/<case>_in.mat
    (MATLAB / Octave)  cd tests/golden; regen_in_matlab('/path/to/nmf-toolbox')
                                                                    # -> tests/golden/matlab_io/<case>_out.mat
    python tests/golden/compare_with_matlab.py [--rewrite]          # oracle fixtures vs the toolbox's outputs

Inputs cannot be regenerated inside MATLAB (NumPy's PCG64 stream is not available there), hence the files.
"""
import os
import sys

import numpy as np
from scipy.io import savemat

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import CASES, inputs  # noqa: E402
from make_golden_large import LARGE, large_inputs  # noqa: E402

OUT = os.path.join(HERE, "matlab_io")


def export(name, alg, V, K, T, cfg):
    d = dict(alg=alg, V=np.asarray(V, np.float64), K=float(K), T=float(T), maxiter=float(cfg["maxiter"]),
             tolerance=float(cfg["tolerance"]), W_init=np.asarray(cfg["W_init"], np.float64),
             H_init=np.asarray(cfg["H_init"], np.float64))
    for key in ("divergence", "alpha", "beta", "W_sparsity", "H_sparsity"):
        if cfg.get(key) is not None:
            d[key] = cfg[key] if isinstance(cfg[key], str) else float(cfg[key])
    savemat(os.path.join(OUT, name + "_in.mat"), d, do_compression=True)
    print("wrote", name)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        alg, V, K, T, cfg = inputs(name)
        export(name, alg, V, K, T, cfg)
    if "--large" in sys.argv:  # several GiB of .mat files and hours of MATLAB time at 16384 x 16384
        for name in LARGE:
            alg, Vt, K, T, cfg = large_inputs(name)
            export(name, alg, Vt.T, K, T, cfg)
