"""Step 3 (see export_inputs_for_matlab.py): compare the committed oracle fixtures with the outputs of the
real toolbox (tests/golden/matlab_io/<case>_out.mat written by regen_in_matlab.m).

Both are float64 runs of the same arithmetic, so they must agree to BLAS summation order:
cost curve 1e-9 relative, reconstruction 1e-8 relative.  --rewrite replaces the fixtures' cost / W / H by the
toolbox's, which turns "parity unpinned"  into a pin against the reference itself.
"""
import os
import sys

import numpy as np
from scipy.io import loadmat

HERE = os.path.dirname(os.path.abspath(__file__))
IO = os.path.join(HERE, "matlab_io")

if __name__ == "__main__":
    rewrite = "--rewrite" in sys.argv
    bad = 0
    for fn in sorted(os.listdir(IO)):
        if not fn.endswith("_out.mat"):
            continue
        name = fn[:-8]
        gpath = os.path.join(HERE, name + ".npz")
        if not os.path.exists(gpath):
            continue
        ref = loadmat(os.path.join(IO, fn))
        g = dict(np.load(gpath))
        cm, co = ref["cost"].ravel(), g["cost"]
        same_len = len(cm) == len(co)
        k = min(len(cm), len(co))
        rel = float(np.max(np.abs(cm[:k] - co[:k]) / np.maximum(np.abs(cm[:k]), 1e-300)))
        ok = same_len and rel < 1e-9
        if "W" in g:  # small fixtures keep the whole factors
            Wm, Hm = ref["W"], ref["H"]
            ok = ok and np.allclose(Wm, g["W"], rtol=1e-5, atol=1e-7) and np.allclose(Hm, g["H"], rtol=1e-5, atol=1e-7)
        print(f"{name}: cost entries {len(cm)} vs {len(co)}, max rel cost diff {rel:.2e} -> {'OK' if ok else 'MISMATCH'}")
        bad += 0 if ok else 1
        if rewrite:
            g["cost"] = cm
            if "W" in g:
                g["W"], g["H"] = ref["W"].astype(np.float32), ref["H"].astype(np.float32)
            np.savez_compressed(gpath, **g)
    sys.exit(1 if bad else 0)
