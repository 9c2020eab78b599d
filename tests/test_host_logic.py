"""CPU tests of the host side: the C-ABI library loads and exports every symbol that
include/nmfb200.h declares, fails loudly without a GPU (no CPU fallback), and the
Python mirror marshals the reference's config conventions.  No compute calls here."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nmfb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nmfb_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from nmf_toolbox_b200 import _lib

    return _lib.load()


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 18, names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nmfb200.h but not exported"


def test_bind_signatures(lib):
    from nmf_toolbox_b200 import api

    api._bind(lib)
    assert lib.nmfb_version().decode().startswith("nmfb200")


def test_no_cpu_fallback(lib):
    """Without a CUDA device creating a handle fails with a message; nothing computes on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nmf_toolbox_b200 import api

    with pytest.raises(api.NmfbError) as e:
        api.Handle(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nmf_toolbox_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from nmf_toolbox_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.NmfbLibraryMissing):
        _lib.load()


def test_config_struct_layout():
    from nmf_toolbox_b200 import api

    # must match struct nmfb_config in include/nmfb200.h (x86-64 SysV layout)
    assert ctypes.sizeof(api._Config) == 128
    assert api._Config.W_init.offset == 24 and api._Config.maxiter.offset == 64
    assert api._Config.tolerance.offset == 72 and api._Config.cost_mode.offset == 88
    assert api._Config.W_sparsity_k.offset == 96 and api._Config.H_fixed_k.offset == 120


def test_multi_source_mapping():
    from nmf_toolbox_b200 import api

    W = [np.ones((4, 2)), 2 * np.ones((4, 3))]
    H = [np.ones((2, 5)), 3 * np.ones((3, 5))]
    cfg = api._multi_source(dict(W_init=W, H_init=H, W_sparsity=[0.1, 0.1], H_fixed=[False, False]), [2, 3])
    assert cfg["W_init"].shape == (4, 5) and cfg["H_init"].shape == (5, 5)
    assert cfg["W_sparsity"] == 0.1 and cfg["H_fixed"] is False
    with pytest.raises(api.NmfbError):  # nmf.m:317-318
        api._multi_source(dict(W_sparsity=[0.1, 0.2, 0.3]), [2, 3])
    with pytest.raises(api.NmfbError):  # nmf.m:301-302
        api._multi_source(dict(W_init=[np.ones((4, 2))]), [2, 3])
    with pytest.raises(api.NmfbError):  # per-source levels: only nmf maps them (cnmf does not)
        api._multi_source(dict(W_sparsity=[0.1, 0.2]), [2, 3])
    # nmf: settings that differ between sources become per-basis vectors (nmfb_config::*_k)
    cfg = api._multi_source(dict(W_sparsity=[0.1, 0.2], H_sparsity=[-1, 0.5], W_fixed=[True, False], H_fixed=[0, 0]),
                            [2, 3], per_basis=True)
    np.testing.assert_allclose(cfg["W_sparsity_k"], [0.1, 0.1, 0.2, 0.2, 0.2])
    np.testing.assert_allclose(cfg["H_sparsity_k"], [0, 0, 0.5, 0.5, 0.5])  # nmf.m:321-333 clamps negatives
    assert list(cfg["W_fixed_k"]) == [1, 1, 0, 0, 0] and cfg["W_fixed"] is None
    assert cfg["H_fixed"] == 0 and "H_fixed_k" not in cfg


def test_shard_bounds_cover_and_partition():
    from nmf_toolbox_b200.distributed import pack_layout, shard_bounds

    for n, P in [(16384, 8), (65536, 8), (1000, 3), (7, 8), (20000, 4)]:
        edges = [shard_bounds(n, P, r) for r in range(P)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(P - 1))
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1
    lay = pack_layout(8192, 128)  # config 3: 8192*128 + 128^2 floats ~ 4.26 MB
    assert lay["total"] == 8192 * 128 + 128 * 128 and lay["G_H"][0] == 8192 * 128


def test_header_is_valid_c99_and_matches_the_python_struct(tmp_path):
    """include/nmfb200.h is the drop-in boundary: it must compile as plain C (the MEX gateway and any C
    caller include it) and nmfb_config must have the layout api._Config marshals."""
    import shutil
    import subprocess

    from nmf_toolbox_b200 import api

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text(
        '#include <stddef.h>\n#include <stdio.h>\n#include "nmfb200.h"\n'
        "int main(void) { nmfb_config c = {0}; (void)c;\n"
        '  printf("%zu %zu %zu %zu %zu\\n", sizeof(nmfb_config), offsetof(nmfb_config, W_init), offsetof(nmfb_config, maxiter),\n'
        "         offsetof(nmfb_config, cost_mode), offsetof(nmfb_config, W_sparsity_k)); return 0; }\n")
    exe = tmp_path / "t"
    inc = os.path.join(ROOT, "include")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    C = api._Config
    assert [int(x) for x in out] == [ctypes.sizeof(C), C.W_init.offset, C.maxiter.offset, C.cost_mode.offset,
                                     C.W_sparsity_k.offset]


def test_kl_split_planner():
    """choose_kl_splits (gemm_host.cu): the column splits of a kl_fused / ab_fused launch are chosen by playing the
    launch through a list scheduler of the 74 resident CTA-pair slots of a B200.  Pure host logic, called through
    the test library's hook (no GPU needed)."""
    import ctypes

    path = os.path.join(ROOT, "tests", "libnmfb200_test.so")
    if not os.path.exists(path):
        pytest.skip("tests/libnmfb200_test.so not built")
    lib = ctypes.CDLL(path)
    per = ctypes.c_int(0)

    def plan(rows, cols, Kp, max_per=0, slots=74, chunk=4):
        pairs, tiles = -(-rows // 256), -(-cols // 64)
        s = lib.nmfb_debug_kl_splits(pairs, tiles, rows, Kp, slots, chunk, max_per, ctypes.byref(per))
        return pairs, tiles, s, per.value

    # BASELINE config 3, W half: 32 row blocks x (460 + 460 + 104 tiles): 64 long items on 64 slots, the 32 short
    # ones on the other ten - instead of 2 equal splits that leave ten slots idle for the whole launch
    assert plan(8192, 65536, 128)[2:] == (3, 460)
    # H half: 256 row blocks x 2 splits = 6.9 waves instead of 3.5 rounded up to 4
    assert plan(65536, 8192, 128)[2:] == (2, 64)
    for rows, cols, Kp, max_per in [(8192, 65536, 128, 0), (65536, 8192, 128, 0), (8192, 8192, 128, 64), (700, 900, 96, 0),
                                    (300, 64, 32, 0), (100000, 130, 128, 0), (2100, 4170, 128, 64), (256, 64 * 1000, 128, 64)]:
        pairs, tiles, s, p = plan(rows, cols, Kp, max_per)
        assert s >= 1 and p >= 4 and p % 4 == 0, (rows, cols, s, p)
        assert (s - 1) * p < tiles <= s * p, "splits must cover the column tiles exactly once"
        if max_per:
            assert p <= max_per
        assert s <= 256 and s * pairs <= 65535  # gridDim.y and the cluster count stay launchable


def test_tail_helper_balance():
    """balance_tail_helpers (gemm_host.cu): where the primaries of a large contraction hand over to the helper CTA
    pairs.  Pure host arithmetic behind plan_tail_helpers, called through the test library's hook."""
    import ctypes

    path = os.path.join(ROOT, "tests", "libnmfb200_test.so")
    if not os.path.exists(path):
        pytest.skip("tests/libnmfb200_test.so not built")
    lib = ctypes.CDLL(path)
    kp = ctypes.c_int(0)

    def plan(tiles, helpers, nkb0):
        h = lib.nmfb_debug_tail_balance(tiles, helpers, nkb0, ctypes.byref(kp))
        return h, kp.value

    # north star: 64 row tiles, 72 - 64 = 8 free pairs, 16384 / 32 = 512 k-blocks: each helper takes the tails of 8
    # tiles, 8 * (512 - kp + 8) = kp  ->  kp = 463 (measured optimum between 456 and 464)
    assert plan(64, 8, 512) == (8, 463)
    assert plan(64, 8, 256) == (8, 235)          # a 2-GPU shard's A contraction: still worth 8 %
    assert plan(64, 8, 128)[0] == 0              # shorter contractions: the per-tile cost of a helper eats the gain
    assert plan(64, 8, 64)[0] == 0
    assert plan(64, 0, 512)[0] == 0 and plan(64, -3, 512)[0] == 0
    h, k = plan(38, 34, 300)                     # more free pairs than half the tiles: two tiles per helper
    assert h == 34 and 0 < k < 300 and 2 * (300 - k + 8) <= k + 2
    h, k = plan(32, 40, 512)                     # never more helpers than tiles
    assert h == 32 and k == 260
    for tiles, helpers, nkb0 in [(64, 8, 512), (50, 22, 1000), (70, 2, 4096), (40, 32, 100)]:
        h, k = plan(tiles, helpers, nkb0)
        if h:
            per = -(-tiles // h)
            assert 1 <= k < nkb0 and abs(per * (nkb0 - k + 8) - k) <= per + 1  # both sides finish together
