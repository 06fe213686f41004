"""CPU suite: the drop-in boundary.  The shared library loads, exports every symbol the header
declares, refuses to compute without a GPU (no CPU fallback), and the product package never imports
the oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "eutropia_b200.h")).read()
    return sorted(set(re.findall(r"EUT_API[^;(]*?\b(eut_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from eutropia_b200 import _lib
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/eutropia_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms
    assert L.eut_abi_version() == 1000
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(syms) <= exported
    # nothing but the C ABI leaks out of the library
    assert all(e.startswith("eut_") for e in exported), sorted(e for e in exported if not e.startswith("eut_"))[:5]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import eutropia_b200 as eb
    m = eb.mesh.make_mesh("Rectangle_12_12", 1.0, 2)
    with pytest.raises(eb.EutropiaError) as ei:
        eb.diffChangePar(m.mat_in, m.mat_out, np.ones((m.nfields, 1)), [1], [1])
    assert ei.value.code == eb._lib.ERR_NO_DEVICE and "no CPU path" in str(ei.value)
    with pytest.raises(eb.EutropiaError):
        eb.Plan(m.mat_in, m.mat_out)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "eutropia_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{os.path.join(dirpath, f)} mentions the oracle"


def test_host_logic_niter_and_mesh_presets():
    import eutropia_b200 as eb
    # R/growthEnvironment.R:380-382 with the defaults D = 75, deltaTime = 1/6 h, fieldSize = 1
    assert eb.diffusion_niter([75.0], 1 / 6, 1.0)[0] == 60
    assert eb.diffusion_niter([10.0, 200.0, 0.0], 1 / 6, 1.0).tolist() == [22.0, 98.0, 0.0]
    p = eb.mesh.universe_polygon_preset("Petri_80")
    assert p.shape == (100, 2) and np.allclose(np.hypot(p[:, 0], p[:, 1]), 80)
    r = eb.mesh.universe_polygon_preset("Rectangle_20_10")
    assert r.tolist() == [[10, 5], [10, -5], [-10, -5], [-10, 5]]
    with pytest.raises(ValueError):
        eb.mesh.universe_polygon_preset("Petri_2")
    h = eb.mesh.harbor_polygon(100.0)
    # non-convex: some cross products change sign
    e = np.diff(np.vstack([h, h[:1]]), axis=0)
    cr = e[:-1, 0] * e[1:, 1] - e[:-1, 1] * e[1:, 0]
    assert (cr > 0).any() and (cr < 0).any()
