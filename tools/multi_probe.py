"""eut_diffchange_par_multi from one process on all visible GPUs, pinned buffers, with the per-device
timeline (EUT_TRACE).  python tools/multi_probe.py [cols_per_gpu]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eutropia_b200 as eb
from eutropia_b200 import synth, _lib

cpg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = _lib.lib()
nd = L.eut_device_count()
mesh = synth.config_mesh("harbor1m")
n = mesh.nfields
adj = np.require(mesh.mat_in, dtype=np.int32, requirements=["F", "A"])
nn = np.require(mesh.mat_out, dtype=np.int32, requirements=["F", "A"])
rng = np.random.default_rng(1)
for devs in ([0], list(range(nd))):
    ct = cpg * len(devs)
    a = torch.from_numpy(rng.uniform(0, 25, (ct, n))).pin_memory()
    b = torch.empty_like(a).pin_memory()
    niter, indvar = np.full(ct, 60, np.int32), np.arange(1, ct + 1, dtype=np.int32)
    dv = np.asarray(devs, np.int32)
    def call():
        _lib.check(L.eut_diffchange_par_multi(_lib.ptr(adj), adj.shape[0], _lib.ptr(nn), nn.shape[0], a.data_ptr(), n, ct,
                                              _lib.ptr(niter), _lib.ptr(indvar), ct, 1, b.data_ptr(), _lib.ptr(dv), len(devs)))
    call(); call()
    os.environ["EUT_TRACE"] = "1"
    t0 = time.perf_counter()
    for _ in range(3):
        call()
    dt = (time.perf_counter() - t0) / 3
    del os.environ["EUT_TRACE"]
    print(f"devices {devs}: {dt * 1e3:.2f} ms per call, {n * ct * 60 / dt / 1e9:.1f} G updates/s", flush=True)
    for var, val in (("EUT_NO_SPECULATION", "1"),):
        os.environ[var] = val
        call()
        t0 = time.perf_counter()
        for _ in range(3):
            call()
        dt = (time.perf_counter() - t0) / 3
        del os.environ[var]
        print(f"devices {devs}, {var}: {dt * 1e3:.2f} ms per call", flush=True)
    del a, b
