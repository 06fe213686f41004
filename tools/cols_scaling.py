"""Substep time vs number of active columns (resident field): python tools/cols_scaling.py [workload]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eutropia_b200 as eb  # noqa: E402
from eutropia_b200 import synth  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "harbor1m"
mesh = synth.config_mesh(workload)
plan = eb.Plan(mesh.mat_in, mesh.mat_out)
rng = np.random.default_rng(0)
field = eb.Field(plan, 64)
field.upload(np.asfortranarray(rng.uniform(0, 25, (mesh.nfields, 64))))
S = 60
for n in (2, 4, 8, 16, 32, 64):
    iv = np.arange(1, n + 1, dtype=np.int32)
    it = np.full(n, S, np.int32)
    field.diffuse(it, iv); field.sync()
    t = time.perf_counter()
    for _ in range(3):
        field.diffuse(it, iv)
    field.sync()
    dt = (time.perf_counter() - t) / 3
    print("cols %2d: %.3f ms per %d substeps, %.4f ms per launch, %.1f G updates/s" % (n, dt * 1e3, S, dt * 1e3 / S, mesh.nfields * n * S / dt / 1e9))
