"""Sweep tile-planner / kernel options and time the diffusion substep with CUDA events.
usage: python tools/sweep.py workload cols "TILE_POINTS,TILE_COLS,nbuf,cpt[,variant]" ..."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eutropia_b200 as eb  # noqa: E402
from eutropia_b200 import synth  # noqa: E402

workload, cols = sys.argv[1], int(sys.argv[2])
mesh = synth.config_mesh(workload)
conc = np.asfortranarray(np.random.default_rng(0).uniform(0, 25, (mesh.nfields, cols)))
niter_w = np.full(cols, 6, np.int32)
niter_t = np.full(cols, 30, np.int32)
iv = np.arange(1, cols + 1, dtype=np.int32)
for spec in sys.argv[3:]:
    parts = [int(x) for x in spec.split(",")]
    tp, tc, nbuf, cpt = parts[:4]
    variant = parts[4] if len(parts) > 4 else 0
    dbg = parts[6] if len(parts) > 6 else 0
    shape = parts[7] if len(parts) > 7 else 0
    os.environ["EUT_TILE_POINTS"] = str(tp)
    os.environ["EUT_TILE_COLS"] = str(tc)
    plan = eb.Plan(mesh.mat_in, mesh.mat_out)
    f = eb.Field(plan, cols)
    f.set_option("nbuf", nbuf)
    f.set_option("cpt", cpt)
    f.set_option("variant", variant)
    f.set_option("debug", dbg)
    f.set_option("shape", shape)
    f.upload(conc)
    st = torch.cuda.ExternalStream(f.stream)
    f.diffuse(niter_w, iv)
    f.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
    f.diffuse(niter_t, iv)
    with torch.cuda.stream(st):
        e1.record()
    f.sync()
    ms = e0.elapsed_time(e1) / 30
    inf = plan.info
    gbs = 16 * mesh.nfields * cols / ms / 1e6
    print(f"{spec:>22s}  tiles {inf.n_tiles:5d} img {inf.img_points:5d} halo {inf.halo_ratio:.3f}  "
          f"{ms * 1e3:8.1f} us/substep  {gbs:7.1f} GB/s  frac {gbs / 6543.7:.3f}", flush=True)
    del f, plan
