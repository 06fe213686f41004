"""Pipeline depth of the segment kernel (image slots x result buffers) on a regular mesh, where every
shape fits in shared memory.  usage: python tools/shape_probe.py [mesh] [cols]   (GPU)"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import torch
    import eutropia_b200 as eb
    from eutropia_b200 import synth
    mesh = synth.config_mesh(sys.argv[2])
    cols, shape = int(sys.argv[3]), int(sys.argv[4])
    plan = eb.Plan(mesh.mat_in, mesh.mat_out)
    f = eb.Field(plan, cols)
    f.upload(np.asfortranarray(np.random.default_rng(0).uniform(0, 25, (mesh.nfields, cols))))
    f.set_option("shape", shape)
    niter, iv = np.full(cols, 60, np.int32), np.arange(1, cols + 1, dtype=np.int32)
    st = torch.cuda.ExternalStream(f.stream, device=0)
    try:
        for _ in range(3):
            f.diffuse(niter, iv)
        f.sync()
    except Exception as e:
        print(json.dumps({"shape": shape, "error": str(e)[:120]}))
        sys.exit(0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        a.record()
    for _ in range(5):
        f.diffuse(niter, iv)
    with torch.cuda.stream(st):
        b.record()
    f.sync()
    ms = a.elapsed_time(b) / 5 / 60
    print(json.dumps({"mesh": sys.argv[2], "shape": shape, "ms_per_substep": ms, "TBps": 16 * mesh.nfields * cols / ms / 1e9,
                      "img_cells": int(plan.info.img_points), "max_tile_points": int(plan.info.max_tile_points)}))
    sys.exit(0)
mesh = sys.argv[1] if len(sys.argv) > 1 else "square:600000"
cols = sys.argv[2] if len(sys.argv) > 2 else "64"
for shape in (0, 1, 2, 3, 5):      # image slots x result buffers: 3x3 (default), 3x2, 4x2, 4x3, 3x4
    r = subprocess.run([sys.executable, __file__, "child", mesh, cols, str(shape)], capture_output=True, text=True)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
