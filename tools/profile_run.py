"""Small driver for ncu: builds the bench workload's plan + field and runs a few diffusion substeps
(and optionally one scatter / gather).  Usage: python tools/profile_run.py [workload] [substeps] [cols] [opt=val ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eutropia_b200 as eb  # noqa: E402
from eutropia_b200 import synth  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "harbor1m"
substeps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cols = int(sys.argv[3]) if len(sys.argv) > 3 else 64
opts = [a.split("=") for a in sys.argv[4:]]
order = 0
mesh = synth.config_mesh(workload)
rng = np.random.default_rng(0)
conc = np.asfortranarray(rng.uniform(0, 25, (mesh.nfields, cols)))
for k, v in opts:
    if k == "order":
        order = int(v)
plan = eb.Plan(mesh.mat_in, mesh.mat_out, order=order)
field = eb.Field(plan, cols)
for k, v in opts:
    if k != "order" and k != "cells":
        field.set_option(k, int(v))
field.upload(conc)
field.diffuse(np.full(cols, substeps, np.int32), np.arange(1, cols + 1, dtype=np.int32))
field.sync()
if any(k == "cells" for k, _ in opts):
    n_cells = int(dict(opts)["cells"])
    cx, cy, size, scv = synth.place_cells(mesh, n_cells, seed=1)
    cells = eb.Cells(plan)
    cells.build_lists(mesh.field_pts, cx, cy, size, scv)
    cells.claim_rescale()
    ex = synth.synthetic_exchanges(n_cells, cols, 10, seed=2)
    cells.scatter(field, mesh.field_vol, None, ex[0], ex[1], ex[2] * 1e-3)
    cells.gather(field, mesh.field_vol, want_host=False)
    field.sync()
print("done", mesh.nfields, cols, substeps, field.launch_count)
inf = plan.info
print("plan: order", inf.order, "tiled", inf.tiled, "segmented", inf.segmented, "segs", inf.n_segments, "generic", inf.n_generic, "tiles", inf.n_tiles, "img_points", inf.img_points,
      "max_tile", inf.max_tile_points, "max_bulk", inf.max_bulk, "max_single", inf.max_single, "halo_ratio %.3f" % inf.halo_ratio,
      "build_ms %.0f" % inf.build_ms, "graph MB %.1f" % (inf.device_bytes / 2**20))
