"""Writes the field points of `Kiel_<size>` (3 layers) into the git-ignored mesh_cache/ so that a gpurun box,
which has no /root/reference, can run `bench.py --workload kiel1m`.  Build container only.
usage: python tools/cache_kiel_points.py [size]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eutropia_b200 import mesh  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1712
poly = mesh._close(mesh.universe_polygon_preset(f"Kiel_{size}"))
pts = mesh.make_field_points(poly, 1.0, 3)
os.makedirs(os.path.join(ROOT, "mesh_cache"), exist_ok=True)
out = os.path.join(ROOT, "mesh_cache", f"kiel_{size}_pts.npz")
np.savez_compressed(out, pts=pts, polygon=poly)
print(out, pts.shape[0], "points", os.path.getsize(out) >> 20, "MiB")
