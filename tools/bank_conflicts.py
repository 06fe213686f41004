"""Shared-memory wavefront model of the tiled kernel on a host-only plan: for every warp-level
LDS.64 (32 consecutive points of a tile, one neighbour slot) count 128-byte wavefronts = max over
the two half-warps of the number of distinct 8-byte words mapped to the same bank pair.
usage: python tools/bank_conflicts.py workload [threads] [tile_points]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eutropia_b200 as eb  # noqa: E402
from eutropia_b200 import _lib, synth  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "harbor:250000"
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 256
if len(sys.argv) > 3:
    os.environ["EUT_TILE_POINTS"] = sys.argv[3]
m = synth.config_mesh(workload)
plan = eb.Plan(m.mat_in, m.mat_out, order=_lib.ORDER_TILED, host_only=True)
inf = plan.info
meta, copies, loc = plan.tiles()
perm = plan.perm()
deg = np.bincount(m.mat_in[:, 1], minlength=m.nfields + 1)[1:]
degq = np.empty(m.nfields, dtype=np.int64); degq[perm] = deg
ideal = actual = 0
for t in range(meta.shape[0]):
    t0, cnt = meta[t, 0], meta[t, 1]
    for w0 in range(0, cnt, 32):
        q = np.arange(t0 + w0, t0 + min(cnt, w0 + 32))
        hi = degq[q].max() > 8
        for k in range(12 if hi else 8):
            offs = loc[k, q].astype(np.int64) // 8          # word index of each lane
            wf = 0
            for half in (offs[:16], offs[16:]):
                if half.size == 0:
                    continue
                words = np.unique(half)                       # same word -> broadcast
                banks = words % 16
                wf += np.bincount(banks, minlength=16).max()
            ideal += 2 if q.size > 16 else 1
            actual += wf
print(f"{workload}: tiles {inf.n_tiles} img {inf.img_points} halo {inf.halo_ratio:.3f}  "
      f"wavefronts ideal {ideal} actual {actual} excess {100 * (actual - ideal) / ideal:.1f}%")
