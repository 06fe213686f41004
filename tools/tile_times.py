"""Per-tile CTA lifetimes of the segment kernel (debug bit 2048) against the tile's properties.
usage: python tools/tile_times.py [workload] [cols]   (GPU)"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eutropia_b200 as eb  # noqa: E402
from eutropia_b200 import _lib, synth  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "harbor1m"
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mesh = synth.config_mesh(workload)
hplan = eb.Plan(mesh.mat_in, mesh.mat_out, order=_lib.ORDER_SEGMENTS, host_only=True)   # the same tables, host side
meta, copies, segs, gens = hplan.segs()
plan = eb.Plan(mesh.mat_in, mesh.mat_out)
assert plan.info.n_tiles == meta.shape[0]
field = eb.Field(plan, cols)
field.upload(np.asfortranarray(np.random.default_rng(0).uniform(0, 25, (mesh.nfields, cols))))
field.set_option("debug", 2048)
os.environ["EUT_NO_GRAPHS"] = "1"
niter, iv = np.full(cols, 9, np.int32), np.arange(1, cols + 1, dtype=np.int32)
field.diffuse(niter, iv)
field.sync()
L = _lib.lib()
buf = (C.c_ulonglong * 8192)()
assert L.eut_debug_tile_ns(buf, 8192) == 0
ns = np.frombuffer(buf, dtype=np.uint64).reshape(2, 4096)[:, :meta.shape[0]].astype(np.float64)
t = ns[0] / 1e3                                        # us per CTA (slice 0)
# conflict factor per tile (as tools/seg_stats.py)
fac = np.zeros(meta.shape[0])
for k in range(meta.shape[0]):
    g0, ng = meta[k, 8], meta[k, 9]
    S = segs[g0:g0 + ng].astype(np.int64)
    pad = (-ng) % 16
    if pad:
        S = np.concatenate([S, np.tile(np.array([[8] + [0] * 23]), (pad, 1))])
    H = S.reshape(-1, 16, 24) // 8
    tot = cnt = 0
    for s in range(7):
        c = H[:, :, s]
        for h in range(H.shape[0]):
            u = np.unique(c[h][c[h] > 0])
            tot += max(1, np.bincount(u % 16, minlength=16).max()); cnt += 1
    fac[k] = tot / max(cnt, 1)
X = np.stack([np.ones(len(t)), meta[:, 1], meta[:, 9], meta[:, 11], meta[:, 3], meta[:, 5], meta[:, 7], (fac - 1) * meta[:, 9]], axis=1).astype(np.float64)
names = ["const", "points", "segments", "generic", "bulk", "single", "image cells", "(conflict-1)*segments"]
coef, *_ = np.linalg.lstsq(X, t, rcond=None)
print("CTA lifetime (us) for", cols // 2 if cols >= 32 else cols, "columns per CTA: min %.1f median %.1f max %.1f" % (t.min(), np.median(t), t.max()))
for n_, c_, x_ in zip(names, coef, X.mean(0)):
    print("  %-24s coef %10.5f   mean contribution %7.2f us" % (n_, c_, c_ * x_))
print("  residual rms %.2f us" % np.sqrt(np.mean((X @ coef - t) ** 2)))
order = np.argsort(t)
for k in list(order[:4]) + list(order[-8:]):
    print("  tile %3d: %6.1f us  points %4d segs %3d gen %3d bulk %2d single %3d img %4d conflict %.2f" % (k, t[k], meta[k, 1], meta[k, 9], meta[k, 11], meta[k, 3], meta[k, 5], meta[k, 7], fac[k]))
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "tile_times.npy"), np.concatenate([X, t[:, None], fac[:, None]], axis=1))
