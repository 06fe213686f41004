"""Host-only statistics of the segment plan of a workload: python tools/seg_stats.py [workload]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eutropia_b200 as eb  # noqa: E402
from eutropia_b200 import _lib, synth  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "harbor1m"
t = time.time()
mesh = synth.config_mesh(workload)
print("mesh %.1fs N=%d" % (time.time() - t, mesh.nfields))
t = time.time()
plan = eb.Plan(mesh.mat_in, mesh.mat_out, order=_lib.ORDER_SEGMENTS, host_only=True)
inf = plan.info
print("plan %.1fs segmented=%d order=%d tiles=%d img=%d max_pts=%d max_segs=%d max_gen=%d bulk=%d single=%d halo=%.3f"
      % (time.time() - t, inf.segmented, inf.order, inf.n_tiles, inf.img_points, inf.max_tile_points, inf.max_tile_segs,
         inf.max_tile_generic, inf.max_bulk, inf.max_single, inf.halo_ratio))
print("n_padded / n_points = %.4f" % (inf.n_padded / inf.n_points))
print("segments %d generic %d fill %.3f" % (inf.n_segments, inf.n_generic, inf.n_seg_covered / max(1, inf.n_segments * inf.seg_points)))
if inf.segmented:
    meta, copies, segs, gens = plan.segs()
    R = inf.seg_points
    print("tile points median %d, segs median %d, img median %d, bulk median %d, single median %d"
          % (np.median(meta[:, 1]), np.median(meta[:, 9]), np.median(meta[:, 7]), np.median(meta[:, 3]), np.median(meta[:, 5])))
    # bank model: wavefronts of one LDS.64 of stream s over a half warp = max multiplicity of (cell mod 16)
    tot = np.zeros(7); cnt = 0
    cls = np.zeros(4, dtype=np.int64)
    for t in range(meta.shape[0]):
        g0, ng = meta[t, 8], meta[t, 9]
        S = segs[g0:g0 + ng].astype(np.int64)
        pad = (-ng) % 16
        if pad:
            S = np.concatenate([S, np.tile(np.array([[8] + [0] * 23]), (pad, 1))])
        H = S.reshape(-1, 16, 24) // 8
        for s in range(7):
            c = H[:, :, s]
            wf = np.zeros(H.shape[0])
            for h in range(H.shape[0]):
                u = np.unique(c[h][c[h] > 0])                         # same address / zero header = broadcast
                wf[h] = max(1, np.bincount(u % 16, minlength=16).max())
            tot[s] += wf.sum()
        cnt += H.shape[0]
    print("modelled wavefronts per half-warp load, streams O A B C D E F:", np.round(tot / cnt, 3))
