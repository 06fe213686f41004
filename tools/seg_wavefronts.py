"""Exact shared-memory wavefront count of the segment kernel's loads, from the plan tables (one column).
Mirrors stencil_seg.cu: per consumer warp (32 consecutive segment slots) and per load instruction,
wavefronts = sum over the two half warps of the max multiplicity of an 8-byte bank pair among distinct
addresses (same address = broadcast).  usage: python tools/seg_wavefronts.py [workload]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eutropia_b200 as eb  # noqa: E402
from eutropia_b200 import _lib, synth  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "harbor1m"
mesh = synth.config_mesh(workload)
plan = eb.Plan(mesh.mat_in, mesh.mat_out, order=_lib.ORDER_SEGMENTS, host_only=True)
inf = plan.info
meta, copies, segs, gens = plan.segs()
R = inf.seg_points
CW = 7 if inf.max_tile_segs <= 224 else 14


def wf_of(cells):
    """cells: [n_halfwarps, 16] int array of 8-byte cell indices -> total wavefronts"""
    tot = 0
    for row in cells:
        u = np.unique(row)
        tot += np.bincount(u % 16, minlength=16).max()
    return tot


tot = ideal = 0
by = {"inline": [0, 0], "ends": [0, 0], "sts": [0, 0]}
for t in range(meta.shape[0]):
    g0, ng = meta[t, 8], meta[t, 9]
    S = np.zeros((CW * 32, 24), dtype=np.int64)
    S[:, 0] = 8                                   # idle lane: O at cell 1
    S[:ng] = segs[g0:g0 + ng]
    C = S // 8
    C[:, 7] = S[:, 7]
    ln = S[:, 7] >> 12
    out = S[:, 7] & 0xfff
    for w in range(CW):
        L = C[32 * w:32 * w + 32]
        if w * 32 >= ng:
            continue
        has_lo = np.any((L[:, 1] | L[:, 2] | L[:, 10] | L[:, 11] | L[:, 12] | L[:, 13]) != 0)
        has_up = np.any((L[:, 5] | L[:, 6] | L[:, 14] | L[:, 15] | L[:, 16] | L[:, 17]) != 0)
        loads_inline = []
        loads_ends = [L[:, 8], L[:, 9]]
        loads_inline += [L[:, 0] + i for i in range(R)]
        for col, red, on in ((1, 10, has_lo), (2, 12, has_lo), (5, 14, has_up), (6, 16, has_up)):
            if on:
                loads_inline += [L[:, col] + i for i in range(1, R)]
                loads_ends += [L[:, red], L[:, red + 1]]
        for col in (3, 4):
            loads_inline += [L[:, col] + i for i in range(R)]
        for kind, lst in (("inline", loads_inline), ("ends", loads_ends)):
            for a in lst:
                w_ = wf_of(a.reshape(2, 16))
                by[kind][0] += w_; by[kind][1] += 2
        lnw, ow = ln[32 * w:32 * w + 32], out[32 * w:32 * w + 32]
        for i in range(R):
            act = lnw > i
            if act.any():
                a = np.where(act, ow + i, -1 - np.arange(32) * 16)      # inactive lanes: no access
                w_ = 0
                for h in range(2):
                    hh = a[16 * h:16 * h + 16]; hh = hh[hh >= 0]
                    if hh.size:
                        w_ += np.bincount(np.unique(hh) % 16, minlength=16).max()
                by["sts"][0] += w_; by["sts"][1] += 2
for k, (a, b) in by.items():
    print("%-7s wavefronts per column %9d  ideal %9d  ratio %.3f" % (k, a, b, a / max(b, 1)))
print("LDS total per column %d (x64 columns = %.2f M), ratio %.3f" % (by["inline"][0] + by["ends"][0], 64e-6 * (by["inline"][0] + by["ends"][0]),
      (by["inline"][0] + by["ends"][0]) / (by["inline"][1] + by["ends"][1])))
