"""Slab-sharded field over all GPUs of the box: ghost ring depth = substeps between two halo exchanges
(EUT_SLAB_DEPTH).  python tools/slab_depth_probe.py [depths ...]   (GPUs)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eutropia_b200 as eb
import bench

nd = torch.cuda.device_count()
peak, _ = bench.peaks()
for depth in (sys.argv[1:] or ["1", "2", "4", "8"]):
    os.environ["EUT_SLAB_DEPTH"] = depth
    r = bench.slab_leg(eb, torch, nd, peak)
    print(json.dumps({"gpus": nd, "depth": int(depth), **{k: r[k] for k in ("value", "ms_per_substep", "frac_of_hbm_peak_per_gpu", "halo_points_per_exchange", "gpu_launches", "bit_identical_to_unsharded")}}), flush=True)
