"""All GPUs of the box from one process: (1) host <-> device copy ceiling (pinned buffers, both directions at
once, 1 .. N GPUs), (2) the slab leg of bench.py, (3) eut_diffchange_par_multi with the per-device timeline.
python tools/n8_probe.py [what]   what: pcie,slab,multi (default all)"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eutropia_b200 as eb
from eutropia_b200 import synth, _lib
import bench

what = (sys.argv[1] if len(sys.argv) > 1 else "pcie,slab,multi").split(",")
nd = torch.cuda.device_count()
print("devices", nd, "host threads", bench.host_threads(), flush=True)
if "pcie" in what:
    mb = 512
    bufs = []
    for d in range(nd):
        torch.cuda.set_device(d)
        bufs.append((torch.empty(mb << 20, dtype=torch.uint8).pin_memory(), torch.empty(mb << 20, dtype=torch.uint8).pin_memory(),
                     torch.empty(mb << 20, dtype=torch.uint8, device=f"cuda:{d}"), torch.empty(mb << 20, dtype=torch.uint8, device=f"cuda:{d}"),
                     torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))
    for k in sorted({1, 2, 4, nd} & set(range(1, nd + 1))):
        for mode in ("h2d", "d2h", "both"):
            for d in range(k):
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            reps = 4
            for _ in range(reps):
                for d in range(k):
                    hi, ho, di, do, s1, s2 = bufs[d]
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(s1):
                            di.copy_(hi, non_blocking=True)
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(s2):
                            ho.copy_(do, non_blocking=True)
            for d in range(k):
                torch.cuda.synchronize(d)
            dt = time.perf_counter() - t0
            per_dir = reps * k * (mb << 20) / dt / 1e9
            print(json.dumps({"probe": "pcie", "gpus": k, "mode": mode, "GBps_per_direction_aggregate": per_dir, "GBps_per_direction_per_gpu": per_dir / k}), flush=True)
    del bufs
if "slab" in what:
    peak, _ = bench.peaks()
    for groups in ("2", "1"):
        os.environ["EUT_SLAB_GROUPS"] = groups
        r = bench.slab_leg(eb, torch, nd, peak)
        print(json.dumps({"probe": "slab", "column_groups": groups, **{k: r[k] for k in ("value", "ms_per_substep", "frac_of_hbm_peak_per_gpu", "gpu_launches", "bit_identical_to_unsharded", "nvlink_bytes_per_substep", "setup_s")}}), flush=True)
    os.environ["EUT_NO_GRAPHS"] = "1"
if "multi" in what:
    L = _lib.lib()
    mesh = synth.config_mesh("harbor1m")
    n = mesh.nfields
    adj = np.require(mesh.mat_in, dtype=np.int32, requirements=["F", "A"])
    nn = np.require(mesh.mat_out, dtype=np.int32, requirements=["F", "A"])
    rng = np.random.default_rng(1)
    for k in sorted({1, 2, 4, nd} & set(range(1, nd + 1))):
        ct = 64 * k
        a = torch.from_numpy(rng.uniform(0, 25, (ct, n))).pin_memory()
        b = torch.empty_like(a).pin_memory()
        niter, indvar = np.full(ct, 60, np.int32), np.arange(1, ct + 1, dtype=np.int32)
        dv = np.arange(k, dtype=np.int32)
        def call():
            _lib.check(L.eut_diffchange_par_multi(_lib.ptr(adj), adj.shape[0], _lib.ptr(nn), nn.shape[0], a.data_ptr(), n, ct,
                                                  _lib.ptr(niter), _lib.ptr(indvar), ct, 1, b.data_ptr(), _lib.ptr(dv), k))
        call(); call()
        if k == nd:
            os.environ["EUT_TRACE"] = "1"
        t0 = time.perf_counter()
        for _ in range(3):
            call()
        dt = (time.perf_counter() - t0) / 3
        os.environ.pop("EUT_TRACE", None)
        print(json.dumps({"probe": "multi", "gpus": k, "ms_per_call": dt * 1e3, "G_updates_per_s": n * ct * 60 / dt / 1e9,
                          "host_GBps_each_way": ct * n * 8 / dt / 1e9}), flush=True)
        del a, b
