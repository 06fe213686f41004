"""Slab-sharded diffusion throughput (BASELINE.json configs[4] shape): one field of
POINTS_PER_GPU x world points x COLS compounds, sharded by spatial slab, one NCCL halo exchange per
substep.  Launch: python -m torch.distributed.run --nproc-per-node N tools/slab_bench.py [ppg] [cols] [substeps]
Prints one JSON line on rank 0 (device time, max over ranks)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eutropia_b200 import synth  # noqa: E402
from eutropia_b200.slab import GpuSlabBackend, SlabDriver, SlabPartition  # noqa: E402

ppg = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 64
substeps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mesh = synth.config_mesh(f"square:{ppg * world}", on_device=True, device=local)   # neighbour maps on this rank's GPU
part = SlabPartition(mesh.mat_in, mesh.mat_out, mesh.nfields, world, rank)
rng = np.random.default_rng(5)
be = GpuSlabBackend(part, cols, local)
be.upload(np.asfortranarray(rng.uniform(0, 25, (part.n_local, cols))))
drv = SlabDriver(part, be, dist)
iv = np.arange(1, cols + 1, dtype=np.int32)
drv.diffuse(np.full(cols, 5), iv)
be.sync(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(be.stream):
    e0.record()
drv.diffuse(np.full(cols, substeps), iv)
with torch.cuda.stream(be.stream):
    e1.record()
be.sync(); torch.cuda.synchronize(); dist.barrier()
t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t[0])
    upd = mesh.nfields * cols * substeps
    print(json.dumps({"workload": "slab", "n_gpus": world, "points": mesh.nfields, "compounds": cols, "substeps": substeps,
                      "ms": ms, "updates_per_s": upd / ms * 1e3, "ghost_points_rank0": int(part.n_local - part.own_local.size),
                      "halo_bytes_per_substep_rank0": int(part.send_all.size * cols * 8),
                      "GB_s_per_gpu_algorithmic": 16 * upd / world / ms / 1e6}), flush=True)
be.close()
dist.destroy_process_group()
