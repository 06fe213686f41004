"""Where the end-to-end time of the one-shot call goes: raw PCIe rates on this box (pinned H2D, D2H,
both at once), then eut_diffchange_par on the bench workload with EUT_TRACE=1 under the settings
given on the command line (KEY=VALUE pairs exported before the library loads).
    python tools/oneshot_probe.py [--workload harbor1m] [--cols 64] [ENV=VALUE ...]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="harbor1m")
    ap.add_argument("--cols", type=int, default=64)
    ap.add_argument("--calls", type=int, default=5)
    ap.add_argument("env", nargs="*")
    args = ap.parse_args()
    for kv in args.env:
        k, v = kv.split("=", 1)
        os.environ[k] = v
    os.environ.setdefault("EUT_TRACE", "1")
    import torch
    import bench
    import eutropia_b200 as eb

    dev = 0
    torch.cuda.set_device(dev)
    mesh_name, cpg, substeps, n_cells, per_cell = bench.WORKLOADS[args.workload]
    cpg = args.cols
    mesh, conc, _, _, _ = bench.build_inputs(args.workload, cpg, seed=1)
    n = mesh.nfields
    pin_in = torch.from_numpy(np.ascontiguousarray(conc.T)).contiguous().pin_memory()
    pin_out = torch.empty_like(pin_in).pin_memory()
    d_a = torch.empty_like(pin_in, device="cuda")
    d_b = torch.empty_like(pin_in, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    nbytes = pin_in.numel() * 8

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(pin_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            pin_out.copy_(d_b, non_blocking=True)

    def both():
        h2d(); d2h()

    for name, fn in (("H2D", h2d), ("D2H", d2h), ("H2D+D2H at once", both)):
        t = timed(fn)
        print(f"{name:>16}: {nbytes / 2**20:.0f} MiB in {t * 1e3:.2f} ms = {nbytes / t / 1e9:.1f} GB/s per direction", flush=True)

    adj = np.require(mesh.mat_in, dtype=np.int32, requirements=["F", "A"])
    nn = np.require(mesh.mat_out, dtype=np.int32, requirements=["F", "A"])
    niter = np.full(cpg, substeps, dtype=np.int32)
    indvar = np.arange(1, cpg + 1, dtype=np.int32)
    L = eb._lib.lib()

    def call():
        eb._lib.check(L.eut_diffchange_par(eb._lib.ptr(adj), adj.shape[0], eb._lib.ptr(nn), nn.shape[0],
                                           pin_in.data_ptr(), n, cpg, eb._lib.ptr(niter), eb._lib.ptr(indvar),
                                           cpg, 1, pin_out.data_ptr(), dev))
    call(); call()
    ts = []
    for _ in range(args.calls):
        t0 = time.perf_counter()
        call()
        ts.append(time.perf_counter() - t0)
    upd = n * cpg * substeps
    print(f"one-shot {args.workload} x {cpg} cols x {substeps} substeps [{' '.join(args.env)}]: "
          f"min {min(ts) * 1e3:.2f} ms, median {np.median(ts) * 1e3:.2f} ms = {upd / np.median(ts) / 1e9:.1f} G updates/s", flush=True)


if __name__ == "__main__":
    main()
