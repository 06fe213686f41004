#!/usr/bin/env python
"""Benchmark of the hot path: field-point x compound diffusion updates/s on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one simulated round of the hot path on a resident field: scatter of the cells' exchange
fluxes (+clamp) -> S diffusion substeps on every compound -> gather of the accessible amounts
(R/run_simulation.R:309-319, :380-383, :243-258).  Workload at N=1: BASELINE.json configs[2]
(non-convex polygon, ~1M points x 64 compounds, 60 substeps, 50k cells); at N GPUs the same mesh with
64*N compounds sharded by compound (configs[3] at N=8), no data-path collective -> weak scaling.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "field-point*compound diffusion updates/s"
UNIT = "updates/s"
BYTES_PER_UPDATE = 16  # one fp64 read + one fp64 write per point, compound and substep

WORKLOADS = {
    # name: (mesh, compounds per GPU, substeps, cells, exchanges per cell)
    "harbor1m": ("harbor1m", 64, 60, 50_000, 10),
    "square100k": ("square100k", 16, 1000, 0, 0),
    "harbor250k": ("harbor:250000", 64, 60, 12_500, 10),
    # configs[2] on the reference's own polygon, Kiel_1712 (999 393 points); needs the reference's kiel.csv or the
    # point cache of tools/cache_kiel_points.py -- the default stays harbor1m, which needs neither
    "kiel1m": ("kiel1m", 64, 60, 50_000, 10),
    "tiny": ("harbor:20000", 8, 10, 500, 4),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons during the timed region: NVML polled from a thread every few
    milliseconds (starts at once, so even a region of a few steps is sampled); `nvidia-smi -lms` as
    the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.handle, self.thread, self.halt = None, None, None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            try:                                    # CUDA ordinal -> NVML device by PCI address
                import torch
                pr = torch.cuda.get_device_properties(index)
                handle = pynvml.nvmlDeviceGetHandleByPciBusId("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
            except Exception:
                handle = None
            if handle is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [v for v in vis.split(",") if v.strip().isdigit()]
                handle = pynvml.nvmlDeviceGetHandleByIndex(int(ids[index]) if index < len(ids) else index)
            self.nvml, self.handle = pynvml, handle
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.handle))
        self.rows.append((sm, self.max_mhz, {nm for nm, bit in self.BITS if mask & bit}))

    def _poll(self):
        while not self.halt.is_set():
            try:
                self._sample()
            except Exception:
                return
            self.halt.wait(0.005)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((float(r[0]), float(r[1]),
                                  {nm for nm, v in zip(names, r[3:7]) if v.lower().startswith("active")}))
            except Exception:
                continue

    def stop(self):
        """Call while the last timed kernels are still in flight (before the closing synchronize)."""
        source = "nvml"
        if self.thread is not None:
            self.halt.set()
            self.thread.join(timeout=1.0)
        elif self.proc is not None:
            source = "nvidia-smi"
            time.sleep(0.15)
            self.proc.terminate()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"], "samples": 0}
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        reasons = set().union(*[r[2] for r in self.rows]) if self.rows else set()
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": source}


def build_inputs(workload: str, n_cols: int, seed: int, want_conc=True):
    from eutropia_b200 import synth
    mesh_name, _, substeps, n_cells, per_cell = WORKLOADS[workload]
    mesh = synth.config_mesh(mesh_name)
    conc = synth.synthetic_concentrations(mesh.field_pts, n_cols, seed=seed) if want_conc else None
    cells = synth.place_cells(mesh, n_cells, seed=seed + 1) if n_cells else None
    ex = synth.synthetic_exchanges(n_cells, n_cols, per_cell, seed=seed + 2) if n_cells else None
    if ex is not None:
        ex = (ex[0], ex[1], ex[2] * 1e-3)   # fmol per round, well below the accessible amounts
    return mesh, conc, cells, ex, substeps


def numa_nodes() -> int:
    try:
        return max(1, sum(1 for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()))
    except OSError:
        return 1


def interleave_host_memory(on: bool) -> bool:
    """MPOL_INTERLEAVE over all NUMA nodes for the pages this thread touches from now on (what `numactl
    --interleave=all R` does for an R session): one process feeding N GPUs from a matrix that lives on ONE node is
    bound by that node's memory channels and the inter-socket link.  Returns whether the policy was set."""
    try:
        import ctypes
        n = numa_nodes()
        if n < 2 and on:
            return False
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong((1 << n) - 1)
        mode = 3 if on else 0                                   # MPOL_INTERLEAVE / MPOL_DEFAULT
        rc = libc.syscall(238, mode, ctypes.byref(mask) if on else None, ctypes.c_ulong(n + 1 if on else 0))   # SYS_set_mempolicy (x86-64)
        return rc == 0
    except Exception:
        return False


def host_threads() -> int:
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which is about torch's
    intra-op threads and not a limit on the box: the reference's policy is `detectCores() - 1`,
    R/run_simulation.R:137-138; the CPU arm gets all cores)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(mesh, substeps_full, n_cols_full, threads, target_s=12.0):
    """Time the reference's CPU implementation of diffChangePar -- its own src/diffChange.cpp from
    oracle/_ref (kind "reference") when that library was built, else the oracle's restatement (kind
    "port"); both run OpenMP over compounds -- on a bounded sample of the workload: same mesh,
    `threads` columns (one per thread, at most the workload's columns), as many substeps as fit in
    about target_s."""
    import oracle
    if oracle.ref_available():
        run, kind, what = oracle.ref_diffChangePar, "reference", "src/diffChange.cpp via oracle/_ref"
    else:
        run, kind, what = oracle.diffChangePar, "port", "oracle/diffchange_oracle.c"
    n = mesh.nfields
    rng = np.random.default_rng(99)
    # calibration: the marginal cost of a substep on one column per thread (two call lengths, so
    # that the per-call copies of the matrix do not count as substep time)
    c1 = int(max(1, min(n_cols_full, threads)))
    conc = np.asfortranarray(rng.uniform(0.0, 25.0, (n, c1)))

    def once(k):
        t0 = time.perf_counter()
        run(mesh.mat_in, mesh.mat_out, conc, np.full(c1, k, np.int32), np.arange(1, c1 + 1), threads)
        return time.perf_counter() - t0
    once(1)                                                # cold pages, thread start-up
    ta, tb = once(1), once(5)
    t1 = max((tb - ta) / 4, 1e-6)
    # sample: full substep count if it fits, on as many column waves (threads columns each) as fit
    s = int(max(1, min(substeps_full, target_s / t1)))
    waves = int(max(1, min(round(target_s / (t1 * s)), n_cols_full // c1)))
    cols = int(min(n_cols_full, c1 * waves))
    conc = np.asfortranarray(rng.uniform(0.0, 25.0, (n, cols)))
    iv = np.arange(1, cols + 1, dtype=np.int32)
    t0 = time.perf_counter()
    run(mesh.mat_in, mesh.mat_out, conc, np.full(cols, s, np.int32), iv, threads)
    dt = time.perf_counter() - t0
    return {"value": n * cols * s / dt, "unit": UNIT, "cores": int(threads), "kind": kind,
            "sample": f"{n} points x {cols} compounds x {s} substeps in {dt:.2f} s "
                      f"({what}, OpenMP over compounds, {threads} threads)"}



def reference_columns(mesh, conc, substeps, threads):
    """`conc` (N x k) diffused `substeps` times by the reference's own diffChangePar (oracle/_ref) or, if
    that library is absent, by the oracle port; the checker of the parity gates."""
    import oracle
    oracle.build()
    run = oracle.ref_diffChangePar if oracle.ref_available() else oracle.diffChangePar
    k = conc.shape[1]
    return run(mesh.mat_in, mesh.mat_out, np.asfortranarray(conc), np.full(k, substeps, np.int32),
               np.arange(1, k + 1, dtype=np.int32), max(1, min(threads, k))), ("reference" if oracle.ref_available() else "port")


def compare(got, ref):
    denom = np.maximum(np.abs(ref), 1e-300)
    return {"bit_equal": bool(np.array_equal(got, ref)), "max_rel": float(np.max(np.abs(got - ref) / denom))}


def parity_gate(eb, mesh, plan, conc, substeps, dev, k=4):
    """BASELINE.md 3 "parity gate on every timed config": k columns of the TIMED mesh x the timed substep
    count through the session tier (the kernels `value` times) and through the one-shot C ABI (what `e2e`
    times), against the reference's diffChangePar on the host.  Outside every timed region."""
    sub = np.asfortranarray(conc[:, :k])
    ref, kind = reference_columns(mesh, sub, substeps, host_threads())
    niter, indvar = np.full(k, substeps, np.int32), np.arange(1, k + 1, dtype=np.int32)
    f = eb.Field(plan, k)
    f.upload(sub)
    f.diffuse(niter, indvar)
    a = compare(f.download(), ref)
    f.close()
    b = compare(eb.diffChangePar(mesh.mat_in, mesh.mat_out, sub, niter, indvar, device=dev), ref)
    out = {"columns": k, "substeps": int(substeps), "points": int(mesh.nfields), "against": kind,
           "session": a, "one_shot": b, "bit_equal": a["bit_equal"] and b["bit_equal"],
           "max_rel": max(a["max_rel"], b["max_rel"]), "tolerance": 1e-12}
    if not (out["max_rel"] <= 1e-12):
        raise SystemExit(f"bench.py: parity gate failed on the timed configuration: {json.dumps(out)}")
    return out


def square100k_leg(eb, torch, dev, peak):
    """BASELINE.json configs[1]: square field, 100 467 points x 16 compounds x 1000 substeps on one GPU
    (L2-resident; bounded by launch latency, not HBM).  Device time by CUDA events; parity on 2 columns x
    1000 substeps against the reference."""
    from eutropia_b200 import synth
    mesh = synth.config_mesh("square100k")
    n, c, s = mesh.nfields, 16, 1000
    conc = synth.synthetic_concentrations(mesh.field_pts, c, seed=7)
    plan = eb.Plan(mesh.mat_in, mesh.mat_out, device=dev)
    field = eb.Field(plan, c)
    field.upload(conc)
    niter, indvar = np.full(c, s, np.int32), np.arange(1, c + 1, dtype=np.int32)
    stream = torch.cuda.ExternalStream(field.stream, device=dev)
    for _ in range(3):
        field.diffuse(niter, indvar)
    field.sync()
    steps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = field.launch_count
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(steps):
        field.diffuse(niter, indvar)
    with torch.cuda.stream(stream):
        e1.record()
    field.sync()
    ms = e0.elapsed_time(e1) / steps
    launches = (field.launch_count - l0) // steps
    field.upload(conc)
    field.diffuse(niter[:2], indvar[:2])
    ref, kind = reference_columns(mesh, conc[:, :2], s, 2)
    par = compare(field.download(cols=[1, 2]), ref)
    field.close(); plan.close()
    if not (par["max_rel"] <= 1e-12):
        raise SystemExit(f"bench.py: parity gate failed on square100k: {json.dumps(par)}")
    value = n * c * s / (ms * 1e-3)
    return {"workload": f"square100k: {n} points x {c} compounds x {s} substeps (configs[1]), field resident in L2",
            "value": value, "unit": UNIT, "ms_per_step": ms, "us_per_substep": ms * 1e3 / s, "launches_per_step": int(launches),
            "algorithmic_GBps": BYTES_PER_UPDATE * value / 1e9, "frac_of_hbm_peak": BYTES_PER_UPDATE * value / 1e9 / peak,
            "parity": dict(par, against=kind, columns=2, substeps=s)}


def slab_leg(eb, torch, world, peak, points_per_gpu=2_000_000, cols=64, substeps=60):
    """BASELINE.json configs[4] in shape: ONE field of 2 M points per GPU x 64 compounds sharded by spatial
    slab over all GPUs of the box, driven by this process through eut_slab_* (halo values stored straight
    into the neighbours' ghost cells over NVLink peer mappings).  Wall clock around the asynchronous
    diffuse call plus the closing synchronize; 2 columns are checked bit for bit against the unsharded
    field on one GPU."""
    from eutropia_b200 import synth
    t0 = time.perf_counter()
    mesh = synth.config_mesh(f"square:{points_per_gpu * world}", on_device=True, device=0)
    n = mesh.nfields
    t_mesh = time.perf_counter() - t0
    rng = np.random.default_rng(5)
    conc = np.empty((n, cols), order="F")
    for c in range(cols):
        conc[:, c] = rng.uniform(0, 25, n)
    sf = eb.SlabField(mesh.mat_in, mesh.mat_out, cols, list(range(world)))
    info = sf.info
    sf.upload(conc)
    iv = np.arange(1, cols + 1, dtype=np.int32)
    for _ in range(2):               # the loop shape is captured as one CUDA graph the second time it is seen
        sf.diffuse(np.full(cols, substeps, np.int32), iv)
        sf.sync()
    sf.upload(conc)
    l0 = sf.launch_count
    t0 = time.perf_counter()
    sf.diffuse(np.full(cols, substeps, np.int32), iv)
    sf.sync()
    dt = time.perf_counter() - t0
    launches = sf.launch_count - l0
    got = sf.download()[:, :2]
    sf.close()
    plan = eb.Plan(mesh.mat_in, mesh.mat_out, device=0)
    f = eb.Field(plan, 2)
    f.upload(conc[:, :2])
    f.diffuse(np.full(2, substeps, np.int32), iv[:2])
    same = bool(np.array_equal(f.download(), got))
    f.close(); plan.close()
    if not same and not os.environ.get("EUT_SKIP_CHECK"):
        raise SystemExit("bench.py: the slab-sharded field differs from the unsharded one")
    value = n * cols * substeps / dt
    gbps = BYTES_PER_UPDATE * value / world / 1e9
    return {"workload": f"slab: one field of {n} points x {cols} compounds x {substeps} substeps, sharded by spatial slab over "
                        f"{world} GPUs from one process (configs[4] shape: 2 M points per GPU)",
            "value": value, "unit": UNIT, "ms_per_substep": dt * 1e3 / substeps, "n_gpus": world,
            "algorithmic_GBps_per_gpu": gbps, "frac_of_hbm_peak_per_gpu": gbps / peak,
            "ghost_rings": int(info.depth), "substeps_between_exchanges": int(info.depth),
            "halo_points_per_exchange": int(info.halo_points),
            "nvlink_bytes_per_substep": int(info.halo_points) * cols * 8 // max(1, int(info.depth)),
            "links": int(info.n_links), "gpu_launches": int(launches), "bit_identical_to_unsharded": same,
            "segment_kernel_on_every_slab": bool(info.segmented), "setup_s": {"mesh": t_mesh, "partition_and_plans": info.build_ms / 1e3},
            "timer": "host wall clock around eut_slab_diffuse + eut_slab_sync"}


def run_reference(args):
    """--impl reference: the reference's own diffChangePar (src/diffChange.cpp compiled into
    oracle/_ref; the oracle port if that library is absent) on this box's host cores, same config
    and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = host_threads()
    mesh_name, cpg, substeps, n_cells, per_cell = WORKLOADS[args.workload]
    mesh, _, _, _, _ = build_inputs(args.workload, 1, seed=1, want_conc=False)
    per_step_target = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    sample = None
    for it in range(args.warmup + args.steps):
        r = cpu_sample(mesh, substeps, cpg * args.gpus, threads, target_s=per_step_target)
        if it >= args.warmup:
            vals.append(r["value"])
        sample = r
    v = float(np.mean(vals))
    n_total = mesh.nfields * cpg * args.gpus * substeps
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": n_total / v * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {mesh.nfields} points x {cpg * args.gpus} compounds x "
                                   f"{substeps} substeps (CPU arm times a bounded sample per step; ms_per_step is "
                                   f"the full step extrapolated from it)"},
            "cpu_baseline": dict(sample, value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import eutropia_b200 as eb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # NCCL prints its version banner on stdout at the VERSION level; stdout is the one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")       # host-side barriers (an NCCL barrier spins on the GPU)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: eutropia_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = local

    mesh_name, cpg, substeps, n_cells, per_cell = WORKLOADS[args.workload]
    # compound sharding: rank r owns columns [r*cpg, (r+1)*cpg) of the global N x (cpg*world) field
    mesh, conc, cellgeo, ex, substeps = build_inputs(args.workload, cpg, seed=1000 + rank)
    n = mesh.nfields
    plan = eb.Plan(mesh.mat_in, mesh.mat_out, device=dev)
    field = eb.Field(plan, cpg)
    for k, v in (args.opt or []):
        field.set_option(k, v)
    field.upload(conc)
    cells = None
    if n_cells:
        cells = eb.Cells(plan)
        cells.build_lists(mesh.field_pts, *cellgeo)
        cells.claim_rescale()
    niter = np.full(cpg, substeps, dtype=np.int32)
    indvar = np.arange(1, cpg + 1, dtype=np.int32)
    stream = torch.cuda.ExternalStream(field.stream, device=dev)

    def step():
        if cells is not None:
            cells.scatter(field, mesh.field_vol, None, *ex)
        field.diffuse(niter, indvar)
        if cells is not None:
            cells.gather(field, mesh.field_vol, want_host=False)

    def barrier():
        field.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing ------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(dev)
    sampler.start()
    l0 = field.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with torch.cuda.stream(stream):
        ev0.record()
    for k in range(args.steps):
        if cells is not None:
            cells.scatter(field, mesh.field_vol, None, *ex)
        with torch.cuda.stream(stream):
            dev_ev[k][0].record()
        field.diffuse(niter, indvar)
        with torch.cuda.stream(stream):
            dev_ev[k][1].record()
        if cells is not None:
            cells.gather(field, mesh.field_vol, want_host=False)
    with torch.cuda.stream(stream):
        ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = field.launch_count - l0
    ms = ev0.elapsed_time(ev1)
    diff_ms = sum(a.elapsed_time(b) for a, b in dev_ev)
    t = torch.tensor([ms, diff_ms], dtype=torch.float64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, diff_ms = float(t[0]), float(t[1])
    updates_per_step = n * cpg * world * substeps
    value = updates_per_step * args.steps / (ms * 1e-3)

    # ---- BASELINE.json configs[3] as a FIXED problem: 512 compounds over the N GPUs (256 / 128 / 64 columns per
    # GPU at N = 2 / 4 / 8; the N = 8 case is the weak-scaling run above).  Diffusion only, device time, max over ranks.
    fixed512 = None
    if world in (2, 4) and not args.no_extra and args.workload == "harbor1m":
        cols_f = 512 // world
        f2 = eb.Field(plan, cols_f)
        for c0 in range(0, cols_f, cpg):
            f2.upload(conc, cols=np.arange(c0 + 1, c0 + cpg + 1, dtype=np.int32))
        nit2, iv2 = np.full(cols_f, substeps, dtype=np.int32), np.arange(1, cols_f + 1, dtype=np.int32)
        st2 = torch.cuda.ExternalStream(f2.stream, device=dev)
        for _ in range(2):
            f2.diffuse(nit2, iv2)
        f2.sync(); torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        with torch.cuda.stream(st2):
            a.record()
        for _ in range(reps):
            f2.diffuse(nit2, iv2)
        with torch.cuda.stream(st2):
            b.record()
        f2.sync(); torch.cuda.synchronize()
        tf = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        f2.close()
        v512 = n * 512 * substeps / (float(tf[0]) * 1e-3)
        fixed512 = {"workload": f"configs[3] as a fixed problem: {n} points x 512 compounds x {substeps} substeps, {cols_f} columns per GPU, diffusion only",
                    "value": v512, "unit": UNIT, "ms_per_step": float(tf[0]), "columns_per_gpu": cols_f,
                    "frac_of_hbm_peak_per_gpu": BYTES_PER_UPDATE * v512 / world / 1e9 / peaks()[0]}

    # ---- parity gate on the timed configuration (outside the timed regions) -------------------
    parity = parity_gate(eb, mesh, plan, conc, substeps, dev) if (rank == 0 and not args.no_parity) else None

    # ---- end to end through the C ABI with host buffers (one-shot diffChangePar) ------------
    # N = 1: eut_diffchange_par on this GPU.  N > 1: ONE process (rank 0) hands the whole N x (64 N)
    # matrix to eut_diffchange_par_multi with all N devices -- the way R would call it -- while the other
    # ranks wait on a host-side (gloo) barrier; the N private per-rank calls of round 1 are kept as
    # `e2e_ranks`.  Pinned and pageable (plain numpy, what R's REALSXP memory is) caller buffers.
    adj = np.require(mesh.mat_in, dtype=np.int32, requirements=["F", "A"])
    nn = np.require(mesh.mat_out, dtype=np.int32, requirements=["F", "A"])
    L = eb._lib.lib()
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 5))

    def host_barrier():
        field.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def time_calls(call, swap):
        for _ in range(2):
            call()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call()
            swap()                                  # next round starts from this round's result
        return time.perf_counter() - t0

    def one_device(buf_in, buf_out):
        state = [buf_in, buf_out]

        def call():
            eb._lib.check(L.eut_diffchange_par(eb._lib.ptr(adj), adj.shape[0], eb._lib.ptr(nn), nn.shape[0],
                                               state[0], n, cpg, eb._lib.ptr(niter), eb._lib.ptr(indvar),
                                               cpg, 1, state[1], dev))
        return call, (lambda: state.reverse())

    e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": n * cpg * world * 8, "d2h_bytes_per_step": n * cpg * world * 8,
           "steps": e2e_steps, "timer": "host wall clock around the blocking calls"}
    e2e_ranks = None
    if e2e_steps:
        pin_in = torch.from_numpy(conc.T).contiguous().pin_memory()      # [cpg, n] == column-major n x cpg
        pin_out = torch.empty_like(pin_in).pin_memory()
        # (a) every rank on its own GPU with its own columns, all at once (N = 1: the headline)
        host_barrier()
        call, swap = one_device(pin_in.data_ptr(), pin_out.data_ptr())
        dt = time_calls(call, swap)
        host_barrier()
        te = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ranks_value = updates_per_step * e2e_steps / float(te[0])
        if world == 1:
            e2e.update(value=ranks_value, api="eut_diffchange_par (one-shot C ABI, pinned host buffers, graph recognised by byte comparison)")
            pg_in = np.array(conc, order="F")                             # plain malloc'ed memory, touched
            pg_out = np.zeros_like(pg_in, order="F")
            call, swap = one_device(pg_in.ctypes.data, pg_out.ctypes.data)
            dt = time_calls(call, swap)
            e2e["pageable_value"] = updates_per_step * e2e_steps / dt
            e2e["pageable_note"] = "same call with pageable (numpy) caller buffers: staged through pinned slot rings by copier threads"
            del pg_in, pg_out
        else:
            e2e_ranks = {"value": ranks_value, "unit": UNIT, "api": f"{world} processes, each eut_diffchange_par on its own GPU and columns (max over ranks)"}
            del pin_out
            # (b) one process, all GPUs (the other ranks idle: this process may use the whole host for its
            # copier pools, as an R session would)
            if rank == 0:
                try:
                    os.environ.pop("LOCAL_WORLD_SIZE", None)
                    ct = cpg * world
                    interleaved = interleave_host_memory(True)
                    big_in = torch.empty((ct, n), dtype=torch.float64).pin_memory()
                    for r in range(world):
                        big_in[r * cpg:(r + 1) * cpg].copy_(pin_in)
                    big_out = torch.empty_like(big_in).pin_memory()
                    nit_all = np.full(ct, substeps, dtype=np.int32)
                    iv_all = np.arange(1, ct + 1, dtype=np.int32)
                    devs = np.arange(world, dtype=np.int32)

                    def multi(bufs, addr):
                        def call():
                            eb._lib.check(L.eut_diffchange_par_multi(eb._lib.ptr(adj), adj.shape[0], eb._lib.ptr(nn), nn.shape[0],
                                                                     addr(bufs[0]), n, ct, eb._lib.ptr(nit_all), eb._lib.ptr(iv_all), ct, 1,
                                                                     addr(bufs[1]), eb._lib.ptr(devs), world))
                        return call, (lambda: bufs.reverse())
                    bufs = [big_in, big_out]
                    del big_in, big_out
                    call, swap = multi(bufs, lambda t: t.data_ptr())
                    dt = time_calls(call, swap)
                    e2e.update(value=updates_per_step * e2e_steps / dt,
                               api=f"eut_diffchange_par_multi: ONE process, {world} GPUs, {ct} columns in one call (pinned host buffers"
                                   + (f", pages interleaved over {numa_nodes()} NUMA nodes)" if interleaved else ")"))
                    # every GPU started from the same 64 columns: the blocks of the result must agree bit for bit
                    e2e["blocks_agree"] = all(bool(torch.equal(bufs[0][:cpg], bufs[0][r * cpg:(r + 1) * cpg])) for r in range(1, world))
                    pg = [bufs[0].numpy().copy(), np.zeros((ct, n))]
                    del bufs
                    call, swap = multi(pg, lambda t: t.ctypes.data)
                    dt = time_calls(call, swap)
                    e2e["pageable_value"] = updates_per_step * e2e_steps / dt
                    e2e["pageable_note"] = "same call with pageable (numpy) caller buffers"
                    del pg
                    interleave_host_memory(False)
                    L.eut_oneshot_reset()
                except Exception as exc:                    # the headline keeps the per-rank figure, and says so
                    interleave_host_memory(False)
                    e2e.update(value=ranks_value, api=f"{world} processes, each eut_diffchange_par on its own GPU (the one-process call failed: {type(exc).__name__}: {exc})"[:400])
            host_barrier()

    # ---- other BASELINE.json configurations, driver-visible ---------------------------------
    peak, peak_src = peaks()
    extra = {}
    if rank == 0 and not args.no_extra:
        # a leg that fails (its own parity gate included) reports the failure under its key, without a value; the
        # headline line above has its own gate and is not affected
        def leg(fn, *a):
            try:
                return fn(*a)
            except (Exception, SystemExit) as e:
                return {"error": f"{type(e).__name__}: {e}"[:400]}
        extra["square100k"] = leg(square100k_leg, eb, torch, dev, peak)
        if world > 1:
            # the other ranks idle on the host barrier below; their GPUs are driven from here
            extra["slab"] = leg(slab_leg, eb, torch, world, peak)
    if world > 1:
        host_barrier()

    if rank == 0:
        n_sub = substeps * args.steps
        launch_ms = diff_ms / n_sub
        achieved = BYTES_PER_UPDATE * n * cpg / (launch_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload)
            except Exception:
                traffic = None
        cpu = cpu_policy = None
        if world == 1 and not args.no_cpu:
            import oracle
            oracle.build()
            cpu = cpu_sample(mesh, substeps, cpg, host_threads())
            # the reference's own default: n.cores = min(10, detectCores() - 1)  (R/run_simulation.R:137-138)
            pol = max(1, min(10, host_threads() - 1))
            cpu_policy = cpu_sample(mesh, substeps, cpg, pol, target_s=6.0) if pol != host_threads() else cpu
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: non-convex polygon mesh, {n} points x {cpg * world} compounds "
                                   f"({cpg}/GPU, sharded by compound), {substeps} substeps/step, {n_cells} cells x "
                                   f"{per_cell} exchanges; step = scatter+clamp -> substeps -> gather on the resident field",
                       "l2": f"inputs larger than L2 ({2 * n * cpg * 8 / 2**20:.0f} MiB ping-pong field per GPU)",
                       "plan": {"order": int(plan.info.order), "ell_width": int(plan.info.ell_width)}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "k_substep_seg" if plan.info.segmented else "k_substep_tile" if plan.info.tiled else ("k_substep_ell" if plan.info.ell_width else "k_substep_csr"),
                         "launch_ms": launch_ms,
                         "algorithmic_bytes_per_launch": BYTES_PER_UPDATE * n * cpg,
                         "whole_step_frac": BYTES_PER_UPDATE * value / world / 1e9 / peak},
            "cpu_baseline": cpu,
            "cpu_baseline_reference_policy": cpu_policy,
            "e2e": e2e,
            "parity": parity,
            "gpu_launches": int(launches), "clocks": clocks,
            "host": {"threads": host_threads(), "numa_nodes": numa_nodes()},
            "breakdown_ms_per_step": {"diffusion_substeps": diff_ms / args.steps,
                                      "scatter_clamp_and_gather": (ms - diff_ms) / args.steps},
        }
        if e2e_ranks is not None:
            line["e2e_ranks"] = e2e_ranks
        if fixed512 is not None:
            line["configs3_fixed_512"] = fixed512
        line.update(extra)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


class QuietStdout:
    """stdout carries exactly one JSON line (the driver parses it).  Libraries print there too -- NCCL's version
    banner comes out at the first collective whatever NCCL_DEBUG says on some builds -- so file descriptor 1
    points at stderr for the whole run and the line is written to the saved descriptor at the end."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text: str):
        sys.stdout.flush()
        os.write(self.saved, (text.rstrip("\n") + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


OUT = None


def emit(line: dict):
    text = json.dumps(line)
    if OUT is not None:
        OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="harbor1m", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate (profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the square100k / slab legs")
    ap.add_argument("--opt", action="append", type=lambda s: (s.split("=")[0], int(s.split("=")[1])),
                    help="field option, e.g. --opt cpt=4")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    with QuietStdout() as out:
        OUT = out
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
        OUT = None


if __name__ == "__main__":
    main()
