"""Slab leg (2 M points per GPU x 64 columns x 60 substeps, all visible GPUs, one process) under the
driver's switches, one subprocess per variant.  python tools/slab_variants.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    import eutropia_b200 as eb
    import bench
    peak, _ = bench.peaks()
    r = bench.slab_leg(eb, torch, torch.cuda.device_count(), peak)
    print(json.dumps({k: r[k] for k in ("value", "ms_per_substep", "frac_of_hbm_peak_per_gpu", "gpu_launches", "bit_identical_to_unsharded")}))
    sys.exit(0)
VARIANTS = [("two groups (default)", {}), ("three groups", {"EUT_SLAB_GROUPS": "3"}), ("four groups", {"EUT_SLAB_GROUPS": "4"}),
            ("four groups, staggered", {"EUT_SLAB_GROUPS": "4", "EUT_SLAB_STAGGER": "1"})]
for name, env in VARIANTS:
    r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
    line = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    print(name, "->", line[-1] if line else (r.stdout[-300:] + r.stderr[-300:]), flush=True)
