"""configs[1] (square100k: 100 467 points x 16 columns x 1000 substeps): columns per CTA and tile size.
python tools/square_sweep.py"""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    import eutropia_b200 as eb
    from eutropia_b200 import synth
    cpt = int(sys.argv[2])
    mesh = synth.config_mesh("square100k")
    n, c, s = mesh.nfields, 16, 1000
    plan = eb.Plan(mesh.mat_in, mesh.mat_out, device=0)
    field = eb.Field(plan, c)
    field.upload(synth.synthetic_concentrations(mesh.field_pts, c, seed=7))
    if cpt:
        field.set_option("cpt", cpt)
    niter, indvar = np.full(c, s, np.int32), np.arange(1, c + 1, dtype=np.int32)
    stream = torch.cuda.ExternalStream(field.stream, device=0)
    for _ in range(3):
        field.diffuse(niter, indvar)
    field.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(3):
        field.diffuse(niter, indvar)
    with torch.cuda.stream(stream):
        e1.record()
    field.sync()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"tile": os.environ.get("EUT_SEG_TILE", "auto"), "cpt": cpt, "tiles": int(plan.info.n_tiles), "us_per_substep": ms, "G_updates_per_s": n * c * s / ms / 1e6}))
    sys.exit(0)
for tile in ("448", "224", "112"):
    for cpt in (0, 8, 4, 2, 1):
        env = dict(os.environ, EUT_SEG_TILE=tile)
        r = subprocess.run([sys.executable, __file__, "child", str(cpt)], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
