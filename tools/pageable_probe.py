"""One-shot call with PAGEABLE caller buffers under different staging settings (copier threads, slot
size) next to pinned buffers.  python tools/pageable_probe.py [points] [cols]"""
import os, sys, time, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eutropia_b200 as eb
from eutropia_b200 import synth, _lib

pts = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 64
print(subprocess.run("nproc; lscpu | grep -E 'Model name|Thread|Core|Socket|NUMA'", shell=True, capture_output=True, text=True).stdout)
mesh = synth.config_mesh(f"harbor:{pts}")
n = mesh.nfields
rng = np.random.default_rng(1)
conc = np.asfortranarray(rng.uniform(0, 25, (n, cols)))
adj = np.require(mesh.mat_in, dtype=np.int32, requirements=["F", "A"])
nn = np.require(mesh.mat_out, dtype=np.int32, requirements=["F", "A"])
niter, indvar = np.full(cols, 60, np.int32), np.arange(1, cols + 1, dtype=np.int32)
L = _lib.lib()

def run(a, b, reps=5):
    def call():
        _lib.check(L.eut_diffchange_par(_lib.ptr(adj), adj.shape[0], _lib.ptr(nn), nn.shape[0], a, n, cols,
                                        _lib.ptr(niter), _lib.ptr(indvar), cols, 1, b, 0))
    call(); call()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    return (time.perf_counter() - t0) / reps * 1e3

pin_in = torch.from_numpy(np.ascontiguousarray(conc.T)).pin_memory()
pin_out = torch.empty_like(pin_in).pin_memory()
print(f"pinned: {run(pin_in.data_ptr(), pin_out.data_ptr()):.2f} ms per call", flush=True)
pg_in, pg_out = np.array(conc, order="F"), np.zeros_like(conc, order="F")
# host memcpy rates of this box
buf = np.empty_like(pg_in)
t0 = time.perf_counter(); np.copyto(buf, pg_in); dt = time.perf_counter() - t0
print(f"single-thread numpy copy: {pg_in.nbytes / dt / 1e9:.1f} GB/s")
for thr, slot in ((8, 8), (4, 8), (12, 8), (8, 16), (6, 4)):
    os.environ["EUT_COPY_THREADS"], os.environ["EUT_STAGE_SLOT_MB"] = str(thr), str(slot)
    L.eut_oneshot_reset()
    os.environ["EUT_TRACE"] = "1" if (thr, slot) == (8, 8) else ""
    if not os.environ["EUT_TRACE"]:
        del os.environ["EUT_TRACE"]
    print(f"pageable, {thr} copier threads per pool, {slot} MB slots: {run(pg_in.ctypes.data, pg_out.ctypes.data):.2f} ms per call", flush=True)
os.environ["EUT_NO_STAGING"] = "1"
L.eut_oneshot_reset()
print(f"pageable, no staging (driver path): {run(pg_in.ctypes.data, pg_out.ctypes.data, 2):.2f} ms per call")
