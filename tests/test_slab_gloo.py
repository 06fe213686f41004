"""CPU suite, world_size 2 over gloo: the slab partition + halo exchange orchestration
(eutropia_b200/slab.py).  The device is replaced by a test-only backend that advances the local graph
with the CPU oracle, so what is verified is the host logic: ownership, local graph, exchange lists,
per-substep exchange.  The sharded result must be bit-identical to the unsharded oracle run."""
import contextlib
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from eutropia_b200 import mesh
from eutropia_b200.slab import SlabDriver, SlabPartition


class OracleBackend:
    """Test stand-in for GpuSlabBackend: same interface, local substeps by the CPU oracle."""

    def __init__(self, part, local_matrix):
        self.part = part
        self.c = np.array(local_matrix, dtype=np.float64, order="F")
        self._send = torch.zeros(max(1, part.send_all.size * self.c.shape[1]), dtype=torch.float64)
        self._recv = torch.zeros(max(1, part.recv_all.size * self.c.shape[1]), dtype=torch.float64)

    def substep(self, cols):
        self.c = oracle.diffChangePar(self.part.mat_in, self.part.mat_out, self.c, np.ones(len(cols)), cols, 1)

    def pack(self, cols):
        n = self.part.send_all.size * len(cols)
        vals = self.c[np.ix_(self.part.send_all - 1, np.asarray(cols) - 1)]       # [point][column]
        self._send[:n] = torch.from_numpy(np.ascontiguousarray(vals).ravel())
        return self._send

    def recv_buffer(self):
        return self._recv

    def unpack(self, cols):
        n = self.part.recv_all.size * len(cols)
        vals = self._recv[:n].numpy().reshape(self.part.recv_all.size, len(cols))
        self.c[np.ix_(self.part.recv_all - 1, np.asarray(cols) - 1)] = vals

    def comm_context(self):
        return contextlib.nullcontext()

    def download_own(self):
        return self.c[self.part.own_local]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir, layers, width):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        m = mesh.make_mesh(mesh.harbor_polygon(width), 1.0, layers)
        rng = np.random.default_rng(7)
        conc = np.asfortranarray(rng.uniform(0, 25, (m.nfields, 4)))
        niter, indvar = [9, 4, 0, 6], [1, 2, 3, 4]
        part = SlabPartition(m.mat_in, m.mat_out, m.nfields, world, rank)
        backend = OracleBackend(part, part.local_rows(conc))
        driver = SlabDriver(part, backend, dist)
        driver.diffuse(niter, indvar)
        own = backend.download_own()
        ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 1)
        own_global = part.local_global[part.own_local]
        ok = np.array_equal(own, ref[own_global])
        # second round after an out-of-band change of owned values (what the cell scatter does)
        backend.c[part.own_local, 0] += 1.0
        ref2 = ref.copy(); ref2[:, 0] += 1.0
        driver.diffuse([3, 0, 0, 0], [1, 2, 3, 4])
        ref2 = oracle.diffChangePar(m.mat_in, m.mat_out, ref2, [3], [1], 1)
        ok2 = np.array_equal(backend.download_own(), ref2[own_global])
        np.savez(os.path.join(outdir, f"r{rank}.npz"), ok=ok, ok2=ok2, n_own=own_global.size,
                 n_ghost=part.n_local - own_global.size, peers=np.array(part.peers),
                 n_send=part.send_all.size, n_recv=part.recv_all.size, own_global=own_global)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,layers,width", [(2, 3, 60.0), (3, 2, 80.0)])
def test_slab_sharded_run_is_bit_identical(world, layers, width):
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d, layers, width), nprocs=world, join=True)
        res = [np.load(os.path.join(d, f"r{r}.npz")) for r in range(world)]
    assert all(bool(r["ok"]) and bool(r["ok2"]) for r in res)
    owned = np.concatenate([r["own_global"] for r in res])
    assert np.array_equal(np.sort(owned), np.arange(owned.size))      # every point owned exactly once
    sizes = [int(r["n_own"]) for r in res]
    assert max(sizes) - min(sizes) <= 2
    assert all(int(r["n_ghost"]) > 0 and int(r["n_send"]) > 0 and int(r["n_recv"]) > 0 for r in res)
    # what one rank sends to a peer is what the peer expects from it
    assert sum(int(r["n_send"]) for r in res) == sum(int(r["n_recv"]) for r in res)
    # slabs are thin interfaces, not random scatterings: ghosts are a small fraction of a slab
    assert all(int(r["n_ghost"]) < 0.35 * int(r["n_own"]) for r in res)


def test_partition_local_graph_properties():
    m = mesh.make_mesh("Rectangle_50_40", 1.0, 3)
    parts = [SlabPartition(m.mat_in, m.mat_out, m.nfields, 4, r) for r in range(4)]
    for p in parts:
        # every in-edge of an owned point is present, in the global row order
        own_global = p.local_global[p.own_local] + 1
        sel = np.isin(m.mat_in[:, 1], own_global)
        assert np.array_equal(p.local_global[p.mat_in[:, 0] - 1] + 1, m.mat_in[sel, 0])
        assert np.array_equal(p.local_global[p.mat_in[:, 1] - 1] + 1, m.mat_in[sel, 1])
        # ghosts: no in-edges, outflow weight zero
        ghosts = np.flatnonzero(~p.is_own) + 1
        assert not np.isin(p.mat_in[:, 1], ghosts).any()
        assert np.all(p.mat_out[ghosts - 1, 1] == 1)
        assert np.array_equal(p.mat_out[p.own_local, 1], m.mat_out[own_global - 1, 1])
