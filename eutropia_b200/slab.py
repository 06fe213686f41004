"""Spatial slab sharding of one field across the GPUs of a box (BASELINE.json configs[4]).

The reference runs on one host and has no counterpart; the `diffChange` signature carries no
coordinates, so slabs are derived from the graph: the tile planner's internal order (compact tiles,
ordered by coarse level band) is cut into `world` contiguous ranges.  Rank r then works on a *local*
graph = its own points plus ghost copies of the remote in-neighbours:

  * local ids follow ascending global id, so raster rows stay runs of consecutive ids and the local
    plan tiles like the global one;
  * local `mat.in` = the rows of the global `mat.in` whose `to` is owned, in the global row order --
    every owned point therefore sums exactly the same values in exactly the same order as in the
    unsharded run, and the result is bit-identical to it;
  * ghosts have no in-edges and `n_neighbors = 1` (outflow weight 0): a local substep leaves them
    alone, the halo exchange overwrites them with the owner's new value after every substep.

`SlabDriver` orchestrates substep -> pack -> send/recv -> unpack through `torch.distributed`
(NCCL over NVLink on GPUs; the host logic is also exercised with gloo on CPU in
tests/test_slab_gloo.py, where a test backend stands in for the device).
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib
from .session import Field, Plan

__all__ = ["SlabPartition", "GpuSlabBackend", "SlabDriver"]


class SlabPartition:
    """Who owns what, the local graph of `rank`, and the exchange lists with every peer."""

    def __init__(self, mat_in, mat_out, n_points: int, world: int, rank: int, global_perm=None):
        mat_in = np.ascontiguousarray(np.asarray(mat_in, dtype=np.int32).reshape(-1, 2))
        mat_out = np.ascontiguousarray(np.asarray(mat_out, dtype=np.int32).reshape(-1, 2))
        n = int(n_points)
        self.world, self.rank, self.n_global = int(world), int(rank), n
        if global_perm is None:
            plan = Plan(mat_in, mat_out, n_points=n, order=_lib.ORDER_TILED, host_only=True)
            global_perm = plan.perm()
            plan.close()
        perm = np.asarray(global_perm, dtype=np.int64)
        # contiguous ranges of the internal order, equal point counts
        bounds = np.linspace(0, n, self.world + 1).round().astype(np.int64)
        owner = (np.searchsorted(bounds, perm, side="right") - 1).astype(np.int32)   # by global id - 1
        self.owner = owner
        frm, to = mat_in[:, 0].astype(np.int64) - 1, mat_in[:, 1].astype(np.int64) - 1
        own_to = owner[to] == self.rank
        own_ids = np.flatnonzero(owner == self.rank)
        ghost_ids = np.unique(frm[own_to & (owner[frm] != self.rank)])
        local_ids = np.union1d(own_ids, ghost_ids)                    # ascending global id (0-based)
        self.local_global = local_ids                                   # local index -> global id - 1
        self.n_local = int(local_ids.shape[0])
        self.is_own = owner[local_ids] == self.rank
        self.own_local = np.flatnonzero(self.is_own)                   # local indices of owned points
        # local graph
        lf = np.searchsorted(local_ids, frm[own_to]) + 1
        lt = np.searchsorted(local_ids, to[own_to]) + 1
        self.mat_in = np.ascontiguousarray(np.stack([lf, lt], axis=1).astype(np.int32))
        nn_by_id = np.ones(n, dtype=np.int32)
        nn_by_id[mat_out[:, 0] - 1] = mat_out[:, 1]
        nn = np.where(self.is_own, nn_by_id[local_ids], 1).astype(np.int32)
        self.mat_out = np.ascontiguousarray(
            np.stack([np.arange(1, self.n_local + 1, dtype=np.int32), nn], axis=1))
        # exchange lists, both sides sorted by global id so they line up without negotiation
        self.peers, self.send_ids, self.recv_ids = [], [], []
        self.send_off, self.recv_off = [0], [0]
        for peer in range(self.world):
            if peer == self.rank:
                continue
            recv_g = ghost_ids[owner[ghost_ids] == peer]
            send_g = np.unique(frm[(owner[to] == peer) & (owner[frm] == self.rank)])
            if recv_g.size == 0 and send_g.size == 0:
                continue
            self.peers.append(peer)
            self.send_ids.append(np.searchsorted(local_ids, send_g) + 1)     # 1-based local ids
            self.recv_ids.append(np.searchsorted(local_ids, recv_g) + 1)
            self.send_off.append(self.send_off[-1] + send_g.size)
            self.recv_off.append(self.recv_off[-1] + recv_g.size)
        cat = lambda xs: (np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)).astype(np.int32)
        self.send_all, self.recv_all = cat(self.send_ids), cat(self.recv_ids)

    def local_rows(self, global_matrix):
        """Rows of a global N x C matrix that this rank holds (own + ghosts), column-major."""
        return np.asfortranarray(np.asarray(global_matrix)[self.local_global])


def exchange_halo(part: SlabPartition, send, recv, n_cols: int, dist):
    """One halo exchange: `send` / `recv` are flat tensors laid out [point][column]; the points of a
    peer are a contiguous slice.  Works for CUDA tensors over NCCL and CPU tensors over gloo."""
    ops = []
    for k, peer in enumerate(part.peers):
        s0, s1 = part.send_off[k] * n_cols, part.send_off[k + 1] * n_cols
        r0, r1 = part.recv_off[k] * n_cols, part.recv_off[k + 1] * n_cols
        if s1 > s0:
            ops.append(dist.P2POp(dist.isend, send[s0:s1], peer))
        if r1 > r0:
            ops.append(dist.P2POp(dist.irecv, recv[r0:r1], peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class GpuSlabBackend:
    """The slab of this rank resident on its GPU (local Plan + Field, halo kernels)."""

    def __init__(self, part: SlabPartition, n_cols: int, device: int, host_staged: bool = False):
        """`host_staged`: exchange through pinned host copies of the buffers (for a process group
        without device support, e.g. gloo when several ranks share one GPU in a test)."""
        import torch
        self.torch = torch
        self.part, self.n_cols, self.device = part, int(n_cols), int(device)
        self.host_staged = bool(host_staged)
        self.plan = Plan(part.mat_in, part.mat_out, n_points=part.n_local, device=device)
        self.field = Field(self.plan, n_cols)
        self.field.halo_set(part.send_all, part.recv_all)
        dev = torch.device("cuda", device)
        self.send = torch.zeros(max(1, part.send_all.size * n_cols), dtype=torch.float64, device=dev)
        self.recv = torch.zeros(max(1, part.recv_all.size * n_cols), dtype=torch.float64, device=dev)
        if self.host_staged:
            self.h_send = torch.zeros_like(self.send, device="cpu")
            self.h_recv = torch.zeros_like(self.recv, device="cpu")
        self.stream = torch.cuda.ExternalStream(self.field.stream, device=dev)
        # overlap of the exchange with the interior tiles: boundary tiles first, their results travel on
        # a second stream while the interior tiles run (needs the segment plan's tile lists)
        self.overlap = (not self.host_staged) and bool(self.plan.info.segmented) and os.environ.get("EUT_SLAB_OVERLAP", "1") != "0"
        self.comm_stream = torch.cuda.Stream(device=dev) if self.overlap else None

    def close(self):
        """Release device objects in a safe order (torch tensors first, then the field's stream)."""
        self.sync()
        self.torch.cuda.synchronize()
        self.send = self.recv = self.stream = None
        self.field.close()
        self.plan.close()

    def upload(self, local_matrix):
        self.field.upload(local_matrix)

    def substep(self, cols):
        self.field.diffuse(np.ones(len(cols), dtype=np.int32), cols)

    def substep_overlapped(self, cols, exchange):
        """One substep with the halo exchange hidden behind the interior tiles: boundary tiles (and the
        buffer flip) on the field's stream; pack -> `exchange()` -> unpack on the comm stream; interior
        tiles on the field's stream meanwhile; the field's stream finally waits for the unpack."""
        torch = self.torch
        ones = np.ones(len(cols), dtype=np.int32)
        self.field.set_option("phase", 1)
        self.field.diffuse(ones, cols)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self.comm_stream.wait_event(ev)
        self.field.set_halo_stream(self.comm_stream.cuda_stream)
        with torch.cuda.stream(self.comm_stream):
            self.field.halo_pack(cols, self.send.data_ptr())
            exchange(self.send, self.recv)
            self.field.halo_unpack(cols, self.recv.data_ptr())
            done = torch.cuda.Event()
            done.record(self.comm_stream)
        self.field.set_halo_stream(None)
        self.field.set_option("phase", 2)
        self.field.diffuse(ones, cols)
        self.field.set_option("phase", 0)
        self.stream.wait_event(done)

    def pack(self, cols):
        self.field.halo_pack(cols, self.send.data_ptr())
        if self.host_staged:
            self.h_send.copy_(self.send)          # blocking copy on the field's stream
            return self.h_send
        return self.send

    def unpack(self, cols):
        if self.host_staged:
            self.recv.copy_(self.h_recv)
        self.field.halo_unpack(cols, self.recv.data_ptr())

    def recv_buffer(self):
        return self.h_recv if self.host_staged else self.recv

    def comm_context(self):
        # NCCL orders its work against torch's *current* stream: make that the field's stream
        return self.torch.cuda.stream(self.stream)

    def download_own(self):
        full = self.field.download()
        return full[self.part.own_local]

    def sync(self):
        self.field.sync()


class SlabDriver:
    """diffChangePar semantics on a slab-sharded field: `niter[i]` substeps on column `indvar[i]`,
    one halo exchange after every substep for the columns that moved."""

    def __init__(self, part: SlabPartition, backend, dist):
        self.part, self.backend, self.dist = part, backend, dist

    def refresh_halo(self, cols):
        """Bring the ghost copies of `cols` up to date (after anything but a substep changed owned
        values, e.g. the cell scatter)."""
        cols = np.asarray(cols, dtype=np.int32)
        if cols.size == 0 or not self.part.peers:
            return
        with self.backend.comm_context():
            send = self.backend.pack(cols)
            exchange_halo(self.part, send, self.backend.recv_buffer(), int(cols.size), self.dist)
            self.backend.unpack(cols)

    def diffuse(self, niter, indvar, refresh: bool = True):
        niter = np.trunc(np.asarray(niter, dtype=np.float64)).astype(np.int64)
        indvar = np.trunc(np.asarray(indvar, dtype=np.float64)).astype(np.int32)
        max_it = int(niter.max()) if niter.size else 0
        if refresh:
            self.refresh_halo(indvar[niter > 0])
        with self.backend.comm_context():
            for s in range(max_it):
                cols = indvar[niter > s]
                if cols.size == 0:
                    break
                if self.part.peers and getattr(self.backend, "overlap", False):
                    n_c = int(cols.size)
                    self.backend.substep_overlapped(
                        cols, lambda send, recv: exchange_halo(self.part, send, recv, n_c, self.dist))
                    continue
                self.backend.substep(cols)
                if self.part.peers:
                    send = self.backend.pack(cols)
                    exchange_halo(self.part, send, self.backend.recv_buffer(), int(cols.size), self.dist)
                    self.backend.unpack(cols)
