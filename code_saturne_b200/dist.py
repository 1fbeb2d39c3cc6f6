"""Rank plumbing for the multi-GPU path (host side only, no arithmetic).

Two ways to run R ranks, matching csb200.h's peer-memory communicator:

* one process per GPU (torchrun / mpirun): `pcomm_from_torch_dist()` exchanges the
  cudaIpc handles of the arenas with torch.distributed and returns an `api.PComm`;
* R host threads in one process (`run_ranks_in_threads`): every thread is a rank with
  its own library context and stream; with a single visible GPU all ranks share it
  (how the `-m gpu` tests exercise the distributed kernels on a one-GPU box).

`build_rank` turns mesh_cartesian.cube_poisson_local output into device objects.
"""
from __future__ import annotations

import threading
import traceback

import numpy as np

from . import api
from .mesh_cartesian import cube_poisson_local


def pcomm_from_torch_dist(n_ghost, device=None):
    """api.PComm over the default torch.distributed group (one process per GPU)."""
    import torch.distributed as dist
    n_ranks, rank = dist.get_world_size(), dist.get_rank()
    arena, window = api.PComm.default_sizes(n_ghost)

    def all_gather(b):
        out = [None] * n_ranks
        dist.all_gather_object(out, b)
        return out

    return api.PComm.from_handles(n_ranks, rank, arena, window, all_gather)


class RankObjects:
    """Device objects of one rank of a partitioned cube system."""

    def __init__(self, system, halo_desc, gids, pcomm, rank):
        self.system, self.halo_desc, self.gids = system, halo_desc, gids
        self.rank, self.pcomm = rank, pcomm
        self.matrix = api.Matrix.from_system(system, share_host=True)   # system outlives it
        self.halo = None
        if pcomm is not None and pcomm.n_ranks > 1:
            self.halo = api.Halo(halo_desc, rank, pcomm=pcomm)
            self.matrix.set_halo(self.halo)

    def multigrid(self):
        mg = api.Multigrid()
        if self.halo is not None:
            mg.set_pcomm(self.pcomm)
        return mg

    def pcg(self, n_max_iter):
        c = api.SlesIt(api.SlesItType.PCG, n_max_iter)
        if self.halo is not None:
            c.set_pcomm(self.pcomm)
        return c

    def destroy(self):
        self.matrix.destroy()
        if self.halo is not None:
            self.halo.destroy()


def build_rank(N, n_ranks, rank, pcomm, bc="neumann"):
    s, h, gids = cube_poisson_local(N, n_ranks, rank, bc)
    return RankObjects(s, h, gids, pcomm, rank)


def run_ranks_in_threads(n_ranks, fn, device_of_rank=None, n_ghost_max=1 << 16):
    """Run fn(rank, pcomm) on n_ranks host threads (one library context each) and return
    the list of results.  device_of_rank(rank) -> CUDA device (default: all on 0).
    Raises the first exception of any rank."""
    group = api.PComm.group_create(n_ranks)
    arena, window = api.PComm.default_sizes(n_ghost_max)
    results = [None] * n_ranks
    errors = [None] * n_ranks

    def body(rank):
        pc = None
        try:
            api.init(device_of_rank(rank) if device_of_rank else 0)
            pc = api.PComm.in_group(group, n_ranks, rank, arena, window)
            results[rank] = fn(rank, pc)
        except BaseException:   # noqa: BLE001 - reported to the caller
            errors[rank] = traceback.format_exc()
        finally:
            if pc is not None and errors[rank] is None:
                pc.destroy()

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(n_ranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    api.PComm.group_destroy(group)
    failed = [(r, e) for r, e in enumerate(errors) if e is not None]
    if failed:
        raise RuntimeError("\n".join(f"rank {r} failed:\n{e}" for r, e in failed))
    return results


def gather_global(values_per_rank, gids_per_rank, n_global):
    """Assemble per-rank local vectors into the global numbering."""
    out = np.zeros(n_global)
    for v, g in zip(values_per_rank, gids_per_rank):
        out[g[:len(v)]] = v
    return out
