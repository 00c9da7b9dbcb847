"""Replicated-data multi-GPU layer, modelled on parallel.c's par_rsum/par_dsum
(src/parallel.c:549-588; call sites src/accel.c:531-535).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Every rank
holds all N positions and the full cell-sorted SoA (the cell build is
replicated: ~0.1 ms at 10^6 sites); rank r of P owns

  * real space:  the r-th contiguous slice of the cell-sorted sites (full
    stencil, owner-computes: its partial force array is non-zero only there),
  * k-space:     its slice of the (charged) sites for ALL k-vectors -- the manual's
    "RIL" scheme (src/moldy.tex:3441-3466): the 8*nslots structure-factor sums are
    all-reduced between the two passes (0.6-1.2 MB), so both the structure-factor
    pass and the back-projection scale with 1/P.  (Moldy's shipped scheme, a block
    of k-vectors per rank for all sites, src/ewald.c:495-496, is what the library
    does when it is driven through force_calc()/ewald() with ithread/nthreads,
    because no exchange is available there before the final sums.)

Real space follows the reference's scheme (cells `icell = ithread mod nthreads`,
src/force.c:856) with a different but equally disjoint assignment.  The partial [forces | pe | stress] blocks are
combined by ONE packed all-reduce (the reference issues three,
src/accel.c:532-534); NCCL returns bit-identical sums on all ranks, which is what
Moldy's DESYNC check (src/main.c:262-273) relies on.  Constants the reference
adds on rank 0 only (eintra, self/sheet energy) are not part of this block.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import lib
from .systems import MoldySystem


def combine(out: torch.Tensor, group=None):
    """par_dsum(pe,2) + par_rsum(stress,9) + par_rsum(site_force,3*nsarray) as one
    in-place SUM all-reduce of the packed result block (works on NCCL and gloo)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def site_slice(nsites: int, rank: int, world: int):
    """Real-space ownership: contiguous slice of the cell-sorted site array."""
    return nsites * rank // world, nsites * (rank + 1) // world


def column_owner(vpos: int, world: int) -> int:
    """k-space ownership: position of an (h,k) column in the l-count-sorted list, mod P."""
    return vpos % world


def recip_sites(eng, d_psum: torch.Tensor, d_out: torch.Tensor, stream: int, group=None):
    """k-space with the site partition: pass 1 on own sites, all-reduce of the structure-factor
    sums, pass 2 (energy/stress on rank 0, forces on own sites)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        eng.force_recip(d_out.data_ptr(), stream)
        return
    eng.recip_partial(d_psum.data_ptr(), stream)
    dist.all_reduce(d_psum, op=dist.ReduceOp.SUM, group=group)
    eng.recip_finish(d_psum.data_ptr(), d_out.data_ptr(), stream)


class SpmdForces:
    """Host-buffer front end of the multi-GPU force evaluation (bench.py `e2e`, N>1)."""

    def __init__(self, ms: MoldySystem, rank: int, world: int, device: int):
        self.ms, self.rank, self.world = ms, rank, world
        torch.cuda.set_device(device)
        self.eng = lib.Engine(device)
        self.eng.configure(ms)
        self.eng.set_partition(rank, world)
        self.n = ms.nsites
        self.d_xyz = torch.empty((3, self.n), dtype=torch.float64, device="cuda")
        self.d_out = torch.zeros(self.eng.out_doubles(), dtype=torch.float64, device="cuda")
        self.d_psum = torch.zeros(self.eng.recip_sum_doubles(), dtype=torch.float64, device="cuda")
        self.h_out = torch.empty(self.eng.out_doubles(), dtype=torch.float64).pin_memory()
        self.eng.set_sites_device(self.d_xyz[0].data_ptr(), self.d_xyz[1].data_ptr(), self.d_xyz[2].data_ptr())

    def step(self, host_sites: torch.Tensor) -> np.ndarray:
        """host_sites: pinned [3,N] float64.  Returns the combined block on the host."""
        st = torch.cuda.current_stream().cuda_stream
        self.d_xyz.copy_(host_sites, non_blocking=True)
        self.eng.set_sites_device(self.d_xyz[0].data_ptr(), self.d_xyz[1].data_ptr(), self.d_xyz[2].data_ptr(), st)
        self.eng.zero_out(self.d_out.data_ptr(), st)
        self.eng.build_cells(st)
        self.eng.force_real(self.d_out.data_ptr(), st)
        recip_sites(self.eng, self.d_psum, self.d_out, st)
        combine(self.d_out)
        self.h_out.copy_(self.d_out, non_blocking=True)
        torch.cuda.synchronize()
        return self.h_out.numpy()
