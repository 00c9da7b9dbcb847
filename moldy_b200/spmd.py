"""Replicated-data multi-GPU layer, modelled on parallel.c's par_rsum/par_dsum
(src/parallel.c:549-588; call sites src/accel.c:531-535).

The sums themselves are kernels of the C library (moldy_b200/csrc/mdb_peer.cu: flag barrier, structure-factor
all-reduce, force reduce-scatter and all-gather over NVLink/NVSwitch peer memory); this module is the Python front end:
`SpmdForces` = one process per GPU (torchrun; CUDA-IPC windows exchanged through torch.distributed, which is used for that
hand-shake only), `PeerGroup` = all ranks as engines of one process (what the library does itself behind
force_calc()/ewald()/eval_forces() with MOLDY_B200_DEVICES, moldy_b200/csrc/mdb_group.cu).  `combine`/`recip_sites` and
`SpmdForces(nccl=True)` keep the round-1 path (packed NCCL all-reduce) for comparison and for the gloo tests.

Every rank holds all N positions and the full cell-sorted SoA (the cell build is replicated: ~0.1 ms at 10^6 sites);
rank r of P owns

  * real space:  the r-th contiguous slice of the BATCHES of the cell-sorted sites (default kernel: the reference's half
    list with Newton's third law, so its partial force array also holds what its batches add to the neighbours j of
    other slices -- the partial arrays of all ranks are summed, as par_rsum does; pair mode 3 is the full stencil,
    owner-computes),
  * k-space:     its slice of the (charged) sites for ALL k-vectors -- the manual's "RIL" scheme
    (src/moldy.tex:3441-3466): the 8*nslots structure-factor sums are all-reduced between the two passes (0.6-1.2 MB),
    so both the structure-factor pass and the back-projection scale with 1/P.  (Moldy's shipped scheme, a block of
    k-vectors per rank for all sites, src/ewald.c:495-496, is what the library does when it is driven through
    force_calc()/ewald() with ithread/nthreads, because no exchange is available there before the final sums.)
  * the result:  the r-th slice of the sites of every force row (reduce-scatter; the 16 scalars are summed by every rank).

Real space follows the reference's scheme (cells `icell = ithread mod nthreads`, src/force.c:856) with a different but
equally disjoint assignment.  The sums are formed in rank order on every rank, so all ranks hold identical bits, which
is what Moldy's DESYNC check (src/main.c:262-273) relies on.  Constants the reference adds on rank 0 only (eintra,
self/sheet energy) are not part of the block.
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


def result_bounds(nsites: int, world: int):
    """Ownership of the RESULT in the peer layer (mdb_peer.cu, default bounds): rank r reduces the original sites
    [b[r], b[r+1]) of every force row, every rank sums the 16 scalars."""
    return [nsites * r // world for r in range(world + 1)]


def molecule_bounds(nmols_per_species, nsites_per_species, world: int):
    """eval_forces() on several ranks (mdb_group.cu): equal numbers of SITES per rank, cut at molecule boundaries.
    Returns (molecule bounds [world+1], site bounds [world+1])."""
    site_of_mol = [0]
    for nm, ns in zip(nmols_per_species, nsites_per_species):
        for _ in range(nm):
            site_of_mol.append(site_of_mol[-1] + ns)
    n, nmols = site_of_mol[-1], len(site_of_mol) - 1
    import bisect
    mb = [min(bisect.bisect_left(site_of_mol, n * r // world), nmols) for r in range(world + 1)]
    mb[world] = nmols
    return mb, [site_of_mol[m] for m in mb]


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
    """Host-buffer front end of the multi-GPU force evaluation, one process per GPU (bench.py `e2e`, N>1).

    The sums run on the library's own peer-memory kernels (mdb_peer.cu): every rank uploads only ITS slice of the
    site rows over its own PCIe link, the slices are all-gathered over NVLink, and after the step every rank
    downloads only its slice of the reduced forces (reduce-scatter) plus the 16 scalars.  `nccl=True` keeps the
    round-1 path (full upload on every rank, NCCL all-reduce, full download) for comparison."""

    def __init__(self, ms: MoldySystem, rank: int, world: int, device: int, nccl: bool = False):
        self.ms, self.rank, self.world, self.nccl = ms, rank, world, nccl
        torch.cuda.set_device(device)
        self.eng = lib.Engine(device)
        self.eng.configure(ms)
        self.eng.set_partition(rank, world)
        self.n = n = ms.nsites
        self.h_force = torch.zeros((3, n), dtype=torch.float64).pin_memory()
        self.h_scal = torch.zeros(16, dtype=torch.float64).pin_memory()
        if nccl:
            self.d_xyz = torch.empty((3, n), dtype=torch.float64, device="cuda")
            self.d_out = torch.zeros(self.eng.out_doubles(), dtype=torch.float64, device="cuda")
            self.d_psum = torch.zeros(self.eng.recip_sum_doubles(), dtype=torch.float64, device="cuda")
            self.h_out = torch.empty(self.eng.out_doubles(), dtype=torch.float64).pin_memory()
            self.eng.set_sites_device(self.d_xyz[0].data_ptr(), self.d_xyz[1].data_ptr(), self.d_xyz[2].data_ptr())
            self.lo, self.hi = 0, n
            return
        self.peer = lib.Peer(self.eng, rank, world)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, self.peer.handle())
            self.peer.open(handles)
            dist.barrier()
        self.lo, self.hi = self.peer.slice()

    def step(self, host_sites: torch.Tensor):
        """host_sites: pinned [3,N] float64 (only this rank's slice is read).  Returns (forces [3,N] pinned -- this
        rank's slice [lo,hi) is valid --, scalars[16] = pe_real, pe_recip, stress[9], lo, hi)."""
        st = torch.cuda.current_stream().cuda_stream
        if self.nccl:
            self.d_xyz.copy_(host_sites, non_blocking=True)
            self.eng.set_sites_device(self.d_xyz[0].data_ptr(), self.d_xyz[1].data_ptr(), self.d_xyz[2].data_ptr(), st)
            self.eng.zero_out(self.d_out.data_ptr(), st)
            self.eng.build_cells(st)
            self.eng.force_real(self.d_out.data_ptr(), st)
            recip_sites(self.eng, self.d_psum, self.d_out, st)
            combine(self.d_out)
            self.h_out.copy_(self.d_out, non_blocking=True)
            torch.cuda.synchronize()
            n = self.n
            self.h_force.copy_(self.h_out[:3 * n].view(3, n))
            self.h_scal.copy_(self.h_out[3 * n:])
            return self.h_force, self.h_scal, 0, n
        p = self.peer
        rows = [host_sites[a].data_ptr() for a in range(3)]
        p.sites_host_slice(rows[0], rows[1], rows[2], st)
        p.barrier(st)
        p.sites_gather(st)
        p.step(lib.REAL | lib.RECIP, False, st)
        f = [self.h_force[a].data_ptr() for a in range(3)]
        p.read_slice_host(f[0], f[1], f[2], self.h_scal.data_ptr(), st)
        return self.h_force, self.h_scal, self.lo, self.hi

    def bytes_per_step(self):
        """(H2D, D2H) bytes this rank moves per step."""
        if self.nccl:
            return 3 * self.n * 8, (3 * self.n + 16) * 8
        return 3 * (self.hi - self.lo) * 8, (3 * (self.hi - self.lo) + 16) * 8

    def close(self):
        if not self.nccl:
            self.peer.close()
        self.eng.close()


class PeerGroup:
    """All ranks as engines of THIS process (one host thread drives them phase by phase): the form the C library uses
    behind force_calc()/ewald()/eval_forces() with MOLDY_B200_DEVICES, and what the single-box GPU tests exercise.
    devices may repeat (several ranks on one GPU: the kernels are the same, the windows are then local)."""

    def __init__(self, ms: MoldySystem, devices):
        self.ms, self.world, self.n = ms, len(devices), ms.nsites
        self.engs, self.peers, self.streams = [], [], []
        for r, d in enumerate(devices):
            torch.cuda.set_device(d)
            e = lib.Engine(d)
            e.configure(ms)
            self.engs.append(e)
            self.peers.append(lib.Peer(e, r, self.world))
            self.streams.append(torch.cuda.Stream(device=d))
        lib.Peer.connect(self.peers)
        self.devices = list(devices)

    def _each(self, fn):
        cur = torch.cuda.current_device()
        try:
            for r, p in enumerate(self.peers):
                fn(p, self.streams[r].cuda_stream)
        finally:
            torch.cuda.set_device(cur)         # the library selects each rank's device (cudaSetDevice)

    def set_sites(self, site_block: np.ndarray):
        blk = np.ascontiguousarray(site_block[:, :self.n])
        self._keep = blk
        self._each(lambda p, st: p.sites_host_all(blk, st))

    def step(self, what=lib.REAL | lib.RECIP, gather=True):
        self._each(lambda p, st: p.phase_a(what, st))
        self._each(lambda p, st: p.barrier(st))
        self._each(lambda p, st: p.phase_b(what, st))
        self._each(lambda p, st: p.barrier(st))
        self._each(lambda p, st: p.phase_c(st))
        if gather:
            self._each(lambda p, st: p.barrier(st))
            self._each(lambda p, st: p.phase_d(st))

    def result(self, rank=0) -> np.ndarray:
        """The reduced block of one rank as a host array (after step(gather=True): complete on every rank)."""
        cur = torch.cuda.current_device()
        torch.cuda.set_device(self.devices[rank])
        try:
            return self.engs[rank].read_out(self.peers[rank].result_ptr(), self.streams[rank].cuda_stream)
        finally:
            torch.cuda.set_device(cur)

    def synchronize(self):
        for s in self.streams:
            s.synchronize()
        for r, p in enumerate(self.peers):
            if p.error(self.streams[r].cuda_stream):
                raise RuntimeError(f"peer barrier of rank {r} timed out")

    def close(self):
        self.synchronize()
        for p in self.peers:
            p.close()
        for e in self.engs:
            e.close()
