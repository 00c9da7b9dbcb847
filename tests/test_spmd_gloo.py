"""CPU tests (-m "not gpu"): the N>1 path of the replicated-data layer on the
`gloo` backend, world_size 2.  Each rank evaluates its slice of the step with the
oracle standing in for the GPU kernels (same ithread/nthreads partition contract),
packs the [forces | pe | stress] block the way the engine lays it out in HBM, and
moldy_b200.spmd.combine() all-reduces it -- the par_dsum/par_rsum of
src/accel.c:531-535.  Every rank must end up with the single-rank result,
bit-identical across ranks (Moldy's DESYNC check, src/main.c:262-273)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from moldy_b200 import spmd
    from oracle import port
    from tests import cases
    ms = cases.GOLDEN_CASES["tips2"]()
    n = ms.nsites
    part = port.run(ms, ithread=rank, nthreads=world)
    block = torch.zeros(3 * n + 16, dtype=torch.float64)
    block[:3 * n] = torch.from_numpy(part["force"].reshape(-1))
    block[3 * n:3 * n + 2] = torch.from_numpy(part["pe"])
    block[3 * n + 2:3 * n + 11] = torch.from_numpy(part["stress"].reshape(-1))
    spmd.combine(block)
    gathered = [torch.zeros_like(block) for _ in range(world)]
    dist.all_gather(gathered, block)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        q.put((block.numpy().copy(), same))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_reproduces_single_rank():
    from oracle import port
    from tests import cases
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    block, same = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same, "ranks disagree after the all-reduce"
    ms = cases.GOLDEN_CASES["tips2"]()
    n = ms.nsites
    whole = port.run(ms)
    assert cases.rel_rms(block[:3 * n].reshape(3, n), whole["force"]) < 1e-13
    assert np.allclose(block[3 * n:3 * n + 2], whole["pe"], rtol=1e-12)
    assert np.allclose(block[3 * n + 2:3 * n + 11].reshape(3, 3), whole["stress"], rtol=1e-11, atol=1e-9)


def _worker_evalf(rank, world, port_no, q):
    """eval_forces under SPMD: partial site forces -> combine() -> the molecular-frame tail on EVERY rank."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from moldy_b200 import spmd
    from oracle import evalf, port
    from tests import cases
    ms = cases.GOLDEN_CASES["mgcl2"]()
    ms.control.surface_dipole = 1
    n = ms.nsites
    part = port.run(ms, ithread=rank, nthreads=world)
    block = torch.zeros(3 * n + 16, dtype=torch.float64)
    block[:3 * n] = torch.from_numpy(part["force"].reshape(-1))
    block[3 * n:3 * n + 2] = torch.from_numpy(part["pe"])
    block[3 * n + 2:3 * n + 11] = torch.from_numpy(part["stress"].reshape(-1))
    spmd.combine(block)
    b = block.numpy()
    out = evalf.tail(ms, b[:3 * n].reshape(3, n), b[3 * n:3 * n + 2], b[3 * n + 2:3 * n + 11].reshape(3, 3))
    flat = torch.from_numpy(np.concatenate([out[k].reshape(-1) for k in ("force", "torque", "pe", "stress", "dip_mom")]))
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        q.put(({k: np.array(v) for k, v in out.items()}, same))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_eval_forces_tail_after_the_allreduce():
    """SURVEY 8f rank 1 under the replicated-data scheme: the molecular-frame steps of eval_forces (surface dipole,
    mol_force/mol_torque, site->molecular virial, distant terms) run after the packed all-reduce, identically on every
    rank, and reproduce the reference's single-process eval_forces()."""
    from oracle import ref as refmod
    from tests import cases
    if not refmod.available(evalf=True):
        pytest.skip("oracle/_ref/libmoldyref_evalf.so not built")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker_evalf, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, same = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same, "ranks disagree on the molecular forces after the all-reduce"
    ms = cases.GOLDEN_CASES["mgcl2"]()
    ms.control.surface_dipole = 1
    want = refmod.RefLib(evalf=True).eval_forces(ms)
    assert cases.rel_rms(out["force"], want["force"]) < 1e-12
    assert cases.rel_rms(out["torque"], want["torque"]) < 1e-12
    assert np.allclose(out["pe"], want["pe"], rtol=1e-12)
    assert np.linalg.norm(out["stress"] - want["stress"]) / np.linalg.norm(want["stress"]) < 1e-11
    assert np.allclose(out["dip_mom"], want["dip_mom"], rtol=1e-10, atol=1e-8)


def test_partition_helpers_cover_everything_once():
    from moldy_b200 import spmd
    for n, w in ((1024000, 8), (1000, 3), (7, 8)):
        sl = [spmd.site_slice(n, r, w) for r in range(w)]
        assert sl[0][0] == 0 and sl[-1][1] == n
        assert all(sl[i][1] == sl[i + 1][0] for i in range(w - 1))
    owners = [spmd.column_owner(v, 4) for v in range(103)]
    assert sorted(set(owners)) == [0, 1, 2, 3]
    assert max(np.bincount(owners)) - min(np.bincount(owners)) <= 1


# ---- the peer layer's scheme (moldy_b200/csrc/mdb_peer.cu) on gloo: reduce-scatter by result slices + all-gather ----
def _worker_slices(rank, world, port_no, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from moldy_b200 import spmd
    from oracle import port
    from tests import cases
    ms = cases.GOLDEN_CASES["mgcl2"]()
    n = ms.nsites
    part = port.run(ms, ithread=rank, nthreads=world)
    f = torch.from_numpy(part["force"].copy())                        # [3, n] partial
    scal = torch.zeros(16, dtype=torch.float64)
    scal[0:2] = torch.from_numpy(part["pe"]); scal[2:11] = torch.from_numpy(part["stress"].reshape(-1))
    b = spmd.result_bounds(n, world)
    # phase C: rank r sums slice r of every row in rank order (fixed order: identical bits everywhere); scalars on all ranks
    red = torch.zeros((3, n), dtype=torch.float64)
    for r in range(world):
        piece = f[:, b[r]:b[r + 1]].contiguous()
        parts = [torch.zeros_like(piece) for _ in range(world)] if rank == r else None
        dist.gather(piece, parts, dst=r)
        if rank == r:
            acc = torch.zeros_like(piece)
            for p in parts:
                acc += p
            red[:, b[r]:b[r + 1]] = acc
    allscal = [torch.zeros_like(scal) for _ in range(world)]
    dist.all_gather(allscal, scal)
    tot = torch.zeros_like(scal)
    for s_ in allscal:
        tot += s_
    # phase D: all-gather of the slices
    for r in range(world):
        piece = red[:, b[r]:b[r + 1]].contiguous()
        dist.broadcast(piece, src=r)
        red[:, b[r]:b[r + 1]] = piece
    gathered = [torch.zeros_like(red) for _ in range(world)]
    dist.all_gather(gathered, red)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        q.put((red.numpy().copy(), tot.numpy().copy(), same))
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_scatter_by_slices_plus_allgather_reproduces_single_rank():
    from oracle import port
    from tests import cases
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker_slices, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    force, scal, same = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert same
    ms = cases.GOLDEN_CASES["mgcl2"]()
    one = port.run(ms)
    assert cases.rel_rms(force, one["force"]) < 1e-13
    assert np.allclose(scal[0:2], one["pe"], rtol=1e-12)
    assert np.allclose(scal[2:11].reshape(3, 3), one["stress"], rtol=1e-11, atol=1e-11 * np.abs(one["stress"]).max())


def test_molecule_bounds_cut_at_molecule_boundaries():
    from moldy_b200 import spmd
    mb, sb = spmd.molecule_bounds([200, 4, 8], [3, 1, 1], 4)            # MgCl2 cell: 200 waters, 4 Mg, 8 Cl = 612 sites
    assert mb[0] == 0 and mb[-1] == 212 and sb[0] == 0 and sb[-1] == 612
    assert all(sb[r] <= sb[r + 1] for r in range(4))
    assert all(s % 3 == 0 for s in sb if s <= 600)                      # inside the waters a cut never splits a molecule
    assert max(sb[r + 1] - sb[r] for r in range(4)) - 612 / 4 <= 3      # balanced to within one molecule
    mb1, sb1 = spmd.molecule_bounds([5], [4], 8)                        # more ranks than molecules: empty shares, no overlap
    assert mb1[-1] == 5 and sorted(mb1) == mb1
