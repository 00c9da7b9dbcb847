"""mdb_force_both(): the real-space pass beside the k-space kernels (a persistent "filler" grid of the pair kernel that
draws batches from a counter until the k-space stream raises its stop flag, then the rest of the pass at full occupancy).
The sums must be those of mdb_force_real + mdb_force_recip -- and the reference's -- whatever share the filler takes."""
import os

import numpy as np
import pytest
import torch

from moldy_b200 import lib
from tests import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(ms, both, fill=(0, 0), mode=4):
    eng = lib.Engine(0)
    eng.set_pair_mode(mode)
    eng.configure(ms)
    eng.set_sites_host(ms.make_sites())
    st = torch.cuda.current_stream().cuda_stream
    out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
    eng.build_cells(st)
    if both:
        eng.set_overlap(*fill)
        eng.force_both(out.data_ptr(), st)
    else:
        eng.force_real(out.data_ptr(), st)
        eng.force_recip(out.data_ptr(), st)
    torch.cuda.synchronize()
    res = lib.unpack(out.cpu().numpy(), ms.nsites)
    eng.close()
    return res


@pytest.mark.parametrize("name", ["tip4p", "tip4p_2", "mgcl2", "quartz", "slab_framework", "tips2_strict", "morse", "argon"])
@pytest.mark.parametrize("fill", [(0, 0), (148, 64), (7, 128), (1000, 32)])
def test_force_both_equals_the_two_phases_and_the_reference(name, fill):
    ms = cases.GOLDEN_CASES[name]()
    f0, pe0, s0 = _run(ms, False)
    f1, pe1, s1 = _run(ms, True, fill)
    assert cases.rel_rms(f1, f0) < 1e-13
    assert (np.abs(pe1 - pe0) <= 1e-12 * np.abs(pe0)).all(), (pe0, pe1)
    assert np.abs(s1 - s0).max() <= 1e-12 * np.abs(s0).max()
    gold = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    assert cases.rel_rms(f1, gold["force"]) < 1e-10       # (the record's energies hold ewald()'s host-side self/sheet terms too)


def test_force_both_without_the_newton3_kernel_falls_back_to_the_phases():
    """pair mode 3 (owner-computes, bit-reproducible) has no filler instantiation: same bits as the two calls."""
    ms = cases.GOLDEN_CASES["tip4p_2"]()
    f0, pe0, s0 = _run(ms, False, mode=3)
    f1, pe1, s1 = _run(ms, True, mode=3)
    assert np.array_equal(f0, f1) and np.array_equal(pe0, pe1) and np.array_equal(s0, s1)


def test_force_both_at_benchmark_size_takes_a_share_in_the_filler():
    """1.024 M sites: the filler grid works through part of the Coulomb pass while k-space runs; the reduced record of
    the compiled reference (tests/golden/large_tip4p_10.npz) pins the sums."""
    path = os.path.join(GOLD, "large_tip4p_10.npz")
    if not os.path.exists(path):
        pytest.skip("large fixture not generated")
    g = np.load(path)
    ms = cases.LARGE_CASES["tip4p_10"]()
    f, pe, s = _run(ms, True, (296, 128))
    smp = g["sample"]
    assert cases.rel_rms(f[:, smp], g["fsample"]) < 1e-10
    assert np.abs((f ** 2).sum(1) / g["fsq"] - 1).max() < 1e-10


def _run_far(ms, far):
    eng = lib.Engine(0)
    eng.set_pair_far(far)
    eng.configure(ms)
    nfar = eng.pair_far_runs()
    eng.set_sites_host(ms.make_sites())
    st = torch.cuda.current_stream().cuda_stream
    out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
    eng.build_cells(st)
    eng.force_real(out.data_ptr(), st)
    torch.cuda.synchronize()
    res = lib.unpack(out.cpu().numpy(), ms.nsites) + (eng.pair_count(st), nfar)
    eng.close()
    return res


def test_far_runs_of_the_buckingham_potential_without_the_exponential():
    """quartz at the benchmark size (995 328 ions, cut-off of ~80 decay lengths of the O-O repulsion): the stencil runs beyond
    52 decay lengths are walked with -C/r^6 + Coulomb only (PT_HIW rows).  Same forces, energy, stress and pair count as with
    the full potential on every run; the comparison with the compiled reference's record is test_gpu_large's."""
    ms = cases.LARGE_CASES["quartz_48"]()
    f0, pe0, s0, n0, far0 = _run_far(ms, False)
    f1, pe1, s1, n1, far1 = _run_far(ms, True)
    assert far0 == 0 and far1 > 0
    assert n0 == n1
    assert cases.rel_rms(f1, f0) < 1e-13
    assert abs(pe1[0] - pe0[0]) <= 1e-13 * abs(pe0[0])
    assert np.abs(s1 - s0).max() <= 1e-13 * np.abs(s0).max()


@pytest.mark.parametrize("name", ["quartz", "morse", "mgcl2", "morse_nocoul"])
def test_far_run_switch_is_neutral_where_no_run_is_far(name):
    ms = cases.GOLDEN_CASES[name]()
    f0, pe0, s0, n0, far0 = _run_far(ms, False)
    f1, pe1, s1, n1, far1 = _run_far(ms, True)
    assert cases.rel_rms(f1, f0) < 1e-13 and n0 == n1 and far0 == 0


@pytest.mark.parametrize("name", ["tip4p_2", "mgcl2", "slab_framework"])
def test_lookahead_kspace_sum_in_slices_through_the_c_abi(name):
    """force_calc() starts ewald()'s k-space kernels ahead; for large systems the force kernel is cut into slices of the
    charged sites and ewald() adds each slice to the caller's rows while the next is computed.  Forced here on small
    golden systems (4 and 3 slices, no size threshold): same result as the reference's record."""
    L = lib.load()
    gold = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    for ns in (4, 3, 1):
        L.mdb_abi_set_kf_slices(ns, 0)
        try:
            ms = cases.GOLDEN_CASES[name]()
            lib.reset()
            out = lib.eval_forces(ms)
        finally:
            L.mdb_abi_set_kf_slices(4, 200000)
        assert cases.rel_rms(out["force"], gold["force"]) < 1e-10, ns
        assert (np.abs(out["pe"] - gold["pe"]) / np.abs(gold["pe"])).max() < 1e-11, ns
        iu = np.triu_indices(3)
        assert np.linalg.norm(out["stress"][iu] - gold["stress"][iu]) / np.linalg.norm(gold["stress"][iu]) < 1e-11, ns
