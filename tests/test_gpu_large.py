"""GPU tests at BASELINE sizes through size-independent properties (the reference needs minutes
to hours per step there): Newton's third law, agreement of the three pair-kernel variants,
invariance under a relabelling of the molecules, and extensivity against the 128k-site system
that IS checked against the reference (tests/test_gpu_parity.py::test_medium...)."""
import numpy as np
import pytest
import torch

from moldy_b200 import lib, systems
from tests import cases

pytestmark = pytest.mark.gpu


def _eval(ms, mode=4, sites=None):
    eng = lib.Engine(0)
    eng.set_pair_mode(mode)
    eng.configure(ms)
    eng.set_sites_host(ms.make_sites() if sites is None else sites)
    st = torch.cuda.current_stream().cuda_stream
    out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
    eng.build_cells(st)
    eng.force_real(out.data_ptr(), st)
    eng.force_recip(out.data_ptr(), st)
    torch.cuda.synchronize()
    res = lib.unpack(out.cpu().numpy(), ms.nsites) + (eng.pair_count(st), eng.n_kvectors())
    eng.close()
    return res


@pytest.mark.parametrize("n", [5, 10])
def test_million_site_properties(n):
    ms = systems.tip4p(n)
    f4, pe4, s4, pairs, nk = _eval(ms, 4)
    fmax = np.abs(f4).max()
    assert np.abs(f4.sum(1)).max() < 1e-9 * fmax * np.sqrt(ms.nsites)          # sum of all forces = 0
    f3, pe3, s3, pairs3, _ = _eval(ms, 3)
    assert pairs == pairs3
    assert cases.rel_rms(f4, f3) < 1e-12
    assert np.allclose(pe4, pe3, rtol=1e-11)
    iu = np.triu_indices(3)
    assert np.linalg.norm(s4[iu] - s3[iu]) < 1e-11 * np.linalg.norm(s3[iu])
    if n == 10:
        assert ms.nsites == 1024000 and nk == 34895                           # BASELINE.md probe of the reference
        assert abs(pairs / 6.43e9 - 1) < 0.01


def test_molecule_relabelling_invariance():
    """Reversing the molecule order permutes the forces and leaves energies/stress unchanged."""
    ms = systems.tip4p(3, seed=5)
    f, pe, s, _, _ = _eval(ms)
    ms2 = systems.tip4p(3, seed=5)
    ms2.c_of_m = ms.c_of_m[::-1].copy()
    ms2.quat = ms.quat[::-1].copy()
    f2, pe2, s2, _, _ = _eval(ms2)
    nm = ms.nmols
    fperm = f.reshape(3, nm, 4)[:, ::-1, :].reshape(3, -1)
    assert cases.rel_rms(f2, fperm) < 1e-12
    assert np.allclose(pe2, pe, rtol=1e-11)
    assert np.allclose(s2, s, rtol=1e-10, atol=1e-10 * np.abs(s).max())


@pytest.mark.parametrize("which", ["mgcl2_7", "quartz_48"])
def test_other_benchmark_families_at_full_size(which):
    """BASELINE.json configs[2] (aqueous MgCl2 replicated 7x7x7 = 278 516 sites, MCY + Ewald, automatic
    cut-offs) and configs[3] (BKS quartz 48x48x48 = 995 328 ions, Buckingham + Ewald, triclinic cell):
    size-independent properties -- zero net force, and the bit-reproducible owner-computes kernel
    (full stencil) against the Newton-3 kernel (half stencil + atomics), two different traversals."""
    ms = systems.mgcl2(7, explicit=False) if which == "mgcl2_7" else systems.quartz(48, pinned_cutoff=False)
    assert ms.nsites == (278516 if which == "mgcl2_7" else 995328)
    f4, pe4, s4, pairs, nk = _eval(ms, 4)
    fmax = np.abs(f4).max()
    assert np.isfinite(f4).all() and np.abs(f4.sum(1)).max() < 1e-9 * fmax * np.sqrt(ms.nsites)
    f3, pe3, s3, pairs3, _ = _eval(ms, 3)
    assert pairs == pairs3 and pairs > 100 * ms.nsites and nk > 1000
    assert cases.rel_rms(f4, f3) < 1e-12
    assert np.allclose(pe4, pe3, rtol=1e-11)
    iu = np.triu_indices(3)
    assert np.linalg.norm(s4[iu] - s3[iu]) < 1e-11 * np.linalg.norm(s3[iu])


def test_kspace_gemm_kernels_are_bit_reproducible():
    """k_sfac_mma / k_kforce_mma stream their operands through mbarrier-guarded TMA rings with warps drifting
    apart; the sums themselves have a fixed order, so three evaluations at 128 000 sites must agree bit for bit
    (a stage overwritten too early or read too early would show up here)."""
    import torch
    ms = systems.tip4p(5)
    eng = lib.Engine(0)
    eng.configure(ms)
    eng.set_sites_host(ms.make_sites())
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for _ in range(3):
        out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
        eng.force_recip(out.data_ptr(), st)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
    eng.close()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert np.abs(outs[0][:3 * ms.nsites]).max() > 0
