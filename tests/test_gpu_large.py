"""GPU tests at the BASELINE.json sizes (configs[1..4]: 128 000 / 278 516 / 995 328 / 1 024 000 sites).

1. Against the compiled reference: tests/golden/large_*.npz hold reduced records of the reference's own
   force_calc()+ewald() at full size (energies, stress, the forces of 16 384 sampled sites, eight weighted sums
   over ALL site forces, the digest of every site's link-cell index), written by
   tests/golden/make_large_fixtures.py from an 8-rank replicated-data run of oracle/_ref summed as par_rsum does.
   The library is driven through the C ABI (force_calc + ewald with host buffers) and must meet the north_star
   tolerances there: cells bit-exact, forces 1e-10 relative RMS, energies and stress 1e-11.
2. Size-independent properties: Newton's third law, agreement of the pair-kernel variants (two different
   traversals), the DMMA and DFMA k-space kernels (two independent implementations), invariance under a
   relabelling of the molecules, bit-reproducibility."""
import hashlib
import os

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


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _weights(n):
    i = np.arange(n, dtype=np.float64)
    return np.stack([np.sin(0.37 * (k + 1) * i + 0.11 * k) for k in range(8)])


def _check_against_record(g, force, pe, stress, what):
    """north_star tolerances against a reduced reference record (tests/golden/make_large_fixtures.py)."""
    smp = g["sample"]
    fr = cases.rel_rms(force[:, smp], g["fsample"])
    assert fr < 1e-10, (what, "sampled forces", fr)
    # weighted sums over ALL sites: |error| against the size of a sum of N terms of the typical magnitude
    proj = _weights(force.shape[1]) @ force.T
    scale = np.sqrt(g["fsq"].sum())
    assert np.abs(proj - g["proj"]).max() < 1e-10 * scale, (what, "projections", np.abs(proj - g["proj"]).max() / scale)
    assert np.abs((force ** 2).sum(1) / g["fsq"] - 1).max() < 1e-10, (what, "sum f^2")
    er = np.abs(pe - g["pe"]) / np.abs(g["pe"])
    assert er.max() < 1e-11, (what, "pe", er)
    iu = np.triu_indices(3)
    sr = np.linalg.norm(stress[iu] - g["stress"][iu]) / np.linalg.norm(g["stress"][iu])
    assert sr < 1e-11, (what, "stress", sr)
    return fr, er.max(), sr


@pytest.mark.parametrize("name", list(cases.LARGE_CASES))
def test_full_size_vs_compiled_reference(name):
    """BASELINE.json configs[1..4] through force_calc()+ewald() of the C ABI against the reference's record."""
    path = os.path.join(GOLD, f"large_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    g = np.load(path)
    ms = cases.LARGE_CASES[name]()
    assert ms.nsites == int(g["nsites"])
    lib.reset()
    out = lib.eval_forces(ms)
    fr, er, sr = _check_against_record(g, out["force"], out["pe"], out["stress"], name)
    print(f"{name}: N={ms.nsites} force relRMS={fr:.2e} pe rel={er:.2e} stress rel={sr:.2e}")
    eng = lib.Engine(0)
    eng.configure(ms)
    eng.set_sites_host(ms.make_sites())
    cid = eng.cell_ids()
    assert eng.n_kvectors() == int(g["n_kvectors"])
    eng.close()
    assert np.array_equal(cid[g["sample"]], g["cell_sample"])
    assert hashlib.sha256(np.ascontiguousarray(cid, dtype=np.int32).tobytes()).hexdigest() == str(g["cell_sha256"])
    lib.reset()


@pytest.mark.parametrize("name", ["tip4p_10", "quartz_48"])
def test_full_size_eval_forces_vs_reference_record(name):
    """One level up: the library's eval_forces() at full size with surface-dipole off, so that a
    molecular force is the plain sum of the molecule's site forces.  Quartz: every molecule is one site, the sampled
    molecular forces must be the record's site forces.  TIP4P: reciprocal-space energy and zero net force."""
    path = os.path.join(GOLD, f"large_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    g = np.load(path)
    ms = cases.LARGE_CASES[name]()
    ms.control.surface_dipole = 0
    lib.reset()
    mol = lib.eval_forces_mol(ms)
    lib.reset()
    assert abs(mol["pe"][1] - g["pe"][1]) < 1e-11 * abs(g["pe"][1])
    if ms.nmols == ms.nsites:
        assert cases.rel_rms(mol["force"][g["sample"]].T, g["fsample"]) < 1e-10
    fmax = np.abs(mol["force"]).max()
    assert np.isfinite(mol["force"]).all() and np.abs(mol["force"].sum(0)).max() < 1e-9 * fmax * np.sqrt(ms.nmols)


@pytest.mark.parametrize("n", [5, 10])
def test_kspace_dmma_vs_dfma_at_full_size(n, monkeypatch):
    """Two independent k-space implementations (DMMA GEMMs and the DFMA kernels, MDB_KSPACE=dfma) at hmax/lmax up to 25."""
    ms = systems.tip4p(n)
    res = {}
    for mode in ("mma", "dfma"):
        monkeypatch.setenv("MDB_KSPACE", mode)
        eng = lib.Engine(0)
        eng.configure(ms)
        eng.set_sites_host(ms.make_sites())
        st = torch.cuda.current_stream().cuda_stream
        out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
        eng.force_recip(out.data_ptr(), st)
        torch.cuda.synchronize()
        res[mode] = lib.unpack(out.cpu().numpy(), ms.nsites)
        eng.close()
    assert cases.rel_rms(res["mma"][0], res["dfma"][0]) < 1e-11
    assert abs(res["mma"][1][1] / res["dfma"][1][1] - 1) < 1e-11
    iu = np.triu_indices(3)
    assert np.linalg.norm(res["mma"][2][iu] - res["dfma"][2][iu]) < 1e-11 * np.linalg.norm(res["dfma"][2][iu])


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
