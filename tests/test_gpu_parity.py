"""GPU parity tests proper: the CUDA path, called through the C ABI
(force_calc / ewald with host buffers, and the mdb_* engine), against

  * the committed golden vectors (outputs of the reference's own
    force_calc()/ewald()/cellbin(), tests/golden/ref_*.npz), and
  * the reference itself (oracle/_ref/libmoldyref.so, prebuilt) on seeded inputs.

Tolerances are the north_star's: cell assignment bit-exact; per-site forces
<= 1e-10 relative RMS; energies and stress <= 1e-11 relative (stress relative
to the Frobenius norm of the tensor: individual off-diagonal components of an
isotropic liquid are sums that cancel to ~0)."""
import os

import numpy as np
import pytest

from moldy_b200 import lib, systems
from tests import cases

pytestmark = pytest.mark.gpu

F_TOL = 1e-10
E_TOL = 1e-11


def _check(out, gold, name):
    fr = cases.rel_rms(out["force"], gold["force"])
    assert fr <= F_TOL, f"{name}: force rel RMS {fr:.3e}"
    for k in range(2):
        ref = gold["pe"][k]
        if ref != 0.0:
            er = abs(out["pe"][k] - ref) / abs(ref)
            assert er <= E_TOL, f"{name}: pe[{k}] rel err {er:.3e} ({out['pe'][k]} vs {ref})"
        else:
            assert out["pe"][k] == 0.0
    iu = np.triu_indices(3)
    sr = np.linalg.norm(out["stress"][iu] - gold["stress"][iu]) / np.linalg.norm(gold["stress"][iu])
    assert sr <= E_TOL, f"{name}: stress rel err {sr:.3e}"
    # the lower triangle is the caller's (src/accel.c:578-579): must stay untouched
    assert out["stress"][1, 0] == 0 and out["stress"][2, 0] == 0 and out["stress"][2, 1] == 0


@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_force_calc_ewald_vs_golden(name, golden_dir):
    ms = cases.GOLDEN_CASES[name]()
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    lib.reset()
    out = lib.eval_forces(ms)
    _check(out, gold, name)


@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_cell_assignment_bit_exact(name, golden_dir):
    ms = cases.GOLDEN_CASES[name]()
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    eng = lib.Engine(0)
    eng.configure(ms)
    eng.set_sites_host(ms.make_sites(wrap=not ms.control.molpbc))
    if ms.control.molpbc:
        eng.set_com_host(ms.c_of_m)
    got = eng.cell_ids()
    assert np.array_equal(got, gold["cell"]), f"{name}: {np.count_nonzero(got != gold['cell'])} sites in the wrong cell"
    eng.close()


@pytest.mark.parametrize("name", ["argon", "tip4p", "tips2", "mgcl2", "quartz", "clay"])
def test_startup_scalars_match_example_outputs(name):
    """#subcells, #neighbour cells and #k-vectors printed by the reference's 1996
    sample outputs (the only goldens its tree holds for this path)."""
    ms = cases.EXAMPLE_SYSTEMS[name]()
    sub, nab, self_e, nk = cases.EXAMPLE_GOLDENS[name]
    eng = lib.Engine(0)
    eng.configure(ms)
    nc, _ = eng.grid()
    if sub is not None:
        assert nc == sub
        assert 2 * eng.n_neighbour_cells() == nab
    if nk is not None:
        assert eng.n_kvectors() == nk
    eng.close()
    if self_e is not None:
        lib.reset()
        lib.eval_forces(ms)
        c = np.zeros(3)
        lib.load().mdb_abi_constants(c.ctypes.data_as(lib.DP))
        assert abs(c[1] * systems.CONV_E - self_e) < 5e-6
        if name == "clay":            # framework-charge sheet correction, src/examples/clay-example.out:77
            assert abs(c[2] * systems.CONV_E / np.linalg.det(ms.h) - cases.CLAY_SHEET_CORRECTION) < 5e-4


def test_partition_sums_to_whole():
    """Replicated-data slices (ithread/nthreads, src/force.c:856, src/ewald.c:495)
    add up to the single-rank result."""
    import torch
    ms = cases.GOLDEN_CASES["tip4p"]()
    eng = lib.Engine(0)
    eng.configure(ms)
    eng.set_sites_host(ms.make_sites())
    n = ms.nsites
    st = torch.cuda.current_stream().cuda_stream

    def run(i, p):
        out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
        eng.set_partition(i, p)
        eng.build_cells(st)
        eng.force_real(out.data_ptr(), st)
        eng.force_recip(out.data_ptr(), st)
        torch.cuda.synchronize()
        return out.cpu().numpy()

    whole = run(0, 1)
    parts = sum(run(i, 3) for i in range(3))
    f0, pe0, s0 = lib.unpack(whole, n)
    f1, pe1, s1 = lib.unpack(parts, n)
    assert cases.rel_rms(f1, f0) < 1e-13
    assert np.allclose(pe1, pe0, rtol=1e-12)
    assert np.allclose(s1, s0, rtol=1e-11, atol=1e-11 * np.abs(s0).max())
    eng.close()


@pytest.mark.parametrize("mode", [2, 3, 4])
@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_every_pair_kernel_variant_vs_golden(name, mode, golden_dir):
    """All three real-space kernel variants (per-thread, tiled, tiled Newton-3) through the
    device-level engine API, real space only, against the reference's force_calc()."""
    import torch
    from oracle import port
    ms = cases.GOLDEN_CASES[name]()
    n = ms.nsites
    eng = lib.Engine(0)
    eng.set_pair_mode(mode)
    eng.configure(ms)
    eng.set_sites_host(ms.make_sites(wrap=not ms.control.molpbc))
    if ms.control.molpbc:
        eng.set_com_host(ms.c_of_m)
    st = torch.cuda.current_stream().cuda_stream
    out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
    eng.build_cells(st)
    eng.force_real(out.data_ptr(), st)
    torch.cuda.synchronize()
    f, pe, s = lib.unpack(out.cpu().numpy(), n)
    gold = port.run(ms, recip=False)            # bit-identical to the compiled reference (test_oracle.py)
    assert cases.rel_rms(f, gold["force"]) <= F_TOL
    ref_pe = gold["pe"][0] + gold["eintra"]     # the engine block carries no first-call constants
    assert abs(pe[0] - ref_pe) <= E_TOL * abs(ref_pe)
    iu = np.triu_indices(3)
    assert np.linalg.norm(s[iu] - gold["stress"][iu]) <= E_TOL * np.linalg.norm(gold["stress"][iu])
    assert abs(eng.pair_count(st) - gold["npairs"]) < 0.5
    eng.close()


def test_site_partitioned_kspace_sums_to_whole():
    """The multi-GPU k-space scheme emulated on one GPU: every rank's structure-factor sums are
    added (the all-reduce), then every rank back-projects onto its own sites; the rank blocks
    add up to the single-rank result (energy and stress come from rank 0 only)."""
    import torch
    ms = cases.GOLDEN_CASES["slab_framework"]()
    n = ms.nsites
    st = torch.cuda.current_stream().cuda_stream
    eng = lib.Engine(0)
    eng.configure(ms)
    eng.set_sites_host(ms.make_sites())
    whole = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
    eng.force_recip(whole.data_ptr(), st)
    P = 3
    psums = []
    for r in range(P):
        eng.set_partition(r, P)
        ps = torch.zeros(eng.recip_sum_doubles(), dtype=torch.float64, device="cuda")
        eng.recip_partial(ps.data_ptr(), st)
        psums.append(ps)
    total = sum(psums)
    parts = torch.zeros_like(whole)
    for r in range(P):
        eng.set_partition(r, P)
        eng.recip_finish(total.data_ptr(), parts.data_ptr(), st)
    torch.cuda.synchronize()
    a, b = whole.cpu().numpy(), parts.cpu().numpy()
    assert cases.rel_rms(b[:3 * n], a[:3 * n]) < 1e-13
    assert np.allclose(b[3 * n:3 * n + 11], a[3 * n:3 * n + 11], rtol=1e-12, atol=1e-9)
    eng.close()


def test_deterministic_repeat():
    """pair_mode 3 (owner-computes, no atomics) is bit-reproducible run to run; the default
    Newton-3 mode accumulates with red.global.add.f64 and may differ in the last bits."""
    import torch
    ms = cases.GOLDEN_CASES["mgcl2"]()
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for mode in (3, 4):
        eng = lib.Engine(0)
        eng.set_pair_mode(mode)
        eng.configure(ms)
        eng.set_sites_host(ms.make_sites())
        outs = []
        for _ in range(2):
            out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
            eng.build_cells(st)
            eng.force_real(out.data_ptr(), st)
            eng.force_recip(out.data_ptr(), st)
            torch.cuda.synchronize()
            outs.append(out.cpu().numpy())
        res[mode] = outs
        eng.close()
    assert np.array_equal(res[3][0], res[3][1])
    n = ms.nsites
    assert cases.rel_rms(res[4][0][:3 * n], res[4][1][:3 * n]) < 1e-13
    assert cases.rel_rms(res[4][0][:3 * n], res[3][0][:3 * n]) < 1e-13


def test_kernel_poteval_dist_pot_vs_reference():
    """The exported scalar/vector potential entry points against the reference's."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    r = ref.RefLib()
    rng = np.random.default_rng(0)
    npar = [2, 3, 4, 6, 3, 1, 7]
    for alpha in (0.35, -1.0):
        r.control.alpha = alpha
        lib.control().alpha = alpha
        for ptype in (0, 1, 2, 3, 4, 6):
            p = np.zeros(8)
            p[:npar[ptype]] = rng.uniform(0.5, 3.0, npar[ptype])
            for rr in (0.9, 2.7, 6.1):
                a = lib.poteval(p, rr, ptype, -0.3)
                b = r.lib.poteval(p.ctypes.data_as(lib.DP), rr, ptype, -0.3)
                assert abs(a - b) <= 1e-12 * max(1.0, abs(b)), (ptype, alpha, rr, a, b)
            a = lib.dist_pot(p, 8.0, ptype)
            b = r.lib.dist_pot(p.ctypes.data_as(lib.DP), 8.0, ptype)
            assert abs(a - b) <= 1e-13 * max(1.0, abs(b))


def test_medium_system_vs_reference_live():
    """27 648-site TIP4P (3x3x3 replica, jittered): the reference takes ~2 s."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    ms = systems.tip4p(3, seed=7)
    gold = ref.RefLib().run(ms)
    lib.reset()
    out = lib.eval_forces(ms)
    _check(out, gold, "tip4p_3")


@pytest.mark.parametrize("mode", [2, 4])
@pytest.mark.parametrize("name", list(cases.RDF_CASES))
def test_rdf_pass_pair_counts_exact(name, mode, golden_dir):
    """RDF pass of force_calc (src/force.c:1302-1313 + src/rdf.c:94-108) on the device: every pair lands
    in the reference's bin (integer work: exact), whole and as a 3-way replicated-data split."""
    import torch
    limit, nbins = cases.RDF_CASES[name]
    ms = cases.GOLDEN_CASES[name]()
    gold = np.load(os.path.join(golden_dir, "ref_rdf.npz"))[name]
    eng = lib.Engine(0)
    eng.set_pair_mode(mode)
    eng.configure(ms)
    eng.set_sites_host(ms.make_sites(wrap=not ms.control.molpbc))
    if ms.control.molpbc:
        eng.set_com_host(ms.c_of_m)
    st = torch.cuda.current_stream().cuda_stream
    cnt = eng.rdf_counts(limit, nbins, st)
    assert cnt.shape == gold.shape
    assert np.array_equal(cnt.astype(np.int64), gold), int(np.abs(cnt.astype(np.int64) - gold).sum())
    parts = np.zeros_like(cnt)
    for r in range(3):
        eng.set_partition(r, 3)
        parts += eng.rdf_counts(limit, nbins, st)
    assert np.array_equal(parts, cnt)
    eng.close()


@pytest.mark.parametrize("name", ["tip4p", "slab_framework", "tips2_molpbc"])
def test_force_calc_accumulates_rdf_store(name, golden_dir):
    """force_calc() with control.rdf_interval on: the float store behind rdf_ptr() holds count/density
    (the reference adds 1/density pair by pair in single precision: compare to float accuracy), and the
    forces are unaffected by the extra pass."""
    limit, nbins = cases.RDF_CASES[name]
    ms = cases.GOLDEN_CASES[name]()
    gold = np.load(os.path.join(golden_dir, "ref_rdf.npz"))[name]
    ref = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    lib.reset()
    out = lib.eval_forces(ms, rdf=(limit, nbins))
    _check(out, ref, name)
    rho = ms.nsites / float(np.linalg.det(ms.h))
    assert out["rdf"].shape == gold.shape
    assert np.allclose(out["rdf"], gold / rho, rtol=1e-6, atol=0)
    lib.reset()


def _real_space(ms, mode, env):
    """force_real through the engine API under the given environment switches"""
    import torch
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        eng = lib.Engine(0)
        eng.set_pair_mode(mode)
        eng.configure(ms)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return eng


@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("mode", [3, 4])
@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_split_pair_passes_vs_golden(name, mode, split, golden_dir):
    """The real-space sum as one fused pass and as two passes by site class (charged x charged with the
    Coulomb term only + potential x potential with the potential only), both forced, against force_calc()."""
    import torch
    from oracle import port
    ms = cases.GOLDEN_CASES[name]()
    n = ms.nsites
    eng = _real_space(ms, mode, {"MDB_PAIR_SPLIT": str(split)})
    eng.set_sites_host(ms.make_sites(wrap=not ms.control.molpbc))
    if ms.control.molpbc:
        eng.set_com_host(ms.c_of_m)
    st = torch.cuda.current_stream().cuda_stream
    out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
    eng.build_cells(st)
    eng.force_real(out.data_ptr(), st)
    torch.cuda.synchronize()
    f, pe, s = lib.unpack(out.cpu().numpy(), n)
    gold = port.run(ms, recip=False)
    assert cases.rel_rms(f, gold["force"]) <= F_TOL
    ref_pe = gold["pe"][0] + gold["eintra"]
    assert abs(pe[0] - ref_pe) <= E_TOL * abs(ref_pe)
    iu = np.triu_indices(3)
    assert np.linalg.norm(s[iu] - gold["stress"][iu]) <= E_TOL * np.linalg.norm(gold["stress"][iu])
    assert abs(eng.pair_count(st) - gold["npairs"]) < 0.5     # the reference's pair count, whatever the passes visit
    eng.close()


@pytest.mark.parametrize("name", [k for k in cases.GOLDEN_CASES if cases.GOLDEN_CASES[k]().control.alpha > 0])
def test_kspace_dmma_and_dfma_kernels_agree(name, golden_dir):
    """Reciprocal space on the FP64 tensor pipe (k_ktables/k_sfac_mma/k_kforce_mma, the default) and with the
    register-operand DFMA kernels (MDB_KSPACE=dfma): same sums, different order."""
    import torch
    ms = cases.GOLDEN_CASES[name]()
    n = ms.nsites
    res = {}
    for kind in ("mma", "dfma"):
        os.environ["MDB_KSPACE"] = kind
        try:
            eng = lib.Engine(0)
            eng.configure(ms)
            eng.set_sites_host(ms.make_sites(wrap=not ms.control.molpbc))
            st = torch.cuda.current_stream().cuda_stream
            out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
            eng.force_recip(out.data_ptr(), st)
            torch.cuda.synchronize()
            res[kind] = lib.unpack(out.cpu().numpy(), n)
            eng.close()
        finally:
            os.environ.pop("MDB_KSPACE", None)
    (f1, pe1, s1), (f2, pe2, s2) = res["mma"], res["dfma"]
    assert cases.rel_rms(f1, f2) <= 1e-12
    assert abs(pe1[1] - pe2[1]) <= 1e-12 * abs(pe2[1])
    assert np.linalg.norm(s1 - s2) <= 1e-12 * np.linalg.norm(s2)


def test_too_close_pairs_counted_alike_by_fused_and_split_passes(golden_dir):
    """src/force.c:939-949 warns about every inter-molecular pair closer than 0.5 A, whatever its types.  A
    hydrogen pushed onto a foreign oxygen (a pair neither split pass visits: O carries no charge, H no
    Lennard-Jones term) must be reported the same by the fused pass and by the split passes + cell scan."""
    import torch
    ms = cases.GOLDEN_CASES["tip4p"]()
    site = ms.make_sites(wrap=True)
    nsm = ms.sysdef.species[0].nsites
    # sites of a TIP4P molecule: O, H, H, M; put an H of molecule 1 next to the O of molecule 0
    site[:, nsm + 1] = site[:, 0] + np.array([0.05, 0.03, 0.02])
    counts = []
    for split in (0, 1):
        eng = _real_space(ms, 4, {"MDB_PAIR_SPLIT": str(split)})
        eng.set_sites_host(site)
        st = torch.cuda.current_stream().cuda_stream
        out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
        eng.build_cells(st)
        eng.force_real(out.data_ptr(), st)
        torch.cuda.synchronize()
        ntc, pair = eng.too_close(st)
        counts.append(ntc)
        assert ntc > 0 and pair[0] // nsm != pair[1] // nsm
        eng.close()
    assert counts[0] == counts[1]


@pytest.mark.parametrize("split", [0, 1])
def test_too_close_count_equals_the_references_warnings(split, golden_dir):
    """The reference prints one TOOCLS warning per inter-molecular pair of its half list closer than 0.5 A
    (src/force.c:939-949).  Three displaced sites -- H on a foreign O (visited by neither split pass), H on a foreign
    H (Coulomb pass), O on a foreign O (Lennard-Jones pass) -- and one intra-molecular contact that must NOT count:
    the library's count through mdb_too_close equals the number of warnings the compiled reference issues."""
    import torch
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    ms = cases.GOLDEN_CASES["tip4p"]()
    site = ms.make_sites(wrap=True)
    nsm = ms.sysdef.species[0].nsites
    d = np.array([0.05, 0.03, 0.02])
    site[:, 1 * nsm + 1] = site[:, 0 * nsm + 0] + d          # H(1) on O(0)
    site[:, 3 * nsm + 2] = site[:, 2 * nsm + 1] + d          # H(3) on H(2)
    site[:, 5 * nsm + 0] = site[:, 4 * nsm + 0] - d          # O(5) on O(4)
    site[:, 6 * nsm + 3] = site[:, 6 * nsm + 0] + d          # M(6) on its own O: same molecule, no warning
    r = ref.RefLib()
    w0 = r.lib.mdref_warnings()
    r.run(ms, recip=False, sites=site)
    nref = r.lib.mdref_warnings() - w0
    assert nref >= 3
    eng = _real_space(ms, 4, {"MDB_PAIR_SPLIT": str(split)})
    eng.set_sites_host(site)
    st = torch.cuda.current_stream().cuda_stream
    out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
    eng.build_cells(st)
    eng.force_real(out.data_ptr(), st)
    torch.cuda.synchronize()
    ntc, pair = eng.too_close(st)
    eng.close()
    assert ntc == nref, (ntc, nref)


def test_ewald_recomputes_when_the_sites_changed_after_force_calc(golden_dir):
    """force_calc() starts the k-space kernels ahead of ewald(); ewald(site, ...) must still compute from the sites it
    is passed (src/ewald.c:280).  The caller rewrites its site array IN PLACE between the two calls: the result must be
    the reciprocal-space sum of the new sites, not the look-ahead result for the old ones."""
    from moldy_b200 import abi
    ms = cases.GOLDEN_CASES["tip4p_2"]()
    n, nsa = ms.nsites, abi.nsarray(ms.nsites)
    ms.control.fill(lib.control())
    lib.set_thread(0, 1)
    lib.reset()
    sysm, spec, pot = ms.cstructs()
    chg = ms.charges()
    site = np.ascontiguousarray(ms.make_sites())
    rng = np.random.default_rng(3)
    moved = site.copy()
    moved[:, :n] += 0.05 * rng.standard_normal((3, n))

    def recip_only(s):
        f, pe, stress = np.zeros((3, nsa)), np.zeros(2), np.zeros((3, 3))
        lib.ewald(s, f, sysm, spec, chg, pe[1:2], stress)
        return f[:, :n].copy(), pe[1], stress.copy()

    f_new, pe_new, s_new = recip_only(moved)                 # no look-ahead: plain ewald() on the moved sites
    f, pe, stress = np.zeros((3, nsa)), np.zeros(2), np.zeros((3, 3))
    buf = site.copy()
    lib.force_calc(buf, f, sysm, spec, chg, pot, pe[0:1], stress)     # starts the k-space kernels for `site`
    buf[:] = moved                                                      # same array, new contents
    f2, stress2 = np.zeros((3, nsa)), np.zeros((3, 3))
    lib.ewald(buf, f2, sysm, spec, chg, pe[1:2], stress2)
    lib.reset()
    assert cases.rel_rms(f2[:, :n], f_new) < 1e-13
    assert abs(pe[1] - pe_new) <= 1e-13 * abs(pe_new)
    assert np.linalg.norm(stress2 - s_new) <= 1e-13 * np.linalg.norm(s_new)


@pytest.mark.parametrize("kind", ["mma", "dfma"])
def test_many_l_slots_are_split_into_l_ranges(kind, golden_dir):
    """k_cutoff large enough for lmax > 32: the structure-factor GEMM then runs a column block over two l-ranges
    of <= 32 slots, the back-projection shrinks its site block to fit the longer tables into shared memory.
    Reciprocal space only, against the oracle (bit-identical to the compiled reference, test_oracle.py)."""
    import torch
    from oracle import port
    ms = cases.GOLDEN_CASES["tip4p"]()
    ms.control.k_cutoff = 2 * np.pi * 34.5 / np.linalg.norm(ms.h[:, 2])        # lmax = 34
    n = ms.nsites
    os.environ["MDB_KSPACE"] = kind
    try:
        eng = lib.Engine(0)
        eng.configure(ms)
        eng.set_sites_host(ms.make_sites(wrap=True))
        st = torch.cuda.current_stream().cuda_stream
        out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
        eng.force_recip(out.data_ptr(), st)
        torch.cuda.synchronize()
        f, pe, s = lib.unpack(out.cpu().numpy(), n)
        nk = eng.n_kvectors()
        eng.close()
    finally:
        os.environ.pop("MDB_KSPACE", None)
    gold = port.run(ms, real=False)
    assert nk == gold["nhkl"] and nk > 70000
    assert cases.rel_rms(f, gold["force"]) <= F_TOL
    ref_pe = gold["pe"][1] + gold["self_energy"] - gold["sheet_energy"] / float(np.linalg.det(ms.h))
    assert abs(pe[1] - ref_pe) <= E_TOL * abs(ref_pe)
    iu = np.triu_indices(3)
    assert np.linalg.norm(s[iu] - gold["stress"][iu]) <= E_TOL * np.linalg.norm(gold["stress"][iu])
