"""CPU tests (-m "not gpu"): the oracle itself.

The C restatement (oracle/moldy_oracle.c) is pinned against
  * every start-up scalar the reference's own example outputs hold for this path,
  * the committed outputs of the reference's compiled force_calc()/ewald()
    (tests/golden/ref_*.npz, written by tests/golden/make_fixtures.py),
  * and, where /root/reference exists, oracle/_ref run live on seeded inputs.
"""
import os

import numpy as np
import pytest

from moldy_b200 import systems
from oracle import port, ref
from tests import cases


@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_restatement_matches_reference_outputs(name, golden_dir):
    ms = cases.GOLDEN_CASES[name]()
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    out = port.run(ms)
    # same operation and summation order => forces agree to the last bit
    assert np.array_equal(out["force"], gold["force"]), cases.rel_rms(out["force"], gold["force"])
    for k in range(2):
        if gold["pe"][k] != 0:
            assert abs(out["pe"][k] - gold["pe"][k]) <= 1e-12 * abs(gold["pe"][k])
    iu = np.triu_indices(3)
    assert np.linalg.norm(out["stress"][iu] - gold["stress"][iu]) <= 1e-13 * np.linalg.norm(gold["stress"][iu])
    assert np.array_equal(port.cell_ids(ms), gold["cell"])


@pytest.mark.parametrize("name", list(cases.EXAMPLE_GOLDENS))
def test_startup_scalars_of_example_outputs(name):
    """src/examples/*-example.out: subcells, neighbour cells, self energy, k-vectors."""
    ms = cases.EXAMPLE_SYSTEMS[name]()
    sub, nab, self_e, nk = cases.EXAMPLE_GOLDENS[name]
    out = port.run(ms)
    if sub is not None:
        assert out["grid"][0] * out["grid"][1] * out["grid"][2] == sub
        assert 2 * out["n_nabors"] == nab
    if nk is not None:
        assert out["nhkl"] == nk
        assert abs(out["self_energy"] * systems.CONV_E - self_e) < 5e-6


def test_clay_framework_sheet_correction():
    """src/examples/clay-example.out:77: 'Framework has net electric charge of -8 - correction of 201.117 kJmol(-1)'."""
    ms = cases.EXAMPLE_SYSTEMS["clay"]()
    out = port.run(ms)
    assert abs(out["sheet_energy"] * systems.CONV_E / np.linalg.det(ms.h) - cases.CLAY_SHEET_CORRECTION) < 5e-4
    assert ms.nsites - ms.nsites_xf == 240 and ms.sysdef.species[-1].framework


def test_more_example_scalars():
    # tip4p-example.out:65 "Intramolecular potential energy correction = -255096" (6 sig. digits)
    e = port.constants(cases.EXAMPLE_SYSTEMS["tip4p"]())[0] * systems.CONV_E
    assert abs(e + 255096) < 1.0
    # quartz as shipped now uses subcell=3 -> 294 subcells / 110 cells
    q = port.run(cases.EXAMPLE_SYSTEMS["quartz"]())
    assert q["grid"] == (7, 6, 7) and 2 * q["n_nabors"] == 110


def test_ewald_parameters_as_startup_derives_them():
    ms = cases.EXAMPLE_SYSTEMS["tip4p"]()                      # 8.938 / 0.3794 / 2.573
    assert abs(ms.control.cutoff - 8.938) < 1e-3 and abs(ms.control.alpha - 0.3794) < 1e-4
    assert abs(ms.control.k_cutoff - 2.573) < 1e-3
    assert abs(ms.h[0, 0] - 19.7055) < 1e-4
    big = systems.tip4p(10)                                     # 28.265 / 0.119977 / 0.813724
    assert big.nsites == 1024000
    assert abs(big.control.cutoff - 28.265) < 2e-3 and abs(big.control.alpha - 0.119977) < 2e-6
    assert abs(big.control.k_cutoff - 0.813724) < 2e-5


def test_cellbin_edges():
    L = port.load()
    eps = 8 * 2.0 ** -52
    cb = lambda s, n: L.orc_cellbin(s, n, float(n), eps, None)
    assert cb(-0.5, 8) == 0 and cb(-0.5 - eps / 2, 8) == 0 and cb(-0.5 + eps / 2, 8) == 0
    assert cb(0.5, 8) == 7 and cb(0.5 - eps, 8) == 7 and cb(0.5 + eps / 2, 8) == 7
    assert cb(0.4999999, 8) == 7 and cb(0.0, 8) == 4 and cb(-1e-9, 8) == 3
    import ctypes
    err = ctypes.c_int(0)
    L.orc_cellbin(0.75, 8, 8.0, eps, ctypes.byref(err))
    assert err.value >= 1                                       # "Co-ordinate out of range in BIN"
    if ref.available():
        r = ref.RefLib()
        for s in (-0.5, 0.5, 0.4999999, -0.5 + eps / 2, 0.5 - eps, 0.1234, -0.25):
            assert cb(s, 8) == r.lib.cellbin(s, 8, 8.0, eps)


def test_err_fn_is_the_polynomial_not_erf():
    L = port.load()
    assert abs(L.orc_err_fn(0.5) - 0.5205000163) < 1e-9         # true erf(0.5) = 0.5204998778
    assert L.orc_err_fn(-0.5) == -L.orc_err_fn(0.5)


def test_partition_of_cells_and_kvectors_sums_to_whole():
    """ithread/nthreads slices (force.c:856, ewald.c:495-496) add up to the full result."""
    ms = cases.GOLDEN_CASES["mgcl2"]()
    whole = port.run(ms)
    parts = [port.run(ms, ithread=i, nthreads=3) for i in range(3)]
    f = sum(p["force"] for p in parts)
    assert cases.rel_rms(f, whole["force"]) < 1e-13
    assert np.allclose(sum(p["pe"] for p in parts), whole["pe"], rtol=1e-12)
    assert np.allclose(sum(p["stress"] for p in parts), whole["stress"], rtol=1e-11, atol=1e-9)


def test_newton_third_law_and_translation():
    ms = cases.GOLDEN_CASES["tips2"]()
    out = port.run(ms)
    assert np.abs(out["force"].sum(1)).max() < 1e-7 * np.abs(out["force"]).max()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("seed", [3, 4])
def test_restatement_vs_reference_live(seed):
    ms = systems.tip4p(2, seed=seed, jitter=0.05)
    gold = ref.RefLib().run(ms)
    out = port.run(ms)
    assert np.array_equal(out["force"], gold["force"])
    assert np.allclose(out["pe"], gold["pe"], rtol=1e-12)
    rng = np.random.default_rng(seed)
    npar = [2, 3, 4, 6, 3, 1, 7]
    r = ref.RefLib()
    import ctypes as C
    for ptype in (0, 1, 2, 3, 4, 6):
        p = np.zeros(8)
        p[:npar[ptype]] = rng.uniform(0.5, 3.0, npar[ptype])
        a = port.load().orc_dist_pot(p.ctypes.data_as(port.DP), 7.5, ptype)
        b = r.lib.dist_pot(p.ctypes.data_as(port.DP), 7.5, ptype)
        assert abs(a - b) <= 1e-14 * abs(b)


@pytest.mark.parametrize("name", list(cases.RDF_CASES))
def test_rdf_restatement_matches_reference_histograms(name, golden_dir):
    """RDF pass (src/force.c:1010-1103, src/rdf.c:94-108): pair counts per (id pair, bin), exact."""
    limit, nbins = cases.RDF_CASES[name]
    ms = cases.GOLDEN_CASES[name]()
    gold = np.load(os.path.join(golden_dir, "ref_rdf.npz"))[name]
    cnt = port.rdf(ms, limit, nbins)
    assert cnt.shape == gold.shape and gold.sum() > 0
    assert np.array_equal(cnt.astype(np.int64), gold)
    # the replicated-data split (icell = rank mod P, src/force.c:1051) sums to the whole
    parts = sum(port.rdf(ms, limit, nbins, ithread=r, nthreads=3) for r in range(3))
    assert np.array_equal(parts, cnt)
