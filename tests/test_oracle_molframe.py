"""Oracle of the NEXT row of the hot-path contract (the molecular-frame steps of eval_forces() that
surround force_calc()/ewald()): our C restatement oracle/molframe.c against the reference's own make_sites /
mol_force / mol_torque (src/algorith.c:111-217, compiled in place into oracle/_ref/libmoldyref_mol.so) on seeded
inputs.  Site positions decide the cell assignment, which the contract wants bit-exact: the restatement keeps the
reference's operation order and must reproduce its sites, forces and torques bit for bit.  CPU only."""
import numpy as np
import pytest

from oracle import molframe

pytestmark = pytest.mark.skipif(not molframe.ref_available(), reason="oracle/_ref/libmoldyref_mol.so not built")


def _cells():
    cubic = np.diag([19.7055, 19.7055, 19.7055])
    tric = np.array([[24.3, 3.1, -2.2], [0.0, 21.7, 4.4], [0.0, 0.0, 18.9]])       # upper triangular, as Moldy keeps h
    return {"cubic": cubic, "triclinic": tric}


def _molecules(rng, nmols, nsites):
    com_s = rng.uniform(-0.5, 0.5, (nmols, 3))
    q = rng.normal(size=(nmols, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    pfs = rng.normal(scale=0.8, size=(nsites, 3))
    return com_s, q, pfs


@pytest.mark.parametrize("cell", ["cubic", "triclinic"])
@pytest.mark.parametrize("nsites", [1, 3, 4, 7])
@pytest.mark.parametrize("sitepbc", [True, False])
def test_make_sites_bit_identical_to_reference(cell, nsites, sitepbc):
    rng = np.random.default_rng(100 * nsites + (7 if sitepbc else 0))
    h = _cells()[cell]
    com_s, q, pfs = _molecules(rng, 257, nsites)
    quat = None if nsites == 1 else q                  # monatomic species carry no quaternions (src/accel.c:497-500)
    a = molframe.make_sites(h, com_s, quat, pfs, sitepbc, impl="port")
    b = molframe.make_sites(h, com_s, quat, pfs, sitepbc, impl="ref")
    assert np.array_equal(a, b)
    if sitepbc:                                        # every site inside the cell (scaled co-ordinates in [-1/2, 1/2])
        s = np.linalg.inv(h) @ a
        assert np.all(np.abs(s) <= 0.5 + 1e-12)


@pytest.mark.parametrize("nsites", [1, 3, 4, 7])
def test_mol_force_and_torque_bit_identical_to_reference(nsites):
    rng = np.random.default_rng(nsites)
    nmols = 311
    _, q, pfs = _molecules(rng, nmols, nsites)
    f = rng.normal(scale=50.0, size=(3, nmols * nsites))
    assert np.array_equal(molframe.mol_force(f, nsites, "port"), molframe.mol_force(f, nsites, "ref"))
    assert np.array_equal(molframe.mol_torque(f, pfs, q, "port"), molframe.mol_torque(f, pfs, q, "ref"))


def test_torque_of_a_pure_translation_vanishes():
    """Property pin that does not need the reference: equal forces on all sites of a molecule whose principal-frame
    sites are centred give no torque, and the molecular force is nsites times the site force."""
    rng = np.random.default_rng(5)
    nmols, nsites = 64, 4
    _, q, pfs = _molecules(rng, nmols, nsites)
    pfs -= pfs.mean(axis=0)
    fm = rng.normal(size=(nmols, 3))
    f = np.repeat(fm.T[:, :, None], nsites, axis=2).reshape(3, nmols * nsites)
    assert np.allclose(molframe.mol_force(f, nsites), nsites * fm, rtol=0, atol=1e-13)
    assert np.abs(molframe.mol_torque(f, pfs, q)).max() < 1e-13
