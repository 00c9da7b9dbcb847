"""Oracle for the next row of the hot-path contract (leapfrog integrator resident on the device): the C
restatement oracle/leapfrog.c -- one fused pass per molecule, the form a GPU kernel takes -- must be BIT-IDENTICAL to
the reference's own leapf_com / leapf_mom / leapf_amom / leapf_quat (src/leapfrog.c, with quaterns.c and matrix.c)
compiled in place into oracle/_ref/libmoldyref_evalf.so."""
import numpy as np
import pytest

from oracle import ref as refmod

pytestmark = pytest.mark.skipif(not refmod.available(evalf=True), reason="oracle/_ref/libmoldyref_evalf.so not built")

H_CUBIC = np.diag([19.7055, 19.7055, 19.7055])
H_TRICLINIC = np.array([[19.612, -9.806, 0.3], [0.0, 16.984, -0.7], [0.0, 0.0, 21.572]])
TIP4P_INERTIA = np.array([0.61457, 1.15511, 1.76968])
SYMMETRIC_TOP = np.array([2.5, 2.5, 0.9])
LINEAR = np.array([11.3, 11.3, 0.0])


def _state(n, seed):
    rng = np.random.default_rng(seed)
    com = rng.uniform(-0.5, 0.5, (n, 3))
    mom = rng.normal(0, 30.0, (n, 3))
    force = rng.normal(0, 2.0e3, (n, 3))
    torque = rng.normal(0, 5.0e2, (n, 3))
    quat = rng.normal(0, 1, (n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    amom = np.concatenate([np.zeros((n, 1)), rng.normal(0, 8.0, (n, 3))], axis=1)
    return com, mom, force, torque, quat, amom


@pytest.mark.parametrize("h", [H_CUBIC, H_TRICLINIC], ids=["cubic", "triclinic"])
def test_translational_substeps_bit_identical(h):
    from oracle import leapfrog as lf
    com, mom, force, torque, quat, amom = _state(777, 1)
    R = lf.Ref()
    for step, s, mass in [(0.00025, 1.0, 18.0154), (0.0005, 1.37, 24.305)]:
        assert np.array_equal(lf.leapf_com(step, com, mom, h, s, mass), R.leapf_com(step, com, mom, h, s, mass))
        assert np.array_equal(lf.leapf_mom(step, h, mom, force), R.leapf_mom(step, h, mom, force))
    # escape(): molecules that leave the cell come back into [-0.5, 0.5)
    far = lf.leapf_com(5.0, com, mom, h, 1.0, 1.0)
    assert np.array_equal(far, R.leapf_com(5.0, com, mom, h, 1.0, 1.0))
    assert far.min() >= -0.5 and far.max() <= 0.5
    assert np.array_equal(lf.leapf_amom(0.00025, amom, torque), R.leapf_amom(0.00025, amom, torque))


@pytest.mark.parametrize("inertia", [TIP4P_INERTIA, SYMMETRIC_TOP, LINEAR], ids=["asymmetric", "symmetric-top", "linear"])
@pytest.mark.parametrize("symmetric", [True, False], ids=["leapf_quat_b", "leapf_quat_a"])
def test_rotational_substep_bit_identical(inertia, symmetric):
    from oracle import leapfrog as lf
    com, mom, force, torque, quat, amom = _state(513, 2)
    amom[7, 1:] = 0.0                                        # a molecule at rest: make_rot_amom's 8*DBL_MIN guard
    R = lf.Ref(nosymmetric_rot=0 if symmetric else 1)
    for step, ts in [(0.00025, 1.0), (0.0005, 0.93)]:
        q_ref, a_ref = R.leapf_quat(step, quat, amom, inertia, ts)
        q, a, bad = lf.leapf_quat(step, quat, amom, inertia, ts, symmetric=symmetric)
        assert bad == 0
        assert np.array_equal(q, q_ref) and np.array_equal(a, a_ref)
        assert np.abs(np.linalg.norm(q, axis=1) - 1.0).max() < 1e-15
        quat, amom = q, a                                    # iterate: second call uses the kept symmetry axis


def test_symmetry_axis_is_chosen_once_from_the_first_species():
    """leapf_quat_b keeps `saxis` in a function static (src/leapfrog.c:272): the axis found for the first species it
    sees is used for every later one.  The restatement takes it as an argument."""
    from oracle import leapfrog as lf
    com, mom, force, torque, quat, amom = _state(64, 3)
    R = lf.Ref()
    first, second = SYMMETRIC_TOP, np.array([0.9, 2.5, 2.5])
    assert lf.symmetry_axis(first) == 2 and lf.symmetry_axis(second) == 0
    R.leapf_quat(0.00025, quat, amom, first, 1.0)
    q_ref, a_ref = R.leapf_quat(0.00025, quat, amom, second, 1.0)
    q, a, _ = lf.leapf_quat(0.00025, quat, amom, second, 1.0, saxis=lf.symmetry_axis(first))
    assert np.array_equal(q, q_ref) and np.array_equal(a, a_ref)
    q2, a2, _ = lf.leapf_quat(0.00025, quat, amom, second, 1.0)           # its own axis: a different splitting
    assert not np.array_equal(a2, a_ref)


@pytest.mark.parametrize("h", [H_CUBIC, H_TRICLINIC], ids=["cubic", "triclinic"])
def test_kinetic_energy_reductions_bit_identical(h):
    """trans_ke / rot_ke / energy_dyad (src/algorith.c:221-284): what tot_ke() and stress_kin() of do_step need from the
    momenta every step."""
    from oracle import leapfrog as lf
    com, mom, force, torque, quat, amom = _state(1001, 4)
    R = lf.Ref()
    for s, mass in [(1.0, 18.0154), (1.21, 35.453)]:
        assert lf.trans_ke(h, mom, s, mass) == lf.trans_ke(h, mom, s, mass, impl="ref", ref=R)
        assert np.array_equal(lf.energy_dyad(h, s, mom, mass), lf.energy_dyad(h, s, mom, mass, impl="ref", ref=R))
    for inertia in (TIP4P_INERTIA, LINEAR):
        assert lf.rot_ke(amom, 1.1, inertia) == lf.rot_ke(amom, 1.1, inertia, impl="ref", ref=R)
    # trace of the dyad is twice the kinetic energy
    assert abs(np.trace(lf.energy_dyad(h, 1.0, mom, 18.0)) - 2 * lf.trans_ke(h, mom, 1.0, 18.0)) < 1e-9 * lf.trans_ke(h, mom, 1.0, 18.0)
