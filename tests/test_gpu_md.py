"""GPU tests of SURVEY 8f rank 4: the NVE leapfrog step of do_step() (src/accel.c:626-827) with the dynamic state resident
in HBM (moldy_b200/csrc/mdb_md.cu), and the library's own do_step() symbol.

Oracles: oracle/leapfrog.c (the restatement, bit-identical to the reference's leapfrog.c / quaterns.c / matrix.c:
tests/test_oracle_leapfrog.py) for the sub-steps, and the reference's own do_step() compiled in place
(oracle/_ref/libmoldyref_evalf.so) for whole steps.  Bar: translational sub-steps bit-exact; rotational sub-steps and
the reductions <= 1e-13 (device sin/cos against libm, tree sums against sequential sums); trajectories over several steps
<= 1e-10 in the co-ordinates, energies <= 1e-11."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from moldy_b200 import lib, systems
from tests import cases

pytestmark = pytest.mark.gpu
STEP = 0.0005


def _setup(name, seed=7):
    ms = cases.GOLDEN_CASES[name]()
    mom, amom = ms.thermal_momenta(seed=seed)
    eng = lib.Engine(0)
    eng.configure(ms)
    md = lib.MdState(eng, ms)
    md.upload(ms.c_of_m, ms.quat, mom, amom)
    return ms, mom, amom, eng, md


def _species_slices(ms):
    m0 = 0
    for s, inert in zip(ms.sysdef.species, ms.principal_inertia()):
        yield s, inert, slice(m0, m0 + s.nmols)
        m0 += s.nmols


@pytest.mark.parametrize("name", ["tip4p", "mgcl2", "quartz"])
def test_substeps_against_the_restatement(name):
    from oracle import leapfrog as lf
    ms, mom, amom, eng, md = _setup(name)
    L, h = eng.L, np.ascontiguousarray(ms.h)
    st = torch.cuda.current_stream().cuda_stream
    # leapf_all_coords(step/2)
    assert L.mdb_md_coords(eng.h, h.ctypes.data, 0.5 * STEP, 1.0, st) == 0
    got = md.download(st)
    saxis = None
    for s, inert, sl in _species_slices(ms):
        want = lf.leapf_com(0.5 * STEP, ms.c_of_m[sl], mom[sl], ms.h, 1.0, s.mass)
        assert np.array_equal(got["com"][sl], want), s.name                      # bit-exact: decides the cell assignment
        if s.rdof:
            if saxis is None:
                saxis = lf.symmetry_axis(inert)
            q, a, bad = lf.leapf_quat(0.5 * STEP, ms.quat[sl], amom[sl], inert, 1.0, symmetric=True, saxis=saxis)
            assert bad == 0
            assert np.abs(got["quat"][sl] - q).max() < 1e-13 and np.abs(got["amom"][sl] - a).max() < 1e-13 * np.abs(a).max()
    # forces of the new configuration, then leapf_all_momenta(step/2)
    assert L.mdb_md_eval_forces(eng.h, h.ctypes.data, ms.control.surface_dipole, int(ms.control.alpha > 1e-7), st) == 0
    before = md.download(st)
    assert L.mdb_md_momenta(eng.h, h.ctypes.data, 0.5 * STEP, st) == 0
    after = md.download(st)
    for s, inert, sl in _species_slices(ms):
        assert np.array_equal(after["mom"][sl], lf.leapf_mom(0.5 * STEP, ms.h, before["mom"][sl], before["force"][sl])), s.name
        if s.rdof:
            assert np.array_equal(after["amom"][sl], lf.leapf_amom(0.5 * STEP, before["amom"][sl], before["torque"][sl]))
    # reductions: trans_ke, energy_dyad, rot_ke
    sums = md.sums_now(st)
    for i, (s, inert, sl) in enumerate(_species_slices(ms)):
        ke = lf.trans_ke(ms.h, after["mom"][sl], 1.0, s.mass)
        assert abs((sums[i, 0] + sums[i, 3] + sums[i, 5]) / (2 * s.mass) - ke) <= 1e-13 * max(ke, 1e-300)
        dy = lf.energy_dyad(ms.h, 1.0, after["mom"][sl], s.mass)
        mine = np.array([[sums[i, 0], sums[i, 1], sums[i, 2]], [sums[i, 1], sums[i, 3], sums[i, 4]], [sums[i, 2], sums[i, 4], sums[i, 5]]]) / s.mass
        assert np.abs(mine - dy).max() <= 1e-13 * np.abs(dy).max()
        if s.rdof:
            rk = lf.rot_ke(after["amom"][sl], 1.0, inert)
            mine_r = 0.5 * sum(sums[i, 6 + k] / inert[k] for k in range(3) if inert[k] > 1e-14)
            assert abs(mine_r - rk) <= 1e-13 * rk
    eng.close()


def _ref_steps(ms, mom, amom, nsteps):
    from oracle import ref
    R = ref.RefLib(evalf=True)
    args, out, set_control = ms.do_step_args(mom, amom, STEP)
    set_control(R.control)
    R.lib.do_step.restype = None
    pes = []
    for k in range(nsteps):
        R.control.istep = 2 + k
        R.lib.do_step(*args)
        pes.append(out["pe"].copy())
    return out, np.array(pes)


@pytest.mark.parametrize("name", ["tip4p", "mgcl2", "quartz", "tips2"])
def test_resident_steps_against_the_references_do_step(name):
    """Five NVE steps: the device-resident integrator against the reference's do_step() run on the CPU."""
    from oracle import ref
    if not ref.available(evalf=True):
        pytest.skip("oracle/_ref/libmoldyref_evalf.so not built")
    nsteps = 5
    ms, mom, amom, eng, md = _setup(name)
    want, pes = _ref_steps(cases.GOLDEN_CASES[name](), mom, amom, nsteps)
    lib.reset()
    ms.control.fill(lib.control())                 # eintra / self energy are the ABI layer's: compare the device scalars
    got_pe = []
    for k in range(nsteps):
        sc = md.step(STEP)
        sd_term = 2 * np.pi / (3 * np.linalg.det(ms.h)) * (sc[:3] ** 2).sum() if ms.control.surface_dipole and ms.control.alpha > 1e-7 else 0.0
        got_pe.append([sc[12], sc[13] + sd_term])        # + the surface-dipole energy eval_forces adds on the host (src/accel.c:557)
    got = md.download()
    eng.close()
    assert np.abs(got["com"] - want["com"]).max() < 1e-10
    assert np.abs(got["mom"] - want["mom"]).max() < 1e-10 * np.abs(want["mom"]).max()
    rot = np.concatenate([np.full(s.nmols, bool(s.rdof)) for s in ms.sysdef.species])
    if rot.any():
        assert np.abs(got["quat"][rot] - want["quat"][rot]).max() < 1e-10
        assert np.abs(got["amom"][rot] - want["amom"][rot]).max() < 1e-10 * np.abs(want["amom"]).max()
    # energies step by step: the device block lacks the constants eval_forces adds on the host; differences of steps do not
    got_pe = np.array(got_pe)
    d_ref = pes.sum(1) - pes.sum(1)[0]
    d_got = got_pe.sum(1) - got_pe.sum(1)[0]
    assert np.abs(d_got - d_ref).max() < 1e-9 * np.abs(pes.sum(1)).max()


@pytest.mark.parametrize("name", ["tip4p", "mgcl2"])
def test_library_do_step_symbol_against_the_references(name):
    """do_step() of libmoldy_b200.so (Moldy's prototype): state arrays advanced in place, pe / stress / meansq / H_0."""
    from oracle import ref
    if not ref.available(evalf=True):
        pytest.skip("oracle/_ref/libmoldyref_evalf.so not built")
    ms = cases.GOLDEN_CASES[name]()
    mom, amom = ms.thermal_momenta(seed=11)
    want, pes = _ref_steps(cases.GOLDEN_CASES[name](), mom, amom, 3)
    lib.reset()
    got = lib.do_step(ms, mom, amom, STEP, nsteps=3)
    lib.reset()
    assert np.abs(got["com"] - want["com"]).max() < 1e-10
    assert np.abs(got["mom"] - want["mom"]).max() < 1e-10 * np.abs(want["mom"]).max()
    assert np.abs(got["quat"] - want["quat"]).max() < 1e-10
    assert np.abs(got["pe"] - want["pe"]).max() < 1e-11 * np.abs(want["pe"]).max()
    assert np.linalg.norm(got["stress"] - want["stress"]) < 1e-10 * np.linalg.norm(want["stress"])
    assert np.allclose(got["meansq"], want["meansq"], rtol=1e-10, atol=1e-12 * np.abs(want["meansq"]).max())
    assert np.allclose(got["dip_mom"], want["dip_mom"], rtol=1e-9, atol=1e-9 * np.abs(want["dip_mom"]).max() + 1e-12)


def test_h0_at_the_first_step():
    """istep == 1: H_0 = KE(half step) + PE (src/accel.c:718-726)."""
    from oracle import ref
    if not ref.available(evalf=True):
        pytest.skip("oracle/_ref/libmoldyref_evalf.so not built")
    ms = cases.GOLDEN_CASES["tip4p"]()
    mom, amom = ms.thermal_momenta(seed=5)
    R = ref.RefLib(evalf=True)
    args, out, set_control = cases.GOLDEN_CASES["tip4p"]().do_step_args(mom, amom, STEP)
    set_control(R.control)
    R.control.istep = 1
    R.lib.do_step.restype = None
    R.lib.do_step(*args)
    h0_ref = out["sysm"].H_0
    lib.reset()
    args2, out2, set2 = ms.do_step_args(mom, amom, STEP)
    set2(lib.control())
    c = lib.control()
    c.istep = 1
    lib.set_thread(0, 1)
    lib.load().do_step(*args2)
    lib.reset()
    assert abs(out2["sysm"].H_0 - h0_ref) < 1e-11 * abs(h0_ref)


def test_resident_and_uploading_paths_agree(monkeypatch):
    """State kept in HBM across steps (mdb_md_step) against do_step() uploading and downloading it every step: the same
    kernels on the same numbers -- bit for bit with the bit-reproducible pair kernel (pair_mode 3; the default Newton-3
    kernel accumulates with atomics and varies in the last bits from run to run)."""
    monkeypatch.setenv("MDB_PAIR_MODE", "3")
    lib.shutdown()
    ms, mom, amom, eng, md = _setup("tip4p", seed=3)
    for _ in range(3):
        md.step(STEP)
    a = md.download()
    eng.close()
    lib.reset()
    b = lib.do_step(cases.GOLDEN_CASES["tip4p"](), mom, amom, STEP, nsteps=3)
    lib.shutdown()
    assert np.array_equal(a["com"], b["com"]) and np.array_equal(a["quat"], b["quat"])
    assert np.array_equal(a["mom"], b["mom"]) and np.array_equal(a["amom"], b["amom"])
