"""TEST INFRASTRUCTURE ONLY.  Restatement of eval_forces() (src/accel.c:398-617) around the reference's own
force_calc()/ewald(): the steps the device path fuses (second make_sites, dipole moment, surface-dipole term,
mol_force/mol_torque, site->molecular virial, distant-potential constants), written the way mdb_molframe.cu evaluates
them -- in particular the virial correction as the per-molecule sum  sum_s f_i (r_j - R_j)  instead of the reference's
difference of two system-wide sums (src/accel.c:585-593).  tests/test_oracle_evalf.py pins it against the reference's
eval_forces() compiled in place (oracle/_ref/libmoldyref_evalf.so).  Nothing under moldy_b200/ may import this."""
from __future__ import annotations

import numpy as np

from oracle import molframe, ref as refmod


def _distant_const(R, ms, iflag):
    """distant_const, src/accel.c:293-326 (dist_pot / poteval are the reference's own)."""
    sd = ms.sysdef
    count = np.zeros(sd.max_id)
    for sp in sd.species:
        for sid in sp.site_id:
            count[sid] += sp.nmols
    rc, c = ms.control.cutoff, 0.0
    import ctypes as C
    DP = C.POINTER(C.c_double)
    for i in range(1, sd.max_id):
        for j in range(1, sd.max_id):
            p = np.ascontiguousarray(sd.potpar[j, i], dtype=np.float64)
            c -= 2 * np.pi * count[i] * count[j] * R.lib.dist_pot(p.ctypes.data_as(DP), rc, sd.ptype)
            if iflag:
                c += 2.0 / 3.0 * np.pi * count[i] * count[j] * rc ** 3 * R.lib.poteval(p.ctypes.data_as(DP), rc, sd.ptype, 0.0)
    return c


def _sites(ms, second):
    """make_sites per species with the reference's own routine: first pass (control.molpbc ? MOLPBC : SITEPBC,
    src/accel.c:500-504) or second pass (framework ? SITEPBC : MOLPBC, :537-542)."""
    from moldy_b200 import abi
    out = np.zeros((3, abi.nsarray(ms.nsites)))
    m0 = s0 = 0
    for sp in ms.sysdef.species:
        q = ms.quat[m0:m0 + sp.nmols] if sp.rdof else None
        pbc = bool(sp.framework) if second else not ms.control.molpbc
        blk = molframe.make_sites(ms.h, ms.c_of_m[m0:m0 + sp.nmols], q, sp.p_f_sites, pbc, impl="ref")
        out[:, s0:s0 + sp.nmols * sp.nsites] = blk
        m0 += sp.nmols
        s0 += sp.nmols * sp.nsites
    return out


def tail(ms, f, pe, stress, R=None):
    """Everything eval_forces() does after the global sums of the site forces (src/accel.c:537-608), in the device path's
    formulation.  f[3,N], pe[2], stress[3,3] (upper triangle) are the COMPLETE sums of force_calc + ewald -- under SPMD
    the all-reduced block -- so every rank that runs this gets the same molecular forces, torques, energies and stress."""
    from moldy_b200.systems import quat_to_rot
    R = R or refmod.RefLib()
    sd = ms.sysdef
    n = ms.nsites
    f, pe, stress = np.array(f, dtype=np.float64), np.array(pe, dtype=np.float64), np.array(stress, dtype=np.float64)
    vol = abs(np.linalg.det(ms.h))
    chg = ms.charges()
    dip = np.zeros(3)
    if ms.control.alpha > 1e-7:
        s2 = _sites(ms, True)[:, :n]
        dip = (s2 * chg).sum(axis=1)
        if ms.control.surface_dipole:
            f -= (4.0 * np.pi / (3.0 * vol) * dip)[:, None] * chg[None, :]
            pe[1] += 2.0 * np.pi / (3.0 * vol) * (dip ** 2).sum()
    force, torque = [], []
    V = np.zeros((3, 3))
    m0 = s0 = 0
    for sp in sd.species:
        ns = sp.nmols * sp.nsites
        fs = f[:, s0:s0 + ns]
        force.append(molframe.mol_force(fs, sp.nsites, impl="ref"))
        pfs = np.asarray(sp.p_f_sites, dtype=np.float64)
        if sp.rdof:
            q = ms.quat[m0:m0 + sp.nmols]
            torque.append(molframe.mol_torque(fs, pfs, q, impl="ref"))
        # d = site - centre of mass of the MOLPBC sites (non-framework) or the principal-frame site (framework)
        if sp.rdof and not sp.framework:
            d = np.einsum("mij,sj->msi", quat_to_rot(q), pfs)
        else:
            d = np.broadcast_to(pfs[None], (sp.nmols, sp.nsites, 3))
        V += np.einsum("ims,msj->ij", fs.reshape(3, sp.nmols, sp.nsites), d)
        m0 += sp.nmols
        s0 += ns
    stress = np.triu(stress) + np.triu(stress, 1).T
    stress -= V
    pe[0] += _distant_const(R, ms, 0) / vol
    stress += np.eye(3) * _distant_const(R, ms, 1) / vol
    return dict(pe=pe, dip_mom=dip, stress=stress, force=np.concatenate(force),
                torque=np.concatenate(torque) if torque else np.zeros((0, 3)))


def eval_forces(ms):
    """pe[2], dip_mom[3], stress[3,3], force[nmols,3], torque[nmols_r,3] from the reference's force_calc/ewald
    site forces plus the restated molecular-frame steps."""
    R = refmod.RefLib()
    r = R.run(ms, sites=_sites(ms, False))
    return tail(ms, r["force"], r["pe"], r["stress"], R)
