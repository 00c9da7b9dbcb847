"""TEST INFRASTRUCTURE ONLY.  ctypes access to (a) the C restatement oracle/molframe.c (in libmoldy_oracle.so) and
(b) the reference's own make_sites / mol_force / mol_torque (src/algorith.c:111-217) compiled in place into
oracle/_ref/libmoldyref_mol.so -- the oracle of the next row of the hot-path contract (the
molecular-frame steps of eval_forces around force_calc/ewald).  Nothing under moldy_b200/ may import this."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DP = C.POINTER(C.c_double)
_ptr = lambda a: a.ctypes.data_as(DP)


def _port():
    L = C.CDLL(os.path.join(HERE, "libmoldy_oracle.so"))
    L.orc_make_sites.argtypes = [DP, DP, DP, DP, DP, DP, DP, C.c_int, C.c_int, C.c_int]
    L.orc_mol_force.argtypes = [DP, DP, DP, DP, C.c_int, C.c_int]
    L.orc_mol_torque.argtypes = [DP, DP, DP, DP, DP, DP, C.c_int, C.c_int]
    return L


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libmoldyref_mol.so"))


def _ref():
    return C.CDLL(os.path.join(HERE, "_ref", "libmoldyref_mol.so"))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_sites(h, com_s, quat, pfs, sitepbc: bool, impl="port"):
    """site rows [3, nmols*nsites] from scaled centres of mass, quaternions (or None) and principal-frame sites."""
    h, com_s, pfs = _c(h), _c(com_s), _c(pfs)
    nmols, nsites = com_s.shape[0], pfs.shape[0]
    site = np.zeros((3, nmols * nsites))
    q = None if quat is None else _c(quat)
    if impl == "port":
        _port().orc_make_sites(_ptr(h), _ptr(com_s), _ptr(q) if q is not None else None, _ptr(pfs),
                               _ptr(site[0]), _ptr(site[1]), _ptr(site[2]), nmols, nsites, 1 if sitepbc else 0)
    else:
        # void make_sites(mat_mt h, vec_mp c_of_m_s, quat_mp quat, vec_mp p_f_sites, real **site, int nmols, int nsites, int molflag)
        rows = (DP * 3)(_ptr(site[0]), _ptr(site[1]), _ptr(site[2]))
        R = _ref()
        R.make_sites.argtypes = [DP, DP, DP, DP, C.POINTER(DP), C.c_int, C.c_int, C.c_int]
        R.make_sites(_ptr(h), _ptr(com_s), _ptr(q) if q is not None else None, _ptr(pfs), rows, nmols, nsites,
                     0 if sitepbc else 1)          # src/defs.h: SITEPBC 0, MOLPBC 1
    return site


def mol_force(site_force, nsites: int, impl="port"):
    f = _c(site_force)
    nmols = f.shape[1] // nsites
    out = np.zeros((nmols, 3))
    if impl == "port":
        _port().orc_mol_force(_ptr(f[0]), _ptr(f[1]), _ptr(f[2]), _ptr(out), nsites, nmols)
    else:
        rows = (DP * 3)(_ptr(f[0]), _ptr(f[1]), _ptr(f[2]))
        R = _ref()
        R.mol_force.argtypes = [C.POINTER(DP), DP, C.c_int, C.c_int]
        R.mol_force(rows, _ptr(out), nsites, nmols)
    return out


def mol_torque(site_force, pfs, quat, impl="port"):
    f, pfs, q = _c(site_force), _c(pfs), _c(quat)
    nsites, nmols = pfs.shape[0], q.shape[0]
    out = np.zeros((nmols, 3))
    if impl == "port":
        _port().orc_mol_torque(_ptr(f[0]), _ptr(f[1]), _ptr(f[2]), _ptr(pfs), _ptr(out), _ptr(q), nsites, nmols)
    else:
        rows = (DP * 3)(_ptr(f[0]), _ptr(f[1]), _ptr(f[2]))
        R = _ref()
        R.mol_torque.argtypes = [C.POINTER(DP), DP, DP, DP, C.c_int, C.c_int]
        R.mol_torque(rows, _ptr(pfs), _ptr(out), _ptr(q), nsites, nmols)
    return out
