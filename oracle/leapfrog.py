"""TEST INFRASTRUCTURE ONLY.  ctypes access to (a) the C restatement oracle/leapfrog.c (in libmoldy_oracle.so) and
(b) the reference's own leapf_com / leapf_mom / leapf_amom / leapf_quat (src/leapfrog.c) compiled in place into
oracle/_ref/libmoldyref_evalf.so -- the oracle of SURVEY 8f rank 4 (a device-resident leapfrog integrator around
eval_forces).  Nothing under moldy_b200/ may import this."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from oracle import ref as refmod

HERE = os.path.dirname(os.path.abspath(__file__))
DP = C.POINTER(C.c_double)
_p = lambda a: a.ctypes.data_as(DP)
_c = lambda a: np.ascontiguousarray(a, dtype=np.float64)


def _port():
    L = C.CDLL(os.path.join(HERE, "libmoldy_oracle.so"))
    L.orc_leapf_com.argtypes = [C.c_double, DP, DP, DP, C.c_double, C.c_double, C.c_int]
    L.orc_leapf_mom.argtypes = [C.c_double, DP, DP, DP, C.c_int]
    L.orc_leapf_amom.argtypes = [C.c_double, DP, DP, C.c_int]
    L.orc_leapf_quat.argtypes = [C.c_double, DP, DP, DP, C.c_double, C.c_int, C.c_int, C.c_int]
    L.orc_symmetry_axis.argtypes = [DP]
    return L


class Ref:
    """One private copy of the reference library: leapf_quat_b keeps its symmetry axis in a function static."""

    def __init__(self, nosymmetric_rot=0):
        self.R = refmod.RefLib(evalf=True)
        self.R.control.nosymmetric_rot = int(nosymmetric_rot)
        self.R.control.const_temp = 0
        L = self.R.lib
        L.leapf_com.argtypes = [C.c_double, DP, DP, DP, C.c_double, C.c_double, C.c_int]
        L.leapf_mom.argtypes = [C.c_double, DP, DP, DP, C.c_int]
        L.leapf_amom.argtypes = [C.c_double, DP, DP, C.c_int]
        L.leapf_quat.argtypes = [C.c_double, DP, DP, DP, DP, C.c_double, C.c_int]
        for f in (L.leapf_com, L.leapf_mom, L.leapf_amom, L.leapf_quat):
            f.restype = None

    def leapf_com(self, step, com, mom, h, s, mass):
        com = _c(com).copy()
        self.R.lib.leapf_com(step, _p(com), _p(_c(mom)), _p(_c(h)), s, mass, len(com))
        return com

    def leapf_mom(self, step, h, mom, force):
        mom = _c(mom).copy()
        self.R.lib.leapf_mom(step, _p(_c(h)), _p(mom), _p(_c(force)), len(mom))
        return mom

    def leapf_amom(self, step, amom, torque):
        amom = _c(amom).copy()
        self.R.lib.leapf_amom(step, _p(amom), _p(_c(torque)), len(amom))
        return amom

    def leapf_quat(self, step, quat, amom, inertia, ts):
        quat, amom = _c(quat).copy(), _c(amom).copy()
        smom = np.zeros(1)
        self.R.lib.leapf_quat(step, _p(quat), _p(amom), _p(_c(inertia)), _p(smom), ts, len(quat))
        return quat, amom


def leapf_com(step, com, mom, h, s, mass):
    com = _c(com).copy()
    _port().orc_leapf_com(step, _p(com), _p(_c(mom)), _p(_c(h)), s, mass, len(com))
    return com


def leapf_mom(step, h, mom, force):
    mom = _c(mom).copy()
    _port().orc_leapf_mom(step, _p(_c(h)), _p(mom), _p(_c(force)), len(mom))
    return mom


def leapf_amom(step, amom, torque):
    amom = _c(amom).copy()
    _port().orc_leapf_amom(step, _p(amom), _p(_c(torque)), len(amom))
    return amom


def symmetry_axis(inertia) -> int:
    return int(_port().orc_symmetry_axis(_p(_c(inertia))))


def leapf_quat(step, quat, amom, inertia, ts, symmetric=True, saxis=None):
    quat, amom = _c(quat).copy(), _c(amom).copy()
    if saxis is None:
        saxis = symmetry_axis(inertia)
    bad = _port().orc_leapf_quat(step, _p(quat), _p(amom), _p(_c(inertia)), ts, len(quat), 1 if symmetric else 0, saxis)
    return quat, amom, bad


# ---- kinetic-energy reductions (trans_ke, rot_ke, energy_dyad: src/algorith.c:221-284) ----
def trans_ke(h, mom, s, mass, impl="port", ref=None):
    h, mom = _c(h), _c(mom)
    if impl == "port":
        L = _port()
        L.orc_trans_ke.restype = C.c_double
        L.orc_trans_ke.argtypes = [DP, DP, C.c_double, C.c_double, C.c_int]
        return L.orc_trans_ke(_p(h), _p(mom), s, mass, len(mom))
    L = (ref or Ref()).R.lib
    L.trans_ke.restype = C.c_double
    L.trans_ke.argtypes = [DP, DP, C.c_double, C.c_double, C.c_int]
    return L.trans_ke(_p(h), _p(mom), s, mass, len(mom))


def rot_ke(amom, s, inertia, impl="port", ref=None):
    amom, inertia = _c(amom), _c(inertia)
    if impl == "port":
        L = _port()
        L.orc_rot_ke.restype = C.c_double
        L.orc_rot_ke.argtypes = [DP, C.c_double, DP, C.c_int]
        return L.orc_rot_ke(_p(amom), s, _p(inertia), len(amom))
    L = (ref or Ref()).R.lib
    L.rot_ke.restype = C.c_double
    L.rot_ke.argtypes = [DP, C.c_double, DP, C.c_int]
    return L.rot_ke(_p(amom), s, _p(inertia), len(amom))


def energy_dyad(h, s, mom, mass, impl="port", ref=None):
    h, mom = _c(h), _c(mom)
    out = np.zeros((3, 3))
    if impl == "port":
        L = _port()
        L.orc_energy_dyad.argtypes = [DP, DP, C.c_double, DP, C.c_double, C.c_int]
        L.orc_energy_dyad(_p(out), _p(h), s, _p(mom), mass, len(mom))
    else:
        L = (ref or Ref()).R.lib
        L.energy_dyad.restype = None
        L.energy_dyad.argtypes = [DP, DP, C.c_double, DP, C.c_double, C.c_int]
        L.energy_dyad(_p(out), _p(h), s, _p(mom), mass, len(mom))
    return out
