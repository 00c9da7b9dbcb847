"""TEST INFRASTRUCTURE ONLY -- ctypes driver for the C restatement of the hot path
(oracle/moldy_oracle.c -> oracle/libmoldy_oracle.so).  Same call shape as
oracle/ref.py: run(ms) -> forces / pe / stress with the reference's constants
(eintra, self and sheet energy) applied the way force_calc()/ewald() apply them."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmoldy_oracle.so")
DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)


class orc_system(C.Structure):
    _fields_ = [("nsites", C.c_int), ("nsites_xf", C.c_int), ("max_id", C.c_int), ("ptype", C.c_int),
                ("n_potpar", C.c_int), ("site_type", IP), ("site_mol", IP), ("chg", DP), ("potpar", DP),
                ("h", C.c_double * 9), ("cutoff", C.c_double), ("subcell", C.c_double), ("alpha", C.c_double),
                ("k_cutoff", C.c_double), ("strict_cutoff", C.c_int), ("molpbc", C.c_int), ("c_of_m", DP),
                ("ithread", C.c_int), ("nthreads", C.c_int)]


class orc_result(C.Structure):
    _fields_ = [("pe", C.c_double), ("stress", C.c_double * 9), ("npairs", C.c_double), ("n_too_close", C.c_int),
                ("n_bin_errors", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("n_nabors", C.c_int),
                ("nhkl", C.c_int)]


_L = None


def load():
    global _L
    if _L is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "moldy_oracle.c")):
            subprocess.check_call(["make", "-s", "-C", HERE, "libmoldy_oracle.so"])
        _L = C.CDLL(LIB)
        _L.orc_eintra.restype = C.c_double
        _L.orc_err_fn.restype = C.c_double
        _L.orc_err_fn.argtypes = [C.c_double]
        _L.orc_dist_pot.restype = C.c_double
        _L.orc_dist_pot.argtypes = [DP, C.c_double, C.c_int]
        _L.orc_cellbin.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, IP]
        _L.orc_pair.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, DP, DP, DP]
    return _L


def _system(ms, ithread=0, nthreads=1):
    sd = ms.sysdef
    keep = dict(ids=np.ascontiguousarray(ms.site_ids(), dtype=np.int32),
                mol=np.ascontiguousarray(ms.molmap(), dtype=np.int32),
                chg=np.ascontiguousarray(ms.charges()),
                pot=np.ascontiguousarray(sd.potpar.reshape(-1), dtype=np.float64),
                com=np.ascontiguousarray(ms.c_of_m, dtype=np.float64))
    s = orc_system()
    s.nsites, s.nsites_xf = ms.nsites, ms.nsites_xf
    s.max_id, s.ptype, s.n_potpar = sd.max_id, sd.ptype, sd.n_potpar
    s.site_type = keep["ids"].ctypes.data_as(IP)
    s.site_mol = keep["mol"].ctypes.data_as(IP)
    s.chg = keep["chg"].ctypes.data_as(DP)
    s.potpar = keep["pot"].ctypes.data_as(DP)
    for i in range(9):
        s.h[i] = float(ms.h.reshape(-1)[i])
    c = ms.control
    s.cutoff, s.subcell, s.alpha, s.k_cutoff, s.strict_cutoff = c.cutoff, c.subcell, c.alpha, c.k_cutoff, c.strict_cutoff
    s.molpbc = int(c.molpbc)
    s.c_of_m = keep["com"].ctypes.data_as(DP)
    s.ithread, s.nthreads = ithread, nthreads
    return s, keep


def _species(ms):
    sp = ms.sysdef.species
    ns = np.array([x.nsites for x in sp], dtype=np.int32)
    nm = np.array([x.nmols for x in sp], dtype=np.int32)
    fw = np.array([int(x.framework) for x in sp], dtype=np.int32)
    pfs = np.ascontiguousarray(np.concatenate([x.p_f_sites for x in sp]), dtype=np.float64)
    return ns, nm, fw, pfs


def constants(ms):
    """(eintra, self_energy, sheet_energy) -- the reference's first-call statics."""
    L = load()
    s, keep = _system(ms)
    ns, nm, fw, pfs = _species(ms)
    args = (C.byref(s), len(ns), ns.ctypes.data_as(IP), nm.ctypes.data_as(IP), fw.ctypes.data_as(IP),
            pfs.ctypes.data_as(DP))
    eintra = L.orc_eintra(*args)
    se, sh = C.c_double(0), C.c_double(0)
    if ms.control.alpha > 1e-7:
        L.orc_self_energy(*args, C.byref(se), C.byref(sh))
    return eintra, se.value, sh.value


def cell_ids(ms, sites=None):
    L = load()
    s, keep = _system(ms)
    site = np.ascontiguousarray(ms.make_sites(wrap=not ms.control.molpbc) if sites is None else sites)
    n = ms.nsites
    x, y, z = (np.ascontiguousarray(site[i, :n]) for i in range(3))
    out = np.empty(n, dtype=np.int32)
    L.orc_cell_ids(C.byref(s), x.ctypes.data_as(DP), y.ctypes.data_as(DP), z.ctypes.data_as(DP), out.ctypes.data_as(IP))
    return out


def run(ms, real=True, recip=True, sites=None, ithread=0, nthreads=1):
    L = load()
    s, keep = _system(ms, ithread, nthreads)
    site = np.ascontiguousarray(ms.make_sites(wrap=not ms.control.molpbc) if sites is None else sites)
    n = ms.nsites
    x, y, z = (np.ascontiguousarray(site[i, :n]) for i in range(3))
    f = np.zeros((3, n))
    ptr = lambda a: a.ctypes.data_as(DP)
    fr = [np.ascontiguousarray(f[i]) for i in range(3)]
    pe = np.zeros(2)
    stress = np.zeros(9)
    info = {}
    eintra, self_e, sheet_e = constants(ms)
    vol = float(np.linalg.det(ms.h))
    if real:
        r = orc_result()
        if L.orc_force_calc(C.byref(s), ptr(x), ptr(y), ptr(z), ptr(fr[0]), ptr(fr[1]), ptr(fr[2]), C.byref(r)):
            raise RuntimeError("Cutoff radius > 1 * cell dimension")
        if ithread == 0:
            pe[0] -= eintra
        pe[0] += r.pe
        stress += np.array(r.stress[:])
        info.update(grid=(r.nx, r.ny, r.nz), n_nabors=r.n_nabors, npairs=r.npairs, too_close=r.n_too_close,
                    bin_errors=r.n_bin_errors)
    if recip and ms.control.alpha > 1e-7:
        r = orc_result()
        if ithread == 0:
            pe[1] -= self_e
            pe[1] += sheet_e / vol
            stress[[0, 4, 8]] += sheet_e / vol
        L.orc_ewald(C.byref(s), ptr(x), ptr(y), ptr(z), ptr(fr[0]), ptr(fr[1]), ptr(fr[2]), C.byref(r))
        pe[1] += r.pe
        stress += np.array(r.stress[:])
        info.update(nhkl=r.nhkl)
    return dict(force=np.stack(fr), pe=pe, stress=stress.reshape(3, 3), eintra=eintra, self_energy=self_e,
                sheet_energy=sheet_e, **info)


def rdf(ms, limit, nbins, sites=None, ithread=0, nthreads=1):
    """Pair counts per (id pair, bin) of the RDF pass: float64 [max_id (max_id-1)/2, nbins]."""
    L = load()
    s, keep = _system(ms, ithread, nthreads)
    site = np.ascontiguousarray(ms.make_sites(wrap=not ms.control.molpbc) if sites is None else sites)
    n = ms.nsites
    x, y, z = (np.ascontiguousarray(site[i, :n]) for i in range(3))
    mid = ms.sysdef.max_id
    counts = np.zeros((mid * (mid - 1) // 2, nbins))
    ptr = lambda a: a.ctypes.data_as(DP)
    if L.orc_rdf(C.byref(s), ptr(x), ptr(y), ptr(z), C.c_double(limit), C.c_int(nbins), ptr(counts)):
        raise RuntimeError("RDF limit > cell dimension")
    return counts
