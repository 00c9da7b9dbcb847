"""TEST INFRASTRUCTURE ONLY -- ctypes driver for the reference's own hot path.

Loads oracle/_ref/libmoldyref.so (force.c, kernel.c, ewald.c of the reference
compiled from where they lie by oracle/Makefile; nothing is copied) and calls
its `force_calc()` / `ewald()` on a `moldy_b200.systems.MoldySystem`.

The reference keeps first-call state in function statics (src/force.c:1146-1149,
src/ewald.c:352-355), so every `RefLib()` instance loads a *private copy* of the
shared object (copied to a temp file) -- one instance per configuration.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

from moldy_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available(fast: bool = False, evalf: bool = False) -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libmoldyref_evalf.so" if evalf else
                                       "libmoldyref_fast.so" if fast else "libmoldyref.so"))


class RefLib:
    def __init__(self, fast: bool = False, evalf: bool = False):
        """evalf: load libmoldyref_evalf.so instead -- the same hot-path files plus the reference's accel.c, algorith.c,
        quaterns.c and leapfrog.c compiled in place, for eval_forces() (src/accel.c:398-617)."""
        src = os.path.join(REF_DIR, "libmoldyref_evalf.so" if evalf else
                           "libmoldyref_fast.so" if fast else "libmoldyref.so")
        if not os.path.exists(src):
            raise FileNotFoundError(f"{src} missing: run `make -C oracle ref` where /root/reference exists")
        fd, self._tmp = tempfile.mkstemp(suffix=".so", prefix="moldyref_")
        os.close(fd)
        shutil.copyfile(src, self._tmp)
        self.lib = C.CDLL(self._tmp)
        os.unlink(self._tmp)                      # mapping stays valid
        L = self.lib
        L.mdref_control.restype = C.POINTER(abi.contr_mt)
        L.mdref_log.restype = C.c_char_p
        for f in ("mdref_sizeof_control", "mdref_sizeof_system", "mdref_sizeof_spec", "mdref_sizeof_pot"):
            getattr(L, f).restype = C.c_size_t
        L.mdref_warnings.restype = C.c_long
        L.poteval.restype = C.c_double
        L.poteval.argtypes = [C.POINTER(C.c_double), C.c_double, C.c_int, C.c_double]
        L.dist_pot.restype = C.c_double
        L.dist_pot.argtypes = [C.POINTER(C.c_double), C.c_double, C.c_int]
        L.err_fn.restype = C.c_double
        L.err_fn.argtypes = [C.c_double]
        L.cellbin.restype = C.c_int
        L.cellbin.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double]
        self.control = L.mdref_control().contents

    def set_thread(self, ithread: int, nthreads: int):
        self.lib.mdref_set_thread(ithread, nthreads)

    def log(self) -> str:
        return self.lib.mdref_log().decode()

    def run(self, ms, real: bool = True, recip: bool = True, sites=None, rdf=None):
        """force_calc (+ ewald when alpha > ALPHAMIN) exactly as eval_forces()
        sequences them (src/accel.c:520-527).  Returns forces[3,N], pe[2], stress[3,3].
        rdf=(limit, nbins): switch on the RDF pass (src/force.c:1302-1313) and return the reference's
        own float histograms (init_rdf / rdf_accum / rdf_ptr of src/rdf.c, compiled in place)."""
        ms.control.fill(self.control)
        if rdf is not None:
            c = self.control
            c.rdf_interval, c.begin_rdf, c.istep = 1, 0, 0
            c.limit, c.nbins = float(rdf[0]), int(rdf[1])
        sysm, spec, pot = ms.cstructs()
        n = ms.nsites
        nsa = abi.nsarray(n)
        site = np.ascontiguousarray(ms.make_sites(wrap=not ms.control.molpbc) if sites is None else sites)
        force = np.zeros((3, nsa))
        chg = ms.charges()
        pe = (C.c_double * 2)(0.0, 0.0)
        stress = np.zeros((3, 3))
        rows = (C.POINTER(C.c_double) * 3)(*[C.cast(site.ctypes.data + 8 * nsa * i, C.POINTER(C.c_double)) for i in range(3)])
        frows = (C.POINTER(C.c_double) * 3)(*[C.cast(force.ctypes.data + 8 * nsa * i, C.POINTER(C.c_double)) for i in range(3)])
        pchg = chg.ctypes.data_as(C.POINTER(C.c_double))
        pstress = stress.ctypes.data_as(C.POINTER(abi.vec_mt))
        if rdf is not None:
            self.lib.init_rdf(C.byref(sysm))
        if real:
            self.lib.force_calc(rows, frows, C.byref(sysm), spec, pchg, pot, pe, pstress)
        if recip and ms.control.alpha > 1e-7:
            self.lib.ewald(rows, frows, C.byref(sysm), spec, pchg, C.byref(pe, 8), pstress)
        out = dict(force=force[:, :n].copy(), pe=np.array([pe[0], pe[1]]), stress=stress.copy(),
                   log=self.log())
        if rdf is not None:
            size = C.c_int(0)
            self.lib.rdf_ptr.restype = C.POINTER(C.c_float)
            base = self.lib.rdf_ptr(C.byref(size))
            out["rdf"] = np.ctypeslib.as_array(base, shape=(size.value,)).copy().reshape(-1, int(rdf[1]))
        return out

    def eval_forces(self, ms):
        """The reference's own eval_forces() on the configuration (needs RefLib(evalf=True)): pe[2], dip_mom[3],
        stress[3,3], molecular force[nmols,3], torque[nmols_r,3]."""
        ms.control.fill(self.control)
        args, out = ms.eval_forces_args()
        self.lib.eval_forces.restype = None
        self.lib.eval_forces(*args)
        out["log"] = self.log()
        return out

    def cell_ids(self, ms, sites=None) -> np.ndarray:
        """Per-site link-cell index through the reference's own cellbin()
        (src/force.c:119-137), with hinv by its own invert() and the product
        order of mat_vec_mul (src/matrix.c:76-83)."""
        h = np.ascontiguousarray(ms.h)
        hinv = np.zeros((3, 3))
        self.lib.invert(h.ctypes.data_as(C.POINTER(abi.vec_mt)), hinv.ctypes.data_as(C.POINTER(abi.vec_mt)))
        site = ms.make_sites(wrap=not ms.control.molpbc) if sites is None else sites
        sub = ms.control.subcell if ms.control.subcell > 0 else ms.control.cutoff / 5.0
        nx = int(h[0, 0] / sub + 0.5)
        ny = int(h[1, 1] / sub + 0.5)
        nz = int(h[2, 2] / sub + 0.5)
        eps = 8.0 * 2.0 ** -52
        n = ms.nsites
        x, y, z = site[0, :n], site[1, :n], site[2, :n]
        s = [(hinv[i, 0] * x + hinv[i, 1] * y) + hinv[i, 2] * z for i in range(3)]
        if ms.control.molpbc:                     # LOCATE(c_of_m[imol]) for non-framework molecules
            nxf, mol = ms.nsites_xf, ms.molmap()
            for i in range(3):
                s[i] = s[i].copy()
                s[i][:nxf] = ms.c_of_m[mol[:nxf], i]
        out = np.empty(n, dtype=np.int32)
        cb = self.lib.cellbin
        for i in range(n):
            out[i] = cb(s[2][i], nz, float(nz), eps) + nz * (cb(s[1][i], ny, float(ny), eps)
                                                               + ny * cb(s[0][i], nx, float(nx), eps))
        return out
