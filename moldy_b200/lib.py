"""ctypes binding of libmoldy_b200.so and the Python mirror of the reference's
operator interface for the hot path.

Two call levels, as in include/moldy_b200.h:

* `force_calc(...)` / `ewald(...)` / `kernel(...)` / `poteval(...)` / `dist_pot(...)`
  take the same arguments, in the same order and with the same accumulate-into
  semantics as the reference functions (src/force.c:1108, src/ewald.c:280,
  src/kernel.c:157,103, src/force.c:632), with numpy arrays standing in for the
  C arrays.  `eval_forces()` sequences them the way src/accel.c:488-535 does.
* `Engine` wraps the device-resident mdb_* layer for callers that keep data in HBM.

There is no fallback: importing works anywhere, but every compute call needs
the CUDA library and a GPU and raises/aborts loudly otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi
from .systems import MoldySystem

_LIB = None
LIB_PATH = os.environ.get("MOLDY_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmoldy_b200.so")

DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)


class mdb_config(C.Structure):
    _fields_ = [
        ("nsites", C.c_int), ("nsites_xf", C.c_int), ("max_id", C.c_int), ("ptype", C.c_int),
        ("n_potpar", C.c_int),
        ("site_type", IP), ("site_mol", IP), ("chg", DP), ("potpar", DP),
        ("h", C.c_double * 9),
        ("cutoff", C.c_double), ("subcell", C.c_double), ("alpha", C.c_double), ("k_cutoff", C.c_double),
        ("strict_cutoff", C.c_int), ("do_recip", C.c_int), ("molpbc", C.c_int), ("nmols", C.c_int),
    ]


def load() -> C.CDLL:
    """Load the CUDA library; fail loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m moldy_b200.build` "
            "(nvcc, sm_100a).  moldy_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.mdb_last_error.restype = C.c_char_p
    L.mdb_create.restype = C.c_void_p
    L.mdb_create.argtypes = [C.c_int]
    L.mdb_destroy.argtypes = [C.c_void_p]
    L.mdb_out_doubles.restype = C.c_size_t
    L.mdb_out_doubles.argtypes = [C.c_int]
    L.mdb_configure.argtypes = [C.c_void_p, C.POINTER(mdb_config)]
    L.mdb_set_partition.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.mdb_set_pair_mode.argtypes = [C.c_void_p, C.c_int]
    L.mdb_set_sites_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mdb_set_sites_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mdb_set_com_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    for f in ("mdb_zero_out", "mdb_force_real", "mdb_force_recip", "mdb_force_both"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mdb_build_cells.argtypes = [C.c_void_p, C.c_void_p]
    L.mdb_recip_sum_doubles.restype = C.c_size_t
    L.mdb_recip_sum_doubles.argtypes = [C.c_void_p]
    L.mdb_recip_partial.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mdb_recip_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mdb_read_out.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mdb_grid.argtypes = [C.c_void_p, IP]
    L.mdb_n_neighbour_cells.argtypes = [C.c_void_p]
    L.mdb_n_kvectors.argtypes = [C.c_void_p]
    L.mdb_pair_split.argtypes = [C.c_void_p]
    L.mdb_make_sites.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.mdb_mol_forces.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mdb_get_sites.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mdb_get_cell_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mdb_rdf_size.restype = C.c_size_t
    L.mdb_rdf_size.argtypes = [C.c_void_p, C.c_int]
    L.mdb_rdf_counts.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    L.mdb_pair_count.restype = C.c_double
    L.mdb_pair_count.argtypes = [C.c_void_p, C.c_void_p]
    L.mdb_kernel_launches.restype = C.c_long
    L.mdb_kernel_launches.argtypes = [C.c_void_p]
    L.mdb_too_close.argtypes = [C.c_void_p, IP, C.c_void_p]
    L.mdb_sizeof.restype = C.c_size_t
    L.mdb_sizeof.argtypes = [C.c_char_p]
    L.mdb_fp64_peak_probe.restype = C.c_double
    L.mdb_fp64_peak_probe.argtypes = [C.c_int, C.c_int]
    L.mdb_dmma_peak_probe.restype = C.c_double
    L.mdb_dmma_peak_probe.argtypes = [C.c_int, C.c_int]
    L.mdb_recip_gemm_flop.restype = C.c_double
    L.mdb_recip_gemm_flop.argtypes = [C.c_void_p]
    L.mdb_control.restype = C.POINTER(abi.contr_mt)
    L.mdb_abi_engine.restype = C.c_void_p
    L.mdb_abi_stream.restype = C.c_void_p
    L.mdb_abi_constants.argtypes = [DP]
    L.poteval.restype = C.c_double
    L.poteval.argtypes = [DP, C.c_double, C.c_int, C.c_double]
    L.dist_pot.restype = C.c_double
    L.dist_pot.argtypes = [DP, C.c_double, C.c_int]
    L.mdb_set_species.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.mdb_eval_result_doubles.restype = C.c_size_t
    L.mdb_eval_result_doubles.argtypes = [C.c_void_p]
    L.mdb_eval_forces_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.eval_forces.restype = None
    L.do_step.restype = None
    L.mdb_md_scalars.restype = C.c_size_t
    L.mdb_md_scalars.argtypes = [C.c_void_p]
    L.mdb_md_set_dynamics.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.mdb_md_upload_state.argtypes = [C.c_void_p] * 6
    L.mdb_md_download_state.argtypes = [C.c_void_p] * 8
    L.mdb_md_step.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.mdb_md_coords.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
    L.mdb_md_momenta.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    L.mdb_md_eval_forces.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.mdb_md_sums_now.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mdb_sites_differ_host.restype = C.c_long
    L.mdb_sites_differ_host.argtypes = [C.c_void_p] * 6
    # multi-GPU peer layer (mdb_peer.cu)
    L.mdb_peer_create.restype = C.c_void_p
    L.mdb_peer_create.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.mdb_peer_destroy.argtypes = [C.c_void_p]
    L.mdb_peer_window_bytes.restype = C.c_size_t
    L.mdb_peer_window_bytes.argtypes = [C.c_void_p]
    L.mdb_peer_handle.argtypes = [C.c_void_p, C.c_void_p]
    L.mdb_peer_open.argtypes = [C.c_void_p, C.c_void_p]
    L.mdb_peer_connect.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.mdb_peer_set_site_bounds.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    L.mdb_peer_slice.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    for f in ("mdb_peer_barrier", "mdb_peer_error", "mdb_peer_sites_gather", "mdb_peer_phase_c", "mdb_peer_phase_d"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
    for f in ("mdb_peer_phase_a", "mdb_peer_phase_b"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.mdb_peer_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    for f in ("mdb_peer_sites_host_slice", "mdb_peer_sites_host_all"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for f in ("mdb_peer_sites", "mdb_peer_in", "mdb_peer_result", "mdb_peer_partial"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_void_p]
    L.mdb_peer_in_host_slice.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.mdb_peer_in_gather.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.mdb_peer_read_slice_host.argtypes = [C.c_void_p] * 6
    L.mdb_peer_barriers.restype = C.c_long
    L.mdb_peer_barriers.argtypes = [C.c_void_p]
    L.kernel.argtypes = [C.c_int, C.c_int, DP, DP, DP, DP, C.c_double, C.c_double, C.c_double, C.c_int,
                         C.POINTER(DP)]
    _LIB = L
    return L


def _err(L):
    return L.mdb_last_error().decode()


# ---------------------------------------------------------------------------
# Moldy-level interface (same names / argument meaning as the reference)
# ---------------------------------------------------------------------------
def control() -> abi.contr_mt:
    """The library's view of Moldy's global `control` record."""
    return load().mdb_control().contents


def set_thread(ithread: int, nthreads: int):
    load().mdb_set_thread(ithread, nthreads)


def _rows(block: np.ndarray):
    assert block.dtype == np.float64 and block.flags.c_contiguous and block.shape[0] == 3
    stride = block.strides[0]
    return (DP * 3)(*[C.cast(block.ctypes.data + stride * i, DP) for i in range(3)])


def force_calc(site, site_force, system, species, chg, potpar, pe, stress):
    """force_calc(site, site_force, system, species, chg, potpar, pe, stress):
    real-space forces accumulated into site_force[3][nsarray], pe[0], stress."""
    L = load()
    L.force_calc(_rows(site), _rows(site_force), C.byref(system), species,
                 chg.ctypes.data_as(DP), potpar, pe.ctypes.data_as(DP),
                 stress.ctypes.data_as(C.POINTER(abi.vec_mt)))


def ewald(site, site_force, system, species, chg, pe, stress):
    """ewald(site, site_force, system, species, chg, pe, stress): reciprocal-space
    part; `pe` is the caller's pe+1 (a length-1 view)."""
    L = load()
    L.ewald(_rows(site), _rows(site_force), C.byref(system), species, chg.ctypes.data_as(DP),
            pe.ctypes.data_as(DP), stress.ctypes.data_as(C.POINTER(abi.vec_mt)))


def kernel(jmin, nnab, forceij, pe, r_sqr, nab_chg, chg, norm, alpha, ptype, pot):
    """kernel(): vector pair-potential evaluation; pot is [n_potpar][nnab]."""
    L = load()
    pot = np.ascontiguousarray(pot, dtype=np.float64)
    rows = (DP * 8)(*[C.cast(pot.ctypes.data + pot.strides[0] * min(i, pot.shape[0] - 1), DP) for i in range(8)])
    L.kernel(jmin, nnab, forceij.ctypes.data_as(DP), pe.ctypes.data_as(DP), r_sqr.ctypes.data_as(DP),
             nab_chg.ctypes.data_as(DP), chg, norm, alpha, ptype, rows)


def poteval(potpar, r, ptype, chgsq):
    p = np.ascontiguousarray(potpar, dtype=np.float64)
    return load().poteval(p.ctypes.data_as(DP), r, ptype, chgsq)


def dist_pot(potpar, cutoff, ptype):
    p = np.ascontiguousarray(potpar, dtype=np.float64)
    return load().dist_pot(p.ctypes.data_as(DP), cutoff, ptype)


def eval_forces(ms: MoldySystem, real=True, recip=True, sites=None, ithread=0, nthreads=1, rdf=None):
    """The hot-path part of eval_forces() (src/accel.c:488-535) for one
    configuration through the Moldy-level C ABI with HOST buffers.
    rdf=(limit, nbins) switches on force_calc's RDF pass for this call (control.rdf_interval = 1) and
    returns the float histograms the library accumulated (what Moldy's rdf_ptr() store would hold)."""
    L = load()
    ms.control.fill(control())
    if rdf is not None:
        c = control()
        c.rdf_interval, c.begin_rdf, c.istep = 1, 0, 0
        c.limit, c.nbins = float(rdf[0]), int(rdf[1])
        mid = ms.sysdef.max_id
        L.mdb_rdf_private_resize(c.nbins * mid * (mid - 1) // 2)
    set_thread(ithread, nthreads)
    sysm, spec, pot = ms.cstructs()
    n = ms.nsites
    nsa = abi.nsarray(n)
    # eval_forces() builds the sites with MOLPBC (no per-site wrap) when control.molpbc (src/accel.c:500-504)
    site = np.ascontiguousarray(ms.make_sites(wrap=not ms.control.molpbc) if sites is None else sites)
    force = np.zeros((3, nsa))
    chg = ms.charges()
    pe = np.zeros(2)
    stress = np.zeros((3, 3))
    if real:
        force_calc(site, force, sysm, spec, chg, pot, pe[0:1], stress)
    if recip and ms.control.alpha > 1e-7:
        ewald(site, force, sysm, spec, chg, pe[1:2], stress)
    out = dict(force=force[:, :n].copy(), pe=pe, stress=stress)
    if rdf is not None:
        size = C.c_int(0)
        L.rdf_ptr.restype = C.POINTER(C.c_float)
        base = L.rdf_ptr(C.byref(size))
        out["rdf"] = np.ctypeslib.as_array(base, shape=(size.value,)).copy().reshape(-1, int(rdf[1]))
    return out


def eval_forces_mol(ms: MoldySystem):
    """Moldy's eval_forces() itself (src/accel.c:398-617) through the library's drop-in symbol: scaled centres of mass
    and quaternions in; pe[2], dip_mom[3], the full virial stress[3,3], molecular force[nmols,3] and
    torque[nmols_r,3] out.  Sites and site forces stay in HBM."""
    L = load()
    ms.control.fill(control())
    set_thread(0, 1)
    args, out = ms.eval_forces_args()
    L.eval_forces(*args)
    return out


def do_step(ms: MoldySystem, mom, amom, step: float, nsteps: int = 1, resident: bool = False):
    """Moldy's do_step() (src/accel.c:626-827, NVE) through the library's drop-in symbol, `nsteps` times: centres of mass,
    quaternions and momenta of `ms` advance in place on the device.  Returns the dict of do_step_args (state arrays, pe,
    stress, molecular forces/torques of the last step, meansq)."""
    L = load()
    args, out, set_control = ms.do_step_args(mom, amom, step)
    set_control(control())
    set_thread(0, 1)
    for k in range(nsteps):
        control().istep = 2 + k
        L.do_step(*args)
    return out


def shutdown():
    """Drop the Moldy-level engine(s); the next call re-reads MOLDY_B200_DEVICE / MOLDY_B200_DEVICES (tests)."""
    L = load()
    L.mdb_abi_reset()
    L.mdb_abi_shutdown()


def n_devices() -> int:
    """Number of GPUs (ranks) behind force_calc()/ewald()/eval_forces() (MOLDY_B200_DEVICES)."""
    return load().mdb_abi_devices()


def reset():
    """Forget the Moldy-level first-call state (lets one test process run several systems)."""
    load().mdb_abi_reset()


# ---------------------------------------------------------------------------
# Device-resident engine
# ---------------------------------------------------------------------------
class Engine:
    def __init__(self, device: int = 0):
        self.L = load()
        self.h = self.L.mdb_create(device)
        if not self.h:
            raise RuntimeError("mdb_create failed: " + _err(self.L))
        self.n = 0
        self._keep = None

    def close(self):
        if self.h:
            self.L.mdb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {_err(self.L)}")

    @staticmethod
    def make_config(ms: MoldySystem):
        """mdb_config of a system + the arrays it points into (keep them alive while it is used)."""
        cfg = mdb_config()
        sd = ms.sysdef
        ids = np.ascontiguousarray(ms.site_ids(), dtype=np.int32)
        mol = np.ascontiguousarray(ms.molmap(), dtype=np.int32)
        chg = np.ascontiguousarray(ms.charges())
        pot = np.ascontiguousarray(sd.potpar.reshape(-1), dtype=np.float64)
        cfg.nsites, cfg.nsites_xf = ms.nsites, ms.nsites_xf
        cfg.max_id, cfg.ptype, cfg.n_potpar = sd.max_id, sd.ptype, sd.n_potpar
        cfg.site_type = ids.ctypes.data_as(IP)
        cfg.site_mol = mol.ctypes.data_as(IP)
        cfg.chg = chg.ctypes.data_as(DP)
        cfg.potpar = pot.ctypes.data_as(DP)
        for i in range(9):
            cfg.h[i] = float(ms.h.reshape(-1)[i])
        c = ms.control
        cfg.cutoff, cfg.subcell, cfg.alpha, cfg.k_cutoff = c.cutoff, c.subcell, c.alpha, c.k_cutoff
        cfg.strict_cutoff = c.strict_cutoff
        cfg.do_recip = int(c.alpha > 1e-7)
        cfg.molpbc, cfg.nmols = int(c.molpbc), ms.nmols
        return cfg, (ids, mol, chg, pot)

    def configure(self, ms: MoldySystem):
        cfg, self._keep = Engine.make_config(ms)
        self._chk(self.L.mdb_configure(self.h, C.byref(cfg)), "mdb_configure")
        self.n = ms.nsites

    def out_doubles(self):
        return self.L.mdb_out_doubles(self.n)

    def set_partition(self, ithread, nthreads):
        self.L.mdb_set_partition(self.h, ithread, nthreads)

    def set_pair_mode(self, mode: int):
        """2: per-thread full stencil, 3: tiled full stencil (default), 4: tiled Newton-3."""
        self.L.mdb_set_pair_mode(self.h, mode)

    def set_sites_host(self, site_block: np.ndarray, stream=0):
        r = [site_block.ctypes.data + site_block.strides[0] * i for i in range(3)]
        self._chk(self.L.mdb_set_sites_host(self.h, r[0], r[1], r[2], stream), "mdb_set_sites_host")

    def set_com_host(self, c_of_m: np.ndarray, stream=0):
        """molecular-cutoff mode: scaled centre-of-mass co-ordinates [nmols,3]."""
        com = np.ascontiguousarray(c_of_m, dtype=np.float64)
        self._com_keep = com
        self._chk(self.L.mdb_set_com_host(self.h, com.ctypes.data, stream), "mdb_set_com_host")

    def set_sites_host_ptrs(self, px, py, pz, stream=0):
        self._chk(self.L.mdb_set_sites_host(self.h, px, py, pz, stream), "mdb_set_sites_host")

    def set_sites_device(self, px, py, pz, stream=0):
        self._chk(self.L.mdb_set_sites_device(self.h, px, py, pz, stream), "mdb_set_sites_device")

    def zero_out(self, d_out, stream=0):
        self._chk(self.L.mdb_zero_out(self.h, d_out, stream), "mdb_zero_out")

    def build_cells(self, stream=0):
        self._chk(self.L.mdb_build_cells(self.h, stream), "mdb_build_cells")

    def force_real(self, d_out, stream=0):
        self._chk(self.L.mdb_force_real(self.h, d_out, stream), "mdb_force_real")

    def force_recip(self, d_out, stream=0):
        self._chk(self.L.mdb_force_recip(self.h, d_out, stream), "mdb_force_recip")

    def force_both(self, d_out, stream=0):
        """real space + k-space of one step, the pair kernel filling the FP64 issue slots beside the k-space GEMMs."""
        self._chk(self.L.mdb_force_both(self.h, d_out, stream), "mdb_force_both")

    def set_overlap(self, fill_blocks: int = 0, fill_threads: int = 0):
        """fill_blocks < 0: one stream; 0: k-space first on a side stream, set-up behind it (default); > 0: filler grid."""
        self.L.mdb_set_overlap(self.h, int(fill_blocks), int(fill_threads))

    def set_pair_far(self, on: bool):
        """before configure(): far stencil runs of the exponential potentials without their exponentials (default on)."""
        self.L.mdb_set_pair_far.argtypes = [C.c_void_p, C.c_int]
        self.L.mdb_set_pair_far(self.h, int(on))

    def pair_far_runs(self) -> int:
        self.L.mdb_pair_far_runs.argtypes = [C.c_void_p]
        return self.L.mdb_pair_far_runs(self.h)

    def overlap_filled(self) -> int:
        self.L.mdb_overlap_filled.restype = C.c_long
        self.L.mdb_overlap_filled.argtypes = [C.c_void_p]
        return self.L.mdb_overlap_filled(self.h)

    def recip_sum_doubles(self) -> int:
        return self.L.mdb_recip_sum_doubles(self.h)

    def recip_partial(self, d_psum, stream=0):
        self._chk(self.L.mdb_recip_partial(self.h, d_psum, stream), "mdb_recip_partial")

    def recip_finish(self, d_psum, d_out, stream=0):
        self._chk(self.L.mdb_recip_finish(self.h, d_psum, d_out, stream), "mdb_recip_finish")

    def read_out(self, d_out, stream=0) -> np.ndarray:
        h = np.empty(self.out_doubles())
        self._chk(self.L.mdb_read_out(self.h, d_out, h.ctypes.data, stream), "mdb_read_out")
        return h

    def grid(self):
        g = (C.c_int * 3)()
        nc = self.L.mdb_grid(self.h, g)
        return nc, tuple(g)

    def n_neighbour_cells(self):
        return self.L.mdb_n_neighbour_cells(self.h)

    def n_kvectors(self):
        return self.L.mdb_n_kvectors(self.h)

    def cell_ids(self, stream=0) -> np.ndarray:
        out = np.empty(self.n, dtype=np.int32)
        self._chk(self.L.mdb_get_cell_ids(self.h, out.ctypes.data, stream), "mdb_get_cell_ids")
        return out

    def rdf_counts(self, limit: float, nbins: int, stream=0) -> np.ndarray:
        """Pair counts per (id pair, bin) of force_calc's RDF pass for the current sites: uint64 [pairs, nbins]."""
        n = self.L.mdb_rdf_size(self.h, nbins)
        out = np.zeros(n, dtype=np.uint64)
        if self.L.mdb_rdf_counts(self.h, limit, nbins, out.ctypes.data, stream):
            raise RuntimeError(_err(self.L))
        return out.reshape(-1, nbins)

    def pair_count(self, stream=0) -> float:
        return self.L.mdb_pair_count(self.h, stream)

    def make_sites(self, h: np.ndarray, d_com_s: int, d_quat: int, d_pfs: int, nmols: int, nsites: int, site_offset: int,
                   sitepbc: bool, stream=0):
        """make_sites() of one species on the device (device pointers; d_quat 0 for species without quaternions)."""
        hh = np.ascontiguousarray(h, dtype=np.float64)
        self._chk(self.L.mdb_make_sites(self.h, hh.ctypes.data, d_com_s, d_quat or None, d_pfs, nmols, nsites, site_offset,
                                        1 if sitepbc else 0, stream), "mdb_make_sites")

    def mol_forces(self, d_out: int, d_quat: int, d_pfs: int, nmols: int, nsites: int, site_offset: int, d_force: int,
                   d_torque: int, stream=0):
        """mol_force() + mol_torque() of one species from a result block (device pointers; d_torque 0: forces only)."""
        self._chk(self.L.mdb_mol_forces(self.h, d_out, d_quat or None, d_pfs, nmols, nsites, site_offset, d_force,
                                        d_torque or None, stream), "mdb_mol_forces")

    def get_sites(self, stream=0) -> np.ndarray:
        out = np.empty((3, self.n))
        self._chk(self.L.mdb_get_sites(self.h, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data, stream), "mdb_get_sites")
        return out

    def pair_split(self) -> int:
        """1 when the real-space sum runs as two passes by site class (charged / with a pair potential)."""
        return self.L.mdb_pair_split(self.h)

    def launches(self) -> int:
        return self.L.mdb_kernel_launches(self.h)

    def too_close(self, stream=0):
        p = (C.c_int * 2)()
        n = self.L.mdb_too_close(self.h, p, stream)
        return n, (p[0], p[1])


PEER_HANDLE_BYTES = 64
REAL, RECIP = 1, 2


class MdState:
    """The resident NVE integrator of an Engine (mdb_md.cu): per-species state arrays in, scalars out."""

    def __init__(self, eng: Engine, ms: MoldySystem, nosymmetric_rot: int = 0):
        self.eng, self.ms, self.L = eng, ms, eng.L
        sd = ms.sysdef
        nsp = len(sd.species)
        sp = (abi.mdb_species * nsp)()
        dyn = (abi.mdb_species_dyn * nsp)()
        pfs = []
        for i, (s, inert) in enumerate(zip(sd.species, ms.principal_inertia())):
            sp[i].nmols, sp[i].nsites, sp[i].framework = s.nmols, s.nsites, int(s.framework)
            sp[i].rotates, sp[i].rdof = int(s.rdof > 0), s.rdof
            dyn[i].mass = s.mass
            for k in range(3):
                dyn[i].inertia[k] = float(inert[k])
            pfs.append(np.asarray(s.p_f_sites, dtype=np.float64).reshape(-1, 3))
        pfs = np.ascontiguousarray(np.concatenate(pfs))
        eng._chk(self.L.mdb_set_species(eng.h, nsp, sp, pfs.ctypes.data), "mdb_set_species")
        eng._chk(self.L.mdb_md_set_dynamics(eng.h, dyn, nosymmetric_rot), "mdb_md_set_dynamics")
        self.nsp = nsp
        self.nscal = self.L.mdb_md_scalars(eng.h)

    def _ptrs(self, arr, width, only_rot=False):
        out = (C.c_void_p * self.nsp)()
        m0 = 0
        for i, s in enumerate(self.ms.sysdef.species):
            if arr is not None and not (only_rot and not s.rdof):
                out[i] = arr.ctypes.data + 8 * width * m0
            m0 += s.nmols
        return out

    def upload(self, com, quat, mom, amom, stream=0):
        self._keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (com, quat, mom, amom)]
        c, q, m, a = self._keep
        self.eng._chk(self.L.mdb_md_upload_state(self.eng.h, self._ptrs(c, 3), self._ptrs(q, 4, True), self._ptrs(m, 3),
                                                 self._ptrs(a, 4, True), stream), "mdb_md_upload_state")

    def download(self, stream=0):
        nm = self.ms.nmols
        com, quat, mom, amom = np.zeros((nm, 3)), np.zeros((nm, 4)), np.zeros((nm, 3)), np.zeros((nm, 4))
        force, torque = np.zeros((nm, 3)), np.zeros((nm, 3))
        self.eng._chk(self.L.mdb_md_download_state(self.eng.h, self._ptrs(com, 3), self._ptrs(quat, 4, True), self._ptrs(mom, 3),
                                                   self._ptrs(amom, 4, True), self._ptrs(force, 3), self._ptrs(torque, 3, True),
                                                   stream), "mdb_md_download_state")
        return dict(com=com, quat=quat, mom=mom, amom=amom, force=force, torque=torque)

    def step(self, step, ts=1.0, surface_dipole=None, half_sums=False, stream=0):
        h = np.ascontiguousarray(self.ms.h, dtype=np.float64)
        sd = self.ms.control.surface_dipole if surface_dipole is None else surface_dipole
        out = np.zeros(self.nscal)
        self.eng._chk(self.L.mdb_md_step(self.eng.h, h.ctypes.data, step, ts, int(sd), int(self.ms.control.alpha > 1e-7),
                                         int(half_sums), out.ctypes.data, stream), "mdb_md_step")
        return out

    def sums_now(self, stream=0):
        h = np.ascontiguousarray(self.ms.h, dtype=np.float64)
        out = np.zeros((self.nsp, 15))
        self.eng._chk(self.L.mdb_md_sums_now(self.eng.h, h.ctypes.data, out.ctypes.data, stream), "mdb_md_sums_now")
        return out


class GroupMd:
    """The resident NVE step on a device group (mdb_group.cu: mdb_group_md_step): one process, several GPUs, every rank
    moving its share of the molecules.  devices may repeat (ranks sharing a GPU)."""

    def __init__(self, ms: MoldySystem, devices, nosymmetric_rot: int = 0):
        self.L, self.ms = load(), ms
        L = self.L
        L.mdb_group_create.restype = C.c_void_p
        L.mdb_group_create.argtypes = [C.c_int, IP]
        for f in ("mdb_group_destroy", "mdb_group_size"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.mdb_group_configure.argtypes = [C.c_void_p, C.POINTER(mdb_config)]
        L.mdb_group_set_species.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mdb_group_md_set_dynamics.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.mdb_group_md_upload_state.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.mdb_group_md_download_state.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.mdb_group_md_scalars.restype = C.c_size_t
        L.mdb_group_md_scalars.argtypes = [C.c_void_p]
        L.mdb_group_md_step.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_double, C.c_int, C.c_void_p]
        dev = (C.c_int * len(devices))(*devices)
        self.g = L.mdb_group_create(len(devices), dev)
        if not self.g:
            raise RuntimeError("mdb_group_create: " + _err(L))
        cfg, self._keep = Engine.make_config(ms)
        self._chk(L.mdb_group_configure(self.g, C.byref(cfg)), "mdb_group_configure")
        sd = ms.sysdef
        nsp = len(sd.species)
        sp = (abi.mdb_species * nsp)()
        dyn = (abi.mdb_species_dyn * nsp)()
        pfs = []
        for i, (s, inert) in enumerate(zip(sd.species, ms.principal_inertia())):
            sp[i].nmols, sp[i].nsites, sp[i].framework = s.nmols, s.nsites, int(s.framework)
            sp[i].rotates, sp[i].rdof = int(s.rdof > 0), s.rdof
            dyn[i].mass = s.mass
            for k in range(3):
                dyn[i].inertia[k] = float(inert[k])
            pfs.append(np.asarray(s.p_f_sites, dtype=np.float64).reshape(-1, 3))
        pfs = np.ascontiguousarray(np.concatenate(pfs))
        self._chk(L.mdb_group_set_species(self.g, nsp, sp, pfs.ctypes.data), "mdb_group_set_species")
        self._chk(L.mdb_group_md_set_dynamics(self.g, dyn, nosymmetric_rot), "mdb_group_md_set_dynamics")
        self.nsp, self.nscal = nsp, L.mdb_group_md_scalars(self.g)

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {_err(self.L)}")

    _ptrs = MdState._ptrs

    def upload(self, com, quat, mom, amom):
        self._state = [np.ascontiguousarray(a, dtype=np.float64) for a in (com, quat, mom, amom)]
        c, q, m, a = self._state
        self._chk(self.L.mdb_group_md_upload_state(self.g, self._ptrs(c, 3), self._ptrs(q, 4, True), self._ptrs(m, 3),
                                                   self._ptrs(a, 4, True)), "mdb_group_md_upload_state")

    def download(self):
        nm = self.ms.nmols
        com, quat, mom, amom = np.zeros((nm, 3)), np.zeros((nm, 4)), np.zeros((nm, 3)), np.zeros((nm, 4))
        force, torque = np.zeros((nm, 3)), np.zeros((nm, 3))
        self._chk(self.L.mdb_group_md_download_state(self.g, self._ptrs(com, 3), self._ptrs(quat, 4, True), self._ptrs(mom, 3),
                                                     self._ptrs(amom, 4, True), self._ptrs(force, 3), self._ptrs(torque, 3, True)),
                  "mdb_group_md_download_state")
        return dict(com=com, quat=quat, mom=mom, amom=amom, force=force, torque=torque)

    def step(self, step, ts=1.0, half_sums=False):
        h = np.ascontiguousarray(self.ms.h, dtype=np.float64)
        out = np.zeros(self.nscal)
        self._chk(self.L.mdb_group_md_step(self.g, h.ctypes.data, step, ts, int(self.ms.control.surface_dipole),
                                           int(self.ms.control.alpha > 1e-7), int(half_sums), out.ctypes.data, 0.0, 0, None),
                  "mdb_group_md_step")
        return out

    def close(self):
        if self.g:
            self.L.mdb_group_destroy(self.g)
            self.g = None


class Peer:
    """One rank of the replicated-data multi-GPU layer (mdb_peer.cu): an Engine plus its peer-mapped window."""

    def __init__(self, eng: Engine, rank: int, world: int):
        self.L, self.eng, self.rank, self.world = eng.L, eng, rank, world
        self.h = self.L.mdb_peer_create(eng.h, rank, world)
        if not self.h:
            raise RuntimeError("mdb_peer_create failed: " + _err(self.L))

    def close(self):
        if self.h:
            self.L.mdb_peer_destroy(self.h)
            self.h = None

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {_err(self.L)}")

    def handle(self) -> bytes:
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        self._chk(self.L.mdb_peer_handle(self.h, buf), "mdb_peer_handle")
        return buf.raw

    def open(self, handles):
        """handles: the `world` 64-byte IPC handles in rank order (one process per rank)."""
        blob = b"".join(handles)
        assert len(blob) == PEER_HANDLE_BYTES * self.world
        self._chk(self.L.mdb_peer_open(self.h, blob), "mdb_peer_open")

    @staticmethod
    def connect(peers):
        """All ranks are engines of this process: map the windows directly."""
        L = peers[0].L
        arr = (C.c_void_p * len(peers))(*[p.h for p in peers])
        if L.mdb_peer_connect(arr, len(peers)):
            raise RuntimeError("mdb_peer_connect: " + _err(L))

    def set_site_bounds(self, bounds):
        b = (C.c_longlong * (self.world + 1))(*[int(x) for x in bounds])
        self._chk(self.L.mdb_peer_set_site_bounds(self.h, b), "mdb_peer_set_site_bounds")

    def slice(self):
        lohi = (C.c_longlong * 2)()
        self.L.mdb_peer_slice(self.h, lohi)
        return int(lohi[0]), int(lohi[1])

    def barrier(self, stream=0):
        self._chk(self.L.mdb_peer_barrier(self.h, stream), "mdb_peer_barrier")

    def error(self, stream=0) -> int:
        return self.L.mdb_peer_error(self.h, stream)

    def sites_host_all(self, site_block: np.ndarray, stream=0):
        r = [site_block.ctypes.data + site_block.strides[0] * i for i in range(3)]
        self._chk(self.L.mdb_peer_sites_host_all(self.h, r[0], r[1], r[2], stream), "mdb_peer_sites_host_all")

    def sites_host_slice(self, px, py, pz, stream=0):
        self._chk(self.L.mdb_peer_sites_host_slice(self.h, px, py, pz, stream), "mdb_peer_sites_host_slice")

    def sites_gather(self, stream=0):
        self._chk(self.L.mdb_peer_sites_gather(self.h, stream), "mdb_peer_sites_gather")

    def phase_a(self, what=REAL | RECIP, stream=0):
        self._chk(self.L.mdb_peer_phase_a(self.h, what, stream), "mdb_peer_phase_a")

    def phase_b(self, what=REAL | RECIP, stream=0):
        self._chk(self.L.mdb_peer_phase_b(self.h, what, stream), "mdb_peer_phase_b")

    def phase_c(self, stream=0):
        self._chk(self.L.mdb_peer_phase_c(self.h, stream), "mdb_peer_phase_c")

    def phase_d(self, stream=0):
        self._chk(self.L.mdb_peer_phase_d(self.h, stream), "mdb_peer_phase_d")

    def step(self, what=REAL | RECIP, gather=False, stream=0):
        self._chk(self.L.mdb_peer_step(self.h, what, 1 if gather else 0, stream), "mdb_peer_step")

    def result_ptr(self) -> int:
        return self.L.mdb_peer_result(self.h)

    def read_slice_host(self, pfx, pfy, pfz, pscal, stream=0):
        self._chk(self.L.mdb_peer_read_slice_host(self.h, pfx, pfy, pfz, pscal, stream), "mdb_peer_read_slice_host")

    def barriers(self) -> int:
        return self.L.mdb_peer_barriers(self.h)


def unpack(h_out: np.ndarray, n: int):
    """Split a result block into forces[3,N], pe[2], stress[3,3]."""
    f = h_out[:3 * n].reshape(3, n).copy()
    pe = h_out[3 * n:3 * n + 2].copy()
    stress = h_out[3 * n + 2:3 * n + 11].reshape(3, 3).copy()
    return f, pe, stress
