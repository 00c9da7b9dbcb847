"""Synthetic Moldy systems for the hot path: sys-spec parsing, unit conversion,
Ewald auto-parameters, configuration generation and replication.

This is the *input side* of force_calc()/ewald(): it produces exactly what
Moldy's start-up hands to eval_forces() -- `system_mt`, `spec_mt[]`,
`pot_mt[max_id**2]`, the `site[3][nsarray]` block and `chg[]` -- in program
units (amu, Angstrom, ps; charge unit such that U = q_i q_j / r).

Reference behaviour followed (not code): sys-spec file grammar
src/input.c:128-366; unit conversion src/convert.c:61-108 with the constants of
src/defs.h:188-227; Ewald auto-parameters src/startup.c:713-772; box from
density src/startup.c (cubic, V = M / rho); lattice cell matrix
src/input.c:405-413; site generation and per-site periodic wrap
src/algorith.c:169-217.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import abi

# ---- constants (CODATA 1986 set, src/defs.h:200-211) -----------------------
AMU = 1.6605402e-27
ELCHG = 1.60217733e-19
ROOT_4_PI_EPS = 1.05482230112e-05
RTAMU = 4.07497263794495e-14
AVOGAD = 6.0221367e23
PROG_UNIT = dict(m=AMU, l=1.0e-10, t=1.0e-12, q=RTAMU * 1.0e-3 * ROOT_4_PI_EPS)
EUNIT = PROG_UNIT["m"] * (PROG_UNIT["l"] / PROG_UNIT["t"]) ** 2
CONV_E = 0.001 * AVOGAD * EUNIT                       # program energy -> kJ/mol
CONV_Q = PROG_UNIT["q"] / ELCHG
KCAL_TIME_UNIT = 4.8888213e-14                         # kcal/mol, amu, A
EV_TIME_UNIT = 1.0181e-14                              # eV, amu, A

# name, number of parameters (src/kernel.c:61-68)
POTSPEC = [("lennard-jones", 2), ("buckingham", 3), ("mcy", 4), ("generic", 6),
           ("hiw", 3), ("reserved for developer", 1), ("morse", 7)]
# powers of (m, l, t) per parameter (src/kernel.c:75-84)
POT_DIM = [
    [(1, 2, -2), (0, 1, 0)],
    [(1, 8, -2), (1, 2, -2), (0, -1, 0)],
    [(1, 2, -2), (0, -1, 0), (1, 2, -2), (0, -1, 0)],
    [(1, 2, -2), (0, -1, 0), (1, 14, -2), (1, 6, -2), (1, 8, -2), (1, 10, -2)],
    [(1, 6, -2), (1, 8, -2), (1, 14, -2)],
    [(0, 0, 0)],
    [(1, 2, -2), (0, 1, 0), (0, -1, 0), (1, 8, -2), (1, 2, -2), (0, -1, 0), (0, 1, 0)],
]


def unit_scale(dim, ufrom, uto):
    """exp(sum dim*(ln from - ln to)), as src/convert.c:68-80."""
    m, l, t = dim[:3]
    q = dim[3] if len(dim) > 3 else 0
    ln = (m * (math.log(ufrom["m"]) - math.log(uto["m"]))
          + l * (math.log(ufrom["l"]) - math.log(uto["l"]))
          + t * (math.log(ufrom["t"]) - math.log(uto["t"]))
          + q * (math.log(ufrom["q"]) - math.log(uto["q"])))
    return math.exp(ln)


# ---- system specifications (our own data tables; parameters are the published
# ---- model constants also used by the reference's example inputs) -----------
SPEC_ARGON = """
Argon 108
1 0 0 0 39.948 0 Ar
end
lennard-jones
1 1 3.984 3.41
end
"""

SPEC_TIP4P = """
Water 256
1 0 0 0 16 0 O
2 0.7569503 0 -0.5858822 1 0.52 H
2 -0.7569503 0 -0.5858822
3 0 0 -0.15 0 -1.04 M
end
lennard-jones
1 1 0.6201667 3.1536
end
"""

SPEC_TIPS2 = """
Water 64
1 0 0 0 16 0 O
2 0.7569503 0 -0.5858822 1 0.535 H
2 -0.7569503 0 -0.5858822
3 0 0 -0.15 0 -1.07 M
end
lennard-jones
1 1 0.51799 3.2407
end
"""

SPEC_MGCL2 = """
Water 200
1 0 0 0 16 0 O
2 0.7569503 0 -0.5858822 1 0.717484 H
2 -0.7569503 0 -0.5858822
3 0 0 -0.2677 0 -1.434968 M
Magnesium 4
4 0 0 0 24.31 2 Mg2+
Chloride 8
5 0 0 0 35.45 -1 Cl-
end
mcy
1 1 1088213.2 5.152712 0 0
1 2 1455.427 2.961895 273.5954 2.233264
2 2 666.3373 2.760844 0 0
1 4 47750.0 3.836 546.3 1.253
2 4 111.0 1.06 0 1.0
1 5 198855.0 3.910 0 0
2 5 1857.0 2.408 77.94 1.369
4 5 28325.5 2.65 0 0
end
"""

SPEC_QUARTZ = """
Oxygen 384
1 0 0 0 16 -1.2 O
Silicon 192
2 0 0 0 28.0855 2.4 Si
end
buckingham
1 1 175.0000 1388.7730 2.76000
1 2 133.5381 18003.7572 4.87318
2 2 0.0 0.0 0.0
end
"""
# alpha-quartz unit cell (a, b, c, alpha, beta, gamma) and fractional basis
QUARTZ_CELL = (4.903, 4.903, 5.393, 90.0, 90.0, 120.0)
QUARTZ_BASIS = [
    ("Oxygen", 0.415, 0.272, 0.120), ("Oxygen", 0.857, 0.585, 0.4533),
    ("Oxygen", 0.728, 0.143, 0.4533), ("Oxygen", 0.143, 0.728, 0.880),
    ("Oxygen", 0.272, 0.415, 0.5467), ("Oxygen", 0.585, 0.857, 0.2133),
    ("Silicon", 0.465, 0.0, 0.0), ("Silicon", 0.535, 0.535, 0.3333),
    ("Silicon", 0.0, 0.465, 0.6667),
]

# A small artificial framework system (exercises frame_type / nfnab / the split
# k-space sums): TIP4P-like water between two charged rigid sheets.
SPEC_SLAB = """
Water 48
1 0 0 0 16 0 O
2 0.7569503 0 -0.5858822 1 0.52 H
2 -0.7569503 0 -0.5858822
3 0 0 -0.15 0 -1.04 M
Cation 6
4 0 0 0 22.99 1 Na
Sheet 1 framework
%s
end
generic
1 1 0 1 600000 0 610 0
1 4 60000 3.2 0 0 0 0
1 5 90000 3.4 0 0 30 0
4 4 50000 3.0 0 0 0 0
4 5 70000 3.1 0 0 10 0
end
"""


@dataclass
class Species:
    name: str
    nmols: int
    framework: bool
    site_id: np.ndarray            # [nsites] int32
    p_f_sites: np.ndarray          # [nsites,3] relative to centre of mass
    mass: float = 0.0
    charge: float = 0.0

    @property
    def nsites(self):
        return len(self.site_id)

    @property
    def rdof(self):
        return 0 if self.nsites == 1 else 3


@dataclass
class SysDef:
    species: list
    site_mass: np.ndarray          # [max_id]
    site_charge: np.ndarray        # [max_id] program units
    site_name: list
    ptype: int
    potpar: np.ndarray             # [max_id, max_id, NPOTP] program units

    @property
    def max_id(self):
        return len(self.site_mass)

    @property
    def n_potpar(self):
        return POTSPEC[self.ptype][1]


def parse_sysdef(text: str, time_unit: float = 1.0e-13, mass_unit: float = AMU,
                 length_unit: float = 1.0e-10, charge_unit: float = ELCHG) -> SysDef:
    """Parse a Moldy system-specification text and convert to program units."""
    lines = [ln.split("#")[0].strip() for ln in text.strip().splitlines()]
    lines = [ln for ln in lines if ln]
    it = iter(lines)
    species, raw = [], []
    info = {}                       # id -> [mass, charge, name]
    cur = None
    for ln in it:
        tok = ln.split()
        if tok[0].lower() == "end":
            break
        if not _is_int(tok[0]):
            cur = dict(name=tok[0], nmols=int(tok[1]),
                       framework=len(tok) > 2 and tok[2].lower() == "framework",
                       ids=[], xyz=[])
            raw.append(cur)
            continue
        sid = int(tok[0])
        cur["ids"].append(sid)
        cur["xyz"].append([float(t) for t in tok[1:4]])
        if len(tok) > 4:
            mass = float(tok[4])
            chg = float(tok[5]) if len(tok) > 5 else 0.0
            name = tok[6] if len(tok) > 6 else ""
            info[sid] = [mass, chg, name]
        elif sid not in info:
            raise ValueError(f"site id {sid} used before its mass/charge were given")
    max_id = max(info) + 1
    pname = next(it).lower().replace("potential parameters", "").strip()
    ptype = [p[0] for p in POTSPEC].index(pname)
    npar = POTSPEC[ptype][1]
    potpar = np.zeros((max_id, max_id, abi.NPOTP))
    for ln in it:
        tok = ln.split()
        if tok[0].lower() == "end":
            break
        i, j = int(tok[0]), int(tok[1])
        p = [float(t) for t in tok[2:2 + npar]]
        potpar[i, j, :len(p)] = p
        potpar[j, i, :len(p)] = p
    ufrom = dict(m=mass_unit, l=length_unit, t=time_unit, q=charge_unit)
    for ip in range(npar):
        potpar[:, :, ip] *= unit_scale(POT_DIM[ptype][ip], ufrom, PROG_UNIT)
    mscale = unit_scale((1, 0, 0, 0), ufrom, PROG_UNIT)
    qscale = unit_scale((0, 0, 0, 1), ufrom, PROG_UNIT)
    lscale = unit_scale((0, 1, 0, 0), ufrom, PROG_UNIT)
    site_mass = np.zeros(max_id)
    site_charge = np.zeros(max_id)
    site_name = [""] * max_id
    for sid, (m, q, nm) in info.items():
        site_mass[sid] = m * mscale
        site_charge[sid] = q * qscale
        site_name[sid] = nm
    for r in raw:
        ids = np.asarray(r["ids"], dtype=np.int32)
        xyz = np.asarray(r["xyz"], dtype=np.float64) * lscale
        m = site_mass[ids]
        mtot = float(m.sum())
        if mtot > 0:
            xyz = xyz - (m[:, None] * xyz).sum(0) / mtot
        species.append(Species(r["name"], r["nmols"], r["framework"], ids, xyz,
                               mass=mtot, charge=float(site_charge[ids].sum())))
    # frameworks last (src/input.c:287)
    species.sort(key=lambda s: s.framework)
    return SysDef(species, site_mass, site_charge, site_name, ptype, potpar)


def _is_int(s):
    try:
        int(s)
        return True
    except ValueError:
        return False


@dataclass
class Control:
    """The subset of contr_mt the hot path reads."""
    cutoff: float = 0.0
    alpha: float = 0.0
    k_cutoff: float = 0.0
    subcell: float = 0.0
    strict_cutoff: int = 0
    surface_dipole: int = 0
    molpbc: int = 0
    ewald_accuracy: float = 1.013e-5
    density: float = 1.0            # g/cm^3, only used to size the box

    def fill(self, c: abi.contr_mt):
        c.cutoff, c.alpha, c.k_cutoff = self.cutoff, self.alpha, self.k_cutoff
        c.subcell, c.strict_cutoff = self.subcell, self.strict_cutoff
        c.surface_dipole, c.molpbc = self.surface_dipole, self.molpbc
        c.ewald_accuracy = self.ewald_accuracy
        c.rdf_interval, c.begin_rdf, c.istep = 0, 1000000, 0
        c.limit, c.nbins = 10.0, 100


def det3(h):
    return float(np.linalg.det(h))


def init_cutoffs(ctl: Control, h: np.ndarray, nsites: int, charged: bool):
    """Ewald parameter selection (behaviour of src/startup.c:713-772, and the
    alpha=-1 rule for uncharged systems, src/startup.c:593-596)."""
    if not charged:
        ctl.alpha = -1.0
        if ctl.cutoff <= 0:
            raise ValueError("cutoff must be given for an uncharged system")
        return ctl
    vol = det3(h)
    sqrt_p = math.sqrt(-math.log(ctl.ewald_accuracy))
    if ctl.alpha == 0.0:
        ctl.alpha = (nsites * math.pi ** 3 / vol ** 2 * 5.5) ** (1.0 / 6.0)
        trial = sqrt_p / ctl.alpha
        if ctl.cutoff > trial:
            ctl.alpha = sqrt_p / ctl.cutoff
        elif ctl.cutoff == 0.0:
            ctl.cutoff = trial
    elif ctl.alpha > 0.0 and ctl.cutoff == 0.0:
        ctl.cutoff = min(sqrt_p / ctl.alpha, min(h[0, 0], h[1, 1], h[2, 2]))
    if ctl.k_cutoff == 0.0:
        ctl.k_cutoff = 2.0 * ctl.alpha * sqrt_p
    return ctl


def cubic_box(sysdef: SysDef, density: float) -> np.ndarray:
    """Cubic MD cell of the requested mass density (g/cm^3)."""
    mtot = sum(s.mass * s.nmols for s in sysdef.species)          # amu
    rho = density * unit_scale((1, -3, 0, 0), dict(m=.001, l=.01, t=1, q=1), PROG_UNIT)
    L = (mtot / rho) ** (1.0 / 3.0)
    return np.diag([L, L, L]).astype(np.float64)


def lattice_h(a, b, c, al, be, ga, nx=1, ny=1, nz=1) -> np.ndarray:
    """Upper-triangular cell matrix from lengths/angles (src/input.c:405-413)."""
    d = math.pi / 180.0
    ca, cb, cg, sg = math.cos(al * d), math.cos(be * d), math.cos(ga * d), math.sin(ga * d)
    h = np.zeros((3, 3))
    h[0, 0] = nx * a
    h[0, 1] = ny * b * cg
    h[1, 1] = ny * b * sg
    h[0, 2] = nz * c * cb
    h[1, 2] = nz * c / sg * (ca - cb * cg)
    h[2, 2] = nz * c / sg * math.sqrt(1 - ca * ca - cb * cb - cg * cg + 2 * ca * cb * cg)
    return h


def quat_to_rot(q: np.ndarray) -> np.ndarray:
    """Rotation matrices [n,3,3] from unit quaternions [n,4] (q0 scalar part)."""
    q0, q1, q2, q3 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((len(q), 3, 3))
    R[:, 0, 0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3
    R[:, 0, 1] = 2 * (q1 * q2 - q0 * q3)
    R[:, 0, 2] = 2 * (q1 * q3 + q0 * q2)
    R[:, 1, 0] = 2 * (q1 * q2 + q0 * q3)
    R[:, 1, 1] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3
    R[:, 1, 2] = 2 * (q2 * q3 - q0 * q1)
    R[:, 2, 0] = 2 * (q1 * q3 - q0 * q2)
    R[:, 2, 1] = 2 * (q2 * q3 + q0 * q1)
    R[:, 2, 2] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3
    return R


def random_quats(rng, n):
    q = rng.standard_normal((n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


@dataclass
class MoldySystem:
    """One configuration: definition + box + molecular coordinates + control."""
    sysdef: SysDef
    h: np.ndarray                   # [3,3] upper triangular
    c_of_m: np.ndarray              # [nmols,3] scaled, [-0.5,0.5)
    quat: np.ndarray                # [nmols,4] (ignored for monatomic species)
    control: Control
    _keep: list = field(default_factory=list, repr=False)

    # ---- sizes ----
    @property
    def nmols(self):
        return sum(s.nmols for s in self.sysdef.species)

    @property
    def nsites(self):
        return sum(s.nmols * s.nsites for s in self.sysdef.species)

    @property
    def nsites_xf(self):
        return sum(s.nmols * s.nsites for s in self.sysdef.species if not s.framework)

    # ---- per-site arrays ----
    def site_ids(self) -> np.ndarray:
        return np.concatenate([np.tile(s.site_id, s.nmols) for s in self.sysdef.species]
                              ).astype(np.int32)

    def charges(self) -> np.ndarray:
        return self.sysdef.site_charge[self.site_ids()].astype(np.float64)

    def molmap(self) -> np.ndarray:
        out, m0 = [], 0
        for s in self.sysdef.species:
            out.append(np.repeat(np.arange(m0, m0 + s.nmols), s.nsites))
            m0 += s.nmols
        return np.concatenate(out).astype(np.int32)

    def make_sites(self, wrap: bool = True) -> np.ndarray:
        """site[3][nsarray] block exactly as eval_forces() builds it
        (site-wise periodic wrap, behaviour of src/algorith.c:169-217)."""
        ns = self.nsites
        out = np.zeros((3, abi.nsarray(ns)))
        hinv = np.linalg.inv(self.h)
        i0 = m0 = 0
        for s in self.sysdef.species:
            com = self.c_of_m[m0:m0 + s.nmols] @ self.h.T                  # [nmols,3]
            if s.rdof:
                R = quat_to_rot(self.quat[m0:m0 + s.nmols])
                rel = np.einsum("mij,sj->msi", R, s.p_f_sites)
            else:
                rel = np.broadcast_to(s.p_f_sites[None], (s.nmols, s.nsites, 3))
            xyz = (com[:, None, :] + rel).reshape(-1, 3)
            if wrap:
                t = np.floor(xyz @ hinv.T + 0.5)
                xyz = xyz - t @ self.h.T
            n = s.nmols * s.nsites
            out[:, i0:i0 + n] = xyz.T
            i0 += n
            m0 += s.nmols
        return out

    # ---- C view (what eval_forces hands to force_calc/ewald) ----
    def cstructs(self):
        sd = self.sysdef
        nsp = len(sd.species)
        spec_arr = (abi.spec_mt * nsp)()
        keep = []
        com = np.ascontiguousarray(self.c_of_m, dtype=np.float64)
        quat = np.ascontiguousarray(self.quat, dtype=np.float64)
        hmat = np.ascontiguousarray(self.h, dtype=np.float64)
        keep += [com, quat, hmat]
        m0 = 0
        for i, s in enumerate(sd.species):
            sp = spec_arr[i]
            sp.mass, sp.charge = s.mass, s.charge
            sp.nsites, sp.nmols = s.nsites, s.nmols
            sp.rdof, sp.framework = s.rdof, int(s.framework)
            sp.name = s.name.encode()[:abi.L_SPEC - 1]
            ids = np.ascontiguousarray(s.site_id, dtype=np.int32)
            pfs = np.ascontiguousarray(s.p_f_sites, dtype=np.float64)
            keep += [ids, pfs]
            sp.site_id = ids.ctypes.data_as(C.POINTER(C.c_int))
            sp.p_f_sites = pfs.ctypes.data_as(C.POINTER(abi.vec_mt))
            sp.c_of_m = C.cast(com.ctypes.data + 24 * m0, C.POINTER(abi.vec_mt))
            sp.quat = C.cast(quat.ctypes.data + 32 * m0, C.POINTER(abi.quat_mt))
            m0 += s.nmols
        sysm = abi.system_mt()
        sysm.nsites, sysm.nmols = self.nsites, self.nmols
        sysm.nmols_r = sum(s.nmols for s in sd.species if s.rdof)
        sysm.nspecies, sysm.max_id = nsp, sd.max_id
        sysm.ptype, sysm.n_potpar = sd.ptype, sd.n_potpar
        sysm.c_of_m = com.ctypes.data_as(C.POINTER(abi.vec_mt))
        sysm.quat = quat.ctypes.data_as(C.POINTER(abi.quat_mt))
        sysm.h = hmat.ctypes.data_as(C.POINTER(abi.vec_mt))
        sysm.ts = 1.0
        pot = (abi.pot_mt * (sd.max_id * sd.max_id))()
        for i in range(sd.max_id):
            for j in range(sd.max_id):
                p = pot[i * sd.max_id + j]
                p.flag = int(np.any(sd.potpar[i, j] != 0.0))
                for k in range(abi.NPOTP):
                    p.p[k] = sd.potpar[i, j, k]
        self._keep = keep + [spec_arr, pot]
        return sysm, spec_arr, pot

    def eval_forces_args(self):
        """Argument list of eval_forces() (src/accel.c:398-407) for this configuration, laid out as Moldy's
        allocate_dynamics does: spec->quat is NULL for species without rotational freedom (src/startup.c:560),
        force[ispec] / torque[ispec] point into one contiguous [nmols][3] / [nmols_r][3] block.
        Returns (args tuple for the C call, dict of the numpy outputs)."""
        sysm, spec, pot = self.cstructs()
        sd = self.sysdef
        nsp = len(sd.species)
        site_info = (abi.site_mt * sd.max_id)()
        for i in range(sd.max_id):
            site_info[i].mass, site_info[i].charge = float(sd.site_mass[i]), float(sd.site_charge[i])
        nmols_r = sum(s.nmols for s in sd.species if s.rdof)
        force = np.zeros((self.nmols, 3))
        torque = np.zeros((max(nmols_r, 1), 3))
        fptr = (C.POINTER(abi.vec_mt) * nsp)()
        tptr = (C.POINTER(abi.vec_mt) * nsp)()
        m0 = r0 = 0
        for i, s in enumerate(sd.species):
            fptr[i] = C.cast(force.ctypes.data + 24 * m0, C.POINTER(abi.vec_mt))
            if s.rdof:
                tptr[i] = C.cast(torque.ctypes.data + 24 * r0, C.POINTER(abi.vec_mt))
                r0 += s.nmols
            else:
                spec[i].quat = C.POINTER(abi.quat_mt)()
            m0 += s.nmols
        pe = np.zeros(abi.NPE)
        dip = np.zeros(3)
        stress = np.zeros((3, 3))
        DP = C.POINTER(C.c_double)
        args = (C.byref(sysm), spec, site_info, pot, pe.ctypes.data_as(DP), dip.ctypes.data_as(DP),
                stress.ctypes.data_as(C.POINTER(abi.vec_mt)), fptr, tptr)
        self._keep += [sysm, site_info, fptr, tptr]
        return args, dict(pe=pe, dip_mom=dip, stress=stress, force=force, torque=torque[:nmols_r])

    def principal_inertia(self):
        """Per species: moments of inertia about the axes of the principal frame the sites are given in,
        sum_s m_s (|r_s|^2 - r_sk^2) (src/startup.c; zero for single-site species)."""
        sd = self.sysdef
        out = []
        for s in sd.species:
            m = np.asarray(sd.site_mass)[np.asarray(s.site_id)]
            r = np.asarray(s.p_f_sites, dtype=np.float64)
            out.append(np.array([(m * ((r ** 2).sum(1) - r[:, k] ** 2)).sum() for k in range(3)]))
        return out

    def thermal_momenta(self, temperature=300.0, seed=7):
        """Maxwell-Boltzmann linear momenta (scaled: h' p) and principal-frame angular momenta [0, L1, L2, L3] for a
        synthetic dynamic state; frameworks at rest."""
        kB = 1.380658e-23 / (1.6605402e-27 * 1.0e4)                                # src/defs.h:226, amu A^2 ps^-2 per K
        rng = np.random.default_rng(seed)
        mom = np.zeros((self.nmols, 3)); amom = np.zeros((self.nmols, 4))
        m0 = 0
        for s, inert in zip(self.sysdef.species, self.principal_inertia()):
            if not s.framework:
                p = rng.standard_normal((s.nmols, 3)) * np.sqrt(s.mass * kB * temperature)
                mom[m0:m0 + s.nmols] = p @ self.h                                 # scaled momentum h' p
                if s.rdof:
                    amom[m0:m0 + s.nmols, 1:] = rng.standard_normal((s.nmols, 3)) * np.sqrt(np.maximum(inert, 0.0) * kB * temperature)
            m0 += s.nmols
        return mom, amom

    def do_step_args(self, mom, amom, step, istep=2):
        """Argument list of do_step() (src/accel.c:626-637) for this configuration with the given momenta (arrays are
        updated in place by the call): returns (args, outputs dict incl. the state arrays, setter for `control`)."""
        args, out = self.eval_forces_args()
        sysm, spec = args[0]._obj, args[1]
        sd = self.sysdef
        com, quat = self._keep[0], self._keep[1]
        mom = np.ascontiguousarray(mom, dtype=np.float64).copy()
        amom = np.ascontiguousarray(amom, dtype=np.float64).copy()
        m0 = 0
        for i, (s, inert) in enumerate(zip(sd.species, self.principal_inertia())):
            spec[i].mom = C.cast(mom.ctypes.data + 24 * m0, C.POINTER(abi.vec_mt))
            for k in range(3):
                spec[i].inertia[k] = float(inert[k])
            if s.rdof:
                spec[i].amom = C.cast(amom.ctypes.data + 32 * m0, C.POINTER(abi.quat_mt))
            m0 += s.nmols
        sysm.mom = mom.ctypes.data_as(C.POINTER(abi.vec_mt))
        sysm.amom = amom.ctypes.data_as(C.POINTER(abi.quat_mt))
        sysm.ts, sysm.tsmom = 1.0, 0.0
        hmom = np.zeros((3, 3))
        sysm.hmom = hmom.ctypes.data_as(C.POINTER(abi.vec_mt))
        sysm.d_of_f = int(sum(s.nmols * (3 + s.rdof) for s in sd.species if not s.framework))
        meansq = np.zeros((len(sd.species), 2, 3))
        self._keep += [mom, amom, meansq, hmom]
        full = (args[0], args[1], args[2], args[3], meansq.ctypes.data_as(C.POINTER(C.c_double)), args[4], args[5], args[6],
                None, 0, 0)
        out.update(com=com, quat=quat, mom=mom, amom=amom, meansq=meansq, sysm=sysm)

        def set_control(c):
            self.control.fill(c)
            c.step, c.istep, c.const_temp, c.const_pressure, c.nosymmetric_rot = float(step), int(istep), 0, 0, 0
            c.dump_interval, c.dump_level = 0, 0
            c.ttmass, c.rtmass, c.pmass, c.temp = 100.0, 100.0, 100.0, 300.0      # the reference's defaults (src/startup.c)
        return full, out, set_control

    # ---- replication ----
    def replicate(self, nx, ny=None, nz=None, jitter=0.0, seed=0) -> "MoldySystem":
        """Periodic replication nx*ny*nz with optional rigid-molecule jitter (A)."""
        ny = nx if ny is None else ny
        nz = nx if nz is None else nz
        reps = nx * ny * nz
        sd = self.sysdef
        new_species = [Species(s.name, s.nmols * reps, s.framework, s.site_id, s.p_f_sites,
                               s.mass, s.charge) for s in sd.species]
        nsd = SysDef(new_species, sd.site_mass, sd.site_charge, sd.site_name, sd.ptype, sd.potpar)
        shifts = np.array([(i, j, k) for i in range(nx) for j in range(ny) for k in range(nz)],
                          dtype=np.float64)
        scale = np.array([nx, ny, nz], dtype=np.float64)
        coms, quats, m0 = [], [], 0
        for s in sd.species:
            c = self.c_of_m[m0:m0 + s.nmols] + 0.5                         # [0,1)
            cc = ((c[None, :, :] + shifts[:, None, :]) / scale).reshape(-1, 3) - 0.5
            coms.append(cc)
            quats.append(np.tile(self.quat[m0:m0 + s.nmols], (reps, 1)))
            m0 += s.nmols
        com = np.concatenate(coms)
        quat = np.concatenate(quats)
        hn = self.h * scale[None, :]
        if jitter > 0:
            rng = np.random.default_rng(seed)
            com = com + (rng.standard_normal(com.shape) * jitter) @ np.linalg.inv(hn).T
            dq = rng.standard_normal(quat.shape) * (jitter * 0.2)
            quat = quat + dq
            quat /= np.linalg.norm(quat, axis=1, keepdims=True)
        com = com - np.floor(com + 0.5)
        com[com >= 0.5] -= 1.0
        ctl = Control(**{**self.control.__dict__})
        return MoldySystem(nsd, hn, com, quat, ctl)


def _grid_coms(rng, nmols, jitter):
    n = int(math.ceil(nmols ** (1.0 / 3.0) - 1e-9))
    g = (np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) / n
    pick = rng.permutation(len(g))[:nmols]
    c = g[np.sort(pick)] + rng.uniform(-jitter, jitter, (nmols, 3)) / n - 0.5
    c = c - np.floor(c + 0.5)
    c[c >= 0.5] -= 1.0
    return c


def build(spec_text: str, control: Control, *, time_unit=1.0e-13, h=None, seed=1,
          jitter=0.25, com=None, quat=None, nmols_override=None, auto_cutoffs=True) -> MoldySystem:
    """Create a configuration.  Molecules sit on a jittered simple-cubic grid of
    scaled positions with random orientations unless `com`/`quat` are given."""
    sd = parse_sysdef(spec_text, time_unit=time_unit)
    if nmols_override:
        for s, n in zip(sd.species, nmols_override):
            s.nmols = n
    if h is None:
        h = cubic_box(sd, control.density)
    nm = sum(s.nmols for s in sd.species)
    rng = np.random.default_rng(seed)
    if com is None:
        com = _grid_coms(rng, nm, jitter)
        # species are interleaved on the grid so that ions are dispersed
        com = com[rng.permutation(nm)]
    if quat is None:
        quat = random_quats(rng, nm)
    ns = sum(s.nmols * s.nsites for s in sd.species)
    charged = bool(np.any(sd.site_charge[np.concatenate([s.site_id for s in sd.species])] != 0))
    ctl = Control(**{**control.__dict__})
    if auto_cutoffs:
        init_cutoffs(ctl, np.asarray(h, dtype=np.float64), ns, charged)
    return MoldySystem(sd, np.asarray(h, dtype=np.float64), np.asarray(com, dtype=np.float64),
                       np.asarray(quat, dtype=np.float64), ctl)


def load_textsave(text: str, control: Control, time_unit=KCAL_TIME_UNIT,
                  auto_cutoffs=True, **units) -> MoldySystem:
    """Read the sys-spec + lattice-start part of a Moldy `text-mode-save` file
    (as written by the reference's print_config, src/output.c:515-606): species
    in their principal frame, then `a b c alpha beta gamma nx ny nz` and one
    `Name fx fy fz [q0 q1 q2 q3]` line per molecule, fractional coordinates."""
    lines = [ln.strip() for ln in text.strip().splitlines() if ln.strip()]
    ends = [i for i, ln in enumerate(lines) if ln.lower() == "end"]
    spec_text = "\n".join(lines[:ends[1] + 1])
    sd = parse_sysdef(spec_text, time_unit=time_unit, **units)
    cell = lines[ends[1] + 1].split()
    a, b, c, al, be, ga = (float(t) for t in cell[:6])
    nx, ny, nz = (int(t) for t in cell[6:9])
    if (nx, ny, nz) != (1, 1, 1):
        raise ValueError("load_textsave expects an unreplicated save file")
    h = lattice_h(a, b, c, al, be, ga)
    by_name = {s.name.lower(): ([], []) for s in sd.species}
    for ln in lines[ends[1] + 2:ends[2]]:
        tok = ln.split()
        f = np.array([float(t) for t in tok[1:4]])
        f -= np.floor(f)                               # "%g" may print 1 for 0.9999999
        q = [float(t) for t in tok[4:8]] if len(tok) >= 8 else [1.0, 0.0, 0.0, 0.0]
        by_name[tok[0].lower()][0].append(f - 0.5)
        by_name[tok[0].lower()][1].append(q)
    com = np.concatenate([np.array(by_name[s.name.lower()][0]).reshape(-1, 3) for s in sd.species])
    quat = np.concatenate([np.array(by_name[s.name.lower()][1]).reshape(-1, 4) for s in sd.species])
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    com[com >= 0.5] -= 1.0
    ns = sum(s.nmols * s.nsites for s in sd.species)
    ctl = Control(**{**control.__dict__})
    if auto_cutoffs:
        init_cutoffs(ctl, h, ns, True)
    return MoldySystem(sd, h, com, quat, ctl)


GOLDEN_DIR = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(
    __import__("os").path.abspath(__file__))), "tests", "golden")


def _eq(name):
    with open(__import__("os").path.join(GOLDEN_DIR, name)) as f:
        return f.read()


# ---- the named benchmark / parity families ----------
def argon(seed=1) -> MoldySystem:
    """C0: 108 LJ atoms at the triple point, strict cutoff, no Ewald."""
    ctl = Control(cutoff=8.5125, subcell=2.0, strict_cutoff=1, density=1.428)
    return build(SPEC_ARGON, ctl, seed=seed, jitter=0.15)


def tip4p(n=1, seed=1, jitter=0.02, equilibrated=True) -> MoldySystem:
    """C1/C4: 256 TIP4P waters replicated n^3 (Ewald parameters re-derived by
    init_cutoffs for the replicated N, as SURVEY 8d prescribes).  The 256-molecule
    cell is the equilibrated fixture tests/golden/tip4p_256_eq.txt (written by the
    reference binary, see tests/golden/make_fixtures.py); replicas get a small
    rigid-body jitter so the big system is not exactly periodic."""
    ctl = Control(subcell=2.5, surface_dipole=1, density=1.0)
    if equilibrated:
        base = load_textsave(_eq("tip4p_256_eq.txt"), ctl, auto_cutoffs=(n == 1))
    else:
        base = build(SPEC_TIP4P, ctl, time_unit=KCAL_TIME_UNIT, seed=seed, auto_cutoffs=(n == 1))
    if n == 1:
        return base
    big = base.replicate(n, jitter=jitter, seed=seed)
    init_cutoffs(big.control, big.h, big.nsites, True)
    return big


def tips2(seed=1, equilibrated=True) -> MoldySystem:
    ctl = Control(subcell=2.5, surface_dipole=1, density=1.0)
    if equilibrated:
        return load_textsave(_eq("tips2_64_eq.txt"), ctl)
    return build(SPEC_TIPS2, ctl, time_unit=KCAL_TIME_UNIT, seed=seed)


def mgcl2(n=1, seed=1, explicit=True) -> MoldySystem:
    """C2: 200 MCY waters + 4 Mg2+ + 8 Cl-."""
    ctl = (Control(cutoff=6.25, k_cutoff=3.0, alpha=0.45, density=1.0) if explicit and n == 1
           else Control(density=1.0))
    base = load_textsave(_eq("mgcl2_812_eq.txt"), ctl, auto_cutoffs=(n == 1))
    if n == 1:
        return base
    big = base.replicate(n, jitter=0.02, seed=seed)
    init_cutoffs(big.control, big.h, big.nsites, True)
    return big


def quartz(n=4, seed=1, jitter=0.03, pinned_cutoff=True) -> MoldySystem:
    """C3: BKS alpha-quartz, n^3 unit cells, triclinic cell (gamma=120)."""
    sd_text = SPEC_QUARTZ.replace("Oxygen 384", f"Oxygen {6 * n ** 3}").replace(
        "Silicon 192", f"Silicon {3 * n ** 3}")
    h = lattice_h(*QUARTZ_CELL, n, n, n)
    coms = {"Oxygen": [], "Silicon": []}
    for name, fx, fy, fz in QUARTZ_BASIS:
        for i in range(n):
            for j in range(n):
                for k in range(n):
                    coms[name].append(((fx + i) / n, (fy + j) / n, (fz + k) / n))
    com = np.array(coms["Oxygen"] + coms["Silicon"]) - 0.5
    rng = np.random.default_rng(seed)
    com = com + (rng.standard_normal(com.shape) * jitter) @ np.linalg.inv(h).T
    com = com - np.floor(com + 0.5)
    com[com >= 0.5] -= 1.0
    ctl = Control(cutoff=8.48 if pinned_cutoff else 0.0, subcell=3.0)
    quat = np.tile([1.0, 0, 0, 0], (len(com), 1))
    return build(sd_text, ctl, time_unit=EV_TIME_UNIT, h=h, com=com, quat=quat)


def clay() -> MoldySystem:
    """The reference's one shipped framework system (src/examples/control.clay): 64 waters + 4 cations around one rigid
    montmorillonite sheet (framework species, net charge -8), MCY potentials, explicit Ewald parameters, the example's own
    units.  Configuration: tests/golden/clay_montmorillonite.txt (the sys-spec part of src/examples/control.clay)."""
    ctl = Control(cutoff=12.6, alpha=0.35, k_cutoff=2.0, subcell=1.8, strict_cutoff=0, surface_dipole=0)
    return load_textsave(_eq("clay_montmorillonite.txt"), ctl, time_unit=1.0e-12, auto_cutoffs=False,
                         mass_unit=1.660565e-27, length_unit=1.0e-10, charge_unit=4.298401e-22)


def slab(seed=3) -> MoldySystem:
    """Framework test system: water + cations around one rigid charged sheet."""
    rng = np.random.default_rng(seed)
    L = 14.0
    rows = []
    for i in range(5):
        for j in range(5):
            x, y = (i + 0.5) * L / 5 - L / 2, (j + 0.5) * L / 5 - L / 2
            q = -0.24
            rows.append(f"5 {x:.4f} {y:.4f} 0 28.0 {q} Sf" if not rows
                        else f"5 {x:.4f} {y:.4f} 0")
    text = SPEC_SLAB % "\n".join(rows)
    ctl = Control(cutoff=6.5, alpha=0.42, k_cutoff=2.6, subcell=1.75)
    h = np.diag([L, L, L + 4.0])
    sd = parse_sysdef(text, time_unit=KCAL_TIME_UNIT)
    nm = sum(s.nmols for s in sd.species)
    com = _grid_coms(rng, nm + 40, 0.2)
    com = com[np.abs(com[:, 2]) > 0.12][:nm - 1]
    com = com[rng.permutation(len(com))]
    com = np.concatenate([com, np.zeros((1, 3))])
    return build(text, ctl, time_unit=KCAL_TIME_UNIT, h=h, com=com,
                 quat=np.concatenate([random_quats(rng, nm - 1), [[1.0, 0, 0, 0]]]))
