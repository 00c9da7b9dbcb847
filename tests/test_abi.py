"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol that
include/moldy_b200.h declares, and agrees with the Python mirror and with the
reference build on struct layouts.  No compute call is made (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from moldy_b200 import abi, lib
from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "moldy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    funcs = re.findall(r"^\s*(?:[A-Za-z_][\w\s\*]*?)\b(\w+)\s*\([^;{]*\)\s*;", text, flags=re.M)
    data = re.findall(r"extern\s+const\s+\w+\s+(\w+)\s*\[", text)
    return sorted(set(funcs) | set(data))


def test_library_exports_every_declared_symbol():
    L = lib.load()
    syms = _declared_symbols()
    for must in ("force_calc", "ewald", "kernel", "poteval", "dist_pot", "potspec", "pot_dim",
                 "mdb_create", "mdb_configure", "mdb_force_real", "mdb_force_recip",
                 "eval_forces", "mdb_eval_forces_moldy", "mdb_set_species", "mdb_eval_forces_host", "mdb_eval_result",
                 "do_step", "mdb_do_step_moldy", "mdb_md_step", "mdb_md_set_dynamics", "mdb_md_upload_state",
                 "mdb_peer_create", "mdb_peer_open", "mdb_peer_connect", "mdb_peer_step", "mdb_peer_read_slice_host",
                 "mdb_group_create", "mdb_group_force_host", "mdb_group_eval_forces_host",
                 "mdb_sites_differ_host", "mdb_dmma_peak_probe", "mdb_recip_gemm_flop",
                 "mdb_force_both", "mdb_set_overlap", "mdb_overlap_filled", "mdb_set_pair_far", "mdb_pair_far_runs"):
        assert must in syms, f"{must} not parsed from the header"
    for s in syms:
        assert hasattr(L, s), f"libmoldy_b200.so does not export {s}"


def test_struct_layouts_agree():
    L = lib.load()
    pairs = {"contr_mt": abi.contr_mt, "system_mt": abi.system_mt, "spec_mt": abi.spec_mt,
             "site_mt": abi.site_mt, "pot_mt": abi.pot_mt, "mdb_config": lib.mdb_config}
    for name, cls in pairs.items():
        assert L.mdb_sizeof(name.encode()) == C.sizeof(cls), name
    if ref.available():
        r = ref.RefLib().lib
        assert r.mdref_sizeof_control() == C.sizeof(abi.contr_mt)
        assert r.mdref_sizeof_system() == C.sizeof(abi.system_mt)
        assert r.mdref_sizeof_spec() == C.sizeof(abi.spec_mt)
        assert r.mdref_sizeof_pot() == C.sizeof(abi.pot_mt)


def test_potspec_table_matches_reference_names():
    L = lib.load()
    tab = (abi.pots_mt * 8).in_dll(L, "potspec")
    names = [(t.name.decode(), t.npar) for t in tab[:7]]
    assert names == [("lennard-jones", 2), ("buckingham", 3), ("mcy", 4), ("generic", 6), ("hiw", 3),
                     ("reserved for developer", 1), ("morse", 7)]
    assert tab[7].name is None and tab[7].npar == 0
    dims = ((abi.dim_mt * 8) * 7).in_dll(L, "pot_dim")
    assert (dims[0][0].m, dims[0][0].l, dims[0][0].t) == (1, 2, -2) and dims[0][1].l == 1
    assert (dims[6][3].m, dims[6][3].l, dims[6][3].t) == (1, 8, -2)


def test_no_cpu_fallback():
    """Without a GPU the engine refuses to exist instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        lib.Engine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "moldy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "ORACLE_DOC/") or f in ("systems.py",), f


def test_far_radius_of_the_exponential_potentials():
    """mdb_far_radius (host only): the distance beyond which exp(-r/rho) is below e^-52 of its amplitude for every
    site-type pair, and the power-law rest of the potential as p0/r^4 + p1/r^6 + p2/r^12 rows (far-run launch)."""
    import ctypes as C
    import numpy as np
    L = lib.load()
    L.mdb_far_radius.restype = C.c_double
    L.mdb_far_radius.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    NP = 8

    def call(ptype, rows):
        mid = int(round(len(rows) ** 0.5))
        pot = np.zeros((mid * mid, NP))
        for k, r in enumerate(rows):
            pot[k, :len(r)] = r
        rest = np.full((mid * mid, NP), np.nan)
        return L.mdb_far_radius(ptype, mid, pot.ctypes.data, rest.ctypes.data), rest

    # Buckingham (ptype 1): -p0/r^6 + p1 exp(-p2 r); the slowest decay of a pair WITH an exponential sets the radius
    r, rest = call(1, [(0, 0, 0), (133.5, 18003.0, 4.873), (133.5, 18003.0, 4.873), (175.0, 1388.8, 2.76)])
    assert abs(r - 52.0 / 2.76) < 1e-12
    assert np.array_equal(rest[3, :3], [0.0, -175.0, 0.0]) and np.array_equal(rest[0], np.zeros(NP))
    # no exponential anywhere, a zero decay constant with an amplitude, or a non-exponential potential type: not applicable
    assert call(1, [(1.0, 0.0, 0.0)])[0] == 0.0
    assert call(1, [(1.0, 5.0, 0.0)])[0] == 0.0
    assert call(0, [(1.0, 3.0)])[0] == 0.0 and call(4, [(1.0, 1.0, 1.0)])[0] == 0.0
    # MCY (2): both exponentials count, nothing is left of the potential beyond
    r, rest = call(2, [(1.0e6, 5.15, 0.0, 0.0), (600.0, 2.76, 270.0, 2.23), (600.0, 2.76, 270.0, 2.23), (0, 0, 0, 0)])
    assert abs(r - 52.0 / 2.23) < 1e-12 and not rest.any()
    # generic (3): p0 exp(-p1 r) + p2/r^12 - p3/r^4 - p4/r^6 - p5/r^8 -> rows (-p3, -p4, p2); an r^-8 term cannot be expressed
    r, rest = call(3, [(10.0, 4.0, 7.0, 3.0, 5.0, 0.0)])
    assert abs(r - 13.0) < 1e-12 and np.array_equal(rest[0, :3], [-3.0, -5.0, 7.0])
    assert call(3, [(10.0, 4.0, 7.0, 3.0, 5.0, 1.0)])[0] == 0.0
    # Morse / BIG (6): p0 exp((p1 - r) p2) - p3/r^6 + p4 (exp(-2 p5 (r - p6)) - 2 exp(-p5 (r - p6)))
    r, rest = call(6, [(2.0, 3.0, 6.0, 11.0, 0.5, 2.0, 1.5)])
    assert abs(r - max(3.0 + 52.0 / 6.0, 1.5 + 52.0 / 2.0)) < 1e-12 and np.array_equal(rest[0, :3], [0.0, -11.0, 0.0])
