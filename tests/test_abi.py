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
