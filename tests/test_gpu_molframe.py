"""SURVEY 8f rank 1, building blocks: make_sites / mol_force / mol_torque on the device (mdb_make_sites,
mdb_mol_forces) against the oracle restatement oracle/molframe.c, which is bit-identical to the reference's own
algorith.c (tests/test_oracle_molframe.py).  Bit-exact: the sites decide the cell assignment."""
import numpy as np
import pytest

from moldy_b200 import lib
from tests import cases

pytestmark = pytest.mark.gpu


def _species_blocks(ms):
    m0 = s0 = 0
    for sp in ms.sysdef.species:
        yield sp, m0, s0
        m0 += sp.nmols
        s0 += sp.nmols * sp.nsites


@pytest.mark.parametrize("sitepbc", [True, False])
@pytest.mark.parametrize("name", ["tip4p", "mgcl2", "quartz", "slab_framework"])
def test_device_make_sites_and_molecular_forces_bit_exact(name, sitepbc, golden_dir):
    import torch
    from oracle import molframe
    ms = cases.GOLDEN_CASES[name]()
    n = ms.nsites
    eng = lib.Engine(0)
    eng.configure(ms)
    st = torch.cuda.current_stream().cuda_stream
    com = torch.from_numpy(np.ascontiguousarray(ms.c_of_m)).cuda()
    quat = torch.from_numpy(np.ascontiguousarray(ms.quat)).cuda()
    ref_sites = np.zeros((3, n))
    keep = []
    for sp, m0, s0 in _species_blocks(ms):
        pfs = torch.from_numpy(np.ascontiguousarray(sp.p_f_sites, dtype=np.float64)).cuda()
        keep.append(pfs)
        q = quat[m0:m0 + sp.nmols] if sp.rdof else None
        eng.make_sites(ms.h, com[m0:m0 + sp.nmols].data_ptr(), q.data_ptr() if q is not None else 0, pfs.data_ptr(),
                       sp.nmols, sp.nsites, s0, sitepbc, st)
        ref_sites[:, s0:s0 + sp.nmols * sp.nsites] = molframe.make_sites(
            ms.h, ms.c_of_m[m0:m0 + sp.nmols], ms.quat[m0:m0 + sp.nmols] if sp.rdof else None, sp.p_f_sites, sitepbc)
    got = eng.get_sites(st)
    assert np.array_equal(got, ref_sites)

    # forces of the device-made sites, then the molecular forces and torques of every species
    if ms.control.molpbc:
        eng.set_com_host(ms.c_of_m)
    out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
    eng.build_cells(st)
    eng.force_real(out.data_ptr(), st)
    if ms.control.alpha > 0:
        eng.force_recip(out.data_ptr(), st)
    torch.cuda.synchronize()
    f_site = out.cpu().numpy()[:3 * n].reshape(3, n)
    assert np.abs(f_site).max() > 0
    for (sp, m0, s0), pfs in zip(_species_blocks(ms), keep):
        d_f = torch.zeros((sp.nmols, 3), dtype=torch.float64, device="cuda")
        d_t = torch.zeros((sp.nmols, 3), dtype=torch.float64, device="cuda")
        q = quat[m0:m0 + sp.nmols] if sp.rdof else None
        eng.mol_forces(out.data_ptr(), q.data_ptr() if q is not None else 0, pfs.data_ptr(), sp.nmols, sp.nsites, s0,
                       d_f.data_ptr(), d_t.data_ptr() if q is not None else 0, st)
        torch.cuda.synchronize()
        fs = f_site[:, s0:s0 + sp.nmols * sp.nsites]
        assert np.array_equal(d_f.cpu().numpy(), molframe.mol_force(fs, sp.nsites))
        if q is not None:
            assert np.array_equal(d_t.cpu().numpy(), molframe.mol_torque(fs, sp.p_f_sites, ms.quat[m0:m0 + sp.nmols]))
    eng.close()
