"""GPU tests of the multi-GPU layer (moldy_b200/csrc/mdb_peer.cu): par_rsum/par_dsum (src/parallel.c:549-588) as
peer-memory kernels.  The P-rank result must equal the 1-rank result (<= 1e-11: different summation order) and the
compiled reference's (golden fixtures / full-size records), and be bit-identical on every rank (Moldy's DESYNC check,
src/main.c:262-273).

  * several ranks on ONE device (always runs: same kernels, local windows);
  * one rank per device in one process (needs >= 2 GPUs);
  * one process per GPU over CUDA IPC (torchrun, needs >= 2 GPUs): tests/peer_worker.py.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from moldy_b200 import lib, spmd, systems
from tests import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _single(ms):
    eng = lib.Engine(0)
    eng.configure(ms)
    eng.set_sites_host(ms.make_sites())
    st = torch.cuda.current_stream().cuda_stream
    out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
    eng.build_cells(st)
    eng.force_real(out.data_ptr(), st)
    eng.force_recip(out.data_ptr(), st)
    torch.cuda.synchronize()
    res = lib.unpack(out.cpu().numpy(), ms.nsites)
    eng.close()
    return res


def _check_group(ms, devices, gold=None):
    g = spmd.PeerGroup(ms, devices)
    g.set_sites(ms.make_sites())
    for _ in range(2):                       # twice: the second step runs on the other half of the double buffer
        g.step(gather=True)
    g.synchronize()
    blocks = [g.result(r) for r in range(len(devices))]
    g.close()
    n = ms.nsites
    for b in blocks[1:]:                     # identical bits on every rank
        assert np.array_equal(b[:3 * n + 11], blocks[0][:3 * n + 11])
    f, pe, s = lib.unpack(blocks[0], n)
    f1, pe1, s1 = _single(ms)
    assert cases.rel_rms(f, f1) < 1e-11
    assert np.allclose(pe, pe1, rtol=1e-11, atol=0.0)
    iu = np.triu_indices(3)
    assert np.linalg.norm(s[iu] - s1[iu]) < 1e-11 * np.linalg.norm(s1[iu])
    return f, pe, s


@pytest.mark.parametrize("name,world", [("tip4p_2", 2), ("tip4p_2", 3), ("mgcl2", 4), ("quartz", 2), ("slab_framework", 2),
                                        ("argon", 2)])
def test_ranks_sharing_one_device_sum_to_the_single_rank_result(name, world):
    ms = cases.GOLDEN_CASES[name]()
    f, pe, s = _check_group(ms, [0] * world)
    gold = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    assert cases.rel_rms(f, gold["force"]) < 1e-10


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("name", ["tip4p_5", "mgcl2_7"])
def test_one_rank_per_device_against_the_reference_record(name):
    from tests.test_gpu_large import _check_against_record
    ms = cases.LARGE_CASES[name]()
    ndev = torch.cuda.device_count()
    f, pe, s = _check_group(ms, list(range(ndev)))
    g = np.load(os.path.join(GOLD, f"large_{name}.npz"))
    # the engine-level block holds the k-space energy without the constants force_calc()/ewald() add on rank 0
    # (intramolecular correction, self energy): compare forces and stress, and the energies through the 1-rank run above
    pe_ref = g["pe"].copy()
    lib.reset()
    one = lib.eval_forces(ms)
    lib.reset()
    _check_against_record(g, f, pe + (one["pe"] - _single(ms)[1]), s, f"{name} x{ndev}")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_one_process_per_gpu_over_cuda_ipc(tmp_path):
    """torchrun, one process per GPU: windows mapped through CUDA IPC handles; host slices in, host slices out."""
    ndev = min(torch.cuda.device_count(), 8)
    out = tmp_path / "peer.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ndev}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "peer_worker.py"), "tip4p_5", str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["world"] == ndev
    assert res["force_relrms_vs_record"] < 1e-10 and res["stress_rel"] < 1e-11 and res["proj_rel"] < 1e-10
    assert res["nccl_vs_peer_relrms"] < 1e-11


# ---- several GPUs behind Moldy's own entry points (MOLDY_B200_DEVICES; mdb_group.cu) ---------------------------------
@pytest.fixture
def abi_devices():
    """Run the Moldy-level library on the given device list; back to one engine afterwards."""
    old = os.environ.get("MOLDY_B200_DEVICES")

    def use(devs):
        lib.shutdown()
        os.environ["MOLDY_B200_DEVICES"] = devs
        lib.reset()
    yield use
    lib.shutdown()
    if old is None:
        os.environ.pop("MOLDY_B200_DEVICES", None)
    else:
        os.environ["MOLDY_B200_DEVICES"] = old


def _devlists():
    nd = torch.cuda.device_count()
    lists = ["0,0", "0,0,0"]
    if nd >= 2:
        lists.append("all")
    return lists


@pytest.mark.parametrize("name", ["tip4p", "tip4p_2", "mgcl2", "quartz", "slab_framework", "argon", "tips2_molpbc",
                                  "tip4p_molpbc_strict", "tips2_strict"])
def test_force_calc_and_ewald_on_a_device_group(name, abi_devices):
    """force_calc()+ewald() through the C ABI with MOLDY_B200_DEVICES: same results as the reference's (goldens)."""
    gold = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    for devs in _devlists():
        abi_devices(devs)
        assert lib.n_devices() == (torch.cuda.device_count() if devs == "all" else len(devs.split(",")))
        ms = cases.GOLDEN_CASES[name]()
        out = lib.eval_forces(ms)
        assert cases.rel_rms(out["force"], gold["force"]) < 1e-10, devs
        scale = np.abs(gold["pe"]).max()
        assert np.abs(out["pe"] - gold["pe"]).max() < 1e-11 * scale, devs
        iu = np.triu_indices(3)
        assert np.linalg.norm(out["stress"][iu] - gold["stress"][iu]) < 1e-11 * np.linalg.norm(gold["stress"][iu]), devs


@pytest.mark.parametrize("name", ["tip4p", "tip4p_2", "mgcl2", "quartz", "slab_framework", "tips2_molpbc"])
@pytest.mark.parametrize("sd", [0, 1])
def test_eval_forces_on_a_device_group(name, sd, abi_devices):
    """The whole of eval_forces() (src/accel.c:398-617) on several ranks: molecular forces, torques, energies, virial."""
    g = np.load(os.path.join(GOLD, f"evalf_{name}_sd{sd}.npz"))
    for devs in _devlists():
        abi_devices(devs)
        ms = cases.GOLDEN_CASES[name]()
        ms.control.surface_dipole = sd
        mol = lib.eval_forces_mol(ms)
        mol = lib.eval_forces_mol(ms)             # (the second call: one host thread per rank when every rank has its own GPU)
        assert cases.rel_rms(mol["force"], g["force"]) < 1e-10, devs
        if g["torque"].size:
            assert cases.rel_rms(mol["torque"], g["torque"]) < 1e-10, devs
        assert np.abs(mol["pe"] - g["pe"]).max() < 1e-11 * np.abs(g["pe"]).max(), devs
        assert np.linalg.norm(mol["stress"] - g["stress"]) < 1e-11 * np.linalg.norm(g["stress"]), devs
        if "dip_mom" in g.files:
            assert np.allclose(mol["dip_mom"], g["dip_mom"], rtol=1e-10, atol=1e-9 * np.abs(g["dip_mom"]).max() + 1e-12), devs


def test_rdf_pass_on_a_device_group(abi_devices):
    """force_calc's RDF pass with the batches split over the ranks: the pair counts add up to the reference's."""
    ref_rdf = np.load(os.path.join(GOLD, "ref_rdf.npz"))
    abi_devices("0,0,0")
    for name in ("tip4p", "quartz"):
        lib.reset()
        limit, nbins = cases.RDF_CASES[name]
        ms = cases.GOLDEN_CASES[name]()
        out = lib.eval_forces(ms, recip=False, rdf=(limit, nbins))
        rho = ms.nsites / float(np.linalg.det(ms.h))
        cnt = np.rint(out["rdf"].astype(np.float64) * rho).astype(np.int64)
        assert np.array_equal(cnt, ref_rdf[name]), name


@pytest.mark.parametrize("name", ["tip4p", "mgcl2", "tip4p_2"])
def test_do_step_on_a_device_group(name, abi_devices):
    """The NVE do_step() of the library with MOLDY_B200_DEVICES: every rank moves its share of the
    molecules, the c-of-m / quaternion block is all-gathered, the kinetic-energy and mean-square sums are added over the
    ranks.  Same trajectory, energies, stress and sums as on one engine (and, through it, as the reference's do_step:
    tests/test_gpu_md.py)."""
    step, nsteps = 0.0005, 3
    ms = cases.GOLDEN_CASES[name]()
    mom, amom = ms.thermal_momenta(seed=11)
    lib.shutdown()
    os.environ.pop("MOLDY_B200_DEVICES", None)
    lib.reset()
    want = lib.do_step(cases.GOLDEN_CASES[name](), mom, amom, step, nsteps=nsteps)
    for devs in _devlists():
        abi_devices(devs)
        got = lib.do_step(cases.GOLDEN_CASES[name](), mom, amom, step, nsteps=nsteps)
        assert np.abs(got["com"] - want["com"]).max() < 1e-11, devs
        assert np.abs(got["mom"] - want["mom"]).max() < 1e-11 * np.abs(want["mom"]).max(), devs
        assert np.abs(got["quat"] - want["quat"]).max() < 1e-11, devs
        if want["amom"].size:
            assert np.abs(got["amom"] - want["amom"]).max() < 1e-11 * max(np.abs(want["amom"]).max(), 1e-300), devs
        assert np.abs(got["pe"] - want["pe"]).max() < 1e-11 * np.abs(want["pe"]).max(), devs
        assert np.linalg.norm(got["stress"] - want["stress"]) < 1e-10 * np.linalg.norm(want["stress"]), devs
        assert np.allclose(got["meansq"], want["meansq"], rtol=1e-10, atol=1e-12 * np.abs(want["meansq"]).max()), devs
        assert np.allclose(got["dip_mom"], want["dip_mom"], rtol=1e-9, atol=1e-9 * np.abs(want["dip_mom"]).max() + 1e-12), devs
