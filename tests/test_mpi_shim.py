"""CPU test (-m "not gpu") of the bench baseline's tooling: the reference's SPMD build -- the UNCHANGED parallel.c with
-DSPMD -DMPI over oracle/mpi_shim (fork + shared-memory MPI stand-in) -- must print the same run as the serial program
(par_rsum/par_dsum of src/parallel.c:549-588 over real processes)."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SERIAL = os.path.join(ROOT, "oracle", "_ref", "moldy")
MPI = os.path.join(ROOT, "oracle", "_ref", "moldy_mpi")

CONTROL = """title=mpi shim test
surface-dipole=1
temperature=300
subcell=2.5
lattice-start=1
sys-spec-file=tip4p_256_eq.txt
scale-interval=1000000
scale-end=0
step=0.0005
nsteps=6
print-interval=3
average-interval=100000000
begin-average=100000000
roll-interval=1
dump-level=0
backup-interval=0
time-unit=4.8888213e-14
end
"""


def _values(text):
    res, lines = {}, text.splitlines()
    for i, ln in enumerate(lines):
        m = re.match(r"=+ Timestep (\d+)\s+Current values", ln)
        if m:
            res[int(m.group(1))] = np.array([float(t) for row in lines[i + 1:i + 4] for t in row.split()])
    return res


@pytest.mark.skipif(not (os.path.exists(SERIAL) and os.path.exists(MPI)), reason="oracle/_ref/moldy{,_mpi} not built")
@pytest.mark.parametrize("nranks", [2, 3])
def test_spmd_build_over_the_mpi_shim_prints_the_serial_run(tmp_path, nranks):
    outs = []
    for binary, env in ((SERIAL, {}), (MPI, {"MOLDY_MPI_NP": str(nranks)})):
        d = tmp_path / (os.path.basename(binary) + env.get("MOLDY_MPI_NP", ""))
        d.mkdir()
        shutil.copy(os.path.join(ROOT, "tests", "golden", "tip4p_256_eq.txt"), d)
        (d / "control").write_text(CONTROL)
        r = subprocess.run([binary, "control"], cwd=d, env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        outs.append(r.stdout)
    a, b = _values(outs[0]), _values(outs[1])
    assert sorted(a) == sorted(b) == [3, 6]
    for k in a:
        assert np.allclose(a[k], b[k], rtol=2e-5, atol=2e-2), (k, a[k], b[k])
    assert f"on {nranks} processors" in outs[1]
