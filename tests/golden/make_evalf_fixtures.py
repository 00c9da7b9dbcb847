"""Generates tests/golden/evalf_<case>_sd<0|1>.npz: outputs of the reference's own eval_forces() (src/accel.c:398-617,
compiled in place into oracle/_ref/libmoldyref_evalf.so by oracle/Makefile) for every parity system of tests/cases.py,
with and without the surface-dipole term.  Run where /root/reference exists:  python tests/golden/make_evalf_fixtures.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref as refmod          # noqa: E402
from tests import cases                   # noqa: E402

if __name__ == "__main__":
    for name, make in cases.GOLDEN_CASES.items():
        for sd in (0, 1):
            ms = make()
            ms.control.surface_dipole = sd
            out = refmod.RefLib(evalf=True).eval_forces(ms)
            np.savez_compressed(os.path.join(HERE, f"evalf_{name}_sd{sd}.npz"), force=out["force"], torque=out["torque"],
                                pe=out["pe"], stress=out["stress"], dip_mom=out["dip_mom"])
            print(name, sd, out["pe"], out["dip_mom"])
