"""Wall time of eval_forces() of libmoldy_b200.so next to force_calc()+ewald() on replicated TIP4P.
usage: python scripts/evalf_probe.py [n=10] [steps=5]      (MOLDY_B200_TIMING=1 prints the host-side split)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from moldy_b200 import lib, systems       # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ms = systems.tip4p(n)
    ms.control.surface_dipole = 1
    L = lib.load()
    ms.control.fill(lib.control())
    lib.set_thread(0, 1)
    args, out = ms.eval_forces_args()
    lib.reset()
    L.eval_forces(*args)
    L.eval_forces(*args)
    t0 = time.perf_counter()
    for _ in range(steps):
        L.eval_forces(*args)
    dt = (time.perf_counter() - t0) / steps
    print(f"eval_forces: N={ms.nsites} nmols={ms.nmols} {1e3 * dt:.2f} ms per call; pe={out['pe']} dip={out['dip_mom']}")
