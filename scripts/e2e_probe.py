"""Wall-clock breakdown of the host-buffer path force_calc()+ewald(). usage: python scripts/e2e_probe.py [n=10]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from moldy_b200 import lib, systems, abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ms = systems.tip4p(n); N = ms.nsites; nsa = abi.nsarray(N)
site = ms.make_sites()
ms.control.fill(lib.control()); lib.set_thread(0, 1)
sysm, spec, pot = ms.cstructs()
hs = torch.empty((3, nsa), dtype=torch.float64).pin_memory(); hf = torch.empty((3, nsa), dtype=torch.float64).pin_memory()
hs.numpy()[:] = site; s_np, f_np = hs.numpy(), hf.numpy()
chg = ms.charges(); pe = np.zeros(2); stress = np.zeros((3, 3))
lib.reset()
for it in range(4):
    t0 = time.perf_counter(); f_np[:] = 0.0; pe[:] = 0; stress[:] = 0
    t1 = time.perf_counter(); lib.force_calc(s_np, f_np, sysm, spec, chg, pot, pe[0:1], stress)
    t2 = time.perf_counter(); lib.ewald(s_np, f_np, sysm, spec, chg, pe[1:2], stress)
    t3 = time.perf_counter()
    print(f"iter {it}: zero {1e3*(t1-t0):.2f} ms  force_calc {1e3*(t2-t1):.2f} ms  ewald {1e3*(t3-t2):.2f} ms  total {1e3*(t3-t0):.2f} ms", file=sys.stderr)
