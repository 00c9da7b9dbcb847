"""Device time of the molecular-frame building blocks (mdb_make_sites, mdb_mol_forces) at bench size.
usage: python scripts/molframe_probe.py [n=10]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from moldy_b200 import lib, systems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ms = systems.tip4p(n); sp = ms.sysdef.species[0]; N = ms.nsites
eng = lib.Engine(0); eng.configure(ms)
st = torch.cuda.current_stream().cuda_stream
com = torch.from_numpy(np.ascontiguousarray(ms.c_of_m)).cuda(); quat = torch.from_numpy(np.ascontiguousarray(ms.quat)).cuda()
pfs = torch.from_numpy(np.ascontiguousarray(sp.p_f_sites, dtype=np.float64)).cuda()
out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
f = torch.zeros((sp.nmols, 3), dtype=torch.float64, device="cuda"); t = torch.zeros_like(f)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, reps=10):
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    return tot / reps
t1 = timed(lambda: eng.make_sites(ms.h, com.data_ptr(), quat.data_ptr(), pfs.data_ptr(), sp.nmols, sp.nsites, 0, True, st))
t2 = timed(lambda: eng.mol_forces(out.data_ptr(), quat.data_ptr(), pfs.data_ptr(), sp.nmols, sp.nsites, 0, f.data_ptr(), t.data_ptr(), st))
b1 = sp.nmols * 56 + N * 24; b2 = N * 24 + sp.nmols * (32 + 48)
print(f"N={N} molecules={sp.nmols}: k_make_sites {t1*1e3:.1f} us = {b1/t1/1e6:.0f} GB/s of {b1/1e6:.1f} MB; "
      f"k_mol_forces {t2*1e3:.1f} us = {b2/t2/1e6:.0f} GB/s of {b2/1e6:.1f} MB")
