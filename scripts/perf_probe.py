"""Quick device-side timing of the hot-path phases (cells / pair / recip) with
CUDA events on the launching stream.  usage: python scripts/perf_probe.py [n=5 | tip4p_10 | quartz_48 | mgcl2_7 ...] [reps=3]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moldy_b200 import lib, systems

arg = sys.argv[1] if len(sys.argv) > 1 else "5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t0 = time.time()
if arg.isdigit():
    n = int(arg); ms = systems.tip4p(n)
else:
    from tests import cases
    n = arg; ms = cases.LARGE_CASES[arg]()
site = ms.make_sites()
print(f"{n}: N={ms.nsites} rc={ms.control.cutoff:.3f} alpha={ms.control.alpha:.5f} kc={ms.control.k_cutoff:.4f} (built in {time.time()-t0:.1f}s)")
eng = lib.Engine(0)
t0 = time.time(); eng.configure(ms); print(f"configure {time.time()-t0:.2f}s  grid={eng.grid()} nabors={2*eng.n_neighbour_cells()} nhkl={eng.n_kvectors()}")
N = ms.nsites
xyz = torch.from_numpy(site[:, :N].copy()).cuda()
out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
eng.set_sites_device(xyz[0].data_ptr(), xyz[1].data_ptr(), xyz[2].data_ptr(), st)

def timed(fn):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)

for r in range(reps):
    eng.zero_out(out.data_ptr(), st)
    tc = timed(lambda: eng.build_cells(st))
    tp = timed(lambda: eng.force_real(out.data_ptr(), st))
    tk = timed(lambda: eng.force_recip(out.data_ptr(), st))
    pairs = eng.pair_count(st)
    print(f"rep {r}: cells {tc:.3f} ms  pair {tp:.2f} ms  recip {tk:.2f} ms  total {tc+tp+tk:.2f} ms | pairs {pairs:.4g}  "
          f"pair-flops {pairs*59/tp/1e9:.2f} TF/s alg  recip (site,k) {N*eng.n_kvectors()/tk/1e6:.1f} G/s "
          f"-> {N*eng.n_kvectors()*18/tk/1e9:.2f} TF/s alg")
h = eng.read_out(out.data_ptr(), st)
f, pe, s = lib.unpack(h, N)
print("pe", pe, "sum f", f.sum(1), "launches", eng.launches())
