"""Real space beside k-space (mdb_force_both): step time against the sequential phases, for several filler grids.
usage: python scripts/overlap_probe.py [n=10] [reps=3] [workload=tip4p]   (MDB_KF_NSB / MDB_MCW builds select the k-space shapes)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moldy_b200 import lib, systems

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = sys.argv[3] if len(sys.argv) > 3 else "tip4p"
ms = getattr(systems, wl)(n)
site = ms.make_sites()
eng = lib.Engine(0)
eng.configure(ms)
N = ms.nsites
xyz = torch.from_numpy(site[:, :N].copy()).cuda()
out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
eng.set_sites_device(xyz[0].data_ptr(), xyz[1].data_ptr(), xyz[2].data_ptr(), st)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def run(fn):
    ts = []
    for r in range(reps + 1):
        flush.zero_()
        eng.zero_out(out.data_ptr(), st)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    h = eng.read_out(out.data_ptr(), st)
    f, pe, s = lib.unpack(h, N)
    return min(ts[1:]), f, pe, s

def seq():
    eng.build_cells(st); eng.force_real(out.data_ptr(), st); eng.force_recip(out.data_ptr(), st)
def both():
    eng.build_cells(st); eng.force_both(out.data_ptr(), st)

t0, f0, pe0, s0 = run(seq)
print(f"{wl} n={n} N={N} nsb={os.environ.get('MDB_KF_NSB','-')}: sequential {t0:.3f} ms  pe {pe0}")
for fb, ft in [(-1, 0), (0, 0), (148, 64), (148, 128), (296, 64), (296, 128), (444, 64)]:
    eng.set_overlap(fb, ft)
    t, f, pe, s = run(both)
    err = np.linalg.norm(f - f0) / np.linalg.norm(f0)
    print(f"  filler {fb:4d} x {ft:3d}: {t:.3f} ms  ({100*(t0-t)/t0:+.1f} %)  drew {eng.overlap_filled()} batches  force rel diff {err:.2e}  pe diff {np.abs(pe-pe0)/np.abs(pe0)}  stress diff {np.abs(s-s0).max()/np.abs(s0).max():.1e}")
