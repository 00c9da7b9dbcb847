"""Device-side phase times of the hot path for the benchmark families of BASELINE.json's configs (CUDA events on the
launching stream, positions resident).  usage: python scripts/perf_families.py [reps=3]  -> markdown table on stdout"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moldy_b200 import lib, systems

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
CASES = [("TIP4P 5x5x5 (configs[1])", lambda: systems.tip4p(5)),
         ("MgCl2/H2O 7x7x7 (configs[2])", lambda: systems.mgcl2(7, explicit=False)),
         ("quartz 48 cells (configs[3])", lambda: systems.quartz(48, pinned_cutoff=False)),
         ("TIP4P 10x10x10 (configs[4], bench.py)", lambda: systems.tip4p(10)),
         ("TIP4P 16x16x16 (configs[4])", lambda: systems.tip4p(16))]


def timed(fn):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


print("| system | sites | potential | grid | k-vectors | real space | cells ms | pair ms | recip ms | step ms | reference pairs/step | 59-flop rate TF/s |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, make in CASES:
    ms = make()
    N = ms.nsites
    site = ms.make_sites()
    eng = lib.Engine(0)
    eng.configure(ms)
    xyz = torch.from_numpy(site[:, :N].copy()).cuda()
    out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    eng.set_sites_device(xyz[0].data_ptr(), xyz[1].data_ptr(), xyz[2].data_ptr(), st)
    best = None
    for r in range(reps + 1):
        eng.zero_out(out.data_ptr(), st)
        tc = timed(lambda: eng.build_cells(st))
        tp = timed(lambda: eng.force_real(out.data_ptr(), st))
        tk = timed(lambda: eng.force_recip(out.data_ptr(), st)) if ms.control.alpha > 0 else 0.0
        if r > 0 and (best is None or tc + tp + tk < sum(best)):
            best = (tc, tp, tk)
    pairs = eng.pair_count(st)
    pt = ["LJ", "Buckingham", "MCY", "generic", "HIW", "", "Morse"][ms.sysdef.ptype]
    tc, tp, tk = best
    print(f"| {name} | {N} | {pt} | {'x'.join(map(str, eng.grid()[1]))} | {eng.n_kvectors()} | "
          f"{'two passes by site class' if eng.pair_split() else 'one fused pass'} | {tc:.3f} | {tp:.2f} | {tk:.2f} | {tc + tp + tk:.2f} | "
          f"{pairs:.4g} | {pairs * 59 / tp / 1e9:.1f} |", flush=True)
    eng.close()
    del xyz, out
    torch.cuda.empty_cache()
