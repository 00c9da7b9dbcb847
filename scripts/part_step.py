"""One rank's share of a P-rank step on ONE GPU (no peers): cells, real-space sum over batches r of P, structure-factor
partial sums and back-projection for the charged sites r of P.  For ncu launch lists of the per-rank fixed costs:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/part_step.py 0 8 tip4p10"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from moldy_b200 import lib  # noqa: E402

r, P = int(sys.argv[1]), int(sys.argv[2])
workload = sys.argv[3] if len(sys.argv) > 3 else "tip4p10"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ms = bench.build_system(workload)
eng = lib.Engine(0)
eng.configure(ms)
eng.set_partition(r, P)
eng.set_sites_host(np.ascontiguousarray(ms.make_sites()[:, :ms.nsites]))
st = torch.cuda.current_stream().cuda_stream
out = torch.zeros(eng.out_doubles(), dtype=torch.float64, device="cuda")
psum = torch.zeros(eng.recip_sum_doubles(), dtype=torch.float64, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for k in range(steps):
    torch.cuda.nvtx.range_push(f"step{k}")
    eng.zero_out(out.data_ptr(), st)
    ev[0].record()
    eng.build_cells(st)
    ev[1].record()
    eng.force_real(out.data_ptr(), st)
    ev[2].record()
    eng.recip_partial(psum.data_ptr(), st)
    ev[3].record()
    eng.recip_finish(psum.data_ptr(), out.data_ptr(), st)
    ev[4].record()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print("step", k, [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(4)], flush=True)
