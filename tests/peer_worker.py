"""torchrun worker of tests/test_gpu_peer.py: one process per GPU, the peer layer over CUDA IPC, host slices in and out.
Rank 0 assembles the slices and compares them with the compiled reference's record and with the NCCL path."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moldy_b200 import spmd  # noqa: E402
from tests import cases  # noqa: E402
from tests.test_gpu_large import _weights  # noqa: E402


def main():
    name, out_path = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ms = cases.LARGE_CASES[name]()
    n = ms.nsites
    hs = torch.from_numpy(np.ascontiguousarray(ms.make_sites()[:, :n])).pin_memory()
    res = {}
    for kind in ("peer", "nccl"):
        ev = spmd.SpmdForces(ms, rank, world, local, nccl=(kind == "nccl"))
        for _ in range(3):
            f, scal, lo, hi = ev.step(hs)
        full = torch.zeros((3, n), dtype=torch.float64, device="cuda")
        full[:, lo:hi] = f[:, lo:hi].cuda()
        if kind == "peer":
            dist.all_reduce(full)                 # assemble the slices (test only)
        res[kind] = (full.cpu().numpy(), scal.numpy().copy())
        ev.close()
    if rank == 0:
        g = np.load(os.path.join(ROOT, "tests", "golden", f"large_{name}.npz"))
        f, scal = res["peer"]
        stress = scal[2:11].reshape(3, 3)
        iu = np.triu_indices(3)
        proj = _weights(n) @ f.T
        out = {"world": world,
               "force_relrms_vs_record": cases.rel_rms(f[:, g["sample"]], g["fsample"]),
               "stress_rel": float(np.linalg.norm(stress[iu] - g["stress"][iu]) / np.linalg.norm(g["stress"][iu])),
               "proj_rel": float(np.abs(proj - g["proj"]).max() / np.sqrt(g["fsq"].sum())),
               "nccl_vs_peer_relrms": cases.rel_rms(f, res["nccl"][0]),
               "pe": [float(scal[0]), float(scal[1])], "pe_nccl": [float(res["nccl"][1][0]), float(res["nccl"][1][1])]}
        with open(out_path, "w") as fh:
            json.dump(out, fh)
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
