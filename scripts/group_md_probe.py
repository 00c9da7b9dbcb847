"""NVE MD steps per second with the whole state resident on the GPUs of ONE process (mdb_group_md_step), against the
one-engine step.  usage: python scripts/group_md_probe.py [n=10] [steps=10] [devices=all|0,1|0,0]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moldy_b200 import lib, systems

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
devs = sys.argv[3] if len(sys.argv) > 3 else "all"
devices = list(range(torch.cuda.device_count())) if devs == "all" else [int(d) for d in devs.split(",")]
ms = systems.tip4p(n)
mom, amom = ms.thermal_momenta(seed=7)

def run(make):
    md = make()
    md.upload(ms.c_of_m, ms.quat, mom, amom)
    md.step(0.0005); md.step(0.0005)
    t0 = time.perf_counter()
    for _ in range(steps):
        sc = md.step(0.0005)
    dt = (time.perf_counter() - t0) / steps
    st = md.download()
    return dt, sc, st

eng = lib.Engine(devices[0]); eng.configure(ms)
dt1, sc1, st1 = run(lambda: lib.MdState(eng, ms))
eng.close()
g = None
def mk():
    global g
    g = lib.GroupMd(ms, devices)
    return g
dtg, scg, stg = run(mk)
g.close()
print(f"tip4p n={n} N={ms.nsites}: one engine {1e3*dt1:.3f} ms/step = {1/dt1:.2f} steps/s | group of {len(devices)} ({devs}) "
      f"{1e3*dtg:.3f} ms/step = {1/dtg:.2f} steps/s | speed-up {dt1/dtg:.2f} | pe {sc1[12]:.10e} vs {scg[12]:.10e} | "
      f"max |com diff| {np.abs(st1['com']-stg['com']).max():.2e}  max |quat diff| {np.abs(st1['quat']-stg['quat']).max():.2e}")
