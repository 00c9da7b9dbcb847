"""Reduced fixtures of the compiled reference at the BASELINE.json sizes (run HERE, where
/root/reference exists; takes ~15 minutes on 8 cores; the GPU box only reads the committed files).

For each named workload the reference's own force_calc()+ewald() (oracle/_ref/libmoldyref.so,
-O2 -ffp-contract=off) run as the P ranks of its replicated-data SPMD split (rank p evaluates cells
p mod P and its block of k-vectors, src/force.c:856, src/ewald.c:495-496), one process per rank
(function statics).  The parent adds the P partial [forces | pe | stress] blocks -- that sum IS
par_rsum/par_dsum (src/accel.c:531-535, src/parallel.c:549-588) -- and keeps a reduced record:

  pe[2], stress[3,3]          complete
  sample[ns], fsample[3,ns]   forces of a seeded sample of sites (ns = 16384)
  proj[8,3]                   sum_i w_k(i) f[a,i] with w_k(i) = sin(0.37 (k+1) i + 0.11 k): every site enters
  fsq[3]                      sum_i f[a,i]^2
  cell_sha256                 digest of the int32 link-cell index of every site (the reference's cellbin())
  cell_sample[ns]             link-cell index of the sampled sites
  n_kvectors, nsites, log

usage: python tests/golden/make_large_fixtures.py [--procs=K] [name ...]   (default: all of cases.LARGE_CASES;
       --procs=K runs the 8 ranks K at a time: tip4p_16 needs ~8 GB per rank)
"""
import hashlib
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
NSAMPLE = 16384


def workloads():
    from tests import cases
    return cases.LARGE_CASES


def weights(n):
    i = np.arange(n, dtype=np.float64)
    return np.stack([np.sin(0.37 * (k + 1) * i + 0.11 * k) for k in range(8)])


def sample_sites(n):
    return np.sort(np.random.default_rng(20261018).choice(n, size=min(NSAMPLE, n), replace=False))


def _rank(args):
    name, p, P = args
    sys.path.insert(0, ROOT)
    from oracle import ref
    ms = workloads()[name]()
    r = ref.RefLib()
    r.set_thread(p, P)
    t0 = time.perf_counter()
    o = r.run(ms)
    return o["force"], o["pe"], o["stress"], o["log"], time.perf_counter() - t0


def make(name, P, procs=None):
    from oracle import ref
    ms = workloads()[name]()
    n = ms.nsites
    t0 = time.perf_counter()
    # one task per worker process (function statics of the reference); `procs` < P runs the ranks in waves (memory)
    with mp.get_context("spawn").Pool(procs or P, maxtasksperchild=1) as pool:
        parts = pool.map(_rank, [(name, p, P) for p in range(P)], chunksize=1)
    force = np.zeros((3, n)); pe = np.zeros(2); stress = np.zeros((3, 3))
    for f, e, s, _, _ in parts:                      # par_rsum / par_dsum
        force += f; pe += e; stress += s
    log = parts[0][3]
    r = ref.RefLib()
    cid = r.cell_ids(ms)
    smp = sample_sites(n)
    nk = [int(w) for ln in log.splitlines() if "K-vectors" in ln for w in ln.split() if w.isdigit()]
    np.savez_compressed(os.path.join(GOLD, f"large_{name}.npz"), pe=pe, stress=stress, sample=smp,
                        fsample=force[:, smp], proj=weights(n) @ force.T, fsq=(force ** 2).sum(1),
                        cell_sha256=np.array(hashlib.sha256(np.ascontiguousarray(cid, dtype=np.int32).tobytes()).hexdigest()),
                        cell_sample=cid[smp], n_kvectors=np.array(nk[0] if nk else 0), nsites=np.array(n),
                        ranks=np.array(P), rank_seconds=np.array([p[4] for p in parts]), log=np.array(log))
    print("wrote large_%s.npz  N=%d pe=%s  %d ranks, %.0f s wall (slowest rank %.0f s)"
          % (name, n, pe, P, time.perf_counter() - t0, max(p[4] for p in parts)), flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--procs=")]
    procs = [int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--procs=")]
    names = args or list(workloads())
    P = len(os.sched_getaffinity(0))
    for nm in names:
        make(nm, P, procs[0] if procs else None)
