"""Regenerate the committed fixtures under tests/golden/ (run HERE, where
/root/reference exists; the GPU box only reads the committed files).

1. *_eq.txt : equilibrated molecular configurations written by the UNMODIFIED
   reference program (oracle/_ref/moldy, built by `make -C oracle ref`) with
   `text-mode-save=1`; only the sys-spec + lattice-start part is kept.
2. ref_*.npz : forces / energies / stress / per-site cell ids computed by the
   reference's own force_calc()/ewald()/cellbin() (oracle/_ref/libmoldyref.so)
   for the parity systems of tests/cases.py.  These pin the C restatement
   (oracle/moldy_oracle.c) and, on the GPU box, the CUDA path.

3. ref_rdf.npz : the reference's own RDF histograms (force_calc's RDF pass binning into init_rdf/rdf_accum of
   src/rdf.c, compiled in place) for tests/cases.py RDF_CASES, stored as pair counts (histogram * density,
   rounded: the float store holds count/density summed pair by pair).

4. clay_montmorillonite.txt : the one framework system the reference ships (control.clay): the sys-spec +
   lattice start that follow the control parameters inside src/examples/control.clay, cut out unchanged (input DATA of
   the example run, in the example's own units).  Pins the framework goldens of that output.

usage: python tests/golden/make_fixtures.py [--eq] [--ref] [--rdf] [--clay]
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
MOLDY = os.path.join(ROOT, "oracle", "_ref", "moldy")
EX = "/root/reference/src/examples"

EQ_RUNS = {
    "tip4p_256_eq.txt": ("tip4p.in", dict(**{"surface-dipole": 1, "subcell": 2.5, "nsteps": 2000,
                                             "scale-end": 1500, "scale-interval": 10})),
    "tips2_64_eq.txt": ("tips2.in", dict(**{"surface-dipole": 1, "subcell": 2.5, "nsteps": 2000,
                                            "scale-end": 1500, "scale-interval": 10})),
    "mgcl2_812_eq.txt": ("mgclh2o.in", dict(**{"cutoff": 6.25, "k-cutoff": 3, "alpha": 0.45,
                                               "nsteps": 800, "scale-end": 700, "scale-interval": 5})),
}


def make_eq():
    for out, (spec, kw) in EQ_RUNS.items():
        with tempfile.TemporaryDirectory() as d:
            subprocess.check_call(["cp", os.path.join(EX, spec), d])
            ctl = {"sys-spec-file": spec, "step": 0.0005, "temperature": 300, "density": 1,
                   "print-interval": 1000, "text-mode-save": 1, "save-file": "save.txt",
                   "backup-interval": 0, "time-unit": 4.8888213e-14, **kw}
            with open(os.path.join(d, "control"), "w") as f:
                f.write("".join(f"{k}={v}\n" for k, v in ctl.items()) + "end\n")
            subprocess.check_call([MOLDY, "control"], cwd=d, stdout=subprocess.DEVNULL)
            lines = open(os.path.join(d, "save.txt")).read().splitlines()
            first_end = next(i for i, ln in enumerate(lines) if ln.strip() == "end")
            with open(os.path.join(GOLD, out), "w") as f:
                f.write("\n".join(lines[first_end + 1:]) + "\n")
        print("wrote", out)


def make_ref():
    from oracle import ref
    from tests import cases
    for name, mk in cases.GOLDEN_CASES.items():
        ms = mk()
        r = ref.RefLib()
        o = r.run(ms)
        cid = r.cell_ids(ms)
        np.savez_compressed(os.path.join(GOLD, f"ref_{name}.npz"), force=o["force"], pe=o["pe"],
                            stress=o["stress"], cell=cid, log=np.array(o["log"]))
        print("wrote ref_%s.npz  N=%d pe=%s" % (name, ms.nsites, o["pe"]))


def make_rdf():
    from oracle import ref
    from tests import cases
    out = {}
    for name, (limit, nbins) in cases.RDF_CASES.items():
        ms = cases.GOLDEN_CASES[name]()
        r = ref.RefLib().run(ms, recip=False, rdf=(limit, nbins))
        rho = ms.nsites / float(np.linalg.det(ms.h))
        cnt = r["rdf"].astype(np.float64) * rho
        assert np.abs(cnt - np.rint(cnt)).max() < 0.2, name
        out[name] = np.rint(cnt).astype(np.int64)
        print("rdf %-20s limit=%g nbins=%d pairs=%d" % (name, limit, nbins, out[name].sum()))
    np.savez_compressed(os.path.join(GOLD, "ref_rdf.npz"), **out)


def make_clay():
    lines = open(os.path.join(EX, "control.clay")).read().splitlines()
    a = next(i for i, ln in enumerate(lines) if ln.strip() == "end") + 1        # the sys-spec follows the control parameters
    with open(os.path.join(GOLD, "clay_montmorillonite.txt"), "w") as f:
        f.write("\n".join(lines[a:]) + "\n")
    print("wrote clay_montmorillonite.txt", len(lines) - a, "lines")


if __name__ == "__main__":
    args = sys.argv[1:] or ["--eq", "--ref", "--rdf"]
    if "--clay" in args:
        make_clay()
    if "--rdf" in args:
        make_rdf()
    if "--eq" in args:
        make_eq()
    if "--ref" in args:
        make_ref()
