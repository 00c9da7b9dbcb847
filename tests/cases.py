"""Parity systems shared by the CPU (`-m "not gpu"`) and GPU (`-m gpu`) tests and by
tests/golden/make_fixtures.py.  Every entry returns a fresh MoldySystem; all
are small enough for the reference to evaluate in well under a second, and
together they cover every potential form of src/kernel.c, Coulomb on/off,
strict/lazy cut-off, orthorhombic/triclinic cells, multi-species systems and a
framework species."""
import numpy as np

from moldy_b200 import systems
from moldy_b200.systems import Control, build, KCAL_TIME_UNIT

SPEC_HIW = """
Water 60
1 0 0 0 16 -0.8 O
2 0.7569503 0 -0.5858822 1 0.4 H
2 -0.7569503 0 -0.5858822
Ion 6
3 0 0 0 52 0 X
end
hiw
1 1 -120 800 900000
1 2 30 -10 300
2 2 5 2 50
1 3 -900 2500 5200000
2 3 450 -300 8000
3 3 0 0 9000000
end
"""

SPEC_MORSE = """
Metal 40
1 0 0 0 55.8 1.2 M
Oxide 60
2 0 0 0 16 -0.8 O
end
morse
1 1 0.3 1.8 6.0 20 0 0 0
1 2 0.3 2.9 6.2 30 9.5 2.0 1.9
2 2 0.3 3.6 5.5 70 0 0 0
end
"""

SPEC_MORSE_NEUTRAL = SPEC_MORSE.replace("55.8 1.2 M", "55.8 0 M").replace("16 -0.8 O", "16 0 O")


def hiw():
    return build(SPEC_HIW, Control(subcell=2.2, density=1.0), time_unit=KCAL_TIME_UNIT, seed=11, jitter=0.1)


def morse():
    return build(SPEC_MORSE, Control(subcell=2.0, density=3.2), time_unit=KCAL_TIME_UNIT, seed=12, jitter=0.1)


def morse_nocoul():
    return build(SPEC_MORSE_NEUTRAL, Control(cutoff=7.0, subcell=2.0, density=3.2),
                 time_unit=KCAL_TIME_UNIT, seed=13, jitter=0.1)


def argon_lazy():
    ms = systems.argon(seed=5)
    ms.control.strict_cutoff = 0
    return ms


def quartz_auto():
    """Triclinic Buckingham crystal with all three Ewald parameters derived."""
    return systems.quartz(n=3, pinned_cutoff=False)


def mcy_auto():
    return systems.mgcl2(explicit=False)


def tip4p_tiny_box():
    """A box so small that the stencil reaches the periodic images of the
    reference cell itself (self-image pairs, SURVEY 8a' item 1)."""
    ms = systems.tips2()
    ms.control.cutoff = 10.5
    ms.control.alpha = 0.32
    ms.control.k_cutoff = 2.4
    ms.control.subcell = 3.2
    return ms


def tips2_molpbc():
    """molecular-cutoff=1: whole molecules binned by their centre of mass, sites not wrapped."""
    ms = systems.tips2()
    ms.control.molpbc = 1
    return ms


def tip4p_molpbc_strict():
    """molecular-cutoff=1 with strict-cutoff=1: strict stencil, but no r^2 clamp (src/force.c:951)."""
    ms = systems.tip4p()
    ms.control.molpbc = 1
    ms.control.strict_cutoff = 1
    return ms


def tips2_strict():
    """strict cut-off with Coulomb: the clamped pairs are evaluated at r = 100 rc."""
    ms = systems.tips2()
    ms.control.strict_cutoff = 1
    return ms


GOLDEN_CASES = {
    "argon": systems.argon,
    "argon_lazy": argon_lazy,
    "tip4p": systems.tip4p,
    "tip4p_2": lambda: systems.tip4p(2),
    "tips2": systems.tips2,
    "tips2_tinybox": tip4p_tiny_box,
    "mgcl2": systems.mgcl2,
    "mgcl2_auto": mcy_auto,
    "quartz": systems.quartz,
    "quartz_auto": quartz_auto,
    "slab_framework": systems.slab,
    "clay": systems.clay,
    "hiw": hiw,
    "morse": morse,
    "morse_nocoul": morse_nocoul,
    "tips2_molpbc": tips2_molpbc,
    "tip4p_molpbc_strict": tip4p_molpbc_strict,
    "tips2_strict": tips2_strict,
}

# start-up scalars the reference's own example outputs pin:
# name -> (subcells, neighbour cells (2 x half list), Ewald self energy kJ/mol, k-vectors)
EXAMPLE_GOLDENS = {
    "argon": (729, 912, None, None),            # src/examples/argon-example.out:58-60
    "tip4p": (512, 204, 1722.888659, 1102),     # src/examples/tip4p-example.out:65-68
    "tips2": (125, 94, 968.356089, 570),        # src/examples/tips2-example.out:65-68
    "mgcl2": (3375, 486, 11332.778403, 1484),   # src/examples/mgclh2o-example.out:79-82
    "quartz": (None, None, 457724.399972, 833), # src/examples/quartz-example.out:68-69
    "clay": (960, 1400, 4773.730122, 381),      # src/examples/clay-example.out:75-79 (framework species)
}
CLAY_SHEET_CORRECTION = 201.117                 # "Framework has net electric charge of -8 - correction of 201.117", :77


# the example systems exactly as the reference's start-up builds them (box from the
# density, not from the 6-digit cell printed in a text-mode save file)
EXAMPLE_SYSTEMS = {
    "argon": systems.argon,
    "tip4p": lambda: systems.tip4p(equilibrated=False),
    "tips2": lambda: systems.tips2(equilibrated=False),
    "mgcl2": lambda: systems.build(systems.SPEC_MGCL2, Control(cutoff=6.25, k_cutoff=3.0, alpha=0.45, density=1.0),
                                   time_unit=KCAL_TIME_UNIT),
    "quartz": systems.quartz,
    "clay": systems.clay,
}


def rel_rms(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


# RDF pass of force_calc (src/force.c:1302-1313): case -> (rdf limit, number of bins).  Covers the lazy and
# strict force stencils (the RDF stencil is always strict), a triclinic cell, a framework, molecular
# cut-off binning and a limit larger than half the box (self-image pairs).
RDF_CASES = {
    "argon": (7.7, 40), "tip4p": (8.8, 100), "tips2": (5.5, 55), "tips2_tinybox": (8.0, 64), "mgcl2": (8.4, 100),
    "quartz": (9.0, 90), "slab_framework": (6.3, 63), "tips2_molpbc": (5.5, 50), "tip4p_molpbc_strict": (6.0, 30),
}


# BASELINE.json configs[1..4] at full size: checked on the GPU against reduced records of the compiled
# reference (tests/golden/large_*.npz, written by tests/golden/make_large_fixtures.py).
LARGE_CASES = {
    "tip4p_5": lambda: systems.tip4p(5),                               # configs[1]: 128 000 sites
    "mgcl2_7": lambda: systems.mgcl2(7, explicit=False),               # configs[2]: 278 516 sites, MCY
    "quartz_48": lambda: systems.quartz(48, pinned_cutoff=False),      # configs[3]: 995 328 ions, Buckingham, triclinic
    "tip4p_10": lambda: systems.tip4p(10),                             # configs[4] / bench workload: 1 024 000 sites
    "tip4p_13": lambda: systems.tip4p(13),                             # configs[4] sweep: 2 249 728 sites
    "tip4p_16": lambda: systems.tip4p(16),                             # configs[4] sweep: 4 194 304 sites
}
