"""ctypes mirror of the Moldy structs that cross the force_calc()/ewald() boundary.

These are *interface* declarations: field order and C types follow the
reference headers exactly (LP64, natural alignment) because the drop-in library
`libmoldy_b200.so` receives pointers to the caller's own structs.

  contr_mt   src/structs.h:27-85     global `control` (read: cutoff, subcell, alpha,
                                     k_cutoff, strict_cutoff, molpbc, surface_dipole ...)
  system_mt  src/structs.h:87-118    nsites, nmols, nspecies, max_id, ptype, n_potpar, h, c_of_m
  spec_mt    src/structs.h:121-145   nsites, nmols, framework, site_id, p_f_sites
  site_mt    src/structs.h:147-154
  pot_mt     src/structs.h:156-161   flag, pad, p[NPOTP]
  sizes      src/defs.h:126-145      NPOTP=8, L_name=128, L_spec=32, L_site=8, NPE=2

The C side of the same declarations is include/moldy_b200.h; tests check that
sizeof() agrees between the two and with the reference build (oracle/_ref).
"""
import ctypes as C

NPOTP = 8
L_NAME = 128
L_SPEC = 32
L_SITE = 8
NPE = 2
NCACHE = 256          # src/defs.h:119
NLINE = 4             # src/defs.h:122

real = C.c_double
vec_mt = real * 3
quat_mt = real * 4
mat_mt = vec_mt * 3


class contr_mt(C.Structure):
    _fields_ = [
        ("title", C.c_char * L_NAME),
        ("istep", C.c_long), ("nsteps", C.c_long),
        ("step", C.c_double),
        ("print_sysdef", C.c_int), ("new_sysdef", C.c_int),
        ("molpbc", C.c_int), ("reset_averages", C.c_int),
        ("scale_options", C.c_int),
        ("surface_dipole", C.c_int), ("lattice_start", C.c_int),
        ("sysdef", C.c_char * L_NAME), ("restart_file", C.c_char * L_NAME),
        ("save_file", C.c_char * L_NAME), ("dump_file", C.c_char * L_NAME),
        ("backup_file", C.c_char * L_NAME), ("temp_file", C.c_char * L_NAME),
        ("spare", C.c_int * 20),
        ("nosymmetric_rot", C.c_int),
        ("ewald_accuracy", C.c_double),
        ("ttmass", C.c_double), ("rtmass", C.c_double),
        ("const_pressure", C.c_int), ("const_temp", C.c_int),
        ("xdr_write", C.c_int), ("strict_cutoff", C.c_int),
        ("strain_mask", C.c_int), ("nbins", C.c_int),
        ("seed", C.c_ulong),
        ("page_width", C.c_int), ("page_length", C.c_int),
        ("scale_interval", C.c_long), ("scale_end", C.c_long),
        ("begin_average", C.c_long), ("average_interval", C.c_long),
        ("begin_dump", C.c_long), ("dump_offset", C.c_long),
        ("dump_interval", C.c_long),
        ("dump_level", C.c_int), ("maxdumps", C.c_int),
        ("backup_interval", C.c_long), ("roll_interval", C.c_long),
        ("print_interval", C.c_long), ("begin_rdf", C.c_long),
        ("rdf_interval", C.c_long), ("rdf_out", C.c_long),
        ("temp", C.c_double), ("pressure", C.c_double), ("pmass", C.c_double),
        ("cutoff", C.c_double), ("subcell", C.c_double), ("density", C.c_double),
        ("alpha", C.c_double), ("k_cutoff", C.c_double), ("limit", C.c_double),
        ("cpu_limit", C.c_double),
    ]


class system_mt(C.Structure):
    _fields_ = [
        ("nsites", C.c_int), ("nmols", C.c_int), ("nmols_r", C.c_int),
        ("nspecies", C.c_int), ("max_id", C.c_int), ("d_of_f", C.c_int),
        ("ptype", C.c_int), ("n_potpar", C.c_int),
        ("c_of_m", C.POINTER(vec_mt)), ("mom", C.POINTER(vec_mt)),
        ("momp", C.POINTER(vec_mt)),
        ("quat", C.POINTER(quat_mt)), ("amom", C.POINTER(quat_mt)),
        ("amomp", C.POINTER(quat_mt)),
        ("h", C.POINTER(vec_mt)), ("hmom", C.POINTER(vec_mt)),
        ("hmomp", C.POINTER(vec_mt)),
        ("ts", real), ("tsmom", real), ("H_0", real), ("rs", real), ("rsmom", real),
    ]


class spec_mt(C.Structure):
    _fields_ = [
        ("inertia", real * 3), ("mass", real), ("dipole", real), ("charge", real),
        ("nsites", C.c_int), ("nmols", C.c_int),
        ("rdof", C.c_int), ("framework", C.c_int),
        ("name", C.c_char * L_SPEC),
        ("site_id", C.POINTER(C.c_int)),
        ("p_f_sites", C.POINTER(vec_mt)),
        ("c_of_m", C.POINTER(vec_mt)), ("mom", C.POINTER(vec_mt)),
        ("momp", C.POINTER(vec_mt)),
        ("quat", C.POINTER(quat_mt)), ("amom", C.POINTER(quat_mt)),
        ("amomp", C.POINTER(quat_mt)),
    ]


class site_mt(C.Structure):
    _fields_ = [
        ("mass", C.c_double), ("charge", C.c_double),
        ("name", C.c_char * L_SITE),
        ("flag", C.c_int), ("pad", C.c_int),
    ]


class pot_mt(C.Structure):
    _fields_ = [("flag", C.c_int), ("pad", C.c_int), ("p", real * NPOTP)]


class pots_mt(C.Structure):
    _fields_ = [("name", C.c_char_p), ("npar", C.c_int)]


class dim_mt(C.Structure):
    _fields_ = [("m", C.c_int), ("l", C.c_int), ("t", C.c_int), ("q", C.c_int)]


class mdb_species(C.Structure):
    _fields_ = [("nmols", C.c_int), ("nsites", C.c_int), ("framework", C.c_int), ("rotates", C.c_int), ("rdof", C.c_int)]


class mdb_species_dyn(C.Structure):
    _fields_ = [("mass", C.c_double), ("inertia", C.c_double * 3)]


def nsarray(nsites: int) -> int:
    """Row stride of the caller's site[3][nsarray] block (src/accel.c:422)."""
    return ((nsites - 1) | (NCACHE - 1)) + 1 + NLINE
