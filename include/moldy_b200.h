/*
 * moldy_b200.h -- C ABI of libmoldy_b200.so, the B200-native replacement for
 * Moldy's force-evaluation hot path (force.c / kernel.c / ewald.c).
 *
 * Two layers, both plain C (pointers and sizes only, no torch/C++ types):
 *
 *  (A) Moldy's own symbols.  Linking libmoldy_b200.so in place of force.o,
 *      kernel.o and ewald.o (moldy_SOURCES, src/Makefile.am:17) leaves the rest
 *      of Moldy unchanged.  Each prototype cites the definition it replaces.
 *      One level up, eval_forces() (src/accel.c:398) is exported as well: with
 *      it the sites and site forces stay in HBM and only molecular data cross
 *      PCIe (declared after the mdb_ building blocks it is made of).
 *  (B) mdb_* : the device-resident engine underneath (A), for callers that keep
 *      positions and forces in HBM (bench.py `value`, multi-GPU ranks that
 *      all-reduce the packed result with NCCL).
 *
 * Struct layouts below MUST stay byte-identical to src/structs.h / src/defs.h of
 * the reference (LP64); tests/test_abi.py checks sizeof() against the Python
 * mirror (moldy_b200/abi.py) and against the reference build.
 */
#ifndef MOLDY_B200_H
#define MOLDY_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ types -- */
#define MDB_NPOTP   8      /* src/defs.h:126   NPOTP                           */
#define MDB_L_NAME  128    /* src/defs.h:139   L_name                          */
#define MDB_L_SPEC  32     /* src/defs.h:140   L_spec                          */
#define MDB_L_SITE  8      /* src/defs.h:141   L_site                          */
#define MDB_NPE     2      /* src/defs.h:144   NPE: real-space & Ewald PE      */

typedef double real;                 /* src/defs.h:259 */
typedef int    boolean;              /* src/defs.h:260 */
typedef real   vec_mt[3];            /* src/defs.h:266 */
typedef vec_mt *vec_mp;
typedef real   quat_mt[4];
typedef quat_mt *quat_mp;
typedef real   mat_mt[3][3];
typedef vec_mt *mat_mp;

/* src/structs.h:27-85.  The hot path reads: cutoff, subcell, alpha, k_cutoff,
 * strict_cutoff, molpbc, surface_dipole(not here: accel.c), rdf_interval,
 * begin_rdf, istep, limit. */
typedef struct {
   char     title[MDB_L_NAME];
   long     istep, nsteps;
   double   step;
   boolean  print_sysdef, new_sysdef, molpbc, reset_averages;
   int      scale_options;
   boolean  surface_dipole, lattice_start;
   char     sysdef[MDB_L_NAME], restart_file[MDB_L_NAME], save_file[MDB_L_NAME],
            dump_file[MDB_L_NAME], backup_file[MDB_L_NAME], temp_file[MDB_L_NAME];
   int      spare[20];
   boolean  nosymmetric_rot;
   double   ewald_accuracy;
   double   ttmass, rtmass;
   int      const_pressure;
   int      const_temp;
   boolean  xdr_write, strict_cutoff;
   int      strain_mask;
   int      nbins;
   unsigned long seed;
   int      page_width, page_length;
   long     scale_interval, scale_end, begin_average, average_interval,
            begin_dump, dump_offset, dump_interval;
   int      dump_level, maxdumps;
   long     backup_interval, roll_interval, print_interval, begin_rdf,
            rdf_interval, rdf_out;
   double   temp, pressure, pmass, cutoff, subcell, density, alpha, k_cutoff,
            limit, cpu_limit;
} contr_mt, *contr_mp;

/* src/structs.h:87-118 */
typedef struct {
   int      nsites, nmols, nmols_r, nspecies, max_id, d_of_f;
   int      ptype, n_potpar;
   vec_mp   c_of_m, mom, momp;
   quat_mp  quat, amom, amomp;
   mat_mp   h, hmom, hmomp;
   real     ts, tsmom;
   real     H_0;
   real     rs, rsmom;
} system_mt, *system_mp;

/* src/structs.h:121-145 */
typedef struct {
   real     inertia[3], mass, dipole, charge;
   int      nsites, nmols;
   int      rdof, framework;
   char     name[MDB_L_SPEC];
   int      *site_id;
   vec_mp   p_f_sites;
   vec_mp   c_of_m, mom, momp;
   quat_mp  quat, amom, amomp;
} spec_mt, *spec_mp;

/* src/structs.h:147-154 */
typedef struct {
   double   mass, charge;
   char     name[MDB_L_SITE];
   int      flag;
   int      pad;
} site_mt, *site_mp;

/* src/structs.h:156-161 */
typedef struct {
   int      flag;
   int      pad;
   real     p[MDB_NPOTP];
} pot_mt, *pot_mp;

/* src/structs.h:163-184 */
typedef struct { char *name; int npar; } pots_mt;
typedef struct { int m, l, t, q; } dim_mt, *dim_mp;

/* ------------------------------------------- (A) Moldy's own entry points -- */

/* Real-space link-cell site-site forces.  Replaces src/force.c:1108-1320
 * (force_calc) and everything it calls (fill_cells, neighbour_list,
 * site_neighbour_list, force_inner, kernel).  Accumulates (+=) into
 * site_force[3][>=nsites], *pe and the upper triangle of stress; reads the
 * globals `control`, `ithread`, `nthreads` (src/main.c:83-84). */
void force_calc(real **site, real **site_force, system_mt *system, spec_mt *species,
                real *chg, pot_mt *potpar, double *pe, mat_mt stress);

/* Reciprocal-space Ewald sum.  Replaces src/ewald.c:280-589.  `pe` is the
 * caller's pe+1 (src/accel.c:526). */
void ewald(real **site, real **site_force, system_mp system, spec_mt *species,
           real *chg, double *pe, real (*stress)[3]);

/* Vector pair-potential evaluation, src/kernel.c:157-463: for j in [jmin,nnab)
 * forceij[j] = -phi'(r_j)/r_j and *pe += sum phi(r_j).  Runs on the GPU (same
 * device function as the pair kernel). */
void kernel(int jmin, int nnab, real *forceij, double *pe, real *r_sqr, real *nab_chg,
            double chg, double norm, double alpha, int ptype, real **pot);

/* Single-pair potential energy, src/force.c:632-650. */
double poteval(real *potpar, double r, int ptype, double chgsq);

/* -int_{rc}^{inf} r^2 U(r) dr, src/kernel.c:103-151. */
double dist_pot(real *potpar, double cutoff, int ptype);

/* Potential name table / parameter dimensions, src/kernel.c:61-84 (read by
 * input.c, output.c, convert.c). */
extern const pots_mt potspec[];
extern const dim_mt  pot_dim[][MDB_NPOTP];

/* --------------------------------------- (B) device-resident engine (mdb_) -- */

typedef struct mdb_engine mdb_engine;

/* Flat description of what force_calc()/ewald() extract from system_mt/spec_mt/
 * pot_mt/control.  All arrays are HOST pointers, copied by mdb_configure. */
typedef struct {
   int          nsites;        /* system->nsites                                    */
   int          nsites_xf;     /* non-framework sites (frameworks are sorted last)  */
   int          max_id;        /* system->max_id                                    */
   int          ptype;         /* system->ptype, index into potspec[]               */
   int          n_potpar;      /* system->n_potpar                                  */
   const int    *site_type;    /* [nsites] site id, src/force.c:1185-1191           */
   const int    *site_mol;     /* [nsites] molecule index (TOO_CLOSE test) or NULL  */
   const double *chg;          /* [nsites]                                          */
   const double *potpar;       /* [max_id*max_id*MDB_NPOTP], row i*max_id+j         */
   double       h[9];          /* MD cell matrix, row major                         */
   double       cutoff, subcell, alpha, k_cutoff;
   int          strict_cutoff;
   int          do_recip;      /* build k-space tables (alpha > ALPHAMIN)           */
   int          molpbc;        /* control.molpbc: bin whole molecules by their c-of-m (src/force.c:456-484) */
   int          nmols;         /* molecules (rows of the c-of-m array), needed when molpbc                  */
} mdb_config;

/* Result block in HBM: [fx(N) | fy(N) | fz(N) | pe_real, pe_recip | stress[9] | pad]
 * (stress row-major, upper triangle only, as the reference writes it). */
#define MDB_OUT_SCALARS 16
size_t      mdb_out_doubles(int nsites);

mdb_engine *mdb_create(int device);
void        mdb_destroy(mdb_engine *e);
const char *mdb_last_error(void);

/* (Re)build everything that depends on the system definition, the cell matrix
 * and the control cut-offs: link-cell grid, neighbour stencil, k-vector tables.
 * Returns 0, or -1 with mdb_last_error() set. */
int  mdb_configure(mdb_engine *e, const mdb_config *cfg);

/* Work partition of this process, the reference's ithread/nthreads
 * (src/force.c:856, src/ewald.c:495-496). */
void mdb_set_partition(mdb_engine *e, int ithread, int nthreads);

/* Real-space kernel variant: 2 = one thread per site, full stencil; 3 = tiled (warp = batch of
 * sites, lanes = neighbours), full stencil, bit-reproducible; 4 = tiled with Newton's third law
 * (half stencil, red.global.add.f64 force accumulation; last-bit run-to-run variation).
 * Default: environment MDB_PAIR_MODE, else 4 (fastest); use 3 for bit-reproducible runs. */
void mdb_set_pair_mode(mdb_engine *e, int mode);

/* Positions: three HOST rows of nsites doubles (copied H2D on `stream`), or
 * three DEVICE rows already resident in HBM. */
int  mdb_set_sites_host(mdb_engine *e, const double *x, const double *y, const double *z, void *stream);
int  mdb_set_sites_device(mdb_engine *e, const double *dx, const double *dy, const double *dz, void *stream);

/* molecular-cutoff mode only: the scaled centre-of-mass co-ordinates system->c_of_m[nmols][3]
 * (HOST pointer, copied on `stream`); call before mdb_build_cells every step. */
int  mdb_set_com_host(mdb_engine *e, const double *c_of_m, void *stream);

/* Hot path.  `d_out` is a DEVICE buffer of mdb_out_doubles(nsites) doubles that
 * is accumulated into (+=); zero it with mdb_zero_out first.  Everything is
 * enqueued on `stream` (a cudaStream_t); nothing synchronises. */
int  mdb_zero_out(mdb_engine *e, double *d_out, void *stream);
int  mdb_build_cells(mdb_engine *e, void *stream);
int  mdb_force_real(mdb_engine *e, double *d_out, void *stream);
int  mdb_force_recip(mdb_engine *e, double *d_out, void *stream);
/* k-space forces in `n` slices of the charged sites (1..8; default 1), an event after each, so that a host-facing caller can
 * bring a slice home while the next one is computed (force_calc()/ewald() of layer (A) do).  After mdb_force_recip:
 * mdb_kforce_slices fills the events (cudaEvent_t, recorded on the launching stream) and the exclusive upper bounds of the
 * original site indices whose k-space forces are complete at each event; returns the number of slices (0: not cut). */
void mdb_set_kforce_slices(mdb_engine *e, int n);
int  mdb_kforce_slices(const mdb_engine *e, void **events, int *site_hi);

/* Both sums of one step on two streams.  Default (mdb_set_overlap(e, 0, 0)): the k-space kernels go first on a side stream
 * while the cell build and the sub-list compaction (launch-latency-bound) are enqueued behind them; the pair passes follow.
 * With a filler grid (`fill_blocks` > 0 blocks of `fill_threads` threads of the pair kernel, persistent, drawing batches
 * from a counter until the k-space chain has ended) the pair kernel works beside the k-space GEMMs on the same SMs and the
 * rest of the real-space pass runs at full occupancy afterwards (measured: +1.4 % at 296 x 128, conserved time otherwise --
 * DFMA and DMMA share the FP64 pipe).  Same sums as mdb_force_real + mdb_force_recip (the k-space block is added last).
 * mdb_set_overlap(e, -1, 0): one stream, real space then k-space.  MDB_OVERLAP=-1 | 0 | blocks,threads sets the default. */
int  mdb_set_overlap(mdb_engine *e, int fill_blocks, int fill_threads);
int  mdb_force_both(mdb_engine *e, double *d_out, void *stream);
long mdb_overlap_filled(mdb_engine *e);
/* Exponential potentials (Buckingham, generic, Morse, MCY): stencil runs whose every cell pair is further apart than 52 decay
 * lengths (exp(-r/rho) < 2.6e-23 of its amplitude for every site-type pair) are evaluated without the exponentials.  Default
 * on (MDB_PAIR_FAR=0 or mdb_set_pair_far(e, 0) before mdb_configure: every run with the full potential). */
void mdb_set_pair_far(mdb_engine *e, int on);
/* host only: that distance for a potential table (max_id x max_id rows of MDB_NPOTP parameters as pot_mt.p holds them) and,
 * if rest != NULL, the power-law rest of the potential as rows of the p0/r^4 + p1/r^6 + p2/r^12 form; 0 = not applicable */
double mdb_far_radius(int ptype, int max_id, const double *potpar, double *rest);
int  mdb_pair_far_runs(const mdb_engine *e);   /* batches the filler drew in the last call (diagnostic; synchronises) */

/* k-space cut by SITES instead of by (h,k) columns (multi-GPU, moldy_b200/spmd.py): pass 1 writes
 * this rank's structure-factor sums (mdb_recip_sum_doubles() doubles) to the DEVICE buffer d_psum,
 * the caller all-reduces d_psum over the ranks, pass 2 adds energy/stress (rank 0) and the forces
 * on this rank's sites.  With one rank the pair equals mdb_force_recip. */
size_t mdb_recip_sum_doubles(const mdb_engine *e);
int  mdb_recip_partial(mdb_engine *e, double *d_psum, void *stream);
int  mdb_recip_finish(mdb_engine *e, const double *d_psum, double *d_out, void *stream);

/* RDF binning pass of force_calc (src/force.c:1302-1313: rdf_inner over the strict neighbour list of
 * radius `limit`, rdf_accum src/rdf.c:94-108).  h_counts[pair][bin] += number of site pairs (this
 * rank's share) with bin = (int)(nbins/limit * r) < nbins; pair = (idi <= idj), idi = 1..max_id-1, in
 * the order init_rdf lays the histograms out (src/rdf.c:82-90); mdb_rdf_size() entries.  force_calc()
 * of layer (A) adds count/density to the host program's array (rdf_ptr(), src/rdf.c:60-64). */
size_t mdb_rdf_size(const mdb_engine *e, int nbins);
int  mdb_rdf_counts(mdb_engine *e, double limit, int nbins, unsigned long long *h_counts, void *stream);
void mdb_rdf_private_resize(int n);      /* host is not Moldy: size (and clear) the fall-back store rdf_ptr() returns */

/* Molecular-frame steps of eval_forces() around force_calc()/ewald() (src/accel.c:497-560), as
 * device-level building blocks; one call per species, all pointers are DEVICE pointers.
 * mdb_make_sites: make_sites() (src/algorith.c:169-217) into the engine's own position rows [site_offset,
 *   site_offset + nmols*nsites): site = h.com_s + R(quat).p_f_site, brought into the cell site by site when sitepbc != 0
 *   (SITEPBC), whole molecules otherwise (MOLPBC); d_quat NULL for species without rotational freedom.  Bit-identical
 *   to the reference's sites (the cell assignment depends on them).  The engine then uses its own rows as the sites.
 * mdb_mol_forces: mol_force() + mol_torque() (src/algorith.c:111-163) of one species from the site forces of a result
 *   block: d_force[nmols][3], d_torque[nmols][3] (NULL: no torques).
 * mdb_get_sites: the engine's current sites to host rows (synchronises). */
int  mdb_make_sites(mdb_engine *e, const double h[9], const double *d_com_s, const double *d_quat, const double *d_pfs,
                    int nmols, int nsites, int site_offset, int sitepbc, void *stream);
int  mdb_mol_forces(mdb_engine *e, const double *d_out, const double *d_quat, const double *d_pfs, int nmols, int nsites,
                    int site_offset, double *d_force, double *d_torque, void *stream);
int  mdb_get_sites(mdb_engine *e, double *hx, double *hy, double *hz, void *stream);

/* The whole of eval_forces() (src/accel.c:398-617) on the device: only the scaled centres of mass
 * and quaternions go in and molecular forces, torques and 23 scalars come out; site positions and site forces never
 * leave HBM.  Sequence, all on `stream`: make_sites per species (SITEPBC, or MOLPBC under molecular-cutoff) -> cell
 * build -> real-space sum -> reciprocal-space sum (do_recip) -> make_sites again (MOLPBC, frameworks SITEPBC;
 * src/accel.c:537-542) -> dipole moment sum chg*site (:550) -> surface-dipole force term (:551-558) folded into ->
 * mol_force/mol_torque (:564-571) and the site->molecular virial sum_sites f_i (r_j - R_j) (:576-601, evaluated per
 * molecule, which is the same sum without the reference's cancellation of two O(N L) terms).
 * mdb_set_species: the species table in system order (frameworks last) and every species' principal-frame sites,
 *   concatenated [sum nsites][3] (HOST).  Call after mdb_configure.
 * mdb_eval_forces_host: com[ispec] = spec->c_of_m (HOST [nmols][3] scaled), quat[ispec] = spec->quat (HOST [nmols][4])
 *   or NULL for a species whose sites are not rotated.  h_result (HOST, mdb_eval_result_doubles() doubles) =
 *   [force 3*nmols | torque 3*nmols_r (species with rdof > 0, in order) | MDB_EVAL_SCALARS scalars]; scalars:
 *   [0..2] dipole moment, [3..11] virial correction V[i][j] (subtract from the symmetrised stress), [12] real-space
 *   energy, [13] reciprocal-space energy (without the surface term), [14..22] stress[3][3] as force_calc+ewald leave it
 *   (upper triangle).  The constants of the first call (intramolecular, self and sheet energies, distant-potential
 *   terms) and the surface-dipole energy are the host program's: eval_forces() of layer (A) adds them.
 *   Synchronises `stream`. */
typedef struct {
   int nmols, nsites;      /* spec->nmols, spec->nsites                          */
   int framework;          /* spec->framework                                    */
   int rotates;            /* spec->quat != NULL: make_sites rotates the sites   */
   int rdof;               /* spec->rdof: torques are returned when > 0          */
} mdb_species;
#define MDB_EVAL_SCALARS 32
int    mdb_set_species(mdb_engine *e, int nspecies, const mdb_species *sp, const double *pfs);
size_t mdb_eval_result_doubles(const mdb_engine *e);
int    mdb_eval_forces_host(mdb_engine *e, const double h[9], const double *const *com, const double *const *quat,
                            int surface_dipole, int do_recip, double *h_result, void *stream);
/* One-shot: the next mdb_eval_forces_host also runs the RDF pass (mdb_rdf_counts) on its cell lists, between the force
 * sums and the second make_sites, into h_counts (HOST, mdb_rdf_size() entries). */
void   mdb_eval_request_rdf(mdb_engine *e, double limit, int nbins, unsigned long long *h_counts);
/* eval_forces() under a name of the library's own (for a trampoline object). */
void   mdb_eval_forces_moldy(system_mp sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, double *pe, real *dip_mom,
                             mat_mt stress, vec_mp *force, vec_mp *torque);
/* The library's own pinned copy of the last result (valid until the next call); pass h_result = NULL above to skip
 * the copy into caller memory. */
const double *mdb_eval_result(const mdb_engine *e);

/* ---- the NVE leapfrog step of do_step() (src/accel.c:626-827) on a state resident in HBM (mdb_md.cu) ----
 * Scaled centres of mass, quaternions, linear and angular momenta live on the device between outputs; a step is
 * leapf_all_coords(step/2) -> eval_forces -> leapf_all_momenta(step/2) x 2 -> framework momenta = 0 -> leapf_all_coords(step/2)
 * and returns scalars only: [MDB_EVAL_SCALARS as mdb_eval_forces_host | per species MDB_MD_SUMS sums at the end of the step |
 * per species MDB_MD_SUMS sums at the half step (half_sums != 0; for H_0, src/accel.c:718-726) | count of quaternions whose
 * norm was off by > 1e-4 (the reference: FATAL)].  Sums of a species: [sum p_i p_j (xx xy xz yy yz zz), p = (h^-1)' mom
 * (trans_ke / energy_dyad, src/algorith.c:221-284) | sum amom_k^2 (rot_ke :244-257) | sum F_k^2 | sum T_k^2 (mean_square :101)].
 * Call order: mdb_configure, mdb_set_species, mdb_md_set_dynamics, mdb_md_upload_state, then mdb_md_step per step. */
typedef struct { double mass, inertia[3]; } mdb_species_dyn;          /* spec->mass, spec->inertia */
#define MDB_MD_SUMS 15
int    mdb_md_set_dynamics(mdb_engine *e, const mdb_species_dyn *dyn, int nosymmetric_rot);
size_t mdb_md_scalars(const mdb_engine *e);
int    mdb_md_upload_state(mdb_engine *e, const double *const *com, const double *const *quat, const double *const *mom,
                           const double *const *amom, void *stream);
int    mdb_md_download_state(mdb_engine *e, double *const *com, double *const *quat, double *const *mom, double *const *amom,
                             double *const *force, double *const *torque, void *stream);       /* NULL entries are skipped */
int    mdb_md_step(mdb_engine *e, const double h[9], double step, double ts, int surface_dipole, int do_recip, int half_sums,
                   double *h_scal, void *stream);
const double *mdb_md_result(const mdb_engine *e);
/* the sub-steps, for tests and for hosts that interleave their own work */
int    mdb_md_coords(mdb_engine *e, const double h[9], double step, double ts, void *stream);
int    mdb_md_momenta(mdb_engine *e, const double h[9], double step, void *stream);
int    mdb_md_eval_forces(mdb_engine *e, const double h[9], int surface_dipole, int do_recip, void *stream);
int    mdb_md_sums_now(mdb_engine *e, const double h[9], double *h_sums, void *stream);

/* Moldy's eval_forces() itself (src/accel.c:398-617) on top of mdb_eval_forces_host: same prototype, same outputs
 * (pe[NPE], dip_mom[3], the full symmetric virial stress, force[ispec][imol], torque[ispec][imol]), same first-call
 * notes.  Linking it in place of accel.c's definition is described in INTEGRATION.md section 5. */
void eval_forces(system_mp sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, double *pe, real *dip_mom,
                 mat_mt stress, vec_mp *force, vec_mp *torque);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU: the reference's replicated-data scheme (every rank holds all sites; par_rsum/par_dsum of the partial force,
 * stress and energy arrays, src/parallel.c:549-588, call sites src/accel.c:531-535) with the sums done by kernels of the
 * library's own over NVLink peer memory (mdb_peer.cu).  A peer = one rank = one engine on one GPU plus a window of device
 * memory the other ranks map: engines of ONE process (mdb_peer_connect) or one process per GPU (mdb_peer_handle, exchange
 * the 64-byte handles, mdb_peer_open).  Rank r owns the r-th slice of the site batches (real space), of the charged sites
 * (k-space, the manual's RIL scheme src/moldy.tex:3441-3466) and of the result (sites [bounds[r], bounds[r+1])).
 * Step = phase A (cells, real-space sum, structure-factor partial sums) | barrier | phase B (structure-factor all-reduce,
 * energy/stress on rank 0, k-space forces on own sites) | barrier | phase C (reduce-scatter of the forces + the scalars
 * on every rank) [| barrier | phase D (all-gather of the forces)].  One host thread may drive all ranks of a process if it
 * enqueues each phase on every rank before the next one; separate processes call mdb_peer_step.  Nothing synchronises
 * unless stated.  `what`: bit 0 real space, bit 1 reciprocal space (phase A also: bit 2 = continue the phase an earlier
 * call began, i.e. no new result block and no cell build). */
#define MDB_MAX_PEERS 16
#define MDB_PEER_HANDLE_BYTES 64
typedef struct mdb_peer mdb_peer;
mdb_peer *mdb_peer_create(mdb_engine *e, int rank, int world);       /* after mdb_configure; sets the engine's partition and sites */
void      mdb_peer_destroy(mdb_peer *p);
size_t    mdb_peer_window_bytes(const mdb_peer *p);
int       mdb_peer_handle(mdb_peer *p, void *handle);                 /* MDB_PEER_HANDLE_BYTES bytes (CUDA IPC)                  */
int       mdb_peer_open(mdb_peer *p, const void *handles);            /* world * MDB_PEER_HANDLE_BYTES bytes, rank order         */
int       mdb_peer_connect(mdb_peer *const *peers, int world);        /* same process: all peers, rank order                     */
int       mdb_peer_set_site_bounds(mdb_peer *p, const long long *bounds);   /* world + 1 values; default nsites r / world        */
void      mdb_peer_slice(const mdb_peer *p, long long lohi[2]);
int       mdb_peer_barrier(mdb_peer *p, void *stream);
int       mdb_peer_error(mdb_peer *p, void *stream);                  /* 1: a barrier timed out (a rank died); synchronises      */
/* inputs: the rank uploads its own slice over its own PCIe link, the rest comes from the peers after a barrier */
int       mdb_peer_sites_host_slice(mdb_peer *p, const double *x, const double *y, const double *z, void *stream);
int       mdb_peer_sites_host_all(mdb_peer *p, const double *x, const double *y, const double *z, void *stream);
int       mdb_peer_sites_gather(mdb_peer *p, void *stream);
double   *mdb_peer_sites(mdb_peer *p);                                /* DEVICE [3][nsites]: the engine's site rows              */
double   *mdb_peer_in(mdb_peer *p);                                   /* DEVICE: generic input block (c-of-m | quaternions)      */
int       mdb_peer_in_host_slice(mdb_peer *p, const double *h_in, size_t len, void *stream);
int       mdb_peer_in_gather(mdb_peer *p, size_t len, void *stream);
int       mdb_peer_in_gather_bounds(mdb_peer *p, size_t off_doubles, long long row_len, const long long *bounds, void *stream);
/* the step */
int       mdb_peer_phase_a(mdb_peer *p, int what, void *stream);
int       mdb_peer_phase_b(mdb_peer *p, int what, void *stream);
int       mdb_peer_phase_c(mdb_peer *p, void *stream);
int       mdb_peer_phase_d(mdb_peer *p, void *stream);
int       mdb_peer_step(mdb_peer *p, int what, int gather, void *stream);   /* A | barrier | B | barrier | C [| barrier | D]     */
double   *mdb_peer_result(mdb_peer *p);    /* DEVICE result block: own slice + scalars after C, all forces after D               */
double   *mdb_peer_partial(mdb_peer *p);   /* DEVICE: this rank's partial block of the current step                              */
int       mdb_peer_read_slice_host(mdb_peer *p, double *fx, double *fy, double *fz, double *scal16, void *stream);
long      mdb_peer_barriers(const mdb_peer *p);

/* Several GPUs behind ONE process (mdb_group.cu): P engines + P peers driven by the calling thread; what force_calc()/
 * ewald()/eval_forces() of layer (A) use when MOLDY_B200_DEVICES names more than one device -- the unmodified Moldy program
 * (no -DSPMD, nthreads = 1: SURVEY 8b) then runs on all of them.  Both calls block until the results are in host memory.
 * mdb_group_force_host: forces of the HOST site rows summed over the ranks, WRITTEN into three host rows (each rank copies
 *   its slice in and out over its own PCIe link), scal16 = [pe_real, pe_recip, stress[9], pad]; c_of_m as mdb_set_com_host
 *   (molecular cut-off) or NULL; tc/tc_pair as mdb_too_close (may be NULL).
 * mdb_group_eval_forces_host: contract and result layout of mdb_eval_forces_host; rdf_counts != NULL adds the RDF pass. */
typedef struct mdb_group mdb_group;
mdb_group  *mdb_group_create(int ndev, const int *devices);
void        mdb_group_destroy(mdb_group *g);
int         mdb_group_size(const mdb_group *g);
mdb_engine *mdb_group_engine(mdb_group *g, int rank);
void       *mdb_group_stream(mdb_group *g, int rank);
int         mdb_group_configure(mdb_group *g, const mdb_config *cfg);
int         mdb_group_force_host(mdb_group *g, const double *x, const double *y, const double *z, const double *c_of_m, int what,
                                 double *fx, double *fy, double *fz, double *scal16, int *tc, int tc_pair[2]);
int         mdb_group_set_species(mdb_group *g, int nspecies, const mdb_species *sp, const double *pfs);
size_t      mdb_group_eval_result_doubles(const mdb_group *g);
const double *mdb_group_eval_result(const mdb_group *g);
int         mdb_group_eval_forces_host(mdb_group *g, const double h[9], const double *const *com, const double *const *quat,
                                       int surface_dipole, int do_recip, double *h_result, double rdf_limit, int rdf_nbins,
                                       unsigned long long *rdf_counts, int *tc, int tc_pair[2]);
/* The resident NVE step (mdb_md_step) on a device group: every rank moves its share of the molecules, the [c-of-m |
 * quaternions] block is all-gathered over NVLink after the first half step of the co-ordinates, the sums are added on the
 * host.  Call order: mdb_group_configure, mdb_group_set_species, mdb_group_md_set_dynamics, mdb_group_md_upload_state,
 * then mdb_group_md_step per step (scalars: layout of mdb_md_step; rdf_counts as mdb_group_eval_forces_host). */
int         mdb_group_too_close(mdb_group *g, int tc_pair[2]);          /* as mdb_too_close, over all ranks */
int         mdb_group_md_set_dynamics(mdb_group *g, const mdb_species_dyn *dyn, int nosymmetric_rot);
int         mdb_group_md_upload_state(mdb_group *g, const double *const *com, const double *const *quat, const double *const *mom,
                                      const double *const *amom);
int         mdb_group_md_download_state(mdb_group *g, double *const *com, double *const *quat, double *const *mom, double *const *amom,
                                        double *const *force, double *const *torque);
size_t      mdb_group_md_scalars(const mdb_group *g);
const double *mdb_group_md_result(const mdb_group *g);
int         mdb_group_md_step(mdb_group *g, const double h[9], double step, double ts, int surface_dipole, int do_recip,
                              int half_sums, double *h_scal, double rdf_limit, int rdf_nbins, unsigned long long *rdf_counts);

/* Number of values in which three HOST rows differ (bit for bit) from the sites the engine currently holds; the rows are
 * staged in `d_scratch` (DEVICE, 3*nsites doubles).  Synchronises `stream`; -1 on error.  ewald() of layer (A) validates
 * the k-space sums that force_calc() started ahead of it with this. */
long mdb_sites_differ_host(mdb_engine *e, const double *x, const double *y, const double *z, double *d_scratch, void *stream);

/* Moldy's do_step() (src/accel.c:626-827) for NVE dynamics (const-temp = const-pressure = 0) on top of mdb_md_step: same
 * prototype (restart_header is only handed on to the host program's dump()), same outputs (meansq_f_t, pe, dip_mom,
 * stress_vir, sys->H_0 at the first step), the host program's c_of_m / quat / mom / amom arrays advanced in place.
 * INTEGRATION.md section 6. */
void do_step(system_mt *sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, vec_mt (*meansq_f_t)[2], double *pe,
             real *dip_mom, mat_mt stress_vir, void *restart_header, int backup_restart, int init_H_0);
void mdb_do_step_moldy(system_mt *sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, vec_mt (*meansq_f_t)[2], double *pe,
                       real *dip_mom, mat_mt stress_vir, void *restart_header, int backup_restart, int init_H_0);

/* Device->host copy of a result block (synchronises `stream`). */
int  mdb_read_out(mdb_engine *e, const double *d_out, double *h_out, void *stream);

/* Introspection used by the parity tests and bench.py. */
int    mdb_grid(const mdb_engine *e, int nxyz[3]);           /* link-cell grid          */
int    mdb_n_neighbour_cells(const mdb_engine *e);            /* half list, as NABORS/2  */
int    mdb_n_kvectors(const mdb_engine *e);                   /* nhkl                    */
int    mdb_pair_split(const mdb_engine *e);                   /* 1: real space runs as one pass per site class (charged / pair potential) */
int    mdb_get_cell_ids(mdb_engine *e, int *h_cell, void *stream);   /* [nsites], NCELL()  */
double mdb_pair_count(mdb_engine *e, void *stream);           /* pairs handed to kernel() per force evaluation
                                                                  (src/force.c:960) for the current sites */
long   mdb_kernel_launches(const mdb_engine *e);              /* our kernels launched so far */
int    mdb_too_close(mdb_engine *e, int pair[2], void *stream);/* count of r^2<0.25 inter-molecular pairs */
size_t mdb_sizeof(const char *struct_name);                   /* "contr_mt", "system_mt", ... */
double mdb_fp64_peak_probe(int device, int iters);            /* measured DFMA rate, flop/s */
double mdb_dmma_peak_probe(int device, int iters);            /* measured DMMA.8x8x4 (FP64 tensor pipe) rate, flop/s */
double mdb_recip_gemm_flop(const mdb_engine *e);              /* flop the k-space GEMM kernels execute per call on this
                                                                  rank (2 per multiply-add of every DMMA issued)      */

#ifdef __cplusplus
}
#endif
#endif /* MOLDY_B200_H */
