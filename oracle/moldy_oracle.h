/*
 * moldy_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of Moldy's force-evaluation hot path (force.c, kernel.c,
 * ewald.c of the reference), written from the reference's behaviour with a flat
 * array interface.  It exists to check the CUDA path (tests/, smoke(), bench.py's
 * cpu_baseline leg) and is itself pinned against
 *   (1) the start-up goldens in the reference's example outputs, and
 *   (2) the reference's own compiled force_calc()/ewald() (oracle/_ref) --
 *       bit-for-bit on forces/energies for the committed fixtures, because the
 *       restatement keeps the reference's operation and summation order.
 * Nothing under moldy_b200/ may include, link or load this.
 */
#ifndef MOLDY_ORACLE_H
#define MOLDY_ORACLE_H

typedef struct {
   /* system (what force_calc/ewald read from system_mt, spec_mt, pot_mt) */
   int nsites, nsites_xf;          /* total sites; non-framework sites (frameworks last) */
   int max_id, ptype, n_potpar;
   const int *site_type;           /* [nsites] site id                                   */
   const int *site_mol;            /* [nsites] molecule index                            */
   const double *chg;              /* [nsites]                                           */
   const double *potpar;           /* [max_id*max_id*8]                                  */
   double h[9];                    /* cell matrix, row major                             */
   /* control */
   double cutoff, subcell, alpha, k_cutoff;
   int strict_cutoff;
   int molpbc;                     /* molecular cut-off: non-framework molecules binned by c-of-m */
   const double *c_of_m;           /* [nmols][3] scaled centre-of-mass co-ordinates (molpbc only)  */
   /* replicated-data partition (the reference's globals ithread/nthreads) */
   int ithread, nthreads;
} orc_system;

typedef struct {
   double pe;                      /* energy summed by this call (no constants applied)  */
   double stress[9];               /* upper triangle only                                */
   double npairs;                  /* pairs handed to the pair kernel                    */
   int    n_too_close;             /* inter-molecular pairs with r^2 < 0.25              */
   int    n_bin_errors;
   int    nx, ny, nz;              /* link-cell grid                                     */
   int    n_nabors;                /* half neighbour-cell list length                    */
   int    nhkl;                    /* k-vectors                                          */
} orc_result;

/* scalar pieces */
double orc_det3(const double a[9]);
void   orc_invert3(const double a[9], double inv[9]);
int    orc_cellbin(double s, int n, double fn, double eps, int *err);
double orc_err_fn(double x);
void   orc_pair(int ptype, double alpha, double norm, double r_sqr, double qq, const double *p,
                double *fij, double *phi);
double orc_dist_pot(const double *p, double cutoff, int ptype);
int    orc_half_list(const double h[9], double cutoff, int strict, int nx, int ny, int nz,
                     int *out, int cap);

/* per-site link-cell index (NCELL order); returns number of binning errors */
int    orc_cell_ids(const orc_system *s, const double *x, const double *y, const double *z, int *cell);

/* real-space forces: f{x,y,z}[nsites] += ; returns 0 or -1 (cut-off too large) */
int    orc_force_calc(const orc_system *s, const double *x, const double *y, const double *z,
                      double *fx, double *fy, double *fz, orc_result *res);
/* RDF binning pass of force_calc: counts[max_id(max_id-1)/2][nbins] += pairs per bin */
int    orc_rdf(const orc_system *s, const double *x, const double *y, const double *z, double limit, int nbins,
               double *counts);
/* reciprocal-space forces */
int    orc_ewald(const orc_system *s, const double *x, const double *y, const double *z,
                 double *fx, double *fy, double *fz, orc_result *res);

/* first-call constants: intramolecular correction, Ewald self energy, sheet energy */
double orc_eintra(const orc_system *s, int nspecies, const int *spec_nsites, const int *spec_nmols,
                  const int *spec_framework, const double *p_f_sites /* concatenated [sum nsites][3] */);
void   orc_self_energy(const orc_system *s, int nspecies, const int *spec_nsites, const int *spec_nmols,
                       const int *spec_framework, const double *p_f_sites, double *self_e, double *sheet_e);
#endif
