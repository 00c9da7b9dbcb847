// mdb_internal.h -- engine state shared by the translation units of libmoldy_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include <vector>
#include "moldy_b200.h"

#define MDB_PI 3.14159265358979323846            /* src/defs.h:184 */
#define MDB_TOO_CLOSE 0.25                       /* src/force.c:89 */
#define MDB_ALPHAMIN 1e-7                        /* src/defs.h:181 */
#define MDB_EPS0 (0.25 / MDB_PI)                 /* src/defs.h:227 */

// ---- host-side tables ------------------------------------------------------
struct StencilRun {            // one contiguous z-run of neighbour cells in column (dx,dy)
   int dx, dy, dzlo, dzhi;     // full stencil = H u (-H), H = reference half list
};

struct HkDesc {                // one (h,k) column of the reciprocal lattice (traversal order)
   double kx, ky, kzt;         // src/ewald.c:445-447
   int    nl;                  // number of l-slots 0..nl-1 (0: nothing inside the cutoff)
   int    slot0;               // first row of this column in the per-slot arrays
   int    h, k;
   int    code;                // how (cos,sin)(h a*.r + k b*.r) follows from the previous column
   int    pad;
};
enum { HK_DIRECT = 0, HK_NEWH = 1, HK_KUP = 2, HK_KDOWN0 = 3, HK_KDOWN = 4 };

struct HostTables {
   // real space
   int nx = 0, ny = 0, nz = 0;
   double hinv[9];
   std::vector<int> half_list;            // (ix,iy,iz) triplets of the reference half list
   std::vector<StencilRun> runs;          // full stencil as z-runs
   std::vector<StencilRun> runs_half;     // the reference half list itself as z-runs (Newton-3 kernel)
   double reloc[27][3];                   // src/force.c:1284-1293
   // reciprocal space
   int hmax = 0, kmax = 0, lmax = 0, nhkl = 0;
   double astar[3], bstar[3], cstar[3];
   double vol = 0;
   std::vector<HkDesc> hk;                // traversal order, includes empty columns
   std::vector<int> hk_valid;             // indices into hk with nl > 0 (sorted by nl, descending)
   std::vector<int> slot_flags;           // per slot: bit0 = +l inside cutoff, bit1 = -l inside
   int nslots = 0;
};

// ---- device-side parameter blocks (passed by value to kernels) --------------
struct CellParams {
   double hinv[9];
   int nx, ny, nz, ncells;
   double fnx, fny, fnz, eps;
   int molpbc, nsites_xf;                 // molecular-cutoff binning for the non-framework sites
};

struct PairParams {
   int nx, ny, nz, nruns;
   double reloc[27][3];
   double alpha, norm, cutoffsq, cutoff100sq;
   int max_id, strict;
   int s_lo, s_hi;                        // slice of cell-sorted sites owned by this rank
};

struct KspaceParams {
   double astar[3], bstar[3], cstar[3];
   double cz2;                            // cstar[2]
   double r4alpha;                        // -1/(4 alpha^2)
   double pref;                           // 2/(EPS0 vol)
   int hmax, kmax, lmax, nlslots;         // nlslots = lmax+1
   int nsites, nsites_xf;
};

// A class of sites (charged / with a non-zero pair potential) as its own cell-sorted list: the tiled
// pair kernel then visits charged x charged pairs with the Coulomb term only and potential x potential
// pairs with the pair potential only, instead of every pair with both (TIP4P: 27 of 54 FP64
// instructions per reference pair).  Compacted from the full cell-sorted list after every cell build.
struct SubList {
   int n = 0;                             // sites in the class (static per system)
   double4 *posq = nullptr; int2 *sinfo = nullptr;
   int *order = nullptr;                  // compacted sorted index -> original site index
   int *start = nullptr;                  // [ncells+1]
   int2 *batches = nullptr; int *nbatch = nullptr; int batch_cap = 0;
   double *fs = nullptr;                  // [3 n] cell-sorted force accumulator
   int cap_n = 0, cap_cells = 0;          // allocated sizes
   bool valid = false;
};

struct mdb_engine {
   int device = 0;
   bool configured = false;
   mdb_config cfg{};
   std::vector<int> h_type, h_mol;
   std::vector<double> h_chg, h_potpar;
   std::vector<unsigned char> h_cls; long n_cls[2] = {0, 0}; bool cls_uploaded = false, mol_is_identity = false;
   std::map<void *, size_t> table_cap;    // bytes allocated behind the device tables that upload() refills
   int slots_cap = 0;
   HostTables T;
   int ithread = 0, nthreads = 1;
   long launches = 0;

   // static per-system device data
   int *d_type = nullptr, *d_mol = nullptr;
   double *d_chg = nullptr, *d_ptab = nullptr;
   // positions (original site order); owned unless borrowed
   double *d_x = nullptr, *d_y = nullptr, *d_z = nullptr;
   double *own_xyz = nullptr;
   bool sites_set = false, cells_valid = false;
   double *d_com = nullptr; int com_cap = 0; bool com_set = false;   // scaled c-of-m (molecular-cutoff mode)
   // link cells
   int ncells = 0, cells_cap = 0;
   int *d_cell = nullptr, *d_count = nullptr, *d_start = nullptr, *d_order = nullptr;
   int *d_scan_tmp = nullptr;
   double4 *d_posq = nullptr;
   int *d_stype = nullptr, *d_scell = nullptr;
   int2 *d_sinfo = nullptr;               // {type | framework bit 30, z index of the cell}, cell-sorted
   StencilRun *d_runs = nullptr, *d_runs_half = nullptr;
   int nruns = 0, nruns_half = 0;
   // Exponential potentials (Buckingham, generic, Morse/BIG, MCY) with a cut-off of many decay lengths: the stencil runs whose
   // EVERY cell pair is further apart than r_far, where exp(-r/rho) has fallen below e^-52 = 2.6e-23 of its amplitude for
   // all site-type pairs, are walked by a second launch with the power-law rest of the potential (as PT_HIW rows) or with the
   // Coulomb term alone.  far_ptype < 0: no such runs / not applicable.
   StencilRun *d_runs_near = nullptr, *d_runs_far = nullptr; int nruns_near = 0, nruns_far = 0;
   int far_ptype = -1, far_enable = -1; double *d_ptab_far = nullptr; double r_far = 0.0;
   int2 *d_batches = nullptr; int *d_nbatch = nullptr; int batch_cap = 0;   // i-site batches of the tiled pair kernel
   double *d_fs = nullptr;                // [3N] cell-sorted force accumulator (Newton-3 mode)
   int pair_mode = -1;                    // 2: per-thread full stencil, 3: tiled full stencil, 4: tiled Newton-3
   bool batches_valid = false;
   int pair_split = 0;                    // tiled kernel: one pass per site class (sub[0] charged, sub[1] potential)
   SubList sub[2];
   unsigned char *d_cls = nullptr;        // per original site: bit 0 charged, bit 1 has a non-zero pair potential
   int *d_sub_flag = nullptr, *d_sub_pos = nullptr, *d_sub_scan = nullptr, *d_sub_cols = nullptr;
   int sub_cap = 0, sub_cols_cap = 0, sub_scan_cap = 0;
   cudaStream_t aux_stream = nullptr; cudaEvent_t ev_cells = nullptr, ev_sub1 = nullptr;   // class-1 compaction beside the class-0 pass
   // RDF pass: strict stencil of the last (limit, grid) and the device histogram
   StencilRun *d_runs_rdf = nullptr; int nruns_rdf = 0; double rdf_limit = -1.0; int rdf_grid[3] = {0, 0, 0};
   double rdf_h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
   unsigned long long *d_rdf = nullptr; size_t rdf_cap = 0;
   // reductions / diagnostics
   double *d_partials = nullptr; int partials_cap = 0;
   long pair_evals = 0;                   // force evaluations since the visit counter was last read (mode 2)
   unsigned long long *d_counters = nullptr;      // [0]=pair visits [1]=too close [2]=bin errors [3..4]=example pair
   // k-space
   HkDesc *d_hk = nullptr; int *d_hk_valid = nullptr; int *d_slot_flags = nullptr;
   double *d_ppart = nullptr; size_t ppart_cap = 0;
   double *d_coef_tot = nullptr, *d_coef_nf = nullptr;
   int *d_cidx = nullptr;                 // charged sites (non-framework first), original indices
   int n_charged = 0, n_charged_nf = 0;
   void *d_sfac_blocks = nullptr; int n_sfac_blocks = 0, sfac_rank = -1, sfac_nranks = -1, sfac_mode = -1;
   void *d_kf_groups = nullptr; int n_kf_groups = 0, kf_rank = -1, kf_nranks = -1;   // k_kforce_mma: coefficient blocks (descriptor + planes) of the column groups, total | non-framework
   int *d_kf_slot_dst = nullptr;          // slot -> entry of its block (k_sfin writes through it)
   void *d_ktab = nullptr; size_t ktab_cap = 0;      // per-site power tables E_l | E_h | E_k of the DMMA kernels
   double *d_kpartials = nullptr;
   double *d_psum = nullptr;              // [2][nslots][4] structure-factor sums (non-framework, framework)
   int n_slabs = 0, n_slabs_nf = 0, slab_sites = 0;
   // k_kforce_mma launched in `kf_chunks` slices of the (non-framework) charged sites, an event after each: the host-facing
   // layer brings a slice of the k-space forces home and adds it to the caller's array while the next slice is computed
   static constexpr int KF_MAXCH = 8;
   int kf_chunks = 1, kf_nch = 0; cudaEvent_t kf_ev[KF_MAXCH] = {}; int kf_site_hi[KF_MAXCH] = {};
   std::vector<int> h_cidx;

   // eval_forces() on the device (mdb_molframe.cu): species table and molecular-frame buffers
   struct MolFrame {
      std::vector<mdb_species> sp;
      std::vector<int> site_off, mol_off, quat_off, torq_off, pfs_off, blk_off;   // per species; quat_off/torq_off -1: none
      int nmols = 0, nmols_q = 0, nmols_r = 0, npfs = 0, nblocks = 0;
      double *d_pfs = nullptr, *d_in = nullptr, *d_res = nullptr, *d_vpart = nullptr, *d_dpart = nullptr;
      double *h_in = nullptr, *h_res = nullptr;                                   // pinned
      size_t in_cap = 0, res_cap = 0;
      double rdf_limit = 0.0; int rdf_nbins = 0; unsigned long long *rdf_counts = nullptr;   // one-shot RDF request
      // resident NVE integrator (mdb_md.cu): momenta, per-species dynamics, sums
      std::vector<mdb_species_dyn> dyn; int nosymmetric_rot = 0, saxis = 0;
      double *d_mom = nullptr, *d_amom = nullptr, *d_mdpart = nullptr, *d_mdscal = nullptr, *h_mdscal = nullptr;
      double *h_state = nullptr; size_t state_cap = 0;      // pinned staging of the state arrays (upload/download)
   } mf;

   // real space beside k-space (mdb_force_both): the k-space chain runs on a high-priority side stream into d_out2, a small
   // persistent pair grid fills the FP64 issue slots its DMMA stream leaves idle; d_ovl_q = {next batch, stop flag}
   int ovl_blocks = -2, ovl_threads = 0; bool ovl_armed = false;     // -2: not chosen yet (MDB_OVERLAP), -1: off, 0: k-space first with the set-up behind it, > 0: filler grid
   // an event the pair passes wait for, after the (launch-latency-bound) cell build and sub-list compaction are enqueued: a
   // k-space kernel on a side stream then hides that set-up and the pair kernel does not share the SMs with it (on one GPU two
   // FP64-bound kernels side by side lose up to 7 % against the same two in sequence, profiles/r02_summary.md; at 8 GPUs
   // waiting and sharing measured the same, 3.339 ms for phase A)
   cudaEvent_t pre_pair_wait = nullptr;
   int *d_ovl_q = nullptr; double *d_out2 = nullptr; size_t out2_cap = 0;
   cudaStream_t ovl_stream = nullptr; cudaEvent_t ev_ovl_fork = nullptr, ev_ovl_join = nullptr;

   // pinned staging for host-facing calls
   double *h_stage = nullptr; size_t stage_cap = 0;
   double *d_out_own = nullptr; size_t out_cap = 0;
};

// ---- host setup (mdb_host.cpp) ---------------------------------------------
void mdb_invert3(const double a[9], double b[9]);
double mdb_det3(const double a[9]);
bool mdb_build_real_tables(const mdb_config &c, HostTables &T, std::string &err);
bool mdb_build_recip_tables(const mdb_config &c, HostTables &T, std::string &err);
bool mdb_build_rdf_runs(const mdb_config &c, const HostTables &T, double limit, std::vector<StencilRun> &runs,
                        std::string &err);
double mdb_err_fn(double x);
void mdb_set_error(const std::string &s);

// ---- kernels launchers (mdb_cells.cu / mdb_pair.cu / mdb_kspace.cu) --------
int mdb_launch_cells(mdb_engine *e, cudaStream_t st);
int mdb_launch_pair(mdb_engine *e, double *d_out, cudaStream_t st);
int mdb_launch_pair_tiled(mdb_engine *e, double *d_out, cudaStream_t st);
int mdb_launch_pair_count_tiled(mdb_engine *e, cudaStream_t st);
int mdb_launch_rdf_tiled(mdb_engine *e, const StencilRun *d_runs, int nruns, double rbin, int nbins,
                         unsigned long long *d_counts, cudaStream_t st);
static constexpr size_t MDB_TILED_TAB_MAX = 28 * 1024;   // pair table of the tiled kernel lives in shared memory
int mdb_launch_batches(mdb_engine *e, cudaStream_t st);
int mdb_need_batches(mdb_engine *e, cudaStream_t st);     // batches of the full cell-sorted list, built once per cell build
int mdb_build_sublist(mdb_engine *e, int k, cudaStream_t st);
int mdb_launch_too_close_scan(mdb_engine *e, cudaStream_t st);
#ifndef MDB_NI_SITES
#define MDB_NI_SITES 4
#endif
static constexpr int MDB_NI = MDB_NI_SITES;   // i-sites per warp in the tiled pair kernel (2..4)
int mdb_launch_recip(mdb_engine *e, double *d_out, cudaStream_t st);
int mdb_launch_recip_partial(mdb_engine *e, double *d_psum, cudaStream_t st);
int mdb_launch_recip_finish(mdb_engine *e, const double *d_psum, double *d_out, cudaStream_t st);
int mdb_launch_kernel_vec(int jmin, int nnab, double *forceij, double *pe, const double *r_sqr,
                          const double *nab_chg, double chg, double norm, double alpha, int ptype,
                          double *const *pot);

// pieces of the NVE step (mdb_md.cu) on a range of molecules, shared with the multi-GPU driver (mdb_group.cu)
struct MdbCopyJob { double *dst; const double *src; size_t bytes; };     // src == nullptr: zero fill
void mdb_run_copy_jobs(const std::vector<MdbCopyJob> &jobs);            // host copies on up to six threads
int mdb_md_coords_range(mdb_engine *e, const double h[9], double step, double ts, double *d_in, int m_lo, int m_hi, cudaStream_t st);
int mdb_md_momenta_range(mdb_engine *e, const double h[9], double step, int m_lo, int m_hi, cudaStream_t st);
int mdb_md_sums_range(mdb_engine *e, const double h[9], int slot, bool with_forces, int m_lo, int m_hi, cudaStream_t st);

// pieces of eval_forces() (mdb_molframe.cu), shared with the multi-GPU driver (mdb_group.cu)
int mdb_evalf_stage_inputs(mdb_engine *e, const double *const *com, const double *const *quat, double *h_in);
int mdb_evalf_make_sites(mdb_engine *e, const double h[9], const double *d_in, bool second, cudaStream_t st);
int mdb_evalf_pre(mdb_engine *e, const double h[9], const double *d_in, cudaStream_t st);
int mdb_evalf_tail(mdb_engine *e, const double h[9], const double *d_in, const double *d_fblock, int m_lo, int m_hi,
                   int surface_dipole, int do_recip, cudaStream_t st);

#define MDB_CUDA(call)                                                                   \
   do {                                                                                  \
      cudaError_t _e = (call);                                                           \
      if (_e != cudaSuccess) {                                                           \
         mdb_set_error(std::string(#call) + ": " + cudaGetErrorString(_e));              \
         return -1;                                                                      \
      }                                                                                  \
   } while (0)
