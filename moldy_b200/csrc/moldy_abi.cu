// moldy_abi.cu -- layer (A) of include/moldy_b200.h: Moldy's own entry points
// (force_calc, ewald, kernel, poteval, dist_pot, potspec, pot_dim, and eval_forces
// one level up) on top of the device engine.  Linked in place of force.o / kernel.o / ewald.o the rest of
// Moldy is unchanged.  There is no CPU path: if no CUDA device
// can be opened every entry point reports FATAL through Moldy's message().
//
// Calling contract reproduced: the caller owns and zeroes
// site_force / pe / stress; we only accumulate (+=); pe[0] gets the real-space
// energy minus the intramolecular correction, the caller's pe+1 gets the
// reciprocal-space energy minus self energy plus sheet term; only the upper
// triangle of stress is touched; rank-0-only constants follow `ithread`.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <thread>
#include <vector>
#include "mdb_internal.h"

#define CONV_E_KJ (0.001 * 6.0221367e23 * 1.6605402e-27 * 1.0e4)     /* src/defs.h:229 */
#define CONV_Q_E  (4.07497263794495e-14 * 1.e-3 * 1.05482230112e-05 / 1.60217733e-19)

enum { SEV_INFO = 0, SEV_WARNING = 1, SEV_ERROR = 2, SEV_FATAL = 3 };       /* src/messages.h */

extern "C" {
// Globals owned by Moldy's main.c (src/main.c:83-84).  Weak here so that a host
// that is not Moldy (ctypes, bench.py) still resolves them; Moldy's own strong
// definitions win when the library is linked into the program.
__attribute__((weak)) contr_mt control;
__attribute__((weak)) int ithread = 0;
__attribute__((weak)) int nthreads = 1;

// Moldy's RDF store (src/rdf.c:60-64): float[nbins * max_id (max_id - 1) / 2].  The weak fall-back keeps a
// private array so that force_calc's RDF pass can be read back when the host program is not Moldy.
static std::vector<float> g_rdf_private;
__attribute__((weak)) void *rdf_ptr(int *size)
{
   *size = (int)g_rdf_private.size();
   return g_rdf_private.data();
}
void mdb_rdf_private_resize(int n) { g_rdf_private.assign((size_t)n, 0.0f); }

// Moldy's message()/note() (src/output.c:131,175); weak fall-backs print the same tags.
__attribute__((weak)) void note(char *text, ...)
{
   if (ithread > 0) return;
   va_list ap;
   va_start(ap, text);
   printf(" *I* ");
   vprintf(text, ap);
   printf("\n");
   va_end(ap);
}
__attribute__((weak)) void message(int *nerrs, ...)
{
   static const char *tag[] = {" *I* ", " *W* ", " *E* ", " *F* "};
   va_list ap;
   va_start(ap, nerrs);
   char *buff = va_arg(ap, char *);
   int sev = va_arg(ap, int);
   char *fmt = va_arg(ap, char *);
   (void)buff;
   if (ithread == 0 || abs(sev) == SEV_FATAL) {
      printf("%s", tag[abs(sev) & 3]);
      vprintf(fmt, ap);
      printf("\n");
   }
   va_end(ap);
   if (sev >= SEV_ERROR && nerrs) (*nerrs)++;
   if (abs(sev) == SEV_FATAL) { fflush(stdout); exit(3); }
}
contr_mt *mdb_control(void) { return &control; }
void mdb_set_thread(int it, int nt) { ithread = it; nthreads = nt; }

const pots_mt potspec[] = {{(char *)"lennard-jones", 2}, {(char *)"buckingham", 3}, {(char *)"mcy", 4},
                           {(char *)"generic", 6},       {(char *)"hiw", 3},
                           {(char *)"reserved for developer", 1}, {(char *)"morse", 7}, {0, 0}};
const dim_mt pot_dim[][MDB_NPOTP] = {
   {{1, 2, -2}, {0, 1, 0}},
   {{1, 8, -2}, {1, 2, -2}, {0, -1, 0}},
   {{1, 2, -2}, {0, -1, 0}, {1, 2, -2}, {0, -1, 0}},
   {{1, 2, -2}, {0, -1, 0}, {1, 14, -2}, {1, 6, -2}, {1, 8, -2}, {1, 10, -2}},
   {{1, 6, -2}, {1, 8, -2}, {1, 14, -2}},
   {{0, 0, 0}},
   {{1, 2, -2}, {0, 1, 0}, {0, -1, 0}, {1, 8, -2}, {1, 2, -2}, {0, -1, 0}, {0, 1, 0}}};
}

#define FATAL_MSG(...) message((int *)0, (char *)0, SEV_FATAL, (char *)__VA_ARGS__)

// ---- process-wide state (the reference keeps the same in function statics) -----
struct AbiState {
   mdb_engine *eng = nullptr;
   mdb_group *group = nullptr;                              // MOLDY_B200_DEVICES names > 1 device: eng = the group's rank 0
   cudaStream_t stream = nullptr, copy_stream = nullptr;   // copy_stream: D2H of the real-space block beside the k-space kernels
   cudaEvent_t ev_real = nullptr, ev_copied = nullptr;
   // look-ahead k-space sum in slices: slice k of the forces is copied home on copy_stream behind its event and added to the
   // caller's rows by ewald() while slice k+1 is computed; check_stream carries ewald()'s validation upload of the site rows
   cudaStream_t check_stream = nullptr;
   int ahead_nslice = 0; int ahead_hi[8] = {0}; cudaEvent_t ev_slice[8] = {nullptr};
   int kf_slices = -1, kf_min_sites = 200000;          // MOLDY_B200_KF_SLICES (default 4) for systems of >= kf_min_sites sites
   bool chg_unchecked = false;                              // sync_config(defer_chg): contents of chg[] still to be compared
   bool real_init = false, recip_init = false;
   double eintra = 0, self_energy = 0, sheet_energy = 0;
   int onabor = 0, onx = 0, ony = 0, onz = 0;
   std::vector<int> type, mol;
   std::vector<double> potflat, chg;
   mdb_config cfg{};
   bool have_cfg = false;
   const void *last_sites = nullptr;
   bool sites_fresh = false;
   double *d_out = nullptr; size_t out_cap = 0;
   double *h_out = nullptr; size_t hout_cap = 0;
   // ewald() results computed ahead of the call, behind force_calc's host-side work (see force_calc)
   double *d_out2 = nullptr, *h_out2 = nullptr;
   cudaEvent_t ev_ahead = nullptr;
   bool ahead_valid = false;
   const void *ahead_sites = nullptr;
   long config_epoch = 0, ahead_epoch = -1, chg_epoch = 0;     // chg_epoch: bumped whenever G.chg is replaced
   int ahead_ithread = 0, ahead_nthreads = 1;
   bool rdf_warned = false;
};
static AbiState G;

// MOLDY_B200_DEVICES = "all" | "0-7" | "0,1,4" (a device may repeat: several ranks on one GPU)
static std::vector<int> parse_devices(const char *s)
{
   std::vector<int> d;
   int ndev = 0;
   cudaGetDeviceCount(&ndev);
   if (!strcmp(s, "all")) { for (int i = 0; i < ndev; i++) d.push_back(i); return d; }
   while (*s) {
      char *end;
      const long a = strtol(s, &end, 10);
      if (end == s) break;
      long b = a;
      if (*end == '-') { s = end + 1; b = strtol(s, &end, 10); }
      for (long v = a; v <= b; v++) d.push_back((int)v);
      s = *end ? end + 1 : end;
   }
   return d;
}

static void ensure_engine()
{
   if (G.eng) return;
   int dev = 0;
   const char *s = getenv("MOLDY_B200_DEVICE");
   if (!s) s = getenv("LOCAL_RANK");
   if (s) dev = atoi(s);
   const char *many = getenv("MOLDY_B200_DEVICES");
   std::vector<int> devs = many ? parse_devices(many) : std::vector<int>();
   if (devs.size() > 1) {
      if (nthreads > 1) FATAL_MSG("libmoldy_b200: MOLDY_B200_DEVICES drives all GPUs from one process; run Moldy without SPMD");
      G.group = mdb_group_create((int)devs.size(), devs.data());
      if (!G.group) FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
      G.eng = mdb_group_engine(G.group, 0);
      dev = devs[0];
      cudaSetDevice(dev);
   } else {
      if (devs.size() == 1) dev = devs[0];
      G.eng = mdb_create(dev);
   }
   if (!G.eng) FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
   if (cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking) != cudaSuccess ||
       cudaStreamCreateWithFlags(&G.copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
       cudaEventCreateWithFlags(&G.ev_real, cudaEventDisableTiming) != cudaSuccess ||
       cudaEventCreateWithFlags(&G.ev_copied, cudaEventDisableTiming) != cudaSuccess)
      FATAL_MSG("libmoldy_b200: cannot create CUDA stream");
}

static int count_xf_sites(const system_mt *system, const spec_mt *species)
{
   int n = 0;
   for (const spec_mt *sp = species; sp < species + system->nspecies && !sp->framework; sp++)
      n += sp->nsites * sp->nmols;
   return n;
}

// Flatten system/species/potpar/control into mdb_config and (re)configure the engine
// when anything that shapes the tables changed (cell matrix changes every step under
// constant-stress dynamics; everything else is constant in a Moldy run).
// defer_chg: do not compare the N charges now (8 MB, 0.6 ms at 10^6 sites); the caller does it with
// chg_changed_late() once its kernels are running and repeats the call if they differ.
static void sync_config(system_mt *system, spec_mt *species, const real *chg, const pot_mt *potpar, bool defer_chg = false,
                        bool chg_known = false)      // chg_known: chg IS G.chg's content (eval_forces' cached array)
{
   ensure_engine();
   const int n = system->nsites, max_id = system->max_id;
   int nfw = 0;
   for (int i = 0; i < system->nspecies; i++) nfw += species[i].framework ? species[i].nmols : 0;
   if (nfw > 1) FATAL_MSG("Multiple framework molecules are not supported");     /* src/force.c:1218 */

   mdb_config c{};
   c.nsites = n; c.nsites_xf = count_xf_sites(system, species);
   c.max_id = max_id; c.ptype = system->ptype; c.n_potpar = system->n_potpar;
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) c.h[3 * i + j] = system->h[i][j];
   c.cutoff = control.cutoff; c.subcell = control.subcell; c.alpha = control.alpha;
   c.k_cutoff = control.k_cutoff; c.strict_cutoff = control.strict_cutoff;
   c.do_recip = control.alpha > MDB_ALPHAMIN;
   c.molpbc = control.molpbc ? 1 : 0; c.nmols = system->nmols;

   bool changed = !G.have_cfg || c.nsites != G.cfg.nsites || c.nsites_xf != G.cfg.nsites_xf ||
                  c.max_id != G.cfg.max_id || c.ptype != G.cfg.ptype || memcmp(c.h, G.cfg.h, sizeof c.h) ||
                  c.cutoff != G.cfg.cutoff || c.subcell != G.cfg.subcell || c.alpha != G.cfg.alpha ||
                  c.k_cutoff != G.cfg.k_cutoff || c.strict_cutoff != G.cfg.strict_cutoff ||
                  c.do_recip != G.cfg.do_recip || c.molpbc != G.cfg.molpbc || c.nmols != G.cfg.nmols;
   if ((int)G.type.size() != n) {
      // site id and molecule maps (src/force.c:1173-1191)
      G.type.resize(n); G.mol.resize(n);
      int js = 0, jm = 0;
      for (const spec_mt *sp = species; sp < species + system->nspecies; sp++)
         for (int im = 0; im < sp->nmols; im++, jm++)
            for (int is = 0; is < sp->nsites; is++, js++) { G.type[js] = sp->site_id[is]; G.mol[js] = jm; }
      changed = true;
   }
   G.chg_unchecked = false;
   if (chg_known && (int)G.chg.size() == n) {
      /* nothing to compare */
   } else if (defer_chg && !changed && (int)G.chg.size() == n) {
      G.chg_unchecked = true;
   } else if ((int)G.chg.size() != n || memcmp(G.chg.data(), chg, sizeof(double) * n)) {
      G.chg.assign(chg, chg + n);
      G.chg_epoch++;
      changed = true;
   }
   if (potpar) {
      std::vector<double> flat((size_t)max_id * max_id * MDB_NPOTP);
      for (int k = 0; k < max_id * max_id; k++) memcpy(&flat[(size_t)k * MDB_NPOTP], potpar[k].p, sizeof(double) * MDB_NPOTP);
      if (flat != G.potflat) { G.potflat.swap(flat); changed = true; }
   } else if (G.potflat.size() != (size_t)max_id * max_id * MDB_NPOTP) {
      G.potflat.assign((size_t)max_id * max_id * MDB_NPOTP, 0.0);
      changed = true;
   }
   if (!changed) return;
   c.site_type = G.type.data(); c.site_mol = G.mol.data(); c.chg = G.chg.data(); c.potpar = G.potflat.data();
   if (G.group ? mdb_group_configure(G.group, &c) : mdb_configure(G.eng, &c)) FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
   if (G.group) cudaSetDevice(G.eng->device);
   G.config_epoch++;
   G.ahead_valid = false;
   G.cfg = c;
   G.have_cfg = true;
   G.sites_fresh = false;
   const size_t need = mdb_out_doubles(n);
   if (need > G.out_cap) {
      if (G.d_out) cudaFree(G.d_out);
      if (G.h_out) cudaFreeHost(G.h_out);
      if (G.d_out2) cudaFree(G.d_out2);
      if (G.h_out2) cudaFreeHost(G.h_out2);
      if (cudaMalloc(&G.d_out, sizeof(double) * need) != cudaSuccess ||
          cudaMallocHost(&G.h_out, sizeof(double) * need) != cudaSuccess ||
          cudaMalloc(&G.d_out2, sizeof(double) * need) != cudaSuccess ||
          cudaMallocHost(&G.h_out2, sizeof(double) * need) != cudaSuccess)
         FATAL_MSG("libmoldy_b200: out of device/pinned memory for %d sites", n);
      G.out_cap = need;
   }
}

static bool chg_changed_late(const real *chg, int n)
{
   if (!G.chg_unchecked) return false;
   G.chg_unchecked = false;
   return memcmp(G.chg.data(), chg, sizeof(double) * n) != 0;
}

static void push_sites(real **site)
{
   if (mdb_set_sites_host(G.eng, site[0], site[1], site[2], G.stream)) FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
   G.last_sites = site[0];
}

static double now_ms()
{
   return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static const bool g_timing = getenv("MOLDY_B200_TIMING") != nullptr;

// D2H of the result block into pinned memory, then site_force[a][i] += f (the caller's arrays
// are only accumulated into, src/accel.c:488-535); the three rows are added by three threads.
static void pull_and_accumulate(real **site_force, double *pe, real (*stress)[3], int n, const double *d_src = nullptr,
                                double *h_dst = nullptr, bool on_host = false)
{
   const double t0 = now_ms();
   if (!d_src) { d_src = G.d_out; h_dst = G.h_out; }
   if (!on_host && mdb_read_out(G.eng, d_src, h_dst, G.stream)) FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
   const double t1 = now_ms();
   // site_force[a][i] += f: 3 rows x 2 halves on six threads (memory-bound, ~25 MB read + 25 MB updated)
   auto add_part = [&](int a, int part, int nparts) {
      const size_t lo = (size_t)n * part / nparts, hi = (size_t)n * (part + 1) / nparts;
      real *dst = site_force[a];
      const double *src = h_dst + (size_t)a * n;
      for (size_t i = lo; i < hi; i++) dst[i] += src[i];
   };
   if (n >= 65536) {
      std::thread th[5];
      for (int k = 1; k < 6; k++) th[k - 1] = std::thread(add_part, k / 2, k % 2, 2);
      add_part(0, 0, 2);
      for (auto &t : th) t.join();
   } else {
      for (int a = 0; a < 3; a++) add_part(a, 0, 1);
   }
   if (g_timing) fprintf(stderr, "[moldy_b200] wait+D2H %.2f ms, host += %.2f ms\n", t1 - t0, now_ms() - t1);
   const double *sc = h_dst + 3 * (size_t)n;
   (void)pe;
   stress[0][0] += sc[2]; stress[0][1] += sc[3]; stress[0][2] += sc[4];
   stress[1][1] += sc[6]; stress[1][2] += sc[7]; stress[2][2] += sc[10];
}

// site_force[a][lo..hi) += rows of a pinned result block, on up to six threads
static void accumulate_rows(real **site_force, const double *h_src, int n, size_t lo, size_t hi)
{
   if (hi <= lo) return;
   auto add_part = [&](int a, int part, int nparts) {
      const size_t l = lo + (hi - lo) * part / nparts, h = lo + (hi - lo) * (part + 1) / nparts;
      real *dst = site_force[a];
      const double *src = h_src + (size_t)a * n;
      for (size_t i = l; i < h; i++) dst[i] += src[i];
   };
   if (hi - lo >= 65536) {
      std::thread th[5];
      for (int k = 1; k < 6; k++) th[k - 1] = std::thread(add_part, k / 2, k % 2, 2);
      add_part(0, 0, 2);
      for (auto &t : th) t.join();
   } else {
      for (int a = 0; a < 3; a++) add_part(a, 0, 1);
   }
}

/* Radial distribution functions, src/force.c:1302-1313 + src/rdf.c:94-108.  The pairs are binned on the device;
 * count/density is added to the host program's float histograms in one step (the reference adds 1/density pair by
 * pair in single precision, so it rounds differently). */
static bool rdf_due()
{
   return control.rdf_interval > 0 && control.istep >= control.begin_rdf && control.istep % control.rdf_interval == 0;
}
static float *rdf_store(size_t nh)
{
   int rsize = 0;
   float *rdf_base = (float *)rdf_ptr(&rsize);
   if (rdf_base && (size_t)rsize >= nh && nh > 0) return rdf_base;
   if (!G.rdf_warned) {
      message((int *)0, (char *)0, SEV_WARNING, (char *)"libmoldy_b200: no RDF store (rdf_ptr) to accumulate into");
      G.rdf_warned = true;
   }
   return nullptr;
}
static void rdf_add_counts(system_mt *system, float *rdf_base, const std::vector<unsigned long long> &cnt)
{
   const double hm[9] = {system->h[0][0], system->h[0][1], system->h[0][2], system->h[1][0], system->h[1][1],
                         system->h[1][2], system->h[2][0], system->h[2][1], system->h[2][2]};
   const double invrho = 1.0 / (system->nsites / mdb_det3(hm));
   for (size_t k = 0; k < cnt.size(); k++)
      if (cnt[k]) rdf_base[k] = (float)(rdf_base[k] + (double)cnt[k] * invrho);
}

// First-call constant and notes of force_calc (src/force.c:1158-1169, 218-223, 1240-1243)
static void real_first_call(system_mt *system, spec_mt *species, const real *chg, pot_mt *potpar)
{
   if (!G.real_init) {                                   /* src/force.c:1158-1169 */
      int isite = 0;
      for (spec_mt *sp = species; sp < species + system->nspecies; sp++) {
         if (!sp->framework) {
            double e = 0.0;
            for (int js = 0; js < sp->nsites; js++)
               for (int is = js + 1; is < sp->nsites; is++) {
                  const double *a = sp->p_f_sites[is], *b = sp->p_f_sites[js];
                  double r = sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) +
                                  (a[2] - b[2]) * (a[2] - b[2]));
                  e += poteval(potpar[sp->site_id[js] * system->max_id + sp->site_id[is]].p, r, system->ptype,
                               chg[isite + is] * chg[isite + js]);
               }
            G.eintra += sp->nmols * e;
         }
         isite += sp->nmols * sp->nsites;
      }
      note((char *)"Intramolecular potential energy correction = %g", G.eintra * CONV_E_KJ);
      G.real_init = true;
   }
   const int nhalf = mdb_n_neighbour_cells(G.eng);
   if (nhalf != G.onabor) {                              /* src/force.c:218-223 */
      note((char *)"Neighbour list contains %d cells", 2 * nhalf);
      G.onabor = nhalf;
   }
   int g[3];
   const int ncells = mdb_grid(G.eng, g);
   if (g[0] != G.onx || g[1] != G.ony || g[2] != G.onz) { /* src/force.c:1240-1243 */
      note((char *)"MD cell divided into %d subcells (%dx%dx%d)", ncells, g[0], g[1], g[2]);
      G.onx = g[0]; G.ony = g[1]; G.onz = g[2];
   }
}

// force_calc()/ewald() on all GPUs of the group: slices of the site rows in, slices of the summed forces out (pinned block),
// then the usual += into the caller's arrays.  what: 1 real space (pe = caller's pe), 2 reciprocal space (pe = caller's pe+1).
static void group_force(real **site, real **site_force, system_mt *system, int what, double *pe, real (*stress)[3])
{
   const int n = system->nsites;
   int tc = 0, pr[2] = {0, 0};
   double *h = G.h_out;
   if (mdb_group_force_host(G.group, site[0], site[1], site[2], control.molpbc ? &system->c_of_m[0][0] : nullptr, what, h, h + n,
                            h + 2 * (size_t)n, h + 3 * (size_t)n, what == 1 ? &tc : nullptr, pr))
      FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
   cudaSetDevice(G.eng->device);
   pull_and_accumulate(site_force, pe, stress, n, G.d_out, h, true);
   *pe += h[3 * (size_t)n + (what == 1 ? 0 : 1)];
   if (tc & (1 << 30))
      message((int *)0, (char *)0, SEV_ERROR, (char *)"Co-ordinate out of range in BIN (fill_cells)");
   if (tc & ~(1 << 30))                                  /* src/force.c:944-946 */
      message((int *)0, (char *)0, SEV_WARNING, (char *)"Sites %d and %d closer than %fA.", pr[0], pr[1],
              sqrt(MDB_TOO_CLOSE));
   G.sites_fresh = false; G.ahead_valid = false;
}

extern "C" void force_calc(real **site, real **site_force, system_mt *system, spec_mt *species, real *chg,
                           pot_mt *potpar, double *pe, mat_mt stress)
{
   const double tc0 = now_ms();
   sync_config(system, species, chg, potpar, true);
   if (g_timing) fprintf(stderr, "[moldy_b200] force_calc: sync_config %.2f ms\n", now_ms() - tc0);
   mdb_set_partition(G.eng, ithread, nthreads);
   const int n = system->nsites;

   real_first_call(system, species, chg, potpar);
   if (ithread == 0) *pe -= G.eintra;

   if (G.group) {                                        /* all GPUs of MOLDY_B200_DEVICES: complete sums (nthreads = 1) */
      chg_changed_late(chg, n);
      group_force(site, site_force, system, 1, pe, stress);
      if (rdf_due()) {
         float *rdf_base = rdf_store(mdb_rdf_size(G.eng, control.nbins));
         if (rdf_base) {
            std::vector<unsigned long long> cnt(mdb_rdf_size(G.eng, control.nbins), 0ULL);
            for (int r = 0; r < mdb_group_size(G.group); r++) {
               mdb_engine *er = mdb_group_engine(G.group, r);
               cudaSetDevice(er->device);
               if (mdb_rdf_counts(er, control.limit, control.nbins, cnt.data(), mdb_group_stream(G.group, r)))
                  FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
            }
            cudaSetDevice(G.eng->device);
            rdf_add_counts(system, rdf_base, cnt);
         }
      }
      return;
   }

   auto launch_real = [&]() {
      G.sites_fresh = false;
      push_sites(site);
      if (control.molpbc && mdb_set_com_host(G.eng, &system->c_of_m[0][0], G.stream))
         FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
      if (mdb_zero_out(G.eng, G.d_out, G.stream) || mdb_build_cells(G.eng, G.stream) ||
          mdb_force_real(G.eng, G.d_out, G.stream))
         FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
      // the result block goes home on the copy stream, so that the k-space kernels started below do not queue behind it
      cudaEventRecord(G.ev_real, G.stream);
      cudaStreamWaitEvent(G.copy_stream, G.ev_real, 0);
      if (cudaMemcpyAsync(G.h_out, G.d_out, sizeof(double) * mdb_out_doubles(n), cudaMemcpyDeviceToHost, G.copy_stream) != cudaSuccess)
         FATAL_MSG("libmoldy_b200: D2H copy failed");
      cudaEventRecord(G.ev_copied, G.copy_stream);
   };
   launch_real();
   if (chg_changed_late(chg, n)) {                       /* compared while the kernels run; never in a Moldy run */
      cudaStreamSynchronize(G.stream);
      cudaStreamSynchronize(G.copy_stream);
      sync_config(system, species, chg, potpar);
      launch_real();
   }
   int pr[2];
   const int tc = mdb_too_close(G.eng, pr, G.stream);     /* synchronises G.stream: the real-space kernels are done */
   G.sites_fresh = true;

   // eval_forces calls ewald() next on the same sites (src/accel.c:520-527): start its kernels now, so
   // that they run while this thread adds the real-space block into the caller's arrays; ewald() then
   // only waits for them.  Discarded if the next ewald() comes with other sites, partition or tables.
   G.ahead_valid = false;
   if (control.alpha > MDB_ALPHAMIN && !getenv("MOLDY_B200_NO_AHEAD")) {
      if (!G.ev_ahead) cudaEventCreateWithFlags(&G.ev_ahead, cudaEventDisableTiming);
      if (G.kf_slices < 0) {
         G.kf_slices = getenv("MOLDY_B200_KF_SLICES") ? atoi(getenv("MOLDY_B200_KF_SLICES")) : 4;
         G.kf_min_sites = 200000;
      }
      mdb_set_kforce_slices(G.eng, n >= G.kf_min_sites ? G.kf_slices : 1);
      const int rc_recip = mdb_zero_out(G.eng, G.d_out2, G.stream) || mdb_force_recip(G.eng, G.d_out2, G.stream);
      mdb_set_kforce_slices(G.eng, 1);
      if (rc_recip) FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
      void *evs[8];
      G.ahead_nslice = mdb_kforce_slices(G.eng, evs, G.ahead_hi);
      if (G.ahead_nslice > 1) {
         // slice k of the k-space forces goes home behind its event while slice k+1 is computed; ewald() adds them one by one
         size_t lo = 0;
         for (int k = 0; k < G.ahead_nslice; k++) {
            const size_t hi = (size_t)G.ahead_hi[k];
            cudaStreamWaitEvent(G.copy_stream, (cudaEvent_t)evs[k], 0);
            for (int a = 0; a < 3 && hi > lo; a++)
               cudaMemcpyAsync(G.h_out2 + (size_t)a * n + lo, G.d_out2 + (size_t)a * n + lo, sizeof(double) * (hi - lo),
                               cudaMemcpyDeviceToHost, G.copy_stream);
            if (k == G.ahead_nslice - 1)
               cudaMemcpyAsync(G.h_out2 + 3 * (size_t)n, G.d_out2 + 3 * (size_t)n, sizeof(double) * (mdb_out_doubles(n) - 3 * (size_t)n),
                               cudaMemcpyDeviceToHost, G.copy_stream);
            if (!G.ev_slice[k]) cudaEventCreateWithFlags(&G.ev_slice[k], cudaEventDisableTiming);
            cudaEventRecord(G.ev_slice[k], G.copy_stream);
            lo = hi;
         }
         cudaStreamWaitEvent(G.stream, G.ev_slice[G.ahead_nslice - 1], 0);
      } else {
         G.ahead_nslice = 0;
         cudaMemcpyAsync(G.h_out2, G.d_out2, sizeof(double) * mdb_out_doubles(n), cudaMemcpyDeviceToHost, G.stream);
      }
      cudaEventRecord(G.ev_ahead, G.stream);
      G.ahead_valid = true; G.ahead_sites = (const void *)site[0]; G.ahead_epoch = G.config_epoch;
      G.ahead_ithread = ithread; G.ahead_nthreads = nthreads;
   }

   if (cudaEventSynchronize(G.ev_copied) != cudaSuccess) FATAL_MSG("libmoldy_b200: D2H copy failed");
   pull_and_accumulate(site_force, pe, stress, n, G.d_out, G.h_out, true);
   *pe += G.h_out[3 * (size_t)n];

   if (tc & (1 << 30))
      message((int *)0, (char *)0, SEV_ERROR, (char *)"Co-ordinate out of range in BIN (fill_cells)");
   if (tc & ~(1 << 30))                                  /* src/force.c:944-946 */
      message((int *)0, (char *)0, SEV_WARNING, (char *)"Sites %d and %d closer than %fA.", pr[0], pr[1],
              sqrt(MDB_TOO_CLOSE));

   if (rdf_due()) {
      float *rdf_base = rdf_store(mdb_rdf_size(G.eng, control.nbins));
      if (rdf_base) {
         std::vector<unsigned long long> cnt(mdb_rdf_size(G.eng, control.nbins), 0ULL);
         if (mdb_rdf_counts(G.eng, control.limit, control.nbins, cnt.data(), G.stream))
            FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
         rdf_add_counts(system, rdf_base, cnt);
      }
   }
}

// First-call constants and notes of ewald (src/ewald.c:367-425)
static void recip_first_call(system_mp system, spec_mt *species, const real *chg, double vol)
{
   const int n = system->nsites;
   if (!G.recip_init) {                                  /* src/ewald.c:367-425 */
      double sqsq = 0, sq = 0, last_intra = 0;
      int ssite = 0;
      spec_mt *sp = species;
      while (sp < species + system->nspecies && !sp->framework) {
         double intra = 0.0;
         for (int is = 0; is < sp->nsites; is++)
            for (int js = is + 1; js < sp->nsites; js++) {
               const double *a = sp->p_f_sites[is], *b = sp->p_f_sites[js];
               double r = sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) +
                               (a[2] - b[2]) * (a[2] - b[2]));
               intra += chg[ssite + is] * chg[ssite + js] * mdb_err_fn(control.alpha * r) / r;
            }
         G.self_energy += sp->nmols * intra;
         ssite += sp->nsites * sp->nmols;
         sp++;
      }
      const int nsitesxf = ssite;
      const bool frame = sp != species + system->nspecies;
      int is = 0;
      for (; is < nsitesxf; is++) { sq += chg[is]; sqsq += chg[is] * chg[is]; }
      G.self_energy += control.alpha / sqrt(MDB_PI) * sqsq;
      const double sqxf = sq;
      for (; is < n; is++) sq += chg[is];
      if (frame) {
         G.sheet_energy = MDB_PI * (sq - sqxf) * (sq - sqxf) / (2.0 * control.alpha * control.alpha);
         message((int *)0, (char *)0, SEV_INFO,
                 (char *)"Framework has net electric charge of %.2g - correction of %g kJ/mol added",
                 (sq - sqxf) * CONV_Q_E, G.sheet_energy / vol * CONV_E_KJ);
      }
      if (fabs(sq) * CONV_Q_E > 1.0e-5) {
         last_intra = MDB_PI * sq * sq / (2.0 * control.alpha * control.alpha);
         G.sheet_energy -= last_intra;
         message((int *)0, (char *)0, SEV_WARNING,
                 (char *)"System has net electric charge of %.2g - correction of %g kJ/mol added", sq * CONV_Q_E,
                 last_intra / vol * CONV_E_KJ);
      }
      note((char *)"Ewald self-energy = %f kJ/mol", G.self_energy * CONV_E_KJ);
      note((char *)"%d K-vectors included in reciprocal-space sum", mdb_n_kvectors(G.eng));
      G.recip_init = true;
   }
}

extern "C" void ewald(real **site, real **site_force, system_mp system, spec_mt *species, real *chg, double *pe,
                      real (*stress)[3])
{
   const double tc0 = now_ms();
   sync_config(system, species, chg, nullptr);
   if (g_timing) fprintf(stderr, "[moldy_b200] ewald: sync_config %.2f ms\n", now_ms() - tc0);
   mdb_set_partition(G.eng, ithread, nthreads);
   const int n = system->nsites;
   double h9[9];
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) h9[3 * i + j] = system->h[i][j];
   const double vol = mdb_det3(h9);

   recip_first_call(system, species, chg, vol);
   if (ithread == 0) {                                   /* src/ewald.c:427-433 */
      *pe -= G.self_energy;
      *pe += G.sheet_energy / vol;
      for (int i = 0; i < 3; i++) stress[i][i] += G.sheet_energy / vol;
   }

   if (G.group) {
      group_force(site, site_force, system, 2, pe, stress);
      return;
   }
   bool ahead = G.ahead_valid && G.sites_fresh && G.ahead_sites == (const void *)site[0] &&
                G.ahead_epoch == G.config_epoch && G.ahead_ithread == ithread && G.ahead_nthreads == nthreads;
   G.ahead_valid = false;
   if (ahead) {
      /* ewald(site,...) must compute from the sites it is passed (src/ewald.c:280): the rows go up again on the copy
       * stream while the kernels started by force_calc run (G.d_out is free by now) and are compared bit for bit on
       * the device with the sites those kernels used; any difference discards the look-ahead result. */
      if (!G.check_stream) cudaStreamCreateWithFlags(&G.check_stream, cudaStreamNonBlocking);
      const long nd = mdb_sites_differ_host(G.eng, site[0], site[1], site[2], G.d_out, G.check_stream);
      if (nd != 0) {
         if (g_timing) fprintf(stderr, "[moldy_b200] ewald: sites changed since force_calc (%ld values), recomputing\n", nd);
         cudaEventSynchronize(G.ev_ahead);
         ahead = false;
      }
   }
   if (ahead && G.ahead_nslice > 1) {                    /* started by force_calc in slices: add each as it arrives */
      const double t0 = now_ms();
      size_t lo = 0;
      for (int k = 0; k < G.ahead_nslice; k++) {
         if (cudaEventSynchronize(G.ev_slice[k]) != cudaSuccess) FATAL_MSG("libmoldy_b200: k-space kernels failed");
         accumulate_rows(site_force, G.h_out2, n, lo, (size_t)G.ahead_hi[k]);
         lo = (size_t)G.ahead_hi[k];
      }
      if (cudaEventSynchronize(G.ev_ahead) != cudaSuccess) FATAL_MSG("libmoldy_b200: k-space kernels failed");
      if (g_timing) fprintf(stderr, "[moldy_b200] ewald: %d slices of the look-ahead k-space sum added in %.2f ms\n", G.ahead_nslice, now_ms() - t0);
      G.sites_fresh = false;
      const double *sc = G.h_out2 + 3 * (size_t)n;
      stress[0][0] += sc[2]; stress[0][1] += sc[3]; stress[0][2] += sc[4];
      stress[1][1] += sc[6]; stress[1][2] += sc[7]; stress[2][2] += sc[10];
      *pe += sc[1];
      return;
   }
   if (ahead) {                                          /* started by force_calc: wait and add */
      const double t0 = now_ms();
      if (cudaEventSynchronize(G.ev_ahead) != cudaSuccess) FATAL_MSG("libmoldy_b200: k-space kernels failed");
      if (g_timing) fprintf(stderr, "[moldy_b200] ewald: waited %.2f ms for the kernels started by force_calc\n", now_ms() - t0);
      G.sites_fresh = false;
      pull_and_accumulate(site_force, pe, stress, n, G.d_out2, G.h_out2, true);
      *pe += G.h_out2[3 * (size_t)n + 1];
      return;
   }
   push_sites(site);
   G.sites_fresh = false;
   if (mdb_zero_out(G.eng, G.d_out, G.stream) || mdb_force_recip(G.eng, G.d_out, G.stream))
      FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
   pull_and_accumulate(site_force, pe, stress, n);
   *pe += G.h_out[3 * (size_t)n + 1];
}

extern "C" void kernel(int jmin, int nnab, real *forceij, double *pe, real *r_sqr, real *nab_chg, double chg,
                       double norm, double alpha, int ptype, real **pot)
{
   ensure_engine();
   if (ptype < 0 || ptype > 6 || ptype == 5)
      FATAL_MSG("KERNEL called with unknown potential type %d", ptype);   /* src/kernel.c:186 */
   if (mdb_launch_kernel_vec(jmin, nnab, forceij, pe, r_sqr, nab_chg, chg, norm, alpha, ptype, pot))
      FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
}

extern "C" double poteval(real *potpar, double r, int ptype, double chgsq)
{
   double pe = 0.0, norm = 2.0 * control.alpha / sqrt(MDB_PI);
   real chgsq_r = chgsq, f, rr = r * r;
   real *pp[MDB_NPOTP];
   for (int i = 0; i < MDB_NPOTP; i++) pp[i] = potpar + i;
   kernel(0, 1, &f, &pe, &rr, &chgsq_r, 1.0, norm, control.alpha, ptype, pp);
   return pe;
}

// Closed forms of -int_rc^inf r^2 U(r) dr per potential (initialisation only).
extern "C" double dist_pot(real *p, double rc, int ptype)
{
   const double tol = 1.0e-7;
   const double rc2 = rc * rc, rc3 = rc2 * rc;
   auto exp_tail = [&](double b) { return rc2 / b + 2 * rc / (b * b) + 2.0 / (b * b * b); };
   switch (ptype) {
      default:
         FATAL_MSG("KERNEL called with unknown potential type %d", ptype);
         return 0.0;                                       /* (a host whose message() returns) */
      case 0: { double s2 = p[1] * p[1] / rc; return p[0] * s2 * s2 * s2 / 3.0; }
      case 1:
         if (p[2] > tol) return p[0] / (3.0 * rc3) - p[1] * exp(-p[2] * rc) * exp_tail(p[2]);
         return p[0] / (3.0 * rc3);
      case 2:
         if (p[3] > tol) return p[2] * exp_tail(p[3]) * exp(-p[3] * rc);
         return 0.0;
      case 3: {
         double tail = -p[2] / (9.0 * rc3 * rc3 * rc3) + p[3] / rc + p[4] / (3.0 * rc3) + p[5] / (5.0 * rc2 * rc3);
         if (p[1] > tol) return -p[0] * exp(-p[1] * rc) * exp_tail(p[1]) + tail;
         return tail;
      }
      case 6:
         if (p[5] != 0.0) return p[3] / (3.0 * rc3) + 2.0 * p[4] * exp_tail(p[5]) * exp(-p[5] * (rc - p[6]));
         return p[3] / (3.0 * rc3);
      case 4:
         return -p[0] / rc - p[1] / rc3 / 3.0 - p[2] / (rc3 * rc3 * rc3) / 9.0;
   }
}

// ---- eval_forces(): the whole of src/accel.c:398-617 with the sites and site forces resident in HBM -------------
// distant_const (src/accel.c:293-326): -2 pi sum_ij Ni Nj A_ij(rc) [+ 2/3 pi Ni Nj rc^3 U_ij(rc) for the pressure term]
static double distant_const_abi(system_mp system, spec_mt *species, pot_mt *potpar, double cutoff, int iflag)
{
   std::vector<int> count(system->max_id, 0);
   for (spec_mt *sp = species; sp < species + system->nspecies; sp++)
      for (int is = 0; is < sp->nsites; is++) count[sp->site_id[is]] += sp->nmols;
   double c = 0.0;
   for (int id = 1; id < system->max_id; id++)
      for (int jd = 1; jd < system->max_id; jd++) {
         c -= 2 * MDB_PI * count[id] * count[jd] * dist_pot(potpar[id + system->max_id * jd].p, cutoff, system->ptype);
         if (iflag)
            c += 2.0 / 3.0 * MDB_PI * count[id] * count[jd] * (cutoff * cutoff * cutoff) *
                 poteval(potpar[id + system->max_id * jd].p, cutoff, system->ptype, 0.0);
      }
   return c;
}

struct EvalState {
   bool init = false;
   double dist = 0, distp = 0;
   std::vector<double> chg, pfs, site_charge;     // site_charge: site_info[].charge the chg array was expanded from
   std::vector<int> layout;                        // nmols, nsites and site ids of every species, ditto
   std::vector<mdb_species> sp;
   long species_epoch = -1, chg_epoch = -1;
};
static EvalState E;

#define CONV_P_MPA (1.6605402e-27 / (1.0e-10 * 1.0e-12 * 1.0e-12) / 1.0e6)   /* src/defs.h:231 CONV_P */

// Internal linkage on purpose: when the host program carries its own `eval_forces` symbol (the trampoline of
// evalf_tramp.c), a call to the exported name from inside the library would bind to the program's definition and
// recurse; both exported names below call this function directly.
// configuration, first-call constants and notes, species table: everything eval_forces() does before the sites are built
static void eval_prepare(system_mp sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, double h9[9], double &vol,
                         bool &do_recip)
{
   const int n = sys->nsites, nspecies = sys->nspecies;
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) h9[3 * i + j] = sys->h[i][j];
   vol = mdb_det3(h9);

   /* site charges, src/accel.c:473-480; expanded again only when site_info or the species layout changed */
   const double tc0 = now_ms();
   bool chg_cached = (int)E.chg.size() == n && (int)E.site_charge.size() == sys->max_id;
   {
      std::vector<int> layout;
      for (spec_mt *sp = species; sp < species + nspecies; sp++) {
         layout.push_back(sp->nmols); layout.push_back(sp->nsites);
         for (int is = 0; is < sp->nsites; is++) layout.push_back(sp->site_id[is]);
      }
      for (int id = 0; chg_cached && id < sys->max_id; id++) chg_cached = site_info[id].charge == E.site_charge[id];
      chg_cached = chg_cached && layout == E.layout;
      if (!chg_cached) {
         E.chg.resize(n);
         double *c = E.chg.data();
         for (spec_mt *sp = species; sp < species + nspecies; sp++)
            for (int im = 0; im < sp->nmols; im++)
               for (int is = 0; is < sp->nsites; is++) *c++ = site_info[sp->site_id[is]].charge;
         E.site_charge.resize(sys->max_id);
         for (int id = 0; id < sys->max_id; id++) E.site_charge[id] = site_info[id].charge;
         E.layout.swap(layout);
      }
   }
   sync_config(sys, species, E.chg.data(), potpar, false, chg_cached && E.chg_epoch == G.chg_epoch);
   E.chg_epoch = G.chg_epoch;
   const double tc1 = now_ms();
   mdb_set_partition(G.eng, 0, 1);            /* complete sums: the library stands for all of the SPMD ranks */

   if (!E.init) {                             /* src/accel.c:444-453 */
      E.dist = distant_const_abi(sys, species, potpar, control.cutoff, 0);
      E.distp = distant_const_abi(sys, species, potpar, control.cutoff, 1);
      note((char *)"Distant potential correction = %f, Pressure correction = %f", CONV_E_KJ * E.dist / vol,
           CONV_P_MPA * E.distp / (vol * vol));
      E.init = true;
   }
   do_recip = control.alpha > MDB_ALPHAMIN;
   real_first_call(sys, species, E.chg.data(), potpar);
   if (do_recip) recip_first_call(sys, species, E.chg.data(), vol);

   if (E.species_epoch != G.config_epoch || (int)E.sp.size() != nspecies) {
      // the engine was reconfigured (every step under constant-stress dynamics: the cell matrix changes); the species table
      // and its device buffers depend on the system definition only and are rebuilt only when that changed
      std::vector<mdb_species> sp_new(nspecies);
      std::vector<double> pfs_new;
      for (int i = 0; i < nspecies; i++) {
         const spec_mt &sp = species[i];
         sp_new[i] = mdb_species{sp.nmols, sp.nsites, sp.framework ? 1 : 0, sp.quat ? 1 : 0, sp.rdof};
         for (int is = 0; is < sp.nsites; is++)
            for (int k = 0; k < 3; k++) pfs_new.push_back(sp.p_f_sites[is][k]);
      }
      const bool same = sp_new.size() == E.sp.size() && pfs_new == E.pfs &&
                        !memcmp(sp_new.data(), E.sp.data(), sizeof(mdb_species) * sp_new.size());
      if (!same) {
         E.sp.swap(sp_new); E.pfs.swap(pfs_new);
         if (G.group ? mdb_group_set_species(G.group, nspecies, E.sp.data(), E.pfs.data())
                     : mdb_set_species(G.eng, nspecies, E.sp.data(), E.pfs.data()))
            FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
      }
      E.species_epoch = G.config_epoch;
   }
   if (g_timing) fprintf(stderr, "[moldy_b200] eval_forces: config %.2f ms\n", tc1 - tc0);
}

// energies, dipole moment and the molecular virial stress from the scalar block of the device
// (src/force.c:1170, src/ewald.c:427-433, src/accel.c:557, :576-601, :606-608)
static void eval_finish_scalars(const double *sc, double vol, bool do_recip, double *pe, real *dip_mom, mat_mt stress)
{
   pe[0] = sc[12] - G.eintra + E.dist / vol;
   pe[1] = 0.0;
   for (int i = 0; i < 3; i++) dip_mom[i] = 0.0;
   memset(&stress[0][0], 0, sizeof(double) * 9);
   stress[0][0] = sc[14]; stress[0][1] = sc[15]; stress[0][2] = sc[16];
   stress[1][1] = sc[18]; stress[1][2] = sc[19]; stress[2][2] = sc[22];
   if (do_recip) {
      pe[1] = sc[13] - G.self_energy + G.sheet_energy / vol;
      for (int i = 0; i < 3; i++) stress[i][i] += G.sheet_energy / vol;
      for (int i = 0; i < 3; i++) dip_mom[i] = sc[i];
      if (control.surface_dipole)
         pe[1] += 2.0 * MDB_PI / (3.0 * vol) * (sc[0] * sc[0] + sc[1] * sc[1] + sc[2] * sc[2]);
   }
   for (int i = 0; i < 3; i++)
      for (int j = i + 1; j < 3; j++) stress[j][i] = stress[i][j];
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) stress[i][j] -= sc[3 + 3 * i + j];
   for (int i = 0; i < 3; i++) stress[i][i] += E.distp / vol;
}

static void report_too_close(int tc, const int pr[2])
{
   if (tc & (1 << 30))
      message((int *)0, (char *)0, SEV_ERROR, (char *)"Co-ordinate out of range in BIN (fill_cells)");
   if (tc & ~(1 << 30))                                  /* src/force.c:944-946 */
      message((int *)0, (char *)0, SEV_WARNING, (char *)"Sites %d and %d closer than %fA.", pr[0], pr[1],
              sqrt(MDB_TOO_CLOSE));
}

static void eval_forces_impl(system_mp sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, double *pe,
                             real *dip_mom, mat_mt stress, vec_mp *force, vec_mp *torque)
{
   const int nspecies = sys->nspecies;
   double h9[9], vol;
   bool do_recip;
   const double tc1 = now_ms();
   eval_prepare(sys, species, site_info, potpar, h9, vol, do_recip);
   std::vector<const double *> com(nspecies), quat(nspecies);
   for (int i = 0; i < nspecies; i++) {
      com[i] = &species[i].c_of_m[0][0];
      quat[i] = species[i].quat ? &species[i].quat[0][0] : nullptr;
   }
   G.sites_fresh = false; G.ahead_valid = false;          /* the engine's sites are no longer the ones force_calc uploaded */
   std::vector<unsigned long long> rdf_cnt;
   float *rdf_base = nullptr;
   if (rdf_due() && (rdf_base = rdf_store(mdb_rdf_size(G.eng, control.nbins))) != nullptr) {
      rdf_cnt.assign(mdb_rdf_size(G.eng, control.nbins), 0ULL);   /* binned between the force sums and the second make_sites */
      if (!G.group) mdb_eval_request_rdf(G.eng, control.limit, control.nbins, rdf_cnt.data());
   }
   int pr[2] = {0, 0}, tc = 0;
   if (G.group) {
      if (mdb_group_eval_forces_host(G.group, h9, com.data(), quat.data(), control.surface_dipole ? 1 : 0, do_recip ? 1 : 0, nullptr,
                                     control.limit, control.nbins, rdf_base ? rdf_cnt.data() : nullptr, &tc, pr))
         FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
      cudaSetDevice(G.eng->device);
   } else if (mdb_eval_forces_host(G.eng, h9, com.data(), quat.data(), control.surface_dipole ? 1 : 0, do_recip ? 1 : 0,
                                   nullptr, G.stream))
      FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
   const double tc2 = now_ms();
   if (rdf_base) rdf_add_counts(sys, rdf_base, rdf_cnt);

   if (!G.group) tc = mdb_too_close(G.eng, pr, G.stream);
   report_too_close(tc, pr);

   /* molecular forces and torques, src/accel.c:564-571 */
   const double *res = G.group ? mdb_group_eval_result(G.group) : mdb_eval_result(G.eng);   /* pinned: copied straight into the caller's arrays */
   int nmols = 0, nmols_r = 0;
   for (int i = 0; i < nspecies; i++) nmols += species[i].nmols;
   for (int i = 0; i < nspecies; i++) nmols_r += species[i].rdof > 0 ? species[i].nmols : 0;
   {
      struct Job { double *dst; const double *src; size_t n; };
      std::vector<Job> jobs;
      const double *f = res, *t = res + 3 * (size_t)nmols;
      for (int i = 0; i < nspecies; i++) {
         jobs.push_back({&force[i][0][0], f, 3 * (size_t)species[i].nmols});
         f += 3 * (size_t)species[i].nmols;
         if (species[i].rdof > 0) {
            jobs.push_back({&torque[i][0][0], t, 3 * (size_t)species[i].nmols});
            t += 3 * (size_t)species[i].nmols;
         }
      }
      auto run = [&](int part, int nparts) {
         for (auto &j : jobs) {
            const size_t lo = j.n * part / nparts, hi = j.n * (part + 1) / nparts;
            memcpy(j.dst + lo, j.src + lo, sizeof(double) * (hi - lo));
         }
      };
      if (nmols >= 65536) {                   /* 12 MB at 10^6 sites: four threads */
         std::thread th[3];
         for (int k = 1; k < 4; k++) th[k - 1] = std::thread(run, k, 4);
         run(0, 4);
         for (auto &th_k : th) th_k.join();
      } else {
         run(0, 1);
      }
   }
   const double *sc = res + 3 * (size_t)nmols + 3 * (size_t)nmols_r;
   eval_finish_scalars(sc, vol, do_recip, pe, dip_mom, stress);
   if (g_timing)
      fprintf(stderr, "[moldy_b200] eval_forces: device (H2D .. D2H) %.2f ms, results %.2f ms\n", tc2 - tc1, now_ms() - tc2);
}

// ---- do_step(): the NVE leapfrog step of src/accel.c:626-827 around the device's eval_forces ----------
// The dynamic state is uploaded from and written back to the host program's arrays on every call, so rescaling, output,
// dumps and restarts of the host program see and may change it between steps; hosts that keep the state in HBM use
// mdb_md_step directly.  Thermostat and cell dynamics are outside this row: FATAL (INTEGRATION.md section 6 shows how a
// host keeps its own do_step for those ensembles).
#define KB_PROG (1.380658e-23 / (1.6605402e-27 * 1.0e4))                 /* src/defs.h:201-226  kB = _kB / EUNIT */
extern "C" __attribute__((weak)) void dump(system_mp, spec_mt *, vec_mt *, vec_mt *, mat_mt, double, void *, int);

struct StepState { std::vector<mdb_species_dyn> dyn; long species_epoch = -1; int nosym = -1; double saved_pe = 0; };
static StepState S;

static void do_step_impl(system_mt *sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, vec_mt (*meansq_f_t)[2],
                         double *pe, real *dip_mom, mat_mt stress_vir, void *restart_header, int backup_restart, int init_H_0)
{
   if (control.const_temp || control.const_pressure)
      FATAL_MSG("libmoldy_b200: do_step on the device covers NVE dynamics only (const-temp=0, const-pressure=0)");
   const int nspecies = sys->nspecies;
   double h9[9], vol;
   bool do_recip;
   eval_prepare(sys, species, site_info, potpar, h9, vol, do_recip);
   std::vector<mdb_species_dyn> dyn(nspecies);
   for (int i = 0; i < nspecies; i++)
      dyn[i] = mdb_species_dyn{species[i].mass, {species[i].inertia[0], species[i].inertia[1], species[i].inertia[2]}};
   if (S.species_epoch != E.species_epoch || S.nosym != control.nosymmetric_rot || dyn.size() != S.dyn.size() ||
       memcmp(dyn.data(), S.dyn.data(), sizeof(mdb_species_dyn) * dyn.size())) {
      if (G.group ? mdb_group_md_set_dynamics(G.group, dyn.data(), control.nosymmetric_rot ? 1 : 0)
                  : mdb_md_set_dynamics(G.eng, dyn.data(), control.nosymmetric_rot ? 1 : 0))
         FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
      S.dyn = dyn; S.species_epoch = E.species_epoch; S.nosym = control.nosymmetric_rot;
   }
   std::vector<const double *> com(nspecies), quat(nspecies), mom(nspecies), amom(nspecies);
   std::vector<double *> wcom(nspecies), wquat(nspecies), wmom(nspecies), wamom(nspecies), wf(nspecies, nullptr), wt(nspecies, nullptr);
   for (int i = 0; i < nspecies; i++) {
      wcom[i] = &species[i].c_of_m[0][0]; wmom[i] = &species[i].mom[0][0];
      wquat[i] = species[i].quat ? &species[i].quat[0][0] : nullptr;
      wamom[i] = species[i].quat && species[i].amom ? &species[i].amom[0][0] : nullptr;
      com[i] = wcom[i]; quat[i] = wquat[i]; mom[i] = wmom[i]; amom[i] = wamom[i];
   }
   G.sites_fresh = false; G.ahead_valid = false;
   std::vector<unsigned long long> rdf_cnt;
   float *rdf_base = nullptr;
   if (rdf_due() && (rdf_base = rdf_store(mdb_rdf_size(G.eng, control.nbins))) != nullptr) {
      rdf_cnt.assign(mdb_rdf_size(G.eng, control.nbins), 0ULL);
      if (!G.group) mdb_eval_request_rdf(G.eng, control.limit, control.nbins, rdf_cnt.data());
   }
   const bool want_h0 = control.istep == 1 || init_H_0;
   int pr[2];
   if (G.group) {                                          /* all GPUs of MOLDY_B200_DEVICES: every rank moves its molecules */
      if (mdb_group_md_upload_state(G.group, com.data(), quat.data(), mom.data(), amom.data()) ||
          mdb_group_md_step(G.group, h9, control.step, sys->ts, control.surface_dipole ? 1 : 0, do_recip ? 1 : 0, want_h0 ? 1 : 0,
                            nullptr, control.limit, control.nbins, rdf_base ? rdf_cnt.data() : nullptr))
         FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
      if (rdf_base) rdf_add_counts(sys, rdf_base, rdf_cnt);
      report_too_close(mdb_group_too_close(G.group, pr), pr);
   } else {
      if (mdb_md_upload_state(G.eng, com.data(), quat.data(), mom.data(), amom.data(), G.stream) ||
          mdb_md_step(G.eng, h9, control.step, sys->ts, control.surface_dipole ? 1 : 0, do_recip ? 1 : 0, want_h0 ? 1 : 0, nullptr,
                      G.stream))
         FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
      if (rdf_base) rdf_add_counts(sys, rdf_base, rdf_cnt);
      report_too_close(mdb_too_close(G.eng, pr, G.stream), pr);
   }
   const bool dump_due = control.dump_interval > 0 && control.dump_level != 0 && control.istep >= control.begin_dump &&
                         (control.istep - control.begin_dump) % control.dump_interval == 0 && ithread == 0 && dump;
   std::vector<double> fbuf, tbuf;
   if (dump_due) {                                         /* the host program's dump wants the forces and torques */
      fbuf.resize(3 * (size_t)sys->nmols); tbuf.resize(3 * (size_t)std::max(sys->nmols_r, 1));
      size_t fo = 0, to = 0;
      for (int i = 0; i < nspecies; i++) {
         wf[i] = fbuf.data() + fo; fo += 3 * (size_t)species[i].nmols;
         if (species[i].rdof > 0) { wt[i] = tbuf.data() + to; to += 3 * (size_t)species[i].nmols; }
      }
   }
   if (G.group ? mdb_group_md_download_state(G.group, wcom.data(), wquat.data(), wmom.data(), wamom.data(), wf.data(), wt.data())
               : mdb_md_download_state(G.eng, wcom.data(), wquat.data(), wmom.data(), wamom.data(), wf.data(), wt.data(), G.stream))
      FATAL_MSG("libmoldy_b200: %s", mdb_last_error());
   const double *r = G.group ? mdb_group_md_result(G.group) : mdb_md_result(G.eng);
   const size_t nscal = G.group ? mdb_group_md_scalars(G.group) : mdb_md_scalars(G.eng);
   if (r[nscal - 1] != 0.0)
      FATAL_MSG("Quaternion %d (%g,%g,%g,%g) - normalisation error in beeman", 0, 0.0, 0.0, 0.0, 0.0);   /* src/leapfrog.c:107 */
   eval_finish_scalars(r, vol, do_recip, pe, dip_mom, stress_vir);
   S.saved_pe = pe[0] + pe[1];
   const double s2 = sys->ts * sys->ts;
   auto species_ke = [&](const double *sum, const spec_mt &sp, bool rot) {      /* trans_ke + rot_ke, src/algorith.c:221-257 */
      double ke = (sum[0] + sum[3] + sum[5]) / (2.0 * sp.mass * s2);
      if (rot && sp.rdof > 0) {
         double k2 = 0.0;
         for (int k = 0; k < 3; k++)
            if (sp.inertia[k] > 1.0e-14) k2 += sum[6 + k] / sp.inertia[k];
         ke += 0.5 * k2 / s2;
      }
      return ke;
   };
   if (want_h0) {                                          /* src/accel.c:718-726, kinetic energy at the half step */
      double ke = 0.0;
      for (int i = 0; i < nspecies; i++)
         ke += species_ke(r + MDB_EVAL_SCALARS + MDB_MD_SUMS * ((size_t)nspecies + i), species[i], true);
      double kc = 0.0;
      if (sys->hmom)
         for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) kc += sys->hmom[i][j] * sys->hmom[i][j];
      sys->H_0 = ke + S.saved_pe + sys->tsmom * sys->tsmom / (2.0 * control.ttmass) + sys->d_of_f * KB_PROG * control.temp * log(sys->ts) +
                 0.5 / control.pmass * kc;
   }
   {  /* src/accel.c:784-790: without GL_THERM the final leapf_nose_therm() is not under `if (control.const_temp)`
       * (no braces), so the reference runs it in NVE too -- a no-op while tsmom = 0; kept for fidelity */
      const double scale = 1.0 + sys->tsmom * (0.5 * control.step) / (2.0 * control.ttmass);
      sys->tsmom /= scale;
      sys->ts *= scale * scale;
   }
   for (int i = 0; i < nspecies; i++) {                    /* mean_square, src/accel.c:806-812 */
      const double *sum = r + MDB_EVAL_SCALARS + MDB_MD_SUMS * (size_t)i;
      for (int k = 0; k < 3; k++) {
         meansq_f_t[i][0][k] = species[i].nmols > 0 ? sum[9 + k] / species[i].nmols : 0.0;
         meansq_f_t[i][1][k] = species[i].rdof > 0 && species[i].nmols > 0 ? sum[12 + k] / species[i].nmols : 0.0;
      }
   }
   if (dump_due)
      dump(sys, species, (vec_mt *)fbuf.data(), sys->nmols_r ? (vec_mt *)tbuf.data() : nullptr, stress_vir, pe[0] + pe[1],
           restart_header, backup_restart);
}

extern "C" void mdb_do_step_moldy(system_mt *sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, vec_mt (*meansq_f_t)[2],
                                  double *pe, real *dip_mom, mat_mt stress_vir, void *restart_header, int backup_restart,
                                  int init_H_0)
{
   do_step_impl(sys, species, site_info, potpar, meansq_f_t, pe, dip_mom, stress_vir, restart_header, backup_restart, init_H_0);
}
extern "C" void do_step(system_mt *sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, vec_mt (*meansq_f_t)[2], double *pe,
                        real *dip_mom, mat_mt stress_vir, void *restart_header, int backup_restart, int init_H_0)
{
   do_step_impl(sys, species, site_info, potpar, meansq_f_t, pe, dip_mom, stress_vir, restart_header, backup_restart, init_H_0);
}

// The same function under a name of the library's own: a host program that keeps its own (weakened) eval_forces can
// still reach this one through a one-line trampoline object (moldy_b200/csrc/evalf_tramp.c).
extern "C" void mdb_eval_forces_moldy(system_mp sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, double *pe,
                                      real *dip_mom, mat_mt stress, vec_mp *force, vec_mp *torque)
{
   eval_forces_impl(sys, species, site_info, potpar, pe, dip_mom, stress, force, torque);
}

extern "C" void eval_forces(system_mp sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, double *pe,
                            real *dip_mom, mat_mt stress, vec_mp *force, vec_mp *torque)
{
   eval_forces_impl(sys, species, site_info, potpar, pe, dip_mom, stress, force, torque);
}

// ---- test/bench accessors ------------------------------------------------------
extern "C" mdb_engine *mdb_abi_engine(void) { ensure_engine(); return G.eng; }
extern "C" int mdb_abi_devices(void) { ensure_engine(); return G.group ? mdb_group_size(G.group) : 1; }
// tests only: drop the engine(s) so that the next call re-reads MOLDY_B200_DEVICE(S)
extern "C" void mdb_abi_shutdown(void)
{
   if (G.stream) cudaStreamSynchronize(G.stream);
   if (G.group) { mdb_group_destroy(G.group); G.group = nullptr; G.eng = nullptr; }
   else if (G.eng) { mdb_destroy(G.eng); G.eng = nullptr; }
   if (G.d_out) cudaFree(G.d_out);
   if (G.h_out) cudaFreeHost(G.h_out);
   if (G.d_out2) cudaFree(G.d_out2);
   if (G.h_out2) cudaFreeHost(G.h_out2);
   G.d_out = G.h_out = G.d_out2 = G.h_out2 = nullptr; G.out_cap = 0;
   if (G.stream) { cudaStreamDestroy(G.stream); G.stream = nullptr; }
   if (G.copy_stream) { cudaStreamDestroy(G.copy_stream); G.copy_stream = nullptr; }
   if (G.check_stream) { cudaStreamDestroy(G.check_stream); G.check_stream = nullptr; }
   for (auto &ev : G.ev_slice) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
   G.have_cfg = false; G.type.clear(); G.chg.clear(); G.potflat.clear();
}
// tests: slices of the look-ahead k-space sum and the smallest system they are used for
extern "C" void mdb_abi_set_kf_slices(int nslices, int min_sites) { G.kf_slices = nslices; G.kf_min_sites = min_sites; }
extern "C" void *mdb_abi_stream(void) { ensure_engine(); return (void *)G.stream; }
extern "C" void mdb_abi_constants(double out[3]) { out[0] = G.eintra; out[1] = G.self_energy; out[2] = G.sheet_energy; }
extern "C" void mdb_abi_reset(void)
{  // forget first-call state so one process can run several systems (tests only;
   // the reference needs a fresh process for that)
   G.real_init = G.recip_init = false;
   G.eintra = G.self_energy = G.sheet_energy = 0;
   G.onabor = G.onx = G.ony = G.onz = 0;
   G.type.clear(); G.mol.clear(); G.chg.clear(); G.potflat.clear();
   G.have_cfg = false; G.sites_fresh = false; G.last_sites = nullptr; G.rdf_warned = false;
   if (G.eng && G.stream) cudaStreamSynchronize(G.stream);
   G.ahead_valid = false;
   S.species_epoch = -1; S.dyn.clear();
   E.init = false; E.species_epoch = -1; E.sp.clear(); E.chg.clear(); E.site_charge.clear(); E.layout.clear();
}
