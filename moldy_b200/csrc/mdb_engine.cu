// mdb_engine.cu -- layer (B) of include/moldy_b200.h: the device-resident engine.
// Owns every HBM buffer of the hot path; all work is enqueued on the caller's
// stream.  HBM layout (N sites, C link cells):
//   x,y,z[N]            positions, original site order (caller's rows)
//   type[N], mol[N], chg[N], ptab[max_id^2][8]          static per system
//   cell[N], count[C+1], start[C+1], order[N]           link-cell tables
//   posq[N] (double4 x,y,z,q), stype[N], scell[N]       cell-sorted SoA, z-fastest cell order
//   runs[], hk[], slot tables, coef[nslots][8], ppart[slabs][nslots][4]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "mdb_internal.h"

#define FREE(p) do { if (p) { cudaFree(p); (p) = nullptr; } } while (0)

extern "C" size_t mdb_out_doubles(int nsites) { return 3 * (size_t)nsites + MDB_OUT_SCALARS; }

extern "C" mdb_engine *mdb_create(int device)
{
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) {
      mdb_set_error("mdb_create: no such CUDA device (libmoldy_b200 has no CPU path)");
      return nullptr;
   }
   if (cudaSetDevice(device) != cudaSuccess) {
      mdb_set_error("mdb_create: cudaSetDevice failed");
      return nullptr;
   }
   mdb_engine *e = new mdb_engine();
   e->device = device;
   if (cudaMalloc(&e->d_counters, 8 * sizeof(unsigned long long)) != cudaSuccess) {
      mdb_set_error("mdb_create: cudaMalloc failed");
      delete e;
      return nullptr;
   }
   cudaMemset(e->d_counters, 0, 8 * sizeof(unsigned long long));
   return e;
}

static void free_system(mdb_engine *e)
{
   e->table_cap.clear(); e->cls_uploaded = false;
   FREE(e->d_type); FREE(e->d_mol); FREE(e->d_chg); FREE(e->d_ptab);
   FREE(e->own_xyz); FREE(e->d_cell); FREE(e->d_order); FREE(e->d_posq);
   FREE(e->d_stype); FREE(e->d_scell); FREE(e->d_sinfo); FREE(e->d_fs); FREE(e->d_com); e->com_cap = 0; e->com_set = false;
   e->d_x = e->d_y = e->d_z = nullptr;
}
static void free_sublists(mdb_engine *e)
{
   for (int k = 0; k < 2; k++) {
      SubList &S = e->sub[k];
      FREE(S.posq); FREE(S.sinfo); FREE(S.order); FREE(S.start); FREE(S.batches); FREE(S.nbatch); FREE(S.fs);
      S.n = 0; S.batch_cap = 0; S.cap_n = 0; S.cap_cells = 0; S.valid = false;
   }
   FREE(e->d_cls); e->table_cap.erase((void *)&e->d_cls); e->cls_uploaded = false; FREE(e->d_sub_flag); FREE(e->d_sub_pos); FREE(e->d_sub_scan); FREE(e->d_sub_cols);
   e->sub_cap = 0; e->sub_cols_cap = 0;
   if (e->aux_stream) { cudaStreamDestroy(e->aux_stream); e->aux_stream = nullptr; }
   if (e->ev_cells) { cudaEventDestroy(e->ev_cells); e->ev_cells = nullptr; }
   if (e->ev_sub1) { cudaEventDestroy(e->ev_sub1); e->ev_sub1 = nullptr; }
}
static void free_grid(mdb_engine *e)
{
   FREE(e->d_count); FREE(e->d_start); FREE(e->d_scan_tmp); FREE(e->d_runs); FREE(e->d_runs_half);
   FREE(e->d_batches); FREE(e->d_nbatch); e->batch_cap = 0;
   FREE(e->d_runs_rdf); e->rdf_limit = -1.0;
   FREE(e->d_runs_near); FREE(e->d_runs_far); FREE(e->d_ptab_far); e->far_ptype = -1;
   e->table_cap.erase((void *)&e->d_runs_near); e->table_cap.erase((void *)&e->d_runs_far); e->table_cap.erase((void *)&e->d_ptab_far);
   e->table_cap.erase((void *)&e->d_runs); e->table_cap.erase((void *)&e->d_runs_half); e->table_cap.erase((void *)&e->d_runs_rdf);
   e->cells_cap = 0;
}
static void free_recip(mdb_engine *e)
{
   FREE(e->d_hk); FREE(e->d_hk_valid); FREE(e->d_slot_flags); FREE(e->d_ppart);
   FREE(e->d_coef_tot); FREE(e->d_coef_nf); FREE(e->d_kpartials); FREE(e->d_cidx); FREE(e->d_sfac_blocks); FREE(e->d_kf_groups); FREE(e->d_kf_slot_dst); FREE(e->d_ktab); e->ktab_cap = 0;
   FREE(e->d_psum);
   e->ppart_cap = 0; e->slots_cap = 0;
   e->table_cap.erase((void *)&e->d_hk); e->table_cap.erase((void *)&e->d_hk_valid);
   e->table_cap.erase((void *)&e->d_slot_flags); e->table_cap.erase((void *)&e->d_cidx);
}

extern "C" void mdb_destroy(mdb_engine *e)
{
   if (e && e->ovl_stream) {
      cudaSetDevice(e->device);
      cudaStreamDestroy(e->ovl_stream); cudaEventDestroy(e->ev_ovl_fork); cudaEventDestroy(e->ev_ovl_join);
      cudaFree(e->d_ovl_q); cudaFree(e->d_out2);
   }
   if (!e) return;
   cudaSetDevice(e->device);
   free_system(e); free_grid(e); free_recip(e); free_sublists(e);
   FREE(e->d_partials); FREE(e->d_counters); FREE(e->d_out_own); FREE(e->d_rdf);
   if (e->h_stage) cudaFreeHost(e->h_stage);
   FREE(e->mf.d_pfs); FREE(e->mf.d_in); FREE(e->mf.d_res); FREE(e->mf.d_vpart); FREE(e->mf.d_dpart);
   if (e->mf.h_in) cudaFreeHost(e->mf.h_in);
   if (e->mf.h_res) cudaFreeHost(e->mf.h_res);
   FREE(e->mf.d_mom); FREE(e->mf.d_amom); FREE(e->mf.d_mdpart); FREE(e->mf.d_mdscal);
   if (e->mf.h_mdscal) cudaFreeHost(e->mf.h_mdscal);
   if (e->mf.h_state) cudaFreeHost(e->mf.h_state);
   for (auto &ev : e->kf_ev) if (ev) cudaEventDestroy(ev);
   delete e;
}

extern "C" void mdb_set_pair_mode(mdb_engine *e, int mode)
{
   e->pair_mode = (mode == 2 || mode == 3) ? mode : 4;
   if (e->pair_mode == 2) e->pair_split = 0;
   if (e->configured && sizeof(double) * e->h_potpar.size() > MDB_TILED_TAB_MAX) e->pair_mode = 2;
   e->cells_valid = false;
}

// far stencil runs of the exponential potentials without the exponentials (default on; MDB_PAIR_FAR=0); takes effect at the
// next mdb_configure.  mdb_pair_far_runs: how many of the half list's runs the current configuration treats that way.
extern "C" void mdb_set_pair_far(mdb_engine *e, int on) { e->far_enable = on ? 1 : 0; }
extern "C" int mdb_pair_far_runs(const mdb_engine *e) { return e->far_ptype >= 0 ? e->nruns_far : 0; }

extern "C" void mdb_set_partition(mdb_engine *e, int ithread, int nthreads)
{
   e->ithread = ithread;
   e->nthreads = nthreads < 1 ? 1 : nthreads;
}

// (Re)fill a device table.  The allocation is kept while it is large enough: under constant-stress dynamics the cell
// matrix changes every step and mdb_configure runs every step -- cudaFree/cudaMalloc would synchronise the device each time.
template <class T>
static int upload(mdb_engine *e, T **dst, const T *src, size_t n)
{
   size_t &cap = e->table_cap[(void *)dst];
   if (*dst && sizeof(T) * n > cap) { cudaFree(*dst); *dst = nullptr; cap = 0; }
   if (n == 0) return 0;
   if (!*dst) {
      MDB_CUDA(cudaMalloc(dst, sizeof(T) * n));
      cap = sizeof(T) * n;
   }
   MDB_CUDA(cudaMemcpy(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice));
   return 0;
}

// ---- far stencil runs of the exponential potentials (see mdb_internal.h) ---------------------------------------------
// kernel.c's forms: Buckingham -p0/r^6 + p1 exp(-p2 r); generic p0 exp(-p1 r) + p2/r^12 - p3/r^4 - p4/r^6 - p5/r^8; Morse/BIG
// p0 exp((p1 - r) p2) - p3/r^6 + p4 (exp(-2 p5 (r - p6)) - 2 exp(-p5 (r - p6))); MCY p0 exp(-p1 r) - p2 exp(-p3 r)
// (src/kernel.c:230-355).  PT_HIW rows are p0/r^4 + p1/r^6 + p2/r^12.
static int mdb_split_far_runs(mdb_engine *e)
{
   const mdb_config &c = e->cfg;
   e->far_ptype = -1; e->nruns_near = e->nruns_far = 0; e->r_far = 0.0;
   if (e->far_enable < 0) e->far_enable = !(getenv("MDB_PAIR_FAR") && atoi(getenv("MDB_PAIR_FAR")) == 0);
   const int pt = c.ptype, id = c.max_id;
   if (!e->far_enable || c.molpbc || e->T.runs_half.empty() || !(pt == 1 || pt == 2 || pt == 3 || pt == 6)) return 0;
   std::vector<double> far(e->h_potpar.size(), 0.0);
   const double r_far = mdb_far_radius(pt, id, e->h_potpar.data(), far.data());
   if (!(r_far > 0.0)) return 0;
   // lower bound of the distance between a point of a cell and a point of the cell (dx,dy,dz) away: |h d| - longest diagonal
   const double inv[3] = {1.0 / e->T.nx, 1.0 / e->T.ny, 1.0 / e->T.nz};
   auto len = [&](double x, double y, double z) {
      double v[3];
      for (int i = 0; i < 3; i++) v[i] = c.h[3 * i] * x * inv[0] + c.h[3 * i + 1] * y * inv[1] + c.h[3 * i + 2] * z * inv[2];
      return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
   };
   double diag = 0.0;
   for (int sgn = 0; sgn < 4; sgn++) diag = std::max(diag, len(1.0, (sgn & 1) ? -1.0 : 1.0, (sgn & 2) ? -1.0 : 1.0));
   std::vector<StencilRun> nearv, farv;
   for (size_t r = 0; r < e->T.runs_half.size(); r++) {
      const StencilRun &run = e->T.runs_half[r];
      double dmin = 1e300;
      for (int dz = run.dzlo; dz <= run.dzhi; dz++) dmin = std::min(dmin, len(run.dx, run.dy, dz));
      if (r > 0 && dmin - diag >= r_far) farv.push_back(run); else nearv.push_back(run);      // run 0 is the prologue's
   }
   if (farv.empty()) return 0;
   bool rest = false;                              // anything left of the potential out there?
   for (double v : far) rest |= v != 0.0;
   if (upload(e, &e->d_runs_near, nearv.data(), nearv.size())) return -1;
   if (upload(e, &e->d_runs_far, farv.data(), farv.size())) return -1;
   if (upload(e, &e->d_ptab_far, far.data(), far.size())) return -1;
   e->nruns_near = (int)nearv.size(); e->nruns_far = (int)farv.size();
   e->far_ptype = rest ? 4 /* PT_HIW */ : 7 /* PT_NONE */;
   e->r_far = r_far;
   return 0;
}

extern "C" int mdb_configure(mdb_engine *e, const mdb_config *cfg)
{
   MDB_CUDA(cudaSetDevice(e->device));
   const int n = cfg->nsites;
   if (n <= 0 || cfg->max_id <= 0) { mdb_set_error("mdb_configure: empty system"); return -1; }
   if (cfg->ptype < 0 || cfg->ptype > 6 || cfg->ptype == 5) {
      mdb_set_error("KERNEL called with unknown potential type");
      return -1;
   }
   const bool new_system = !e->configured || e->cfg.nsites != n || e->cfg.max_id != cfg->max_id;
   // the system definition (site ids, molecule map, charges, potential parameters) is static in a run; only the cell
   // matrix and with it the tables below change under constant-stress dynamics: skip the per-site work then
   const int np_chk = cfg->max_id * cfg->max_id * MDB_NPOTP;
   const bool same_def = !new_system && e->cfg.ptype == cfg->ptype && e->cfg.nsites_xf == cfg->nsites_xf &&
                         (int)e->h_type.size() == n && !memcmp(e->h_type.data(), cfg->site_type, sizeof(int) * n) &&
                         (cfg->site_mol ? !memcmp(e->h_mol.data(), cfg->site_mol, sizeof(int) * n) : e->mol_is_identity) &&
                         !memcmp(e->h_chg.data(), cfg->chg, sizeof(double) * n) && (int)e->h_potpar.size() == np_chk &&
                         !memcmp(e->h_potpar.data(), cfg->potpar, sizeof(double) * np_chk);
   e->cfg = *cfg;
   e->cfg.site_type = nullptr; e->cfg.site_mol = nullptr; e->cfg.chg = nullptr; e->cfg.potpar = nullptr;
   e->cells_valid = false;

   // ---- static per-site data ----
   if (!same_def) {
      e->h_type.assign(cfg->site_type, cfg->site_type + n);
      if (cfg->site_mol) e->h_mol.assign(cfg->site_mol, cfg->site_mol + n);
      else { e->h_mol.resize(n); for (int i = 0; i < n; i++) e->h_mol[i] = i; }
      e->mol_is_identity = cfg->site_mol == nullptr;
      e->h_chg.assign(cfg->chg, cfg->chg + n);
      e->h_potpar.assign(cfg->potpar, cfg->potpar + np_chk);
      for (int i = 0; i < n; i++)
         if (e->h_type[i] < 0 || e->h_type[i] >= cfg->max_id) {
            mdb_set_error("mdb_configure: site id out of range");
            return -1;
         }
   }
   // pair table as the kernel wants it (LJ: sigma^2 and 6 eps precomputed, src/kernel.c:206-210)
   std::vector<double> ptab(e->h_potpar);
   if (cfg->ptype == 0)
      for (int k = 0; k < cfg->max_id * cfg->max_id; k++) {
         double *p = &ptab[(size_t)k * MDB_NPOTP];
         p[2] = 6.0 * p[0];
         p[1] = p[1] * p[1];
      }
   if (new_system) {
      free_system(e);
      MDB_CUDA(cudaMalloc(&e->own_xyz, sizeof(double) * 3 * (size_t)n));
      MDB_CUDA(cudaMalloc(&e->d_cell, sizeof(int) * (size_t)n));
      MDB_CUDA(cudaMalloc(&e->d_order, sizeof(int) * (size_t)n));
      MDB_CUDA(cudaMalloc(&e->d_posq, sizeof(double4) * (size_t)n));
      MDB_CUDA(cudaMalloc(&e->d_stype, sizeof(int) * (size_t)n));
      MDB_CUDA(cudaMalloc(&e->d_scell, sizeof(int) * (size_t)n));
      MDB_CUDA(cudaMalloc(&e->d_sinfo, sizeof(int2) * (size_t)n));
      e->sites_set = false;
   }
   if (!same_def) {
      if (upload(e, &e->d_type, e->h_type.data(), n)) return -1;
      if (upload(e, &e->d_mol, e->h_mol.data(), n)) return -1;
      if (upload(e, &e->d_chg, e->h_chg.data(), n)) return -1;
      if (upload(e, &e->d_ptab, ptab.data(), ptab.size())) return -1;
   }

   // ---- link-cell grid + stencil ----
   std::string err;
   if (!mdb_build_real_tables(e->cfg, e->T, err)) { mdb_set_error(err); return -1; }
   e->ncells = e->T.nx * e->T.ny * e->T.nz;
   if (std::max(e->ncells + 1, 2 * (e->T.nx * e->T.ny + 1)) > e->cells_cap) {
      free_grid(e);
      e->cells_cap = std::max(e->ncells + 1, 2 * (e->T.nx * e->T.ny + 1));
      MDB_CUDA(cudaMalloc(&e->d_count, sizeof(int) * (size_t)e->cells_cap));
      MDB_CUDA(cudaMalloc(&e->d_start, sizeof(int) * (size_t)e->cells_cap));
      MDB_CUDA(cudaMalloc(&e->d_scan_tmp, sizeof(int) * (size_t)(e->cells_cap / 2048 + 2)));
   }
   e->nruns = (int)e->T.runs.size();
   if (upload(e, &e->d_runs, e->T.runs.data(), e->T.runs.size())) return -1;
   e->nruns_half = (int)e->T.runs_half.size();
   if (upload(e, &e->d_runs_half, e->T.runs_half.data(), e->T.runs_half.size())) return -1;
   if (mdb_split_far_runs(e)) return -1;
   {
      const int need = n / MDB_NI + e->T.nx * e->T.ny + 8;
      if (need > e->batch_cap) {
         FREE(e->d_batches); FREE(e->d_nbatch);
         MDB_CUDA(cudaMalloc(&e->d_batches, sizeof(int2) * (size_t)need));
         MDB_CUDA(cudaMalloc(&e->d_nbatch, sizeof(int)));
         e->batch_cap = need;
      }
      if (e->pair_mode < 0) {
         const char *m = getenv("MDB_PAIR_MODE");
         const int mm = m ? atoi(m) : 4;
         e->pair_mode = (mm == 2 || mm == 3) ? mm : 4;
      }
      // the tiled kernel keeps the pair table in shared memory; very many site types -> per-thread kernel
      if (sizeof(double) * ptab.size() > MDB_TILED_TAB_MAX) e->pair_mode = 2;
      if (new_system) { FREE(e->d_fs); }
      if (!e->d_fs) MDB_CUDA(cudaMalloc(&e->d_fs, sizeof(double) * 3 * (size_t)n));
      for (auto &r : e->T.runs)
         if (r.dzlo < -500 || r.dzhi > 500) { mdb_set_error("stencil z extent > 500 cells"); return -1; }
   }
   // ---- site classes for the split pair passes.  A pair whose charge product is zero gets an exact
   // zero from the Coulomb term, one whose potential parameters all vanish an exact zero from kernel()'s
   // potential (src/kernel.c:182-461), so charged x charged pairs with the Coulomb term alone plus
   // potential x potential pairs with the potential alone give the reference's sums term for term.
   {
      const int id = cfg->max_id;
      auto amp_zero = [&](const double *p) {          // amplitudes of the potential, by type
         switch (cfg->ptype) {
            case 0: return p[0] == 0.0;                                              // eps
            case 1: return p[0] == 0.0 && p[1] == 0.0;                               // -A/r^6 + B exp(-C r)
            case 2: return p[0] == 0.0 && p[2] == 0.0;                               // A exp(-B r) - C exp(-D r)
            case 3: return p[0] == 0.0 && p[2] == 0.0 && p[3] == 0.0 && p[4] == 0.0 && p[5] == 0.0;
            case 4: return p[0] == 0.0 && p[1] == 0.0 && p[2] == 0.0;
            default: return p[0] == 0.0 && p[3] == 0.0 && p[4] == 0.0;               // Morse / BIG
         }
      };
      std::vector<char> active(id, 0);
      for (int a = 0; a < id; a++)
         for (int b = 0; b < id; b++)
            if (!amp_zero(&e->h_potpar[((size_t)a * id + b) * MDB_NPOTP])) active[a] = 1;
      std::vector<unsigned char> &cls = e->h_cls;
      if (!same_def || (int)cls.size() != n) {
         cls.resize(n);
         e->n_cls[0] = e->n_cls[1] = 0;
         for (int i = 0; i < n; i++) {
            cls[i] = (e->h_chg[i] != 0.0 ? 1 : 0) | (active[e->h_type[i]] ? 2 : 0);
            e->n_cls[0] += cls[i] & 1; e->n_cls[1] += (cls[i] >> 1) & 1;
         }
         e->cls_uploaded = false;
      }
      const long nc = e->n_cls[0], np2 = e->n_cls[1];
      // FP64 instructions per visit (profiles/): 22 shared (r^2, 1/r, accumulation), 23 Coulomb, 9..30 potential
      static const double ptc[7] = {9, 19, 27, 30, 12, 0, 40};
      const bool coul = cfg->alpha > 0.0;
      const double fused = (double)n * n * (22 + (coul ? 23 : 0) + ptc[cfg->ptype]);
      const double split = (double)nc * nc * (22 + 23) + (double)np2 * np2 * (22 + ptc[cfg->ptype]);
      const char *sp = getenv("MDB_PAIR_SPLIT");
      e->pair_split = coul && e->pair_mode >= 3 && (sp ? atoi(sp) != 0 : split < 0.85 * fused);
      // molecular cut-off: whole molecules are binned by their centre of mass, so two close sites of different molecules
      // can lie more than one cell apart and the adjacent-cell TOO_CLOSE scan of the unvisited pairs would miss them
      if (cfg->molpbc) e->pair_split = 0;
      // the TOO_CLOSE scan of the unvisited pairs looks at adjacent cells only: cells must be >= 0.5 A thick
      const int ng[3] = {e->T.nx, e->T.ny, e->T.nz};
      for (int d = 0; d < 3; d++) {
         const double *r = &e->T.hinv[3 * d];
         if (1.0 / sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) / ng[d] < 0.5) e->pair_split = 0;
      }
      for (int k = 0; k < 2; k++) e->sub[k].valid = false;
      if (e->pair_split) {
         if (!e->cls_uploaded || !e->d_cls) {
            if (upload(e, &e->d_cls, cls.data(), n)) return -1;
            e->cls_uploaded = true;
         }
         const int ncols = e->T.nx * e->T.ny;
         if (n + 1 > e->sub_cap || new_system || n / 2048 + ncols / 2048 + 4 > e->sub_scan_cap) {
            FREE(e->d_sub_flag); FREE(e->d_sub_pos); FREE(e->d_sub_scan);
            MDB_CUDA(cudaMalloc(&e->d_sub_flag, sizeof(int) * 2 * (size_t)(n + 1)));
            MDB_CUDA(cudaMalloc(&e->d_sub_pos, sizeof(int) * 2 * (size_t)(n + 1)));
            MDB_CUDA(cudaMalloc(&e->d_sub_scan, sizeof(int) * 2 * (size_t)(n / 2048 + ncols / 2048 + 4)));
            e->sub_cap = n + 1; e->sub_scan_cap = n / 2048 + ncols / 2048 + 4;
         }
         if (2 * (ncols + 1) > e->sub_cols_cap) {
            FREE(e->d_sub_cols);
            MDB_CUDA(cudaMalloc(&e->d_sub_cols, sizeof(int) * 4 * (size_t)(ncols + 1)));
            e->sub_cols_cap = 2 * (ncols + 1);
         }
         const long cnt[2] = {nc, np2};
         for (int k = 0; k < 2; k++) {
            // (re)allocate only when the class or the grid grew: constant-stress runs reconfigure every step
            SubList &S = e->sub[k];
            const int need_b = (int)cnt[k] / MDB_NI + ncols + 8;
            if ((int)cnt[k] > S.cap_n || e->ncells + 1 > S.cap_cells || need_b > S.batch_cap) {
               FREE(S.posq); FREE(S.sinfo); FREE(S.order); FREE(S.start); FREE(S.batches); FREE(S.nbatch); FREE(S.fs);
               const size_t m = (size_t)std::max((int)cnt[k], 1);
               MDB_CUDA(cudaMalloc(&S.posq, sizeof(double4) * m));
               MDB_CUDA(cudaMalloc(&S.sinfo, sizeof(int2) * m));
               MDB_CUDA(cudaMalloc(&S.order, sizeof(int) * m));
               MDB_CUDA(cudaMalloc(&S.start, sizeof(int) * (size_t)(e->ncells + 1)));
               MDB_CUDA(cudaMalloc(&S.batches, sizeof(int2) * (size_t)need_b));
               MDB_CUDA(cudaMalloc(&S.nbatch, sizeof(int)));
               MDB_CUDA(cudaMalloc(&S.fs, sizeof(double) * 3 * m));
               S.cap_n = (int)cnt[k]; S.cap_cells = e->ncells + 1; S.batch_cap = need_b;
            }
            S.n = (int)cnt[k];
         }
      }
   }

   // ---- reciprocal space ----
   if (cfg->do_recip) {
      if (!mdb_build_recip_tables(e->cfg, e->T, err)) { mdb_set_error(err); return -1; }
      HostTables &T = e->T;
      for (size_t v = 0; v < T.hk_valid.size(); v++) T.hk[T.hk_valid[v]].pad = (int)v;
      for (auto &d : T.hk) if (d.nl == 0) d.pad = -1;
      std::vector<int> slotinfo(2 * (size_t)T.nslots);
      for (int s = 0; s < T.nslots; s++) slotinfo[s] = T.slot_flags[s];
      for (size_t i = 0; i < T.hk.size(); i++)
         for (int l = 0; l < T.hk[i].nl; l++) slotinfo[T.nslots + T.hk[i].slot0 + l] = (int)i;
      if (!same_def || !e->d_cidx) {
         std::vector<int> cidx;
         for (int i = 0; i < n; i++) {
            if (i == cfg->nsites_xf) e->n_charged_nf = (int)cidx.size();
            if (e->h_chg[i] != 0.0) cidx.push_back(i);
         }
         if (cfg->nsites_xf >= n) e->n_charged_nf = (int)cidx.size();
         e->n_charged = (int)cidx.size();
         if (upload(e, &e->d_cidx, cidx.data(), cidx.size())) return -1;
         e->h_cidx = cidx;
      }
      if (upload(e, &e->d_hk, T.hk.data(), T.hk.size())) return -1;
      if (upload(e, &e->d_hk_valid, T.hk_valid.data(), T.hk_valid.size())) return -1;
      if (upload(e, &e->d_slot_flags, slotinfo.data(), slotinfo.size())) return -1;
      e->sfac_rank = -1; e->kf_rank = -1;
      if (T.nslots > e->slots_cap) {                       // grown (never shrunk: a breathing cell changes nslots a little)
         FREE(e->d_coef_tot); FREE(e->d_coef_nf); FREE(e->d_kpartials); FREE(e->d_psum);
         const size_t cap = (size_t)std::max(T.nslots, 1) * 5 / 4 + 64;
         MDB_CUDA(cudaMalloc(&e->d_coef_tot, sizeof(double) * 8 * cap));
         MDB_CUDA(cudaMalloc(&e->d_coef_nf, sizeof(double) * 8 * cap));
         MDB_CUDA(cudaMalloc(&e->d_kpartials, sizeof(double) * 8 * (cap / 256 + 1)));
         MDB_CUDA(cudaMalloc(&e->d_psum, sizeof(double) * 8 * cap));
         e->slots_cap = (int)cap;
      }
   } else {
      e->T.nhkl = 0; e->T.hk.clear(); e->T.hk_valid.clear(); e->T.nslots = 0;
   }
   e->configured = true;
   return 0;
}

extern "C" int mdb_set_sites_host(mdb_engine *e, const double *x, const double *y, const double *z, void *stream)
{
   cudaStream_t st = (cudaStream_t)stream;
   const size_t n = e->cfg.nsites;
   MDB_CUDA(cudaMemcpyAsync(e->own_xyz, x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
   MDB_CUDA(cudaMemcpyAsync(e->own_xyz + n, y, sizeof(double) * n, cudaMemcpyHostToDevice, st));
   MDB_CUDA(cudaMemcpyAsync(e->own_xyz + 2 * n, z, sizeof(double) * n, cudaMemcpyHostToDevice, st));
   e->d_x = e->own_xyz; e->d_y = e->own_xyz + n; e->d_z = e->own_xyz + 2 * n;
   e->sites_set = true; e->cells_valid = false;
   return 0;
}

extern "C" int mdb_set_com_host(mdb_engine *e, const double *c_of_m, void *stream)
{
   const int nm = e->cfg.nmols;
   if (nm <= 0) { mdb_set_error("mdb_set_com_host: nmols not configured"); return -1; }
   if (nm > e->com_cap) {
      FREE(e->d_com);
      MDB_CUDA(cudaMalloc(&e->d_com, sizeof(double) * 3 * (size_t)nm));
      e->com_cap = nm;
   }
   MDB_CUDA(cudaMemcpyAsync(e->d_com, c_of_m, sizeof(double) * 3 * (size_t)nm, cudaMemcpyHostToDevice, (cudaStream_t)stream));
   e->com_set = true; e->cells_valid = false;
   return 0;
}

extern "C" int mdb_set_sites_device(mdb_engine *e, const double *dx, const double *dy, const double *dz, void *stream)
{
   (void)stream;
   e->d_x = const_cast<double *>(dx); e->d_y = const_cast<double *>(dy); e->d_z = const_cast<double *>(dz);
   e->sites_set = true; e->cells_valid = false;
   return 0;
}

// Are the engine's current sites bit-identical to three HOST rows?  The rows are copied H2D into `d_scratch` (3 N
// doubles of device memory) on `stream` and compared there; returns the number of differing values (synchronises
// `stream`), or -1.  moldy_abi.cu uses it to validate the k-space sums it started ahead of ewald().
__global__ void __launch_bounds__(256) k_sites_differ(const unsigned long long *__restrict__ a, const unsigned long long *__restrict__ b,
                                                      size_t n, unsigned int *__restrict__ ndiff)
{
   unsigned int d = 0;
   for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) d += a[i] != b[i];
   if (__syncthreads_or(d != 0) && d) atomicAdd(ndiff, d);
}
extern "C" long mdb_sites_differ_host(mdb_engine *e, const double *x, const double *y, const double *z, double *d_scratch,
                                      void *stream)
{
   cudaStream_t st = (cudaStream_t)stream;
   const size_t n = e->cfg.nsites;
   if (!e->sites_set || e->d_x != e->own_xyz) return -1;
   MDB_CUDA(cudaMemcpyAsync(d_scratch, x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
   MDB_CUDA(cudaMemcpyAsync(d_scratch + n, y, sizeof(double) * n, cudaMemcpyHostToDevice, st));
   MDB_CUDA(cudaMemcpyAsync(d_scratch + 2 * n, z, sizeof(double) * n, cudaMemcpyHostToDevice, st));
   unsigned int *d_nd = reinterpret_cast<unsigned int *>(e->d_counters + 7);
   MDB_CUDA(cudaMemsetAsync(d_nd, 0, sizeof(unsigned int), st));
   k_sites_differ<<<592, 256, 0, st>>>(reinterpret_cast<const unsigned long long *>(d_scratch),
                                       reinterpret_cast<const unsigned long long *>(e->own_xyz), 3 * n, d_nd);
   e->launches++;
   unsigned int nd = 0;
   MDB_CUDA(cudaMemcpyAsync(&nd, d_nd, sizeof nd, cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaStreamSynchronize(st));
   return (long)nd;
}

extern "C" int mdb_zero_out(mdb_engine *e, double *d_out, void *stream)
{
   MDB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * mdb_out_doubles(e->cfg.nsites), (cudaStream_t)stream));
   return 0;
}

extern "C" int mdb_build_cells(mdb_engine *e, void *stream)
{
   if (!e->configured || !e->sites_set) { mdb_set_error("mdb_build_cells: engine not configured / no sites"); return -1; }
   return mdb_launch_cells(e, (cudaStream_t)stream);
}

extern "C" int mdb_force_real(mdb_engine *e, double *d_out, void *stream)
{
   if (!e->configured || !e->sites_set) { mdb_set_error("mdb_force_real: engine not configured / no sites"); return -1; }
   if (!e->cells_valid && mdb_launch_cells(e, (cudaStream_t)stream)) return -1;
   if (e->pair_mode >= 3) return mdb_launch_pair_tiled(e, d_out, (cudaStream_t)stream);
   e->pair_evals++;
   return mdb_launch_pair(e, d_out, (cudaStream_t)stream);
}

// k-space forces in slices (see mdb_engine::kf_chunks): set before mdb_force_recip; afterwards mdb_kforce_slices returns the
// number of slices, their events (recorded on the launching stream) and the exclusive upper bounds of the ORIGINAL site
// indices whose k-space forces are complete at each event (the last one is nsites).  0 slices: the launch was not cut.
extern "C" void mdb_set_kforce_slices(mdb_engine *e, int n) { e->kf_chunks = std::max(1, std::min(n, (int)mdb_engine::KF_MAXCH)); }
extern "C" int mdb_kforce_slices(const mdb_engine *e, void **events, int *site_hi)
{
   for (int k = 0; k < e->kf_nch; k++) { events[k] = (void *)e->kf_ev[k]; site_hi[k] = e->kf_site_hi[k]; }
   return e->kf_nch;
}

extern "C" int mdb_force_recip(mdb_engine *e, double *d_out, void *stream)
{
   if (!e->configured || !e->sites_set) { mdb_set_error("mdb_force_recip: engine not configured / no sites"); return -1; }
   e->kf_nch = 0;
   if (!e->cfg.do_recip) return 0;
   return mdb_launch_recip(e, d_out, (cudaStream_t)stream);
}

// ---- real space beside k-space -------------------------------------------------------------------------------------
// DFMA (pair kernel) and DMMA (k-space GEMMs) share one FP64 execution resource per SM sub-partition
// (profiles/r02_ubench_mix.txt), so running the phases side by side cannot beat the sum of their pipe times -- but each
// phase alone leaves the pipe idle a quarter to a third of the time (in-order warps waiting on their own latencies).
// A few pair warps resident beside the GEMM blocks take those slots.
__global__ void __launch_bounds__(256) k_add_block(double *__restrict__ out, const double *__restrict__ add, size_t n)
{
   for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] += add[i];
}

extern "C" int mdb_set_overlap(mdb_engine *e, int fill_blocks, int fill_threads)
{
   if (fill_blocks <= 0) { e->ovl_blocks = fill_blocks < 0 ? -1 : 0; e->ovl_threads = 0; return 0; }
   e->ovl_blocks = fill_blocks;
   e->ovl_threads = fill_threads > 0 ? std::min(128, (fill_threads + 31) / 32 * 32) : 64;
   return 0;
}

// batches the filler grid drew in the last mdb_force_both (synchronises the device); -1: no overlapped step yet
extern "C" long mdb_overlap_filled(mdb_engine *e)
{
   if (!e->d_ovl_q) return -1;
   int q[2] = {0, 0};
   if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(q, e->d_ovl_q, sizeof q, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
   return q[0];
}

extern "C" int mdb_force_both(mdb_engine *e, double *d_out, void *stream)
{
   cudaStream_t st = (cudaStream_t)stream;
   if (!e->configured || !e->sites_set) { mdb_set_error("mdb_force_both: engine not configured / no sites"); return -1; }
   if (e->ovl_blocks == -2) {
      const char *s = getenv("MDB_OVERLAP");             // "-1": off, "0": k-space first (default), "blocks,threads": filler grid
      int b = 0, t = 0;
      if (s) sscanf(s, "%d,%d", &b, &t);
      mdb_set_overlap(e, b, t);
   }
   if (e->ovl_blocks < 0 || !e->cfg.do_recip) {
      if (mdb_force_real(e, d_out, stream)) return -1;
      return mdb_force_recip(e, d_out, stream);
   }
   const bool filler = e->ovl_blocks > 0 && e->pair_mode == 4;                  // (the filler is a Newton-3 instantiation)
   const size_t nd = mdb_out_doubles(e->cfg.nsites);
   if (!e->ovl_stream) {
      int lo = 0, hi = 0;
      MDB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      MDB_CUDA(cudaStreamCreateWithPriority(&e->ovl_stream, cudaStreamNonBlocking, hi));
      MDB_CUDA(cudaEventCreateWithFlags(&e->ev_ovl_fork, cudaEventDisableTiming));
      MDB_CUDA(cudaEventCreateWithFlags(&e->ev_ovl_join, cudaEventDisableTiming));
      MDB_CUDA(cudaMalloc(&e->d_ovl_q, 2 * sizeof(int)));
   }
   if (nd > e->out2_cap) {
      FREE(e->d_out2);
      MDB_CUDA(cudaMalloc(&e->d_out2, sizeof(double) * nd));
      e->out2_cap = nd;
   }
   MDB_CUDA(cudaMemsetAsync(e->d_ovl_q, 0, 2 * sizeof(int), st));
   MDB_CUDA(cudaEventRecord(e->ev_ovl_fork, st));                    // the sites are complete at this point of `st`
   MDB_CUDA(cudaStreamWaitEvent(e->ovl_stream, e->ev_ovl_fork, 0));
   MDB_CUDA(cudaMemsetAsync(e->d_out2, 0, sizeof(double) * nd, e->ovl_stream));
   if (mdb_launch_recip(e, e->d_out2, e->ovl_stream)) return -1;
   MDB_CUDA(cudaMemsetAsync(e->d_ovl_q + 1, 1, sizeof(int), e->ovl_stream));      // stop flag: the filler grid ends
   MDB_CUDA(cudaEventRecord(e->ev_ovl_join, e->ovl_stream));
   // filler: the pair kernel shares the SMs with the k-space kernels; otherwise the k-space chain goes first, with the cell
   // build and the sub-list compaction (0.3 ms of launch latencies at 10^6 sites) hidden behind it, and the pair passes follow
   e->ovl_armed = filler;
   e->pre_pair_wait = filler ? nullptr : e->ev_ovl_join;
   const int rc = mdb_force_real(e, d_out, stream);
   e->ovl_armed = false;
   e->pre_pair_wait = nullptr;
   if (rc) return -1;
   MDB_CUDA(cudaStreamWaitEvent(st, e->ev_ovl_join, 0));
   k_add_block<<<592, 256, 0, st>>>(d_out, e->d_out2, nd);
   e->launches++;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

// k-space with the SITE partition (moldy_b200/spmd.py): pass 1 leaves this rank's structure-factor
// sums in d_psum (mdb_recip_sum_doubles() doubles, device memory of the caller); the caller
// all-reduces them over the ranks; pass 2 turns them into energy/stress (rank 0 only) and the
// forces on this rank's sites.
extern "C" size_t mdb_recip_sum_doubles(const mdb_engine *e) { return 8 * (size_t)std::max(e->T.nslots, 1); }
extern "C" int mdb_recip_partial(mdb_engine *e, double *d_psum, void *stream)
{
   if (!e->configured || !e->sites_set) { mdb_set_error("mdb_recip_partial: engine not configured / no sites"); return -1; }
   if (!e->cfg.do_recip) return 0;
   return mdb_launch_recip_partial(e, d_psum, (cudaStream_t)stream);
}
extern "C" int mdb_recip_finish(mdb_engine *e, const double *d_psum, double *d_out, void *stream)
{
   if (!e->configured || !e->sites_set) { mdb_set_error("mdb_recip_finish: engine not configured / no sites"); return -1; }
   if (!e->cfg.do_recip) return 0;
   return mdb_launch_recip_finish(e, d_psum, d_out, (cudaStream_t)stream);
}

extern "C" int mdb_read_out(mdb_engine *e, const double *d_out, double *h_out, void *stream)
{
   cudaStream_t st = (cudaStream_t)stream;
   MDB_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(double) * mdb_out_doubles(e->cfg.nsites), cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaStreamSynchronize(st));
   return 0;
}

extern "C" int mdb_grid(const mdb_engine *e, int nxyz[3])
{
   nxyz[0] = e->T.nx; nxyz[1] = e->T.ny; nxyz[2] = e->T.nz;
   return e->ncells;
}
extern "C" int mdb_n_neighbour_cells(const mdb_engine *e) { return (int)e->T.half_list.size() / 3; }
extern "C" int mdb_n_kvectors(const mdb_engine *e) { return e->T.nhkl; }
extern "C" int mdb_pair_split(const mdb_engine *e) { return e->pair_split; }
extern "C" long mdb_kernel_launches(const mdb_engine *e) { return e->launches; }

extern "C" int mdb_get_cell_ids(mdb_engine *e, int *h_cell, void *stream)
{
   cudaStream_t st = (cudaStream_t)stream;
   if (!e->cells_valid && mdb_build_cells(e, stream)) return -1;
   MDB_CUDA(cudaMemcpyAsync(h_cell, e->d_cell, sizeof(int) * (size_t)e->cfg.nsites, cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaStreamSynchronize(st));
   return 0;
}

static int read_counters(mdb_engine *e, unsigned long long c[8], bool reset, cudaStream_t st)
{
   MDB_CUDA(cudaMemcpyAsync(c, e->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaStreamSynchronize(st));
   if (reset) MDB_CUDA(cudaMemsetAsync(e->d_counters, 0, 8 * sizeof(unsigned long long), st));
   return 0;
}

// pairs handed to kernel() per force evaluation (src/force.c:960) for the current configuration.
// Tiled modes: the force kernel's traversal and window tests re-run without the arithmetic;
// mode 2: the visits counted by the force kernel itself, averaged over the evaluations since the last call.
extern "C" double mdb_pair_count(mdb_engine *e, void *stream)
{
   cudaStream_t st = (cudaStream_t)stream;
   unsigned long long c[8], z = 0;
   double evals = 1.0;
   if (e->pair_mode >= 3) {
      if (!e->configured || !e->sites_set) { mdb_set_error("mdb_pair_count: engine not configured / no sites"); return -1.0; }
      if (!e->cells_valid && mdb_launch_cells(e, st)) return -1.0;
      cudaMemcpyAsync(e->d_counters, &z, sizeof z, cudaMemcpyHostToDevice, st);
      if (mdb_launch_pair_count_tiled(e, st)) return -1.0;
   } else {
      evals = e->pair_evals > 0 ? (double)e->pair_evals : 1.0;
      e->pair_evals = 0;
   }
   if (read_counters(e, c, false, st)) return -1.0;
   cudaMemcpyAsync(e->d_counters, &z, sizeof z, cudaMemcpyHostToDevice, st);
   cudaStreamSynchronize(st);
   return 0.5 * (double)c[0] / evals;
}

// RDF binning pass of force_calc (src/force.c:1302-1313, rdf_inner + rdf_accum): h_counts[pair][bin] +=
// number of site pairs of this rank's share whose distance falls into the bin, pair = (idi <= idj)
// in the order of init_rdf (src/rdf.c:82-90), nbins bins of width limit/nbins.
extern "C" size_t mdb_rdf_size(const mdb_engine *e, int nbins)
{
   return (size_t)nbins * (e->cfg.max_id * (e->cfg.max_id - 1) / 2);
}
extern "C" int mdb_rdf_counts(mdb_engine *e, double limit, int nbins, unsigned long long *h_counts, void *stream)
{
   cudaStream_t st = (cudaStream_t)stream;
   if (!e->configured || !e->sites_set) { mdb_set_error("mdb_rdf_counts: engine not configured / no sites"); return -1; }
   if (limit <= 0.0 || nbins <= 0) { mdb_set_error("mdb_rdf_counts: bad limit / nbins"); return -1; }
   for (int i = 0; i < e->cfg.nsites; i++)
      if (e->h_type[i] < 1) { mdb_set_error("mdb_rdf_counts: site id 0 has no RDF slot (src/rdf.c:78-90)"); return -1; }
   if (!e->cells_valid && mdb_launch_cells(e, st)) return -1;
   if (e->pair_mode < 3 && mdb_launch_batches(e, st)) return -1;      // the pass walks the tiled kernel's batches
   const bool same = e->d_runs_rdf && e->rdf_limit == limit && e->rdf_grid[0] == e->T.nx && e->rdf_grid[1] == e->T.ny &&
                     e->rdf_grid[2] == e->T.nz && !memcmp(e->rdf_h, e->cfg.h, sizeof e->rdf_h);
   if (!same) {
      std::vector<StencilRun> runs;
      std::string err;
      if (!mdb_build_rdf_runs(e->cfg, e->T, limit, runs, err)) { mdb_set_error(err); return -1; }
      if (upload(e, &e->d_runs_rdf, runs.data(), runs.size())) return -1;
      e->nruns_rdf = (int)runs.size(); e->rdf_limit = limit;
      e->rdf_grid[0] = e->T.nx; e->rdf_grid[1] = e->T.ny; e->rdf_grid[2] = e->T.nz;
      memcpy(e->rdf_h, e->cfg.h, sizeof e->rdf_h);
   }
   const size_t nh = mdb_rdf_size(e, nbins);
   if (nh == 0) return 0;
   if (nh > e->rdf_cap) {
      FREE(e->d_rdf);
      MDB_CUDA(cudaMalloc(&e->d_rdf, sizeof(unsigned long long) * nh));
      e->rdf_cap = nh;
   }
   MDB_CUDA(cudaMemsetAsync(e->d_rdf, 0, sizeof(unsigned long long) * nh, st));
   if (mdb_launch_rdf_tiled(e, e->d_runs_rdf, e->nruns_rdf, (double)nbins / limit, nbins, e->d_rdf, st)) return -1;
   std::vector<unsigned long long> tmp(nh);
   MDB_CUDA(cudaMemcpyAsync(tmp.data(), e->d_rdf, sizeof(unsigned long long) * nh, cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaStreamSynchronize(st));
   for (size_t k = 0; k < nh; k++) h_counts[k] += tmp[k];
   return 0;
}

extern "C" int mdb_too_close(mdb_engine *e, int pair[2], void *stream)
{
   unsigned long long c[8];
   if (read_counters(e, c, false, (cudaStream_t)stream)) return -1;
   pair[0] = (int)(c[3] >> 32); pair[1] = (int)(c[3] & 0xffffffffu);
   unsigned long long z[3] = {0, 0, 0};
   cudaMemcpyAsync(e->d_counters + 1, z, sizeof z, cudaMemcpyHostToDevice, (cudaStream_t)stream);
   cudaStreamSynchronize((cudaStream_t)stream);
   // every inter-molecular close pair is visited from both ends; [2] = bin errors
   return (int)(c[1] / 2) + (c[2] ? (1 << 30) : 0);
}

extern "C" size_t mdb_sizeof(const char *name)
{
   if (!strcmp(name, "contr_mt")) return sizeof(contr_mt);
   if (!strcmp(name, "system_mt")) return sizeof(system_mt);
   if (!strcmp(name, "spec_mt")) return sizeof(spec_mt);
   if (!strcmp(name, "site_mt")) return sizeof(site_mt);
   if (!strcmp(name, "pot_mt")) return sizeof(pot_mt);
   if (!strcmp(name, "mdb_config")) return sizeof(mdb_config);
   if (!strcmp(name, "HkDesc")) return sizeof(HkDesc);
   return 0;
}

// ---- FP64 roofline probe: a pure dependent-chain-free DFMA loop ----------------
__global__ void __launch_bounds__(256) k_dfma_probe(double *out, int iters, double a, double b)
{
   double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
   for (int i = 0; i < iters; i++) {
      v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
      v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
}

extern "C" double mdb_fp64_peak_probe(int device, int iters)
{
   if (cudaSetDevice(device) != cudaSuccess) return -1.0;
   cudaDeviceProp prop;
   cudaGetDeviceProperties(&prop, device);
   const int blocks = prop.multiProcessorCount * 8, threads = 256;
   double *d = nullptr;
   if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   k_dfma_probe<<<blocks, threads>>>(d, iters / 8, 0.999999, 1e-9);
   double best = 0.0;
   for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      k_dfma_probe<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      double flops = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3);
      best = std::max(best, flops);
   }
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   cudaFree(d);
   return best;
}

// ---- FP64 tensor-pipe probe: mma.sync.m8n8k4.f64 (DMMA.8x8x4) with 8 independent accumulator tiles per warp ----
__global__ void __launch_bounds__(256) k_dmma_probe(double *out, int iters)
{
   double a = 1e-3 * (threadIdx.x % 7 + 1), b = 1e-3 * (threadIdx.x % 5 + 1), c[8][2];
   for (int k = 0; k < 8; k++) { c[k][0] = k; c[k][1] = k + 1; }
   for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < 8; k++)
         asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                      : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
   }
   double s = 0;
   for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" double mdb_dmma_peak_probe(int device, int iters)
{
   if (cudaSetDevice(device) != cudaSuccess) return -1.0;
   cudaDeviceProp prop;
   cudaGetDeviceProperties(&prop, device);
   const int blocks = prop.multiProcessorCount * 2, threads = 256;
   double *d = nullptr;
   if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   k_dmma_probe<<<blocks, threads>>>(d, iters / 8);
   double best = 0.0;
   for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      k_dmma_probe<<<blocks, threads>>>(d, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      // 8 DMMA per iteration per warp, 2*8*8*4 flop each
      best = std::max(best, 512.0 * 8.0 * (double)iters * blocks * (threads / 32) / (ms * 1e-3));
   }
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   cudaFree(d);
   return best;
}
