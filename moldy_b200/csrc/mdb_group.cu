// mdb_group.cu -- several GPUs behind ONE process: the form in which the unmodified Moldy program (built without -DSPMD,
// nthreads = 1; SURVEY 8b "Threading") uses more than one device.  A group = P engines + P peers (mdb_peer.cu), driven by
// the calling thread phase by phase: every phase is enqueued on all ranks before the next one, so no blocking call can wait
// for a barrier whose partner has not been enqueued.  Selected with MOLDY_B200_DEVICES=0,1,.. / 0-7 / all (moldy_abi.cu).
//
//   mdb_group_force_host       force_calc()/ewald() level: every rank uploads its slice of the caller's site rows over its
//                              own PCIe link, the slices are all-gathered over NVLink, the partial sums are reduce-scattered
//                              and every rank writes its slice of the summed forces into the caller's rows.
//   mdb_group_eval_forces_host eval_forces() level (src/accel.c:398-617): c-of-m/quaternion slices in; every rank builds all
//                              sites, sums its share of the pairs and of the charged sites, receives the complete forces of
//                              ITS molecules from the reduce-scatter, and returns their molecular forces and torques.
#include <string.h>
#include <algorithm>
#include <vector>
#include "mdb_internal.h"

struct mdb_group {
   int world = 0;
   std::vector<int> dev;
   std::vector<mdb_engine *> eng;
   std::vector<mdb_peer *> peer;
   std::vector<cudaStream_t> st;
   size_t peer_n = 0; int peer_nslots = 0;
   double *h_in = nullptr, *h_res = nullptr, *h_scal = nullptr;    // pinned (portable)
   size_t in_cap = 0, res_cap = 0;
   std::vector<int> mol_lo;                                         // molecule bounds of the ranks (eval_forces)
   bool species_set = false;
};

#define GFOR(r) for (int r = 0; r < g->world; r++)

extern "C" mdb_group *mdb_group_create(int ndev, const int *devices)
{
   if (ndev < 1 || ndev > MDB_MAX_PEERS) { mdb_set_error("mdb_group_create: 1..16 devices"); return nullptr; }
   mdb_group *g = new mdb_group();
   g->world = ndev;
   for (int r = 0; r < ndev; r++) {
      mdb_engine *e = mdb_create(devices[r]);
      if (!e) { delete g; return nullptr; }
      cudaStream_t s = nullptr;
      if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { mdb_set_error("mdb_group_create: stream"); delete g; return nullptr; }
      g->dev.push_back(devices[r]); g->eng.push_back(e); g->st.push_back(s);
   }
   if (cudaHostAlloc(&g->h_scal, sizeof(double) * (MDB_EVAL_SCALARS + MDB_OUT_SCALARS) * ndev, cudaHostAllocPortable) != cudaSuccess) {
      mdb_set_error("mdb_group_create: pinned memory");
      delete g;
      return nullptr;
   }
   return g;
}

static void drop_peers(mdb_group *g)
{
   for (auto *p : g->peer) mdb_peer_destroy(p);
   g->peer.clear();
}

extern "C" void mdb_group_destroy(mdb_group *g)
{
   if (!g) return;
   drop_peers(g);
   GFOR(r) { cudaSetDevice(g->dev[r]); cudaStreamDestroy(g->st[r]); mdb_destroy(g->eng[r]); }
   if (g->h_in) cudaFreeHost(g->h_in);
   if (g->h_res) cudaFreeHost(g->h_res);
   if (g->h_scal) cudaFreeHost(g->h_scal);
   delete g;
}

extern "C" int mdb_group_size(const mdb_group *g) { return g->world; }
extern "C" mdb_engine *mdb_group_engine(mdb_group *g, int r) { return g->eng[r]; }
extern "C" void *mdb_group_stream(mdb_group *g, int r) { return (void *)g->st[r]; }

extern "C" int mdb_group_configure(mdb_group *g, const mdb_config *cfg)
{
   GFOR(r) if (mdb_configure(g->eng[r], cfg)) return -1;
   const mdb_engine *e0 = g->eng[0];
   const bool fits = !g->peer.empty() && g->peer_n == (size_t)cfg->nsites && e0->T.nslots <= g->peer_nslots;
   if (!fits) {
      drop_peers(g);
      GFOR(r) {
         mdb_peer *p = mdb_peer_create(g->eng[r], r, g->world);
         if (!p) return -1;
         g->peer.push_back(p);
      }
      if (mdb_peer_connect(g->peer.data(), g->world)) return -1;
      g->peer_n = (size_t)cfg->nsites;
      g->peer_nslots = std::max(e0->T.nslots, 1) * 3 / 2;
   } else {
      GFOR(r) mdb_set_partition(g->eng[r], r, g->world);
   }
   return 0;
}

static int barrier_all(mdb_group *g) { GFOR(r) if (mdb_peer_barrier(g->peer[r], g->st[r])) return -1; return 0; }

// Forces of the caller's HOST site rows, summed over the ranks, written (=, not +=) into three HOST rows, and the 16
// scalars [pe_real, pe_recip, stress[9], ..].  what: bit 0 real, bit 1 reciprocal space.  c_of_m: scaled centres of mass
// (molecular cut-off mode) or NULL.  tc: TOO_CLOSE count and example pair as mdb_too_close.  Blocks until done.
extern "C" int mdb_group_force_host(mdb_group *g, const double *x, const double *y, const double *z, const double *c_of_m,
                                    int what, double *fx, double *fy, double *fz, double *scal16, int *tc, int tc_pair[2])
{
   if (g->peer.empty()) { mdb_set_error("mdb_group_force_host: group not configured"); return -1; }
   GFOR(r) if (mdb_peer_sites_host_slice(g->peer[r], x, y, z, g->st[r])) return -1;
   if (c_of_m) GFOR(r) { cudaSetDevice(g->dev[r]); if (mdb_set_com_host(g->eng[r], c_of_m, g->st[r])) return -1; }
   if (barrier_all(g)) return -1;
   GFOR(r) if (mdb_peer_sites_gather(g->peer[r], g->st[r])) return -1;
   GFOR(r) if (mdb_peer_phase_a(g->peer[r], what, g->st[r])) return -1;
   if (barrier_all(g)) return -1;
   GFOR(r) if (mdb_peer_phase_b(g->peer[r], what, g->st[r])) return -1;
   if (barrier_all(g)) return -1;
   GFOR(r) if (mdb_peer_phase_c(g->peer[r], g->st[r])) return -1;
   // D2H of every rank's slice into the caller's rows; the copies of all ranks are in flight before the first wait
   GFOR(r) {
      long long lohi[2];
      mdb_peer_slice(g->peer[r], lohi);
      const double *red = mdb_peer_result(g->peer[r]);
      const size_t n = g->peer_n;
      double *rows[3] = {fx, fy, fz};
      MDB_CUDA(cudaSetDevice(g->dev[r]));
      for (int a = 0; a < 3 && lohi[1] > lohi[0]; a++)
         MDB_CUDA(cudaMemcpyAsync(rows[a] + lohi[0], red + a * n + lohi[0], sizeof(double) * (size_t)(lohi[1] - lohi[0]),
                                  cudaMemcpyDeviceToHost, g->st[r]));
      if (r == 0) MDB_CUDA(cudaMemcpyAsync(g->h_scal, red + 3 * n, sizeof(double) * MDB_OUT_SCALARS, cudaMemcpyDeviceToHost, g->st[r]));
   }
   GFOR(r) { MDB_CUDA(cudaSetDevice(g->dev[r])); MDB_CUDA(cudaStreamSynchronize(g->st[r])); }
   GFOR(r) if (mdb_peer_error(g->peer[r], g->st[r]) != 0) { mdb_set_error("mdb_group: a peer barrier timed out"); return -1; }
   memcpy(scal16, g->h_scal, sizeof(double) * MDB_OUT_SCALARS);
   if (tc) {
      *tc = 0;
      GFOR(r) {
         int pr[2];
         MDB_CUDA(cudaSetDevice(g->dev[r]));
         const int t = mdb_too_close(g->eng[r], pr, g->st[r]);
         if (t < 0) return -1;
         if ((t & ~(1 << 30)) && tc_pair) { tc_pair[0] = pr[0]; tc_pair[1] = pr[1]; }
         *tc = ((*tc & ~(1 << 30)) + (t & ~(1 << 30))) | ((*tc | t) & (1 << 30));      // counts add up, bit 30 = bin error
      }
   }
   return 0;
}

// ---- eval_forces() ---------------------------------------------------------------------------------------------------
extern "C" int mdb_group_set_species(mdb_group *g, int nspecies, const mdb_species *sp, const double *pfs)
{
   GFOR(r) { MDB_CUDA(cudaSetDevice(g->dev[r])); if (mdb_set_species(g->eng[r], nspecies, sp, pfs)) return -1; }
   const auto &M = g->eng[0]->mf;
   // molecule shares: equal numbers of SITES per rank, cut at molecule boundaries
   std::vector<long long> site_of_mol(1, 0);
   for (int i = 0; i < nspecies; i++)
      for (int m = 0; m < sp[i].nmols; m++) site_of_mol.push_back(site_of_mol.back() + sp[i].nsites);
   const long long n = site_of_mol.back();
   g->mol_lo.assign(g->world + 1, 0);
   long long bounds[MDB_MAX_PEERS + 1];
   for (int r = 0; r <= g->world; r++) {
      const long long target = n * r / g->world;
      const int m = (int)(std::lower_bound(site_of_mol.begin(), site_of_mol.end(), target) - site_of_mol.begin());
      g->mol_lo[r] = std::min(m, M.nmols);
      bounds[r] = site_of_mol[g->mol_lo[r]];
   }
   g->mol_lo[g->world] = M.nmols; bounds[g->world] = n;
   GFOR(r) if (mdb_peer_set_site_bounds(g->peer[r], bounds)) return -1;
   const size_t need_in = 3 * (size_t)M.nmols + 4 * (size_t)M.nmols_q, need_res = mdb_eval_result_doubles(g->eng[0]);
   if (need_in > g->in_cap) {
      if (g->h_in) cudaFreeHost(g->h_in);
      MDB_CUDA(cudaHostAlloc(&g->h_in, sizeof(double) * need_in, cudaHostAllocPortable));
      g->in_cap = need_in;
   }
   if (need_res > g->res_cap) {
      if (g->h_res) cudaFreeHost(g->h_res);
      MDB_CUDA(cudaHostAlloc(&g->h_res, sizeof(double) * need_res, cudaHostAllocPortable));
      g->res_cap = need_res;
   }
   g->species_set = true;
   return 0;
}

extern "C" size_t mdb_group_eval_result_doubles(const mdb_group *g) { return mdb_eval_result_doubles(g->eng[0]); }
extern "C" const double *mdb_group_eval_result(const mdb_group *g) { return g->h_res; }

// Same contract and result layout as mdb_eval_forces_host.  rdf_*: when rdf_counts != NULL the RDF pass runs on every
// rank's share of the batches and the counts are added up.  tc / tc_pair as above.
extern "C" int mdb_group_eval_forces_host(mdb_group *g, const double h[9], const double *const *com, const double *const *quat,
                                          int surface_dipole, int do_recip, double *h_result, double rdf_limit, int rdf_nbins,
                                          unsigned long long *rdf_counts, int *tc, int tc_pair[2])
{
   if (g->peer.empty() || !g->species_set) { mdb_set_error("mdb_group_eval_forces_host: group not configured"); return -1; }
   mdb_engine *e0 = g->eng[0];
   const auto &M0 = e0->mf;
   const size_t len_in = 3 * (size_t)M0.nmols + 4 * (size_t)M0.nmols_q, n = g->peer_n;
   const int what = 1 | (do_recip ? 2 : 0);
   if (mdb_evalf_stage_inputs(e0, com, quat, g->h_in)) return -1;
   GFOR(r) if (mdb_peer_in_host_slice(g->peer[r], g->h_in, len_in, g->st[r])) return -1;
   if (barrier_all(g)) return -1;
   GFOR(r) if (mdb_peer_in_gather(g->peer[r], len_in, g->st[r])) return -1;
   GFOR(r) { MDB_CUDA(cudaSetDevice(g->dev[r])); if (mdb_evalf_pre(g->eng[r], h, mdb_peer_in(g->peer[r]), g->st[r])) return -1; }
   GFOR(r) if (mdb_peer_phase_a(g->peer[r], what, g->st[r])) return -1;
   if (barrier_all(g)) return -1;
   GFOR(r) if (mdb_peer_phase_b(g->peer[r], what, g->st[r])) return -1;
   if (barrier_all(g)) return -1;
   GFOR(r) if (mdb_peer_phase_c(g->peer[r], g->st[r])) return -1;
   // RDF pass of force_calc (src/force.c:1302-1313) on this step's cell lists, before the second make_sites replaces the
   // sites; every rank bins its share of the batches (blocks on each rank in turn: all barriers above are enqueued)
   if (rdf_counts) GFOR(r) {
      MDB_CUDA(cudaSetDevice(g->dev[r]));
      if (mdb_rdf_counts(g->eng[r], rdf_limit, rdf_nbins, rdf_counts, g->st[r])) return -1;
   }
   const size_t res_scal = 3 * (size_t)M0.nmols + 3 * (size_t)M0.nmols_r;
   GFOR(r) {
      mdb_engine *e = g->eng[r];
      auto &M = e->mf;
      MDB_CUDA(cudaSetDevice(g->dev[r]));
      const int m_lo = g->mol_lo[r], m_hi = g->mol_lo[r + 1];
      if (mdb_evalf_tail(e, h, mdb_peer_in(g->peer[r]), mdb_peer_result(g->peer[r]), m_lo, m_hi, surface_dipole, do_recip, g->st[r]))
         return -1;
      // this rank's molecules: forces, torques (per species the rows of [m_lo, m_hi)), and its scalars
      for (size_t i = 0; i < M.sp.size(); i++) {
         const int a = std::max(m_lo, M.mol_off[i]) - M.mol_off[i], b = std::min(m_hi, M.mol_off[i] + M.sp[i].nmols) - M.mol_off[i];
         if (b <= a) continue;
         const size_t fo = 3 * ((size_t)M.mol_off[i] + a);
         MDB_CUDA(cudaMemcpyAsync(g->h_res + fo, M.d_res + fo, sizeof(double) * 3 * (size_t)(b - a), cudaMemcpyDeviceToHost, g->st[r]));
         if (M.torq_off[i] >= 0) {
            const size_t to = 3 * (size_t)M.nmols + 3 * ((size_t)M.torq_off[i] + a);
            MDB_CUDA(cudaMemcpyAsync(g->h_res + to, M.d_res + to, sizeof(double) * 3 * (size_t)(b - a), cudaMemcpyDeviceToHost, g->st[r]));
         }
      }
      MDB_CUDA(cudaMemcpyAsync(g->h_scal + (size_t)r * MDB_EVAL_SCALARS, M.d_res + res_scal, sizeof(double) * MDB_EVAL_SCALARS,
                               cudaMemcpyDeviceToHost, g->st[r]));
   }
   GFOR(r) { MDB_CUDA(cudaSetDevice(g->dev[r])); MDB_CUDA(cudaStreamSynchronize(g->st[r])); }
   GFOR(r) if (mdb_peer_error(g->peer[r], g->st[r]) != 0) { mdb_set_error("mdb_group: a peer barrier timed out"); return -1; }
   // scalars: dipole moment, energies and stress are complete and identical on every rank; the virial partial sums add up
   double *sc = g->h_res + res_scal;
   memcpy(sc, g->h_scal, sizeof(double) * MDB_EVAL_SCALARS);
   for (int r = 1; r < g->world; r++)
      for (int k = 3; k < 12; k++) sc[k] += g->h_scal[(size_t)r * MDB_EVAL_SCALARS + k];
   if (tc) {
      *tc = 0;
      GFOR(r) {
         int pr[2];
         MDB_CUDA(cudaSetDevice(g->dev[r]));
         const int t = mdb_too_close(g->eng[r], pr, g->st[r]);
         if (t < 0) return -1;
         if ((t & ~(1 << 30)) && tc_pair) { tc_pair[0] = pr[0]; tc_pair[1] = pr[1]; }
         *tc = ((*tc & ~(1 << 30)) + (t & ~(1 << 30))) | ((*tc | t) & (1 << 30));      // counts add up, bit 30 = bin error
      }
   }
   if (h_result) memcpy(h_result, g->h_res, sizeof(double) * mdb_eval_result_doubles(e0));
   (void)n;
   return 0;
}
