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
#include <stdlib.h>
#include <thread>
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
   // resident NVE step: molecule bounds as element bounds of the c-of-m row and of the quaternion row, staging
   long long com_b[MDB_MAX_PEERS + 1] = {0}, quat_b[MDB_MAX_PEERS + 1] = {0};
   double *h_md = nullptr, *h_state = nullptr; size_t md_cap = 0, state_cap = 0;
   bool md_set = false;
   long md_steps = 0, ev_calls = 0;                                 // steps / evaluations since the set-up (the first one allocates)
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
   if (g->h_md) cudaFreeHost(g->h_md);
   if (g->h_state) cudaFreeHost(g->h_state);
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
   g->species_set = true; g->ev_calls = 0;
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
   // one host thread per rank for the launches up to the reduce-scatter, as in mdb_group_md_step (every rank on its own GPU,
   // buffers allocated by an earlier call, no RDF pass)
   bool distinct = g->world > 1;
   GFOR(r) for (int q = 0; q < r; q++) distinct = distinct && g->dev[q] != g->dev[r];
   static const bool threads_env = !(getenv("MDB_GROUP_THREADS") && atoi(getenv("MDB_GROUP_THREADS")) == 0);
   const bool threaded = distinct && threads_env && !rdf_counts && g->ev_calls > 0;
   g->ev_calls++;
   if (threaded) {
      std::vector<int> rc(g->world, 0);
      auto seq = [&](int r) {
         mdb_peer *p = g->peer[r];
         cudaStream_t st = g->st[r];
         rc[r] = cudaSetDevice(g->dev[r]) != cudaSuccess || mdb_peer_in_host_slice(p, g->h_in, len_in, st) || mdb_peer_barrier(p, st) ||
                 mdb_peer_in_gather(p, len_in, st) || mdb_evalf_pre(g->eng[r], h, mdb_peer_in(p), st) || mdb_peer_phase_a(p, what, st) ||
                 mdb_peer_barrier(p, st) || mdb_peer_phase_b(p, what, st) || mdb_peer_barrier(p, st) || mdb_peer_phase_c(p, st);
      };
      std::vector<std::thread> th;
      for (int r = 1; r < g->world; r++) th.emplace_back(seq, r);
      seq(0);
      for (auto &t : th) t.join();
      GFOR(r) if (rc[r]) { if (!*mdb_last_error()) mdb_set_error("mdb_group_eval_forces_host: a rank failed"); return -1; }
   } else {
   GFOR(r) if (mdb_peer_in_host_slice(g->peer[r], g->h_in, len_in, g->st[r])) return -1;
   if (barrier_all(g)) return -1;
   GFOR(r) if (mdb_peer_in_gather(g->peer[r], len_in, g->st[r])) return -1;
   GFOR(r) { MDB_CUDA(cudaSetDevice(g->dev[r])); if (mdb_evalf_pre(g->eng[r], h, mdb_peer_in(g->peer[r]), g->st[r])) return -1; }
   GFOR(r) if (mdb_peer_phase_a(g->peer[r], what, g->st[r])) return -1;
   if (barrier_all(g)) return -1;
   GFOR(r) if (mdb_peer_phase_b(g->peer[r], what, g->st[r])) return -1;
   if (barrier_all(g)) return -1;
   GFOR(r) if (mdb_peer_phase_c(g->peer[r], g->st[r])) return -1;
   }
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

// TOO_CLOSE count / example pair / bin-error bit over all ranks (as mdb_too_close)
extern "C" int mdb_group_too_close(mdb_group *g, int tc_pair[2])
{
   int tc = 0;
   GFOR(r) {
      int pr[2];
      if (cudaSetDevice(g->dev[r]) != cudaSuccess) return -1;
      const int t = mdb_too_close(g->eng[r], pr, g->st[r]);
      if (t < 0) return -1;
      if ((t & ~(1 << 30)) && tc_pair) { tc_pair[0] = pr[0]; tc_pair[1] = pr[1]; }
      tc = ((tc & ~(1 << 30)) + (t & ~(1 << 30))) | ((tc | t) & (1 << 30));
   }
   return tc;
}

// ---- the NVE step with the state resident on the GPUs of the group (mdb_md.cu on every rank's share of the molecules) ---
// Every rank keeps the whole [c-of-m | quaternions] block in its peer window and the momenta of ITS molecules; a step is
//   coords(step/2) on own molecules -> barrier, all-gather of the block -> make_sites, phases A | B | C, molecular forces and
//   torques of own molecules -> momenta(step/2) x 2 [sums at the half step] -> coords(step/2) -> sums;
// the per-rank sums (kinetic-energy dyads, mean squares, virial pieces) are added on the host.
extern "C" int mdb_group_md_set_dynamics(mdb_group *g, const mdb_species_dyn *dyn, int nosymmetric_rot)
{
   if (!g->species_set) { mdb_set_error("mdb_group_md_set_dynamics: mdb_group_set_species was not called"); return -1; }
   GFOR(r) { MDB_CUDA(cudaSetDevice(g->dev[r])); if (mdb_md_set_dynamics(g->eng[r], dyn, nosymmetric_rot)) return -1; }
   const auto &M = g->eng[0]->mf;
   for (int r = 0; r <= g->world; r++) {
      const int m = g->mol_lo[r];
      long long q = 0;                                    // molecules with quaternions below molecule m
      for (size_t i = 0; i < M.sp.size(); i++)
         if (M.quat_off[i] >= 0) q += std::max(0, std::min(m, M.mol_off[i] + M.sp[i].nmols) - M.mol_off[i]);
      g->com_b[r] = 3LL * m; g->quat_b[r] = 4LL * q;
   }
   const size_t ns = mdb_md_scalars(g->eng[0]) * (size_t)g->world;
   if (ns > g->md_cap) {
      if (g->h_md) cudaFreeHost(g->h_md);
      MDB_CUDA(cudaHostAlloc(&g->h_md, sizeof(double) * (ns + mdb_md_scalars(g->eng[0])), cudaHostAllocPortable));
      g->md_cap = ns;
   }
   g->md_set = true; g->md_steps = 0;
   return 0;
}

// pinned staging block of the group: [c-of-m 3 nmols | quaternions 4 nmols_q | mom 3 nmols | amom 4 nmols_q | force | torque]
static int group_state_block(mdb_group *g, size_t off[6])
{
   const auto &M0 = g->eng[0]->mf;
   const size_t nm = (size_t)M0.nmols, nq = (size_t)M0.nmols_q, nr = (size_t)M0.nmols_r;
   off[0] = 0; off[1] = 3 * nm; off[2] = off[1] + 4 * nq; off[3] = off[2] + 3 * nm; off[4] = off[3] + 4 * nq; off[5] = off[4] + 3 * nm;
   const size_t need = off[5] + 3 * nr + 8;
   if (need > g->state_cap) {
      if (g->h_state) cudaFreeHost(g->h_state);
      g->h_state = nullptr; g->state_cap = 0;
      MDB_CUDA(cudaHostAlloc(&g->h_state, sizeof(double) * need, cudaHostAllocPortable));
      g->state_cap = need;
   }
   return 0;
}

// The caller's per-species arrays are staged ONCE in pinned memory (six host threads); every rank then copies the block over
// its own PCIe link, all ranks at the same time.
extern "C" int mdb_group_md_upload_state(mdb_group *g, const double *const *com, const double *const *quat, const double *const *mom,
                                         const double *const *amom)
{
   if (!g->md_set) { mdb_set_error("mdb_group_md_upload_state: mdb_group_md_set_dynamics was not called"); return -1; }
   const auto &M0 = g->eng[0]->mf;
   const size_t nm = (size_t)M0.nmols, nq = (size_t)M0.nmols_q;
   size_t off[6];
   if (group_state_block(g, off)) return -1;
   double *hs = g->h_state;
   std::vector<MdbCopyJob> jobs;
   for (size_t i = 0; i < M0.sp.size(); i++) {
      const size_t n = (size_t)M0.sp[i].nmols;
      if (n == 0) continue;
      jobs.push_back({hs + off[0] + 3 * (size_t)M0.mol_off[i], com[i], sizeof(double) * 3 * n});
      jobs.push_back({hs + off[2] + 3 * (size_t)M0.mol_off[i], mom[i], sizeof(double) * 3 * n});
      if (M0.quat_off[i] >= 0) {
         if (!quat || !quat[i]) { mdb_set_error("mdb_group_md_upload_state: quaternions missing"); return -1; }
         jobs.push_back({hs + off[1] + 4 * (size_t)M0.quat_off[i], quat[i], sizeof(double) * 4 * n});
         jobs.push_back({hs + off[3] + 4 * (size_t)M0.quat_off[i], amom && amom[i] ? amom[i] : nullptr, sizeof(double) * 4 * n});
      }
   }
   mdb_run_copy_jobs(jobs);
   GFOR(r) {
      auto &M = g->eng[r]->mf;
      MDB_CUDA(cudaSetDevice(g->dev[r]));
      MDB_CUDA(cudaMemcpyAsync(mdb_peer_in(g->peer[r]), hs + off[0], sizeof(double) * (3 * nm + 4 * nq), cudaMemcpyHostToDevice, g->st[r]));
      MDB_CUDA(cudaMemcpyAsync(M.d_mom, hs + off[2], sizeof(double) * 3 * nm, cudaMemcpyHostToDevice, g->st[r]));
      if (nq) MDB_CUDA(cudaMemcpyAsync(M.d_amom, hs + off[3], sizeof(double) * 4 * nq, cudaMemcpyHostToDevice, g->st[r]));
   }
   GFOR(r) { MDB_CUDA(cudaSetDevice(g->dev[r])); MDB_CUDA(cudaStreamSynchronize(g->st[r])); }
   return 0;
}

static int gather_state(mdb_group *g)
{
   const auto &M0 = g->eng[0]->mf;
   if (barrier_all(g)) return -1;
   GFOR(r) {
      if (mdb_peer_in_gather_bounds(g->peer[r], 0, 3LL * M0.nmols, g->com_b, g->st[r])) return -1;
      if (M0.nmols_q > 0 && mdb_peer_in_gather_bounds(g->peer[r], 3 * (size_t)M0.nmols, 4LL * M0.nmols_q, g->quat_b, g->st[r])) return -1;
   }
   return 0;
}

extern "C" size_t mdb_group_md_scalars(const mdb_group *g) { return mdb_md_scalars(g->eng[0]); }
extern "C" const double *mdb_group_md_result(const mdb_group *g) { return g->h_md + mdb_md_scalars(g->eng[0]) * (size_t)g->world; }

// One NVE step (contract of mdb_md_step; h_scal may be NULL: mdb_group_md_result).  Blocks until the scalars are on the host.
extern "C" int mdb_group_md_step(mdb_group *g, const double h[9], double step, double ts, int surface_dipole, int do_recip,
                                 int half_sums, double *h_scal, double rdf_limit, int rdf_nbins, unsigned long long *rdf_counts)
{
   if (g->peer.empty() || !g->md_set) { mdb_set_error("mdb_group_md_step: group not configured"); return -1; }
   const int what = 1 | (do_recip ? 2 : 0);
   const size_t ns = mdb_md_scalars(g->eng[0]);
   auto each = [&](auto fn) -> int {
      GFOR(r) { MDB_CUDA(cudaSetDevice(g->dev[r])); if (fn(r, g->eng[r], g->mol_lo[r], g->mol_lo[r + 1], g->st[r])) return -1; }
      return 0;
   };
   // One host thread per rank once the buffers exist and every rank has a device of its own: each enqueues its whole step (the
   // barriers are kernels, nothing in the sequence blocks), so the ~70 launches per rank and step no longer queue behind one
   // another on one thread.  (Ranks sharing a device stay on the phase-by-phase path: a device-wide call of one rank could
   // wait there for a barrier kernel whose partner is not enqueued yet.)
   bool distinct = g->world > 1;
   GFOR(r) for (int q = 0; q < r; q++) distinct = distinct && g->dev[q] != g->dev[r];
   static const bool threads_env = !(getenv("MDB_GROUP_THREADS") && atoi(getenv("MDB_GROUP_THREADS")) == 0);
   const bool threaded = distinct && threads_env && !rdf_counts && g->md_steps > 0;
   g->md_steps++;
   if (threaded) {
      const auto &M0 = g->eng[0]->mf;
      std::vector<int> rc(g->world, 0);
      auto seq = [&](int r) {
         mdb_engine *e = g->eng[r];
         mdb_peer *p = g->peer[r];
         cudaStream_t st = g->st[r];
         const int lo = g->mol_lo[r], hi = g->mol_lo[r + 1];
         auto &M = e->mf;
         int bad = cudaSetDevice(g->dev[r]) != cudaSuccess;
         bad = bad || mdb_md_coords_range(e, h, 0.5 * step, ts, mdb_peer_in(p), lo, hi, st) || mdb_peer_barrier(p, st) ||
               mdb_peer_in_gather_bounds(p, 0, 3LL * M0.nmols, g->com_b, st) ||
               (M0.nmols_q > 0 && mdb_peer_in_gather_bounds(p, 3 * (size_t)M0.nmols, 4LL * M0.nmols_q, g->quat_b, st)) ||
               mdb_evalf_pre(e, h, mdb_peer_in(p), st) || mdb_peer_phase_a(p, what, st) || mdb_peer_barrier(p, st) ||
               mdb_peer_phase_b(p, what, st) || mdb_peer_barrier(p, st) || mdb_peer_phase_c(p, st) ||
               mdb_evalf_tail(e, h, mdb_peer_in(p), mdb_peer_result(p), lo, hi, surface_dipole, do_recip, st) ||
               mdb_md_momenta_range(e, h, 0.5 * step * ts, lo, hi, st) || (half_sums && mdb_md_sums_range(e, h, 1, false, lo, hi, st)) ||
               mdb_md_momenta_range(e, h, 0.5 * step * ts, lo, hi, st);
         for (size_t i = 0; !bad && i < M.sp.size(); i++)
            if (M.sp[i].framework && M.sp[i].nmols > 0)
               bad = cudaMemsetAsync(M.d_mom + 3 * (size_t)M.mol_off[i], 0, sizeof(double) * 3 * (size_t)M.sp[i].nmols, st) != cudaSuccess;
         bad = bad || mdb_md_coords_range(e, h, 0.5 * step, ts, mdb_peer_in(p), lo, hi, st) || mdb_md_sums_range(e, h, 0, true, lo, hi, st) ||
               cudaMemcpyAsync(M.d_mdscal, M.d_res + 3 * (size_t)M.nmols + 3 * (size_t)M.nmols_r, sizeof(double) * MDB_EVAL_SCALARS,
                               cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
               cudaMemcpyAsync(g->h_md + ns * (size_t)r, M.d_mdscal, sizeof(double) * ns, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
               cudaStreamSynchronize(st) != cudaSuccess;
         rc[r] = bad;
      };
      std::vector<std::thread> th;
      for (int r = 1; r < g->world; r++) th.emplace_back(seq, r);
      seq(0);
      for (auto &t : th) t.join();
      GFOR(r) if (rc[r]) { if (!*mdb_last_error()) mdb_set_error("mdb_group_md_step: a rank failed"); return -1; }
   } else {
   if (each([&](int r, mdb_engine *e, int lo, int hi, cudaStream_t st) { return mdb_md_coords_range(e, h, 0.5 * step, ts, mdb_peer_in(g->peer[r]), lo, hi, st); })) return -1;
   if (gather_state(g)) return -1;
   if (each([&](int r, mdb_engine *e, int, int, cudaStream_t st) { return mdb_evalf_pre(e, h, mdb_peer_in(g->peer[r]), st); })) return -1;
   GFOR(r) if (mdb_peer_phase_a(g->peer[r], what, g->st[r])) return -1;
   if (barrier_all(g)) return -1;
   GFOR(r) if (mdb_peer_phase_b(g->peer[r], what, g->st[r])) return -1;
   if (barrier_all(g)) return -1;
   GFOR(r) if (mdb_peer_phase_c(g->peer[r], g->st[r])) return -1;
   if (rdf_counts) GFOR(r) {
      MDB_CUDA(cudaSetDevice(g->dev[r]));
      if (mdb_rdf_counts(g->eng[r], rdf_limit, rdf_nbins, rdf_counts, g->st[r])) return -1;
   }
   if (each([&](int r, mdb_engine *e, int lo, int hi, cudaStream_t st) {
          return mdb_evalf_tail(e, h, mdb_peer_in(g->peer[r]), mdb_peer_result(g->peer[r]), lo, hi, surface_dipole, do_recip, st); })) return -1;
   if (each([&](int, mdb_engine *e, int lo, int hi, cudaStream_t st) { return mdb_md_momenta_range(e, h, 0.5 * step * ts, lo, hi, st); })) return -1;
   if (half_sums && each([&](int, mdb_engine *e, int lo, int hi, cudaStream_t st) { return mdb_md_sums_range(e, h, 1, false, lo, hi, st); })) return -1;
   if (each([&](int, mdb_engine *e, int lo, int hi, cudaStream_t st) { return mdb_md_momenta_range(e, h, 0.5 * step * ts, lo, hi, st); })) return -1;
   GFOR(r) {                                               /* framework constraint, src/accel.c:751-755 */
      auto &M = g->eng[r]->mf;
      MDB_CUDA(cudaSetDevice(g->dev[r]));
      for (size_t i = 0; i < M.sp.size(); i++)
         if (M.sp[i].framework && M.sp[i].nmols > 0)
            MDB_CUDA(cudaMemsetAsync(M.d_mom + 3 * (size_t)M.mol_off[i], 0, sizeof(double) * 3 * (size_t)M.sp[i].nmols, g->st[r]));
   }
   if (each([&](int r, mdb_engine *e, int lo, int hi, cudaStream_t st) { return mdb_md_coords_range(e, h, 0.5 * step, ts, mdb_peer_in(g->peer[r]), lo, hi, st); })) return -1;
   if (each([&](int, mdb_engine *e, int lo, int hi, cudaStream_t st) { return mdb_md_sums_range(e, h, 0, true, lo, hi, st); })) return -1;
   GFOR(r) {
      auto &M = g->eng[r]->mf;
      MDB_CUDA(cudaSetDevice(g->dev[r]));
      MDB_CUDA(cudaMemcpyAsync(M.d_mdscal, M.d_res + 3 * (size_t)M.nmols + 3 * (size_t)M.nmols_r, sizeof(double) * MDB_EVAL_SCALARS,
                               cudaMemcpyDeviceToDevice, g->st[r]));
      MDB_CUDA(cudaMemcpyAsync(g->h_md + ns * (size_t)r, M.d_mdscal, sizeof(double) * ns, cudaMemcpyDeviceToHost, g->st[r]));
   }
   GFOR(r) { MDB_CUDA(cudaSetDevice(g->dev[r])); MDB_CUDA(cudaStreamSynchronize(g->st[r])); }
   }
   GFOR(r) if (mdb_peer_error(g->peer[r], g->st[r]) != 0) { mdb_set_error("mdb_group: a peer barrier timed out"); return -1; }
   // combine: energies, dipole moment and stress of the force evaluation are complete and identical on every rank, its
   // virial pieces (3..11), all sums over molecules and the bad-quaternion counts add up
   double *out = g->h_md + ns * (size_t)g->world;
   memcpy(out, g->h_md, sizeof(double) * ns);
   unsigned int bad_tot = 0;
   GFOR(r) {
      const double *sr = g->h_md + ns * (size_t)r;
      unsigned int bad;
      memcpy(&bad, sr + ns - 1, sizeof bad);
      bad_tot += bad;
      if (bad) { MDB_CUDA(cudaSetDevice(g->dev[r])); MDB_CUDA(cudaMemsetAsync(g->eng[r]->mf.d_mdscal + ns - 1, 0, sizeof(double), g->st[r])); }
      if (r == 0) continue;
      for (int k = 3; k < 12; k++) out[k] += sr[k];
      for (size_t k = MDB_EVAL_SCALARS; k + 1 < ns; k++) out[k] += sr[k];
   }
   out[ns - 1] = (double)bad_tot;
   if (h_scal) memcpy(h_scal, out, sizeof(double) * ns);
   return 0;
}

// The state (and, if asked for, the molecular forces and torques of the last step) back into the caller's per-species arrays:
// every rank sends the rows of its own molecules through one pinned block.
extern "C" int mdb_group_md_download_state(mdb_group *g, double *const *com, double *const *quat, double *const *mom, double *const *amom,
                                           double *const *force, double *const *torque)
{
   if (!g->md_set) { mdb_set_error("mdb_group_md_download_state: group not configured"); return -1; }
   const auto &M0 = g->eng[0]->mf;
   const size_t nm = (size_t)M0.nmols;
   size_t off[6];
   if (group_state_block(g, off)) return -1;
   double *hs = g->h_state;
   GFOR(r) {
      auto &M = g->eng[r]->mf;
      const double *in = mdb_peer_in(g->peer[r]);
      MDB_CUDA(cudaSetDevice(g->dev[r]));
      const int m_lo = g->mol_lo[r], m_hi = g->mol_lo[r + 1];
      for (size_t i = 0; i < M.sp.size(); i++) {
         const int a = std::max(m_lo, M.mol_off[i]) - M.mol_off[i], b = std::min(m_hi, M.mol_off[i] + M.sp[i].nmols) - M.mol_off[i];
         if (b <= a) continue;
         const size_t cnt = (size_t)(b - a), mo = (size_t)M.mol_off[i] + a;
         auto d2h = [&](double *dst, const double *src, size_t n) { return cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToHost, g->st[r]); };
         MDB_CUDA(d2h(hs + off[0] + 3 * mo, in + 3 * mo, 3 * cnt));
         MDB_CUDA(d2h(hs + off[2] + 3 * mo, M.d_mom + 3 * mo, 3 * cnt));
         MDB_CUDA(d2h(hs + off[4] + 3 * mo, M.d_res + 3 * mo, 3 * cnt));
         if (M.quat_off[i] >= 0) {
            const size_t qo = (size_t)M.quat_off[i] + a;
            MDB_CUDA(d2h(hs + off[1] + 4 * qo, in + 3 * nm + 4 * qo, 4 * cnt));
            MDB_CUDA(d2h(hs + off[3] + 4 * qo, M.d_amom + 4 * qo, 4 * cnt));
         }
         if (M.torq_off[i] >= 0) {
            const size_t to = (size_t)M.torq_off[i] + a;
            MDB_CUDA(d2h(hs + off[5] + 3 * to, M.d_res + 3 * nm + 3 * to, 3 * cnt));
         }
      }
   }
   GFOR(r) { MDB_CUDA(cudaSetDevice(g->dev[r])); MDB_CUDA(cudaStreamSynchronize(g->st[r])); }
   std::vector<MdbCopyJob> jobs;
   for (size_t i = 0; i < M0.sp.size(); i++) {
      const size_t n = (size_t)M0.sp[i].nmols, mo = (size_t)M0.mol_off[i];
      if (n == 0) continue;
      if (com && com[i]) jobs.push_back({com[i], hs + off[0] + 3 * mo, sizeof(double) * 3 * n});
      if (mom && mom[i]) jobs.push_back({mom[i], hs + off[2] + 3 * mo, sizeof(double) * 3 * n});
      if (force && force[i]) jobs.push_back({force[i], hs + off[4] + 3 * mo, sizeof(double) * 3 * n});
      if (M0.quat_off[i] >= 0) {
         if (quat && quat[i]) jobs.push_back({quat[i], hs + off[1] + 4 * (size_t)M0.quat_off[i], sizeof(double) * 4 * n});
         if (amom && amom[i]) jobs.push_back({amom[i], hs + off[3] + 4 * (size_t)M0.quat_off[i], sizeof(double) * 4 * n});
      }
      if (M0.torq_off[i] >= 0 && torque && torque[i]) jobs.push_back({torque[i], hs + off[5] + 3 * (size_t)M0.torq_off[i], sizeof(double) * 3 * n});
   }
   mdb_run_copy_jobs(jobs);
   return 0;
}
