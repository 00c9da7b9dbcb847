// mdb_pair_tiled.cu -- tiled FP64 real-space pair kernel (pair_mode 3 and 4).
//
// Same pair set and arithmetic as k_pair (mdb_pair.cu), different mapping:
//   * a WARP owns a batch of <= NI consecutive cell-sorted sites i of one z-column;
//     their data sit in shared memory (re-read per use), their force accumulators in registers;
//   * the 32 LANES own neighbour sites j.  For 64 stencil runs at a time the warp
//     builds, in shared memory, the list of contiguous j-segments that the union of
//     the batch's windows covers (cells are sorted z-fastest, so a run is one
//     contiguous range per periodic piece), prefix-sums their lengths and walks the
//     flattened list 32 sites per step: lanes are always full, control flow is
//     warp-uniform, each j is loaded once and reused for the NI sites i;
//   * a (i,j) visit is masked by the exact window test  dzlo <= cz_j - cz_i <= dzhi,
//     one unsigned compare per visit (the batch spans ~2 cells, ~10 % of the visits are masked).
// mode 3 (full stencil): every reference pair is visited from both ends, forces are
//   written by their owner only -> no atomics, bit-reproducible.
// mode 4 (Newton-3): the reference's half list; the lane also accumulates the force
//   on j over the NI visits and adds it to a cell-sorted accumulator with
//   red.global.add.f64 (coalesced); results vary in the last bits run to run.  The
//   triangular part (own column, central image: j > i) is walked by a short prologue
//   so that the main loop carries no index test.
// COUNT instantiations run the same traversal and window tests without the arithmetic and only
// count the visits: mdb_pair_count() (the number of pairs handed to kernel(), src/force.c:960).
#include <type_traits>
#include "mdb_internal.h"
#include "mdb_math.cuh"

#ifndef MDB_TILED_MINB
#define MDB_TILED_MINB 3
#endif
static constexpr int TW = 4;                   // warps per block
static constexpr int RB = 64;                  // stencil runs per shared-memory pass
static constexpr int NSLOT = 3 * RB;           // segment slots per pass (3 periodic pieces per run)
static constexpr int NI = MDB_NI;
static constexpr int NRED = 8;
static constexpr int BIGZ = 1 << 20;           // larger than any cell index / window offset

// ---- shared-memory layout (dynamic, one block): every access goes through a 32-bit shared
// address held in a register plus a compile-time offset, so the hot loop never re-derives a base.
struct WarpSm {                                // per warp
   int4 q[NSLOT];                              // {jstart - pre, dzlo - zoff, dzhi - dzlo, kimg | selfcol<<5}
   int pre[NSLOT + 8];                         // running prefix of the segment lengths, INT_MAX sentinels behind
   double4 ipos[NI];                           // batch sites: x,y,z,q  (warp-uniform, re-read per use to keep
   int4 iint[NI];                              //   them out of registers): cz, row offset, framework flag, sorted index
};
static constexpr int OFF_Q = 0, OFF_PRE = OFF_Q + 16 * NSLOT, OFF_IPOS = OFF_PRE + 4 * (NSLOT + 8),
                     OFF_IINT = OFF_IPOS + 32 * NI, WARP_SM = OFF_IINT + 16 * NI;
static_assert(sizeof(WarpSm) == WARP_SM && WARP_SM % 16 == 0, "WarpSm layout");
static constexpr int OFF_RELOC = TW * WARP_SM;           // double4[27]: image translations
static constexpr int NE2 = MDB_EXP_F32TAIL ? 256 : 64;    // entries of the exp table in shared memory
static constexpr int OFF_E2 = OFF_RELOC + 32 * 28;       // double[NE2]: 2^(j/NE2), mdb_exp_tab_v / mdb_exp_tab8_v
static constexpr int OFF_TAB = OFF_E2 + 8 * NE2;         // pair-parameter table
static constexpr size_t TILED_SMEM = OFF_TAB;

template <int OFF> __device__ __forceinline__ int lds_i(unsigned a)
{
   int v;
   asm volatile("ld.shared.s32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
   return v;
}
template <int OFF> __device__ __forceinline__ int4 lds_i4(unsigned a)
{
   int4 v;
   asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a), "n"(OFF));
   return v;
}
template <int OFF> __device__ __forceinline__ double4 lds_d4(unsigned a)
{
   double4 v;
   asm volatile("ld.shared.v2.f64 {%0,%1}, [%4+%5];\n\tld.shared.v2.f64 {%2,%3}, [%4+%5+16];"
                : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "r"(a), "n"(OFF));
   return v;
}
template <int OFF> __device__ __forceinline__ void lds_d3(unsigned a, double &x, double &y, double &z)
{
   asm volatile("ld.shared.v2.f64 {%0,%1}, [%3+%4];\n\tld.shared.f64 %2, [%3+%4+16];"
                : "=d"(x), "=d"(y), "=d"(z) : "r"(a), "n"(OFF));
}
template <int OFF> __device__ __forceinline__ void sts_i(unsigned a, int v)
{
   asm volatile("st.shared.s32 [%0+%1], %2;" :: "r"(a), "n"(OFF), "r"(v) : "memory");
}
template <int OFF> __device__ __forceinline__ void sts_i4(unsigned a, int4 v)
{
   asm volatile("st.shared.v4.s32 [%0+%1], {%2,%3,%4,%5};" :: "r"(a), "n"(OFF), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ double wsum(double v)
{
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
   return v;
}

struct JBuf {                                  // one lane's neighbour for one step (double-buffered)
   double4 pj;
   int2 sj;                                    // {type | framework bit, z index}
   int j;
   unsigned qa;                                // shared address of the segment's q entry
};

enum { TM_FORCE = 0, TM_COUNT = 1, TM_RDF = 2 };     // what a visit does
struct RdfParams { double rbin; int nbins, hist_smem; unsigned long long *counts; };
// How warps get their batches.  mode 0: warp (block, w) owns batch block * warps + w (one batch per warp).
// mode 1 ("filler", the QUEUE instantiations): a small persistent grid (one block per SM) draws batches from the counter `next` until the list is
// exhausted or `stop` is raised; it runs beside the k-space GEMM kernels, whose DMMA stream leaves a quarter of the FP64
// pipe idle.  mode 2 ("remainder"): the static mapping again, starting at the batch the filler stopped at.
struct PairQueue { int mode; int *next; const int *stop; int prow0; int nopro; };
// nopro = 1: a launch over a subset of the stencil runs that does not contain run 0 (the far runs): no Newton-3 prologue, and
// its first run is an ordinary one.

template <int PT, bool COUL, bool STRICT, bool FW, bool N3, int MODE, bool QUEUE = false>
// (the power-law potential-only instantiations fit 128 registers without spills: 4 blocks per SM, 2.10 -> 1.92 ms for the LJ
//  pass; with the Coulomb term 4 blocks spill and lose, 16.6 -> 17.9 ms; the exponential potentials alone (MCY: 200-300
//  bytes of spills, LDL/STL inside the hot loop at 128 registers) stay at 3 blocks as well)
__global__ void __launch_bounds__(TW * 32, (COUL || MODE != TM_FORCE || (PT != PT_LJ && PT != PT_HIW)) ? MDB_TILED_MINB : MDB_TILED_MINB + 1)
k_pair_tiled(PairParams P, int nsites, int nout, const double4 *__restrict__ posq, const int2 *__restrict__ sinfo,
             const int *__restrict__ cstart, const int *__restrict__ order,
             const int *__restrict__ mol, const StencilRun *__restrict__ runs, const double *__restrict__ ptab,
             const int2 *__restrict__ batches, const int *__restrict__ nbatch_p, int rank, int nranks,
             double *__restrict__ out, double *__restrict__ fs, double *__restrict__ partials,
             unsigned long long *__restrict__ counters, RdfParams R, PairQueue Q)
{
   constexpr bool COUNT = MODE == TM_COUNT;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   {
      double4 *s_reloc = reinterpret_cast<double4 *>(smem_raw + OFF_RELOC);
      double *s_e2 = reinterpret_cast<double *>(smem_raw + OFF_E2), *s_ptab = reinterpret_cast<double *>(smem_raw + OFF_TAB);
      for (int k = threadIdx.x; k < 27; k += blockDim.x) s_reloc[k] = make_double4(P.reloc[k][0], P.reloc[k][1], P.reloc[k][2], 0.0);
      for (int k = threadIdx.x; k < NE2; k += blockDim.x) s_e2[k] = MDB_EXP_F32TAIL ? c_exp2tab256[k & (MDB_EXP_F32TAIL ? 255 : 0)] : c_exp2tab[k & 63];
      if (MODE == TM_RDF) {                      // block-local histogram in place of the pair table
         unsigned int *hist = reinterpret_cast<unsigned int *>(s_ptab);
         const int nh = R.hist_smem ? R.nbins * (P.max_id * (P.max_id - 1) / 2) : 0;
         for (int k = threadIdx.x; k < nh; k += blockDim.x) hist[k] = 0u;
      } else {
         const int ntab = P.max_id * P.max_id * MDB_NPOTP;
         for (int k = threadIdx.x; k < ntab; k += blockDim.x) s_ptab[k] = ptab[k];
      }
   }
   __syncthreads();
   // the two base addresses of the hot loop; passed through a shuffle so that they stay in registers
   // (a plain symbol address is re-derived from SR_CgaCtaId at every use under register pressure)
   const unsigned sb = __shfl_sync(0xffffffffu, (unsigned)__cvta_generic_to_shared(smem_raw), 0);
   const unsigned sw = __shfl_sync(0xffffffffu, sb + w * WARP_SM, 0);

   double fix[NI], fiy[NI], fiz[NI];
   double pe = 0, w00 = 0, w01 = 0, w02 = 0, w11 = 0, w12 = 0, w22 = 0;
   unsigned int visits = 0;
   int s0 = 0, cnt = 1, col = 0, cx = 0, cy = 0, cz_lo = 0, cz_hi = 0, nruns = 0;
   bool idle = false;

   // ---- one step: lane's neighbour j against the NI batch sites --------------------------------
   // u = cz_j - (dzlo - zoff) (or a value no window accepts for padding lanes), wid = dzhi - dzlo:
   // the visit (i,j) is inside the stencil run iff (unsigned)(u - cz_i) <= wid.
   auto step = [&](auto selft_tag, const JBuf &J, const bool valid) {
      constexpr bool SELFT = decltype(selft_tag)::value;
      const int4 sq = lds_i4<OFF_Q>(J.qa);
      const int j = J.j, wid = sq.z;
      const int u = valid ? J.sj.y - sq.y : 3 * BIGZ;
      const int kimg = sq.w & 31;
      const int fwj = J.sj.x >> 30;
      const unsigned tjb = sb + OFF_TAB + (unsigned)(J.sj.x & 0x3fffffff) * (MDB_NPOTP * 8);
      // mode 3: j == s_i can only happen in the own column's central image
      const int jself = N3 ? j : (sq.w == (13 | 32) ? j : -1);
      bool in[NI];
      if (COUNT) {
#pragma unroll
         for (int k = 0; k < NI; k++) {
            const int4 ik = k == 0 ? lds_i4<OFF_IINT>(sw) : k == 1 ? lds_i4<OFF_IINT + 16>(sw)
                          : k == 2 ? lds_i4<OFF_IINT + 32>(sw) : lds_i4<OFF_IINT + 48>(sw);
            in[k] = (unsigned)(u - ik.x) <= (unsigned)wid;
            if (SELFT) in[k] = in[k] && (N3 ? jself > ik.w : jself != ik.w);
            if (FW) in[k] = in[k] && !(ik.z & fwj);
            visits += in[k] ? 1u : 0u;
         }
         return;
      }
      if (MODE == TM_RDF) {
         // rdf_inner + rdf_accum (src/force.c:1084-1096, src/rdf.c:94-108): the reference's operation
         // order with explicit roundings so that every pair lands in the same bin, bit for bit
         double rlx, rly, rlz;
         lds_d3<OFF_RELOC>(sb + 32 * kimg, rlx, rly, rlz);
         const int tj = J.sj.x & 0x3fffffff;
         unsigned int *hist = reinterpret_cast<unsigned int *>(smem_raw + OFF_TAB);
#pragma unroll
         for (int k = 0; k < NI; k++) {
            const double4 pi = k == 0 ? lds_d4<OFF_IPOS>(sw) : k == 1 ? lds_d4<OFF_IPOS + 32>(sw)
                             : k == 2 ? lds_d4<OFF_IPOS + 64>(sw) : lds_d4<OFF_IPOS + 96>(sw);
            const int4 ik = k == 0 ? lds_i4<OFF_IINT>(sw) : k == 1 ? lds_i4<OFF_IINT + 16>(sw)
                          : k == 2 ? lds_i4<OFF_IINT + 32>(sw) : lds_i4<OFF_IINT + 48>(sw);
            in[k] = (unsigned)(u - ik.x) <= (unsigned)wid;
            if (SELFT) in[k] = in[k] && (N3 ? jself > ik.w : jself != ik.w);
            if (FW) in[k] = in[k] && !(ik.z & fwj);
            const double rx = __dadd_rn(__dsub_rn(J.pj.x, pi.x), rlx), ry = __dadd_rn(__dsub_rn(J.pj.y, pi.y), rly),
                         rz = __dadd_rn(__dsub_rn(J.pj.z, pi.z), rlz);
            const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
            const int bin = (int)__dmul_rn(R.rbin, __dsqrt_rn(rsq));
            if (in[k] && bin < R.nbins && min(ik.y, tj) >= 1) {
               const int a = min(ik.y, tj), bb = max(ik.y, tj);          // rdf[idi][idj] = rdf[idj][idi], src/rdf.c:82-90
               const int slot = ((a - 1) * P.max_id - (a - 1) * a / 2 + (bb - a)) * R.nbins + bin;
               if (R.hist_smem) atomicAdd(&hist[slot], 1u);
               else atomicAdd(&R.counts[slot], 1ULL);
            }
         }
         return;
      }
      double4 pj = J.pj;
      double rlx, rly, rlz;
      lds_d3<OFF_RELOC>(sb + 32 * kimg, rlx, rly, rlz);
      pj.x += rlx; pj.y += rly; pj.z += rlz;
      double gx = 0, gy = 0, gz = 0;
      double dx[NI], dy[NI], dzz[NI], r2[NI], qq[NI], fij[NI], phi[NI];
      MdbSmemRow prow[NI];
#pragma unroll
      for (int k = 0; k < NI; k++) {
         const double4 pi = k == 0 ? lds_d4<OFF_IPOS>(sw) : k == 1 ? lds_d4<OFF_IPOS + 32>(sw)
                          : k == 2 ? lds_d4<OFF_IPOS + 64>(sw) : lds_d4<OFF_IPOS + 96>(sw);
         const int4 ik = k == 0 ? lds_i4<OFF_IINT>(sw) : k == 1 ? lds_i4<OFF_IINT + 16>(sw)
                       : k == 2 ? lds_i4<OFF_IINT + 32>(sw) : lds_i4<OFF_IINT + 48>(sw);
         in[k] = (unsigned)(u - ik.x) <= (unsigned)wid;
         if (SELFT) in[k] = in[k] && (N3 ? jself > ik.w : jself != ik.w);
         if (FW) in[k] = in[k] && !(ik.z & fwj);
         dx[k] = pj.x - pi.x; dy[k] = pj.y - pi.y; dzz[k] = pj.z - pi.z;
         qq[k] = pi.w * pj.w;
         prow[k].a = tjb + (unsigned)ik.y;
      }
#pragma unroll
      for (int k = 0; k < NI; k++) r2[k] = fma(dx[k], dx[k], fma(dy[k], dy[k], dzz[k] * dzz[k]));
      int r2min = __double2hiint(r2[0]);
#pragma unroll
      for (int k = 1; k < NI; k++) r2min = min(r2min, __double2hiint(r2[k]));
      int close = 0;
      if (r2min < 0x3fd00000) {                 // some r^2 < 0.25 in this step (rare): look closer
#pragma unroll
         for (int k = 0; k < NI; k++) close |= (in[k] && r2[k] < MDB_TOO_CLOSE) ? (1 << k) : 0;
      }
      if (STRICT) {
#pragma unroll
         for (int k = 0; k < NI; k++) r2[k] = r2[k] > P.cutoffsq ? P.cutoff100sq : r2[k];
      }
#ifdef MDB_EXP_NOMATH                           /* experiment: traversal + accumulation without the pair arithmetic */
#pragma unroll
      for (int k = 0; k < NI; k++) { fij[k] = r2[k] * prow[k][0]; phi[k] = qq[k]; }
#else
      mdb_pair_eval_v<PT, COUL, NI, true>(r2, qq, prow, P.alpha, P.norm, fij, phi, sb + OFF_E2);
#endif
#pragma unroll
      for (int k = 0; k < NI; k++) {
         if (in[k]) {                         // predicated FP64 accumulation, no selects
            const double f = fij[k];
            pe += phi[k];
            fix[k] = fma(-f, dx[k], fix[k]); fiy[k] = fma(-f, dy[k], fiy[k]); fiz[k] = fma(-f, dzz[k], fiz[k]);
            gx = fma(f, dx[k], gx); gy = fma(f, dy[k], gy); gz = fma(f, dzz[k], gz);
         }
      }
      if (close) {                           // TOO_CLOSE diagnostics, off the hot path (src/force.c:939-949)
         const int mj = mol[order[j]];
         for (int k = 0; k < NI; k++)
            if (((close >> k) & 1) && mol[order[s0 + k]] != mj) {
               atomicAdd(&counters[1], N3 ? 2ULL : 1ULL);
               counters[3] = ((unsigned long long)(unsigned)order[s0 + k] << 32) | (unsigned)order[j];
            }
      }
#ifdef MDB_EXP_NORED                            /* experiment: no scatter of the force on j */
      pe += gx + gy + gz;
#else
      if (N3) {
         if (valid) {
            atomicAdd(&fs[j], gx);
            atomicAdd(&fs[(size_t)nsites + j], gy);
            atomicAdd(&fs[2 * (size_t)nsites + j], gz);
         }
      }
#endif
      if (kimg != 13) {                      // Bekker image-force virial (src/force.c:983-991)
         const double sc = N3 ? 1.0 : 0.5;
         const double rx = sc * rlx, ry = sc * rly, rz = sc * rlz;
         w00 = fma(rx, gx, w00); w01 = fma(ry, gx, w01); w02 = fma(rz, gx, w02);
         w11 = fma(ry, gy, w11); w12 = fma(rz, gy, w12); w22 = fma(rz, gz, w22);
      }
   };

   for (;;) {                                   // one batch per pass (filler: until the list is exhausted or `stop` is raised)
   // (bounds and the queue state are re-derived here so that nothing of them stays in registers across the pass)
   const int nbatch = *nbatch_p;
   const int b_lo = (int)((long)nbatch * rank / nranks), b_hi = (int)((long)nbatch * (rank + 1) / nranks);
   int b;
   if (QUEUE) {
      int t = 0;
      if (lane == 0) t = *reinterpret_cast<const volatile int *>(Q.stop) ? 0x3fffffff : atomicAdd(Q.next, 1);
      b = b_lo + min(__shfl_sync(0xffffffffu, t, 0), 0x3fffffff);
   } else
      b = b_lo + (Q.mode == 2 ? min(*Q.next, b_hi - b_lo) : 0) + (int)blockIdx.x * (int)(blockDim.x >> 5) + w;
   idle = b >= b_hi;
   if (idle) break;                             // (an idle warp still publishes its zero row below)
   {
      const int2 bt = batches[b];               // {first sorted site, count | column << 3}
      s0 = bt.x; cnt = bt.y & 7; col = bt.y >> 3;
      cx = col / P.ny; cy = col - cx * P.ny;
      cz_lo = sinfo[s0].y; cz_hi = sinfo[s0 + cnt - 1].y;
      // ---- batch (warp-uniform) data
      if (lane < NI) {
         const int k = lane, sk = s0 + min(k, cnt - 1);
         WarpSm *W = reinterpret_cast<WarpSm *>(smem_raw) + w;
         W->ipos[k] = posq[sk];
         const int2 si = sinfo[sk];
         // padded entries never pass the window test; row offset in bytes from the table start
         W->iint[k] = make_int4(k < cnt ? si.y : BIGZ,
                                MODE == TM_RDF ? (si.x & 0x3fffffff) : (si.x & 0x3fffffff) * P.max_id * (MDB_NPOTP * 8),
                                si.x >> 30, k < cnt ? s0 + k : 0x7fffffff);
      }
#pragma unroll
      for (int k = 0; k < NI; k++) fix[k] = fiy[k] = fiz[k] = 0.0;
      __syncwarp();
      nruns = P.nruns;
   }

   // ---- Newton-3 prologue: own column, central image, j > s_i (run 0 of the half list starts at dz = 0)
   if (N3 && !Q.nopro) {
      const StencilRun r0 = runs[0];
      const int jb = s0 + 1, je = cstart[col * P.nz + min(cz_hi + r0.dzhi, P.nz - 1) + 1];
      if (lane == 0) sts_i4<OFF_Q>(sw, make_int4(0, -BIGZ, r0.dzhi + BIGZ, 13 | 32));
      __syncwarp();
      for (int pj = jb + lane; pj - lane < je; pj += 32) {
         JBuf J;
         J.j = min(pj, je - 1);
         J.sj = sinfo[J.j];
         J.pj = posq[J.j];
         J.qa = sw;
         step(std::true_type{}, J, pj < je);
      }
   }

   for (int rb = 0; rb < nruns; rb += RB) {
      // ---- segment table for runs rb .. rb+RB-1: lane handles runs rb+lane and rb+32+lane
      __syncwarp();
      int total = 0, nseg = 0;
#pragma unroll
      for (int hh = 0; hh < RB / 32; hh++) {
         const int lr = hh * 32 + lane, r = rb + lr;
         int cnt3[3] = {0, 0, 0}, js3[3] = {0, 0, 0}, img3[3] = {0, 0, 0}, dzlo = 0, dzw = 0;
         if (r < nruns) {
            const StencilRun run = runs[r];
            int tx = cx + run.dx, ty = cy + run.dy, ii = 0, jj = 0;
            if (tx < 0) { tx += P.nx; ii = -1; } else if (tx >= P.nx) { tx -= P.nx; ii = 1; }
            if (ty < 0) { ty += P.ny; jj = -1; } else if (ty >= P.ny) { ty -= P.ny; jj = 1; }
            const int colb = (tx * P.ny + ty) * P.nz;
            const int z0 = cz_lo + run.dzlo, z1 = cz_hi + run.dzhi;
            const int selfcol = (run.dx == 0 && run.dy == 0) ? 1 : 0;
            dzlo = run.dzlo; dzw = run.dzhi - run.dzlo;
#pragma unroll
            for (int kk = -1; kk <= 1; kk++) {
               const int zoff = kk * P.nz;
               const int a = max(z0, zoff), bb = min(z1, zoff + P.nz - 1);
               // Newton-3: the central piece of run 0 is the prologue's
               if (a <= bb && !(N3 && r == 0 && kk == 0 && !Q.nopro)) {
                  const int jb = cstart[colb + a - zoff], jn = cstart[colb + bb - zoff + 1];
                  js3[kk + 1] = jb; cnt3[kk + 1] = jn - jb;
                  img3[kk + 1] = (9 * (ii + 1) + 3 * (jj + 1) + (kk + 1)) | (selfcol << 5);
               }
            }
         }
         // append the non-empty pieces (compacted) and their running prefix
         const int ne = (cnt3[0] > 0) + (cnt3[1] > 0) + (cnt3[2] > 0);
         const int tot = cnt3[0] + cnt3[1] + cnt3[2];
         int inc = tot, ninc = ne;
#pragma unroll
         for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d), u = __shfl_up_sync(0xffffffffu, ninc, d);
            if (lane >= d) { inc += t; ninc += u; }
         }
         int base = total + inc - tot, slot = nseg + ninc - ne;
#pragma unroll
         for (int q = 0; q < 3; q++)
            if (cnt3[q] > 0) {
               sts_i4<OFF_Q>(sw + 16 * slot, make_int4(js3[q] - base, dzlo - (q - 1) * P.nz, dzw, img3[q]));
               sts_i<OFF_PRE>(sw + 4 * slot, base);
               base += cnt3[q];
               slot++;
            }
         total += __shfl_sync(0xffffffffu, inc, 31);
         nseg += __shfl_sync(0xffffffffu, ninc, 31);
      }
      if (lane < 4) sts_i<OFF_PRE>(sw + 4 * (nseg + lane), lane == 0 ? total : 0x7fffffff);
      __syncwarp();

      // ---- walk the flattened j list, 32 sites per step: the loads of step n+1 are issued
      // before the arithmetic of step n (uniform control flow makes this free)
      unsigned sega = sw;                       // address of this lane's current segment: pre[seg] at sega + OFF_PRE
      auto fetch = [&](int p, JBuf &J) {
         const int pp = min(p, total - 1);
         // mostly 0..2 segment boundaries per step: three independent look-ahead loads, a loop only beyond
         const int a1 = lds_i<OFF_PRE + 4>(sega), a2 = lds_i<OFF_PRE + 8>(sega), a3 = lds_i<OFF_PRE + 12>(sega);
         sega += (pp >= a1 ? 4u : 0u) + (pp >= a2 ? 4u : 0u) + (pp >= a3 ? 4u : 0u);
         if (pp >= a3)
            while (pp >= lds_i<OFF_PRE + 4>(sega)) sega += 4;
         J.qa = sw + 4 * (sega - sw);
         J.j = lds_i<OFF_Q>(J.qa) + pp;
#ifdef MDB_EXP_NOFETCH                          /* experiment: no per-step global loads */
         J.pj = make_double4(0.37 * J.j, 1.1 * lane, 3.0, 0.4); J.sj = make_int2(lane & 3, cz_lo);
#else
         J.pj = posq[J.j];
         J.sj = sinfo[J.j];
#endif
      };
      JBuf A, B;
      int pl = lane;                            // this lane's position in the flattened list (step of A)
      if (total > 0) fetch(pl, A);
      // (a two-step unrolled ping-pong of A/B measured 10 % slower: spills and a 17 KB loop body)
      for (int base = 0; base < total; base += 32, pl += 32) {
         B = A;
         // unconditional (the position is clamped inside): with the loads under a branch ptxas makes the
         // first instruction after the join wait for them, which serialises the prefetch with the step
         fetch(pl + 32, A);
         step(std::integral_constant<bool, !N3>{}, B, pl < total);
      }
   }

   // ---- forces on the batch sites: reduce the per-lane partial sums
   if (MODE == TM_FORCE) {
      const WarpSm *W = reinterpret_cast<const WarpSm *>(smem_raw) + w;
#pragma unroll
      for (int k = 0; k < NI; k++) {
         const double fx = wsum(fix[k]), fy = wsum(fiy[k]), fz = wsum(fiz[k]);
         if (lane == 0 && k < cnt) {
            if (N3) {
               atomicAdd(&fs[s0 + k], fx);
               atomicAdd(&fs[(size_t)nsites + s0 + k], fy);
               atomicAdd(&fs[2 * (size_t)nsites + s0 + k], fz);
            } else {
               const int o = order[s0 + k];          // nsites = length of this pass's site list, nout = all sites
               out[o] += fx;
               out[(size_t)nout + o] += fy;
               out[2 * (size_t)nout + o] += fz;
               const double4 pk4 = W->ipos[k];
               const double px = pk4.x, py = pk4.y, pz = pk4.z;
               w00 = fma(px, fx, w00); w01 = fma(py, fx, w01); w02 = fma(pz, fx, w02);
               w11 = fma(py, fy, w11); w12 = fma(pz, fy, w12); w22 = fma(pz, fz, w22);
            }
         }
      }
   }
   if (!QUEUE) break;
   __syncwarp();                                // the batch data in shared memory are rewritten by the next pass
   }

   if (MODE == TM_RDF) {
      __syncthreads();
      if (R.hist_smem) {
         const unsigned int *hist = reinterpret_cast<const unsigned int *>(smem_raw + OFF_TAB);
         const int nh = R.nbins * (P.max_id * (P.max_id - 1) / 2);
         for (int k = threadIdx.x; k < nh; k += blockDim.x)
            if (hist[k]) atomicAdd(&R.counts[k], (unsigned long long)hist[k]);
      }
      return;
   }
   if (COUNT) {
      unsigned int vs = visits;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, d);
      if (lane == 0) atomicAdd(&counters[0], (unsigned long long)vs * (N3 ? 2ULL : 1ULL));
      return;
   }
   double v[7] = {pe, w00, w01, w02, w11, w12, w22};
#pragma unroll
   for (int k = 0; k < 7; k++) {
      const double t = wsum(v[k]);
      if (lane == 0) partials[((size_t)Q.prow0 + (size_t)blockIdx.x * (blockDim.x >> 5) + w) * NRED + k] = t;
   }
}

// Newton-3 mode: move the cell-sorted accumulator back to the caller's site order and add the
// site virial sum_i r_i (x) F_i (src/force.c:973-982), one row of partial sums per block.
__global__ void __launch_bounds__(256) k_unsort_virial(int n, int nout, const double4 *__restrict__ posq,
                                                       const int *__restrict__ order, const double *__restrict__ fs,
                                                       double *__restrict__ out, double *__restrict__ partials)
{
   const int s = blockIdx.x * 256 + threadIdx.x;
   double v[6] = {0, 0, 0, 0, 0, 0};
   if (s < n) {
      const double fx = fs[s], fy = fs[(size_t)n + s], fz = fs[2 * (size_t)n + s];
      const double4 p = posq[s];
      const int o = order[s];
      out[o] += fx; out[(size_t)nout + o] += fy; out[2 * (size_t)nout + o] += fz;
      v[0] = p.x * fx; v[1] = p.y * fx; v[2] = p.z * fx; v[3] = p.y * fy; v[4] = p.z * fy; v[5] = p.z * fz;
   }
   __shared__ double red[8][6];
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
   for (int k = 0; k < 6; k++) {
      const double t = wsum(v[k]);
      if (lane == 0) red[w][k] = t;
   }
   __syncthreads();
   if (threadIdx.x < 6) {
      double t = 0;
      for (int k = 0; k < 8; k++) t += red[k][threadIdx.x];
      partials[(size_t)blockIdx.x * NRED + 1 + threadIdx.x] = t;
   }
   if (threadIdx.x == 6) partials[(size_t)blockIdx.x * NRED] = 0.0;
}

// fixed-order sum of partial rows, two stages: RS1 blocks each fold a contiguous span of rows
// into one row, then one block folds those and adds pe_real (scaled) and the stress sums.
static constexpr int RS1 = 128;
__global__ void __launch_bounds__(256) k_rows_stage1(const double *__restrict__ partials, int nrows,
                                                     double *__restrict__ rows1)
{
   __shared__ double sm[256];
   const int per = (nrows + RS1 - 1) / RS1;
   const int r0 = blockIdx.x * per, r1 = min(nrows, r0 + per);
   double acc[7] = {0, 0, 0, 0, 0, 0, 0};
   for (int b = r0 + threadIdx.x; b < r1; b += 256)
#pragma unroll
      for (int k = 0; k < 7; k++) acc[k] += partials[(size_t)b * NRED + k];
   for (int k = 0; k < 7; k++) {
      sm[threadIdx.x] = acc[k];
      __syncthreads();
      for (int d = 128; d > 0; d >>= 1) {
         if (threadIdx.x < d) sm[threadIdx.x] += sm[threadIdx.x + d];
         __syncthreads();
      }
      if (threadIdx.x == 0) rows1[(size_t)blockIdx.x * NRED + k] = sm[0];
      __syncthreads();
   }
}

__global__ void __launch_bounds__(RS1) k_rows_finish(const double *__restrict__ rows1, int nsites, double pe_scale,
                                                     double *__restrict__ out)
{
   __shared__ double sm[RS1];
   double tot[7];
   for (int k = 0; k < 7; k++) {
      sm[threadIdx.x] = rows1[(size_t)threadIdx.x * NRED + k];
      __syncthreads();
      for (int d = RS1 / 2; d > 0; d >>= 1) {
         if (threadIdx.x < d) sm[threadIdx.x] += sm[threadIdx.x + d];
         __syncthreads();
      }
      tot[k] = sm[0];
      __syncthreads();
   }
   if (threadIdx.x == 0) {
      double *sc = out + 3 * (size_t)nsites;
      sc[0] += pe_scale * tot[0];
      sc[2 + 0] += tot[1]; sc[2 + 1] += tot[2]; sc[2 + 2] += tot[3];
      sc[2 + 4] += tot[4]; sc[2 + 5] += tot[5]; sc[2 + 8] += tot[6];
   }
}

// the cell-sorted site list a pass runs on: all sites, or one site class (SubList)
struct SiteList {
   int n; const double4 *posq; const int2 *sinfo; const int *start; const int *order; const int2 *batches; const int *nbatch;
   double *fs;
};
static SiteList full_list(const mdb_engine *e)
{
   return SiteList{e->cfg.nsites, e->d_posq, e->d_sinfo, e->d_start, e->d_order, e->d_batches, e->d_nbatch, e->d_fs};
}
static SiteList sub_list(const mdb_engine *e, int k)
{
   const SubList &S = e->sub[k];
   return SiteList{S.n, S.posq, S.sinfo, S.start, S.order, S.batches, S.nbatch, S.fs};
}

#define TILED_ARGS P, L.n, c.nsites, L.posq, L.sinfo, L.start, L.order, e->d_mol, runs, (ptab_dev ? ptab_dev : e->d_ptab), \
                   L.batches, L.nbatch, e->ithread, e->nthreads, d_out, L.fs, e->d_partials, e->d_counters, R, Q

template <int PT, bool COUL, int MODE>
static void launch_tiled(bool strict, bool fw, bool n3, dim3 g, cudaStream_t st, PairParams &P, mdb_engine *e, const SiteList &L,
                         const StencilRun *runs, double *d_out, RdfParams R = RdfParams{0.0, 0, 0, nullptr}, size_t shm_extra = 0,
                         PairQueue Q = PairQueue{0, nullptr, nullptr, 0, 0}, int threads = TW * 32, const double *ptab_dev = nullptr)
{
   const mdb_config &c = e->cfg;
   const size_t shm = TILED_SMEM + (MODE == TM_RDF ? shm_extra : sizeof(double) * MDB_NPOTP * (size_t)c.max_id * c.max_id);
   // The filler shares its SMs with the k-space GEMM blocks (90-220 KB of shared memory each): an SM has ONE L1/shared split at
   // a time, so both sides ask for the largest shared-memory carve-out -- otherwise a filler block only gets onto an SM
   // between two GEMM blocks, when the split can change.
#define GO(S, F, N) do { if (MODE == TM_FORCE && N && Q.mode == 1) { \
                           auto kq = k_pair_tiled<PT, COUL, S, F, N, MODE, (MODE == TM_FORCE && N)>; \
                           cudaFuncSetAttribute(kq, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
                           kq<<<g, threads, shm, st>>>(TILED_ARGS); \
                        } else k_pair_tiled<PT, COUL, S, F, N, MODE, false><<<g, threads, shm, st>>>(TILED_ARGS); } while (0)
   if (strict && MODE == TM_FORCE) {
      if (fw) { if (n3) GO(true, true, true); else GO(true, true, false); }
      else    { if (n3) GO(true, false, true); else GO(true, false, false); }
   } else {
      if (fw) { if (n3) GO(false, true, true); else GO(false, true, false); }
      else    { if (n3) GO(false, false, true); else GO(false, false, false); }
   }
   (void)threads;
#undef GO
}

static void tiled_params(mdb_engine *e, bool n3, int nlist, PairParams &P, const StencilRun *&runs, int &nblocks)
{
   const mdb_config &c = e->cfg;
   P.nx = e->T.nx; P.ny = e->T.ny; P.nz = e->T.nz;
   P.nruns = n3 ? e->nruns_half : e->nruns;
   runs = n3 ? e->d_runs_half : e->d_runs;
   for (int k = 0; k < 27; k++)
      for (int a = 0; a < 3; a++) P.reloc[k][a] = e->T.reloc[k][a];
   P.alpha = c.alpha;
   P.norm = 2.0 * c.alpha / sqrt(MDB_PI);
   P.cutoffsq = c.cutoff * c.cutoff;
   P.cutoff100sq = 10000.0 * P.cutoffsq;
   P.max_id = c.max_id;
   P.strict = c.strict_cutoff;
   P.s_lo = 0; P.s_hi = nlist;
   // upper bound on this rank's batches (the exact count lives on the device)
   const long nb_max = (long)nlist / NI + (long)e->T.nx * e->T.ny + 1;
   const int my_max = (int)((nb_max + e->nthreads - 1) / e->nthreads) + 1;
   nblocks = (my_max + TW - 1) / TW;
}

// pairs handed to kernel() for the current cells: the traversal of the force kernel without its arithmetic
int mdb_launch_pair_count_tiled(mdb_engine *e, cudaStream_t st)
{
   const mdb_config &c = e->cfg;
   const bool n3 = e->pair_mode == 4, fw = c.nsites_xf < c.nsites;
   PairParams P;
   const StencilRun *runs;
   int nblocks;
   if (mdb_need_batches(e, st)) return -1;
   tiled_params(e, n3, c.nsites, P, runs, nblocks);
   double *d_out = nullptr;
   const SiteList L = full_list(e);
   launch_tiled<PT_LJ, false, TM_COUNT>(false, fw, n3, dim3(nblocks), st, P, e, L, runs, d_out);
   e->launches += 1;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

// RDF binning pass (src/force.c:1302-1313): the Newton-3 traversal over the strict stencil of radius
// `limit` (runs built by the caller), counts[pair(idi,idj)][bin] += 1 for this rank's batches
int mdb_launch_rdf_tiled(mdb_engine *e, const StencilRun *d_runs, int nruns, double rbin, int nbins,
                         unsigned long long *d_counts, cudaStream_t st)
{
   const mdb_config &c = e->cfg;
   const bool fw = c.nsites_xf < c.nsites;
   PairParams P;
   const StencilRun *runs;
   int nblocks;
   if (mdb_need_batches(e, st)) return -1;
   tiled_params(e, true, c.nsites, P, runs, nblocks);
   runs = d_runs; P.nruns = nruns;
   const SiteList L = full_list(e);
   RdfParams R;
   R.rbin = rbin; R.nbins = nbins; R.counts = d_counts;
   const size_t hbytes = sizeof(unsigned int) * (size_t)nbins * (c.max_id * (c.max_id - 1) / 2);
   R.hist_smem = hbytes <= MDB_TILED_TAB_MAX ? 1 : 0;
   double *d_out = nullptr;
   launch_tiled<PT_LJ, false, TM_RDF>(false, fw, true, dim3(nblocks), st, P, e, L, runs, d_out, R, R.hist_smem ? hbytes : 0);
   e->launches += 1;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

// one pass of the force kernel over a site list: pair potential `ptype` (PT_NONE: none) and/or the Coulomb term.
// filler: the pass is split into a small persistent grid that draws batches from a counter until the k-space stream
// raises the stop flag (mdb_force_both), and a full-size launch for the batches that are left.
static int pair_pass(mdb_engine *e, const SiteList &L, int ptype, bool coul, double *d_out, cudaStream_t st, bool filler = false)
{
   const mdb_config &c = e->cfg;
   if (L.n <= 0) return 0;
   const bool n3 = e->pair_mode == 4;
   filler = filler && n3;
   PairParams P;
   const StencilRun *runs;
   int nblocks;
   tiled_params(e, n3, L.n, P, runs, nblocks);
   const int fill_wpb = filler ? e->ovl_threads / 32 : 0, fill_rows = filler ? e->ovl_blocks * fill_wpb : 0;
   // exponential potentials: the runs beyond r_far get their own launch with what is left of the potential there
   // (mdb_split_far_runs; nothing at all for a potential-only pass whose rest is zero)
   const bool far = n3 && e->far_ptype >= 0 && ptype == c.ptype;
   const bool far_launch = far && (coul || e->far_ptype != PT_NONE);
   const int far_row0 = fill_rows + nblocks * TW;
   const int nrows_pair = far_row0 + (far_launch ? nblocks * TW : 0);
   const int nblocks_u = (L.n + 255) / 256;
   const int nrows = nrows_pair + (n3 ? nblocks_u : 0);
   if (nrows + RS1 > e->partials_cap) {
      if (e->d_partials) cudaFree(e->d_partials);
      MDB_CUDA(cudaMalloc(&e->d_partials, sizeof(double) * NRED * (size_t)(nrows + RS1)));
      e->partials_cap = nrows + RS1;
   }
   if (n3) MDB_CUDA(cudaMemsetAsync(L.fs, 0, sizeof(double) * 3 * (size_t)L.n, st));
   const bool strict = c.strict_cutoff != 0 && !c.molpbc, fw = c.nsites_xf < c.nsites;   // src/force.c:951
   const RdfParams R0{0.0, 0, 0, nullptr};
   auto go = [&](int pt, dim3 g, PairQueue Q, int threads, const double *ptab_dev) -> int {
#define PT_CASE(X) case X: if (coul) launch_tiled<X, true, TM_FORCE>(strict, fw, n3, g, st, P, e, L, runs, d_out, R0, 0, Q, threads, ptab_dev); \
                           else launch_tiled<X, false, TM_FORCE>(strict, fw, n3, g, st, P, e, L, runs, d_out, R0, 0, Q, threads, ptab_dev); break
      switch (pt) {
         PT_CASE(PT_LJ);
#ifndef MDB_DEV_LJ_ONLY
         PT_CASE(PT_E6); PT_CASE(PT_MCY); PT_CASE(PT_GEN); PT_CASE(PT_HIW); PT_CASE(PT_MOR);
#endif
         case PT_NONE: launch_tiled<PT_NONE, true, TM_FORCE>(strict, fw, n3, g, st, P, e, L, runs, d_out, R0, 0, Q, threads, ptab_dev); break;
         default:
            mdb_set_error("KERNEL called with unknown potential type");
            return -1;
      }
#undef PT_CASE
      e->launches += 1;
      return 0;
   };
   if (far) { runs = e->d_runs_near; P.nruns = e->nruns_near; }
   if (filler) {
      if (go(ptype, dim3(e->ovl_blocks), PairQueue{1, e->d_ovl_q, e->d_ovl_q + 1, 0, 0}, e->ovl_threads, nullptr)) return -1;
      if (go(ptype, dim3(nblocks), PairQueue{2, e->d_ovl_q, e->d_ovl_q + 1, fill_rows, 0}, TW * 32, nullptr)) return -1;
   } else if (go(ptype, dim3(nblocks), PairQueue{0, nullptr, nullptr, 0, 0}, TW * 32, nullptr))
      return -1;
   if (far_launch) {
      runs = e->d_runs_far; P.nruns = e->nruns_far;
      if (go(e->far_ptype, dim3(nblocks), PairQueue{0, nullptr, nullptr, far_row0, 1}, TW * 32, e->d_ptab_far)) return -1;
   }
   if (n3) {
      k_unsort_virial<<<nblocks_u, 256, 0, st>>>(L.n, c.nsites, L.posq, L.order, L.fs, d_out,
                                                 e->d_partials + (size_t)nrows_pair * NRED);
      e->launches += 1;
   }
   double *rows1 = e->d_partials + (size_t)nrows * NRED;
   k_rows_stage1<<<RS1, 256, 0, st>>>(e->d_partials, nrows, rows1);
   k_rows_finish<<<1, RS1, 0, st>>>(rows1, c.nsites, n3 ? 1.0 : 0.5, d_out);
   e->launches += 2;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

int mdb_launch_pair_tiled(mdb_engine *e, double *d_out, cudaStream_t st)
{
   const mdb_config &c = e->cfg;
   const bool coul = c.alpha > 0.0;
   if (!e->pair_split) {
      if (mdb_need_batches(e, st)) return -1;
      if (e->pre_pair_wait) MDB_CUDA(cudaStreamWaitEvent(st, e->pre_pair_wait, 0));
      return pair_pass(e, full_list(e), c.ptype, coul, d_out, st, e->ovl_armed);
   }
   // split passes: charged x charged with the Coulomb term only, potential x potential with the potential only
   // the potential-site list is compacted on a second stream while the Coulomb pass runs
   bool wait_sub1 = false;
   if (!e->sub[1].valid) {
      if (!e->aux_stream) {
         MDB_CUDA(cudaStreamCreateWithFlags(&e->aux_stream, cudaStreamNonBlocking));
         MDB_CUDA(cudaEventCreateWithFlags(&e->ev_cells, cudaEventDisableTiming));
         MDB_CUDA(cudaEventCreateWithFlags(&e->ev_sub1, cudaEventDisableTiming));
      }
      MDB_CUDA(cudaEventRecord(e->ev_cells, st));
      MDB_CUDA(cudaStreamWaitEvent(e->aux_stream, e->ev_cells, 0));
      if (mdb_build_sublist(e, 1, e->aux_stream)) return -1;
      MDB_CUDA(cudaEventRecord(e->ev_sub1, e->aux_stream));
      wait_sub1 = true;
   }
   if (mdb_build_sublist(e, 0, st)) return -1;
   if (e->pre_pair_wait) MDB_CUDA(cudaStreamWaitEvent(st, e->pre_pair_wait, 0));
   if (pair_pass(e, sub_list(e, 0), PT_NONE, true, d_out, st, e->ovl_armed)) return -1;
   if (wait_sub1) MDB_CUDA(cudaStreamWaitEvent(st, e->ev_sub1, 0));
   if (pair_pass(e, sub_list(e, 1), c.ptype, false, d_out, st)) return -1;
   return mdb_launch_too_close_scan(e, st);
}

// ---- micro-benchmark: throughput of the pair arithmetic alone (no memory traffic) -----------------
template <int NV>
__global__ void __launch_bounds__(128) k_pairmath_probe(double *out, int iters, double alpha, double norm, double seed)
{
   double r2[NV], qq[NV], f[NV], ph[NV], acc = 0.0;
   __shared__ double tabp[8];
   if (threadIdx.x < 8) tabp[threadIdx.x] = 0.5 + 0.1 * threadIdx.x;
   __syncthreads();
   const double *pr[NV];
#pragma unroll
   for (int k = 0; k < NV; k++) { r2[k] = seed + 0.37 * k + 1e-3 * threadIdx.x; qq[k] = 0.3 + 0.01 * k; pr[k] = tabp; }
   for (int i = 0; i < iters; i++) {
      mdb_pair_eval_v<PT_LJ, true, NV>(r2, qq, pr, alpha, norm, f, ph);
#pragma unroll
      for (int k = 0; k < NV; k++) { acc += f[k] + ph[k]; r2[k] = fma(1e-9, f[k], r2[k]); }
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

extern "C" double mdb_pair_math_probe(int device, int iters, int nv, int blocks_per_sm)
{
   if (cudaSetDevice(device) != cudaSuccess) return -1.0;
   cudaDeviceProp prop;
   cudaGetDeviceProperties(&prop, device);
   const int blocks = prop.multiProcessorCount * blocks_per_sm, threads = 128;
   double *d = nullptr;
   if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   double best = 1e30;
   for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      if (nv == 1) k_pairmath_probe<1><<<blocks, threads>>>(d, iters, 0.12, 0.135, 9.0);
      else if (nv == 2) k_pairmath_probe<2><<<blocks, threads>>>(d, iters, 0.12, 0.135, 9.0);
      else if (nv == 4) k_pairmath_probe<4><<<blocks, threads>>>(d, iters, 0.12, 0.135, 9.0);
      else k_pairmath_probe<8><<<blocks, threads>>>(d, iters, 0.12, 0.135, 9.0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
   }
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   cudaFree(d);
   // pair evaluations per second
   return (double)iters * (nv == 1 ? 1 : nv == 2 ? 2 : nv == 4 ? 4 : 8) * blocks * threads / (best * 1e-3);
}
