// mdb_pair.cu -- FP64 real-space pair kernel over the link-cell structure.
//
// Replaces the loop nest of force_inner() (src/force.c:856-971): site_neighbour_list
// + gather + mk_r_sqr + kernel + mk_forces + scatter_forces, and the site virial
// (src/force.c:973-997).
//
// Formulation: one thread owns one site i of the cell-sorted
// array and walks the FULL neighbour-cell stencil F = H u (-H), where H is the
// reference's half list.  Cells are sorted z-fastest, so every z-run of stencil
// cells in a column (dx,dy) is ONE contiguous range of the sorted site array
// (split in at most three pieces by the periodic wrap in z).  Each reference
// pair (i,j,image) is therefore visited twice, once from each end; the visit
// adds the force on i only (no scatter, no atomics, deterministic), half the
// pair energy, and the image-translation part of the virial
//      stress = sum_i r_i (x) F_i + 1/2 sum_visits reloc_k (x) f_visit
// which equals the reference's sum_i r_i (x) F_i + sum_k reloc_k (x) gforce_k.
// The pair set is exactly the reference's: cells x stencil, no distance test,
// no minimum image, self-image pairs kept, only (j == i in the central image)
// skipped (SURVEY 8a' item 1).
#include "mdb_internal.h"
#include "mdb_math.cuh"

static constexpr int PB = 32;            // one warp per block: no block-level barrier, no tail warps
static constexpr int NRED = 8;           // pe + 6 stress + pad

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
   return v;
}

// Each lane walks its own flattened sequence of stencil segments (run x z-piece):
// when a lane exhausts a segment it advances on its own, so lanes with short and
// long segments do not wait for each other run by run; the warp only needs the
// total visit counts of its 32 sites to be similar (they are, to ~1 %).
// Two visits (j, j+1) are evaluated per iteration for instruction-level
// parallelism; the second is masked off at the end of a segment.
template <int PT, bool COUL, bool STRICT, bool FW>
__global__ void __launch_bounds__(PB)
k_pair(PairParams P, int nsites, const double4 *__restrict__ posq, const int *__restrict__ stype,
       const int *__restrict__ scell, const int *__restrict__ cstart, const int *__restrict__ order,
       const int *__restrict__ mol, const StencilRun *__restrict__ runs, const double *__restrict__ ptab,
       double *__restrict__ out, double *__restrict__ partials, unsigned long long *__restrict__ counters)
{
   const int s = P.s_lo + blockIdx.x * PB + threadIdx.x;
   const bool active = s < P.s_hi;
   double4 pi = make_double4(0, 0, 0, 0);
   int cx = 0, cy = 0, cz = 0, ti = 0, fwi = 0;
   if (active) {
      pi = posq[s];
      ti = stype[s];
      fwi = ti >> 30;                     // framework flag rides in bit 30 of the sorted type
      ti &= 0x3fffffff;
      int c = scell[s];
      cz = c % P.nz;
      int t = c / P.nz;
      cy = t % P.ny;
      cx = t / P.ny;
   }
   const double *__restrict__ prow = ptab + (size_t)ti * P.max_id * MDB_NPOTP;
   double fx = 0, fy = 0, fz = 0, pe = 0;
   double w00 = 0, w01 = 0, w02 = 0, w11 = 0, w12 = 0, w22 = 0;
   double gx = 0, gy = 0, gz = 0, sx = 0, sy = 0, sz = 0;
   unsigned int visits = 0;
   // segment iterator
   int r = 0, kk = 2, j = 0, je = 0, pend_j = 0, pend_je = 0, kimg = 13;
   int col = 0, z0 = 0, z1 = 0, ii = 0, jj = 0;
   bool done = !active;

   for (;;) {
      if (!done && j >= je) {
         // ---- close the finished segment: fold its force sum into F and the image virial
         fx += gx; fy += gy; fz += gz;
         if (kimg != 13) {
            const double rx = -0.5 * P.reloc[kimg][0], ry = -0.5 * P.reloc[kimg][1], rz = -0.5 * P.reloc[kimg][2];
            w00 = fma(rx, gx, w00); w01 = fma(ry, gx, w01); w02 = fma(rz, gx, w02);
            w11 = fma(ry, gy, w11); w12 = fma(rz, gy, w12); w22 = fma(rz, gz, w22);
         }
         gx = gy = gz = 0.0;
         if (pend_je > pend_j) {          // second half of the central segment (after the self site)
            j = pend_j; je = pend_je; pend_je = 0;
         } else {
            pend_je = 0;
            for (;;) {
               kk++;
               if (kk > 1 || kk * P.nz > z1) {      // open the next run
                  if (r >= P.nruns) { done = true; break; }
                  const StencilRun run = runs[r++];
                  int tx = cx + run.dx, ty = cy + run.dy;
                  ii = 0; jj = 0;
                  if (tx < 0) { tx += P.nx; ii = -1; } else if (tx >= P.nx) { tx -= P.nx; ii = 1; }
                  if (ty < 0) { ty += P.ny; jj = -1; } else if (ty >= P.ny) { ty -= P.ny; jj = 1; }
                  col = (tx * P.ny + ty) * P.nz;
                  z0 = cz + run.dzlo; z1 = cz + run.dzhi;
                  kk = z0 < 0 ? -1 : 0;
               }
               const int zoff = kk * P.nz;
               const int a = max(z0, zoff), b = min(z1, zoff + P.nz - 1);
               if (a > b) continue;
               const int jb = cstart[col + a - zoff], jn = cstart[col + b - zoff + 1];
               if (jb == jn) continue;
               kimg = 9 * (ii + 1) + 3 * (jj + 1) + (kk + 1);
               sx = pi.x - P.reloc[kimg][0]; sy = pi.y - P.reloc[kimg][1]; sz = pi.z - P.reloc[kimg][2];
               j = jb; je = jn;
               visits += (unsigned)(jn - jb);
               if (kimg == 13 && jb <= s && s < jn) {   // the reference cell itself: skip j == i
                  visits--;
                  je = s; pend_j = s + 1; pend_je = jn;
                  if (j >= je) { j = pend_j; je = pend_je; pend_je = 0; }
                  if (j >= je) continue;
               }
               break;
            }
         }
      }
      if (__all_sync(0xffffffffu, done)) break;
      if (!done) {
         const bool two = j + 1 < je;
         const int ja = j, jb2 = two ? j + 1 : j;
         const double4 pa = posq[ja], pb = posq[jb2];
         int ta = stype[ja], tb = stype[jb2];
         double ma = 1.0, mb = two ? 1.0 : 0.0;
         if (FW) {                         // framework sites never interact with each other (src/force.c:904-912)
            if (fwi & (ta >> 30)) { ma = 0.0; visits--; }
            if (two && (fwi & (tb >> 30))) { mb = 0.0; visits--; }
            ta &= 0x3fffffff; tb &= 0x3fffffff;
         }
         const double dxa = pa.x - sx, dya = pa.y - sy, dza = pa.z - sz;
         const double dxb = pb.x - sx, dyb = pb.y - sy, dzb = pb.z - sz;
         double r2a = fma(dxa, dxa, fma(dya, dya, dza * dza));
         double r2b = fma(dxb, dxb, fma(dyb, dyb, dzb * dzb));
         if (min(__double2hiint(r2a), __double2hiint(r2b)) < 0x3fd00000) {      // r^2 < 0.25 (rare)
            const int mi = mol[order[s]];
            if (r2a < MDB_TOO_CLOSE && ma != 0.0 && mi != mol[order[ja]]) {
               atomicAdd(&counters[1], 1ULL);
               counters[3] = ((unsigned long long)(unsigned)order[s] << 32) | (unsigned)order[ja];
            }
            if (two && r2b < MDB_TOO_CLOSE && mb != 0.0 && mi != mol[order[jb2]]) {
               atomicAdd(&counters[1], 1ULL);
               counters[3] = ((unsigned long long)(unsigned)order[s] << 32) | (unsigned)order[jb2];
            }
         }
         if (STRICT) {
            r2a = r2a > P.cutoffsq ? P.cutoff100sq : r2a;
            r2b = r2b > P.cutoffsq ? P.cutoff100sq : r2b;
         }
         const PairOut oa = mdb_pair_eval<PT, COUL>(r2a, pi.w * pa.w, prow + ta * MDB_NPOTP, P.alpha, P.norm);
         const PairOut ob = mdb_pair_eval<PT, COUL>(r2b, pi.w * pb.w, prow + tb * MDB_NPOTP, P.alpha, P.norm);
         const double fa = FW ? oa.fij * ma : oa.fij, fb = ob.fij * mb;
         pe += (FW ? oa.phi * ma : oa.phi) + ob.phi * mb;
         gx = fma(-fa, dxa, gx); gy = fma(-fa, dya, gy); gz = fma(-fa, dza, gz);
         gx = fma(-fb, dxb, gx); gy = fma(-fb, dyb, gy); gz = fma(-fb, dzb, gz);
         j += 2;
      }
   }

   if (active) {
      const int o = order[s];
      out[o] += fx;
      out[(size_t)nsites + o] += fy;
      out[2 * (size_t)nsites + o] += fz;
      w00 = fma(pi.x, fx, w00); w01 = fma(pi.y, fx, w01); w02 = fma(pi.z, fx, w02);
      w11 = fma(pi.y, fy, w11); w12 = fma(pi.z, fy, w12); w22 = fma(pi.z, fz, w22);
   }
   // warp reduction -> one row of partials per warp (deterministic second stage)
   double v[7] = {pe, w00, w01, w02, w11, w12, w22};
#pragma unroll
   for (int k = 0; k < 7; k++) {
      const double t = warp_sum(v[k]);
      if (threadIdx.x == 0) partials[(size_t)blockIdx.x * NRED + k] = t;
   }
   unsigned int vs = visits;
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, d);
   if (threadIdx.x == 0) atomicAdd(&counters[0], (unsigned long long)vs);
}

// Fixed-order sum of the per-block rows; adds 1/2 sum(phi) to pe_real and the six
// upper-triangle virial sums to stress (layout of include/moldy_b200.h).
__global__ void __launch_bounds__(256) k_pair_finish(const double *__restrict__ partials, int nblocks, int nsites,
                                                     double *__restrict__ out)
{
   __shared__ double sm[256];
   double acc[7] = {0, 0, 0, 0, 0, 0, 0};
   for (int b = threadIdx.x; b < nblocks; b += 256)
#pragma unroll
      for (int k = 0; k < 7; k++) acc[k] += partials[(size_t)b * NRED + k];
   double tot[7];
   for (int k = 0; k < 7; k++) {
      sm[threadIdx.x] = acc[k];
      __syncthreads();
      for (int d = 128; d > 0; d >>= 1) {
         if (threadIdx.x < d) sm[threadIdx.x] += sm[threadIdx.x + d];
         __syncthreads();
      }
      tot[k] = sm[0];
      __syncthreads();
   }
   if (threadIdx.x == 0) {
      double *sc = out + 3 * (size_t)nsites;
      sc[0] += 0.5 * tot[0];
      sc[2 + 0] += tot[1]; sc[2 + 1] += tot[2]; sc[2 + 2] += tot[3];
      sc[2 + 4] += tot[4]; sc[2 + 5] += tot[5]; sc[2 + 8] += tot[6];
   }
}

#define PAIR_ARGS P, e->cfg.nsites, e->d_posq, e->d_stype, e->d_scell, e->d_start, e->d_order, e->d_mol, e->d_runs, \
                  e->d_ptab, d_out, e->d_partials, e->d_counters
template <int PT, bool COUL>
static void launch_pair_t(bool strict, dim3 g, cudaStream_t st, PairParams &P, mdb_engine *e, double *d_out)
{
   const bool fw = e->cfg.nsites_xf < e->cfg.nsites;
   if (strict) {
      if (fw) k_pair<PT, COUL, true, true><<<g, PB, 0, st>>>(PAIR_ARGS);
      else k_pair<PT, COUL, true, false><<<g, PB, 0, st>>>(PAIR_ARGS);
   } else {
      if (fw) k_pair<PT, COUL, false, true><<<g, PB, 0, st>>>(PAIR_ARGS);
      else k_pair<PT, COUL, false, false><<<g, PB, 0, st>>>(PAIR_ARGS);
   }
}

template <int PT>
static void launch_pair_c(bool coul, bool strict, dim3 g, cudaStream_t st, PairParams &P, mdb_engine *e, double *d_out)
{
   if (coul) launch_pair_t<PT, true>(strict, g, st, P, e, d_out);
   else launch_pair_t<PT, false>(strict, g, st, P, e, d_out);
}

int mdb_launch_pair(mdb_engine *e, double *d_out, cudaStream_t st)
{
   const mdb_config &c = e->cfg;
   PairParams P;
   P.nx = e->T.nx; P.ny = e->T.ny; P.nz = e->T.nz; P.nruns = e->nruns;
   for (int k = 0; k < 27; k++)
      for (int a = 0; a < 3; a++) P.reloc[k][a] = e->T.reloc[k][a];
   P.alpha = c.alpha;
   P.norm = 2.0 * c.alpha / sqrt(MDB_PI);               // src/force.c:833
   P.cutoffsq = c.cutoff * c.cutoff;
   P.cutoff100sq = 10000.0 * P.cutoffsq;                // src/force.c:834-835
   P.max_id = c.max_id;
   P.strict = c.strict_cutoff;
   // contiguous slice of the cell-sorted sites for this rank (replicated-data SPMD)
   const long n = c.nsites;
   P.s_lo = (int)(n * e->ithread / e->nthreads);
   P.s_hi = (int)(n * (e->ithread + 1) / e->nthreads);
   const int cnt = P.s_hi - P.s_lo;
   if (cnt <= 0) return 0;
   const int nblocks = (cnt + PB - 1) / PB;
   if (nblocks > e->partials_cap) {
      if (e->d_partials) cudaFree(e->d_partials);
      MDB_CUDA(cudaMalloc(&e->d_partials, sizeof(double) * NRED * (size_t)nblocks));
      e->partials_cap = nblocks;
   }
   const bool coul = c.alpha > 0.0;                     // src/kernel.c:182
   const bool strict = c.strict_cutoff != 0 && !c.molpbc; // src/force.c:951
   dim3 g(nblocks);
   switch (c.ptype) {
      case PT_LJ:  launch_pair_c<PT_LJ>(coul, strict, g, st, P, e, d_out); break;
      case PT_E6:  launch_pair_c<PT_E6>(coul, strict, g, st, P, e, d_out); break;
      case PT_MCY: launch_pair_c<PT_MCY>(coul, strict, g, st, P, e, d_out); break;
      case PT_GEN: launch_pair_c<PT_GEN>(coul, strict, g, st, P, e, d_out); break;
      case PT_HIW: launch_pair_c<PT_HIW>(coul, strict, g, st, P, e, d_out); break;
      case PT_MOR: launch_pair_c<PT_MOR>(coul, strict, g, st, P, e, d_out); break;
      default:
         mdb_set_error("KERNEL called with unknown potential type");
         return -1;
   }
   k_pair_finish<<<1, 256, 0, st>>>(e->d_partials, nblocks, c.nsites, d_out);
   e->launches += 2;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

// ---- kernel() of the C ABI: the same device function over caller-supplied vectors
template <int PT, bool COUL>
__global__ void __launch_bounds__(256) k_kernel_vec(int jmin, int nnab, double *__restrict__ forceij,
                                                    double *__restrict__ pe_block, const double *__restrict__ r_sqr,
                                                    const double *__restrict__ nab_chg, double chg, double norm,
                                                    double alpha, const double *__restrict__ pot, int npar_rows)
{
   int j = jmin + blockIdx.x * 256 + threadIdx.x;
   double phi = 0.0;
   if (j < nnab) {
      __align__(16) double p[MDB_NPOTP];
#pragma unroll
      for (int k = 0; k < MDB_NPOTP; k++) p[k] = k < npar_rows ? pot[(size_t)k * nnab + j] : 0.0;
      if (PT == PT_LJ) { p[2] = 6.0 * p[0]; p[1] = p[1] * p[1]; }
      PairOut o = mdb_pair_eval<PT, COUL>(r_sqr[j], nab_chg[j] * chg, p, alpha, norm);
      forceij[j] = o.fij;
      phi = o.phi;
   }
   __shared__ double red[8];
   double t = warp_sum(phi);
   if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
   __syncthreads();
   if (threadIdx.x == 0) {
      double s = 0;
      for (int k = 0; k < 8; k++) s += red[k];
      pe_block[blockIdx.x] = s;
   }
}

template <int PT>
static void launch_vec(bool coul, int nb, int jmin, int nnab, double *f, double *peb, const double *r2,
                       const double *q, double chg, double norm, double alpha, const double *pot, int rows)
{
   if (coul) k_kernel_vec<PT, true><<<nb, 256>>>(jmin, nnab, f, peb, r2, q, chg, norm, alpha, pot, rows);
   else k_kernel_vec<PT, false><<<nb, 256>>>(jmin, nnab, f, peb, r2, q, chg, norm, alpha, pot, rows);
}

static int npar_of(int ptype)
{
   static const int np[] = {2, 3, 4, 6, 3, 1, 7};
   return (ptype >= 0 && ptype <= 6) ? np[ptype] : -1;
}

int mdb_launch_kernel_vec(int jmin, int nnab, double *forceij, double *pe, const double *r_sqr,
                          const double *nab_chg, double chg, double norm, double alpha, int ptype,
                          double *const *pot)
{
   const int cnt = nnab - jmin;
   if (cnt <= 0) return 0;
   const int rows = npar_of(ptype);
   if (rows < 0 || ptype == PT_RSV) {
      mdb_set_error("KERNEL called with unknown potential type");
      return -1;
   }
   const int nb = (cnt + 255) / 256;
   const size_t n = (size_t)nnab;
   struct Scratch {                             // freed on every return path
      double *p = nullptr;
      ~Scratch() { if (p) cudaFree(p); }
   } scratch;
   MDB_CUDA(cudaMalloc(&scratch.p, sizeof(double) * (n * (3 + rows) + nb)));
   double *d = scratch.p;
   double *d_f = d, *d_r2 = d + n, *d_q = d + 2 * n, *d_pot = d + 3 * n, *d_peb = d + (3 + rows) * n;
   MDB_CUDA(cudaMemcpy(d_r2, r_sqr, sizeof(double) * n, cudaMemcpyHostToDevice));
   MDB_CUDA(cudaMemcpy(d_q, nab_chg, sizeof(double) * n, cudaMemcpyHostToDevice));
   for (int k = 0; k < rows; k++)
      MDB_CUDA(cudaMemcpy(d_pot + k * n, pot[k], sizeof(double) * n, cudaMemcpyHostToDevice));
   const bool coul = alpha > 0.0;
   switch (ptype) {
      case PT_LJ:  launch_vec<PT_LJ>(coul, nb, jmin, nnab, d_f, d_peb, d_r2, d_q, chg, norm, alpha, d_pot, rows); break;
      case PT_E6:  launch_vec<PT_E6>(coul, nb, jmin, nnab, d_f, d_peb, d_r2, d_q, chg, norm, alpha, d_pot, rows); break;
      case PT_MCY: launch_vec<PT_MCY>(coul, nb, jmin, nnab, d_f, d_peb, d_r2, d_q, chg, norm, alpha, d_pot, rows); break;
      case PT_GEN: launch_vec<PT_GEN>(coul, nb, jmin, nnab, d_f, d_peb, d_r2, d_q, chg, norm, alpha, d_pot, rows); break;
      case PT_HIW: launch_vec<PT_HIW>(coul, nb, jmin, nnab, d_f, d_peb, d_r2, d_q, chg, norm, alpha, d_pot, rows); break;
      default:     launch_vec<PT_MOR>(coul, nb, jmin, nnab, d_f, d_peb, d_r2, d_q, chg, norm, alpha, d_pot, rows); break;
   }
   MDB_CUDA(cudaGetLastError());
   std::vector<double> peb(nb);
   MDB_CUDA(cudaMemcpy(forceij + jmin, d_f + jmin, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
   MDB_CUDA(cudaMemcpy(peb.data(), d_peb, sizeof(double) * nb, cudaMemcpyDeviceToHost));
   double s = 0;
   for (int k = 0; k < nb; k++) s += peb[k];
   *pe += s;
   return 0;
}
