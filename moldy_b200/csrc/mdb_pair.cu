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

static constexpr int PB = 128;           // threads per block
static constexpr int NRED = 8;           // pe + 6 stress + pad

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
   return v;
}

template <int PT, bool COUL, bool STRICT>
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
   unsigned int visits = 0;

   if (active)
      for (int r = 0; r < P.nruns; r++) {
         const StencilRun run = runs[r];
         int tx = cx + run.dx, ty = cy + run.dy, ii = 0, jj = 0;
         if (tx < 0) { tx += P.nx; ii = -1; } else if (tx >= P.nx) { tx -= P.nx; ii = 1; }
         if (ty < 0) { ty += P.ny; jj = -1; } else if (ty >= P.ny) { ty -= P.ny; jj = 1; }
         const int col = (tx * P.ny + ty) * P.nz;
         const int z0 = cz + run.dzlo, z1 = cz + run.dzhi;
#pragma unroll 1
         for (int kk = -1; kk <= 1; kk++) {
            const int zoff = kk * P.nz;
            const int a = max(z0, zoff), b = min(z1, zoff + P.nz - 1);
            if (a > b) continue;
            const int jb = cstart[col + a - zoff], je = cstart[col + b - zoff + 1];
            const int kimg = 9 * (ii + 1) + 3 * (jj + 1) + (kk + 1);
            const bool central = kimg == 13;
            const double sx = pi.x - P.reloc[kimg][0], sy = pi.y - P.reloc[kimg][1],
                         sz = pi.z - P.reloc[kimg][2];
            double gx = 0, gy = 0, gz = 0;
            visits += (unsigned)(je - jb);
#pragma unroll 2
            for (int j = jb; j < je; j++) {
               if (central && j == s) { visits--; continue; }
               const double4 pj = posq[j];
               int tj = stype[j];
               // framework sites never interact with each other (src/force.c:904-912)
               if (fwi & (tj >> 30)) { visits--; continue; }
               tj &= 0x3fffffff;
               const double dx = pj.x - sx, dy = pj.y - sy, dz = pj.z - sz;
               double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
               if (r2 < MDB_TOO_CLOSE) {
                  if (mol[order[s]] != mol[order[j]]) {
                     atomicAdd(&counters[1], 1ULL);
                     counters[3] = ((unsigned long long)(unsigned)order[s] << 32) | (unsigned)order[j];
                  }
               }
               if (STRICT) r2 = r2 > P.cutoffsq ? P.cutoff100sq : r2;
               const PairOut o = mdb_pair_eval<PT, COUL>(r2, pi.w * pj.w, prow + tj * MDB_NPOTP, P.alpha, P.norm);
               pe += o.phi;
               gx = fma(-o.fij, dx, gx);
               gy = fma(-o.fij, dy, gy);
               gz = fma(-o.fij, dz, gz);
            }
            fx += gx; fy += gy; fz += gz;
            if (!central) {
               const double rx = -0.5 * P.reloc[kimg][0], ry = -0.5 * P.reloc[kimg][1],
                            rz = -0.5 * P.reloc[kimg][2];
               w00 = fma(rx, gx, w00); w01 = fma(ry, gx, w01); w02 = fma(rz, gx, w02);
               w11 = fma(ry, gy, w11); w12 = fma(rz, gy, w12); w22 = fma(rz, gz, w22);
            }
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
   // block reduction -> one row of partials per block (deterministic second stage)
   __shared__ double red[PB / 32][NRED];
   __shared__ unsigned int vred[PB / 32];
   double v[7] = {pe, w00, w01, w02, w11, w12, w22};
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
   for (int k = 0; k < 7; k++) {
      double t = warp_sum(v[k]);
      if (lane == 0) red[w][k] = t;
   }
   unsigned int vs = visits;
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, d);
   if (lane == 0) vred[w] = vs;
   __syncthreads();
   if (threadIdx.x < 7) {
      double t = 0;
#pragma unroll
      for (int k = 0; k < PB / 32; k++) t += red[k][threadIdx.x];
      partials[(size_t)blockIdx.x * NRED + threadIdx.x] = t;
   }
   if (threadIdx.x == 0) {
      unsigned long long t = 0;
      for (int k = 0; k < PB / 32; k++) t += vred[k];
      atomicAdd(&counters[0], t);
   }
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

template <int PT, bool COUL>
static void launch_pair_t(bool strict, dim3 g, cudaStream_t st, PairParams &P, mdb_engine *e, double *d_out)
{
   if (strict)
      k_pair<PT, COUL, true><<<g, PB, 0, st>>>(P, e->cfg.nsites, e->d_posq, e->d_stype, e->d_scell, e->d_start,
                                               e->d_order, e->d_mol, e->d_runs, e->d_ptab, d_out, e->d_partials,
                                               e->d_counters);
   else
      k_pair<PT, COUL, false><<<g, PB, 0, st>>>(P, e->cfg.nsites, e->d_posq, e->d_stype, e->d_scell, e->d_start,
                                                e->d_order, e->d_mol, e->d_runs, e->d_ptab, d_out, e->d_partials,
                                                e->d_counters);
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
   const bool strict = c.strict_cutoff != 0;            // src/force.c:951 (molpbc unsupported)
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
      double p[MDB_NPOTP];
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
   double *d = nullptr;
   const size_t n = (size_t)nnab;
   MDB_CUDA(cudaMalloc(&d, sizeof(double) * (n * (3 + rows) + nb)));
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
   cudaFree(d);
   return 0;
}
