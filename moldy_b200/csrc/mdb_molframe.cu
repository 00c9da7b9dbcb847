// mdb_molframe.cu -- the molecular-frame steps of eval_forces() that surround force_calc()/ewald(), on the device
// (SURVEY 8f rank 1): building blocks, and below them the whole of eval_forces() with only centres of mass and
// quaternions in, molecular forces, torques and 23 scalars out (mdb_set_species / mdb_eval_forces_host):
//
//   mdb_make_sites   site co-ordinates of one species from scaled centres of mass, quaternions and principal-frame
//                    sites, written into the engine's own position rows (make_sites, src/algorith.c:169-217; rotate
//                    :76-97; q_to_rot src/quaterns.c:129-156; mat_vec_mul src/matrix.c:76-83)
//   mdb_mol_forces   molecular forces and torques of one species from the site forces of a result block
//                    (mol_force src/algorith.c:111-128, mol_torque :133-163)
//
// Site positions decide the cell assignment, which must be bit-identical to the reference's: every product and sum is
// an explicit round-to-nearest intrinsic in the reference's operation order, so nvcc cannot contract a*b+c into an FMA.
// HBM-bound: 56 B read per molecule + 24 B written per site, 24 B read per site + 48 B written per molecule.
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>
#include "mdb_internal.h"

static constexpr int MB = 256;

#define MUL(a, b) __dmul_rn(a, b)
#define ADD(a, b) __dadd_rn(a, b)
#define SUB(a, b) __dsub_rn(a, b)
// m0 a0 + m1 a1 + m2 a2, left to right (mat_vec_mul, MATMUL)
#define DOT3(m0, m1, m2, a0, a1, a2) ADD(ADD(MUL(m0, a0), MUL(m1, a1)), MUL(m2, a2))

struct Mat3 { double m[9]; };

__device__ __forceinline__ void q_to_rot(const double *__restrict__ q, double r[9])
{
   double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
   const double a01 = MUL(MUL(2.0, q0), q1), a02 = MUL(MUL(2.0, q0), q2), a03 = MUL(MUL(2.0, q0), q3);
   const double a12 = MUL(MUL(2.0, q1), q2), a13 = MUL(MUL(2.0, q1), q3), a23 = MUL(MUL(2.0, q2), q3);
   r[1] = SUB(a12, a03); r[2] = ADD(a13, a02);
   r[3] = ADD(a12, a03); r[5] = SUB(a23, a01);
   r[6] = SUB(a13, a02); r[7] = ADD(a23, a01);
   q0 = MUL(q0, q0); q1 = MUL(q1, q1); q2 = MUL(q2, q2); q3 = MUL(q3, q3);
   r[0] = SUB(SUB(ADD(q0, q1), q2), q3);
   r[4] = SUB(ADD(SUB(q0, q1), q2), q3);
   r[8] = ADD(SUB(SUB(q0, q1), q2), q3);
}

__global__ void __launch_bounds__(MB)
k_make_sites(Mat3 H, Mat3 HI, const double *__restrict__ com_s, const double *__restrict__ quat,
             const double *__restrict__ pfs, int nmols, int nsites, int sitepbc, double *__restrict__ x,
             double *__restrict__ y, double *__restrict__ z)
{
   const int k = blockIdx.x * MB + threadIdx.x;                 // site of the species, molecule-major
   if (k >= nmols * nsites) return;
   const int imol = k / nsites, is = k - imol * nsites;
   const double *h = H.m, *hi = HI.m;
   const double s0 = com_s[3 * imol], s1 = com_s[3 * imol + 1], s2 = com_s[3 * imol + 2];
   const double c0 = DOT3(h[0], h[1], h[2], s0, s1, s2), c1 = DOT3(h[3], h[4], h[5], s0, s1, s2),
                c2 = DOT3(h[6], h[7], h[8], s0, s1, s2);
   const double p0 = pfs[3 * is], p1 = pfs[3 * is + 1], p2 = pfs[3 * is + 2];
   double r0 = p0, r1 = p1, r2 = p2;
   if (quat) {
      double rot[9];
      q_to_rot(quat + 4 * (size_t)imol, rot);
      r0 = DOT3(rot[0], rot[1], rot[2], p0, p1, p2);
      r1 = DOT3(rot[3], rot[4], rot[5], p0, p1, p2);
      r2 = DOT3(rot[6], rot[7], rot[8], p0, p1, p2);
   }
   double sx = ADD(r0, c0), sy = ADD(r1, c1), sz = ADD(r2, c2);
   if (sitepbc) {
      const double tx = floor(ADD(DOT3(hi[0], hi[1], hi[2], sx, sy, sz), 0.5));
      const double ty = floor(ADD(DOT3(hi[3], hi[4], hi[5], sx, sy, sz), 0.5));
      const double tz = floor(ADD(DOT3(hi[6], hi[7], hi[8], sx, sy, sz), 0.5));
      const double dx = DOT3(h[0], h[1], h[2], tx, ty, tz), dy = DOT3(h[3], h[4], h[5], tx, ty, tz),
                   dz = DOT3(h[6], h[7], h[8], tx, ty, tz);
      sx = SUB(sx, dx); sy = SUB(sy, dy); sz = SUB(sz, dz);
   }
   x[k] = sx; y[k] = sy; z[k] = sz;
}

__global__ void __launch_bounds__(MB)
k_mol_forces(const double *__restrict__ fx, const double *__restrict__ fy, const double *__restrict__ fz,
             const double *__restrict__ quat, const double *__restrict__ pfs, int nmols, int nsites,
             double *__restrict__ force, double *__restrict__ torque)
{
   const int imol = blockIdx.x * MB + threadIdx.x;
   if (imol >= nmols) return;
   const size_t b = (size_t)imol * nsites;
   double f0 = 0.0, f1 = 0.0, f2 = 0.0;
   for (int is = 0; is < nsites; is++) { f0 = ADD(f0, fx[b + is]); f1 = ADD(f1, fy[b + is]); f2 = ADD(f2, fz[b + is]); }
   force[3 * (size_t)imol] = f0; force[3 * (size_t)imol + 1] = f1; force[3 * (size_t)imol + 2] = f2;
   if (!torque) return;
   double rot[9];
   q_to_rot(quat + 4 * (size_t)imol, rot);
   double t[3] = {0.0, 0.0, 0.0};
   for (int is = 0; is < nsites; is++) {
      const double a0 = fx[b + is], a1 = fy[b + is], a2 = fz[b + is];
      // principal-frame force = transposed rotation matrix x site force
      const double g0 = DOT3(rot[0], rot[3], rot[6], a0, a1, a2), g1 = DOT3(rot[1], rot[4], rot[7], a0, a1, a2),
                   g2 = DOT3(rot[2], rot[5], rot[8], a0, a1, a2);
      const double p0 = pfs[3 * is], p1 = pfs[3 * is + 1], p2 = pfs[3 * is + 2];
      t[0] = ADD(t[0], SUB(MUL(p1, g2), MUL(p2, g1)));
      t[1] = ADD(t[1], SUB(MUL(p2, g0), MUL(p0, g2)));
      t[2] = ADD(t[2], SUB(MUL(p0, g1), MUL(p1, g0)));
   }
   torque[3 * (size_t)imol] = t[0]; torque[3 * (size_t)imol + 1] = t[1]; torque[3 * (size_t)imol + 2] = t[2];
}

// det / invert in the reference's operation order (src/matrix.c:160-190); host code is built with -ffp-contract=off
static void invert_ref(const double a[9], double b[9])
{
   double d = 0.0;
   for (int i = 0, j = 1, k = 2; i < 3; i++, j = (j + 1) % 3, k = (k + 1) % 3)
      d += a[i] * (a[3 + j] * a[6 + k] - a[3 + k] * a[6 + j]);
   const double deter = 1.0 / d;
   for (int i = 0, j = 1, k = 2; i < 3; i++, j = (j + 1) % 3, k = (k + 1) % 3)
      for (int l = 0, m = 1, n = 2; l < 3; l++, m = (m + 1) % 3, n = (n + 1) % 3)
         b[3 * l + i] = deter * (a[3 * j + m] * a[3 * k + n] - a[3 * j + n] * a[3 * k + m]);
}

extern "C" int mdb_make_sites(mdb_engine *e, const double h[9], const double *d_com_s, const double *d_quat,
                              const double *d_pfs, int nmols, int nsites, int site_offset, int sitepbc, void *stream)
{
   if (!e->configured) { mdb_set_error("mdb_make_sites: engine not configured"); return -1; }
   const size_t n = e->cfg.nsites;
   if (site_offset < 0 || (size_t)site_offset + (size_t)nmols * nsites > n) { mdb_set_error("mdb_make_sites: site range"); return -1; }
   Mat3 H, HI;
   for (int i = 0; i < 9; i++) H.m[i] = h[i];
   invert_ref(H.m, HI.m);
   const int ns = nmols * nsites;
   if (ns > 0)
      k_make_sites<<<(ns + MB - 1) / MB, MB, 0, (cudaStream_t)stream>>>(H, HI, d_com_s, d_quat, d_pfs, nmols, nsites, sitepbc,
                                                                          e->own_xyz + site_offset, e->own_xyz + n + site_offset,
                                                                          e->own_xyz + 2 * n + site_offset);
   e->launches++;
   e->d_x = e->own_xyz; e->d_y = e->own_xyz + n; e->d_z = e->own_xyz + 2 * n;
   e->sites_set = true; e->cells_valid = false;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

extern "C" int mdb_mol_forces(mdb_engine *e, const double *d_out, const double *d_quat, const double *d_pfs, int nmols,
                              int nsites, int site_offset, double *d_force, double *d_torque, void *stream)
{
   const size_t n = e->cfg.nsites;
   if (site_offset < 0 || (size_t)site_offset + (size_t)nmols * nsites > n) { mdb_set_error("mdb_mol_forces: site range"); return -1; }
   if (d_torque && !d_quat) { mdb_set_error("mdb_mol_forces: torques need quaternions"); return -1; }
   if (nmols > 0)
      k_mol_forces<<<(nmols + MB - 1) / MB, MB, 0, (cudaStream_t)stream>>>(d_out + site_offset, d_out + n + site_offset,
                                                                            d_out + 2 * n + site_offset, d_quat, d_pfs, nmols,
                                                                            nsites, d_force, d_torque);
   e->launches++;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

extern "C" int mdb_get_sites(mdb_engine *e, double *hx, double *hy, double *hz, void *stream)
{
   if (!e->sites_set) { mdb_set_error("mdb_get_sites: no sites"); return -1; }
   const size_t n = e->cfg.nsites;
   cudaStream_t st = (cudaStream_t)stream;
   MDB_CUDA(cudaMemcpyAsync(hx, e->d_x, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaMemcpyAsync(hy, e->d_y, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaMemcpyAsync(hz, e->d_z, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaStreamSynchronize(st));
   return 0;
}

// =====================================================================================================================
// eval_forces() on the device (src/accel.c:398-617): species table, dipole moment, and the fused tail pass
// (surface-dipole term + mol_force + mol_torque + site->molecular virial).  HBM-bound, one pass over the site forces.
// =====================================================================================================================
static constexpr int DIP_BLOCKS = 592;            // 4 x 148 SMs, grid-stride

__device__ __forceinline__ double block_sum(double v, double *sh)
{
   for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
   const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
   __syncthreads();
   if (l == 0) sh[w] = v;
   __syncthreads();
   double s = 0.0;
   if (threadIdx.x == 0)
      for (int k = 0; k < MB / 32; k++) s += sh[k];
   return s;                                       // valid in thread 0
}

// dip[i] = sum_sites chg * site_i (src/accel.c:549-550), deterministic two-level sum
__global__ void __launch_bounds__(MB)
k_dipole_partial(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                 const double *__restrict__ chg, int n, double *__restrict__ part)
{
   __shared__ double sh[MB / 32];
   double a = 0.0, b = 0.0, c = 0.0;
   for (int i = blockIdx.x * MB + threadIdx.x; i < n; i += gridDim.x * MB) {
      const double q = chg[i];
      a = fma(q, x[i], a); b = fma(q, y[i], b); c = fma(q, z[i], c);
   }
   a = block_sum(a, sh); b = block_sum(b, sh); c = block_sum(c, sh);
   if (threadIdx.x == 0) { part[3 * blockIdx.x] = a; part[3 * blockIdx.x + 1] = b; part[3 * blockIdx.x + 2] = c; }
}

__global__ void __launch_bounds__(MB)
k_dipole_finish(const double *__restrict__ part, int nb, double *__restrict__ scal)
{
   __shared__ double sh[MB / 32];
   const int c = blockIdx.x;                       // one block per component
   double v = 0.0;
   for (int k = threadIdx.x; k < nb; k += MB) v += part[3 * k + c];
   v = block_sum(v, sh);
   if (threadIdx.x == 0) scal[c] = v;
}

// One thread per molecule of one species.  f' = f - (coef dip_i) chg (surface-dipole term, src/accel.c:553-555);
// F = sum f' (mol_force); torque in the principal frame (mol_torque); virial partials V[i][j] = sum f'_i d_j with
// d = R(quat) p (non-framework; = site - c.o.m. of the MOLPBC sites) or d = p (framework, src/accel.c:595-599).
__global__ void __launch_bounds__(MB)
k_mol_frame(const double *__restrict__ fx, const double *__restrict__ fy, const double *__restrict__ fz,
            const double *__restrict__ chg, const double *__restrict__ scal, double coef,
            const double *__restrict__ quat, const double *__restrict__ pfs, int nmols, int nsites, int framework,
            double *__restrict__ force, double *__restrict__ torque, double *__restrict__ vpart)
{
   __shared__ double sh[MB / 32];
   const int imol = blockIdx.x * MB + threadIdx.x;
   double v[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
   if (imol < nmols) {
      const size_t b = (size_t)imol * nsites;
      double k0 = 0.0, k1 = 0.0, k2 = 0.0;
      if (coef != 0.0) { k0 = MUL(coef, scal[0]); k1 = MUL(coef, scal[1]); k2 = MUL(coef, scal[2]); }
      double rot[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
      if (quat) q_to_rot(quat + 4 * (size_t)imol, rot);
      double f0 = 0.0, f1 = 0.0, f2 = 0.0, t0 = 0.0, t1 = 0.0, t2 = 0.0;
      for (int is = 0; is < nsites; is++) {
         double a0 = fx[b + is], a1 = fy[b + is], a2 = fz[b + is];
         if (coef != 0.0) {
            const double q = chg[b + is];
            a0 = SUB(a0, MUL(k0, q)); a1 = SUB(a1, MUL(k1, q)); a2 = SUB(a2, MUL(k2, q));
         }
         f0 = ADD(f0, a0); f1 = ADD(f1, a1); f2 = ADD(f2, a2);
         const double p0 = pfs[3 * is], p1 = pfs[3 * is + 1], p2 = pfs[3 * is + 2];
         double d0 = p0, d1 = p1, d2 = p2;
         if (quat && !framework) {
            d0 = DOT3(rot[0], rot[1], rot[2], p0, p1, p2);
            d1 = DOT3(rot[3], rot[4], rot[5], p0, p1, p2);
            d2 = DOT3(rot[6], rot[7], rot[8], p0, p1, p2);
         }
         v[0] = fma(a0, d0, v[0]); v[1] = fma(a0, d1, v[1]); v[2] = fma(a0, d2, v[2]);
         v[3] = fma(a1, d0, v[3]); v[4] = fma(a1, d1, v[4]); v[5] = fma(a1, d2, v[5]);
         v[6] = fma(a2, d0, v[6]); v[7] = fma(a2, d1, v[7]); v[8] = fma(a2, d2, v[8]);
         if (torque) {
            const double g0 = DOT3(rot[0], rot[3], rot[6], a0, a1, a2), g1 = DOT3(rot[1], rot[4], rot[7], a0, a1, a2),
                         g2 = DOT3(rot[2], rot[5], rot[8], a0, a1, a2);
            t0 = ADD(t0, SUB(MUL(p1, g2), MUL(p2, g1)));
            t1 = ADD(t1, SUB(MUL(p2, g0), MUL(p0, g2)));
            t2 = ADD(t2, SUB(MUL(p0, g1), MUL(p1, g0)));
         }
      }
      force[3 * (size_t)imol] = f0; force[3 * (size_t)imol + 1] = f1; force[3 * (size_t)imol + 2] = f2;
      if (torque) { torque[3 * (size_t)imol] = t0; torque[3 * (size_t)imol + 1] = t1; torque[3 * (size_t)imol + 2] = t2; }
   }
   for (int c = 0; c < 9; c++) {
      const double s = block_sum(v[c], sh);
      if (threadIdx.x == 0) vpart[9 * (size_t)blockIdx.x + c] = s;
   }
}

// scal[3..11] = sum of the virial partials (one block per component); scal[12..22] = pe_real, pe_recip, stress[9] of the
// result block
__global__ void __launch_bounds__(MB)
k_eval_finish(const double *__restrict__ vpart, int nb, const double *__restrict__ out_scal, double *__restrict__ scal)
{
   __shared__ double sh[MB / 32];
   const int c = blockIdx.x;
   double v = 0.0;
   for (int k = threadIdx.x; k < nb; k += MB) v += vpart[9 * (size_t)k + c];
   v = block_sum(v, sh);
   if (threadIdx.x == 0) scal[3 + c] = v;
   if (c == 0 && threadIdx.x < 11) scal[12 + threadIdx.x] = out_scal[threadIdx.x];
}

extern "C" int mdb_set_species(mdb_engine *e, int nspecies, const mdb_species *sp, const double *pfs)
{
   if (!e->configured) { mdb_set_error("mdb_set_species: engine not configured"); return -1; }
   auto &M = e->mf;
   M.sp.assign(sp, sp + nspecies);
   M.site_off.assign(nspecies, 0); M.mol_off.assign(nspecies, 0); M.quat_off.assign(nspecies, -1);
   M.torq_off.assign(nspecies, -1); M.pfs_off.assign(nspecies, 0); M.blk_off.assign(nspecies, 0);
   int so = 0, mo = 0, qo = 0, to = 0, po = 0, bo = 0;
   for (int i = 0; i < nspecies; i++) {
      if (sp[i].nmols < 0 || sp[i].nsites <= 0) { mdb_set_error("mdb_set_species: bad species"); return -1; }
      M.site_off[i] = so; M.mol_off[i] = mo; M.pfs_off[i] = po; M.blk_off[i] = bo;
      if (sp[i].rotates) { M.quat_off[i] = qo; qo += sp[i].nmols; }
      if (sp[i].rdof > 0) {
         if (!sp[i].rotates) { mdb_set_error("mdb_set_species: species with rotational freedom needs quaternions"); return -1; }
         M.torq_off[i] = to; to += sp[i].nmols;
      }
      so += sp[i].nmols * sp[i].nsites; mo += sp[i].nmols; po += sp[i].nsites; bo += (sp[i].nmols + MB - 1) / MB;
   }
   if (so != e->cfg.nsites) { mdb_set_error("mdb_set_species: species do not add up to nsites"); return -1; }
   M.nmols = mo; M.nmols_q = qo; M.nmols_r = to; M.npfs = po; M.nblocks = bo;
   if (M.d_pfs) cudaFree(M.d_pfs);
   if (M.d_vpart) cudaFree(M.d_vpart);
   M.d_pfs = M.d_vpart = nullptr;
   if (!M.d_dpart) MDB_CUDA(cudaMalloc(&M.d_dpart, sizeof(double) * 3 * DIP_BLOCKS));
   MDB_CUDA(cudaMalloc(&M.d_pfs, sizeof(double) * 3 * (size_t)std::max(po, 1)));
   MDB_CUDA(cudaMalloc(&M.d_vpart, sizeof(double) * 9 * (size_t)std::max(bo, 1)));
   MDB_CUDA(cudaMemcpy(M.d_pfs, pfs, sizeof(double) * 3 * (size_t)po, cudaMemcpyHostToDevice));
   const size_t need_in = 3 * (size_t)mo + 4 * (size_t)qo, need_res = 3 * (size_t)mo + 3 * (size_t)to + MDB_EVAL_SCALARS;
   if (need_in > M.in_cap) {
      if (M.d_in) cudaFree(M.d_in);
      if (M.h_in) cudaFreeHost(M.h_in);
      M.d_in = M.h_in = nullptr; M.in_cap = 0;
      MDB_CUDA(cudaMalloc(&M.d_in, sizeof(double) * need_in));
      MDB_CUDA(cudaMallocHost(&M.h_in, sizeof(double) * need_in));
      M.in_cap = need_in;
   }
   if (need_res > M.res_cap) {
      if (M.d_res) cudaFree(M.d_res);
      if (M.h_res) cudaFreeHost(M.h_res);
      M.d_res = M.h_res = nullptr; M.res_cap = 0;
      MDB_CUDA(cudaMalloc(&M.d_res, sizeof(double) * need_res));
      MDB_CUDA(cudaMemset(M.d_res, 0, sizeof(double) * need_res));    // forces/torques read back before the first evaluation are 0, not noise
      MDB_CUDA(cudaMallocHost(&M.h_res, sizeof(double) * need_res));
      M.res_cap = need_res;
   }
   const size_t need_out = mdb_out_doubles(e->cfg.nsites);
   if (need_out > e->out_cap) {
      if (e->d_out_own) cudaFree(e->d_out_own);
      e->d_out_own = nullptr; e->out_cap = 0;
      MDB_CUDA(cudaMalloc(&e->d_out_own, sizeof(double) * need_out));
      e->out_cap = need_out;
   }
   return 0;
}

extern "C" void mdb_eval_request_rdf(mdb_engine *e, double limit, int nbins, unsigned long long *h_counts)
{
   e->mf.rdf_limit = limit; e->mf.rdf_nbins = nbins; e->mf.rdf_counts = h_counts;
}

extern "C" size_t mdb_eval_result_doubles(const mdb_engine *e)
{
   return 3 * (size_t)e->mf.nmols + 3 * (size_t)e->mf.nmols_r + MDB_EVAL_SCALARS;
}

// ---- the pieces of eval_forces(), shared by the one-GPU path below and the multi-GPU driver (mdb_group.cu) ----------
// d_in = [scaled c-of-m 3 nmols | quaternions 4 nmols_q] on the device

int mdb_evalf_make_sites(mdb_engine *e, const double h[9], const double *d_in, bool second, cudaStream_t st)
{
   auto &M = e->mf;
   const double *d_com = d_in, *d_quat = d_in + 3 * (size_t)M.nmols;
   for (size_t i = 0; i < M.sp.size(); i++) {
      const mdb_species &s = M.sp[i];
      // first pass: control.molpbc ? MOLPBC : SITEPBC (src/accel.c:500-504); second: framework ? SITEPBC : MOLPBC (:537-542)
      const int sitepbc = second ? (s.framework ? 1 : 0) : (e->cfg.molpbc ? 0 : 1);
      if (mdb_make_sites(e, h, d_com + 3 * (size_t)M.mol_off[i], M.quat_off[i] >= 0 ? d_quat + 4 * (size_t)M.quat_off[i] : nullptr,
                         M.d_pfs + 3 * (size_t)M.pfs_off[i], s.nmols, s.nsites, M.site_off[i], sitepbc, st))
         return -1;
   }
   return 0;
}

// host staging of the caller's (pageable) per-species arrays into one pinned block, by up to four threads
int mdb_evalf_stage_inputs(mdb_engine *e, const double *const *com, const double *const *quat, double *h_in)
{
   auto &M = e->mf;
   struct Job { double *dst; const double *src; size_t bytes; };
   std::vector<Job> jobs;
   for (size_t i = 0; i < M.sp.size(); i++) {
      jobs.push_back({h_in + 3 * (size_t)M.mol_off[i], com[i], sizeof(double) * 3 * (size_t)M.sp[i].nmols});
      if (M.quat_off[i] >= 0) {
         if (!quat || !quat[i]) { mdb_set_error("mdb_eval_forces_host: quaternions missing"); return -1; }
         jobs.push_back({h_in + 3 * (size_t)M.nmols + 4 * (size_t)M.quat_off[i], quat[i],
                         sizeof(double) * 4 * (size_t)M.sp[i].nmols});
      }
   }
   size_t total = 0;
   for (auto &j : jobs) total += j.bytes;
   auto run = [&](int part, int nparts) {
      for (auto &j : jobs) {
         const size_t lo = j.bytes / 8 * part / nparts * 8, hi = j.bytes / 8 * (part + 1) / nparts * 8;
         memcpy((char *)j.dst + lo, (const char *)j.src + lo, hi - lo);
      }
   };
   if (total >= (4u << 20)) {
      std::thread th[3];
      for (int k = 1; k < 4; k++) th[k - 1] = std::thread(run, k, 4);
      run(0, 4);
      for (auto &t : th) t.join();
   } else {
      run(0, 1);
   }
   return 0;
}

// sites from the inputs (first make_sites) and, under molecular cut-off, the c-of-m the cell build bins by
int mdb_evalf_pre(mdb_engine *e, const double h[9], const double *d_in, cudaStream_t st)
{
   auto &M = e->mf;
   if (mdb_evalf_make_sites(e, h, d_in, false, st)) return -1;
   if (e->cfg.molpbc) {
      if (M.nmols > e->com_cap) {
         if (e->d_com) cudaFree(e->d_com);
         e->d_com = nullptr;
         MDB_CUDA(cudaMalloc(&e->d_com, sizeof(double) * 3 * (size_t)M.nmols));
         e->com_cap = M.nmols;
      }
      MDB_CUDA(cudaMemcpyAsync(e->d_com, d_in, sizeof(double) * 3 * (size_t)M.nmols, cudaMemcpyDeviceToDevice, st));
      e->com_set = true;
   }
   return 0;
}

// After the force sums: second make_sites, dipole moment, and for the molecules [m_lo, m_hi) (global molecule index) the
// surface-dipole term + mol_force + mol_torque + virial partials -> M.d_res (same layout as the one-GPU result; only the
// rows of those molecules are written) and the scalars.  d_fblock: result block [fx|fy|fz|16 scalars] holding the COMPLETE
// site forces of those molecules.
int mdb_evalf_tail(mdb_engine *e, const double h[9], const double *d_in, const double *d_fblock, int m_lo, int m_hi,
                   int surface_dipole, int do_recip, cudaStream_t st)
{
   auto &M = e->mf;
   const size_t n = e->cfg.nsites;
   double *scal = M.d_res + 3 * (size_t)M.nmols + 3 * (size_t)M.nmols_r;
   MDB_CUDA(cudaMemsetAsync(scal, 0, sizeof(double) * MDB_EVAL_SCALARS, st));
   double coef = 0.0;
   if (do_recip) {
      if (mdb_evalf_make_sites(e, h, d_in, true, st)) return -1;
      k_dipole_partial<<<DIP_BLOCKS, MB, 0, st>>>(e->d_x, e->d_y, e->d_z, e->d_chg, (int)n, M.d_dpart);
      k_dipole_finish<<<3, MB, 0, st>>>(M.d_dpart, DIP_BLOCKS, scal);
      e->launches += 2;
      if (surface_dipole) coef = 4.0 * MDB_PI / (3.0 * mdb_det3(h));
   }
   int nblk = 0;
   for (size_t i = 0; i < M.sp.size(); i++) {
      const mdb_species &s = M.sp[i];
      const int a = std::max(m_lo, M.mol_off[i]) - M.mol_off[i], b = std::min(m_hi, M.mol_off[i] + s.nmols) - M.mol_off[i];
      if (b <= a) continue;
      const int cnt = b - a;
      const size_t so = (size_t)M.site_off[i] + (size_t)a * s.nsites;
      k_mol_frame<<<(cnt + MB - 1) / MB, MB, 0, st>>>(
         d_fblock + so, d_fblock + n + so, d_fblock + 2 * n + so, e->d_chg + so, scal, coef,
         M.quat_off[i] >= 0 ? d_in + 3 * (size_t)M.nmols + 4 * ((size_t)M.quat_off[i] + a) : nullptr,
         M.d_pfs + 3 * (size_t)M.pfs_off[i], cnt, s.nsites, s.framework, M.d_res + 3 * ((size_t)M.mol_off[i] + a),
         M.torq_off[i] >= 0 ? M.d_res + 3 * (size_t)M.nmols + 3 * ((size_t)M.torq_off[i] + a) : nullptr,
         M.d_vpart + 9 * (size_t)nblk);
      nblk += (cnt + MB - 1) / MB;
      e->launches++;
   }
   k_eval_finish<<<9, MB, 0, st>>>(M.d_vpart, nblk, d_fblock + 3 * n, scal);
   e->launches++;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

extern "C" int mdb_eval_forces_host(mdb_engine *e, const double h[9], const double *const *com, const double *const *quat,
                                    int surface_dipole, int do_recip, double *h_result, void *stream)
{
   auto &M = e->mf;
   if (!e->configured || M.sp.empty()) { mdb_set_error("mdb_eval_forces_host: mdb_set_species was not called"); return -1; }
   cudaStream_t st = (cudaStream_t)stream;
   // centres of mass and quaternions: the only per-step input (56 B per molecule)
   if (mdb_evalf_stage_inputs(e, com, quat, M.h_in)) return -1;
   MDB_CUDA(cudaMemcpyAsync(M.d_in, M.h_in, sizeof(double) * (3 * (size_t)M.nmols + 4 * (size_t)M.nmols_q),
                            cudaMemcpyHostToDevice, st));
   if (mdb_evalf_pre(e, h, M.d_in, st)) return -1;
   double *d_out = e->d_out_own;
   if (mdb_zero_out(e, d_out, stream) || mdb_build_cells(e, stream) || mdb_force_real(e, d_out, stream)) return -1;
   if (do_recip && mdb_force_recip(e, d_out, stream)) return -1;
   if (M.rdf_counts) {                         // RDF pass of force_calc on this step's cell lists (synchronises)
      unsigned long long *dst = M.rdf_counts;
      M.rdf_counts = nullptr;
      if (mdb_rdf_counts(e, M.rdf_limit, M.rdf_nbins, dst, stream)) return -1;
   }
   if (mdb_evalf_tail(e, h, M.d_in, d_out, 0, M.nmols, surface_dipole, do_recip, st)) return -1;
   const size_t nres = mdb_eval_result_doubles(e);
   MDB_CUDA(cudaMemcpyAsync(M.h_res, M.d_res, sizeof(double) * nres, cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaStreamSynchronize(st));
   if (h_result) memcpy(h_result, M.h_res, sizeof(double) * nres);
   return 0;
}

extern "C" const double *mdb_eval_result(const mdb_engine *e) { return e->mf.h_res; }
