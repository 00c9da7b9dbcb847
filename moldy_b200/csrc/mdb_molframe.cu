// mdb_molframe.cu -- the molecular-frame steps of eval_forces() that surround force_calc()/ewald(), on the device
// (building blocks; the eval_forces() entry point itself is not built yet):
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
