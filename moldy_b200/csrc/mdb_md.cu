// mdb_md.cu -- the NVE leapfrog step of do_step() (src/accel.c:626-827) with the whole dynamic state resident in HBM
// (SURVEY 8f rank 4): scaled centres of mass, quaternions, linear and angular momenta never cross PCIe between outputs;
// a step returns ~100 scalars.
//
//   k_leap_coords    leapf_com (src/leapfrog.c:134-146, escape :117-129) + leapf_quat (:262-392; make_rot :205-219,
//                    make_rot_amom :224-241, rot_substep :247-253, normalise :96-113, q_mul/q_conj_mul src/quaterns.c:33-98)
//                    of one species: ONE fused pass per molecule
//   k_leap_momenta   leapf_mom (:150-158) + leapf_amom (:408-419)
//   k_md_sums        what values()/do_step reduce from the momenta and forces: sum p_i p_j with p = (h^-1)' mom (trans_ke,
//                    energy_dyad: src/algorith.c:221-284), sum amom_i^2 (rot_ke :244-257), sum F_i^2, sum T_i^2 (mean_square
//                    :101-106): per-block partials, folded in a fixed order by k_md_sums_finish
// Step (mdb_md_step): coords(step/2) -> eval_forces on the device (mdb_molframe.cu: make_sites .. mol_force/mol_torque) ->
// momenta(step/2) [-> sums at the half step, for H_0] -> momenta(step/2) -> framework momenta = 0 -> coords(step/2) -> sums.
//
// Parity: every product and sum of the translational part is an explicit round-to-nearest intrinsic in the reference's
// operation order (bit-identical to oracle/leapfrog.c, which is bit-identical to the reference); the rotational part goes
// through the device's sin/cos (<= 2 ulp against libm) and is compared to 1e-13; the sums are tree reductions (1e-13).
#include <float.h>
#include <math.h>
#include <algorithm>
#include <thread>
#include <vector>
#include "mdb_internal.h"

static constexpr int MB = 256;
#define MUL(a, b) __dmul_rn(a, b)
#define ADD(a, b) __dadd_rn(a, b)
#define SUB(a, b) __dsub_rn(a, b)
#define INERTIA_MIN 1.0e-14                     /* src/leapfrog.c:50 */

struct Mat3 { double m[9]; };
struct RotPar { double ri[3]; int saxis, symmetric; };

// y += a x, mvaxpy of src/matrix.c:211-225: y0 + a00 x0 + a01 x1 + a02 x2, left to right
__device__ __forceinline__ void mvaxpy1(const double *a, const double *x, double *y)
{
   const double y0 = ADD(ADD(ADD(y[0], MUL(a[0], x[0])), MUL(a[1], x[1])), MUL(a[2], x[2]));
   const double y1 = ADD(ADD(ADD(y[1], MUL(a[3], x[0])), MUL(a[4], x[1])), MUL(a[5], x[2]));
   const double y2 = ADD(ADD(ADD(y[2], MUL(a[6], x[0])), MUL(a[7], x[1])), MUL(a[8], x[2]));
   y[0] = y0; y[1] = y1; y[2] = y2;
}

// r = p q (conj: p -> p^-1), src/quaterns.c:33-98; r may alias p or q
__device__ __forceinline__ void qmul(const double *p, const double *q, double *r, bool conj_p)
{
   const double p0 = p[0], p1 = conj_p ? -p[1] : p[1], p2 = conj_p ? -p[2] : p[2], p3 = conj_p ? -p[3] : p[3];
   const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
   r[0] = SUB(SUB(SUB(MUL(p0, q0), MUL(p1, q1)), MUL(p2, q2)), MUL(p3, q3));
   r[1] = ADD(SUB(ADD(MUL(p1, q0), MUL(p0, q1)), MUL(p3, q2)), MUL(p2, q3));
   r[2] = SUB(ADD(ADD(MUL(p2, q0), MUL(p3, q1)), MUL(p0, q2)), MUL(p1, q3));
   r[3] = ADD(ADD(SUB(MUL(p3, q0), MUL(p2, q1)), MUL(p1, q2)), MUL(p0, q3));
}

__device__ __forceinline__ void axis_substep(double step, int axis, double rinertia, double *amom, double *quat)
{
   double rot[4] = {0.0, 0.0, 0.0, 0.0}, s, c;
   const double angle = MUL(MUL(MUL(0.5, step), rinertia), amom[axis + 1]);       /* make_rot */
   sincos(angle, &s, &c);
   rot[0] = c; rot[axis + 1] = s;
   qmul(rot, amom, amom, true);                                                   /* rot_substep */
   qmul(amom, rot, amom, false);
   qmul(quat, rot, quat, false);
}

__global__ void __launch_bounds__(MB)
k_leap_coords(Mat3 GI, RotPar R, double stepdts, int nmols, double *__restrict__ com, const double *__restrict__ mom,
              double *__restrict__ quat, double *__restrict__ amom, unsigned int *__restrict__ bad)
{
   const int m = blockIdx.x * MB + threadIdx.x;
   if (m >= nmols) return;
   {
      double c[3] = {com[3 * (size_t)m], com[3 * (size_t)m + 1], com[3 * (size_t)m + 2]};
      const double x[3] = {mom[3 * (size_t)m], mom[3 * (size_t)m + 1], mom[3 * (size_t)m + 2]};
      mvaxpy1(GI.m, x, c);
      for (int k = 0; k < 3; k++) c[k] = SUB(c[k], floor(ADD(c[k], 0.5)));        /* escape */
      com[3 * (size_t)m] = c[0]; com[3 * (size_t)m + 1] = c[1]; com[3 * (size_t)m + 2] = c[2];
   }
   if (!quat) return;
   double q[4], a[4];
   for (int k = 0; k < 4; k++) { q[k] = quat[4 * (size_t)m + k]; a[k] = amom[4 * (size_t)m + k]; }
   if (R.symmetric) {                                      /* leapf_quat_b */
      const int sx = R.saxis, o1 = (sx + 1) % 3, o2 = (sx + 2) % 3;
      axis_substep(MUL(0.5, stepdts), o1, SUB(R.ri[o1], R.ri[o2]), a, q);
      axis_substep(stepdts, sx, SUB(R.ri[sx], R.ri[o2]), a, q);
      {                                                    /* make_rot_amom + q_mul */
         double rot[4], sa, ca;
         const double samom = sqrt(ADD(ADD(MUL(a[1], a[1]), MUL(a[2], a[2])), MUL(a[3], a[3])));
         const double ramom = 1.0 / ADD(samom, 8 * DBL_MIN);
         const double angle = MUL(MUL(MUL(0.5, stepdts), R.ri[o2]), samom);
         sincos(angle, &sa, &ca);
         rot[0] = ca; rot[1] = MUL(MUL(sa, ramom), a[1]); rot[2] = MUL(MUL(sa, ramom), a[2]); rot[3] = MUL(MUL(sa, ramom), a[3]);
         qmul(q, rot, q, false);
      }
      axis_substep(MUL(0.5, stepdts), o1, SUB(R.ri[o1], R.ri[o2]), a, q);
   } else {                                                /* leapf_quat_a */
      axis_substep(MUL(0.5, stepdts), 0, R.ri[0], a, q);
      axis_substep(MUL(0.5, stepdts), 1, R.ri[1], a, q);
      axis_substep(stepdts, 2, R.ri[2], a, q);
      axis_substep(MUL(0.5, stepdts), 1, R.ri[1], a, q);
      axis_substep(MUL(0.5, stepdts), 0, R.ri[0], a, q);
   }
   double norm = 0.0;                                      /* normalise */
   for (int j = 0; j < 4; j++) norm = ADD(norm, MUL(q[j], q[j]));
   norm = sqrt(norm);
   if (fabs(norm - 1.0) > 1.0e-4) atomicAdd(bad, 1u);
   for (int j = 0; j < 4; j++) { quat[4 * (size_t)m + j] = q[j] / norm; amom[4 * (size_t)m + j] = a[j]; }
}

__global__ void __launch_bounds__(MB)
k_leap_momenta(Mat3 HT, double step, int nmols, double *__restrict__ mom, const double *__restrict__ force,
               double *__restrict__ amom, const double *__restrict__ torque)
{
   const int m = blockIdx.x * MB + threadIdx.x;
   if (m >= nmols) return;
   double y[3] = {mom[3 * (size_t)m], mom[3 * (size_t)m + 1], mom[3 * (size_t)m + 2]};
   const double x[3] = {force[3 * (size_t)m], force[3 * (size_t)m + 1], force[3 * (size_t)m + 2]};
   mvaxpy1(HT.m, x, y);
   mom[3 * (size_t)m] = y[0]; mom[3 * (size_t)m + 1] = y[1]; mom[3 * (size_t)m + 2] = y[2];
   if (amom)
      for (int k = 0; k < 3; k++)
         amom[4 * (size_t)m + 1 + k] = ADD(amom[4 * (size_t)m + 1 + k], MUL(step, torque[3 * (size_t)m + k]));
}

// per block: [pxpx pxpy pxpz pypy pypz pzpz | a1^2 a2^2 a3^2 | Fx^2 Fy^2 Fz^2 | Tx^2 Ty^2 Tz^2]
static constexpr int NSUM = MDB_MD_SUMS;
__global__ void __launch_bounds__(MB)
k_md_sums(Mat3 HI, int nmols, const double *__restrict__ mom, const double *__restrict__ amom, const double *__restrict__ force,
          const double *__restrict__ torque, double *__restrict__ part)
{
   __shared__ double sh[MB / 32][NSUM];
   const int m = blockIdx.x * MB + threadIdx.x;
   double v[NSUM];
   for (int k = 0; k < NSUM; k++) v[k] = 0.0;
   if (m < nmols) {
      const double x0 = mom[3 * (size_t)m], x1 = mom[3 * (size_t)m + 1], x2 = mom[3 * (size_t)m + 2];
      const double *hi = HI.m;             /* p = transposed inverse x mom, mat_vec_mul order */
      const double p0 = ADD(ADD(MUL(hi[0], x0), MUL(hi[3], x1)), MUL(hi[6], x2));
      const double p1 = ADD(ADD(MUL(hi[1], x0), MUL(hi[4], x1)), MUL(hi[7], x2));
      const double p2 = ADD(ADD(MUL(hi[2], x0), MUL(hi[5], x1)), MUL(hi[8], x2));
      v[0] = p0 * p0; v[1] = p0 * p1; v[2] = p0 * p2; v[3] = p1 * p1; v[4] = p1 * p2; v[5] = p2 * p2;
      if (amom) for (int k = 0; k < 3; k++) { const double a = amom[4 * (size_t)m + 1 + k]; v[6 + k] = a * a; }
      if (force) for (int k = 0; k < 3; k++) { const double f = force[3 * (size_t)m + k]; v[9 + k] = f * f; }
      if (torque) for (int k = 0; k < 3; k++) { const double t = torque[3 * (size_t)m + k]; v[12 + k] = t * t; }
   }
   const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
   for (int k = 0; k < NSUM; k++) {
      double t = v[k];
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if (l == 0) sh[w][k] = t;
   }
   __syncthreads();
   if (threadIdx.x < NSUM) {
      double t = 0.0;
      for (int k = 0; k < MB / 32; k++) t += sh[k][threadIdx.x];
      part[(size_t)blockIdx.x * NSUM + threadIdx.x] = t;
   }
}

__global__ void __launch_bounds__(MB) k_md_sums_finish(const double *__restrict__ part, int nb, double *__restrict__ out)
{
   __shared__ double sh[MB / 32];
   const int c = blockIdx.x;
   double v = 0.0;
   for (int k = threadIdx.x; k < nb; k += MB) v += part[(size_t)k * NSUM + c];
   for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
   if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
   __syncthreads();
   if (threadIdx.x == 0) {
      double s = 0.0;
      for (int k = 0; k < MB / 32; k++) s += sh[k];
      out[c] = s;
   }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
static void mat_mul3(const double a[9], const double b[9], double c[9])          /* src/matrix.c mat_mul */
{
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
static void invert3(const double a[9], double b[9])                               /* src/matrix.c:160-190 */
{
   double d = 0.0;
   for (int i = 0, j = 1, k = 2; i < 3; i++, j = (j + 1) % 3, k = (k + 1) % 3)
      d += a[i] * (a[3 + j] * a[6 + k] - a[3 + k] * a[6 + j]);
   const double deter = 1.0 / d;
   for (int i = 0, j = 1, k = 2; i < 3; i++, j = (j + 1) % 3, k = (k + 1) % 3)
      for (int l = 0, m = 1, n = 2; l < 3; l++, m = (m + 1) % 3, n = (n + 1) % 3)
         b[3 * l + i] = deter * (a[3 * j + m] * a[3 * k + n] - a[3 * j + n] * a[3 * k + m]);
}
static void rinertia(const double *inertia, double r[3])                          /* src/leapfrog.c:274-280 */
{
   for (int i = 0; i < 3; i++)
      r[i] = inertia[i] / (inertia[(i + 1) % 3] + inertia[(i + 2) % 3]) < INERTIA_MIN ? 0.0 : 1.0 / inertia[i];
}

extern "C" int mdb_md_set_dynamics(mdb_engine *e, const mdb_species_dyn *dyn, int nosymmetric_rot)
{
   auto &M = e->mf;
   if (M.sp.empty()) { mdb_set_error("mdb_md_set_dynamics: mdb_set_species was not called"); return -1; }
   MDB_CUDA(cudaSetDevice(e->device));
   M.dyn.assign(dyn, dyn + M.sp.size());
   M.nosymmetric_rot = nosymmetric_rot;
   // the near-symmetry axis leapf_quat_b picks from the FIRST species it is called for and keeps (function static, :272,286-301)
   M.saxis = 0;
   for (size_t i = 0; i < M.sp.size(); i++)
      if (M.sp[i].rdof > 0) {
         double r[3], idmin = DBL_MAX;
         rinertia(dyn[i].inertia, r);
         for (int k = 0; k < 3; k++) {
            const double idiff = fabs(r[(k + 1) % 3] - r[(k + 2) % 3]);
            if (idiff < idmin) { idmin = idiff; M.saxis = k; }
         }
         break;
      }
   const size_t nm = (size_t)std::max(M.nmols, 1), nq = (size_t)std::max(M.nmols_q, 1);
   if (M.d_mom) cudaFree(M.d_mom);
   if (M.d_amom) cudaFree(M.d_amom);
   if (M.d_mdpart) cudaFree(M.d_mdpart);
   if (M.d_mdscal) cudaFree(M.d_mdscal);
   if (M.h_mdscal) cudaFreeHost(M.h_mdscal);
   M.d_mom = M.d_amom = M.d_mdpart = M.d_mdscal = M.h_mdscal = nullptr;
   MDB_CUDA(cudaMalloc(&M.d_mom, sizeof(double) * 3 * nm));
   MDB_CUDA(cudaMalloc(&M.d_amom, sizeof(double) * 4 * nq));
   MDB_CUDA(cudaMalloc(&M.d_mdpart, sizeof(double) * NSUM * (size_t)std::max(M.nblocks, 1)));
   const size_t nscal = mdb_md_scalars(e);
   MDB_CUDA(cudaMalloc(&M.d_mdscal, sizeof(double) * nscal));
   MDB_CUDA(cudaMallocHost(&M.h_mdscal, sizeof(double) * nscal));
   MDB_CUDA(cudaMemset(M.d_mdscal, 0, sizeof(double) * nscal));
   return 0;
}

// [MDB_EVAL_SCALARS of eval_forces | per species MDB_MD_SUMS at the end of the step | per species MDB_MD_SUMS at the half
//  step (only when asked for) | bad quaternion count]
extern "C" size_t mdb_md_scalars(const mdb_engine *e) { return MDB_EVAL_SCALARS + 2 * NSUM * e->mf.sp.size() + 1; }

// the dynamic state, per species HOST arrays as Moldy holds them: c_of_m[nmols][3] (scaled), quat[nmols][4], mom[nmols][3],
// amom[nmols][4]; quat/amom NULL for species without quaternions
// Host <-> device transfer of the per-species state arrays.  The caller's arrays are pageable (Moldy's own malloc'ed
// c_of_m/quat/mom/amom): cudaMemcpy from them runs at 2-3 GB/s, which cost the do_step() of the C ABI 24 ms per step at
// 256 000 molecules.  They are staged through ONE pinned block instead: up to six host threads copy the pieces, the device
// copies then run at PCIe speed.  Layout of the block: [c-of-m 3 nmols | quaternions 4 nmols_q | mom 3 nmols | amom 4 nmols_q
// | force 3 nmols | torque 3 nmols_r].
typedef MdbCopyJob CopyJob;
void mdb_run_copy_jobs(const std::vector<MdbCopyJob> &jobs)
{
   size_t total = 0;
   for (auto &j : jobs) total += j.bytes;
   auto run = [&](int part, int nparts) {
      for (auto &j : jobs) {
         const size_t lo = j.bytes / 8 * part / nparts * 8, hi = j.bytes / 8 * (part + 1) / nparts * 8;
         if (j.src) memcpy((char *)j.dst + lo, (const char *)j.src + lo, hi - lo);
         else memset((char *)j.dst + lo, 0, hi - lo);
      }
   };
   if (total >= (4u << 20)) {
      std::thread th[5];
      for (int k = 1; k < 6; k++) th[k - 1] = std::thread(run, k, 6);
      run(0, 6);
      for (auto &t : th) t.join();
   } else {
      run(0, 1);
   }
}
static int state_block(mdb_engine *e, size_t off[6])
{
   auto &M = e->mf;
   const size_t nm = (size_t)M.nmols, nq = (size_t)M.nmols_q, nr = (size_t)M.nmols_r;
   off[0] = 0; off[1] = 3 * nm; off[2] = off[1] + 4 * nq; off[3] = off[2] + 3 * nm; off[4] = off[3] + 4 * nq; off[5] = off[4] + 3 * nm;
   const size_t need = off[5] + 3 * nr + 8;
   if (need > M.state_cap) {
      if (M.h_state) cudaFreeHost(M.h_state);
      M.h_state = nullptr; M.state_cap = 0;
      MDB_CUDA(cudaMallocHost(&M.h_state, sizeof(double) * need));
      M.state_cap = need;
   }
   return 0;
}

// the dynamic state, per species HOST arrays as Moldy holds them: c_of_m[nmols][3] (scaled), quat[nmols][4], mom[nmols][3],
// amom[nmols][4]; quat/amom NULL for species without quaternions.  Returns when the state is on the device.
extern "C" int mdb_md_upload_state(mdb_engine *e, const double *const *com, const double *const *quat, const double *const *mom,
                                   const double *const *amom, void *stream)
{
   auto &M = e->mf;
   cudaStream_t st = (cudaStream_t)stream;
   if (!M.d_mom) { mdb_set_error("mdb_md_upload_state: mdb_md_set_dynamics was not called"); return -1; }
   MDB_CUDA(cudaSetDevice(e->device));
   size_t off[6];
   if (state_block(e, off)) return -1;
   double *h = M.h_state;
   std::vector<CopyJob> jobs;
   for (size_t i = 0; i < M.sp.size(); i++) {
      const size_t nm = (size_t)M.sp[i].nmols;
      if (nm == 0) continue;
      jobs.push_back({h + off[0] + 3 * (size_t)M.mol_off[i], com[i], sizeof(double) * 3 * nm});
      jobs.push_back({h + off[2] + 3 * (size_t)M.mol_off[i], mom[i], sizeof(double) * 3 * nm});
      if (M.quat_off[i] >= 0) {
         if (!quat || !quat[i]) { mdb_set_error("mdb_md_upload_state: quaternions missing"); return -1; }
         jobs.push_back({h + off[1] + 4 * (size_t)M.quat_off[i], quat[i], sizeof(double) * 4 * nm});
         jobs.push_back({h + off[3] + 4 * (size_t)M.quat_off[i], amom && amom[i] ? amom[i] : nullptr, sizeof(double) * 4 * nm});
      }
   }
   mdb_run_copy_jobs(jobs);
   const size_t nm = (size_t)M.nmols, nq = (size_t)M.nmols_q;
   MDB_CUDA(cudaMemcpyAsync(M.d_in, h + off[0], sizeof(double) * (3 * nm + 4 * nq), cudaMemcpyHostToDevice, st));
   MDB_CUDA(cudaMemcpyAsync(M.d_mom, h + off[2], sizeof(double) * 3 * nm, cudaMemcpyHostToDevice, st));
   if (nq) MDB_CUDA(cudaMemcpyAsync(M.d_amom, h + off[3], sizeof(double) * 4 * nq, cudaMemcpyHostToDevice, st));
   MDB_CUDA(cudaStreamSynchronize(st));                   // the staging block is free again
   return 0;
}

extern "C" int mdb_md_download_state(mdb_engine *e, double *const *com, double *const *quat, double *const *mom, double *const *amom,
                                     double *const *force, double *const *torque, void *stream)
{
   auto &M = e->mf;
   cudaStream_t st = (cudaStream_t)stream;
   MDB_CUDA(cudaSetDevice(e->device));
   size_t off[6];
   if (state_block(e, off)) return -1;
   double *h = M.h_state;
   const size_t nm = (size_t)M.nmols, nq = (size_t)M.nmols_q, nr = (size_t)M.nmols_r;
   bool want_f = false, want_t = false;
   for (size_t i = 0; i < M.sp.size(); i++) { want_f |= force && force[i]; want_t |= torque && torque[i]; }
   MDB_CUDA(cudaMemcpyAsync(h + off[0], M.d_in, sizeof(double) * (3 * nm + 4 * nq), cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaMemcpyAsync(h + off[2], M.d_mom, sizeof(double) * 3 * nm, cudaMemcpyDeviceToHost, st));
   if (nq) MDB_CUDA(cudaMemcpyAsync(h + off[3], M.d_amom, sizeof(double) * 4 * nq, cudaMemcpyDeviceToHost, st));
   if (want_f) MDB_CUDA(cudaMemcpyAsync(h + off[4], M.d_res, sizeof(double) * 3 * nm, cudaMemcpyDeviceToHost, st));
   if (want_t && nr) MDB_CUDA(cudaMemcpyAsync(h + off[5], M.d_res + 3 * nm, sizeof(double) * 3 * nr, cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaStreamSynchronize(st));
   std::vector<CopyJob> jobs;
   for (size_t i = 0; i < M.sp.size(); i++) {
      const size_t n = (size_t)M.sp[i].nmols;
      if (n == 0) continue;
      if (com && com[i]) jobs.push_back({com[i], h + off[0] + 3 * (size_t)M.mol_off[i], sizeof(double) * 3 * n});
      if (mom && mom[i]) jobs.push_back({mom[i], h + off[2] + 3 * (size_t)M.mol_off[i], sizeof(double) * 3 * n});
      if (force && force[i]) jobs.push_back({force[i], h + off[4] + 3 * (size_t)M.mol_off[i], sizeof(double) * 3 * n});
      if (M.quat_off[i] >= 0) {
         if (quat && quat[i]) jobs.push_back({quat[i], h + off[1] + 4 * (size_t)M.quat_off[i], sizeof(double) * 4 * n});
         if (amom && amom[i]) jobs.push_back({amom[i], h + off[3] + 4 * (size_t)M.quat_off[i], sizeof(double) * 4 * n});
      }
      if (M.torq_off[i] >= 0 && torque && torque[i])
         jobs.push_back({torque[i], h + off[5] + 3 * (size_t)M.torq_off[i], sizeof(double) * 3 * n});
   }
   mdb_run_copy_jobs(jobs);
   return 0;
}

// leapf_all_coords(step) of src/accel.c:360-372 on the resident state; the _range forms work on the molecules [m_lo, m_hi)
// of the state block d_in (a rank's share in a device group, mdb_group.cu)
int mdb_md_coords_range(mdb_engine *e, const double h[9], double step, double ts, double *d_in, int m_lo, int m_hi, cudaStream_t st)
{
   auto &M = e->mf;
   double ht[9], G[9], Gi[9];
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) ht[3 * j + i] = h[3 * i + j];
   mat_mul3(ht, h, G);
   invert3(G, Gi);
   unsigned int *bad = reinterpret_cast<unsigned int *>(M.d_mdscal + mdb_md_scalars(e) - 1);
   for (size_t i = 0; i < M.sp.size(); i++) {
      const mdb_species &s = M.sp[i];
      const int a = std::max(m_lo, M.mol_off[i]) - M.mol_off[i], b = std::min(m_hi, M.mol_off[i] + s.nmols) - M.mol_off[i];
      if (b <= a) continue;
      Mat3 GI;
      const double f = step / (M.dyn[i].mass * ts);                 /* mat_sca_mul(step/(mass*s), Ginv, Ginv), :142 */
      for (int k = 0; k < 9; k++) GI.m[k] = f * Gi[k];
      RotPar R;
      rinertia(M.dyn[i].inertia, R.ri);
      R.saxis = M.saxis; R.symmetric = M.nosymmetric_rot ? 0 : 1;
      const bool rot = s.rdof > 0 && M.quat_off[i] >= 0;
      k_leap_coords<<<(b - a + MB - 1) / MB, MB, 0, st>>>(
         GI, R, step / ts, b - a, d_in + 3 * ((size_t)M.mol_off[i] + a), M.d_mom + 3 * ((size_t)M.mol_off[i] + a),
         rot ? d_in + 3 * (size_t)M.nmols + 4 * ((size_t)M.quat_off[i] + a) : nullptr,
         rot ? M.d_amom + 4 * ((size_t)M.quat_off[i] + a) : nullptr, bad);
      e->launches++;
   }
   MDB_CUDA(cudaGetLastError());
   return 0;
}
extern "C" int mdb_md_coords(mdb_engine *e, const double h[9], double step, double ts, void *stream)
{
   return mdb_md_coords_range(e, h, step, ts, e->mf.d_in, 0, e->mf.nmols, (cudaStream_t)stream);
}

// leapf_all_momenta(step) of src/accel.c:376-388 (the caller passes step*ts) from the forces/torques of the last eval
int mdb_md_momenta_range(mdb_engine *e, const double h[9], double step, int m_lo, int m_hi, cudaStream_t st)
{
   auto &M = e->mf;
   Mat3 HT;
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) HT.m[3 * j + i] = step * h[3 * i + j];         /* transpose + mat_sca_mul, :154-155 */
   for (size_t i = 0; i < M.sp.size(); i++) {
      const mdb_species &s = M.sp[i];
      const int a = std::max(m_lo, M.mol_off[i]) - M.mol_off[i], b = std::min(m_hi, M.mol_off[i] + s.nmols) - M.mol_off[i];
      if (b <= a) continue;
      const bool rot = s.rdof > 0 && M.quat_off[i] >= 0 && M.torq_off[i] >= 0;
      k_leap_momenta<<<(b - a + MB - 1) / MB, MB, 0, st>>>(
         HT, step, b - a, M.d_mom + 3 * ((size_t)M.mol_off[i] + a), M.d_res + 3 * ((size_t)M.mol_off[i] + a),
         rot ? M.d_amom + 4 * ((size_t)M.quat_off[i] + a) : nullptr,
         rot ? M.d_res + 3 * (size_t)M.nmols + 3 * ((size_t)M.torq_off[i] + a) : nullptr);
      e->launches++;
   }
   MDB_CUDA(cudaGetLastError());
   return 0;
}
extern "C" int mdb_md_momenta(mdb_engine *e, const double h[9], double step, void *stream)
{
   return mdb_md_momenta_range(e, h, step, 0, e->mf.nmols, (cudaStream_t)stream);
}

// sums of the momenta (and, with_forces, of the molecular forces/torques) per species -> d_mdscal[slot]
int mdb_md_sums_range(mdb_engine *e, const double h[9], int slot, bool with_forces, int m_lo, int m_hi, cudaStream_t st)
{
   auto &M = e->mf;
   Mat3 HI;
   invert3(h, HI.m);
   for (size_t i = 0; i < M.sp.size(); i++) {
      const mdb_species &s = M.sp[i];
      double *dst = M.d_mdscal + MDB_EVAL_SCALARS + NSUM * ((size_t)slot * M.sp.size() + i);
      const int a = std::max(m_lo, M.mol_off[i]) - M.mol_off[i], b = std::min(m_hi, M.mol_off[i] + s.nmols) - M.mol_off[i];
      if (b <= a) { MDB_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * NSUM, st)); continue; }
      const int nb = (b - a + MB - 1) / MB;
      const bool rot = s.rdof > 0 && M.quat_off[i] >= 0;
      k_md_sums<<<nb, MB, 0, st>>>(HI, b - a, M.d_mom + 3 * ((size_t)M.mol_off[i] + a),
                                   rot ? M.d_amom + 4 * ((size_t)M.quat_off[i] + a) : nullptr,
                                   with_forces ? M.d_res + 3 * ((size_t)M.mol_off[i] + a) : nullptr,
                                   with_forces && M.torq_off[i] >= 0 ? M.d_res + 3 * (size_t)M.nmols + 3 * ((size_t)M.torq_off[i] + a) : nullptr,
                                   M.d_mdpart);
      k_md_sums_finish<<<NSUM, MB, 0, st>>>(M.d_mdpart, nb, dst);
      e->launches += 2;
   }
   MDB_CUDA(cudaGetLastError());
   return 0;
}
static int md_sums(mdb_engine *e, const double h[9], int slot, bool with_forces, cudaStream_t st)
{
   return mdb_md_sums_range(e, h, slot, with_forces, 0, e->mf.nmols, st);
}

// eval_forces() on the resident state: no input upload, no force download (mdb_molframe.cu pieces)
extern "C" int mdb_md_eval_forces(mdb_engine *e, const double h[9], int surface_dipole, int do_recip, void *stream)
{
   auto &M = e->mf;
   cudaStream_t st = (cudaStream_t)stream;
   if (mdb_evalf_pre(e, h, M.d_in, st)) return -1;
   double *d_out = e->d_out_own;
   if (mdb_zero_out(e, d_out, stream) || mdb_build_cells(e, stream) || mdb_force_real(e, d_out, stream)) return -1;
   if (do_recip && mdb_force_recip(e, d_out, stream)) return -1;
   if (M.rdf_counts) {
      unsigned long long *dst = M.rdf_counts;
      M.rdf_counts = nullptr;
      if (mdb_rdf_counts(e, M.rdf_limit, M.rdf_nbins, dst, stream)) return -1;
   }
   return mdb_evalf_tail(e, h, M.d_in, d_out, 0, M.nmols, surface_dipole, do_recip, st);
}

// One NVE step of do_step() (src/accel.c:705-768, const_temp = const_pressure = 0) on the resident state.
// half_sums != 0: also the sums at the half step (for H_0, :718-726).  h_scal (HOST, mdb_md_scalars() doubles, may be
// NULL: use mdb_md_result) receives the scalars; synchronises `stream`.
extern "C" int mdb_md_step(mdb_engine *e, const double h[9], double step, double ts, int surface_dipole, int do_recip,
                           int half_sums, double *h_scal, void *stream)
{
   auto &M = e->mf;
   cudaStream_t st = (cudaStream_t)stream;
   if (!M.d_mom) { mdb_set_error("mdb_md_step: mdb_md_set_dynamics / mdb_md_upload_state were not called"); return -1; }
   MDB_CUDA(cudaSetDevice(e->device));
   if (mdb_md_coords(e, h, 0.5 * step, ts, stream)) return -1;
   if (mdb_md_eval_forces(e, h, surface_dipole, do_recip, stream)) return -1;
   if (mdb_md_momenta(e, h, 0.5 * step * ts, stream)) return -1;
   if (half_sums && md_sums(e, h, 1, false, st)) return -1;
   if (mdb_md_momenta(e, h, 0.5 * step * ts, stream)) return -1;
   for (size_t i = 0; i < M.sp.size(); i++)                         /* framework constraint, :751-755 */
      if (M.sp[i].framework && M.sp[i].nmols > 0)
         MDB_CUDA(cudaMemsetAsync(M.d_mom + 3 * (size_t)M.mol_off[i], 0, sizeof(double) * 3 * (size_t)M.sp[i].nmols, st));
   if (mdb_md_coords(e, h, 0.5 * step, ts, stream)) return -1;
   if (md_sums(e, h, 0, true, st)) return -1;
   const size_t ns = mdb_md_scalars(e);
   MDB_CUDA(cudaMemcpyAsync(M.d_mdscal, M.d_res + 3 * (size_t)M.nmols + 3 * (size_t)M.nmols_r, sizeof(double) * MDB_EVAL_SCALARS,
                            cudaMemcpyDeviceToDevice, st));
   MDB_CUDA(cudaMemcpyAsync(M.h_mdscal, M.d_mdscal, sizeof(double) * ns, cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaStreamSynchronize(st));
   // the bad-quaternion counter is an integer in the last slot
   unsigned int bad;
   memcpy(&bad, M.h_mdscal + ns - 1, sizeof bad);
   M.h_mdscal[ns - 1] = (double)bad;
   if (bad) MDB_CUDA(cudaMemsetAsync(M.d_mdscal + ns - 1, 0, sizeof(double), st));
   if (h_scal) memcpy(h_scal, M.h_mdscal, sizeof(double) * ns);
   return 0;
}

extern "C" const double *mdb_md_result(const mdb_engine *e) { return e->mf.h_mdscal; }
extern "C" int mdb_md_sums_now(mdb_engine *e, const double h[9], double *h_sums, void *stream)
{  // sums of the current momenta (slot 0, without forces) -> HOST [nspecies][MDB_MD_SUMS]; synchronises
   auto &M = e->mf;
   cudaStream_t st = (cudaStream_t)stream;
   MDB_CUDA(cudaSetDevice(e->device));
   if (md_sums(e, h, 0, false, st)) return -1;
   MDB_CUDA(cudaMemcpyAsync(h_sums, M.d_mdscal + MDB_EVAL_SCALARS, sizeof(double) * NSUM * M.sp.size(), cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaStreamSynchronize(st));
   return 0;
}
