/* oracle/molframe.c -- TEST INFRASTRUCTURE ONLY (not linked into, imported by or shipped with the product).
 *
 * CPU restatement of the molecular-frame steps that surround force_calc()/ewald() inside eval_forces()
 * (the next row of the hot-path contract): site generation from centres of mass and
 * quaternions, and the reduction of site forces to molecular forces and torques.  Written from the
 * reference's behaviour with flat arrays; every function cites the lines it follows and keeps their
 * operation order, so that it can be compared bit for bit with the reference's own algorith.c compiled in
 * place (oracle/_ref/libmoldyref_mol.so, tests/test_oracle_molframe.py).
 *
 * Layouts: h[9] row-major cell matrix; com_s[nmols][3] scaled centres of mass; quat[nmols][4] or NULL;
 * pfs[nsites][3] principal-frame sites; site/force rows x,y,z of length nmols*nsites (molecule-major).
 */
#include <math.h>
#include <stddef.h>

/* src/quaterns.c:129-156 q_to_rot */
static void orc_q_to_rot(const double *q, double r[3][3])
{
   double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
   const double a01 = 2.0 * q0 * q1, a02 = 2.0 * q0 * q2, a03 = 2.0 * q0 * q3;
   const double a12 = 2.0 * q1 * q2, a13 = 2.0 * q1 * q3, a23 = 2.0 * q2 * q3;
   r[0][1] = a12 - a03; r[0][2] = a13 + a02;
   r[1][0] = a12 + a03; r[1][2] = a23 - a01;
   r[2][0] = a13 - a02; r[2][1] = a23 + a01;
   q0 = q0 * q0; q1 = q1 * q1; q2 = q2 * q2; q3 = q3 * q3;
   r[0][0] = q0 + q1 - q2 - q3;
   r[1][1] = q0 - q1 + q2 - q3;
   r[2][2] = q0 - q1 - q2 + q3;
}

/* src/matrix.c:160-168 det, :174-190 invert */
static double orc_det(const double a[3][3])
{
   double d = 0.0;
   for (int i = 0, j = 1, k = 2; i < 3; i++, j = (j + 1) % 3, k = (k + 1) % 3)
      d += a[0][i] * (a[1][j] * a[2][k] - a[1][k] * a[2][j]);
   return d;
}
static void orc_invert(const double a[3][3], double b[3][3])
{
   const double deter = 1.0 / orc_det(a);
   for (int i = 0, j = 1, k = 2; i < 3; i++, j = (j + 1) % 3, k = (k + 1) % 3)
      for (int l = 0, m = 1, n = 2; l < 3; l++, m = (m + 1) % 3, n = (n + 1) % 3)
         b[l][i] = deter * (a[j][m] * a[k][n] - a[j][n] * a[k][m]);
}

/* src/algorith.c:169-217 make_sites (rotate :76-97, mat_vec_mul src/matrix.c:76-83).
 * sitepbc != 0: every site is brought into the cell on its own; 0 (MOLPBC): molecules stay whole. */
void orc_make_sites(const double *h9, const double *com_s, const double *quat, const double *pfs,
                    double *x, double *y, double *z, int nmols, int nsites, int sitepbc)
{
   double h[3][3], hinv[3][3];
   double *site[3] = {x, y, z};
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) h[i][j] = h9[3 * i + j];
   orc_invert(h, hinv);
   for (int imol = 0; imol < nmols; imol++) {
      const double *s = com_s + 3 * imol;
      double c[3], rot[3][3];
      for (int i = 0; i < 3; i++) c[i] = h[i][0] * s[0] + h[i][1] * s[1] + h[i][2] * s[2];
      if (quat) orc_q_to_rot(quat + 4 * imol, rot);
      for (int is = 0; is < nsites; is++) {
         const double *p = pfs + 3 * is;
         for (int i = 0; i < 3; i++) {
            const double rel = quat ? rot[i][0] * p[0] + rot[i][1] * p[1] + rot[i][2] * p[2] : p[i];
            site[i][imol * nsites + is] = rel + c[i];
         }
      }
   }
   if (sitepbc)
      for (int k = 0; k < nmols * nsites; k++) {
         const double tx = floor(hinv[0][0] * x[k] + hinv[0][1] * y[k] + hinv[0][2] * z[k] + 0.5);
         const double ty = floor(hinv[1][0] * x[k] + hinv[1][1] * y[k] + hinv[1][2] * z[k] + 0.5);
         const double tz = floor(hinv[2][0] * x[k] + hinv[2][1] * y[k] + hinv[2][2] * z[k] + 0.5);
         x[k] -= h[0][0] * tx + h[0][1] * ty + h[0][2] * tz;
         y[k] -= h[1][0] * tx + h[1][1] * ty + h[1][2] * tz;
         z[k] -= h[2][0] * tx + h[2][1] * ty + h[2][2] * tz;
      }
}

/* src/algorith.c:111-128 mol_force: force[imol][i] = sum over the molecule's sites, in site order */
void orc_mol_force(const double *fx, const double *fy, const double *fz, double *force, int nsites, int nmols)
{
   const double *f[3] = {fx, fy, fz};
   for (int imol = 0; imol < nmols; imol++)
      for (int i = 0; i < 3; i++) {
         double a = 0.0;
         for (int is = 0; is < nsites; is++) a += f[i][is + imol * nsites];
         force[3 * imol + i] = a;
      }
}

/* src/algorith.c:133-163 mol_torque: site forces rotated into the principal frame (transposed rotation
 * matrix), torque = sum r x f over the sites in the principal frame */
void orc_mol_torque(const double *fx, const double *fy, const double *fz, const double *pfs, double *torque,
                    const double *quat, int nsites, int nmols)
{
   const double *f[3] = {fx, fy, fz};
   for (int imol = 0; imol < nmols; imol++) {
      double rot[3][3], t[3][3];
      orc_q_to_rot(quat + 4 * imol, rot);
      for (int i = 0; i < 3; i++)
         for (int j = 0; j < 3; j++) t[i][j] = rot[j][i];
      for (int i = 0, j = 1, k = 2; i < 3; i++, j = (j + 1) % 3, k = (k + 1) % 3) {
         double torq = 0.0;
         for (int is = 0; is < nsites; is++) {
            const double a0 = f[0][is + imol * nsites], a1 = f[1][is + imol * nsites], a2 = f[2][is + imol * nsites];
            const double pk = t[k][0] * a0 + t[k][1] * a1 + t[k][2] * a2;
            const double pj = t[j][0] * a0 + t[j][1] * a1 + t[j][2] * a2;
            torq += pfs[3 * is + j] * pk - pfs[3 * is + k] * pj;
         }
         torque[3 * imol + i] = torq;
      }
   }
}
