/*
 * oracle/leapfrog.c -- TEST INFRASTRUCTURE ONLY.  "parity pinned": tests/test_oracle_leapfrog.py checks every function
 * bit for bit against the reference's own leapfrog.c / quaterns.c / matrix.c compiled in place
 * (oracle/_ref/libmoldyref_evalf.so).
 *
 * Restatement of the NVE leapfrog sub-steps that surround eval_forces() in do_step() (src/accel.c:626-827; SURVEY 8f
 * rank 4, the next row of the hot-path contract): the oracle for a device-resident integrator.  Written the way a GPU
 * kernel would evaluate them -- ONE fused pass per molecule instead of the reference's array-wide sub-passes -- which
 * is bit-identical because every molecule is independent and each expression keeps the reference's operation order
 * (compile with -ffp-contract=off; sin/cos/sqrt are libm's on both sides).
 *
 *   orc_leapf_com    leapf_com  src/leapfrog.c:134-146  (G = h'h, G^-1 step/(mass s), mvaxpy src/matrix.c:211-225, escape :117-129)
 *   orc_leapf_mom    leapf_mom  src/leapfrog.c:150-158  (mom += step h' force)
 *   orc_leapf_amom   leapf_amom src/leapfrog.c:408-419
 *   orc_leapf_quat   leapf_quat_b :262-337 (symmetric=1) / leapf_quat_a :345-392 (symmetric=0), const_temp = 0;
 *                    make_rot :205-219, make_rot_amom :224-241, rot_substep :247-253, normalise :96-113,
 *                    q_mul / q_conj_mul src/quaterns.c:33-98
 *   orc_symmetry_axis  the near-symmetry axis leapf_quat_b picks on its FIRST call and keeps for every species
 *                    (function-static `saxis`, src/leapfrog.c:272,286-301)
 * Not restated (out of the NVE path): thermostat (leapf_s, leapf_smom_*, gleap_therm), cell dynamics (gleap_cell,
 * update_hmom, leapf_h/hmom).
 */
#include <float.h>
#include <math.h>

#define INERTIA_MIN 1.0e-14

static void lf_mat_mul(const double a[3][3], const double b[3][3], double c[3][3])
{
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) c[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
}
static void lf_invert(const double a[3][3], double b[3][3])          /* src/matrix.c:160-190 */
{
   double d = 0.0;
   for (int i = 0, j = 1, k = 2; i < 3; i++, j = (j + 1) % 3, k = (k + 1) % 3)
      d += a[0][i] * (a[1][j] * a[2][k] - a[1][k] * a[2][j]);
   const double deter = 1.0 / d;
   for (int i = 0, j = 1, k = 2; i < 3; i++, j = (j + 1) % 3, k = (k + 1) % 3)
      for (int l = 0, m = 1, n = 2; l < 3; l++, m = (m + 1) % 3, n = (n + 1) % 3)
         b[l][i] = deter * (a[j][m] * a[k][n] - a[j][n] * a[k][m]);
}
static void lf_mvaxpy1(const double a[3][3], const double *x, double *y)
{
   const double y0 = y[0] + a[0][0] * x[0] + a[0][1] * x[1] + a[0][2] * x[2];
   const double y1 = y[1] + a[1][0] * x[0] + a[1][1] * x[1] + a[1][2] * x[2];
   const double y2 = y[2] + a[2][0] * x[0] + a[2][1] * x[1] + a[2][2] * x[2];
   y[0] = y0; y[1] = y1; y[2] = y2;
}

void orc_leapf_com(double step, double *c_of_m, const double *mom, const double *h9, double s, double mass, int nmols)
{
   double h[3][3], ht[3][3], G[3][3], Gi[3][3];
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) { h[i][j] = h9[3 * i + j]; ht[j][i] = h9[3 * i + j]; }
   lf_mat_mul(ht, h, G);
   lf_invert(G, Gi);
   const double f = step / (mass * s);
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Gi[i][j] = f * Gi[i][j];
   for (int m = 0; m < nmols; m++) {
      double *c = c_of_m + 3 * m;
      lf_mvaxpy1(Gi, mom + 3 * m, c);
      c[0] -= floor(c[0] + 0.5); c[1] -= floor(c[1] + 0.5); c[2] -= floor(c[2] + 0.5);
   }
}

void orc_leapf_mom(double step, const double *h9, double *mom, const double *force, int nmols)
{
   double ht[3][3];
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) ht[j][i] = step * h9[3 * i + j];
   for (int m = 0; m < nmols; m++) lf_mvaxpy1(ht, force + 3 * m, mom + 3 * m);
}

void orc_leapf_amom(double step, double *amom, const double *torque, int nmols)
{
   for (int m = 0; m < nmols; m++)
      for (int k = 0; k < 3; k++) amom[4 * m + 1 + k] += step * torque[3 * m + k];
}

/* r = p q ; conj: p -> p^-1 (src/quaterns.c:33-98); r may alias p or q */
static void lf_qmul(const double *p, const double *q, double *r, int conj_p)
{
   const double p0 = p[0], p1 = conj_p ? -p[1] : p[1], p2 = conj_p ? -p[2] : p[2], p3 = conj_p ? -p[3] : p[3];
   const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
   r[0] = p0 * q0 - p1 * q1 - p2 * q2 - p3 * q3;
   r[1] = p1 * q0 + p0 * q1 - p3 * q2 + p2 * q3;
   r[2] = p2 * q0 + p3 * q1 + p0 * q2 - p1 * q3;
   r[3] = p3 * q0 - p2 * q1 + p1 * q2 + p0 * q3;
}
static void lf_axis_substep(double step, int axis, double rinertia, double *amom, double *quat)
{
   double rot[4] = {0.0, 0.0, 0.0, 0.0};
   const double angle = 0.5 * step * rinertia * amom[axis + 1];      /* make_rot */
   rot[0] = cos(angle); rot[axis + 1] = sin(angle);
   lf_qmul(rot, amom, amom, 1);                                       /* rot_substep */
   lf_qmul(amom, rot, amom, 0);
   lf_qmul(quat, rot, quat, 0);
}

static void lf_rinertia(const double *inertia, double r[3])
{
   for (int i = 0; i < 3; i++)
      r[i] = inertia[i] / (inertia[(i + 1) % 3] + inertia[(i + 2) % 3]) < INERTIA_MIN ? 0.0 : 1.0 / inertia[i];
}

int orc_symmetry_axis(const double *inertia)
{
   double r[3], idmin = DBL_MAX;
   int saxis = 0;
   lf_rinertia(inertia, r);
   for (int i = 0; i < 3; i++) {
      const double idiff = fabs(r[(i + 1) % 3] - r[(i + 2) % 3]);
      if (idiff < idmin) { idmin = idiff; saxis = i; }
   }
   return saxis;
}

/* returns the number of quaternions whose norm was off by more than 1e-4 before normalisation (the reference: FATAL) */
int orc_leapf_quat(double step, double *quat, double *amom, const double *inertia, double ts, int nmols, int symmetric,
                   int saxis)
{
   const double stepdts = step / ts;
   double ri[3];
   int bad = 0;
   lf_rinertia(inertia, ri);
   const int o1 = (saxis + 1) % 3, o2 = (saxis + 2) % 3;
   for (int m = 0; m < nmols; m++) {
      double *q = quat + 4 * m, *a = amom + 4 * m;
      if (symmetric) {                                     /* leapf_quat_b */
         lf_axis_substep(0.5 * stepdts, o1, ri[o1] - ri[o2], a, q);
         lf_axis_substep(stepdts, saxis, ri[saxis] - ri[o2], a, q);
         {                                                 /* make_rot_amom + q_mul */
            double rot[4];
            const double samom = sqrt(a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
            const double ramom = 1.0 / (samom + (8 * DBL_MIN));
            const double angle = 0.5 * stepdts * ri[o2] * samom;
            const double ca = cos(angle), sa = sin(angle);
            rot[0] = ca; rot[1] = sa * ramom * a[1]; rot[2] = sa * ramom * a[2]; rot[3] = sa * ramom * a[3];
            lf_qmul(q, rot, q, 0);
         }
         lf_axis_substep(0.5 * stepdts, o1, ri[o1] - ri[o2], a, q);
      } else {                                             /* leapf_quat_a */
         lf_axis_substep(0.5 * stepdts, 0, ri[0], a, q);
         lf_axis_substep(0.5 * stepdts, 1, ri[1], a, q);
         lf_axis_substep(stepdts, 2, ri[2], a, q);
         lf_axis_substep(0.5 * stepdts, 1, ri[1], a, q);
         lf_axis_substep(0.5 * stepdts, 0, ri[0], a, q);
      }
      double norm = 0.0;                                   /* normalise */
      for (int j = 0; j < 4; j++) norm += q[j] * q[j];
      norm = sqrt(norm);
      if (fabs(norm - 1.0) > 1.0e-4) bad++;
      for (int j = 0; j < 4; j++) q[j] /= norm;
   }
   return bad;
}

/* ---- kinetic-energy reductions the step needs from the momenta (values()/do_step: tot_ke, stress_kin, src/accel.c:330-357)
 *   orc_trans_ke     trans_ke    src/algorith.c:221-239   p = (h^-1)' mom (mat_vec_mul src/matrix.c:46-90), sum p.p / (2 m s^2)
 *   orc_rot_ke       rot_ke      src/algorith.c:244-257
 *   orc_energy_dyad  energy_dyad src/algorith.c:261-284   ke_dyad[i][j] += sum p_i p_j / (m s^2)
 * Sequential sums in the reference's order (vdot, src/auxil.c: generic version); a device reduction is compared with
 * these to a tolerance, not bit for bit. */
static void lf_real_mom(const double *h9, const double *mom, int m, double p[3])
{
   double h[3][3], hi[3][3];
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) h[i][j] = h9[3 * i + j];
   lf_invert(h, hi);
   const double *x = mom + 3 * m;
   for (int i = 0; i < 3; i++) p[i] = hi[0][i] * x[0] + hi[1][i] * x[1] + hi[2][i] * x[2];     /* transposed inverse */
}

double orc_trans_ke(const double *h9, const double *mom, double s, double mass, int nmols)
{
   double ke = 0.0, p[3];
   for (int m = 0; m < nmols; m++) {
      lf_real_mom(h9, mom, m, p);
      ke += p[0] * p[0]; ke += p[1] * p[1]; ke += p[2] * p[2];
   }
   return ke / (2.0 * mass * (s * s));
}

double orc_rot_ke(const double *amom, double s, const double *inertia, int nmols)
{
   double ke = 0.0;
   for (int i = 0; i < 3; i++)
      if (inertia[i] > INERTIA_MIN) {
         double dot = 0.0;
         for (int m = 0; m < nmols; m++) dot += amom[4 * m + i + 1] * amom[4 * m + i + 1];
         ke += dot / inertia[i];
      }
   return 0.5 * ke / (s * s);
}

void orc_energy_dyad(double *ke_dyad9, const double *h9, double s, const double *mom, double mass, int nmols)
{
   double dot[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, p[3];
   for (int m = 0; m < nmols; m++) {
      lf_real_mom(h9, mom, m, p);
      for (int i = 0; i < 3; i++)
         for (int j = 0; j < 3; j++) dot[i][j] += p[i] * p[j];
   }
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) ke_dyad9[3 * i + j] += dot[i][j] / (mass * (s * s));
}
