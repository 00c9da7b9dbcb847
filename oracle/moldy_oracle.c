/*
 * moldy_oracle.c -- TEST INFRASTRUCTURE ONLY (see moldy_oracle.h).
 *
 * CPU restatement of the reference's link-cell real-space force loop, pair
 * potentials and reciprocal-space Ewald sum.  Data structures are ours (CSR
 * cell lists, flat arrays); the arithmetic -- operation order inside a pair,
 * order in which pairs are summed -- follows the reference so that results can
 * be compared bit-for-bit with oracle/_ref.  Compile with -ffp-contract=off.
 *
 * Reference lines followed are cited at each function (paths under
 * /root/reference/src).
 */
#include "moldy_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI_ 3.14159265358979323846
#define SQ(x) ((x) * (x))
#define H(a, i, j) ((a)[3 * (i) + (j)])

/* ---- 3x3 algebra: matrix.c:159-185 (det by first-row cofactors accumulated from
 * 0.0, inverse = adjoint * (1/det)) ------------------------------------------- */
double orc_det3(const double a[9])
{
   double d = 0.0;
   int i;
   for (i = 0; i < 3; i++) {
      int j = (i + 1) % 3, k = (i + 2) % 3;
      d += H(a, 0, i) * (H(a, 1, j) * H(a, 2, k) - H(a, 1, k) * H(a, 2, j));
   }
   return d;
}

void orc_invert3(const double a[9], double inv[9])
{
   double rd = 1.0 / orc_det3(a);
   int c, r;
   for (c = 0; c < 3; c++) {
      int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      for (r = 0; r < 3; r++) {
         int r1 = (r + 1) % 3, r2 = (r + 2) % 3;
         H(inv, r, c) = rd * (H(a, c1, r1) * H(a, c2, r2) - H(a, c1, r2) * H(a, c2, r1));
      }
   }
}

/* ---- A&S 7.1.26, auxil.c:586-593 -------------------------------------------- */
static const double AS1 = 0.254829592, AS2 = -0.284496736, AS3 = 1.421413741, AS4 = -1.453152027,
                    AS5 = 1.061405429, ASP = 0.3275911;
static double as_poly(double t) { return t * (AS1 + t * (AS2 + t * (AS3 + t * (AS4 + t * AS5)))); }

double orc_err_fn(double x)
{
   if (x < 0.0) return -orc_err_fn(-x);
   return 1.0 - as_poly(1.0 / (1.0 + ASP * x)) * exp(-x * x);
}

/* ---- safe binning, force.c:119-137 ------------------------------------------- */
int orc_cellbin(double s, int n, double fn, double eps, int *err)
{
   int b;
   if (s < -0.5 + eps || s >= 0.5 - eps) {
      if (s < -0.5 + eps && s >= -0.5 - eps) s = -0.5;
      else if (s >= 0.5 - eps && s <= 0.5 + eps) s = 0.5 - eps;
      else if (err) (*err)++;
   }
   b = (int)floor((s + 0.5) * fn);
   if ((b >= n || b < 0) && err) (*err)++;
   return b;
}

/* ---- one site pair: kernel.c:182-461.  fij = -phi'(r)/r ------------------------ */
void orc_pair(int ptype, double alpha, double norm, double r_sqr, double qq, const double *p,
              double *fij, double *phi)
{
   double r, rinv, rinv2, coul_e = 0.0, coul_f = 0.0;
   double e1, e2, e3, i4, i6, i8, i12;
   if (alpha > 0.0) {
      double ar, t, scr;
      r = sqrt(r_sqr);
      ar = alpha * r;
      t = 1.0 / (1.0 + ASP * ar);
      scr = qq * exp(-SQ(ar));
      rinv = 1.0 / r;
      coul_e = as_poly(t) * scr * rinv;
      coul_f = coul_e + norm * scr;
      rinv2 = SQ(rinv);
   } else if (ptype == 0) {
      r = rinv = 0.0;
      rinv2 = 1.0 / r_sqr;
   } else {
      r = sqrt(r_sqr);
      rinv = 1.0 / r;
      rinv2 = SQ(rinv);
   }
   switch (ptype) {
   case 0:                                    /* Lennard-Jones, kernel.c:188-213, 362-373 */
      i6 = SQ(p[1]) * rinv2;
      i6 = i6 * i6 * i6;
      i12 = SQ(i6);
      if (alpha > 0.0) {
         *phi = coul_e + p[0] * (i12 - i6);
         *fij = rinv2 * (6.0 * p[0] * (2 * i12 - i6) + coul_f);
      } else {
         *phi = p[0] * (i12 - i6);
         *fij = rinv2 * 6.0 * p[0] * (2 * i12 - i6);
      }
      break;
   case 1:                                    /* Buckingham, kernel.c:214-238, 374-387 */
      e1 = p[1] * exp(-p[2] * r);
      if (alpha > 0.0) {
         i6 = p[0] * (rinv2 * rinv2 * rinv2);
         *phi = coul_e - i6 + e1;
         *fij = rinv2 * (-6.0 * i6 + coul_f) + p[2] * e1 * rinv;
      } else {
         i6 = p[0] * rinv2 * rinv2 * rinv2;
         *phi = -i6 + e1;
         *fij = -rinv2 * 6.0 * i6 + p[2] * e1 * rinv;
      }
      break;
   case 2:                                    /* MCY, kernel.c:239-263, 388-399 */
      e1 = p[0] * exp(-p[1] * r);
      e2 = -p[2] * exp(-p[3] * r);
      if (alpha > 0.0) {
         *phi = coul_e + e1 + e2;
         *fij = (p[1] * e1 + p[3] * e2) * rinv + coul_f * rinv2;
      } else {
         *phi = e1 + e2;
         *fij = (p[1] * e1 + p[3] * e2) * rinv;
      }
      break;
   case 3:                                    /* generic, kernel.c:264-295, 400-420 */
      e1 = p[0] * exp(-p[1] * r);
      i4 = SQ(rinv2);
      i6 = rinv2 * i4;
      i8 = p[5] * SQ(i4);
      i12 = p[2] * SQ(i6);
      i4 *= p[3];
      i6 *= p[4];
      if (alpha > 0.0) {
         *phi = coul_e + e1 + i12 - i4 - i6 - i8;
         *fij = rinv2 * (12.0 * i12 - 4.0 * i4 - 6.0 * i6 - 8.0 * i8 + coul_f) + p[1] * e1 * rinv;
      } else {
         *phi = e1 + i12 - i4 - i6 - i8;
         *fij = rinv2 * (12.0 * i12 - 4.0 * i4 - 6.0 * i6 - 8.0 * i8) + p[1] * e1 * rinv;
      }
      break;
   case 4:                                    /* HIW, kernel.c:325-354, 442-460 */
      i4 = SQ(rinv2);
      i6 = rinv2 * i4;
      if (alpha > 0.0) i12 = p[2] * SQ(i6);
      else i12 = SQ(i6) * p[2];
      i6 *= p[1];
      i4 *= p[0];
      *phi = (alpha > 0.0 ? coul_e + i4 : i4) + i6 + i12;
      if (alpha > 0.0) *fij = rinv2 * (4.0 * i4 + 6.0 * i6 + 12.0 * i12 + coul_f);
      else *fij = rinv2 * (4.0 * i4 + 6.0 * i6 + 12.0 * i12);
      break;
   default:                                   /* Morse/BIG (6), kernel.c:296-324, 421-441 */
      e1 = p[0] * exp((p[1] - r) * p[2]);
      i6 = p[3] * (rinv2 * rinv2 * rinv2);
      e2 = p[4] * exp(-2.0 * p[5] * (r - p[6]));
      e3 = -p[4] * 2.0 * exp(-p[5] * (r - p[6]));
      if (alpha > 0.0) {
         *phi = coul_e + e1 - i6 + e2 + e3;
         *fij = rinv2 * (-6.0 * i6 + coul_f) + rinv * (p[2] * e1 + (2.0 * p[5]) * e2 + p[5] * e3);
      } else {
         *phi = e1 - i6 + e2 + e3;
         *fij = rinv2 * (-6.0 * i6) + rinv * (p[2] * e1 + (2.0 * p[5]) * e2 + p[5] * e3);
      }
      break;
   }
}

/* ---- long-range correction integrals, kernel.c:103-151 ------------------------- */
static double tail3(double rc, double b) { return SQ(rc) / b + 2 * rc / SQ(b) + 2.0 / (b * b * b); }
double orc_dist_pot(const double *p, double rc, int ptype)
{
   double rc3 = rc * rc * rc;
   switch (ptype) {
   default:
   case 0: { double s = SQ(p[1]) / rc; return p[0] * (s * s * s) / 3.0; }
   case 1: return p[2] > 1.0e-7 ? p[0] / (3.0 * rc3) - p[1] * exp(-p[2] * rc) * tail3(rc, p[2]) : p[0] / (3.0 * rc3);
   case 2: return p[3] > 1.0e-7 ? p[2] * tail3(rc, p[3]) * exp(-p[3] * rc) : 0.0;
   case 3: {
      double t = -p[2] / (9.0 * (rc3 * rc3 * rc3)) + p[3] / rc + p[4] / (3.0 * rc3) + p[5] / (5.0 * SQ(rc) * rc3);
      return p[1] > 1.0e-7 ? -p[0] * exp(-p[1] * rc) * tail3(rc, p[1]) + t : t;
   }
   case 6: return p[5] != 0.0 ? p[3] / (3.0 * rc3) + 2.0 * p[4] * tail3(rc, p[5]) * exp(-p[5] * (rc - p[6]))
                              : p[3] / (3.0 * rc3);
   case 4: return -p[0] / rc - p[1] / rc3 / 3.0 - p[2] / (rc3 * rc3 * rc3) / 9.0;
   }
}

/* ---- neighbour-cell half lists, force.c:167-226 (lazy) and :273-421 (strict) ---- */
typedef struct { double G[9], hti[9]; int mx, my, mz; } metric_t;

static void metric(const double h[9], double rc, int nx, int ny, int nz, metric_t *m)
{
   double ht[9];
   int i, j;
   for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) H(ht, j, i) = H(h, i, j);
   for (i = 0; i < 3; i++)
      for (j = 0; j < 3; j++)
         H(m->G, i, j) = H(ht, i, 0) * H(h, 0, j) + H(ht, i, 1) * H(h, 1, j) + H(ht, i, 2) * H(h, 2, j);
   orc_invert3(ht, m->hti);
   m->mx = (int)ceil(rc * nx * sqrt(SQ(H(m->hti, 0, 0)) + SQ(H(m->hti, 1, 0)) + SQ(H(m->hti, 2, 0))));
   m->my = (int)ceil(rc * ny * sqrt(SQ(H(m->hti, 0, 1)) + SQ(H(m->hti, 1, 1)) + SQ(H(m->hti, 2, 1))));
   m->mz = (int)ceil(rc * nz * sqrt(SQ(H(m->hti, 0, 2)) + SQ(H(m->hti, 1, 2)) + SQ(H(m->hti, 2, 2))));
}

static double gdist(const metric_t *m, int ix, int iy, int iz, int nx, int ny, int nz)
{
   double s[3], d = 0.0;
   int i, j;
   s[0] = (double)ix / nx; s[1] = (double)iy / ny; s[2] = (double)iz / nz;
   for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) d += s[i] * H(m->G, i, j) * s[j];
   return d;
}

int orc_half_list(const double h[9], double rc, int strict, int nx, int ny, int nz, int *out, int cap)
{
   metric_t m;
   int n = 0, ix, iy, iz, i, j, k;
   metric(h, rc, nx, ny, nz, &m);
   if (!strict) {
      for (ix = 0; ix < m.mx; ix++)
         for (iy = (ix == 0 ? 0 : -m.my); iy < m.my; iy++)
            for (iz = (ix == 0 && iy == 0 ? 0 : -m.mz); iz < m.mz; iz++)
               if (gdist(&m, ix, iy, iz, nx, ny, nz) < SQ(rc)) {
                  if (ix > nx || iy > ny || iy < -ny || iz > nz || iz < -nz) return -1;
                  if (3 * n + 3 > cap) return -2;
                  out[3 * n] = ix; out[3 * n + 1] = iy; out[3 * n + 2] = iz; n++;
               }
      return n;
   }
   {
      const int mx = m.mx, my = m.my, mz = m.mz, wy = 2 * my + 2, wz = 2 * mz + 2;
      const int nn[3] = {nx, ny, nz}, mm[3] = {mx, my, mz};
      char *map = (char *)calloc((size_t)(mx + 1) * wy * wz, 1);
      int a;
#define MAP(x, y, z) map[((size_t)(x) * wy + ((y) + my + 1)) * wz + ((z) + mz + 1)]
      for (ix = 0; ix < mx; ix++)
         for (iy = (ix == 0 ? 0 : -my); iy < my; iy++)
            for (iz = (ix == 0 && iy == 0 ? 0 : -mz); iz < mz; iz++)
               if (gdist(&m, ix, iy, iz, nx, ny, nz) < SQ(rc))
                  for (i = 0; i <= 1; i++) for (j = -1; j <= 1; j++) for (k = -1; k <= 1; k++)
                     MAP(ix + i, iy + j, iz + k) = 1;
      for (a = 0; a < 3; a++) {
         int b = (a + 1) % 3, g = (b + 1) % 3, fc[4][3], f;
         double proj[3] = {0, 0, 0}, mod = 0.0;
         for (i = 0; i < 3; i++) {
            mod += H(m.hti, i, a);
            proj[i] += H(m.hti, i, a) * H(m.hti, i, (a + i) % 3);
         }
         for (i = 0; i < 3; i++) proj[i] *= (mm[a] - 1) * nn[i] / (nn[a] * mod);
         for (i = 0; i < 3; i++) fc[0][i] = fc[1][i] = fc[2][i] = fc[3][i] = (int)floor(proj[i]);
         for (f = 0; f < 4; f++) fc[f][a] = mm[a];
         fc[1][b] = fc[3][b] = (int)ceil(proj[b]);
         fc[2][g] = fc[3][g] = (int)ceil(proj[g]);
         for (f = 0; f < 4; f++) {
            if (fc[f][0] < 0) for (j = 0; j < 3; j++) fc[f][j] = -fc[f][j];
            MAP(fc[f][0], fc[f][1], fc[f][2]) = 1;
         }
      }
      for (ix = 0; ix <= mx; ix++)
         for (iy = (ix == 0 ? 0 : -my - 1); iy <= my; iy++)
            for (iz = (ix == 0 && iy == 0 ? 0 : -mz - 1); iz <= mz; iz++)
               if (MAP(ix, iy, iz)) {
                  if (ix > nx || iy > ny || iy < -ny || iz > nz || iz < -nz) { free(map); return -1; }
                  if (3 * n + 3 > cap) { free(map); return -2; }
                  out[3 * n] = ix; out[3 * n + 1] = iy; out[3 * n + 2] = iz; n++;
               }
      free(map);
   }
   return n;
}

/* ---- grid, force.c:1153-1157 ---------------------------------------------------- */
static void grid_dims(const orc_system *s, int *nx, int *ny, int *nz)
{
   double sub = s->subcell;
   if (sub <= 0.0) sub = s->cutoff / 5.0;
   *nx = (int)(H(s->h, 0, 0) / sub + 0.5);
   *ny = (int)(H(s->h, 1, 1) / sub + 0.5);
   *nz = (int)(H(s->h, 2, 2) / sub + 0.5);
}

/* ---- per-site cell index: fill_cells force.c:460-472 with mat_vec_mul's product
 * order (matrix.c:76-83) and eps = 8*DBL_EPSILON (force.c:437) -------------------- */
int orc_cell_ids(const orc_system *s, const double *x, const double *y, const double *z, int *cell)
{
   double hinv[9], eps = 8.0 * 2.220446049250313e-16;
   int nx, ny, nz, i, err = 0;
   grid_dims(s, &nx, &ny, &nz);
   orc_invert3(s->h, hinv);
   for (i = 0; i < s->nsites; i++) {
      double s0, s1, s2;
      if (s->molpbc && i < s->nsites_xf) {      /* whole molecule by its scaled c-of-m, force.c:474-484 */
         const double *c = s->c_of_m + 3 * s->site_mol[i];
         s0 = c[0]; s1 = c[1]; s2 = c[2];
      } else {
         s0 = hinv[0] * x[i] + hinv[1] * y[i] + hinv[2] * z[i];
         s1 = hinv[3] * x[i] + hinv[4] * y[i] + hinv[5] * z[i];
         s2 = hinv[6] * x[i] + hinv[7] * y[i] + hinv[8] * z[i];
      }
      int bx = orc_cellbin(s0, nx, (double)nx, eps, &err);
      int by = orc_cellbin(s1, ny, (double)ny, eps, &err);
      int bz = orc_cellbin(s2, nz, (double)nz, eps, &err);
      cell[i] = bz + nz * (by + ny * bx);
   }
   return err;
}

/* ---- real-space loop: force.c:856-997 -------------------------------------------- */
int orc_force_calc(const orc_system *s, const double *x, const double *y, const double *z,
                   double *fx, double *fy, double *fz, orc_result *res)
{
   const int n = s->nsites;
   int nx, ny, nz, ncells, nhalf, i, c, k;
   int *cell = (int *)malloc(sizeof(int) * (size_t)n), *start, *members, *half;
   double reloc[27][3], gforce[27][3];
   const double norm = 2.0 * s->alpha / sqrt(PI_);
   const double rc2 = SQ(s->cutoff), rc2far = 10000.0 * rc2;
   int cap, *nab, *img;
   double *nx_, *ny_, *nz_, *nq, *fjx, *fjy, *fjz, *rx, *ry, *rz, *r2, *fij;
   double s00 = 0, s01 = 0, s02 = 0, s11 = 0, s12 = 0, s22 = 0, pe = 0.0, npairs = 0.0;
   int too_close = 0, rc = 0;
   const int has_fw = s->nsites_xf < n;

   grid_dims(s, &nx, &ny, &nz);
   ncells = nx * ny * nz;
   memset(res, 0, sizeof *res);
   res->nx = nx; res->ny = ny; res->nz = nz;
   res->n_bin_errors = orc_cell_ids(s, x, y, z, cell);
   /* CSR cell lists; members of a cell in DESCENDING site order (the reference
    * prepends to a linked list, force.c:470-471) */
   start = (int *)calloc((size_t)ncells + 1, sizeof(int));
   members = (int *)malloc(sizeof(int) * (size_t)n);
   for (i = 0; i < n; i++) start[cell[i] + 1]++;
   for (c = 0; c < ncells; c++) start[c + 1] += start[c];
   {
      int *cur = (int *)malloc(sizeof(int) * (size_t)ncells);
      memcpy(cur, start, sizeof(int) * (size_t)ncells);
      /* list order = reverse insertion order: molecules from last to first; a per-site node list
       * (site mode, framework) gives its sites in descending order, a per-molecule node (molpbc,
       * force.c:474-484) keeps its sites ascending */
      i = n - 1;
      while (i >= 0) {
         int lo = i, k;
         while (lo > 0 && s->site_mol[lo - 1] == s->site_mol[i]) lo--;
         if (s->molpbc && i < s->nsites_xf)
            for (k = lo; k <= i; k++) members[cur[cell[k]]++] = k;
         else
            for (k = i; k >= lo; k--) members[cur[cell[k]]++] = k;
         i = lo - 1;
      }
      free(cur);
   }
   cap = 3 * 4 * 64 * 64 * 64;
   half = (int *)malloc(sizeof(int) * (size_t)cap);
   nhalf = orc_half_list(s->h, s->cutoff, s->strict_cutoff, nx, ny, nz, half, cap);
   if (nhalf < 0) { rc = -1; goto done0; }
   res->n_nabors = nhalf;
   k = 0;
   {
      int a, b, g;
      for (a = -1; a <= 1; a++) for (b = -1; b <= 1; b++) for (g = -1; g <= 1; g++, k++) {
         reloc[k][0] = H(s->h, 0, 0) * a + H(s->h, 0, 1) * b + H(s->h, 0, 2) * g;
         reloc[k][1] = H(s->h, 1, 0) * a + H(s->h, 1, 1) * b + H(s->h, 1, 2) * g;
         reloc[k][2] = H(s->h, 2, 0) * a + H(s->h, 2, 1) * b + H(s->h, 2, 2) * g;
      }
   }
   memset(gforce, 0, sizeof gforce);
   /* neighbour-site scratch, sized by the largest possible list */
   {
      int maxc = 0;
      for (c = 0; c < ncells; c++) if (start[c + 1] - start[c] > maxc) maxc = start[c + 1] - start[c];
      cap = maxc * nhalf * 2 + 8;
   }
   nab = (int *)malloc(sizeof(int) * (size_t)cap); img = (int *)malloc(sizeof(int) * (size_t)cap);
   nx_ = (double *)malloc(sizeof(double) * (size_t)cap * 12);
   ny_ = nx_ + cap; nz_ = ny_ + cap; nq = nz_ + cap; fjx = nq + cap; fjy = fjx + cap; fjz = fjy + cap;
   rx = fjz + cap; ry = rx + cap; rz = ry + cap; r2 = rz + cap; fij = r2 + cap;

   for (c = s->ithread; c < ncells; c += s->nthreads) {
      int cx, cy, cz, nnab = 0, nnf = 0, pass, e, m, jmin = 0;
      if (start[c] == start[c + 1]) continue;
      cx = c / (ny * nz); cy = c / nz - ny * cx; cz = c - nz * (cy + ny * cx);
      /* site_neighbour_list, force.c:521-569: non-framework sites of every stencil
       * cell first, then framework sites of stencil cells 1.. */
      for (pass = 0; pass < (has_fw ? 2 : 1); pass++) {
         for (e = (pass == 0 ? 0 : 1); e < nhalf; e++) {
            int tx = cx + half[3 * e], ty = cy + half[3 * e + 1], tz = cz + half[3 * e + 2];
            int ia = 1, ib = 1, ig = 1, tc, kimg;
            if (tx < 0) { tx += nx; ia = 0; } else if (tx >= nx) { tx -= nx; ia = 2; }
            if (ty < 0) { ty += ny; ib = 0; } else if (ty >= ny) { ty -= ny; ib = 2; }
            if (tz < 0) { tz += nz; ig = 0; } else if (tz >= nz) { tz -= nz; ig = 2; }
            tc = tz + nz * (ty + ny * tx);
            kimg = 9 * ia + 3 * ib + ig;
            for (m = start[tc]; m < start[tc + 1]; m++) {
               int j = members[m];
               if ((j >= s->nsites_xf) == pass) { nab[nnab] = j; img[nnab] = kimg; nnab++; }
            }
         }
         if (pass == 0) nnf = nnab;
      }
      if (!has_fw) nnf = nnab;
      for (m = 0; m < nnab; m++) {
         int j = nab[m];
         nx_[m] = x[j]; ny_[m] = y[j]; nz_[m] = z[j]; nq[m] = s->chg[j];
         fjx[m] = fjy[m] = fjz[m] = 0.0;
      }
      for (m = start[c]; m < start[c + 1]; m++) {
         const int isite = members[m];
         const int fw = isite >= s->nsites_xf;
         int jmax, j;
         const double *prow = s->potpar + (size_t)s->site_type[isite] * s->max_id * 8;
         double ppe = 0.0, a0 = 0.0, a1 = 0.0, a2 = 0.0;
         if (fw) { jmin = 0; jmax = nnf; } else { jmax = nnab; jmin++; }
         for (j = 0; j < jmax; j++) {                       /* mk_r_sqr, force.c:692-711 */
            const double *rv = reloc[img[j]];
            double dx = nx_[j] - x[isite] + rv[0], dy = ny_[j] - y[isite] + rv[1], dz = nz_[j] - z[isite] + rv[2];
            r2[j] = dx * dx + dy * dy + dz * dz;
            rx[j] = dx; ry[j] = dy; rz[j] = dz;
         }
         for (j = jmin; j < jmax; j++)                      /* TOO_CLOSE, force.c:939-949 */
            if (r2[j] < 0.25 && s->site_mol[isite] != s->site_mol[nab[j]]) too_close++;
         if (s->strict_cutoff && !s->molpbc)                /* force.c:951-954 */
            for (j = jmin; j < jmax; j++) if (r2[j] > rc2) r2[j] = rc2far;
         for (j = jmin; j < jmax; j++) {                    /* kernel + mk_forces */
            double f, phi;
            orc_pair(s->ptype, s->alpha, norm, r2[j], nq[j] * s->chg[isite],
                     prow + (size_t)s->site_type[nab[j]] * 8, &f, &phi);
            ppe += phi;
            fij[j] = f;
         }
         for (j = jmin; j < jmax; j++) {
            double cxf = fij[j] * rx[j], cyf = fij[j] * ry[j], czf = fij[j] * rz[j];
            a0 -= cxf; a1 -= cyf; a2 -= czf;
            cxf += fjx[j]; cyf += fjy[j]; czf += fjz[j];
            fjx[j] = cxf; fjy[j] = cyf; fjz[j] = czf;
         }
         pe += ppe;
         fx[isite] += a0; fy[isite] += a1; fz[isite] += a2;
         if (jmax > jmin) npairs += jmax - jmin;
      }
      for (m = 0; m < nnab; m++) {                          /* scatter_forces, force.c:750-778 */
         int j = nab[m];
         double *g = gforce[img[m]];
         fx[j] = fjx[m] + fx[j]; fy[j] = fjy[m] + fy[j]; fz[j] = fjz[m] + fz[j];
         g[0] = fjx[m] + g[0]; g[1] = fjy[m] + g[1]; g[2] = fjz[m] + g[2];
      }
   }
   /* site virial, force.c:973-997 (note: uses the force arrays as they now stand) */
   for (i = 0; i < n; i++) {
      s00 += x[i] * fx[i]; s01 += y[i] * fx[i]; s02 += z[i] * fx[i];
      s11 += y[i] * fy[i]; s12 += z[i] * fy[i]; s22 += z[i] * fz[i];
   }
   for (k = 0; k < 27; k++) {
      s00 += reloc[k][0] * gforce[k][0]; s01 += reloc[k][1] * gforce[k][0]; s02 += reloc[k][2] * gforce[k][0];
      s11 += reloc[k][1] * gforce[k][1]; s12 += reloc[k][2] * gforce[k][1]; s22 += reloc[k][2] * gforce[k][2];
   }
   res->stress[0] = s00; res->stress[1] = s01; res->stress[2] = s02;
   res->stress[4] = s11; res->stress[5] = s12; res->stress[8] = s22;
   res->pe = pe; res->npairs = npairs; res->n_too_close = too_close;
   free(nab); free(img); free(nx_);
done0:
   free(cell); free(start); free(members); free(half);
   return rc;
}

/* ---- RDF binning pass: force.c:1010-1103 (rdf_inner: same cell walk as the force loop, over the
 * STRICT neighbour list of radius `limit`, force.c:1306-1308) with rdf.c:94-108 (rdf_accum:
 * bin = (int)(nbins/limit * sqrt(r^2)), counted when bin < nbins).  counts[pair][bin] += 1 with
 * pair = (idi <= idj) in the order init_rdf lays the histograms out (rdf.c:82-90). ------------- */
int orc_rdf(const orc_system *s, const double *x, const double *y, const double *z, double limit, int nbins,
            double *counts)
{
   const int n = s->nsites;
   int nx, ny, nz, ncells, nhalf, i, c, k;
   int *cell = (int *)malloc(sizeof(int) * (size_t)n), *start, *members, *half;
   double reloc[27][3];
   const double rbin = nbins / limit;
   int cap, *nab, *img, rc = 0;
   const int has_fw = s->nsites_xf < n;

   grid_dims(s, &nx, &ny, &nz);
   ncells = nx * ny * nz;
   orc_cell_ids(s, x, y, z, cell);
   start = (int *)calloc((size_t)ncells + 1, sizeof(int));
   members = (int *)malloc(sizeof(int) * (size_t)n);
   for (i = 0; i < n; i++) start[cell[i] + 1]++;
   for (c = 0; c < ncells; c++) start[c + 1] += start[c];
   {
      int *cur = (int *)malloc(sizeof(int) * (size_t)ncells);
      memcpy(cur, start, sizeof(int) * (size_t)ncells);
      i = n - 1;
      while (i >= 0) {                                   /* list order as in orc_force_calc */
         int lo = i, kk;
         while (lo > 0 && s->site_mol[lo - 1] == s->site_mol[i]) lo--;
         if (s->molpbc && i < s->nsites_xf)
            for (kk = lo; kk <= i; kk++) members[cur[cell[kk]]++] = kk;
         else
            for (kk = i; kk >= lo; kk--) members[cur[cell[kk]]++] = kk;
         i = lo - 1;
      }
      free(cur);
   }
   cap = 3 * 4 * 64 * 64 * 64;
   half = (int *)malloc(sizeof(int) * (size_t)cap);
   nhalf = orc_half_list(s->h, limit, 1, nx, ny, nz, half, cap);
   if (nhalf < 0) { rc = -1; goto done; }
   k = 0;
   {
      int a, b, g;
      for (a = -1; a <= 1; a++) for (b = -1; b <= 1; b++) for (g = -1; g <= 1; g++, k++) {
         reloc[k][0] = H(s->h, 0, 0) * a + H(s->h, 0, 1) * b + H(s->h, 0, 2) * g;
         reloc[k][1] = H(s->h, 1, 0) * a + H(s->h, 1, 1) * b + H(s->h, 1, 2) * g;
         reloc[k][2] = H(s->h, 2, 0) * a + H(s->h, 2, 1) * b + H(s->h, 2, 2) * g;
      }
   }
   {
      int maxc = 0;
      for (c = 0; c < ncells; c++) if (start[c + 1] - start[c] > maxc) maxc = start[c + 1] - start[c];
      cap = maxc * nhalf * 2 + 8;
   }
   nab = (int *)malloc(sizeof(int) * (size_t)cap); img = (int *)malloc(sizeof(int) * (size_t)cap);
   for (c = s->ithread; c < ncells; c += s->nthreads) {
      int cx, cy, cz, nnab = 0, nnf = 0, pass, e, m, jmin = 0;
      if (start[c] == start[c + 1]) continue;
      cx = c / (ny * nz); cy = c / nz - ny * cx; cz = c - nz * (cy + ny * cx);
      for (pass = 0; pass < (has_fw ? 2 : 1); pass++) {   /* site_neighbour_list, force.c:521-569 */
         for (e = (pass == 0 ? 0 : 1); e < nhalf; e++) {
            int tx = cx + half[3 * e], ty = cy + half[3 * e + 1], tz = cz + half[3 * e + 2];
            int ia = 1, ib = 1, ig = 1, tc, kimg;
            if (tx < 0) { tx += nx; ia = 0; } else if (tx >= nx) { tx -= nx; ia = 2; }
            if (ty < 0) { ty += ny; ib = 0; } else if (ty >= ny) { ty -= ny; ib = 2; }
            if (tz < 0) { tz += nz; ig = 0; } else if (tz >= nz) { tz -= nz; ig = 2; }
            tc = tz + nz * (ty + ny * tx);
            kimg = 9 * ia + 3 * ib + ig;
            for (m = start[tc]; m < start[tc + 1]; m++) {
               int j = members[m];
               if ((j >= s->nsites_xf) == pass) { nab[nnab] = j; img[nnab] = kimg; nnab++; }
            }
         }
         if (pass == 0) nnf = nnab;
      }
      if (!has_fw) nnf = nnab;
      for (m = start[c]; m < start[c + 1]; m++) {
         const int isite = members[m], idi = s->site_type[isite];
         const int fw = isite >= s->nsites_xf;
         int jmax, j;
         if (fw) { jmin = 0; jmax = nnf; } else { jmax = nnab; jmin++; }
         for (j = jmin; j < jmax; j++) {
            const double *rv = reloc[img[j]];
            const int jj = nab[j], idj = s->site_type[jj];
            double dx = x[jj] - x[isite] + rv[0], dy = y[jj] - y[isite] + rv[1], dz = z[jj] - z[isite] + rv[2];
            double rsq = dx * dx + dy * dy + dz * dz;
            int bin = rbin * sqrt(rsq);
            if (bin < nbins) {
               const int a = idi < idj ? idi : idj, b = idi < idj ? idj : idi;
               counts[(size_t)((a - 1) * s->max_id - (a - 1) * a / 2 + (b - a)) * nbins + bin] += 1.0;
            }
         }
      }
   }
   free(nab); free(img);
done:
   free(cell); free(start); free(members); free(half);
   return rc;
}

/* ---- 4-lane sum of auxil.c:222-256 ----------------------------------------------- */
static double lane_sum(int n, const double *v)
{
   double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0;
   int i = 0;
   for (; i < n - 3; i += 4) { l0 += v[i]; l1 += v[i + 1]; l2 += v[i + 2]; l3 += v[i + 3]; }
   for (; i < n; i++) l0 += v[i];
   return l0 + l1 + l2 + l3;
}

/* ---- reciprocal space: ewald.c:309-311, 435-583 ------------------------------------ */
int orc_ewald(const orc_system *s, const double *x, const double *y, const double *z,
              double *fx, double *fy, double *fz, orc_result *res)
{
   const int n = s->nsites, nxf = s->nsites_xf;
   double hp[9], vol = orc_det3(s->h);
   const double *as = hp, *bs = hp + 3, *cs = hp + 6;
   const double r4a = -1.0 / (4.0 * s->alpha * s->alpha), kc2 = SQ(s->k_cutoff);
   int hmax, kmax, lmax, h, k, l, i, nhkl = 0, per, q0, q1, q, hl = -1000000, kl = -1000000;
   double *tab, **ch, **sh, **ck, **sk, **cl, **sl, *qc, *qs, *chk, *shk;
   typedef struct { double kx, ky, kz; int h, k, l; } kv_t;
   kv_t *kv;
   double pe = 0.0;

   memset(res, 0, sizeof *res);
   orc_invert3(s->h, hp);
   for (i = 0; i < 9; i++) hp[i] = 2 * PI_ * hp[i];
   hmax = (int)floor(s->k_cutoff / (2 * PI_) * H(s->h, 0, 0));
   kmax = (int)floor(s->k_cutoff / (2 * PI_) * sqrt(SQ(H(s->h, 0, 1)) + SQ(H(s->h, 1, 1))));
   lmax = (int)floor(s->k_cutoff / (2 * PI_) * sqrt(SQ(H(s->h, 0, 2)) + SQ(H(s->h, 1, 2)) + SQ(H(s->h, 2, 2))));
   kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(4 * (hmax + 1) * (kmax + 1) * (lmax + 1)));
   for (h = 0; h <= hmax; h++)
      for (k = (h == 0 ? 0 : -kmax); k <= kmax; k++) {
         double kx = h * as[0] + k * bs[0], ky = h * as[1] + k * bs[1], kzt = h * as[2] + k * bs[2];
         double ksq = SQ(kx) + SQ(ky);
         for (l = (h == 0 && k == 0 ? 1 : -lmax); l <= lmax; l++) {
            double kz = kzt + l * cs[2];
            if (SQ(kz) + ksq < kc2) {
               kv[nhkl].h = h; kv[nhkl].k = k; kv[nhkl].l = l;
               kv[nhkl].kx = kx; kv[nhkl].ky = ky; kv[nhkl].kz = kz;
               nhkl++;
            }
         }
      }
   res->nhkl = nhkl;
   /* power tables by libm cos/sin + angle addition, ewald.c:148-193 */
   tab = (double *)malloc(sizeof(double) * (size_t)n * (2 * (hmax + kmax + lmax + 3) + 4));
   ch = (double **)malloc(sizeof(double *) * (size_t)(2 * (hmax + kmax + lmax + 3)));
   sh = ch + hmax + 1; ck = sh + hmax + 1; sk = ck + kmax + 1; cl = sk + kmax + 1; sl = cl + lmax + 1;
   {
      double *p = tab;
      int ax;
      for (i = 0; i < 2 * (hmax + kmax + lmax + 3); i++, p += n) ch[i] = p;
      qc = p; qs = p + n; chk = p + 2 * n; shk = p + 3 * n;
      for (ax = 0; ax < 3; ax++) {
         double **c = ax == 0 ? ch : ax == 1 ? ck : cl, **sn = ax == 0 ? sh : ax == 1 ? sk : sl;
         const double *ks = ax == 0 ? as : ax == 1 ? bs : cs;
         int mmax = ax == 0 ? hmax : ax == 1 ? kmax : lmax, m;
         for (i = 0; i < n; i++) { c[0][i] = 1.0; sn[0][i] = 0.0; }
         if (mmax >= 1) {
            for (i = 0; i < n; i++) {
               double kr = ks[0] * x[i] + ks[1] * y[i] + ks[2] * z[i];
               c[1][i] = cos(kr); sn[1][i] = sin(kr);
            }
            for (m = 2; m <= mmax; m++)
               for (i = 0; i < n; i++) {
                  double cc = c[m - 1][i] * c[1][i] - sn[m - 1][i] * sn[1][i];
                  sn[m][i] = sn[m - 1][i] * c[1][i] + c[m - 1][i] * sn[1][i];
                  c[m][i] = cc;
               }
         }
      }
      for (l = 0; l <= lmax; l++)
         for (i = 0; i < n; i++) { cl[l][i] *= s->chg[i]; sl[l][i] *= s->chg[i]; }
   }
   /* this rank's block of k-vectors, ewald.c:495-496 */
   per = (nhkl + s->nthreads - 1) / s->nthreads;
   q0 = s->ithread * per;
   q1 = (s->ithread + 1) * per < nhkl ? (s->ithread + 1) * per : nhkl;
   for (q = q0; q < q1; q++) {
      const kv_t *v = &kv[q];
      const double *chh = ch[v->h], *shh = sh[v->h], *ckk = ck[abs(v->k)], *skk = sk[abs(v->k)];
      const double *cll = cl[abs(v->l)], *sll = sl[abs(v->l)];
      double ksq = v->kx * v->kx + v->ky * v->ky + v->kz * v->kz;
      double coeff = 2.0 / ((0.25 / PI_) * vol) * exp(ksq * r4a) / ksq;
      double coeff2 = 2.0 * (1.0 - ksq * r4a) / ksq;
      double cn, sn_, cf, sf, ct, st, pek;
      if (v->h != hl || v->k != kl) {                       /* qsincos, ewald.c:95-142 */
         if (v->k >= 0)
            for (i = 0; i < n; i++) {
               double t = chh[i] * ckk[i] - shh[i] * skk[i];
               shk[i] = shh[i] * ckk[i] + chh[i] * skk[i];
               chk[i] = t;
            }
         else
            for (i = 0; i < n; i++) {
               double t = chh[i] * ckk[i] + shh[i] * skk[i];
               shk[i] = shh[i] * ckk[i] - chh[i] * skk[i];
               chk[i] = t;
            }
      }
      hl = v->h; kl = v->k;
      if (v->l >= 0)
         for (i = 0; i < n; i++) {
            double t = chk[i] * cll[i] - shk[i] * sll[i];
            qs[i] = shk[i] * cll[i] + chk[i] * sll[i];
            qc[i] = t;
         }
      else
         for (i = 0; i < n; i++) {
            double t = chk[i] * cll[i] + shk[i] * sll[i];
            qs[i] = shk[i] * cll[i] - chk[i] * sll[i];
            qc[i] = t;
         }
      cn = lane_sum(nxf, qc); sn_ = lane_sum(nxf, qs);
      cf = lane_sum(n - nxf, qc + nxf); sf = lane_sum(n - nxf, qs + nxf);
      ct = cn + cf; st = sn_ + sf;
      pek = 0.5 * coeff * (cn * (cn + cf + cf) + sn_ * (sn_ + sf + sf));      /* ewald.c:539-541 */
      pe += pek;
      st *= coeff; ct *= coeff; sn_ *= coeff; cn *= coeff;
      res->stress[0] += pek - pek * coeff2 * v->kx * v->kx;
      res->stress[1] -= pek * coeff2 * v->kx * v->ky;
      res->stress[2] -= pek * coeff2 * v->kx * v->kz;
      res->stress[4] += pek - pek * coeff2 * v->ky * v->ky;
      res->stress[5] -= pek * coeff2 * v->ky * v->kz;
      res->stress[8] += pek - pek * coeff2 * v->kz * v->kz;
      for (i = 0; i < nxf; i++) {                           /* ewald.c:558-566 */
         double fc = qs[i] * ct - qc[i] * st;
         fx[i] = fx[i] + v->kx * fc; fy[i] = fy[i] + v->ky * fc; fz[i] += v->kz * fc;
      }
      for (i = nxf; i < n; i++) {                           /* ewald.c:571-579 */
         double fc = qs[i] * cn - qc[i] * sn_;
         fx[i] = fx[i] + v->kx * fc; fy[i] = fy[i] + v->ky * fc; fz[i] += v->kz * fc;
      }
   }
   res->pe = pe;
   free(tab); free(ch); free(kv);
   return 0;
}

/* ---- first-call constants ---------------------------------------------------------- */
static double dist3(const double *a, const double *b)
{
   return sqrt(SQ(a[0] - b[0]) + SQ(a[1] - b[1]) + SQ(a[2] - b[2]));
}

/* force.c:1158-1169 + :654-668 (pe_intra through poteval: unit reference charge,
 * product of charges on the neighbour side) */
double orc_eintra(const orc_system *s, int nspecies, const int *spec_nsites, const int *spec_nmols,
                  const int *spec_framework, const double *pfs)
{
   const double norm = 2.0 * s->alpha / sqrt(PI_);
   double e = 0.0;
   int sp, site0 = 0, p0 = 0;
   for (sp = 0; sp < nspecies; sp++) {
      if (!spec_framework[sp]) {
         double em = 0.0;
         int a, b;
         for (a = 0; a < spec_nsites[sp]; a++)
            for (b = a + 1; b < spec_nsites[sp]; b++) {
               double r = dist3(pfs + 3 * (p0 + b), pfs + 3 * (p0 + a)), f, phi;
               const double *p = s->potpar + ((size_t)s->site_type[site0 + a] * s->max_id + s->site_type[site0 + b]) * 8;
               orc_pair(s->ptype, s->alpha, norm, SQ(r), (s->chg[site0 + b] * s->chg[site0 + a]) * 1.0, p, &f, &phi);
               em += 0.0 + phi;
            }
         e += spec_nmols[sp] * em;
      }
      site0 += spec_nmols[sp] * spec_nsites[sp];
      p0 += spec_nsites[sp];
   }
   return e;
}

/* ewald.c:367-425 */
void orc_self_energy(const orc_system *s, int nspecies, const int *spec_nsites, const int *spec_nmols,
                     const int *spec_framework, const double *pfs, double *self_e, double *sheet_e)
{
   double self = 0.0, sheet = 0.0, sq = 0.0, sqsq = 0.0, sqxf;
   int sp = 0, site0 = 0, p0 = 0, i;
   while (sp < nspecies && !spec_framework[sp]) {
      double intra = 0.0;
      int a, b;
      for (a = 0; a < spec_nsites[sp]; a++)
         for (b = a + 1; b < spec_nsites[sp]; b++) {
            double r = dist3(pfs + 3 * (p0 + a), pfs + 3 * (p0 + b));
            intra += s->chg[site0 + a] * s->chg[site0 + b] * orc_err_fn(s->alpha * r) / r;
         }
      self += spec_nmols[sp] * intra;
      site0 += spec_nsites[sp] * spec_nmols[sp];
      p0 += spec_nsites[sp];
      sp++;
   }
   for (i = 0; i < site0; i++) { sq += s->chg[i]; sqsq += SQ(s->chg[i]); }
   self += s->alpha / sqrt(PI_) * sqsq;
   sqxf = sq;
   for (; i < s->nsites; i++) sq += s->chg[i];
   if (sp != nspecies) sheet = PI_ * SQ(sq - sqxf) / (2.0 * SQ(s->alpha));
   if (fabs(sq) * (4.07497263794495e-14 * 1.e-3 * 1.05482230112e-05 / 1.60217733e-19) > 1.0e-5)
      sheet -= PI_ * SQ(sq) / (2.0 * SQ(s->alpha));
   *self_e = self; *sheet_e = sheet;
}
