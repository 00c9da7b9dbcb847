// mdb_host.cpp -- host-side set-up of the hot path: everything that decides a
// *discrete* set (link-cell grid, neighbour-cell stencil, k-vector list) is
// evaluated here in plain IEEE double arithmetic, in the same operation order
// as the reference so the sets come out identical (compile WITHOUT fp
// contraction).  Nothing here is per-site work.
//
// Behaviour followed (file:line under /root/reference):
//   3x3 inverse by adjoint / determinant ........ src/matrix.c:159-185
//   grid nx,ny,nz from subcell ................... src/force.c:1153-1157
//   half neighbour-cell list (lazy) .............. src/force.c:167-226
//   strict neighbour-cell list ................... src/force.c:273-421
//   image translation vectors .................... src/force.c:1284-1293
//   hmax/kmax/lmax, k-vector half space .......... src/ewald.c:309-311, 435-461
//   A&S 7.1.26 error function .................... src/auxil.c:586-593
#include <math.h>
#include <stdio.h>
#include <algorithm>
#include <map>
#include "mdb_internal.h"

static std::string g_err;
void mdb_set_error(const std::string &s) { g_err = s; }
extern "C" const char *mdb_last_error(void) { return g_err.c_str(); }

// ---- 3x3 helpers (row-major double[9]) -------------------------------------
#define M(a, i, j) ((a)[3 * (i) + (j)])

double mdb_det3(const double a[9])
{
   double d = 0.0;
   for (int i = 0; i < 3; i++) {
      int j = (i + 1) % 3, k = (i + 2) % 3;
      d += M(a, 0, i) * (M(a, 1, j) * M(a, 2, k) - M(a, 1, k) * M(a, 2, j));
   }
   return d;
}

void mdb_invert3(const double a[9], double b[9])
{
   double rdet = 1.0 / mdb_det3(a);
   for (int i = 0; i < 3; i++) {
      int j = (i + 1) % 3, k = (i + 2) % 3;
      for (int l = 0; l < 3; l++) {
         int m = (l + 1) % 3, n = (l + 2) % 3;
         M(b, l, i) = rdet * (M(a, j, m) * M(a, k, n) - M(a, j, n) * M(a, k, m));
      }
   }
}

static void transpose3(const double a[9], double b[9])
{
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) M(b, j, i) = M(a, i, j);
}

static void matmul3(const double a[9], const double b[9], double c[9])
{
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
         M(c, i, j) = M(a, i, 0) * M(b, 0, j) + M(a, i, 1) * M(b, 1, j) + M(a, i, 2) * M(b, 2, j);
}

static double colnorm(const double a[9], int c)
{
   return sqrt(M(a, 0, c) * M(a, 0, c) + M(a, 1, c) * M(a, 1, c) + M(a, 2, c) * M(a, 2, c));
}

double mdb_err_fn(double x)
{
   if (x < 0.0) return -mdb_err_fn(-x);
   const double E1 = 0.254829592, E2 = -0.284496736, E3 = 1.421413741, E4 = -1.453152027,
                E5 = 1.061405429, PP = 0.3275911;
   double t = 1.0 / (1.0 + PP * x);
   double poly = t * (E1 + t * (E2 + t * (E3 + t * (E4 + t * E5))));
   return 1.0 - poly * exp(-x * x);
}

// ---- neighbour-cell stencil ------------------------------------------------
struct CellMetric {
   double G[9];      // h^T h
   double hti[9];    // (h^T)^-1
   int mx, my, mz;
};

static CellMetric cell_metric(const double h[9], double cutoff, int nx, int ny, int nz)
{
   CellMetric c;
   double htr[9];
   transpose3(h, htr);
   matmul3(htr, h, c.G);
   mdb_invert3(htr, c.hti);
   c.mx = (int)ceil(cutoff * nx * colnorm(c.hti, 0));
   c.my = (int)ceil(cutoff * ny * colnorm(c.hti, 1));
   c.mz = (int)ceil(cutoff * nz * colnorm(c.hti, 2));
   return c;
}

static inline double grid_dist2(const CellMetric &c, int ix, int iy, int iz, int nx, int ny, int nz)
{
   double s[3] = {(double)ix / nx, (double)iy / ny, (double)iz / nz};
   double d = 0.0;
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) d += s[i] * M(c.G, i, j) * s[j];
   return d;
}

static bool outside_images(int ix, int iy, int iz, int nx, int ny, int nz)
{  // the reference supports exactly one shell of periodic images (IMCELL_XTRA = 1)
   return ix > nx || ix < -nx || iy > ny || iy < -ny || iz > nz || iz < -nz;
}

static bool half_list_lazy(const double h[9], double cutoff, int nx, int ny, int nz,
                           std::vector<int> &out, std::string &err)
{
   CellMetric c = cell_metric(h, cutoff, nx, ny, nz);
   double rc2 = cutoff * cutoff;
   for (int ix = 0; ix < c.mx; ix++)
      for (int iy = (ix == 0 ? 0 : -c.my); iy < c.my; iy++)
         for (int iz = (ix == 0 && iy == 0 ? 0 : -c.mz); iz < c.mz; iz++)
            if (grid_dist2(c, ix, iy, iz, nx, ny, nz) < rc2) {
               if (outside_images(ix, iy, iz, nx, ny, nz)) {
                  err = "Cutoff radius > 1 * cell dimension.";
                  return false;
               }
               out.push_back(ix); out.push_back(iy); out.push_back(iz);
            }
   return true;
}

static bool half_list_strict(const double h[9], double cutoff, int nx, int ny, int nz,
                             std::vector<int> &out, std::string &err)
{
   CellMetric c = cell_metric(h, cutoff, nx, ny, nz);
   const int mx = c.mx, my = c.my, mz = c.mz;
   double rc2 = cutoff * cutoff;
   // occupancy map over ix in [0,mx], iy in [-my-1,my], iz in [-mz-1,mz]
   const int wy = 2 * my + 2, wz = 2 * mz + 2;
   std::vector<char> map((size_t)(mx + 1) * wy * wz, 0);
   auto at = [&](int ix, int iy, int iz) -> char & {
      return map[((size_t)ix * wy + (iy + my + 1)) * wz + (iz + mz + 1)];
   };
   // every cell that has a corner-to-corner vector shorter than the cutoff
   for (int ix = 0; ix < mx; ix++)
      for (int iy = (ix == 0 ? 0 : -my); iy < my; iy++)
         for (int iz = (ix == 0 && iy == 0 ? 0 : -mz); iz < mz; iz++)
            if (grid_dist2(c, ix, iy, iz, nx, ny, nz) < rc2)
               for (int i = 0; i <= 1; i++)
                  for (int j = -1; j <= 1; j++)
                     for (int k = -1; k <= 1; k++) at(ix + i, iy + j, iz + k) = 1;
   // outermost cells that touch the cutoff sphere face-on
   const int nn[3] = {nx, ny, nz}, mm[3] = {mx, my, mz};
   for (int a = 0; a < 3; a++) {
      int b = (a + 1) % 3, g = (b + 1) % 3;
      double proj[3] = {0, 0, 0}, mod = 0.0;
      for (int i = 0; i < 3; i++) {
         mod += M(c.hti, i, a);
         proj[i] += M(c.hti, i, a) * M(c.hti, i, (a + i) % 3);
      }
      for (int i = 0; i < 3; i++) proj[i] *= (mm[a] - 1) * nn[i] / (nn[a] * mod);
      int fc[4][3];
      for (int i = 0; i < 3; i++)
         fc[0][i] = fc[1][i] = fc[2][i] = fc[3][i] = (int)floor(proj[i]);
      for (int f = 0; f < 4; f++) fc[f][a] = mm[a];
      fc[1][b] = fc[3][b] = (int)ceil(proj[b]);
      fc[2][g] = fc[3][g] = (int)ceil(proj[g]);
      for (int f = 0; f < 4; f++) {
         if (fc[f][0] < 0)
            for (int j = 0; j < 3; j++) fc[f][j] = -fc[f][j];
         if (fc[f][0] > mx || fc[f][1] < -my - 1 || fc[f][1] > my || fc[f][2] < -mz - 1 || fc[f][2] > mz) {
            err = "strict neighbour list: face cell outside map";
            return false;
         }
         at(fc[f][0], fc[f][1], fc[f][2]) = 1;
      }
   }
   for (int ix = 0; ix <= mx; ix++)
      for (int iy = (ix == 0 ? 0 : -my - 1); iy <= my; iy++)
         for (int iz = (ix == 0 && iy == 0 ? 0 : -mz - 1); iz <= mz; iz++)
            if (at(ix, iy, iz)) {
               if (outside_images(ix, iy, iz, nx, ny, nz)) {
                  err = "Cutoff radius > 1 * cell dimension.";
                  return false;
               }
               out.push_back(ix); out.push_back(iy); out.push_back(iz);
            }
   return true;
}

static void to_runs(std::map<std::pair<int, int>, std::vector<int>> &m, std::vector<StencilRun> &out)
{
   out.clear();
   for (auto &kv : m) {
      std::vector<int> &z = kv.second;
      std::sort(z.begin(), z.end());
      size_t i = 0;
      while (i < z.size()) {
         size_t j = i;
         while (j + 1 < z.size() && z[j + 1] <= z[j] + 1) j++;
         out.push_back({kv.first.first, kv.first.second, z[i], z[j]});
         i = j + 1;
      }
   }
}

bool mdb_build_real_tables(const mdb_config &c, HostTables &T, std::string &err)
{
   const double *h = c.h;
   double subcell = c.subcell;
   if (subcell <= 0.0) subcell = c.cutoff / 5.0;
   T.nx = (int)(M(h, 0, 0) / subcell + 0.5);
   T.ny = (int)(M(h, 1, 1) / subcell + 0.5);
   T.nz = (int)(M(h, 2, 2) / subcell + 0.5);
   if (T.nx < 1 || T.ny < 1 || T.nz < 1) {
      err = "link-cell grid has a zero dimension";
      return false;
   }
   mdb_invert3(h, T.hinv);
   T.vol = mdb_det3(h);
   T.half_list.clear();
   bool ok = c.strict_cutoff ? half_list_strict(h, c.cutoff, T.nx, T.ny, T.nz, T.half_list, err)
                             : half_list_lazy(h, c.cutoff, T.nx, T.ny, T.nz, T.half_list, err);
   if (!ok) return false;
   if (T.half_list.size() < 3 || T.half_list[0] != 0 || T.half_list[1] != 0 || T.half_list[2] != 0) {
      err = "neighbour list does not start with the reference cell";
      return false;
   }
   // Full stencil F = H u (-H) grouped into z-runs per (dx,dy) column.  Every
   // non-zero offset of H appears in exactly one of H, -H (H lives in the
   // half space ix>0 | ix=0,iy>0 | ix=iy=0,iz>=0), so visiting F from every
   // site touches each reference pair exactly twice.
   std::map<std::pair<int, int>, std::vector<int>> cols;
   for (size_t i = 0; i < T.half_list.size(); i += 3) {
      int ix = T.half_list[i], iy = T.half_list[i + 1], iz = T.half_list[i + 2];
      cols[{ix, iy}].push_back(iz);
      if (ix || iy || iz) cols[{-ix, -iy}].push_back(-iz);
   }
   std::map<std::pair<int, int>, std::vector<int>> hcols;
   for (size_t i = 0; i < T.half_list.size(); i += 3)
      hcols[{T.half_list[i], T.half_list[i + 1]}].push_back(T.half_list[i + 2]);
   to_runs(cols, T.runs);
   to_runs(hcols, T.runs_half);
   int k = 0;
   for (int ii = -1; ii <= 1; ii++)
      for (int jj = -1; jj <= 1; jj++)
         for (int kk = -1; kk <= 1; kk++, k++)
            for (int a = 0; a < 3; a++)
               T.reloc[k][a] = M(h, a, 0) * ii + M(h, a, 1) * jj + M(h, a, 2) * kk;
   return true;
}

// Strict half list of radius `limit` on the current grid as z-runs: the neighbour cells of the RDF
// pass (src/force.c:1306-1308, strict_neighbour_list with the RDF limit).
bool mdb_build_rdf_runs(const mdb_config &c, const HostTables &T, double limit, std::vector<StencilRun> &runs,
                        std::string &err)
{
   std::vector<int> half;
   if (!half_list_strict(c.h, limit, T.nx, T.ny, T.nz, half, err)) return false;
   if (half.size() < 3 || half[0] != 0 || half[1] != 0 || half[2] != 0) {
      err = "RDF neighbour list does not start with the reference cell";
      return false;
   }
   std::map<std::pair<int, int>, std::vector<int>> hcols;
   for (size_t i = 0; i < half.size(); i += 3) hcols[{half[i], half[i + 1]}].push_back(half[i + 2]);
   to_runs(hcols, runs);
   return true;
}

// ---- reciprocal lattice ----------------------------------------------------
bool mdb_build_recip_tables(const mdb_config &c, HostTables &T, std::string &err)
{
   const double *h = c.h;
   double hinvp[9];
   mdb_invert3(h, hinvp);
   for (int i = 0; i < 9; i++) hinvp[i] = 2 * MDB_PI * hinvp[i];
   for (int a = 0; a < 3; a++) {
      T.astar[a] = M(hinvp, 0, a);
      T.bstar[a] = M(hinvp, 1, a);
      T.cstar[a] = M(hinvp, 2, a);
   }
   // upper-triangular-h shortcuts for |a|,|b|,|c| exactly as the reference takes them
   double moda = M(h, 0, 0);
   double modb = sqrt(M(h, 0, 1) * M(h, 0, 1) + M(h, 1, 1) * M(h, 1, 1));
   double modc = sqrt(M(h, 0, 2) * M(h, 0, 2) + M(h, 1, 2) * M(h, 1, 2) + M(h, 2, 2) * M(h, 2, 2));
   T.hmax = (int)floor(c.k_cutoff / (2 * MDB_PI) * moda);
   T.kmax = (int)floor(c.k_cutoff / (2 * MDB_PI) * modb);
   T.lmax = (int)floor(c.k_cutoff / (2 * MDB_PI) * modc);
   T.vol = mdb_det3(h);
   const double kcsq = c.k_cutoff * c.k_cutoff;
   T.hk.clear(); T.hk_valid.clear(); T.slot_flags.clear();
   T.nhkl = 0; T.nslots = 0;
   const int L = T.lmax;
   std::vector<int> flags(L + 1);
   for (int hh = 0; hh <= T.hmax; hh++) {
      // sweep k = 0,1,..,kmax then (h>0) k = -1,..,-kmax; trailing empty columns are trimmed
      for (int dir = 0; dir < (hh == 0 ? 1 : 2); dir++) {
         std::vector<HkDesc> sweep;
         std::vector<std::vector<int>> sweep_flags;
         int last_valid = -1;
         for (int step = 0; step <= T.kmax; step++) {
            int kk = dir == 0 ? step : -step;
            if (dir == 1 && step == 0) continue;
            HkDesc d{};
            d.h = hh; d.k = kk;
            d.kx = hh * T.astar[0] + kk * T.bstar[0];
            d.ky = hh * T.astar[1] + kk * T.bstar[1];
            d.kzt = hh * T.astar[2] + kk * T.bstar[2];
            double ksq = d.kx * d.kx + d.ky * d.ky;
            int nl = 0;
            std::fill(flags.begin(), flags.end(), 0);
            for (int l = (hh == 0 && kk == 0 ? 1 : -L); l <= L; l++) {
               double kz = d.kzt + l * T.cstar[2];
               if (kz * kz + ksq < kcsq) {
                  flags[abs(l)] |= (l >= 0) ? 1 : 2;
                  nl = std::max(nl, abs(l) + 1);
                  T.nhkl++;
               }
            }
            d.nl = nl;
            if (dir == 0) d.code = step == 0 ? HK_NEWH : HK_KUP;
            else d.code = step == 1 ? HK_KDOWN0 : HK_KDOWN;
            if (nl > 0) last_valid = (int)sweep.size();
            sweep.push_back(d);
            sweep_flags.push_back(std::vector<int>(flags.begin(), flags.begin() + nl));
         }
         // always keep the k=0 column (it carries the h recurrence)
         int keep = std::max(last_valid, dir == 0 ? 0 : -1);
         for (int i = 0; i <= keep; i++) {
            HkDesc d = sweep[i];
            d.slot0 = T.nslots;
            if (d.nl > 0) T.hk_valid.push_back((int)T.hk.size());
            T.hk.push_back(d);
            T.slot_flags.insert(T.slot_flags.end(), sweep_flags[i].begin(), sweep_flags[i].end());
            T.nslots += d.nl;
         }
      }
   }
   std::stable_sort(T.hk_valid.begin(), T.hk_valid.end(),
                    [&](int a, int b) { return T.hk[a].nl > T.hk[b].nl; });
   (void)err;
   return true;
}

// Distance beyond which every exponential of an exponential pair potential (src/kernel.c:230-355) has fallen below
// e^-52 = 2.6e-23 of its amplitude for all site-type pairs, and the power-law rest of the potential as PT_HIW rows
// (p0/r^4 + p1/r^6 + p2/r^12) in `rest` (max_id^2 x MDB_NPOTP, may be NULL).  0: not applicable (no exponential term, a
// non-positive decay constant, or a term the rows cannot express).  Host only; mdb_split_far_runs (mdb_engine.cu) uses it.
//   Buckingham -p0/r^6 + p1 exp(-p2 r); MCY p0 exp(-p1 r) - p2 exp(-p3 r); generic p0 exp(-p1 r) + p2/r^12 - p3/r^4 - p4/r^6
//   - p5/r^8; Morse/BIG p0 exp((p1 - r) p2) - p3/r^6 + p4 (exp(-2 p5 (r - p6)) - 2 exp(-p5 (r - p6)))
extern "C" double mdb_far_radius(int ptype, int max_id, const double *potpar, double *rest)
{
   const double FAR_EXPONENT = 52.0;
   if (!(ptype == 1 || ptype == 2 || ptype == 3 || ptype == 6) || max_id <= 0 || !potpar) return 0.0;
   double r_far = 0.0;
   bool any = false;
   for (int a = 0; a < max_id; a++)
      for (int b = 0; b < max_id; b++) {
         const double *p = potpar + ((size_t)a * max_id + b) * MDB_NPOTP;
         double q[MDB_NPOTP] = {0};
         bool ok = true;
         auto term = [&](double amp, double decay, double shift) {         // amp exp(-decay (r - shift))
            if (amp == 0.0) return;
            if (!(decay > 0.0)) { ok = false; return; }
            r_far = std::max(r_far, shift + FAR_EXPONENT / decay);
            any = true;
         };
         if (ptype == 1) { term(p[1], p[2], 0.0); q[1] = -p[0]; }
         else if (ptype == 2) { term(p[0], p[1], 0.0); term(p[2], p[3], 0.0); }
         else if (ptype == 3) { term(p[0], p[1], 0.0); ok = ok && p[5] == 0.0; q[0] = -p[3]; q[1] = -p[4]; q[2] = p[2]; }
         else { term(p[0], p[2], p[1]); term(p[4], p[5], p[6]); q[1] = -p[3]; }
         if (!ok) return 0.0;
         if (rest) memcpy(rest + ((size_t)a * max_id + b) * MDB_NPOTP, q, sizeof q);
      }
   return any ? r_far : 0.0;
}
