// mdb_cells.cu -- GPU link-cell builder (HBM-bound).
//
//   k_cell_ids     per site: scaled coordinate hinv.r, safe binning, cell count
//   k_scan_*       exclusive scan of the per-cell counts -> cell_start[ncells+1]
//   k_fill         site -> slot inside its cell (atomic cursor)
//   k_sort_gather  per cell: order the slots by site index (deterministic),
//                  gather {x,y,z,q}, type, cell id, {type, z index} into cell-sorted SoA
//
// Cell assignment must be bit-identical to the reference (src/force.c:119-137
// cellbin, :460-472 fill_cells, product order of mat_vec_mul src/matrix.c:76-83):
// all arithmetic uses explicit round-to-nearest mul/add intrinsics so that nvcc
// cannot contract a*b+c into an FMA.
#include "mdb_internal.h"

static constexpr int CB = 256;

__device__ __forceinline__ int safe_bin(double rc, int nc, double fnc, double eps, int *bad)
{
   const double lo = __dadd_rn(-0.5, eps), hi = __dadd_rn(0.5, -eps);
   if (rc < lo || rc >= hi) {
      if (rc < lo && rc >= __dadd_rn(-0.5, -eps))
         rc = -0.5;
      else if (rc >= hi && rc <= __dadd_rn(0.5, eps))
         rc = hi;
      else
         *bad = 1;                       // "Co-ordinate out of range in BIN"
   }
   int ibin = (int)floor(__dmul_rn(__dadd_rn(rc, 0.5), fnc));
   if (ibin >= nc || ibin < 0) {
      *bad = 1;                          // "Rounding problem in BIN"
      ibin = min(max(ibin, 0), nc - 1);  // keep the run alive; the host reports the error
   }
   return ibin;
}

__global__ void __launch_bounds__(CB) k_cell_ids(CellParams P, int n, const double *__restrict__ x,
                                                 const double *__restrict__ y, const double *__restrict__ z,
                                                 const int *__restrict__ mol, const double *__restrict__ com,
                                                 int *__restrict__ cell, int *__restrict__ count,
                                                 unsigned long long *__restrict__ counters)
{
   int i = blockIdx.x * CB + threadIdx.x;
   if (i >= n) return;
   double s0, s1, s2;
   if (P.molpbc && i < P.nsites_xf) {
      // molecular cut-off: the whole molecule goes into the cell of its (already scaled) centre of
      // mass, src/force.c:474-484; framework sites are always binned one by one
      const int m = mol[i];
      s0 = com[3 * m]; s1 = com[3 * m + 1]; s2 = com[3 * m + 2];
   } else {
      const double a0 = x[i], a1 = y[i], a2 = z[i];
      s0 = __dadd_rn(__dadd_rn(__dmul_rn(P.hinv[0], a0), __dmul_rn(P.hinv[1], a1)), __dmul_rn(P.hinv[2], a2));
      s1 = __dadd_rn(__dadd_rn(__dmul_rn(P.hinv[3], a0), __dmul_rn(P.hinv[4], a1)), __dmul_rn(P.hinv[5], a2));
      s2 = __dadd_rn(__dadd_rn(__dmul_rn(P.hinv[6], a0), __dmul_rn(P.hinv[7], a1)), __dmul_rn(P.hinv[8], a2));
   }
   int bad = 0;
   int ix = safe_bin(s0, P.nx, P.fnx, P.eps, &bad);
   int iy = safe_bin(s1, P.ny, P.fny, P.eps, &bad);
   int iz = safe_bin(s2, P.nz, P.fnz, P.eps, &bad);
   int c = iz + P.nz * (iy + P.ny * ix);
   cell[i] = c;
   atomicAdd(&count[c], 1);
   if (bad) atomicAdd(&counters[2], 1ULL);
}

// ---- exclusive scan, three passes over tiles of SCAN_TILE ints ---------------
static constexpr int SCAN_T = 256, SCAN_PER = 8, SCAN_TILE = SCAN_T * SCAN_PER;

__device__ __forceinline__ int block_excl_scan(int v, int *total)
{
   __shared__ int wsum[SCAN_T / 32];
   int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   int inc = v;
#pragma unroll
   for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
   }
   if (lane == 31) wsum[w] = inc;
   __syncthreads();
   if (w == 0) {
      int s = lane < SCAN_T / 32 ? wsum[lane] : 0;
#pragma unroll
      for (int d = 1; d < SCAN_T / 32; d <<= 1) {
         int t = __shfl_up_sync(0xffffffffu, s, d);
         if (lane >= d) s += t;
      }
      if (lane < SCAN_T / 32) wsum[lane] = s;
   }
   __syncthreads();
   int base = w ? wsum[w - 1] : 0;
   *total = wsum[SCAN_T / 32 - 1];
   __syncthreads();
   return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_tiles(const int *__restrict__ in, int n, int *__restrict__ tile_sum)
{
   int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER;
   int s = 0;
#pragma unroll
   for (int j = 0; j < SCAN_PER; j++) s += (base + j < n) ? in[base + j] : 0;
   int tot;
   block_excl_scan(s, &tot);
   if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_sums(int *__restrict__ tile_sum, int ntiles)
{  // single block; ntiles is small (ncells / 2048)
   __shared__ int carry;
   if (threadIdx.x == 0) carry = 0;
   __syncthreads();
   for (int b = 0; b < ntiles; b += SCAN_T) {
      int i = b + threadIdx.x;
      int v = i < ntiles ? tile_sum[i] : 0;
      int tot;
      int ex = block_excl_scan(v, &tot);
      if (i < ntiles) tile_sum[i] = carry + ex;
      __syncthreads();
      if (threadIdx.x == 0) carry += tot;
      __syncthreads();
   }
}

__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const int *__restrict__ in, int n, const int *__restrict__ tile_sum,
                                                        int *__restrict__ out)
{
   int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER;
   int v[SCAN_PER], s = 0;
#pragma unroll
   for (int j = 0; j < SCAN_PER; j++) {
      v[j] = (base + j < n) ? in[base + j] : 0;
      s += v[j];
   }
   int tot;
   int ex = block_excl_scan(s, &tot) + tile_sum[blockIdx.x];
#pragma unroll
   for (int j = 0; j < SCAN_PER; j++) {
      if (base + j < n) out[base + j] = ex;
      ex += v[j];
   }
   if (base <= n - 1 && n - 1 < base + SCAN_PER) out[n] = ex;   // cell_start[ncells] = nsites
}

__global__ void __launch_bounds__(CB) k_fill(int n, const int *__restrict__ cell, const int *__restrict__ start,
                                             int *__restrict__ cursor, int *__restrict__ order)
{
   int i = blockIdx.x * CB + threadIdx.x;
   if (i >= n) return;
   int c = cell[i];
   order[start[c] + atomicAdd(&cursor[c], 1)] = i;
}

__global__ void __launch_bounds__(CB) k_sort_gather(int ncells, const int *__restrict__ start, int *__restrict__ order,
                                                    const double *__restrict__ x, const double *__restrict__ y,
                                                    const double *__restrict__ z, const double *__restrict__ q,
                                                    const int *__restrict__ type, int nsites_xf,
                                                    double4 *__restrict__ posq, int *__restrict__ stype,
                                                    int *__restrict__ scell, int2 *__restrict__ sinfo, int nz)
{
   int c = blockIdx.x * CB + threadIdx.x;
   if (c >= ncells) return;
   int b = start[c], e = start[c + 1];
   for (int i = b + 1; i < e; i++) {     // insertion sort, a handful of sites per cell
      int v = order[i], j = i - 1;
      while (j >= b && order[j] > v) {
         order[j + 1] = order[j];
         j--;
      }
      order[j + 1] = v;
   }
   for (int s = b; s < e; s++) {
      int o = order[s];
      posq[s] = make_double4(x[o], y[o], z[o], q[o]);
      stype[s] = type[o] | (o >= nsites_xf ? 0x40000000 : 0);   // bit 30: framework site
      scell[s] = c;
      sinfo[s] = make_int2(stype[s], c % nz);                  // what the tiled pair kernel reads per neighbour
   }
}

// Batches of <= NI consecutive cell-sorted sites that share a z-column (cx,cy): the unit
// of work of the tiled pair kernel.  Per-column batch counts are scanned with the same
// three-pass scan as the cell counts, so the batch list (and with it every summation
// order downstream) is deterministic.
__global__ void __launch_bounds__(CB) k_col_batches(int ncols, int nz, int ni, const int *__restrict__ start,
                                                    int *__restrict__ nb)
{
   const int col = blockIdx.x * CB + threadIdx.x;
   if (col < ncols) nb[col] = (start[(col + 1) * nz] - start[col * nz] + ni - 1) / ni;
}

__global__ void __launch_bounds__(CB) k_fill_batches(int ncols, int nz, int ni, const int *__restrict__ start,
                                                     const int *__restrict__ off, int2 *__restrict__ batches,
                                                     int *__restrict__ nbatch)
{
   const int col = blockIdx.x * CB + threadIdx.x;
   if (col >= ncols) return;
   const int s0 = start[col * nz], cnt = start[(col + 1) * nz] - s0, o = off[col];
   const int nb = (cnt + ni - 1) / ni;
   for (int k = 0; k < nb; k++) batches[o + k] = make_int2(s0 + k * ni, min(ni, cnt - k * ni) | (col << 3));
   if (col == ncols - 1) *nbatch = off[ncols];
}

static int launch_batches(mdb_engine *e, const int *start, int *scratch, int *scan_tmp, int2 *batches, int *nbatch,
                          cudaStream_t st)
{
   const int ncols = e->T.nx * e->T.ny;
   int *nb = scratch, *off = scratch + ncols + 1;
   const int ntiles = (ncols + SCAN_TILE - 1) / SCAN_TILE;
   k_col_batches<<<(ncols + CB - 1) / CB, CB, 0, st>>>(ncols, e->T.nz, MDB_NI, start, nb);
   k_scan_tiles<<<ntiles, SCAN_T, 0, st>>>(nb, ncols, scan_tmp);
   k_scan_sums<<<1, SCAN_T, 0, st>>>(scan_tmp, ntiles);
   k_scan_apply<<<ntiles, SCAN_T, 0, st>>>(nb, ncols, scan_tmp, off);
   k_fill_batches<<<(ncols + CB - 1) / CB, CB, 0, st>>>(ncols, e->T.nz, MDB_NI, start, off, batches, nbatch);
   e->launches += 5;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

int mdb_need_batches(mdb_engine *e, cudaStream_t st)
{
   if (e->batches_valid) return 0;
   if (mdb_launch_batches(e, st)) return -1;
   e->batches_valid = true;
   return 0;
}

int mdb_launch_batches(mdb_engine *e, cudaStream_t st)
{
   const int ncols = e->T.nx * e->T.ny;
   if (2 * (ncols + 1) > e->cells_cap) { mdb_set_error("batch scratch too small"); return -1; }
   // the cell-count array is free after the fill
   return launch_batches(e, e->d_start, e->d_count, e->d_scan_tmp, e->d_batches, e->d_nbatch, st);
}

// ---- site-class sublists: stream compaction of the cell-sorted list (order within a cell is kept, so a
// cell's members stay contiguous and start_K[c] is the scanned flag count at start[c])
__global__ void __launch_bounds__(CB) k_sub_flags(int n, const int *__restrict__ order, const unsigned char *__restrict__ cls,
                                                  int bit, int *__restrict__ flag)
{
   const int s = blockIdx.x * CB + threadIdx.x;
   if (s < n) flag[s] = (cls[order[s]] >> bit) & 1;
}
__global__ void __launch_bounds__(CB) k_sub_scatter(int n, const int *__restrict__ flag, const int *__restrict__ pos,
                                                    const double4 *__restrict__ posq, const int2 *__restrict__ sinfo,
                                                    const int *__restrict__ order, double4 *__restrict__ posq_k,
                                                    int2 *__restrict__ sinfo_k, int *__restrict__ order_k)
{
   const int s = blockIdx.x * CB + threadIdx.x;
   if (s >= n || !flag[s]) return;
   const int d = pos[s];
   posq_k[d] = posq[s]; sinfo_k[d] = sinfo[s]; order_k[d] = order[s];
}
__global__ void __launch_bounds__(CB) k_sub_start(int ncells, const int *__restrict__ start, const int *__restrict__ pos,
                                                  int *__restrict__ start_k)
{
   const int c = blockIdx.x * CB + threadIdx.x;
   if (c <= ncells) start_k[c] = pos[start[c]];          // pos[n] = class size
}

int mdb_build_sublist(mdb_engine *e, int k, cudaStream_t st)
{
   SubList &S = e->sub[k];
   if (S.valid) return 0;
   const int n = e->cfg.nsites, nc = e->ncells, ncols = e->T.nx * e->T.ny;
   const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
   // each class has its own scratch: class 1 is compacted on a second stream beside the class-0 pair pass
   int *flag = e->d_sub_flag + (size_t)k * (n + 1), *pos = e->d_sub_pos + (size_t)k * (n + 1);
   int *scan = e->d_sub_scan + (size_t)k * e->sub_scan_cap, *cols = e->d_sub_cols + (size_t)k * 2 * (ncols + 1);
   k_sub_flags<<<(n + CB - 1) / CB, CB, 0, st>>>(n, e->d_order, e->d_cls, k, flag);
   k_scan_tiles<<<ntiles, SCAN_T, 0, st>>>(flag, n, scan);
   k_scan_sums<<<1, SCAN_T, 0, st>>>(scan, ntiles);
   k_scan_apply<<<ntiles, SCAN_T, 0, st>>>(flag, n, scan, pos);
   k_sub_scatter<<<(n + CB - 1) / CB, CB, 0, st>>>(n, flag, pos, e->d_posq, e->d_sinfo, e->d_order, S.posq, S.sinfo, S.order);
   k_sub_start<<<(nc + 1 + CB - 1) / CB, CB, 0, st>>>(nc, e->d_start, pos, S.start);
   e->launches += 6;
   MDB_CUDA(cudaGetLastError());
   if (launch_batches(e, S.start, cols, scan, S.batches, S.nbatch, st)) return -1;
   S.valid = true;
   return 0;
}

// ---- TOO_CLOSE diagnostics for the pairs the split passes never visit (src/force.c:939-949 looks at every
// pair of the stencil).  Two sites closer than 0.5 A sit in the same or in adjacent cells (mdb_configure
// refuses the split for cells thinner than that), so 27 cells per site are enough.
struct CloseParams { int nx, ny, nz; double reloc[27][3]; };
__global__ void __launch_bounds__(CB) k_too_close_scan(CloseParams P, int s_lo, int n, const double4 *__restrict__ posq,
                                                       const int *__restrict__ scell, const int *__restrict__ start,
                                                       const int *__restrict__ order, const int *__restrict__ mol,
                                                       const unsigned char *__restrict__ cls,
                                                       unsigned long long *__restrict__ counters)
{
   const int s = s_lo + blockIdx.x * CB + threadIdx.x;      // this rank's slice of the sorted sites; partners t > s anywhere
   if (s >= n) return;
   const int c = scell[s], cz = c % P.nz, cy = (c / P.nz) % P.ny, cx = c / (P.nz * P.ny);
   const double4 pi = posq[s];
   const int oi = order[s], mi = mol[oi], ci = cls[oi];
   for (int d = 0; d < 27; d++) {
      int tx = cx + d / 9 - 1, ty = cy + (d / 3) % 3 - 1, tz = cz + d % 3 - 1, ii = 0, jj = 0, kk = 0;
      if (tx < 0) { tx += P.nx; ii = -1; } else if (tx >= P.nx) { tx -= P.nx; ii = 1; }
      if (ty < 0) { ty += P.ny; jj = -1; } else if (ty >= P.ny) { ty -= P.ny; jj = 1; }
      if (tz < 0) { tz += P.nz; kk = -1; } else if (tz >= P.nz) { tz -= P.nz; kk = 1; }
      const int img = 9 * (ii + 1) + 3 * (jj + 1) + (kk + 1), cc = tz + P.nz * (ty + P.ny * tx);
      for (int t = max(start[cc], s + 1); t < start[cc + 1]; t++) {
         const double4 pj = posq[t];
         const double dx = pj.x + P.reloc[img][0] - pi.x, dy = pj.y + P.reloc[img][1] - pi.y, dz = pj.z + P.reloc[img][2] - pi.z;
         if (dx * dx + dy * dy + dz * dz < MDB_TOO_CLOSE) {
            const int oj = order[t], both = ci & cls[oj];
            if (!both && mol[oj] != mi) {             // not seen by the charged or the potential pass
               atomicAdd(&counters[1], 2ULL);
               counters[3] = ((unsigned long long)(unsigned)oi << 32) | (unsigned)oj;
            }
         }
      }
   }
}

int mdb_launch_too_close_scan(mdb_engine *e, cudaStream_t st)
{
   CloseParams P;
   P.nx = e->T.nx; P.ny = e->T.ny; P.nz = e->T.nz;
   for (int k = 0; k < 27; k++)
      for (int a = 0; a < 3; a++) P.reloc[k][a] = e->T.reloc[k][a];
   const long nall = e->cfg.nsites;
   const int s_lo = (int)(nall * e->ithread / e->nthreads), n = (int)(nall * (e->ithread + 1) / e->nthreads);
   if (n <= s_lo) return 0;
   k_too_close_scan<<<(n - s_lo + CB - 1) / CB, CB, 0, st>>>(P, s_lo, n, e->d_posq, e->d_scell, e->d_start, e->d_order, e->d_mol,
                                                      e->d_cls, e->d_counters);
   e->launches += 1;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

int mdb_launch_cells(mdb_engine *e, cudaStream_t st)
{
   const int n = e->cfg.nsites, nc = e->ncells;
   CellParams P;
   for (int i = 0; i < 9; i++) P.hinv[i] = e->T.hinv[i];
   P.nx = e->T.nx; P.ny = e->T.ny; P.nz = e->T.nz; P.ncells = nc;
   P.fnx = P.nx; P.fny = P.ny; P.fnz = P.nz;
   P.eps = 8.0 * 2.220446049250313e-16;                 // 8 * precision(), src/force.c:437
   P.molpbc = e->cfg.molpbc; P.nsites_xf = e->cfg.nsites_xf;
   if (P.molpbc && !e->com_set) { mdb_set_error("molecular-cutoff mode: mdb_set_com_host was not called"); return -1; }
   MDB_CUDA(cudaMemsetAsync(e->d_count, 0, sizeof(int) * (size_t)(nc + 1), st));
   k_cell_ids<<<(n + CB - 1) / CB, CB, 0, st>>>(P, n, e->d_x, e->d_y, e->d_z, e->d_mol, e->d_com, e->d_cell, e->d_count,
                                                e->d_counters);
   int ntiles = (nc + SCAN_TILE - 1) / SCAN_TILE;
   k_scan_tiles<<<ntiles, SCAN_T, 0, st>>>(e->d_count, nc, e->d_scan_tmp);
   k_scan_sums<<<1, SCAN_T, 0, st>>>(e->d_scan_tmp, ntiles);
   k_scan_apply<<<ntiles, SCAN_T, 0, st>>>(e->d_count, nc, e->d_scan_tmp, e->d_start);
   MDB_CUDA(cudaMemsetAsync(e->d_count, 0, sizeof(int) * (size_t)(nc + 1), st));   // now the fill cursor
   k_fill<<<(n + CB - 1) / CB, CB, 0, st>>>(n, e->d_cell, e->d_start, e->d_count, e->d_order);
   k_sort_gather<<<(nc + CB - 1) / CB, CB, 0, st>>>(nc, e->d_start, e->d_order, e->d_x, e->d_y, e->d_z, e->d_chg,
                                                   e->d_type, e->cfg.nsites_xf, e->d_posq, e->d_stype, e->d_scell, e->d_sinfo,
                                                   e->T.nz);
   e->launches += 6;
   MDB_CUDA(cudaGetLastError());
   // batches of the full list: the fused pass, the counting pass and the RDF pass want them; with split
   // passes they are built on demand (mdb_need_batches)
   e->batches_valid = false;
   if (e->pair_mode >= 3 && !e->pair_split && mdb_need_batches(e, st)) return -1;
   e->sub[0].valid = e->sub[1].valid = false;
   e->cells_valid = true;
   return 0;
}
