// mdb_kspace.cu -- reciprocal-space Ewald sum (replaces the k-vector loop of
// ewald(), src/ewald.c:469-583: trig_recur, qsincos, sum, energy/stress, forces).
//
// The reference walks the k-vector list and, per k, makes three passes over
// N-long trig arrays (HBM-bound, N x (hmax+kmax+lmax) doubles of tables).  Here
// the phase factor is kept factorised,
//        exp(i k.r) = E_hk(r) * E_l(r),   E_hk = E_h E_k,
// and a k-vector pair (h,k,+l),(h,k,-l) shares its four real products
//        P1 = sum c_hk C_l   P2 = sum s_hk S_l   P3 = sum s_hk C_l   P4 = sum c_hk S_l
// (C_l = q cos(l c*.r), S_l = q sin(l c*.r)):  C(+-l) = P1 -+ P2,  S(+-l) = P3 +- P4.
//
// Default path (FP64 tensor pipe, further down): k_ktables -> k_sfac_mma -> k_slab_sum -> k_sfin -> k_kforce_mma,
// both heavy kernels as DMMA.8x8x4 GEMMs fed by TMA bulk copies.  The register-operand DFMA kernels below
// (k_sfac, k_kforce) are the earlier formulation, kept selectable with MDB_KSPACE=dfma for A/B timing:
//
//  k_sfac    structure factors: thread owns one (h,k) column and up to 8 l-slots
//            (32 FP64 accumulators), sites stream through shared memory in
//            chunks with their E_h/E_k/E_l power tables built by recurrence.
//            2 DFMA per (site,k-vector).  Per-slab partial sums, no atomics.
//  k_sfin    fixed-order slab reduction -> C,S per k; energy, stress
//            (src/ewald.c:511-553) and the eight back-projection coefficients
//            per l-slot.
//  k_kforce  forces: thread owns a site, walks the (h,k) columns by recurrence
//            and accumulates X,Y,Xz,Yz over the l-slots: 4 DFMA + one complex
//            recurrence step per slot (2 k-vectors); f += k (s X + c Y) ...
//            (src/ewald.c:558-579).
// Framework sites (src/ewald.c:528-579) are handled by keeping their slabs
// separate in k_sfac and using the non-framework-only coefficient set for them.
#include <algorithm>
#include <cstdlib>
#include "mdb_internal.h"

// dynamic shared-memory sizes already granted to the kernels, PER DEVICE (cudaFuncSetAttribute acts on the current
// device only, and a later smaller request must not shrink an earlier grant)
static size_t g_shm_set[64][4];
static int g_max_smem[64];

static constexpr int KT = 256;          // threads per block (k_sfac)
static constexpr int SC = 32;           // sites per shared-memory chunk
static constexpr int LCH = 8;           // l-slots per thread
static constexpr int KF = 128;          // threads per block (k_kforce)

struct SfacArgs {
   KspaceParams K;
   int HKB;                             // shared-memory row stride of the E_hk tile (max columns per block)
   int nvalid, rank, nranks;
   int nslots, slab_sites, n_slabs_nf;
   int nf_lo, nf_hi, fw_lo, fw_hi;      // this rank's ranges in the compacted charged-site list
};

// per-block work descriptor of k_sfac: columns e0..e0+hkb-1 of this rank's share of the
// (nl-sorted) valid list, split over nlc l-chunks of lch slots each (hkb*nlc <= 256)
struct SfacBlock { int e0, hkb, nlc, lch; };

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
   return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.y, b.x, a.x * b.y));
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b)   // a * conj(b)
{
   return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}

__global__ void __launch_bounds__(KT)
k_sfac(SfacArgs A, const SfacBlock *__restrict__ blocks, const int *__restrict__ cidx,
       const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
       const double *__restrict__ chg, const HkDesc *__restrict__ hk, const int *__restrict__ hk_valid,
       double *__restrict__ ppart)
{
   extern __shared__ double2 smem[];
   const KspaceParams &K = A.K;
   const int NL = K.nlslots, NH = K.hmax + 1, NK = K.kmax + 1;
   double2 *sA = smem;                          // [SC][HKB]
   double2 *sB = sA + SC * A.HKB;               // [SC][NL]
   double2 *sH = sB + SC * NL;                  // [SC][NH]
   double2 *sK = sH + SC * NH;                  // [SC][NK]

   const SfacBlock B = blocks[blockIdx.x];
   const int tid = threadIdx.x;
   const int hkl = tid % B.hkb, lc = tid / B.hkb;
   // this thread's (h,k) column: entry of the rank's interleaved share of the valid list
   const int v = A.rank + A.nranks * (B.e0 + hkl);
   const bool have = v < A.nvalid && lc < B.nlc;
   int h = 0, ka = 0, ksgn = 1, nl = 0, slot0 = 0;
   if (v < A.nvalid) {
      const HkDesc d = hk[hk_valid[v]];
      h = d.h; ka = abs(d.k); ksgn = d.k < 0 ? -1 : 1; nl = d.nl; slot0 = d.slot0;
   }
   const int l0 = lc * B.lch;
   const int lcnt = have ? min(max(nl - l0, 0), B.lch) : 0;
   const int wl = __reduce_max_sync(0xffffffffu, lcnt);       // warp-uniform trip count
   const int ngen = KT / B.hkb;                               // threads per column in phase 2

   // slab -> range of the compacted charged-site list (framework sites in their own slabs)
   const int slab = blockIdx.y;
   int s0, s1;
   if (slab < A.n_slabs_nf) {
      s0 = A.nf_lo + slab * A.slab_sites;
      s1 = min(s0 + A.slab_sites, A.nf_hi);
   } else {
      s0 = A.fw_lo + (slab - A.n_slabs_nf) * A.slab_sites;
      s1 = min(s0 + A.slab_sites, A.fw_hi);
   }

   double acc[LCH][4];
#pragma unroll
   for (int j = 0; j < LCH; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;

   for (int base = s0; base < s1; base += SC) {
      // phase 1: power tables of the three base phase factors, one (site,axis) per thread
      if (tid < 3 * SC) {
         const int sl = tid % SC, axis = tid / SC, g = base + sl;
         double q = 0.0, kr = 0.0;
         const double *ks = axis == 0 ? K.astar : axis == 1 ? K.bstar : K.cstar;
         if (g < s1) {
            const int i = cidx[g];
            kr = ks[0] * x[i] + ks[1] * y[i] + ks[2] * z[i];
            q = chg[i];
         }
         double s1v, c1v;
         sincos(kr, &s1v, &c1v);
         const double2 e1 = make_double2(c1v, s1v);
         double2 *tab = axis == 0 ? sH + sl * NH : axis == 1 ? sK + sl * NK : sB + sl * NL;
         const int nmax = axis == 0 ? NH : axis == 1 ? NK : NL;
         const double amp = axis == 2 ? q : 1.0;
         double2 e = make_double2(1.0, 0.0);
         tab[0] = make_double2(amp, 0.0);
         for (int m = 1; m < nmax; m++) {
            e = cmul(e, e1);
            tab[m] = make_double2(amp * e.x, amp * e.y);
         }
      }
      __syncthreads();
      // phase 2: E_hk = E_h * E_k (conjugate for k < 0) for this thread's column
      if (tid < ngen * B.hkb)
         for (int sl = lc; sl < SC; sl += ngen) {
            const double2 eh = sH[sl * NH + h], ek = sK[sl * NK + ka];
            sA[sl * A.HKB + hkl] = ksgn > 0 ? cmul(eh, ek) : cmulc(eh, ek);
         }
      __syncthreads();
      // phase 3: rank-1 updates, 4 DFMA per (site, l-slot)
      if (wl > 0) {
#pragma unroll 2
         for (int sl = 0; sl < SC; sl++) {
            const double2 a = sA[sl * A.HKB + hkl];
            const double2 *bp = sB + sl * NL + l0;
#pragma unroll
            for (int j = 0; j < LCH; j++) {
               if (j < wl) {
                  const double2 b = bp[j];
                  acc[j][0] = fma(a.x, b.x, acc[j][0]);
                  acc[j][1] = fma(a.y, b.y, acc[j][1]);
                  acc[j][2] = fma(a.y, b.x, acc[j][2]);
                  acc[j][3] = fma(a.x, b.y, acc[j][3]);
               }
            }
         }
      }
      __syncthreads();
   }
#pragma unroll
   for (int j = 0; j < LCH; j++)
      if (j < lcnt) {
         double *o = ppart + ((size_t)slab * A.nslots + slot0 + l0 + j) * 4;
         o[0] = acc[j][0]; o[1] = acc[j][1]; o[2] = acc[j][2]; o[3] = acc[j][3];
      }
}

// ---- structure factors on the FP64 tensor pipe (DMMA.8x8x4) -------------------
// The four sums of a column are one real GEMM over the sites:
//    [c_hk ; s_hk](2 x sites) . [C_l , S_l](sites x 2)  =  [P1 P4 ; P3 P2].
// mma.sync.m8n8k4.f64 sustains the nominal FP64 rate on sm_100a (16.3 cycles per warp
// instruction = 256 FMA, scripts/ubench_dmma.cu) with one issue slot and four operand
// registers per 256 FMAs, where the register-operand DFMA form above stops at ~60 %.
//  k_ktables   once per step: per charged site the power tables E_h, E_k (unit modulus) and
//              q E_l, written to HBM in exactly the padded row layout the GEMM kernel wants in
//              shared memory (1.4 KB per site; they are re-read once per column block, mostly from L2).
//  k_sfac_mma  block = 96 columns x one l-range of <= 32 slots, one block per SM: 12 consumer warps of
//              8 columns each (two 8-row m-tiles of 4 columns x {c,s}, NT n-tiles of 4 slots x {C,S},
//              k = 4 sites per instruction) and one producer warp that brings the tables of the next
//              32-site chunk in by three TMA bulk copies (full / empty mbarriers on a double buffer; no
//              block barrier in the loop).  Versions that built the tables in the block had 34 % / 21 % of
//              all warp samples waiting at a barrier.  The A fragment (E_hk of the lane's column and site)
//              is formed in registers from E_h and E_k.  12 warps measured faster than 8, 10, 14, 15, 16.
#ifndef MDB_MCW
#define MDB_MCW 12
#endif
static constexpr int MCW = MDB_MCW;      // consumer warps per block
static constexpr int MC = 8 * MCW;      // columns per block
#ifndef MDB_MSC
#define MDB_MSC 32
#endif
static constexpr int MSC = MDB_MSC;      // sites per shared-memory chunk
static constexpr int MT = 32 * MCW + 32;  // + one producer warp

struct SfacMBlock { int e0, ncols, l0, nt; };

struct KtabLayout {                     // row strides of the per-site tables
   int SB, SH, SK, NLP;                 // sB in doubles (= 4 mod 16), sH/sK in double2 (odd); padded slot count
};

struct SfacMArgs {
   KspaceParams K;
   KtabLayout L;
   int nvalid, rank, nranks;
   int nslots, slab_sites, n_slabs_nf;
   int nf_lo, nf_hi, fw_lo, fw_hi;
   int nrows;                           // rows allocated in the tables (charged sites + MSC of padding)
   int n_slabs, nblk, sgroup;           // 1-D grid order: slab groups of `sgroup` slabs; inside a group column blocks, heaviest first
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
   asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
       : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip_sign(double v, int mask)      // mask = 0 or 0x80000000: no FP64 op
{
   return __hiloint2double(__double2hiint(v) ^ mask, __double2loint(v));
}
// ---- 1-D bulk copies (TMA unit, UBLKCP) completing on an mbarrier
__device__ __forceinline__ void mbar_init(unsigned mb, int count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mb, unsigned bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned mb)
{
   const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(d), "l"(src), "r"(bytes), "r"(mb) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mb, unsigned parity)
{
   asm volatile("{\n\t.reg .pred p;\n\tMBAR_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra MBAR_DONE;\n\tbra MBAR_WAIT;\n\tMBAR_DONE:\n\t}" ::"r"(mb), "r"(parity) : "memory");
}

// power tables of the three base phase factors for rows [g0, g1) of the charged-site list: one (site,axis)
// per thread builds its row in shared memory by recurrence, then the block writes the 32 rows of each
// table as one contiguous, coalesced piece (row-at-a-time 16-byte stores ran at 1.1 ms for 1.08 GB)
static constexpr int KTS = 32;          // sites per block
__global__ void __launch_bounds__(128)
k_ktables(KspaceParams K, KtabLayout L, int g0, int g1, const int *__restrict__ cidx, const double *__restrict__ x,
          const double *__restrict__ y, const double *__restrict__ z, const double *__restrict__ chg,
          double *__restrict__ tE, double2 *__restrict__ tH, double2 *__restrict__ tK)
{
   extern __shared__ double2 smem[];
   const int vE = L.SB / 2;                                     // double2 per row
   double2 *sE = smem, *sH = sE + KTS * vE, *sK = sH + KTS * L.SH;
   const int base = g0 + blockIdx.x * KTS, nv = min(KTS, g1 - base);
   const int tid = threadIdx.x;
   if (tid < 3 * KTS) {
      const int sl = tid % KTS, axis = tid / KTS;
      if (sl < nv) {
         const double *ks = axis == 0 ? K.astar : axis == 1 ? K.bstar : K.cstar;
         const int i = cidx[base + sl];
         const double kr = ks[0] * x[i] + ks[1] * y[i] + ks[2] * z[i];
         const double amp = axis == 2 ? chg[i] : 1.0;
         double s1v, c1v;
         sincos(kr, &s1v, &c1v);
         const double2 e1 = make_double2(c1v, s1v);
         double2 *tab = axis == 0 ? sH + sl * L.SH : axis == 1 ? sK + sl * L.SK : sE + sl * vE;
         const int nmax = axis == 0 ? K.hmax + 1 : axis == 1 ? K.kmax + 1 : K.nlslots;
         const int nrow = axis == 0 ? L.SH : axis == 1 ? L.SK : vE;
         double2 e = make_double2(1.0, 0.0);
         tab[0] = make_double2(amp, 0.0);
         for (int m = 1; m < nmax; m++) {
            e = cmul(e, e1);
            tab[m] = make_double2(amp * e.x, amp * e.y);
         }
         for (int m = nmax; m < nrow; m++) tab[m] = make_double2(0.0, 0.0);
      }
   }
   __syncthreads();
   double2 *gE = reinterpret_cast<double2 *>(tE + (size_t)base * L.SB), *gH = tH + (size_t)base * L.SH, *gK = tK + (size_t)base * L.SK;
   for (int u = tid; u < nv * vE; u += 128) gE[u] = sE[u];
   for (int u = tid; u < nv * L.SH; u += 128) gH[u] = sH[u];
   for (int u = tid; u < nv * L.SK; u += 128) gK[u] = sK[u];
}

template <int NT>
__device__ __forceinline__ void sfac_mma_body(const SfacMArgs &A, const SfacMBlock &B, const double *__restrict__ tE,
                                              const double2 *__restrict__ tH, const double2 *__restrict__ tK,
                                              const HkDesc *__restrict__ hk, const int *__restrict__ hk_valid,
                                              double *__restrict__ ppart, double2 *smem, const int slab)
{
   const KtabLayout &L = A.L;
   const size_t buf_doubles = (size_t)MSC * L.SB + 2 * (size_t)MSC * (L.SH + L.SK);
   double *buf0 = reinterpret_cast<double *>(smem);
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int g = lane >> 2, kq = lane & 3, comp = g & 1;

   int s0, s1;
   if (slab < A.n_slabs_nf) {
      s0 = A.nf_lo + slab * A.slab_sites;
      s1 = min(s0 + A.slab_sites, A.nf_hi);
   } else {
      s0 = A.fw_lo + (slab - A.n_slabs_nf) * A.slab_sites;
      s1 = min(s0 + A.slab_sites, A.fw_hi);
   }

   int ch[2] = {0, 0}, ck[2] = {0, 0}, cnl[2] = {0, 0}, cslot[2] = {0, 0}, csign[2] = {0, 0};
#pragma unroll
   for (int mt = 0; mt < 2; mt++) {
      const int cl = warp * 8 + mt * 4 + (g >> 1);
      const int v = A.rank + A.nranks * (B.e0 + cl);
      bool neg = false;
      if (cl < B.ncols && v < A.nvalid) {
         const HkDesc d = hk[hk_valid[v]];
         ch[mt] = d.h; ck[mt] = abs(d.k); neg = d.k < 0; cnl[mt] = d.nl; cslot[mt] = d.slot0;
      }
      // c = eh.x ek.x - sg eh.y ek.y    s = eh.y ek.x + sg eh.x ek.y    (sg = -1 for k < 0)
      csign[mt] = ((comp == 0) != neg) ? (int)0x80000000 : 0;
   }
   const bool warp_on = warp * 8 < B.ncols;

   double acc[2][NT][2];
#pragma unroll
   for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int n = 0; n < NT; n++) acc[mt][n][0] = acc[mt][n][1] = 0.0;

   auto bufB = [&](int b) { return buf0 + b * buf_doubles; };
   auto bufH = [&](int b) { return reinterpret_cast<double2 *>(buf0 + b * buf_doubles + (size_t)MSC * L.SB); };
   auto bufK = [&](int b) { return reinterpret_cast<double2 *>(buf0 + b * buf_doubles + (size_t)MSC * L.SB) + (size_t)MSC * L.SH; };
   // chunk [base, base+MSC) of the table rows -> buffer b: three bulk copies issued by one thread, completing on
   // the buffer's mbarrier (as 16-byte cp.async pieces the staging cost 0.5 ms of LSU time per launch).  Rows past
   // the slab end belong to other slabs (or the zeroed padding): their q E_l row is zeroed instead of copied,
   // E_h/E_k are copied as they are (finite).
   __shared__ __align__(8) unsigned long long mbar_store[4];         // full[2], empty[2]
   const unsigned mb0 = (unsigned)__cvta_generic_to_shared(mbar_store);
   if (tid == 0) {
      mbar_init(mb0, 1); mbar_init(mb0 + 8, 1);
      mbar_init(mb0 + 16, MCW); mbar_init(mb0 + 24, MCW);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncthreads();
   const int nchunks = (s1 - s0 + MSC - 1) / MSC;
   if (warp == MCW) {
      // producer warp: chunk c -> buffer c & 1 once every consumer warp has released it
      const int vE = L.SB / 2;                                        // 16-byte units per E row
      for (int c = 0; c < nchunks; c++) {
         const int b = c & 1, base = s0 + c * MSC, nv = min(MSC, s1 - base);
         if (lane == 0) mbar_wait(mb0 + 16 + 8 * b, ((c >> 1) & 1) ^ 1);
         __syncwarp();
         double2 *dE = reinterpret_cast<double2 *>(bufB(b));
         for (int u = nv * vE + lane; u < MSC * vE; u += 32) dE[u] = make_double2(0.0, 0.0);
         __syncwarp();
         if (lane == 0) {
            const unsigned bE = (unsigned)nv * L.SB * 8u, bH = (unsigned)MSC * L.SH * 16u, bK = (unsigned)MSC * L.SK * 16u;
            const unsigned mb = mb0 + 8u * b;
            mbar_expect_tx(mb, bE + bH + bK);                         // (an arrive: releases the zero stores above)
            bulk_g2s(dE, tE + (size_t)base * L.SB, bE, mb);
            bulk_g2s(bufH(b), tH + (size_t)base * L.SH, bH, mb);
            bulk_g2s(bufK(b), tK + (size_t)base * L.SK, bK, mb);
         }
      }
      return;
   }
   for (int c = 0; c < nchunks; c++) {
      const int b = c & 1;
      mbar_wait(mb0 + 8u * b, (c >> 1) & 1);
      if (warp_on) {
         const double *sB = bufB(b) + 2 * B.l0 + g;
         const double2 *sH = bufH(b), *sK = bufK(b);
         auto make_a = [&](int t, double (&a)[2]) {
            const int sl = 4 * t + kq;
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
               const double2 eh = sH[sl * L.SH + ch[mt]];
               const double2 ek = sK[sl * L.SK + ck[mt]];
               const double p = comp ? eh.y : eh.x, r = comp ? eh.x : eh.y;
               a[mt] = fma(p, ek.x, flip_sign(r * ek.y, csign[mt]));
            }
         };
#pragma unroll 2
         for (int t = 0; t < MSC / 4; t++) {
            double a[2];
            make_a(t, a);
            const double *bp = sB + (4 * t + kq) * L.SB;
#pragma unroll
            for (int n = 0; n < NT; n++) {
               const double bv = bp[8 * n];
               dmma884(acc[0][n], a[0], bv);
               dmma884(acc[1][n], a[1], bv);
            }
         }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb0 + 16 + 8 * b) : "memory");
   }
   // C fragment: row g = (column g/2, c|s), columns 2 kq + e = (slot kq of the n-tile, C|S)
   if (warp_on)
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
         for (int n = 0; n < NT; n++) {
            const int l = B.l0 + 4 * n + kq;
            if (l < cnl[mt]) {
               double *o = ppart + ((size_t)slab * A.nslots + cslot[mt] + l) * 4;
               // (c,C) = P1 -> 0   (c,S) = P4 -> 3   (s,C) = P3 -> 2   (s,S) = P2 -> 1
               o[comp ? 2 : 0] = acc[mt][n][0];
               o[comp ? 1 : 3] = acc[mt][n][1];
            }
         }
}

__global__ void __launch_bounds__(MT, 1)
k_sfac_mma(SfacMArgs A, const SfacMBlock *__restrict__ blocks, const double *__restrict__ tE,
           const double2 *__restrict__ tH, const double2 *__restrict__ tK, const HkDesc *__restrict__ hk,
           const int *__restrict__ hk_valid, double *__restrict__ ppart)
{
   extern __shared__ double2 smem[];
   // 1-D grid, ordered [slab group][column block][slab of the group]: inside a group of `sgroup` slabs the column
   // blocks with the most slots (the list is sorted) come first, so the tail of the grid is made of the cheapest
   // blocks, and the blocks that run at the same time share the table chunks of one or two slab groups through L2.
   // Measured (1.024 M sites, MDB_SFAC_GROUP): groups of 4 / 8 / 16 / 32 / all slabs = 4.87 / 4.88 / 4.71 / 4.69 / 4.57 ms
   // with 1.3 / 1.8 / 3.3 / 6.0 / 10.2 GB read from HBM (one global heavy-first order re-reads the 1 GB of tables once
   // per column block); column-block-fastest order: 4.79 ms, 1.1 GB.  Default 16.
   const int per_group = A.sgroup * A.nblk, sg = blockIdx.x / per_group, rem = blockIdx.x - sg * per_group;
   const int gsz = min(A.sgroup, A.n_slabs - sg * A.sgroup);
   const int cb = rem / gsz, slab = sg * A.sgroup + rem - cb * gsz;
   if (cb >= A.nblk) return;                                      // (short last group)
   const SfacMBlock B = blocks[cb];
   switch (B.nt) {          // block-uniform: straight-line DMMA sequences, accumulators in registers
      case 1: sfac_mma_body<1>(A, B, tE, tH, tK, hk, hk_valid, ppart, smem, slab); break;
      case 2: sfac_mma_body<2>(A, B, tE, tH, tK, hk, hk_valid, ppart, smem, slab); break;
      case 3: sfac_mma_body<3>(A, B, tE, tH, tK, hk, hk_valid, ppart, smem, slab); break;
      case 4: sfac_mma_body<4>(A, B, tE, tH, tK, hk, hk_valid, ppart, smem, slab); break;
      case 5: sfac_mma_body<5>(A, B, tE, tH, tK, hk, hk_valid, ppart, smem, slab); break;
      case 6: sfac_mma_body<6>(A, B, tE, tH, tK, hk, hk_valid, ppart, smem, slab); break;
      case 7: sfac_mma_body<7>(A, B, tE, tH, tK, hk, hk_valid, ppart, smem, slab); break;
      default: sfac_mma_body<8>(A, B, tE, tH, tK, hk, hk_valid, ppart, smem, slab); break;
   }
}

// ---- fixed-order sum of the slab partials: psum[0] = non-framework slabs, psum[1] = framework slabs.
// This 8*nslots block is what ranks all-reduce when k-space is partitioned by sites.
__global__ void __launch_bounds__(256) k_slab_sum(int nslots, int n_slabs, int n_slabs_nf,
                                                  const double *__restrict__ ppart, double *__restrict__ psum)
{
   const int t = blockIdx.x * 256 + threadIdx.x;            // one thread per (slot, component)
   if (t >= nslots * 4) return;
   double a = 0.0, b = 0.0;
   for (int sb = 0; sb < n_slabs_nf; sb++) a += ppart[(size_t)sb * nslots * 4 + t];
   for (int sb = n_slabs_nf; sb < n_slabs; sb++) b += ppart[(size_t)sb * nslots * 4 + t];
   psum[t] = a;
   psum[(size_t)nslots * 4 + t] = b;
}

// ---- per k-vector: energy, stress, back-projection coefficients --------------
struct SfinArgs {
   KspaceParams K;
   int nslots, n_slabs, n_slabs_nf, rank, nranks, framework;
   int mma;                // coefficients go to k_kforce_mma's padded group blocks (slot_dst), in its order:
                           // plane 0 {X,Y} x {el.x, el.y}, plane 1 {Xz,Yz} x {el.x, el.y}
   int plane;              // double2 between the two planes of a block
};

__global__ void __launch_bounds__(256)
k_sfin(SfinArgs A, const HkDesc *__restrict__ hk, const int *__restrict__ slot_hk,
       const int *__restrict__ slot_flags, const double *__restrict__ ppart, double *__restrict__ coef_tot,
       double *__restrict__ coef_nf, double *__restrict__ kpartials, const int *__restrict__ slot_dst,
       double2 *__restrict__ blk_tot, double2 *__restrict__ blk_nf)
{
   const KspaceParams &K = A.K;
   const int slot = blockIdx.x * 256 + threadIdx.x;
   double red[7] = {0, 0, 0, 0, 0, 0, 0};
   if (slot < A.nslots) {
      const HkDesc d = hk[slot_hk[slot]];
      const bool mine = (d.pad % A.nranks) == A.rank;          // pad = position in the valid list
      double ct[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cn[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (mine) {
         const int l = slot - d.slot0, flags = slot_flags[slot];
         double pn[4] = {0, 0, 0, 0}, pf[4] = {0, 0, 0, 0};
         for (int sb = 0; sb < A.n_slabs; sb++) {
            const double *p = ppart + ((size_t)sb * A.nslots + slot) * 4;
            double *dst = sb < A.n_slabs_nf ? pn : pf;
            dst[0] += p[0]; dst[1] += p[1]; dst[2] += p[2]; dst[3] += p[3];
         }
         double Ct[2] = {0, 0}, St[2] = {0, 0}, Cn2[2] = {0, 0}, Sn2[2] = {0, 0};
#pragma unroll
         for (int sg = 0; sg < 2; sg++) {
            if (!(flags & (1 << sg))) continue;
            const double sgn = sg == 0 ? 1.0 : -1.0;
            const double Cn = pn[0] - sgn * pn[1], Sn = pn[2] + sgn * pn[3];
            const double Cf = pf[0] - sgn * pf[1], Sf = pf[2] + sgn * pf[3];
            const double kx = d.kx, ky = d.ky, kz = d.kzt + (sg == 0 ? l : -l) * K.cz2;
            const double ksq = kx * kx + ky * ky + kz * kz;
            const double coeff = K.pref * exp(ksq * K.r4alpha) / ksq;
            const double coeff2 = 2.0 * (1.0 - ksq * K.r4alpha) / ksq;
            const double pe_k = 0.5 * coeff * (Cn * (Cn + Cf + Cf) + Sn * (Sn + Sf + Sf));
            red[0] += pe_k;
            red[1] += pe_k - pe_k * coeff2 * kx * kx;
            red[2] -= pe_k * coeff2 * kx * ky;
            red[3] -= pe_k * coeff2 * kx * kz;
            red[4] += pe_k - pe_k * coeff2 * ky * ky;
            red[5] -= pe_k * coeff2 * ky * kz;
            red[6] += pe_k - pe_k * coeff2 * kz * kz;
            Ct[sg] = coeff * (Cn + Cf); St[sg] = coeff * (Sn + Sf);
            Cn2[sg] = coeff * Cn;       Sn2[sg] = coeff * Sn;
         }
         const double fl = (double)l;
         ct[0] = Ct[0] + Ct[1]; ct[1] = St[0] + St[1]; ct[2] = Ct[0] - Ct[1]; ct[3] = St[0] - St[1];
         ct[4] = fl * ct[2]; ct[5] = fl * ct[1]; ct[6] = fl * ct[0]; ct[7] = fl * ct[3];
         cn[0] = Cn2[0] + Cn2[1]; cn[1] = Sn2[0] + Sn2[1]; cn[2] = Cn2[0] - Cn2[1]; cn[3] = Sn2[0] - Sn2[1];
         cn[4] = fl * cn[2]; cn[5] = fl * cn[1]; cn[6] = fl * cn[0]; cn[7] = fl * cn[3];
      }
      if (A.mma) {
         // X = sum el.x ct0 + el.y ct3   Y = sum -el.x ct1 + el.y ct2   Xz = sum el.x ct4 + el.y ct5   Yz = sum -el.x ct7 + el.y ct6
         const int dst = slot_dst[slot];
         if (dst >= 0) {
            blk_tot[dst] = make_double2(ct[0], -ct[1]); blk_tot[dst + 1] = make_double2(ct[3], ct[2]);
            blk_tot[dst + A.plane] = make_double2(ct[4], -ct[7]); blk_tot[dst + A.plane + 1] = make_double2(ct[5], ct[6]);
            if (A.framework) {
               blk_nf[dst] = make_double2(cn[0], -cn[1]); blk_nf[dst + 1] = make_double2(cn[3], cn[2]);
               blk_nf[dst + A.plane] = make_double2(cn[4], -cn[7]); blk_nf[dst + A.plane + 1] = make_double2(cn[5], cn[6]);
            }
         }
      } else {
#pragma unroll
         for (int k = 0; k < 8; k++) coef_tot[(size_t)slot * 8 + k] = ct[k];
         if (A.framework)
#pragma unroll
            for (int k = 0; k < 8; k++) coef_nf[(size_t)slot * 8 + k] = cn[k];
      }
   }
   __shared__ double sm[8][7];
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
   for (int k = 0; k < 7; k++) {
      double t = red[k];
#pragma unroll
      for (int dd = 16; dd > 0; dd >>= 1) t += __shfl_xor_sync(0xffffffffu, t, dd);
      if (lane == 0) sm[w][k] = t;
   }
   __syncthreads();
   if (threadIdx.x < 7) {
      double t = 0;
      for (int k = 0; k < 8; k++) t += sm[k][threadIdx.x];
      kpartials[(size_t)blockIdx.x * 8 + threadIdx.x] = t;
   }
}

__global__ void __launch_bounds__(256) k_recip_finish(const double *__restrict__ kpartials, int nblocks, int nsites,
                                                      double *__restrict__ out)
{
   __shared__ double sm[256];
   double acc[7] = {0, 0, 0, 0, 0, 0, 0};
   for (int b = threadIdx.x; b < nblocks; b += 256)
#pragma unroll
      for (int k = 0; k < 7; k++) acc[k] += kpartials[(size_t)b * 8 + k];
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
      sc[1] += tot[0];
      sc[2 + 0] += tot[1]; sc[2 + 1] += tot[2]; sc[2 + 2] += tot[3];
      sc[2 + 4] += tot[4]; sc[2 + 5] += tot[5]; sc[2 + 8] += tot[6];
   }
}

// ---- forces -------------------------------------------------------------------
// One thread per CHARGED site (cidx: compacted list, uncharged sites get no k-space
// force).  The (h,k) descriptors and their back-projection coefficients are the same
// for every site, so the block stages them through shared memory in batches
// (coalesced copy, broadcast reads) instead of each warp fetching them from L2.
struct KfArgs {
   KspaceParams K;
   int c0, c1;            // range in the compacted charged-site list
   int nhk, rank, nranks;
   int hb;                // descriptors per shared-memory batch
   int max_slots;         // capacity of the coefficient stage (slots)
};

// NS sites per thread: the coefficient loads (8 broadcast LDS.128 per slot pair) are shared by NS
// sites, which takes the shared-memory pipe off the critical path (at NS = 1 it is as busy as FP64)
// and lets consecutive DFMAs share the coefficient operand (.reuse): a DFMA that reads three
// distinct vector registers issues every 3.2 cycles on sm_100a, 2.1 otherwise (scripts/ubench2.cu).
// Reading the coefficients through the uniform datapath instead (__constant__ chunks, LDCU -> UR
// operands) was measured slower: LDCU.128 sustains only ~1 per 6 cycles per SM.
template <int NS>
__global__ void __launch_bounds__(KF, NS == 1 ? 4 : NS == 2 ? 3 : 2)
k_kforce(KfArgs A, const int *__restrict__ cidx, const double *__restrict__ x, const double *__restrict__ y,
         const double *__restrict__ z, const double *__restrict__ chg, const HkDesc *__restrict__ hk,
         const double *__restrict__ coef, double *__restrict__ out)
{
   extern __shared__ double2 smem[];
   double4 *s_coef = reinterpret_cast<double4 *>(smem);                        // [max_slots][2]
   HkDesc *s_hk = reinterpret_cast<HkDesc *>(s_coef + 2 * (size_t)A.max_slots); // [hb]
   const KspaceParams &K = A.K;
   int i[NS];
   bool active[NS];
   double q[NS];
   double2 ea[NS], eb[NS], ec[NS], eh[NS], ehk[NS];
   double fx[NS], fy[NS], fz[NS];
#pragma unroll
   for (int s = 0; s < NS; s++) {
      const int t = A.c0 + (blockIdx.x * NS + s) * KF + threadIdx.x;
      active[s] = t < A.c1;
      i[s] = 0;
      double xi = 0, yi = 0, zi = 0;
      q[s] = 0.0;
      if (active[s]) {
         i[s] = cidx[t];
         q[s] = chg[i[s]]; xi = x[i[s]]; yi = y[i[s]]; zi = z[i[s]];
      }
      sincos(K.astar[0] * xi + K.astar[1] * yi + K.astar[2] * zi, &ea[s].y, &ea[s].x);
      sincos(K.bstar[0] * xi + K.bstar[1] * yi + K.bstar[2] * zi, &eb[s].y, &eb[s].x);
      sincos(K.cstar[0] * xi + K.cstar[1] * yi + K.cstar[2] * zi, &ec[s].y, &ec[s].x);
      eh[s] = make_double2(1.0, 0.0); ehk[s] = eh[s];
      fx[s] = fy[s] = fz[s] = 0.0;
   }

   for (int c0 = 0; c0 < A.nhk; c0 += A.hb) {
      const int nb = min(A.hb, A.nhk - c0);
      __syncthreads();
      if (threadIdx.x < nb) {
         HkDesc d = hk[c0 + threadIdx.x];
         if (d.nl > 0 && (d.pad % A.nranks) != A.rank) d.nl = -d.nl;     // not ours: keep slot count, skip work
         s_hk[threadIdx.x] = d;
      }
      __syncthreads();
      // slots of a batch are contiguous (slot0 grows along the traversal)
      int slot_lo = 0, nsl = 0;
      {
         const HkDesc &f = s_hk[0], &l = s_hk[nb - 1];
         slot_lo = f.slot0;
         nsl = l.slot0 + abs(l.nl) - slot_lo;
      }
      const double4 *src = reinterpret_cast<const double4 *>(coef) + 2 * (size_t)slot_lo;
      for (int k = threadIdx.x; k < 2 * nsl; k += KF) s_coef[k] = src[k];
      __syncthreads();
      // two (h,k) columns per pass: they share the l-recurrence of E_l (4 of the 12 FP64 ops per
      // slot) and give the scheduler 8 NS independent accumulator chains
      for (int c = 0; c < nb; c += 2) {
         const HkDesc &d1 = s_hk[c];
         const HkDesc &d2 = s_hk[min(c + 1, nb - 1)];
         const bool two = c + 1 < nb;
         double2 ehk1[NS], ehk2[NS];
#pragma unroll
         for (int s = 0; s < NS; s++) {
            switch (d1.code) {
               case HK_NEWH:   if (d1.h > 0) eh[s] = cmul(eh[s], ea[s]); ehk[s] = eh[s]; break;
               case HK_KUP:    ehk[s] = cmul(ehk[s], eb[s]); break;
               case HK_KDOWN0: ehk[s] = cmulc(eh[s], eb[s]); break;
               default:        ehk[s] = cmulc(ehk[s], eb[s]); break;
            }
            ehk1[s] = ehk[s];
            if (two) {
               switch (d2.code) {
                  case HK_NEWH:   if (d2.h > 0) eh[s] = cmul(eh[s], ea[s]); ehk[s] = eh[s]; break;
                  case HK_KUP:    ehk[s] = cmul(ehk[s], eb[s]); break;
                  case HK_KDOWN0: ehk[s] = cmulc(eh[s], eb[s]); break;
                  default:        ehk[s] = cmulc(ehk[s], eb[s]); break;
               }
            }
            ehk2[s] = ehk[s];
         }
         const int nl1 = max(d1.nl, 0), nl2 = two ? max(d2.nl, 0) : 0;
         if (nl1 + nl2 == 0) continue;
         const double4 *cf1 = s_coef + 2 * (d1.slot0 - slot_lo);
         const double4 *cf2 = s_coef + 2 * (d2.slot0 - slot_lo);
         double2 el[NS];
         double X1[NS], Y1[NS], Xz1[NS], Yz1[NS], X2[NS], Y2[NS], Xz2[NS], Yz2[NS];
#pragma unroll
         for (int s = 0; s < NS; s++) {
            el[s] = make_double2(q[s], 0.0);
            X1[s] = Y1[s] = Xz1[s] = Yz1[s] = X2[s] = Y2[s] = Xz2[s] = Yz2[s] = 0.0;
         }
         const int nj = min(nl1, nl2);
         int l = 0;
#pragma unroll 2
         for (; l < nj; l++) {
            const double4 a1 = cf1[2 * l], b1 = cf1[2 * l + 1], a2 = cf2[2 * l], b2 = cf2[2 * l + 1];
#pragma unroll
            for (int s = 0; s < NS; s++) {
               X1[s] = fma(el[s].x, a1.x, fma(el[s].y, a1.w, X1[s]));
               Y1[s] = fma(el[s].y, a1.z, fma(-el[s].x, a1.y, Y1[s]));
               Xz1[s] = fma(el[s].x, b1.x, fma(el[s].y, b1.y, Xz1[s]));
               Yz1[s] = fma(el[s].y, b1.z, fma(-el[s].x, b1.w, Yz1[s]));
               X2[s] = fma(el[s].x, a2.x, fma(el[s].y, a2.w, X2[s]));
               Y2[s] = fma(el[s].y, a2.z, fma(-el[s].x, a2.y, Y2[s]));
               Xz2[s] = fma(el[s].x, b2.x, fma(el[s].y, b2.y, Xz2[s]));
               Yz2[s] = fma(el[s].y, b2.z, fma(-el[s].x, b2.w, Yz2[s]));
               el[s] = cmul(el[s], ec[s]);
            }
         }
         for (; l < nl1; l++) {
            const double4 a1 = cf1[2 * l], b1 = cf1[2 * l + 1];
#pragma unroll
            for (int s = 0; s < NS; s++) {
               X1[s] = fma(el[s].x, a1.x, fma(el[s].y, a1.w, X1[s]));
               Y1[s] = fma(el[s].y, a1.z, fma(-el[s].x, a1.y, Y1[s]));
               Xz1[s] = fma(el[s].x, b1.x, fma(el[s].y, b1.y, Xz1[s]));
               Yz1[s] = fma(el[s].y, b1.z, fma(-el[s].x, b1.w, Yz1[s]));
               el[s] = cmul(el[s], ec[s]);
            }
         }
         for (; l < nl2; l++) {
            const double4 a2 = cf2[2 * l], b2 = cf2[2 * l + 1];
#pragma unroll
            for (int s = 0; s < NS; s++) {
               X2[s] = fma(el[s].x, a2.x, fma(el[s].y, a2.w, X2[s]));
               Y2[s] = fma(el[s].y, a2.z, fma(-el[s].x, a2.y, Y2[s]));
               Xz2[s] = fma(el[s].x, b2.x, fma(el[s].y, b2.y, Xz2[s]));
               Yz2[s] = fma(el[s].y, b2.z, fma(-el[s].x, b2.w, Yz2[s]));
               el[s] = cmul(el[s], ec[s]);
            }
         }
#pragma unroll
         for (int s = 0; s < NS; s++) {
            if (nl1 > 0) {
               const double T = fma(ehk1[s].y, X1[s], ehk1[s].x * Y1[s]), Tz = fma(ehk1[s].y, Xz1[s], ehk1[s].x * Yz1[s]);
               fx[s] = fma(d1.kx, T, fx[s]);
               fy[s] = fma(d1.ky, T, fy[s]);
               fz[s] = fma(d1.kzt, T, fma(K.cz2, Tz, fz[s]));
            }
            if (nl2 > 0) {
               const double T = fma(ehk2[s].y, X2[s], ehk2[s].x * Y2[s]), Tz = fma(ehk2[s].y, Xz2[s], ehk2[s].x * Yz2[s]);
               fx[s] = fma(d2.kx, T, fx[s]);
               fy[s] = fma(d2.ky, T, fy[s]);
               fz[s] = fma(d2.kzt, T, fma(K.cz2, Tz, fz[s]));
            }
         }
      }
   }
#pragma unroll
   for (int s = 0; s < NS; s++)
      if (active[s]) {
         out[i[s]] += fx[s];
         out[(size_t)K.nsites + i[s]] += fy[s];
         out[2 * (size_t)K.nsites + i[s]] += fz[s];
      }
}

// ---- forces on the FP64 tensor pipe --------------------------------------------------
// For a site j and a column c the four l-sums of k_kforce are one real GEMM over the slots,
//    [X Y Xz Yz](j,c) = sum_(l,comp) E_l(j)[l,comp] . coef(c)[l,comp][X Y Xz Yz],
// m = 8 sites, n = 8 columns, k = 2 slots x {cos,sin} per DMMA.8x8x4, one accumulator tile per
// quantity.  The C fragment leaves a thread with ONE site and TWO columns, so the back-projection
//    T = s X + c Y,  Tz = s Xz + c Yz,   Fa += h T,  Fb += k T,  Fc += Tz,   F = a* Fa + b* Fb + c* Fc
// stays thread-local (k.x = h a*.x + k b*.x, k.z adds l c*.z which is folded into Xz, Yz);
// lanes are summed once at the end.  Columns come in groups of 8 in descending slot count (the valid
// list of k_sfac), so a group runs max(nl)/2 steps with 3 % padding; E_hk = E_h E_k is read from per-site
// power tables in shared memory, which is what limits a block to 128 sites (one block per SM).
// Warps 0..NSB-1 own 16 sites each and walk the even groups, warps NSB..2NSB-1 the same sites and
// the odd groups: the coefficient fragments (the same for every site) are then loaded once per 16
// sites and four warps per scheduler keep the DMMA pipe fed.
// k_sfin stores the coefficients group-major and zero-padded ("blocks": descriptor + two planes of 8
// columns), so a group arrives by ONE TMA bulk copy issued by a producer warp into a two-deep ring per
// half; consumers wait on the stage's `full` mbarrier and release it (`empty`) after their DMMA loop,
// before the back-projection.  No block-wide barrier in the loop: the warps of a scheduler drift apart
// and one warp's epilogue overlaps the others' DMMAs (with a barrier per round and cp.async staging by the
// consumers the kernel lost ~2 200 of ~7 100 cycles per round: 9.72 ms).
struct KfGroup {            // 8 columns, padded with nl = 0
   int nsteps, pad0, pad1, pad2;
   int h[8], k[8], nl[8], slot0[8];
};
static constexpr int KFB_HDR = 16;      // double2 units in front of the planes (descriptor, 256 bytes)

struct KfMArgs {
   KspaceParams K;
   int c0, c1;              // range in the compacted charged-site list
   int ngroups, nsb;        // nsb: site-warps per block (block = 2 nsb consumer warps + 1 producer warp, 16 nsb sites)
   int SE, SH, SK, LPS;     // table strides: E_l in doubles (= 4 mod 16), E_h/E_k in double2 (odd); stage slots (= 2 mod 4)
   int blk2;                // double2 per coefficient block: KFB_HDR + 2 planes x 8 columns x 2 LPS
};

__global__ void __launch_bounds__(544, 1)
k_kforce_mma(KfMArgs A, const int *__restrict__ cidx, const double *__restrict__ x, const double *__restrict__ y,
             const double *__restrict__ z, const double *__restrict__ chg, const double2 *__restrict__ blocks,
             double *__restrict__ out)
{
   extern __shared__ double2 smem[];
   const KspaceParams &K = A.K;
   const int nsites_b = 16 * A.nsb, ncons = 64 * A.nsb, nthreads = ncons + 32;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int g = lane >> 2, kq = lane & 3;
   const bool producer = warp >= 2 * A.nsb;
   const int half = producer ? 0 : warp / A.nsb, sw = warp % A.nsb;
   // shared memory: E_l [sites][SE] doubles | E_h [sites][SH] | E_k [sites][SK] | ring [2 halves][2 stages][blk2] double2 | F [2][sites][3] | mbarriers
   double *sE = reinterpret_cast<double *>(smem);
   double2 *sH = reinterpret_cast<double2 *>(sE + (size_t)nsites_b * A.SE);
   double2 *sK = sH + (size_t)nsites_b * A.SH;
   double2 *sC = sK + (size_t)nsites_b * A.SK;
   const int colstride = 2 * A.LPS;                      // double2 per (plane, column)
   double *sF = reinterpret_cast<double *>(sC + 4 * (size_t)A.blk2);
   const unsigned mb0 = (unsigned)__cvta_generic_to_shared(sF + 2 * (size_t)nsites_b * 3);
   // full[half][stage] at mb0 + 8 (2 half + stage), empty[half][stage] 32 bytes further
   if (tid == 0) {
      for (int k = 0; k < 4; k++) { mbar_init(mb0 + 8 * k, 1); mbar_init(mb0 + 32 + 8 * k, A.nsb); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }

   const int site0 = A.c0 + blockIdx.x * nsites_b;
   // ---- power tables, one (site, axis) per thread
   for (int job = tid; job < 3 * nsites_b; job += nthreads) {
      const int sl = job % nsites_b, axis = job / nsites_b, t = site0 + sl;
      double q = 0.0, kr = 0.0;
      const double *ks = axis == 0 ? K.astar : axis == 1 ? K.bstar : K.cstar;
      if (t < A.c1) {
         const int i = cidx[t];
         kr = ks[0] * x[i] + ks[1] * y[i] + ks[2] * z[i];
         q = chg[i];
      }
      double s1v, c1v;
      sincos(kr, &s1v, &c1v);
      const double2 e1 = make_double2(c1v, s1v);
      double2 *tab = axis == 0 ? sH + (size_t)sl * A.SH : axis == 1 ? sK + (size_t)sl * A.SK
                                                                    : reinterpret_cast<double2 *>(sE + (size_t)sl * A.SE);
      const int nmax = axis == 0 ? K.hmax + 1 : axis == 1 ? K.kmax + 1 : K.nlslots;
      const double amp = axis == 2 ? q : 1.0;
      double2 e = make_double2(1.0, 0.0);
      tab[0] = make_double2(amp, 0.0);
      for (int m = 1; m < nmax; m++) {
         e = cmul(e, e1);
         tab[m] = make_double2(amp * e.x, amp * e.y);
      }
      if (axis == 2)
         for (int m = nmax; m < A.SE / 2; m++) tab[m] = make_double2(0.0, 0.0);
   }
   __syncthreads();                                      // tables and mbarriers are ready

   const int nrounds = (A.ngroups + 1) / 2;
   double Fa[2] = {0, 0}, Fb[2] = {0, 0}, Fc[2] = {0, 0};
   const int sl0 = sw * 16 + g;                          // this thread's site of tile 0 (tile 1: +8)
   if (producer) {
      if (lane == 0) {
         const unsigned bytes = (unsigned)A.blk2 * 16u;
         for (int r = 0; r < nrounds; r++)
            for (int hf = 0; hf < 2; hf++) {
               const int gi = 2 * r + hf, st = r & 1, slot = 2 * hf + st;
               if (gi >= A.ngroups) continue;
               mbar_wait(mb0 + 32 + 8 * slot, ((r >> 1) & 1) ^ 1);     // stage free (passes at once the first time round)
               mbar_expect_tx(mb0 + 8 * slot, bytes);
               bulk_g2s(sC + (size_t)slot * A.blk2, blocks + (size_t)gi * A.blk2, bytes, mb0 + 8 * slot);
            }
      }
   } else {
      const double *ea0 = sE + (size_t)sl0 * A.SE + kq, *ea1 = ea0 + 8 * (size_t)A.SE;
      for (int r = 0; r < nrounds; r++) {
         const int gi = 2 * r + half, st = r & 1, slot = 2 * half + st;
         if (gi >= A.ngroups) break;
         mbar_wait(mb0 + 8 * slot, (r >> 1) & 1);
         const double2 *blk = sC + (size_t)slot * A.blk2;
         const KfGroup &G = *reinterpret_cast<const KfGroup *>(blk);
         const int nsteps = G.nsteps;
         const int hc0 = G.h[2 * kq], kc0 = G.k[2 * kq], hc1 = G.h[2 * kq + 1], kc1 = G.k[2 * kq + 1];
         const double2 *c0 = blk + KFB_HDR + g * colstride + kq;
         const double2 *c1 = c0 + 8 * colstride;
         double acc[2][4][2];
#pragma unroll
         for (int s = 0; s < 2; s++)
#pragma unroll
            for (int qn = 0; qn < 4; qn++) acc[s][qn][0] = acc[s][qn][1] = 0.0;
#pragma unroll 2
         for (int t = 0; t < nsteps; t++) {
            const double2 q01 = c0[4 * t], q23 = c1[4 * t];
            const double a0 = ea0[4 * t], a1 = ea1[4 * t];
            dmma884(acc[0][0], a0, q01.x); dmma884(acc[0][1], a0, q01.y);
            dmma884(acc[0][2], a0, q23.x); dmma884(acc[0][3], a0, q23.y);
            dmma884(acc[1][0], a1, q01.x); dmma884(acc[1][1], a1, q01.y);
            dmma884(acc[1][2], a1, q23.x); dmma884(acc[1][3], a1, q23.y);
         }
         __syncwarp();
         if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb0 + 32 + 8 * slot) : "memory");
         // back-projection: C fragment = (site g of the tile, columns 2 kq + e)
#pragma unroll
         for (int e = 0; e < 2; e++) {
            const int hc = e ? hc1 : hc0, kc = e ? kc1 : kc0;
            const double fh = (double)hc, fk = (double)kc;
            const int ka = abs(kc), sg = kc < 0 ? (int)0x80000000 : 0;
#pragma unroll
            for (int s = 0; s < 2; s++) {
               const int sl = sl0 + 8 * s;
               const double2 eh = sH[(size_t)sl * A.SH + hc];
               double2 ek = sK[(size_t)sl * A.SK + ka];
               ek.y = flip_sign(ek.y, sg);
               const double2 ehk = cmul(eh, ek);
               const double T = fma(ehk.y, acc[s][0][e], ehk.x * acc[s][1][e]);
               Fa[s] = fma(fh, T, Fa[s]);
               Fb[s] = fma(fk, T, Fb[s]);
               Fc[s] = fma(ehk.y, acc[s][2][e], fma(ehk.x, acc[s][3][e], Fc[s]));
            }
         }
      }
   }
   // ---- sum the four lanes of a site, then the two halves, and add to the caller's force rows
   if (!producer)
#pragma unroll
      for (int s = 0; s < 2; s++) {
#pragma unroll
         for (int d = 1; d <= 2; d <<= 1) {
            Fa[s] += __shfl_xor_sync(0xffffffffu, Fa[s], d);
            Fb[s] += __shfl_xor_sync(0xffffffffu, Fb[s], d);
            Fc[s] += __shfl_xor_sync(0xffffffffu, Fc[s], d);
         }
         if (kq == 0) {
            double *f = sF + ((size_t)half * nsites_b + sl0 + 8 * s) * 3;
            f[0] = Fa[s]; f[1] = Fb[s]; f[2] = Fc[s];
         }
      }
   __syncthreads();
   for (int sl = tid; sl < nsites_b; sl += nthreads) {
      const int t = site0 + sl;
      if (t >= A.c1) continue;
      const double *f0 = sF + (size_t)sl * 3, *f1 = f0 + (size_t)nsites_b * 3;
      const double fa = f0[0] + f1[0], fb = f0[1] + f1[1], fc = f0[2] + f1[2];
      const int i = cidx[t];
      out[i] += K.astar[0] * fa + K.bstar[0] * fb + K.cstar[0] * fc;
      out[(size_t)K.nsites + i] += K.astar[1] * fa + K.bstar[1] * fb + K.cstar[1] * fc;
      out[2 * (size_t)K.nsites + i] += K.astar[2] * fa + K.bstar[2] * fb + K.cstar[2] * fc;
   }
}

#ifndef MDB_KF_NS
#define MDB_KF_NS 2
#endif

// How one rank's share of the k-space work is cut.
//  column mode (Moldy's own scheme, src/ewald.c:495-496): all sites x every P-th (h,k) column;
//              no exchange before the final force sum, but the per-site set-up is replicated.
//  site mode   (the manual's "RIL" scheme, src/moldy.tex:3441-3466): own sites x all columns; the
//              8*nslots structure-factor sums are all-reduced between the two passes; both
//              passes then scale with 1/P.
// MDB_KSPACE=dfma selects the register-operand DFMA kernels (k_sfac, k_kforce) for A/B timing;
// the default is the DMMA pair (k_sfac_mma, k_kforce_mma).  Both are device paths.
static bool kspace_use_mma()
{
   const char *m = getenv("MDB_KSPACE");           // read per call: the tests switch it between engines
   return !(m && std::string(m) == "dfma");
}

struct RecipPlan {
   int col_rank, col_nranks;
   int nf_lo, nf_hi, fw_lo, fw_hi;
   int slab_sites, n_slabs_nf, n_slabs;
   bool add_scalars;
};

static void kspace_params(const mdb_engine *e, KspaceParams &K)
{
   const mdb_config &c = e->cfg;
   const HostTables &T = e->T;
   for (int a = 0; a < 3; a++) { K.astar[a] = T.astar[a]; K.bstar[a] = T.bstar[a]; K.cstar[a] = T.cstar[a]; }
   K.cz2 = T.cstar[2];
   K.r4alpha = -1.0 / (4.0 * c.alpha * c.alpha);
   K.pref = 2.0 / (MDB_EPS0 * T.vol);
   K.hmax = T.hmax; K.kmax = T.kmax; K.lmax = T.lmax; K.nlslots = T.lmax + 1;
   K.nsites = c.nsites; K.nsites_xf = c.nsites_xf;
}

static int make_plan(mdb_engine *e, bool by_sites, RecipPlan &P, cudaStream_t st)
{
   const HostTables &T = e->T;
   const int nvalid = (int)T.hk_valid.size();
   const int r = e->ithread, np = e->nthreads;
   if (by_sites) {
      P.col_rank = 0; P.col_nranks = 1;
      const long nf = e->n_charged_nf, fw = e->n_charged - e->n_charged_nf;
      P.nf_lo = (int)(nf * r / np); P.nf_hi = (int)(nf * (r + 1) / np);
      P.fw_lo = e->n_charged_nf + (int)(fw * r / np); P.fw_hi = e->n_charged_nf + (int)(fw * (r + 1) / np);
      P.add_scalars = r == 0;
   } else {
      P.col_rank = r; P.col_nranks = np;
      P.nf_lo = 0; P.nf_hi = e->n_charged_nf; P.fw_lo = e->n_charged_nf; P.fw_hi = e->n_charged;
      P.add_scalars = true;
   }
   // per-block column table for this column partition
   const int mode = kspace_use_mma() ? 1 : 0;
   if (e->sfac_rank != P.col_rank || e->sfac_nranks != P.col_nranks || e->sfac_mode != mode || !e->d_sfac_blocks) {
      std::vector<SfacBlock> blocks;          // SfacMBlock has the same four ints
      const int my_cols = nvalid > P.col_rank ? (nvalid - P.col_rank + P.col_nranks - 1) / P.col_nranks : 0;
      int e0 = 0;
      while (e0 < my_cols) {
         const int nlmax = T.hk[T.hk_valid[P.col_rank + P.col_nranks * e0]].nl;
         if (mode == 1) {
            const int ncols = std::min(MC, my_cols - e0);
            for (int l0 = 0; l0 < nlmax; l0 += 32)
               blocks.push_back({e0, ncols, l0, (std::min(32, nlmax - l0) + 3) / 4});
            e0 += ncols;
            continue;
         }
         int nlc = std::max(2, (nlmax + LCH - 1) / LCH);
         if (nlc > 8) { mdb_set_error("k_cutoff gives lmax > 63: not supported by this build of k_sfac"); return -1; }
         const int hkb = KT / nlc;
         blocks.push_back({e0, hkb, nlc, (nlmax + nlc - 1) / nlc});
         e0 += hkb;
      }
      if (e->d_sfac_blocks) { cudaFree(e->d_sfac_blocks); e->d_sfac_blocks = nullptr; }
      e->n_sfac_blocks = (int)blocks.size();
      if (!blocks.empty()) {
         MDB_CUDA(cudaMalloc(&e->d_sfac_blocks, sizeof(SfacBlock) * blocks.size()));
         MDB_CUDA(cudaMemcpyAsync(e->d_sfac_blocks, blocks.data(), sizeof(SfacBlock) * blocks.size(),
                                  cudaMemcpyHostToDevice, st));
         MDB_CUDA(cudaStreamSynchronize(st));
      }
      e->sfac_rank = P.col_rank; e->sfac_nranks = P.col_nranks; e->sfac_mode = mode;
   }
   // site slabs: enough blocks for about two waves on 148 SMs x 2 resident blocks
   const int own = (P.nf_hi - P.nf_lo) + (P.fw_hi - P.fw_lo);
   int want = std::max(1, (4 * 148 + std::max(1, e->n_sfac_blocks) - 1) / std::max(1, e->n_sfac_blocks));
   if (mode == 1) {
      // k_sfac_mma: one block per SM; never a few blocks more than a whole number of waves
      static int waves = getenv("MDB_SFAC_WAVES") ? atoi(getenv("MDB_SFAC_WAVES")) : 8;
      want = std::max(1, waves * 148 / std::max(1, e->n_sfac_blocks));
   }
   int slab = (own + want - 1) / want;
   slab = std::max(SC, ((slab + SC - 1) / SC) * SC);
   P.slab_sites = slab;
   P.n_slabs_nf = (P.nf_hi - P.nf_lo + slab - 1) / slab;
   P.n_slabs = P.n_slabs_nf + (P.fw_hi - P.fw_lo + slab - 1) / slab;
   const size_t pp = (size_t)std::max(1, P.n_slabs) * T.nslots * 4;
   if (pp > e->ppart_cap) {
      if (e->d_ppart) cudaFree(e->d_ppart);
      MDB_CUDA(cudaMalloc(&e->d_ppart, sizeof(double) * pp));
      e->ppart_cap = pp;
   }
   return 0;
}

// pass 1: structure-factor sums of this rank's share -> psum[2][nslots][4]
static int recip_partial(mdb_engine *e, const RecipPlan &P, double *d_psum, cudaStream_t st)
{
   const HostTables &T = e->T;
   SfacArgs A;
   kspace_params(e, A.K);
   A.HKB = KT / 2;
   A.nvalid = (int)T.hk_valid.size(); A.rank = P.col_rank; A.nranks = P.col_nranks;
   A.nslots = T.nslots; A.slab_sites = P.slab_sites; A.n_slabs_nf = P.n_slabs_nf;
   A.nf_lo = P.nf_lo; A.nf_hi = P.nf_hi; A.fw_lo = P.fw_lo; A.fw_hi = P.fw_hi;
   const size_t shm = sizeof(double2) * (size_t)SC * (A.HKB + A.K.nlslots + A.K.hmax + 1 + A.K.kmax + 1);
   if (P.col_nranks > 1)      // slots of other ranks' columns are never written: keep them defined
      MDB_CUDA(cudaMemsetAsync(e->d_ppart, 0, sizeof(double) * (size_t)std::max(1, P.n_slabs) * T.nslots * 4, st));
   if (e->n_sfac_blocks > 0 && P.n_slabs > 0 && kspace_use_mma()) {
      SfacMArgs M;
      M.K = A.K;
      M.nvalid = A.nvalid; M.rank = A.rank; M.nranks = A.nranks;
      M.nslots = A.nslots; M.slab_sites = A.slab_sites; M.n_slabs_nf = A.n_slabs_nf;
      M.nf_lo = A.nf_lo; M.nf_hi = A.nf_hi; M.fw_lo = A.fw_lo; M.fw_hi = A.fw_hi;
      M.L.NLP = (A.K.nlslots + 3) / 4 * 4;
      M.L.SB = 2 * M.L.NLP; while (M.L.SB % 16 != 4) M.L.SB += 2;
      M.L.SH = (A.K.hmax + 1) | 1; M.L.SK = (A.K.kmax + 1) | 1;
      M.nrows = e->n_charged + MSC;
      // per-site power tables (zeroed once: rows a rank never computes must stay finite)
      const size_t row_bytes = sizeof(double) * M.L.SB + sizeof(double2) * (M.L.SH + M.L.SK);
      const size_t need = row_bytes * (size_t)M.nrows;
      if (need > e->ktab_cap) {
         if (e->d_ktab) cudaFree(e->d_ktab);
         e->d_ktab = nullptr; e->ktab_cap = 0;
         MDB_CUDA(cudaMalloc(&e->d_ktab, need));
         MDB_CUDA(cudaMemsetAsync(e->d_ktab, 0, need, st));
         e->ktab_cap = need;
      }
      double *tE = reinterpret_cast<double *>(e->d_ktab);
      double2 *tH = reinterpret_cast<double2 *>(tE + (size_t)M.nrows * M.L.SB);
      double2 *tK = tH + (size_t)M.nrows * M.L.SH;
      size_t &tshm_set = g_shm_set[e->device & 63][0];
      if (KTS * row_bytes > 48 * 1024 && KTS * row_bytes > tshm_set) {
         MDB_CUDA(cudaFuncSetAttribute(k_ktables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(KTS * row_bytes)));
         tshm_set = KTS * row_bytes;
      }
      for (int part = 0; part < 2; part++) {
         const int g0 = part == 0 ? P.nf_lo : P.fw_lo, g1 = part == 0 ? P.nf_hi : P.fw_hi;
         if (g1 <= g0) continue;
         k_ktables<<<(g1 - g0 + KTS - 1) / KTS, 128, KTS * row_bytes, st>>>(M.K, M.L, g0, g1, e->d_cidx, e->d_x, e->d_y, e->d_z,
                                                                        e->d_chg, tE, tH, tK);
         e->launches++;
      }
      const size_t mshm = 2 * (size_t)MSC * row_bytes;
      size_t &mshm_set = g_shm_set[e->device & 63][1];
      if (mshm > mshm_set) {
         MDB_CUDA(cudaFuncSetAttribute(k_sfac_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mshm));
         // largest shared-memory carve-out: the pair kernel's filler blocks share the SM (mdb_force_both)
         MDB_CUDA(cudaFuncSetAttribute(k_sfac_mma, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
         mshm_set = mshm;
      }
      static const int sgroup = getenv("MDB_SFAC_GROUP") ? std::max(1, atoi(getenv("MDB_SFAC_GROUP"))) : 16;
      M.n_slabs = P.n_slabs; M.nblk = e->n_sfac_blocks; M.sgroup = std::min(sgroup, P.n_slabs);
      const int ngrp = (P.n_slabs + M.sgroup - 1) / M.sgroup;
      dim3 g(ngrp * M.sgroup * M.nblk);
      k_sfac_mma<<<g, MT, mshm, st>>>(M, (const SfacMBlock *)e->d_sfac_blocks, tE, tH, tK, e->d_hk, e->d_hk_valid,
                                      e->d_ppart);
      e->launches++;
   } else if (e->n_sfac_blocks > 0 && P.n_slabs > 0) {
      size_t &shm_set = g_shm_set[e->device & 63][2];
      if (shm > shm_set) {
         MDB_CUDA(cudaFuncSetAttribute(k_sfac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
         shm_set = shm;
      }
      dim3 g(e->n_sfac_blocks, P.n_slabs);
      k_sfac<<<g, KT, shm, st>>>(A, (const SfacBlock *)e->d_sfac_blocks, e->d_cidx, e->d_x, e->d_y, e->d_z, e->d_chg,
                                 e->d_hk, e->d_hk_valid, e->d_ppart);
      e->launches++;
   }
   k_slab_sum<<<(T.nslots * 4 + 255) / 256, 256, 0, st>>>(T.nslots, P.n_slabs, P.n_slabs_nf, e->d_ppart, d_psum);
   e->launches++;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

// pass 2: per-k energy/stress/coefficients from the (complete) sums, then forces on this rank's sites
static int recip_finish(mdb_engine *e, const RecipPlan &P, const double *d_psum, double *d_out, cudaStream_t st)
{
   const mdb_config &c = e->cfg;
   e->kf_nch = 0;                                   // (slices of the force kernel: set again below when it is cut)
   const HostTables &T = e->T;
   SfinArgs F;
   kspace_params(e, F.K);
   F.nslots = T.nslots; F.n_slabs = 2; F.n_slabs_nf = 1;
   F.rank = P.col_rank; F.nranks = P.col_nranks; F.framework = c.nsites_xf < c.nsites;
   F.mma = kspace_use_mma() ? 1 : 0;
   const int fb = (T.nslots + 255) / 256;
   KfMArgs Q{};
   size_t kshm = 0;
   if (F.mma) {
      // groups of 8 columns of this rank's share of the valid list (descending slot count); their coefficient
      // blocks [descriptor | plane 0 | plane 1], zero-padded, and the slot -> block entry map k_sfin writes through
      Q.K = F.K;
      const int nle = (F.K.nlslots + 1) / 2 * 2;
      Q.SE = 2 * nle; while (Q.SE % 16 != 4) Q.SE += 2;
      Q.SH = (F.K.hmax + 1) | 1; Q.SK = (F.K.kmax + 1) | 1;
      Q.LPS = nle; while (Q.LPS % 4 != 2) Q.LPS += 2;
      const int colstride = 2 * Q.LPS;
      Q.blk2 = KFB_HDR + 2 * 8 * colstride;
      F.plane = 8 * colstride;
      if (e->kf_rank != P.col_rank || e->kf_nranks != P.col_nranks || !e->d_kf_groups) {
         const int nvalid = (int)T.hk_valid.size();
         std::vector<KfGroup> groups;
         std::vector<int> slot_dst((size_t)std::max(T.nslots, 1), -1);
         for (int v = P.col_rank; v < nvalid; v += 8 * P.col_nranks) {
            KfGroup G{};
            int mx = 0;
            const size_t base = groups.size() * (size_t)Q.blk2 + KFB_HDR;
            for (int j = 0; j < 8; j++) {
               const int vv = v + j * P.col_nranks;
               if (vv >= nvalid) break;
               const HkDesc &d = T.hk[T.hk_valid[vv]];
               G.h[j] = d.h; G.k[j] = d.k; G.nl[j] = d.nl; G.slot0[j] = d.slot0;
               mx = std::max(mx, d.nl);
               for (int l = 0; l < d.nl; l++) slot_dst[d.slot0 + l] = (int)(base + (size_t)j * colstride + 2 * l);
            }
            G.nsteps = (mx + 1) / 2;
            groups.push_back(G);
         }
         if (e->d_kf_groups) { cudaFree(e->d_kf_groups); e->d_kf_groups = nullptr; }
         if (e->d_kf_slot_dst) { cudaFree(e->d_kf_slot_dst); e->d_kf_slot_dst = nullptr; }
         e->n_kf_groups = (int)groups.size();
         const size_t nb2 = (size_t)std::max(e->n_kf_groups, 1) * Q.blk2;
         // both coefficient sets (total | non-framework) in one allocation, zeroed once; headers = descriptors
         MDB_CUDA(cudaMalloc(&e->d_kf_groups, sizeof(double2) * 2 * nb2));
         MDB_CUDA(cudaMemsetAsync(e->d_kf_groups, 0, sizeof(double2) * 2 * nb2, st));
         for (int set = 0; set < 2 && !groups.empty(); set++)
            MDB_CUDA(cudaMemcpy2DAsync((double2 *)e->d_kf_groups + set * nb2, sizeof(double2) * Q.blk2, groups.data(), sizeof(KfGroup),
                                       sizeof(KfGroup), groups.size(), cudaMemcpyHostToDevice, st));
         MDB_CUDA(cudaMalloc(&e->d_kf_slot_dst, sizeof(int) * slot_dst.size()));
         MDB_CUDA(cudaMemcpyAsync(e->d_kf_slot_dst, slot_dst.data(), sizeof(int) * slot_dst.size(), cudaMemcpyHostToDevice, st));
         MDB_CUDA(cudaStreamSynchronize(st));
         e->kf_rank = P.col_rank; e->kf_nranks = P.col_nranks;
      }
      Q.ngroups = e->n_kf_groups;
   }
   const size_t nb2 = (size_t)std::max(e->n_kf_groups, 1) * (size_t)Q.blk2;
   double2 *blk_tot = (double2 *)e->d_kf_groups, *blk_nf = blk_tot ? blk_tot + nb2 : nullptr;
   k_sfin<<<fb, 256, 0, st>>>(F, e->d_hk, e->d_slot_flags + T.nslots, e->d_slot_flags, d_psum, e->d_coef_tot,
                              e->d_coef_nf, e->d_kpartials, e->d_kf_slot_dst, blk_tot, blk_nf);
   e->launches++;
   if (P.add_scalars) {
      k_recip_finish<<<1, 256, 0, st>>>(e->d_kpartials, fb, c.nsites, d_out);
      e->launches++;
   }
   if (F.mma) {
      const size_t per_site = sizeof(double) * Q.SE + sizeof(double2) * (Q.SH + Q.SK) + 6 * sizeof(double);
      const size_t stage_bytes = 4 * (size_t)Q.blk2 * sizeof(double2) + 64;
      int &max_smem = g_max_smem[e->device & 63];
      if (!max_smem) MDB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device));
      Q.nsb = 8;
      while (Q.nsb > 1 && stage_bytes + 16 * (size_t)Q.nsb * per_site > (size_t)max_smem) Q.nsb--;
      {
         // few waves (a rank's share of the sites at 8 GPUs: 750 blocks = 5.07 waves of 148): pick the block
         // size whose last wave is fullest, cost = waves x sites per block
         int nsm = 148;
         cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, e->device);
         const long nown = std::max(P.nf_hi - P.nf_lo, P.fw_hi - P.fw_lo);
         long best = -1;
         int best_nsb = Q.nsb;
         // a block lasts as long as its busiest scheduler: the 2 nb consumer warps sit ceil(2 nb / 4) deep on the four
         // DMMA pipes (measured at 8 ranks: nb = 7 takes as long per block as nb = 8, 192 us)
         for (int nb = Q.nsb; nb >= std::max(4, Q.nsb - 4) && nown > 0; nb--) {
            const long blocks = (nown + 16 * nb - 1) / (16 * nb), waves = (blocks + nsm - 1) / nsm;
            if (waves > 16) break;                         // many waves: the tail does not matter
            const long cost = waves * ((2 * nb + 3) / 4);
            if (best < 0 || cost < best) { best = cost; best_nsb = nb; }
         }
         if (getenv("MDB_KF_NSB")) best_nsb = std::min(Q.nsb, std::max(1, atoi(getenv("MDB_KF_NSB"))));
         Q.nsb = best_nsb;
      }
      kshm = stage_bytes + 16 * (size_t)Q.nsb * per_site;
      if (kshm > (size_t)max_smem) { mdb_set_error("k_cutoff too large for the shared-memory tables of k_kforce_mma"); return -1; }
      size_t &kshm_set = g_shm_set[e->device & 63][3];
      if (kshm > kshm_set) {
         MDB_CUDA(cudaFuncSetAttribute(k_kforce_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kshm));
         MDB_CUDA(cudaFuncSetAttribute(k_kforce_mma, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
         kshm_set = kshm;
      }
      // the non-framework sites in kf_chunks slices of whole blocks, an event after each (kf_site_hi: the forces of the
      // ORIGINAL sites below it are complete then -- d_cidx ascends); the framework sites go with the last slice
      const int nch = std::max(1, std::min(e->kf_chunks, (int)mdb_engine::KF_MAXCH));
      const bool events = nch > 1 && P.col_nranks == 1 && (int)e->h_cidx.size() == e->n_charged && Q.ngroups > 0;
      e->kf_nch = 0;
      for (int part = 0; part < 2; part++) {
         const int c0 = part == 0 ? P.nf_lo : P.fw_lo, c1 = part == 0 ? P.nf_hi : P.fw_hi;
         if (c1 <= c0 || Q.ngroups == 0) continue;
         const int spb = 16 * Q.nsb, nblk = (c1 - c0 + spb - 1) / spb, parts = part == 0 && events ? nch : 1;
         for (int k = 0; k < parts; k++) {
            const long b_lo = (long)nblk * k / parts, b_hi = (long)nblk * (k + 1) / parts;
            if (b_hi <= b_lo) continue;
            Q.c0 = c0 + (int)b_lo * spb; Q.c1 = std::min(c1, c0 + (int)b_hi * spb);
            k_kforce_mma<<<(unsigned)(b_hi - b_lo), 64 * Q.nsb + 32, kshm, st>>>(
               Q, e->d_cidx, e->d_x, e->d_y, e->d_z, e->d_chg, part == 0 ? blk_tot : blk_nf, d_out);
            e->launches++;
            if (events && part == 0 && k + 1 < parts) {
               if (!e->kf_ev[e->kf_nch]) MDB_CUDA(cudaEventCreateWithFlags(&e->kf_ev[e->kf_nch], cudaEventDisableTiming));
               MDB_CUDA(cudaEventRecord(e->kf_ev[e->kf_nch], st));
               e->kf_site_hi[e->kf_nch] = Q.c1 < e->n_charged ? e->h_cidx[Q.c1] : c.nsites;
               e->kf_nch++;
            }
         }
      }
      if (events) {                                      // the last slice (and the framework sites): everything is complete
         if (!e->kf_ev[e->kf_nch]) MDB_CUDA(cudaEventCreateWithFlags(&e->kf_ev[e->kf_nch], cudaEventDisableTiming));
         MDB_CUDA(cudaEventRecord(e->kf_ev[e->kf_nch], st));
         e->kf_site_hi[e->kf_nch] = c.nsites;
         e->kf_nch++;
      }
      MDB_CUDA(cudaGetLastError());
      return 0;
   }
   KfArgs Qd;
   Qd.K = F.K; Qd.nhk = (int)T.hk.size(); Qd.rank = P.col_rank; Qd.nranks = P.col_nranks;
   Qd.hb = std::max(1, std::min(16, 512 / F.K.nlslots));
   Qd.max_slots = Qd.hb * F.K.nlslots;
   kshm = sizeof(double4) * 2 * (size_t)Qd.max_slots + sizeof(HkDesc) * (size_t)Qd.hb;
   if (P.nf_hi > P.nf_lo) {
      Qd.c0 = P.nf_lo; Qd.c1 = P.nf_hi;
      k_kforce<MDB_KF_NS><<<(Qd.c1 - Qd.c0 + KF * MDB_KF_NS - 1) / (KF * MDB_KF_NS), KF, kshm, st>>>(
         Qd, e->d_cidx, e->d_x, e->d_y, e->d_z, e->d_chg, e->d_hk, e->d_coef_tot, d_out);
      e->launches++;
   }
   if (P.fw_hi > P.fw_lo) {
      Qd.c0 = P.fw_lo; Qd.c1 = P.fw_hi;
      k_kforce<MDB_KF_NS><<<(Qd.c1 - Qd.c0 + KF * MDB_KF_NS - 1) / (KF * MDB_KF_NS), KF, kshm, st>>>(
         Qd, e->d_cidx, e->d_x, e->d_y, e->d_z, e->d_chg, e->d_hk, e->d_coef_nf, d_out);
      e->launches++;
   }
   MDB_CUDA(cudaGetLastError());
   return 0;
}

// one call, no exchange: column partition (or everything, for a single rank)
int mdb_launch_recip(mdb_engine *e, double *d_out, cudaStream_t st)
{
   if (e->T.hk_valid.empty()) return 0;
   RecipPlan P;
   if (make_plan(e, false, P, st)) return -1;
   if (recip_partial(e, P, e->d_psum, st)) return -1;
   return recip_finish(e, P, e->d_psum, d_out, st);
}

// split form for the site partition: the caller all-reduces d_psum between the two calls
int mdb_launch_recip_partial(mdb_engine *e, double *d_psum, cudaStream_t st)
{
   if (e->T.hk_valid.empty()) return 0;
   RecipPlan P;
   if (make_plan(e, true, P, st)) return -1;
   return recip_partial(e, P, d_psum, st);
}
int mdb_launch_recip_finish(mdb_engine *e, const double *d_psum, double *d_out, cudaStream_t st)
{
   if (e->T.hk_valid.empty()) return 0;
   RecipPlan P;
   if (make_plan(e, true, P, st)) return -1;
   return recip_finish(e, P, d_psum, d_out, st);
}

// Flop the two GEMM kernels execute per call for this rank's share of the charged sites (site partition; with one rank:
// everything): 512 = 2 x 8 x 8 x 4 per DMMA.8x8x4.
//   k_sfac_mma: block = column block (<= MC columns, 8 per warp) x l-range (nt n-tiles of 4 slots); per 4 sites every
//               active warp issues 2 m-tiles x nt n-tiles;
//   k_kforce_mma: group of 8 columns, nsteps = ceil(max nl / 2) k-steps, 4 accumulator tiles (X, Y, Xz, Yz) per 8 sites.
// Padding (sites to whole chunks, columns to whole warps, slots to whole tiles) is counted: the pipe executes it.
extern "C" double mdb_recip_gemm_flop(const mdb_engine *e)
{
   const HostTables &T = e->T;
   const int nvalid = (int)T.hk_valid.size();
   if (nvalid == 0) return 0.0;
   const int r = e->ithread, np = e->nthreads;
   const long nf = e->n_charged_nf, fw = e->n_charged - e->n_charged_nf;
   const long own_nf = nf * (r + 1) / np - nf * r / np, own_fw = fw * (r + 1) / np - fw * r / np;
   auto pad = [](long v, long m) { return (v + m - 1) / m * m; };
   double sfac = 0.0, kf = 0.0;
   for (int e0 = 0; e0 < nvalid; e0 += MC) {
      const int ncols = std::min(MC, nvalid - e0), nlmax = T.hk[T.hk_valid[e0]].nl;
      const int warps = (ncols + 7) / 8;
      for (int l0 = 0; l0 < nlmax; l0 += 32) {
         const int nt = (std::min(32, nlmax - l0) + 3) / 4;
         sfac += 512.0 * warps * 2 * nt * (double)(pad(own_nf, MSC) + pad(own_fw, MSC)) / 4.0;
      }
   }
   for (int v = 0; v < nvalid; v += 8) {
      int mx = 0;
      for (int j = 0; j < 8 && v + j < nvalid; j++) mx = std::max(mx, T.hk[T.hk_valid[v + j]].nl);
      kf += 512.0 * 4 * ((mx + 1) / 2) * (double)(pad(own_nf, 16) + pad(own_fw, 16)) / 8.0;
   }
   return sfac + kf;
}
