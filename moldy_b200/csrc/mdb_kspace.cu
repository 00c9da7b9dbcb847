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
#include "mdb_internal.h"

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
};

__global__ void __launch_bounds__(256)
k_sfin(SfinArgs A, const HkDesc *__restrict__ hk, const int *__restrict__ slot_hk,
       const int *__restrict__ slot_flags, const double *__restrict__ ppart, double *__restrict__ coef_tot,
       double *__restrict__ coef_nf, double *__restrict__ kpartials)
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
#pragma unroll
      for (int k = 0; k < 8; k++) coef_tot[(size_t)slot * 8 + k] = ct[k];
      if (A.framework)
#pragma unroll
         for (int k = 0; k < 8; k++) coef_nf[(size_t)slot * 8 + k] = cn[k];
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

#ifndef MDB_KF_NS
#define MDB_KF_NS 2
#endif

// How one rank's share of the k-space work is cut.
//  column mode (Moldy's own scheme, src/ewald.c:495-496): all sites x every P-th (h,k) column;
//              no exchange before the final force sum, but the per-site set-up is replicated.
//  site mode   (the manual's "RIL" scheme, src/moldy.tex:3441-3466): own sites x all columns; the
//              8*nslots structure-factor sums are all-reduced between the two passes; both
//              passes then scale with 1/P.
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
   if (e->sfac_rank != P.col_rank || e->sfac_nranks != P.col_nranks || !e->d_sfac_blocks) {
      std::vector<SfacBlock> blocks;
      const int my_cols = nvalid > P.col_rank ? (nvalid - P.col_rank + P.col_nranks - 1) / P.col_nranks : 0;
      int e0 = 0;
      while (e0 < my_cols) {
         const int nlmax = T.hk[T.hk_valid[P.col_rank + P.col_nranks * e0]].nl;
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
      e->sfac_rank = P.col_rank; e->sfac_nranks = P.col_nranks;
   }
   // site slabs: enough blocks for about two waves on 148 SMs x 2 resident blocks
   const int own = (P.nf_hi - P.nf_lo) + (P.fw_hi - P.fw_lo);
   const int want = std::max(1, (4 * 148 + std::max(1, e->n_sfac_blocks) - 1) / std::max(1, e->n_sfac_blocks));
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
   if (e->n_sfac_blocks > 0 && P.n_slabs > 0) {
      static size_t shm_set = 0;
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
   const HostTables &T = e->T;
   SfinArgs F;
   kspace_params(e, F.K);
   F.nslots = T.nslots; F.n_slabs = 2; F.n_slabs_nf = 1;
   F.rank = P.col_rank; F.nranks = P.col_nranks; F.framework = c.nsites_xf < c.nsites;
   const int fb = (T.nslots + 255) / 256;
   k_sfin<<<fb, 256, 0, st>>>(F, e->d_hk, e->d_slot_flags + T.nslots, e->d_slot_flags, d_psum, e->d_coef_tot,
                              e->d_coef_nf, e->d_kpartials);
   e->launches++;
   if (P.add_scalars) {
      k_recip_finish<<<1, 256, 0, st>>>(e->d_kpartials, fb, c.nsites, d_out);
      e->launches++;
   }
   KfArgs Q;
   Q.K = F.K; Q.nhk = (int)T.hk.size(); Q.rank = P.col_rank; Q.nranks = P.col_nranks;
   Q.hb = std::max(1, std::min(16, 512 / F.K.nlslots));
   Q.max_slots = Q.hb * F.K.nlslots;
   const size_t kshm = sizeof(double4) * 2 * (size_t)Q.max_slots + sizeof(HkDesc) * (size_t)Q.hb;
   if (P.nf_hi > P.nf_lo) {
      Q.c0 = P.nf_lo; Q.c1 = P.nf_hi;
      k_kforce<MDB_KF_NS><<<(Q.c1 - Q.c0 + KF * MDB_KF_NS - 1) / (KF * MDB_KF_NS), KF, kshm, st>>>(
         Q, e->d_cidx, e->d_x, e->d_y, e->d_z, e->d_chg, e->d_hk, e->d_coef_tot, d_out);
      e->launches++;
   }
   if (P.fw_hi > P.fw_lo) {
      Q.c0 = P.fw_lo; Q.c1 = P.fw_hi;
      k_kforce<MDB_KF_NS><<<(Q.c1 - Q.c0 + KF * MDB_KF_NS - 1) / (KF * MDB_KF_NS), KF, kshm, st>>>(
         Q, e->d_cidx, e->d_x, e->d_y, e->d_z, e->d_chg, e->d_hk, e->d_coef_nf, d_out);
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
