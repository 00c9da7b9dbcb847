// mdb_math.cuh -- FP64 building blocks of the pair kernel.
//
// The pair loop is bound by the FP64 pipe (64 DFMA/clk/SM on sm_100), so the
// transcendental pieces are written as short DFMA chains around the MUFU
// seeds instead of calling libdevice (whose sqrt / division / exp carry
// slow-path branches and IEEE rounding we do not need: 1e-13 relative is
// plenty for the 1e-10 force tolerance, we keep ~2e-16).
#pragma once
#include <cuda_runtime.h>

#ifndef MDB_LIBM_MATH
#define MDB_LIBM_MATH 0        /* 1: fall back to libdevice (debugging) */
#endif
#ifndef MDB_EXP_F32TAIL
#define MDB_EXP_F32TAIL 0      /* tiled pair kernel's exp: 0 = 64-entry table + degree-5 polynomial in FP64 (10 FP64 instructions);
                                  1 = 256-entry table + FP32 Taylor tail (6 FP64 + 12 integer/FP32 instructions, 2e-13): measured
                                  SLOWER on B200 (TIP4P Coulomb pass 18.75 -> 19.17 ms, quartz 35.5 -> 36.3): kept for A/B only */
#endif

// Polynomial / reduction constants live in __constant__ memory so that DFMA takes
// them as c[bank][offset] operands (a 64-bit immediate would cost two UMOVs per use).
__constant__ double c_exp[16] = {
   2.08767569878680989792e-09,  // 1/12!
   2.50521083854417187751e-08,  // 1/11!
   2.75573192239858906526e-07,  // 1/10!
   2.75573192239858906526e-06,  // 1/9!
   2.48015873015873015873e-05,  // 1/8!
   1.98412698412698412698e-04,  // 1/7!
   1.38888888888888888889e-03,  // 1/6!
   8.33333333333333333333e-03,  // 1/5!
   4.16666666666666666667e-02,  // 1/4!
   1.66666666666666666667e-01,  // 1/3!
   1.4426950408889634074,       // [10] log2(e)
   6755399441055744.0,          // [11] 1.5 * 2^52
   -6.93147180369123816490e-01, // [12] -ln2 (high part)
   -1.90821492927058770002e-10, // [13] -ln2 (low part)
   0.0, 0.0};
__constant__ double c_as[6] = {0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429, 0.3275911};

// Table-driven exp for the tiled pair kernel: x = n ln2/64 + r, exp(x) = 2^(n>>6) * 2^((n&63)/64) * e^r,
// |r| <= ln2/128 so a degree-5 Taylor polynomial leaves r^6/720 < 3.5e-17.  10 FP64 ops instead of 15;
// the 64-entry table is copied to shared memory by the kernel (one LDS.64 per evaluation).
__constant__ double c_expt[8] = {
   0x1.71547652b82fep+6,        // [0] 64 log2(e)
   6755399441055744.0,          // [1] 1.5 * 2^52
   -0x1.62e42fee00000p-7,       // [2] -ln2/64 (high part, 21 significant bits: n * hi is exact)
   -0x1.a39ef35793c76p-39,      // [3] -ln2/64 (low part)
   8.33333333333333333333e-03,  // [4] 1/5!
   4.16666666666666666667e-02,  // [5] 1/4!
   1.66666666666666666667e-01,  // [6] 1/3!
   0.0};
__constant__ double c_exp2tab[64] = {
   0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
   0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
   0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
   0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
   0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
   0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
   0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
   0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
   0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
   0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
   0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
   0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
   0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
   0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
   0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
   0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

// Same reduction with a 256-entry table and the tail of the Taylor series in FP32 (the tiled kernel's default): the pair
// loop is bound by FP64 instruction throughput while a third of the issue slots and the whole FP32 pipe are idle
// (profiles/r02_summary.md).  x = n ln2/256 + r, |r| <= ln2/512 = 1.35e-3:  exp(r) = 1 + r + tail(r), tail = r^2/2 + r^3/6 +
// r^4/24 <= 9.2e-7 (r^5/120 < 4e-17 dropped), so FP32 arithmetic on the tail (r and the tail cross between the formats by
// integer operations on their bit patterns) leaves ~1e-13 relative to the result.  6 FP64 instructions instead of 10.
__constant__ double c_expt8[4] = {
   0x1.71547652b82fep+8,        // [0] 256 log2(e)
   6755399441055744.0,          // [1] 1.5 * 2^52
   -0x1.62e42fee00000p-9,       // [2] -ln2/256 (high part, 33 significant bits: n * hi is exact for |n| < 2^20)
   -0x1.a39ef35793c76p-41};     // [3] -ln2/256 (low part)
#if MDB_EXP_F32TAIL
__constant__ double c_exp2tab256[256] = {
   0x1.0000000000000p+0, 0x1.00b1afa5abcbfp+0, 0x1.0163da9fb3335p+0, 0x1.02168143b0281p+0,
   0x1.02c9a3e778061p+0, 0x1.037d42e11bbccp+0, 0x1.04315e86e7f85p+0, 0x1.04e5f72f654b1p+0,
   0x1.059b0d3158574p+0, 0x1.0650a0e3c1f89p+0, 0x1.0706b29ddf6dep+0, 0x1.07bd42b72a836p+0,
   0x1.0874518759bc8p+0, 0x1.092bdf66607e0p+0, 0x1.09e3ecac6f383p+0, 0x1.0a9c79b1f3919p+0,
   0x1.0b5586cf9890fp+0, 0x1.0c0f145e46c85p+0, 0x1.0cc922b7247f7p+0, 0x1.0d83b23395decp+0,
   0x1.0e3ec32d3d1a2p+0, 0x1.0efa55fdfa9c5p+0, 0x1.0fb66affed31bp+0, 0x1.1073028d7233ep+0,
   0x1.11301d0125b51p+0, 0x1.11edbab5e2ab6p+0, 0x1.12abdc06c31ccp+0, 0x1.136a814f204abp+0,
   0x1.1429aaea92de0p+0, 0x1.14e95934f312ep+0, 0x1.15a98c8a58e51p+0, 0x1.166a45471c3c2p+0,
   0x1.172b83c7d517bp+0, 0x1.17ed48695bbc0p+0, 0x1.18af9388c8deap+0, 0x1.1972658375d2fp+0,
   0x1.1a35beb6fcb75p+0, 0x1.1af99f8138a1cp+0, 0x1.1bbe084045cd4p+0, 0x1.1c82f95281c6bp+0,
   0x1.1d4873168b9aap+0, 0x1.1e0e75eb44027p+0, 0x1.1ed5022fcd91dp+0, 0x1.1f9c18438ce4dp+0,
   0x1.2063b88628cd6p+0, 0x1.212be3578a819p+0, 0x1.21f49917ddc96p+0, 0x1.22bdda27912d1p+0,
   0x1.2387a6e756238p+0, 0x1.2451ffb82140ap+0, 0x1.251ce4fb2a63fp+0, 0x1.25e85711ece75p+0,
   0x1.26b4565e27cddp+0, 0x1.2780e341ddf29p+0, 0x1.284dfe1f56381p+0, 0x1.291ba7591bb70p+0,
   0x1.29e9df51fdee1p+0, 0x1.2ab8a66d10f13p+0, 0x1.2b87fd0dad990p+0, 0x1.2c57e39771b2fp+0,
   0x1.2d285a6e4030bp+0, 0x1.2df961f641589p+0, 0x1.2ecafa93e2f56p+0, 0x1.2f9d24abd886bp+0,
   0x1.306fe0a31b715p+0, 0x1.31432edeeb2fdp+0, 0x1.32170fc4cd831p+0, 0x1.32eb83ba8ea32p+0,
   0x1.33c08b26416ffp+0, 0x1.3496266e3fa2dp+0, 0x1.356c55f929ff1p+0, 0x1.36431a2de883bp+0,
   0x1.371a7373aa9cbp+0, 0x1.37f26231e754ap+0, 0x1.38cae6d05d866p+0, 0x1.39a401b7140efp+0,
   0x1.3a7db34e59ff7p+0, 0x1.3b57fbfec6cf4p+0, 0x1.3c32dc313a8e5p+0, 0x1.3d0e544ede173p+0,
   0x1.3dea64c123422p+0, 0x1.3ec70df1c5175p+0, 0x1.3fa4504ac801cp+0, 0x1.40822c367a024p+0,
   0x1.4160a21f72e2ap+0, 0x1.423fb2709468ap+0, 0x1.431f5d950a897p+0, 0x1.43ffa3f84b9d4p+0,
   0x1.44e086061892dp+0, 0x1.45c2042a7d232p+0, 0x1.46a41ed1d0057p+0, 0x1.4786d668b3237p+0,
   0x1.486a2b5c13cd0p+0, 0x1.494e1e192aed2p+0, 0x1.4a32af0d7d3dep+0, 0x1.4b17dea6db7d7p+0,
   0x1.4bfdad5362a27p+0, 0x1.4ce41b817c114p+0, 0x1.4dcb299fddd0dp+0, 0x1.4eb2d81d8abffp+0,
   0x1.4f9b2769d2ca7p+0, 0x1.508417f4531eep+0, 0x1.516daa2cf6642p+0, 0x1.5257de83f4eefp+0,
   0x1.5342b569d4f82p+0, 0x1.542e2f4f6ad27p+0, 0x1.551a4ca5d920fp+0, 0x1.56070dde910d2p+0,
   0x1.56f4736b527dap+0, 0x1.57e27dbe2c4cfp+0, 0x1.58d12d497c7fdp+0, 0x1.59c0827ff07ccp+0,
   0x1.5ab07dd485429p+0, 0x1.5ba11fba87a03p+0, 0x1.5c9268a5946b7p+0, 0x1.5d84590998b93p+0,
   0x1.5e76f15ad2148p+0, 0x1.5f6a320dceb71p+0, 0x1.605e1b976dc09p+0, 0x1.6152ae6cdf6f4p+0,
   0x1.6247eb03a5585p+0, 0x1.633dd1d1929fdp+0, 0x1.6434634ccc320p+0, 0x1.652b9febc8fb7p+0,
   0x1.6623882552225p+0, 0x1.671c1c70833f6p+0, 0x1.68155d44ca973p+0, 0x1.690f4b19e9538p+0,
   0x1.6a09e667f3bcdp+0, 0x1.6b052fa75173ep+0, 0x1.6c012750bdabfp+0, 0x1.6cfdcddd47645p+0,
   0x1.6dfb23c651a2fp+0, 0x1.6ef9298593ae5p+0, 0x1.6ff7df9519484p+0, 0x1.70f7466f42e87p+0,
   0x1.71f75e8ec5f74p+0, 0x1.72f8286ead08ap+0, 0x1.73f9a48a58174p+0, 0x1.74fbd35d7cbfdp+0,
   0x1.75feb564267c9p+0, 0x1.77024b1ab6e09p+0, 0x1.780694fde5d3fp+0, 0x1.790b938ac1cf6p+0,
   0x1.7a11473eb0187p+0, 0x1.7b17b0976cfdbp+0, 0x1.7c1ed0130c132p+0, 0x1.7d26a62ff86f0p+0,
   0x1.7e2f336cf4e62p+0, 0x1.7f3878491c491p+0, 0x1.80427543e1a12p+0, 0x1.814d2add106d9p+0,
   0x1.82589994cce13p+0, 0x1.8364c1eb941f7p+0, 0x1.8471a4623c7adp+0, 0x1.857f4179f5b21p+0,
   0x1.868d99b4492edp+0, 0x1.879cad931a436p+0, 0x1.88ac7d98a6699p+0, 0x1.89bd0a478580fp+0,
   0x1.8ace5422aa0dbp+0, 0x1.8be05bad61778p+0, 0x1.8cf3216b5448cp+0, 0x1.8e06a5e0866d9p+0,
   0x1.8f1ae99157736p+0, 0x1.902fed0282c8ap+0, 0x1.9145b0b91ffc6p+0, 0x1.925c353aa2fe2p+0,
   0x1.93737b0cdc5e5p+0, 0x1.948b82b5f98e5p+0, 0x1.95a44cbc8520fp+0, 0x1.96bdd9a7670b3p+0,
   0x1.97d829fde4e50p+0, 0x1.98f33e47a22a2p+0, 0x1.9a0f170ca07bap+0, 0x1.9b2bb4d53fe0dp+0,
   0x1.9c49182a3f090p+0, 0x1.9d674194bb8d5p+0, 0x1.9e86319e32323p+0, 0x1.9fa5e8d07f29ep+0,
   0x1.a0c667b5de565p+0, 0x1.a1e7aed8eb8bbp+0, 0x1.a309bec4a2d33p+0, 0x1.a42c980460ad8p+0,
   0x1.a5503b23e255dp+0, 0x1.a674a8af46052p+0, 0x1.a799e1330b358p+0, 0x1.a8bfe53c12e59p+0,
   0x1.a9e6b5579fdbfp+0, 0x1.ab0e521356ebap+0, 0x1.ac36bbfd3f37ap+0, 0x1.ad5ff3a3c2774p+0,
   0x1.ae89f995ad3adp+0, 0x1.afb4ce622f2ffp+0, 0x1.b0e07298db666p+0, 0x1.b20ce6c9a8952p+0,
   0x1.b33a2b84f15fbp+0, 0x1.b468415b749b1p+0, 0x1.b59728de5593ap+0, 0x1.b6c6e29f1c52ap+0,
   0x1.b7f76f2fb5e47p+0, 0x1.b928cf22749e4p+0, 0x1.ba5b030a1064ap+0, 0x1.bb8e0b79a6f1fp+0,
   0x1.bcc1e904bc1d2p+0, 0x1.bdf69c3f3a207p+0, 0x1.bf2c25bd71e09p+0, 0x1.c06286141b33dp+0,
   0x1.c199bdd85529cp+0, 0x1.c2d1cd9fa652cp+0, 0x1.c40ab5fffd07ap+0, 0x1.c544778fafb22p+0,
   0x1.c67f12e57d14bp+0, 0x1.c7ba88988c933p+0, 0x1.c8f6d9406e7b5p+0, 0x1.ca3405751c4dbp+0,
   0x1.cb720dcef9069p+0, 0x1.ccb0f2e6d1675p+0, 0x1.cdf0b555dc3fap+0, 0x1.cf3155b5bab74p+0,
   0x1.d072d4a07897cp+0, 0x1.d1b532b08c968p+0, 0x1.d2f87080d89f2p+0, 0x1.d43c8eacaa1d6p+0,
   0x1.d5818dcfba487p+0, 0x1.d6c76e862e6d3p+0, 0x1.d80e316c98398p+0, 0x1.d955d71ff6075p+0,
   0x1.da9e603db3285p+0, 0x1.dbe7cd63a8315p+0, 0x1.dd321f301b460p+0, 0x1.de7d5641c0658p+0,
   0x1.dfc97337b9b5fp+0, 0x1.e11676b197d17p+0, 0x1.e264614f5a129p+0, 0x1.e3b333b16ee12p+0,
   0x1.e502ee78b3ff6p+0, 0x1.e653924676d76p+0, 0x1.e7a51fbc74c83p+0, 0x1.e8f7977cdb740p+0,
   0x1.ea4afa2a490dap+0, 0x1.eb9f4867cca6ep+0, 0x1.ecf482d8e67f1p+0, 0x1.ee4aaa2188510p+0,
   0x1.efa1bee615a27p+0, 0x1.f0f9c1cb6412ap+0, 0x1.f252b376bba97p+0, 0x1.f3ac948dd7274p+0,
   0x1.f50765b6e4540p+0, 0x1.f6632798844f8p+0, 0x1.f7bfdad9cbe14p+0, 0x1.f91d802243c89p+0,
   0x1.fa7c1819e90d8p+0, 0x1.fbdba3692d514p+0, 0x1.fd3c22b8f71f1p+0, 0x1.fe9d96b2a23d9p+0};
#else
__constant__ double c_exp2tab256[1] = {1.0};
#endif

// All helpers below work on NV independent values at once and are written
// step-by-step across the NV lanes ("structure of arrays" in registers): ptxas keeps
// that order, so the NV dependent DFMA chains are issued interleaved and the FP64
// pipe (2 issue cycles per warp-instruction, ~8 cycles latency) stays busy even with
// two or three resident warps per scheduler.
#define MDB_V for (int k = 0; k < NV; k++)

// 1/sqrt(x), x > 0 normal.  MUFU.RSQ64H seed (~2^-20) + one cubic step (5 FP64 ops).
template <int NV>
__device__ __forceinline__ void mdb_rsqrt_v(const double (&x)[NV], double (&y)[NV])
{
#if MDB_LIBM_MATH
#pragma unroll
   MDB_V y[k] = rsqrt(x[k]);
#else
   // one cubic (Householder) step from the ~2^-20 seed: e = 1 - x y^2, y' = y + y e (1/2 + 3/8 e);
   // remaining relative error 5/16 e^3 < 1e-18
   double t[NV], e[NV];
#pragma unroll
   MDB_V asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[k]) : "d"(x[k]));
#pragma unroll
   MDB_V t[k] = x[k] * y[k];
#pragma unroll
   MDB_V e[k] = fma(-t[k], y[k], 1.0);
#pragma unroll
   MDB_V t[k] = fma(0.375, e[k], 0.5);
#pragma unroll
   MDB_V e[k] = y[k] * e[k];
#pragma unroll
   MDB_V y[k] = fma(e[k], t[k], y[k]);
#endif
}

// 1/x, x normal.  MUFU.RCP64H seed + one cubic step (3 FP64 ops).
template <int NV>
__device__ __forceinline__ void mdb_rcp_v(const double (&x)[NV], double (&y)[NV])
{
#if MDB_LIBM_MATH
#pragma unroll
   MDB_V y[k] = 1.0 / x[k];
#else
   // one cubic step: e = 1 - x y, y' = y + y (e + e^2); remaining relative error e^3
   double e[NV];
#pragma unroll
   MDB_V asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[k]) : "d"(x[k]));
#pragma unroll
   MDB_V e[k] = fma(-x[k], y[k], 1.0);
#pragma unroll
   MDB_V e[k] = fma(e[k], e[k], e[k]);
#pragma unroll
   MDB_V y[k] = fma(y[k], e[k], y[k]);
#endif
}

// exp(x) for x <= ~700: n = rint(x log2e), r = x - n ln2 in two pieces, degree-11
// Taylor on |r| <= 0.3466 (remainder < 6.3e-15), scale by adding n to the exponent
// field.  x < -708 returns 0 through an integer test on the high word (no FP64
// min/max); large positive x is the caller's responsibility (never occurs for
// -a^2 r^2, -p r).
template <int NV>
__device__ __forceinline__ void mdb_exp_v(const double (&x)[NV], double (&out)[NV])
{
#if MDB_LIBM_MATH
#pragma unroll
   MDB_V out[k] = exp(x[k]);
#else
   double t[NV], r[NV], p[NV];
#pragma unroll
   MDB_V t[k] = fma(x[k], c_exp[10], c_exp[11]);
#pragma unroll
   MDB_V r[k] = t[k] - c_exp[11];
#pragma unroll
   MDB_V p[k] = fma(r[k], c_exp[12], x[k]);
#pragma unroll
   MDB_V r[k] = fma(r[k], c_exp[13], p[k]);
#pragma unroll
   MDB_V p[k] = fma(c_exp[1], r[k], c_exp[2]);     // degree 11: remainder r^12/12! < 6.3e-15
#pragma unroll
   for (int m = 3; m < 10; m++) {
#pragma unroll
      MDB_V p[k] = fma(p[k], r[k], c_exp[m]);
   }
#pragma unroll
   MDB_V p[k] = fma(p[k], r[k], 0.5);
#pragma unroll
   MDB_V p[k] = fma(p[k], r[k], 1.0);
#pragma unroll
   MDB_V p[k] = fma(p[k], r[k], 1.0);
#pragma unroll
   MDB_V {
      const int hi = __double2hiint(p[k]) + (__double2loint(t[k]) << 20), lo = __double2loint(p[k]);
      const bool tiny = (unsigned)__double2hiint(x[k]) > 0xc0862000u;     // x < -708
      out[k] = __hiloint2double(tiny ? 0 : hi, tiny ? 0 : lo);
   }
#endif
}

// exp(x) through the shared-memory copy of c_exp2tab at 32-bit shared address `e2s` (see c_expt).
// x below -708 gives ~2^-1022 (the integer part is clamped) instead of 0: either vanishes in every sum.
template <int NV>
__device__ __forceinline__ void mdb_exp_tab_v(const double (&x)[NV], double (&out)[NV], unsigned e2s)
{
   double t[NV], r[NV], p[NV], T[NV];
   int n[NV];
#pragma unroll
   MDB_V t[k] = fma(x[k], c_expt[0], c_expt[1]);
#pragma unroll
   MDB_V {
      n[k] = max(__double2loint(t[k]), -65408);
      asm("ld.shared.f64 %0, [%1];" : "=d"(T[k]) : "r"(e2s + (unsigned)(n[k] & 63) * 8u));
   }
#pragma unroll
   MDB_V r[k] = t[k] - c_expt[1];
#pragma unroll
   MDB_V p[k] = fma(r[k], c_expt[2], x[k]);
#pragma unroll
   MDB_V r[k] = fma(r[k], c_expt[3], p[k]);
#pragma unroll
   MDB_V p[k] = fma(c_expt[4], r[k], c_expt[5]);
#pragma unroll
   MDB_V p[k] = fma(p[k], r[k], c_expt[6]);
#pragma unroll
   MDB_V p[k] = fma(p[k], r[k], 0.5);
#pragma unroll
   MDB_V p[k] = fma(p[k], r[k], 1.0);
#pragma unroll
   MDB_V p[k] = fma(p[k], r[k], 1.0);
#pragma unroll
   MDB_V p[k] = p[k] * T[k];
#pragma unroll
   MDB_V {
      int hi;
      asm("mad.lo.s32 %0, %1, 16384, %2;" : "=r"(hi) : "r"(n[k] & ~63), "r"(__double2hiint(p[k])));
      out[k] = __hiloint2double(hi, __double2loint(p[k]));
   }
}

// exp(x) through the 256-entry table at shared address `e2s`, Taylor tail in FP32 (see c_expt8).  x <= 0 in every use
// (-a^2 r^2, -r/rho); x below -708 gives ~2^-1022 (the integer part is clamped) instead of 0.
template <int NV>
__device__ __forceinline__ void mdb_exp_tab8_v(const double (&x)[NV], double (&out)[NV], unsigned e2s)
{
   double t[NV], r[NV], p[NV], T[NV];
   int n[NV];
#pragma unroll
   MDB_V t[k] = fma(x[k], c_expt8[0], c_expt8[1]);
#pragma unroll
   MDB_V {
      n[k] = max(__double2loint(t[k]), -261632);
      asm("ld.shared.f64 %0, [%1];" : "=d"(T[k]) : "r"(e2s + (unsigned)(n[k] & 255) * 8u));
   }
#pragma unroll
   MDB_V r[k] = t[k] - c_expt8[1];
#pragma unroll
   MDB_V p[k] = fma(r[k], c_expt8[2], x[k]);
#pragma unroll
   MDB_V r[k] = fma(r[k], c_expt8[3], p[k]);
#pragma unroll
   MDB_V {
      // r -> float by bit operations (|r| < 2^-9; exactly 0 or >= 2^-100 by construction): drop 29 mantissa bits, rebias the
      // exponent by clearing bit 7 of its low byte (896 = 0x380), keep the sign
      const unsigned hi = (unsigned)__double2hiint(r[k]), lo = (unsigned)__double2loint(r[k]);
      const unsigned w = __funnelshift_l(lo, hi, 3);
      const float rf = __uint_as_float((w & 0x3fffffffu) | (hi & 0x80000000u));
      float q = fmaf(rf, 0.041666668f, 0.16666667f);
      q = fmaf(rf, q, 0.5f);
      const unsigned v = __float_as_uint((rf * rf) * q);                  // tail >= 0
      p[k] = __hiloint2double((int)((v >> 3) + 0x38000000u), (int)(v << 29));
   }
#pragma unroll
   MDB_V r[k] = r[k] + p[k];
#pragma unroll
   MDB_V p[k] = fma(T[k], r[k], T[k]);
#pragma unroll
   MDB_V {
      int hi;
      asm("mad.lo.s32 %0, %1, 4096, %2;" : "=r"(hi) : "r"(n[k] & ~255), "r"(__double2hiint(p[k])));
      out[k] = __hiloint2double(hi, __double2loint(p[k]));
   }
}

// dispatch: TAB selects the table-driven exp (tiled pair kernel), otherwise the self-contained one
template <int NV, bool TAB>
__device__ __forceinline__ void mdb_exp_sel(const double (&x)[NV], double (&out)[NV], unsigned e2s)
{
   if constexpr (TAB) {
      if (MDB_EXP_F32TAIL) mdb_exp_tab8_v<NV>(x, out, e2s);
      else mdb_exp_tab_v<NV>(x, out, e2s);
   } else mdb_exp_v<NV>(x, out);
}

// ---- pair potentials --------------------------------------------------------
// -phi'(r)/r and phi(r) for one site pair, the arithmetic of src/kernel.c:182-461
// (one instantiation per potential type x {Coulomb on, off}; no run-time branch
// inside).  `p` points at the 8 parameters of the (type_i,type_j) entry as
// Moldy stores them in pot_mt.p (LJ: derived values, see PairTable).  The
// Coulomb part is the reference's erfc-screened term with its own A&S 7.1.26
// polynomial (src/kernel.c:88-96,195-201) -- NOT erfc().
enum { PT_LJ = 0, PT_E6 = 1, PT_MCY = 2, PT_GEN = 3, PT_HIW = 4, PT_RSV = 5, PT_MOR = 6,
       PT_NONE = 7 };   // PT_NONE: Coulomb term only (charged-site pass of the split pair kernel)

struct PairOut { double fij, phi; };

// A pair-parameter row either behind an ordinary pointer or at a 32-bit shared-memory address
// (tiled pair kernel: the table lives in shared memory and is read with ld.shared + immediate offset).
struct MdbSmemRow {
   unsigned a;
   __device__ __forceinline__ double operator[](int n) const
   {
      double v;
      asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a + 8u * (unsigned)n));
      return v;
   }
};
__device__ __forceinline__ double2 mdb_row_ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ double2 mdb_row_ld2(MdbSmemRow r)
{
   double2 v;
   asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(r.a));
   return v;
}

// NV pairs at once: r2[k], qq[k] = q_i q_j, p[k] -> parameter row of the (type_i,type_j) entry.
template <int PT, bool COUL, int NV, bool TAB = false, class ROW = const double *>
__device__ __forceinline__ void mdb_pair_eval_v(const double (&r2)[NV], const double (&qq)[NV],
                                                const ROW (&p)[NV], double alpha, double norm,
                                                double (&fij)[NV], double (&phi)[NV], unsigned e2s = 0)
{
   double r_r[NV], r[NV], r_sqr_r[NV], erfc_term[NV], t[NV];
   if (COUL || PT != PT_LJ) {
      mdb_rsqrt_v<NV>(r2, r_r);
#pragma unroll
      MDB_V r[k] = r2[k] * r_r[k];
#pragma unroll
      MDB_V r_sqr_r[k] = r_r[k] * r_r[k];
   } else {
      mdb_rcp_v<NV>(r2, r_sqr_r);
#pragma unroll
      MDB_V r_r[k] = r[k] = 0.0;
   }
#pragma unroll
   MDB_V erfc_term[k] = t[k] = 0.0;
   if (COUL) {
      double u[NV], tt[NV], x[NV], e[NV], poly[NV];
      const double ppa = c_as[5] * alpha, na2 = -(alpha * alpha);
#pragma unroll
      MDB_V u[k] = fma(ppa, r[k], 1.0);
#pragma unroll
      MDB_V x[k] = na2 * r2[k];
      mdb_rcp_v<NV>(u, tt);
      mdb_exp_sel<NV, TAB>(x, e, e2s);
#pragma unroll
      MDB_V e[k] = qq[k] * e[k];
#pragma unroll
      MDB_V poly[k] = fma(tt[k], c_as[4], c_as[3]);
#pragma unroll
      MDB_V poly[k] = fma(tt[k], poly[k], c_as[2]);
#pragma unroll
      MDB_V poly[k] = fma(tt[k], poly[k], c_as[1]);
#pragma unroll
      MDB_V poly[k] = fma(tt[k], poly[k], c_as[0]);
#pragma unroll
      MDB_V poly[k] = tt[k] * poly[k];
#pragma unroll
      MDB_V t[k] = poly[k] * e[k] * r_r[k];
#pragma unroll
      MDB_V erfc_term[k] = fma(norm, e[k], t[k]);
   }
   if (PT == PT_NONE) {                     // what every potential below gives with all-zero parameters
#pragma unroll
      MDB_V { phi[k] = t[k]; fij[k] = r_sqr_r[k] * erfc_term[k]; }
   } else if (PT == PT_LJ) {                // p[0]=eps, p[1]=sigma^2, p[2]=6 eps
      double r6[NV], r12[NV];
      double2 p01[NV];
      double p2[NV];
#pragma unroll
      MDB_V { p01[k] = mdb_row_ld2(p[k]); p2[k] = p[k][2]; }   // rows are 64-byte aligned
#pragma unroll
      MDB_V r6[k] = p01[k].y * r_sqr_r[k];
#pragma unroll
      MDB_V r6[k] = r6[k] * r6[k] * r6[k];
#pragma unroll
      MDB_V r12[k] = r6[k] * r6[k];
#pragma unroll
      MDB_V phi[k] = fma(p01[k].x, r12[k] - r6[k], t[k]);
#pragma unroll
      MDB_V fij[k] = r_sqr_r[k] * fma(p2[k], fma(2.0, r12[k], -r6[k]), erfc_term[k]);
   } else if (PT == PT_E6) {                // -p0/r^6 + p1 exp(-p2 r)
      double x[NV], e1[NV], r6[NV];
#pragma unroll
      MDB_V x[k] = -p[k][2] * r[k];
      mdb_exp_sel<NV, TAB>(x, e1, e2s);
#pragma unroll
      MDB_V e1[k] = p[k][1] * e1[k];
#pragma unroll
      MDB_V r6[k] = p[k][0] * r_sqr_r[k] * r_sqr_r[k] * r_sqr_r[k];
#pragma unroll
      MDB_V phi[k] = t[k] - r6[k] + e1[k];
#pragma unroll
      MDB_V fij[k] = fma(r_sqr_r[k], fma(-6.0, r6[k], erfc_term[k]), p[k][2] * e1[k] * r_r[k]);
   } else if (PT == PT_MCY) {               // p0 exp(-p1 r) - p2 exp(-p3 r)
      double x1[NV], x2[NV], e1[NV], e2[NV];
#pragma unroll
      MDB_V { x1[k] = -p[k][1] * r[k]; x2[k] = -p[k][3] * r[k]; }
      mdb_exp_sel<NV, TAB>(x1, e1, e2s);
      mdb_exp_sel<NV, TAB>(x2, e2, e2s);
#pragma unroll
      MDB_V { e1[k] = p[k][0] * e1[k]; e2[k] = -p[k][2] * e2[k]; }
#pragma unroll
      MDB_V phi[k] = t[k] + e1[k] + e2[k];
#pragma unroll
      MDB_V fij[k] = fma(fma(p[k][1], e1[k], p[k][3] * e2[k]), r_r[k], erfc_term[k] * r_sqr_r[k]);
   } else if (PT == PT_GEN) {               // p0 exp(-p1 r) + p2/r^12 - p3/r^4 - p4/r^6 - p5/r^8
      double x[NV], e1[NV];
#pragma unroll
      MDB_V x[k] = -p[k][1] * r[k];
      mdb_exp_sel<NV, TAB>(x, e1, e2s);
#pragma unroll
      MDB_V {
         e1[k] = p[k][0] * e1[k];
         double r4 = r_sqr_r[k] * r_sqr_r[k];
         double r6 = r_sqr_r[k] * r4;
         const double r8 = p[k][5] * r4 * r4;
         const double r12 = p[k][2] * r6 * r6;
         r4 *= p[k][3];
         r6 *= p[k][4];
         phi[k] = t[k] + e1[k] + r12 - r4 - r6 - r8;
         fij[k] = fma(r_sqr_r[k], 12.0 * r12 - 4.0 * r4 - 6.0 * r6 - 8.0 * r8 + erfc_term[k], p[k][1] * e1[k] * r_r[k]);
      }
   } else if (PT == PT_HIW) {               // p0/r^4 + p1/r^6 + p2/r^12
#pragma unroll
      MDB_V {
         double r4 = r_sqr_r[k] * r_sqr_r[k];
         double r6 = r_sqr_r[k] * r4;
         const double r12 = p[k][2] * r6 * r6;
         r6 *= p[k][1];
         r4 *= p[k][0];
         phi[k] = t[k] + r4 + r6 + r12;
         fij[k] = r_sqr_r[k] * (4.0 * r4 + 6.0 * r6 + 12.0 * r12 + erfc_term[k]);
      }
   } else {                                 // PT_MOR: Busing-Ida-Gilbert + Morse
      double x1[NV], x2[NV], x3[NV], e1[NV], e2[NV], e3[NV];
#pragma unroll
      MDB_V {
         x1[k] = (p[k][1] - r[k]) * p[k][2];
         x2[k] = -2.0 * p[k][5] * (r[k] - p[k][6]);
         x3[k] = -p[k][5] * (r[k] - p[k][6]);
      }
      mdb_exp_sel<NV, TAB>(x1, e1, e2s);
      mdb_exp_sel<NV, TAB>(x2, e2, e2s);
      mdb_exp_sel<NV, TAB>(x3, e3, e2s);
#pragma unroll
      MDB_V {
         e1[k] = p[k][0] * e1[k];
         e2[k] = p[k][4] * e2[k];
         e3[k] = -p[k][4] * 2.0 * e3[k];
         const double r6 = p[k][3] * r_sqr_r[k] * r_sqr_r[k] * r_sqr_r[k];
         phi[k] = t[k] + e1[k] - r6 + e2[k] + e3[k];
         fij[k] = fma(r_sqr_r[k], fma(-6.0, r6, erfc_term[k]),
                      r_r[k] * (p[k][2] * e1[k] + (2.0 * p[k][5]) * e2[k] + p[k][5] * e3[k]));
      }
   }
}

// single pair (kernel() of the C ABI, per-thread pair kernel)
template <int PT, bool COUL>
__device__ __forceinline__ PairOut mdb_pair_eval(double r2, double qq, const double *__restrict__ p,
                                                 double alpha, double norm)
{
   const double r2v[1] = {r2}, qqv[1] = {qq};
   const double *const pv[1] = {p};
   double f[1], ph[1];
   mdb_pair_eval_v<PT, COUL, 1>(r2v, qqv, pv, alpha, norm, f, ph);
   PairOut o;
   o.fij = f[0]; o.phi = ph[0];
   return o;
}
