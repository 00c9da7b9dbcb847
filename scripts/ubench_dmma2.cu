// DMMA.8x8x4 issue rate with realistic operand traffic (sm_100a): distinct A/B registers per instruction,
// operands reloaded from shared memory every step, and interleaved DFMA work, at 4 warps per scheduler.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/bin/ubench_dmma2 scripts/ubench_dmma2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// MODE 0: 16 accumulators, a[2] x b[8] all in registers, constant          (operand variety only)
// MODE 1: same, a and b reloaded from shared memory each step (LDS.64)     (k_sfac_mma's pattern)
// MODE 2: MODE 1 + 4 DFMA/DMUL per step forming a from two loaded values   (k_sfac_mma's A generation)
// MODE 4: MODE 2 with the warps of a scheduler desynchronised (each starts after a different delay)
// MODE 5: MODE 4 but the FP64 A generation of all 8 steps is done in one burst ahead of the 128 DMMAs
// MODE 3: 8 accumulators, a[2] x b[4] from LDS each step                   (k_kforce_mma's pattern)
template <int MODE>
__global__ void k(double *out, int iters)
{
   extern __shared__ double sm[];
   for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1e-3 * (i % 13);
   __syncthreads();
   const int lane = threadIdx.x & 31;
   constexpr int NA = 2, NB = MODE == 3 ? 4 : 8;
   double a[NA], b[NB], c[NA][NB][2];
   for (int i = 0; i < NA; i++) a[i] = 1e-3 * (lane + i);
   for (int j = 0; j < NB; j++) b[j] = 1e-3 * (lane % 5 + j);
   for (int i = 0; i < NA; i++)
      for (int j = 0; j < NB; j++) c[i][j][0] = c[i][j][1] = i + j;
   const double *p = sm + lane;
   if (MODE >= 4) {                     // skew the warps: up to ~3000 cycles
      const long long t0 = clock64();
      while (clock64() - t0 < 211 * (threadIdx.x >> 5)) { }
   }
   if (MODE == 5) {
      for (int it = 0; it < iters; it++) {
         double aa[8][NA];
#pragma unroll
         for (int t = 0; t < 8; t++)
#pragma unroll
            for (int i = 0; i < NA; i++) {
               const double u = p[(t * 68 + i * 40 + 1024) & 4095 - 31], v = p[(t * 68 + i * 40 + 2048) & 4095 - 31];
               aa[t][i] = fma(u, v, u * v);
            }
#pragma unroll
         for (int t = 0; t < 8; t++) {
#pragma unroll
            for (int j = 0; j < NB; j++) b[j] = p[(t * 64 + j * 36) & 4095 - 31];
#pragma unroll
            for (int j = 0; j < NB; j++)
#pragma unroll
               for (int i = 0; i < NA; i++) dmma(c[i][j], aa[t][i], b[j]);
         }
      }
   } else
   for (int it = 0; it < iters; it++) {
#pragma unroll 2
      for (int t = 0; t < 8; t++) {
         if (MODE >= 1) {
#pragma unroll
            for (int j = 0; j < NB; j++) b[j] = p[(t * 64 + j * 36) & 4095 - 31];
            if (MODE == 2 || MODE == 4) {
#pragma unroll
               for (int i = 0; i < NA; i++) {
                  const double u = p[(t * 68 + i * 40 + 1024) & 4095 - 31], v = p[(t * 68 + i * 40 + 2048) & 4095 - 31];
                  a[i] = fma(u, v, u * b[0]);
               }
            } else {
#pragma unroll
               for (int i = 0; i < NA; i++) a[i] = p[(t * 68 + i * 40 + 1024) & 4095 - 31];
            }
         }
#pragma unroll
         for (int j = 0; j < NB; j++)
#pragma unroll
            for (int i = 0; i < NA; i++) dmma(c[i][j], a[i], b[j]);
      }
   }
   double s = 0;
   for (int i = 0; i < NA; i++)
      for (int j = 0; j < NB; j++) s += c[i][j][0] + c[i][j][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char *name, int nsm, int wps, double *d)
{
   const int threads = 32 * 4 * wps, iters = 500;
   constexpr int NB = MODE == 3 ? 4 : 8;
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   k<MODE><<<nsm, threads, 32768>>>(d, 5);
   cudaEventRecord(e0);
   k<MODE><<<nsm, threads, 32768>>>(d, iters);
   cudaEventRecord(e1); cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1);
   int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
   const double n_per_warp = (double)iters * 8 * 2 * NB;
   printf("%-44s warps/sched=%d  %8.3f ms  %6.2f cycles per DMMA per scheduler  %6.2f TFLOP/s (%s)\n", name, wps, ms,
          ms * 1e-3 * clk * 1e3 / (n_per_warp * wps), n_per_warp * 4 * wps * nsm * 512 / (ms * 1e-3) * 1e-12,
          cudaGetErrorString(cudaGetLastError()));
}

int main()
{
   cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
   const int nsm = p.multiProcessorCount;
   double *d; cudaMalloc(&d, sizeof(double) * nsm * 1024);
   for (int wps = 2; wps <= 4; wps += 2) {
      run<0>("16 acc, operands in registers", nsm, wps, d);
      run<1>("16 acc, a,b from LDS.64 every step", nsm, wps, d);
      run<2>("16 acc, LDS + DFMA/DMUL A generation", nsm, wps, d);
      run<3>("8 acc, a,b from LDS.64 every step", nsm, wps, d);
      run<4>("16 acc, LDS + FP64 A gen, warps skewed", nsm, wps, d);
      run<5>("16 acc, A gen of 8 steps in one burst, skewed", nsm, wps, d);
   }
   return 0;
}
