// Do DFMA (FP64 CUDA-core pipe) and DMMA.8x8x4 (FP64 tensor pipe) share execution resources on sm_100a?
// 32 warps per block, one block per SM (4 + 4 warps per scheduler, enough to saturate either pipe alone): warps with
// ((warp >> 2) & 1) == 0 run 8 independent DFMA chains, the others 8 independent DMMA accumulator tiles.  Timed: DFMA warps alone, DMMA warps alone, both together.  If "both" ~ max(alone) the pipes are
// independent and a DFMA-bound kernel (real-space pair sum) can overlap a DMMA-bound one (k-space GEMMs) on the same SMs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/bin/ubench_mix scripts/ubench_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) k_mix(double *out, int it_dfma, int it_dmma, int mode)
{
   const int warp = threadIdx.x >> 5;
   const bool is_dfma = ((warp >> 2) & 1) == 0;      // warps 4k..4k+3 alternate: every scheduler (warp % 4) gets both kinds
   double s = 0;
   if (is_dfma) {
      if (!(mode & 1)) return;
      double v[8], a = 0.999999, b = 1e-9;
      for (int k = 0; k < 8; k++) v[k] = threadIdx.x + k;
      for (int i = 0; i < it_dfma; i++) {
#pragma unroll
         for (int k = 0; k < 8; k++) v[k] = fma(v[k], a, b);
      }
      for (int k = 0; k < 8; k++) s += v[k];
   } else {
      if (!(mode & 2)) return;
      double a = 1e-3 * (threadIdx.x % 7 + 1), b = 1e-3 * (threadIdx.x % 5 + 1), c[8][2];
      for (int k = 0; k < 8; k++) { c[k][0] = k; c[k][1] = k + 1; }
      for (int i = 0; i < it_dmma; i++) {
#pragma unroll
         for (int k = 0; k < 8; k++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
      }
      for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1];
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
   cudaDeviceProp p;
   cudaGetDeviceProperties(&p, 0);
   const int blocks = p.multiProcessorCount, threads = 1024;
   double *d;
   cudaMalloc(&d, sizeof(double) * blocks * threads);
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   const int it_dfma = 200000, it_dmma = 25000;      // 8 DFMA (2 cycles each) vs 8 DMMA (16 cycles each) per iteration
   for (int mode = 1; mode <= 3; mode++) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; rep++) {
         cudaEventRecord(e0);
         k_mix<<<blocks, threads>>>(d, it_dfma, it_dmma, mode);
         cudaEventRecord(e1);
         cudaEventSynchronize(e1);
         float ms;
         cudaEventElapsedTime(&ms, e0, e1);
         best = ms < best ? ms : best;
      }
      const double fl_dfma = (mode & 1) ? 2.0 * 8 * it_dfma * 32.0 * 16 * blocks : 0, fl_dmma = (mode & 2) ? 512.0 * 8 * it_dmma * 16 * blocks : 0;
      printf("%-22s %8.3f ms   DFMA %.2f TFLOP/s   DMMA %.2f TFLOP/s   (%s)\n",
             mode == 1 ? "DFMA warps alone" : mode == 2 ? "DMMA warps alone" : "both together", best,
             fl_dfma / best / 1e9, fl_dmma / best / 1e9, cudaGetErrorString(cudaGetLastError()));
   }
   return 0;
}
