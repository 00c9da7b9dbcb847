// FP64 pipe micro-benchmarks for sm_100a: dependent-chain latency and throughput of DFMA streams as a
// function of resident warps per scheduler and independent chains per warp, with and without interleaved
// integer / shared-memory instructions.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/ubench.cu -o ubench
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MIX>
__global__ void k_chain(double *out, int iters, double a, double b, int *ibuf)
{
   __shared__ double sm[256];
   sm[threadIdx.x % 256] = threadIdx.x;
   __syncthreads();
   double v[ILP];
   int x = threadIdx.x, y = 0;
#pragma unroll
   for (int k = 0; k < ILP; k++) v[k] = threadIdx.x + k;
   for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int r = 0; r < 8; r++) {
#pragma unroll
         for (int k = 0; k < ILP; k++) v[k] = fma(v[k], a, b);
         if (MIX == 1) { x = x * 3 + r; y ^= x; }                       // 2 integer ops per ILP DFMAs
         if (MIX == 2) { v[0] += sm[(x + r) & 255]; }                   // 1 LDS + DADD
         if (MIX == 3) { x = x * 3 + r; y ^= x; x += y >> 3; y += x & 7; }     // 4 integer ops
      }
   }
   double s = 0;
#pragma unroll
   for (int k = 0; k < ILP; k++) s += v[k];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s + y;
   if (ibuf) ibuf[0] = x;
}

template <int ILP, int MIX>
static void run(const char *name, int warps_per_sched, int nsm)
{
   double *d; cudaMalloc(&d, sizeof(double) * nsm * 1024);
   const int threads = 32 * 4 * warps_per_sched, iters = 4000;
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   k_chain<ILP, MIX><<<nsm, threads>>>(d, 10, 0.999, 1e-3, nullptr);
   cudaEventRecord(e0);
   k_chain<ILP, MIX><<<nsm, threads>>>(d, iters, 0.999, 1e-3, nullptr);
   cudaEventRecord(e1); cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1);
   int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
   const double cycles = ms * 1e-3 * clk * 1e3;
   const double dfma_per_sched = (double)iters * 8 * ILP * warps_per_sched;
   printf("%-10s ILP=%d warps/sched=%d : %.2f cycles per DFMA warp-instr per scheduler (%.1f%% of 2.0)  [%.3f ms]\n", name, ILP,
          warps_per_sched, cycles / dfma_per_sched, 200.0 / (cycles / dfma_per_sched), ms);
   cudaFree(d);
}

int main()
{
   cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
   printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
   const int nsm = p.multiProcessorCount;
   printf("-- pure DFMA chains (ILP=1, 1 warp: cycles = dependent-issue latency)\n");
   run<1, 0>("dfma", 1, nsm); run<2, 0>("dfma", 1, nsm); run<4, 0>("dfma", 1, nsm); run<8, 0>("dfma", 1, nsm);
   run<1, 0>("dfma", 2, nsm); run<1, 0>("dfma", 3, nsm); run<1, 0>("dfma", 4, nsm); run<1, 0>("dfma", 8, nsm);
   run<4, 0>("dfma", 2, nsm); run<4, 0>("dfma", 3, nsm); run<4, 0>("dfma", 4, nsm);
   printf("-- 2 integer ops per ILP DFMAs\n");
   run<4, 1>("dfma+2int", 1, nsm); run<4, 1>("dfma+2int", 3, nsm); run<2, 1>("dfma+2int", 3, nsm); run<1, 1>("dfma+2int", 3, nsm);
   printf("-- 4 integer ops per ILP DFMAs\n");
   run<4, 3>("dfma+4int", 3, nsm); run<2, 3>("dfma+4int", 3, nsm); run<1, 3>("dfma+4int", 3, nsm); run<1, 3>("dfma+4int", 8, nsm);
   printf("-- 1 LDS + DADD per ILP DFMAs\n");
   run<4, 2>("dfma+lds", 3, nsm); run<2, 2>("dfma+lds", 3, nsm);
   return 0;
}
