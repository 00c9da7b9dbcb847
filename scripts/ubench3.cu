// Does operand sharing between consecutive DFMAs (register reuse cache) or a uniform-register / constant operand
// remove the 3-register-operand penalty on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double c_tab[512];

template <int MODE>
__global__ void k_ops(double *out, int iters, const double *__restrict__ g)
{
   double v[8], w[4], u[4];
#pragma unroll
   for (int k = 0; k < 8; k++) v[k] = threadIdx.x + k;
#pragma unroll
   for (int k = 0; k < 4; k++) { w[k] = 0.999 + 1e-6 * (threadIdx.x + k); u[k] = 1e-3 * (k + 1) + 1e-9 * threadIdx.x; }
   for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
         if (MODE == 1) {           // shared operand A across 8 consecutive accumulates
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = fma(w[r], u[k & 3], v[k]);
         }
         if (MODE == 2) {           // all three operands differ from the previous instruction
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = fma(w[(k + r) & 3], u[(k + 1) & 3], v[k]);
         }
         if (MODE == 3) {           // coefficient from constant memory with a uniform index (LDCU + UR operand?)
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = fma(w[k & 3], c_tab[(i & 31) * 8 + k + r], v[k]);
         }
         if (MODE == 4) {           // coefficient from global memory with a uniform address
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = fma(w[k & 3], __ldg(&g[(i & 31) * 8 + k + r]), v[k]);
         }
      }
   }
   double s = 0;
#pragma unroll
   for (int k = 0; k < 8; k++) s += v[k];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s + w[0] + u[0];
}

template <int MODE>
static void run(const char *name, int nsm, const double *g)
{
   double *d; cudaMalloc(&d, sizeof(double) * nsm * 1024);
   const int threads = 32 * 4 * 3, iters = 4000;
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   k_ops<MODE><<<nsm, threads>>>(d, 10, g);
   cudaEventRecord(e0);
   k_ops<MODE><<<nsm, threads>>>(d, iters, g);
   cudaEventRecord(e1); cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1);
   int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
   const double cycles = ms * 1e-3 * clk * 1e3;
   printf("%-52s %.2f cycles per DFMA warp-instr per scheduler\n", name, cycles / ((double)iters * 4 * 8 * 3));
   cudaFree(d);
}

int main()
{
   cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
   const int nsm = p.multiProcessorCount;
   double h[512]; for (int i = 0; i < 512; i++) h[i] = 1e-3 * i;
   cudaMemcpyToSymbol(c_tab, h, sizeof h);
   double *g; cudaMalloc(&g, sizeof h); cudaMemcpy(g, h, sizeof h, cudaMemcpyHostToDevice);
   run<1>("acc += w*u, w shared by 8 consecutive DFMAs", nsm, g);
   run<2>("acc += w*u, no operand shared", nsm, g);
   run<3>("acc += w*c[uniform idx] (constant memory)", nsm, g);
   run<4>("acc += w*g[uniform idx] (global, __ldg)", nsm, g);
   return 0;
}
