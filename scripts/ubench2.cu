// FP64 operand-bandwidth micro-benchmark for sm_100a: cost of a DFMA as a function of how many distinct
// 64-bit register operands it reads (1, 2 or 3), 4 independent chains per warp, 3 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k_ops(double *out, int iters, double a, double b)
{
   double v[4], w[4], u[4];
#pragma unroll
   for (int k = 0; k < 4; k++) { v[k] = threadIdx.x + k; w[k] = 0.999 + 1e-6 * (threadIdx.x + k); u[k] = 1e-3 * (k + 1) + 1e-9 * threadIdx.x; }
   for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int r = 0; r < 8; r++) {
#pragma unroll
         for (int k = 0; k < 4; k++) {
            if (MODE == 1) v[k] = fma(v[k], a, b);                 // 1 register operand, 2 constants
            if (MODE == 2) v[k] = fma(v[k], w[k], b);              // 2 register operands
            if (MODE == 3) v[k] = fma(v[k], w[k], u[k]);           // 3 distinct register operands
            if (MODE == 4) v[k] = fma(v[k], v[k], u[k]);           // 2 distinct (one repeated)
            if (MODE == 5) v[k] = v[k] * w[k];                     // DMUL, 2 registers
            if (MODE == 6) v[k] = v[k] + w[k];                     // DADD, 2 registers
            if (MODE == 7) v[k] = fma(w[k], u[k], v[k]);           // accumulate form: a*b + acc
            if (MODE == 8) v[k] = fma(w[(k + r) & 3], u[k], v[k]); // accumulate, operand a varies
         }
      }
   }
   double s = 0;
#pragma unroll
   for (int k = 0; k < 4; k++) s += v[k] + w[k] + u[k];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char *name, int nsm)
{
   double *d; cudaMalloc(&d, sizeof(double) * nsm * 1024);
   const int threads = 32 * 4 * 3, iters = 4000;
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   k_ops<MODE><<<nsm, threads>>>(d, 10, 0.999, 1e-3);
   cudaEventRecord(e0);
   k_ops<MODE><<<nsm, threads>>>(d, iters, 0.999, 1e-3);
   cudaEventRecord(e1); cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1);
   int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
   const double cycles = ms * 1e-3 * clk * 1e3;
   printf("%-44s %.2f cycles per FP64 warp-instr per scheduler\n", name, cycles / ((double)iters * 8 * 4 * 3));
   cudaFree(d);
}

int main()
{
   cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
   const int nsm = p.multiProcessorCount;
   run<1>("DFMA v = v*c + c   (1 reg)", nsm);
   run<2>("DFMA v = v*w + c   (2 regs)", nsm);
   run<3>("DFMA v = v*w + u   (3 regs)", nsm);
   run<4>("DFMA v = v*v + u   (2 distinct regs)", nsm);
   run<5>("DMUL v = v*w       (2 regs)", nsm);
   run<6>("DADD v = v+w       (2 regs)", nsm);
   run<7>("DFMA v = w*u + v   (3 regs, accumulate)", nsm);
   run<8>("DFMA v = w'*u + v  (3 regs, a rotates)", nsm);
   return 0;
}
