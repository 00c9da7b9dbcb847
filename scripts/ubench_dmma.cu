// FP64 tensor-core (DMMA, mma.sync .f64) throughput probe for sm_100a: is an FP64 contraction faster through
// mma.sync than through DFMA chains on this machine?  Shapes m8n8k4, m16n8k4, m16n8k8, m16n8k16; NACC independent
// accumulator tiles per warp, W warps per scheduler.  Prints TFLOP/s (2*M*N*K per instruction).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_dmma scripts/ubench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE> struct Sh;
template <> struct Sh<0> { static constexpr int NA = 1, NB = 1, NC = 2; static constexpr double FL = 2.0 * 8 * 8 * 4; };
template <> struct Sh<1> { static constexpr int NA = 2, NB = 1, NC = 4; static constexpr double FL = 2.0 * 16 * 8 * 4; };
template <> struct Sh<2> { static constexpr int NA = 4, NB = 2, NC = 4; static constexpr double FL = 2.0 * 16 * 8 * 8; };
template <> struct Sh<3> { static constexpr int NA = 8, NB = 4, NC = 4; static constexpr double FL = 2.0 * 16 * 8 * 16; };

template <int SHAPE>
__device__ __forceinline__ void dmma(double *c, const double *a, const double *b)
{
   if constexpr (SHAPE == 0)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
   if constexpr (SHAPE == 1)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
   if constexpr (SHAPE == 2)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
   if constexpr (SHAPE == 3)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int SHAPE, int NACC>
__global__ void k_dmma(double *out, int iters)
{
   using S = Sh<SHAPE>;
   double a[S::NA], b[S::NB], c[NACC][S::NC];
   for (int i = 0; i < S::NA; i++) a[i] = 1e-3 * (threadIdx.x % 7 + i);
   for (int i = 0; i < S::NB; i++) b[i] = 1e-3 * (threadIdx.x % 5 + i);
   for (int k = 0; k < NACC; k++)
      for (int i = 0; i < S::NC; i++) c[k][i] = k + i;
   for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
#pragma unroll
         for (int k = 0; k < NACC; k++) dmma<SHAPE>(c[k], a, b);
      }
   }
   double s = 0;
   for (int k = 0; k < NACC; k++)
      for (int i = 0; i < S::NC; i++) s += c[k][i];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA reference in the same harness: NACC*4 independent accumulate-form chains (acc += w*u, three registers)
template <int NACC>
__global__ void k_dfma(double *out, int iters, double w0)
{
   double c[NACC], w = w0 + 1e-9 * threadIdx.x, u = 1e-3 + 1e-9 * threadIdx.x;
   for (int k = 0; k < NACC; k++) c[k] = k;
   for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
#pragma unroll
         for (int k = 0; k < NACC; k++) c[k] = fma(w, u, c[k]);
      }
   }
   double s = 0;
   for (int k = 0; k < NACC; k++) s += c[k];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NACC>
static void run(const char *name, int nsm, int wps, double *d)
{
   const int threads = 32 * 4 * wps, iters = 2000;
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   k_dmma<SHAPE, NACC><<<nsm, threads>>>(d, 10);
   cudaEventRecord(e0);
   k_dmma<SHAPE, NACC><<<nsm, threads>>>(d, iters);
   cudaEventRecord(e1); cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1);
   int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
   const double ninst = (double)iters * 4 * NACC * 4 * wps * nsm;
   const double cyc = ms * 1e-3 * clk * 1e3 / ((double)iters * 4 * NACC * wps);
   printf("%-10s acc=%d warps/sched=%d  %8.3f ms  %7.2f TFLOP/s  %6.2f cycles per instr per scheduler  (err %s)\n", name, NACC, wps, ms,
          ninst * Sh<SHAPE>::FL / (ms * 1e-3) * 1e-12, cyc, cudaGetErrorString(cudaGetLastError()));
}

template <int NACC>
static void run_dfma(int nsm, int wps, double *d)
{
   const int threads = 32 * 4 * wps, iters = 2000;
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   k_dfma<NACC><<<nsm, threads>>>(d, 10, 0.999);
   cudaEventRecord(e0);
   k_dfma<NACC><<<nsm, threads>>>(d, iters, 0.999);
   cudaEventRecord(e1); cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1);
   const double ninst = (double)iters * 4 * NACC * 4 * wps * nsm;
   printf("DFMA acc+=w*u acc=%d warps/sched=%d  %8.3f ms  %7.2f TFLOP/s\n", NACC, wps, ms, ninst * 64 / (ms * 1e-3) * 1e-12);
}

int main()
{
   cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
   const int nsm = p.multiProcessorCount;
   double *d; cudaMalloc(&d, sizeof(double) * nsm * 1024);
   printf("%s, %d SMs\n", p.name, nsm);
   for (int wps = 1; wps <= 4; wps *= 2) {
      run<0, 1>("m8n8k4", nsm, wps, d);  run<0, 4>("m8n8k4", nsm, wps, d);  run<0, 8>("m8n8k4", nsm, wps, d);
      run<1, 1>("m16n8k4", nsm, wps, d); run<1, 4>("m16n8k4", nsm, wps, d); run<1, 8>("m16n8k4", nsm, wps, d);
      run<2, 1>("m16n8k8", nsm, wps, d); run<2, 4>("m16n8k8", nsm, wps, d); run<2, 8>("m16n8k8", nsm, wps, d);
      run<3, 1>("m16n8k16", nsm, wps, d); run<3, 4>("m16n8k16", nsm, wps, d); run<3, 8>("m16n8k16", nsm, wps, d);
      run_dfma<8>(nsm, wps, d); run_dfma<16>(nsm, wps, d);
   }
   return 0;
}
