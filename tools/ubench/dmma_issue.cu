// Microbenchmark: fp64 tensor-core MMA (mma.sync m8n8k4 f64 = DMMA.884) throughput on sm_100a next to DFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_issue dmma_issue.cu && ./dmma_issue
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int CH>
__global__ void k(double * out, int iters, double x, double y, long long * cycles)
{
  double c[CH][2];
  #pragma unroll
  for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
  const long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < iters; ++it)
  {
    #pragma unroll
    for (int rep = 0; rep < 4; ++rep)
      #pragma unroll
      for (int i = 0; i < CH; ++i) dmma(c[i], x, y);
  }
  const long long t1 = clock64();
  double s = 0;
  #pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CH>
void run(int warps_per_smsp, double * out, long long * cyc, int sms)
{
  const int iters = 2000, threads = warps_per_smsp * 128;
  k<CH><<<sms, threads>>>(out, 10, 1.0000001, 1e-9, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<CH><<<sms, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[1024]; cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
  const double n = (double) iters * 4 * CH * warps_per_smsp * 4;     // warp-level DMMAs per SM
  printf("DMMA.884 chains %d warps/SMSP %d: %.4f warp-instr/clk/SMSP = %.1f FMA/clk/SM, %.2f TFLOP/s fp64\n", CH, warps_per_smsp,
         n / 4 / c, n * 256 / c, n * 256 * 2 * sms / (ms * 1e-3) / 1e12);
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double * out; long long * cyc;
  cudaMalloc(&out, sms * 1024 * sizeof(double)); cudaMalloc(&cyc, 1024 * sizeof(long long));
  printf("%s, %d SMs\n", p.name, sms);
  for (int w : {1, 2, 4, 8}) { run<1>(w, out, cyc, sms); run<4>(w, out, cyc, sms); run<8>(w, out, cyc, sms); }
  return 0;
}
