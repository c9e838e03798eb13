// Microbenchmark: DFMA throughput on sm_100a and what else can issue next to it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu && ./fp64_issue
// For warps/SMSP = 1, 2, 3, 4, 8 and instruction mixes (per 8 DFMA: 0, 4, 8, 16 FFMA / IMAD / LDS),
// prints DFMA per clock per SM and total warp instructions per clock per SMSP.
#include <cstdio>
#include <cuda_runtime.h>

template <int MIX, int KIND, int CH>
__global__ void k(double * out, float * fout, int iters, double x, double y, float fx, float fy, long long * cycles)
{
  __shared__ float sm[1024];
  double a[CH];
  float f[16];
  int q[16];
  #pragma unroll
  for (int i = 0; i < CH; ++i) a[i] = threadIdx.x * 1e-3 + i;
  #pragma unroll
  for (int i = 0; i < 16; ++i) { f[i] = threadIdx.x * 1e-3f + i; q[i] = threadIdx.x + i; }
  sm[threadIdx.x & 1023] = fx;
  __syncthreads();
  const long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < iters; ++it)
  {
    #pragma unroll
    for (int rep = 0; rep < 4; ++rep)
    {
      #pragma unroll
      for (int i = 0; i < CH; ++i) a[i] = fma(a[i], x, y);
      #pragma unroll
      for (int i = 0; i < MIX * CH / 8; ++i)
      {
        if (KIND == 0) f[i & 15] = fmaf(f[i & 15], fx, fy);
        else if (KIND == 1) q[i & 15] = q[i & 15] * 3 + it;
        else f[i & 15] += sm[(threadIdx.x + i * 32 + q[0]) & 1023];
      }
    }
  }
  const long long t1 = clock64();
  double s = 0; float fs = 0;
  #pragma unroll
  for (int i = 0; i < CH; ++i) s += a[i];
  #pragma unroll
  for (int i = 0; i < 16; ++i) fs += f[i] + q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  fout[blockIdx.x * blockDim.x + threadIdx.x] = fs;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MIX, int KIND, int CH>
void run(const char * name, int warps_per_smsp, double * out, float * fout, long long * cyc, int sms)
{
  const int iters = 4000;
  const int threads = warps_per_smsp * 4 * 32;
  k<MIX, KIND, CH><<<sms, threads>>>(out, fout, 10, 1.0000001, 1e-9, 1.0001f, 1e-6f, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MIX, KIND, CH><<<sms, threads>>>(out, fout, iters, 1.0000001, 1e-9, 1.0001f, 1e-6f, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[1024]; cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
  const double dfma_warp = (double) iters * 4 * CH * warps_per_smsp * 4;       // per SM
  const double other_warp = (double) iters * 4 * (MIX * CH / 8) * warps_per_smsp * 4;
  printf("%-6s ch %2d warps/SMSP %d mix %2d/8: DFMA lanes/clk/SM %6.1f  warp-instr/clk/SMSP: dfma %.3f other %.3f total %.3f  (%.3f ms, %.2f TFLOP/s fp64)\n",
         name, CH, warps_per_smsp, MIX, dfma_warp * 32 / c, dfma_warp / 4 / c, other_warp / 4 / c, (dfma_warp + other_warp) / 4 / c, ms,
         dfma_warp * 32 * 2 * sms / (ms * 1e-3) / 1e12);
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double * out; float * fout; long long * cyc;
  cudaMalloc(&out, sms * 1024 * sizeof(double)); cudaMalloc(&fout, sms * 1024 * sizeof(float)); cudaMalloc(&cyc, 1024 * sizeof(long long));
  printf("%s, %d SMs, %d MHz\n", p.name, sms, p.clockRate / 1000);
  for (int w : {1, 2, 3, 4, 8})
  {
    run<0, 0, 8>("pure", w, out, fout, cyc, sms);
    run<0, 0, 4>("pure", w, out, fout, cyc, sms);
    run<0, 0, 2>("pure", w, out, fout, cyc, sms);
    run<0, 0, 1>("pure", w, out, fout, cyc, sms);
    run<4, 0, 8>("ffma", w, out, fout, cyc, sms);
    run<8, 0, 8>("ffma", w, out, fout, cyc, sms);
    run<16, 0, 8>("ffma", w, out, fout, cyc, sms);
    run<8, 1, 8>("imad", w, out, fout, cyc, sms);
    run<16, 1, 8>("imad", w, out, fout, cyc, sms);
    run<4, 2, 8>("lds", w, out, fout, cyc, sms);
    run<8, 2, 8>("lds", w, out, fout, cyc, sms);
  }
  return 0;
}
