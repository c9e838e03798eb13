// Checks the fragment layout of mma.sync.m8n8k4.f64 assumed by kernels_blo_aa.cuh:
// A[row = lane / 4][col = lane % 4], B[k = lane % 4][n = lane / 4], C[row = lane / 4][col = 2 (lane % 4) + {0, 1}].
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const double * A, const double * B, double * C)
{
  const int lane = threadIdx.x;
  double c[2] = {0.0, 0.0};
  const double a = A[(lane >> 2) * 4 + (lane & 3)], b = B[(lane & 3) * 8 + (lane >> 2)];
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
  C[(lane >> 2) * 8 + 2 * (lane & 3)] = c[0];
  C[(lane >> 2) * 8 + 2 * (lane & 3) + 1] = c[1];
}
int main()
{
  double hA[32], hB[32], hC[64], ref[64] = {};
  for (int i = 0; i < 32; ++i) { hA[i] = 1 + i * 0.37; hB[i] = 2 - i * 0.11; }
  for (int r = 0; r < 8; ++r) for (int n = 0; n < 8; ++n) for (int kk = 0; kk < 4; ++kk) ref[r * 8 + n] += hA[r * 4 + kk] * hB[kk * 8 + n];
  double * dA, * dB, * dC;
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  k<<<1, 32>>>(dA, dB, dC);
  cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost);
  double worst = 0; for (int i = 0; i < 64; ++i) { double d = hC[i] - ref[i]; if (d < 0) d = -d; if (d > worst) worst = d; }
  printf("dmma layout check: max |diff| = %g (%s)\n", worst, worst < 1e-12 ? "layout as assumed" : "LAYOUT MISMATCH");
  return worst < 1e-12 ? 0 : 1;
}
