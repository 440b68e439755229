// micro-benchmark: can other instructions issue in the shadow of a half-rate DFMA on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NF>
__global__ void k(double *out, int iters, long long *cyc)
{
  double a[ND > 0 ? ND : 1]; float f[NF > 0 ? NF : 1];
  for (int i = 0; i < ND; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for (int i = 0; i < NF; ++i) f[i] = 1.0f + 1e-3f * (threadIdx.x + i);
  const double b = 1.0000001, c = 1e-12; const float fb = 1.0001f, fc = 1e-6f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < ND; ++i) a[i] = fma(a[i], b, c);
#pragma unroll
      for (int i = 0; i < NF; ++i) f[i] = fmaf(f[i], fb, fc);
    }
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < ND; ++i) s += a[i]; for (int i = 0; i < NF; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ND, int NF> void run(int warps)
{
  double *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 2048;
  k<ND, NF><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  k<ND, NF><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 4);   // cycles per (ND dfma + NF ffma) group per warp
  double wps = warps / 4.0;
  printf("warps/SMSP %.0f: %d DFMA + %d FFMA per group: %.2f cyc/group/warp -> per SMSP per cycle: %.3f DFMA, %.3f FFMA, %.3f total\n",
         wps, ND, NF, per, ND * wps / per, NF * wps / per, (ND + NF) * wps / per);
  cudaFree(out); cudaFree(cyc);
}
int main()
{
  run<4, 0>(16); run<0, 4>(16); run<4, 4>(16); run<4, 8>(16); run<2, 8>(16); run<4, 12>(16);
  return 0;
}
