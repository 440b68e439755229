// micro-benchmark: DFMA dependent-issue latency and throughput per SM sub-partition on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, long long *cyc)
{
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double b = 1.0000001, c = 1e-12;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int warps)
{
  double *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<ILP><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  k<ILP><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / iters;   // cycles per loop iteration (ILP dfma per warp)
  printf("warps/SM %2d (per SMSP %4.1f) ILP %d: %.2f cyc/iter -> %.3f DFMA warp-inst/cyc/SMSP\n", warps, warps / 4.0, ILP, per, ILP * (warps / 4.0) / per);
  cudaFree(out); cudaFree(cyc);
}
int main()
{
  for (int w : {4, 8, 12, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
