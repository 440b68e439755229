// How much HBM bandwidth does a streaming kernel get when its bytes are spread over N separate arrays (the thermal step kernel reads
// 13 arrays and writes 1)?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o streams streams.cu && ./streams
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void __launch_bounds__(128) rd(const double *const *in, double *out, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) s += in[k][i];
  out[i] = s;
}
template <int N> void run(double **d_ptrs, double **h_ptrs, double *out, long long n)
{
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int blocks = (int)((n + 127) / 128);
  for (int w = 0; w < 3; ++w) rd<N><<<blocks, 128>>>(d_ptrs, out, n);
  cudaEventRecord(a);
  for (int r = 0; r < 10; ++r) rd<N><<<blocks, 128>>>(d_ptrs, out, n);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
  printf("%2d input arrays + 1 output: %.3f ms, %.0f GB/s\n", N, ms, (N + 1) * 8.0 * n / (ms * 1e-3) / 1e9);
}
int main()
{
  const long long n = 15LL << 20;                      // 1 Mi columns x 15 cells
  double *h_ptrs[16], **d_ptrs, *out;
  for (int k = 0; k < 16; ++k) { cudaMalloc(&h_ptrs[k], n * 8); cudaMemset(h_ptrs[k], 0, n * 8); }
  cudaMalloc(&out, n * 8); cudaMalloc(&d_ptrs, sizeof(h_ptrs)); cudaMemcpy(d_ptrs, h_ptrs, sizeof(h_ptrs), cudaMemcpyHostToDevice);
  run<1>(d_ptrs, h_ptrs, out, n); run<2>(d_ptrs, h_ptrs, out, n); run<4>(d_ptrs, h_ptrs, out, n); run<8>(d_ptrs, h_ptrs, out, n);
  run<13>(d_ptrs, h_ptrs, out, n); run<16>(d_ptrs, h_ptrs, out, n);
  return 0;
}
