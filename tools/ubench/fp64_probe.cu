// Micro-probe (not part of the product): FP64 FMA dependent-chain latency and per-SM throughput, LDS latency, on the GPU
// the build runs on.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double* out, long long* cyc, int n) {
  double a = out[0], b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) a = fma(a, b, c);   // one dependent chain
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_tput(double* out, long long* cyc, int n) {
  double a0 = out[0], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  __syncthreads();
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc, int n) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i * 17 + 5) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) p = nxt[p];   // dependent shared-memory loads
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 1 << 22); cudaMalloc(&c, 64); cudaMemset(d, 0, 1 << 22);
  long long h; const int n = 4096;
  k_lat<<<1, 32>>>(d, c, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("DFMA dependent chain: %.2f cycles per FMA (1 warp)\n", (double)h / n);
  for (int thr : {128, 256, 512, 1024}) {
    k_tput<<<1, thr>>>(d, c, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA throughput, 1 CTA of %4d threads, 8 chains/thread: %.2f FMA/clk/SM\n", thr, (double)thr * 8 * n / h);
  }
  k_lds<<<1, 32>>>(d, c, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("LDS dependent chain: %.2f cycles per load\n", (double)h / n);
  return 0;
}
