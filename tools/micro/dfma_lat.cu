// tools/micro/dfma_lat.cu -- fp64 pipe of one SM sub-partition on B200: latency
// of a dependent DFMA chain and throughput with C independent chains per warp,
// for 1..4 warps per sub-partition (block = 128 * W threads on one SM).
// Also: LDS.128 + STS.128 round trip through shared memory and bar.sync cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dfma_lat dfma_lat.cu
#include <cstdio>
#include <cstdint>

template <int C>
__global__ void dfma_kernel(double *out, unsigned long long *cyc, int iters) {
  double a[C];
#pragma unroll
  for (int c = 0; c < C; c++) a[c] = 1.0 + threadIdx.x * 1e-9 + c;
  const double m = 1.0000001, b = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int c = 0; c < C; c++) a[c] = fma(a[c], m, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; c++) s += a[c];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = (unsigned long long)(t1 - t0);
}

__global__ void smem_kernel(double2 *out, unsigned long long *cyc, int iters) {
  __shared__ double2 buf[1024];
  double2 v[16];
  const int t = threadIdx.x & 63;
#pragma unroll
  for (int q = 0; q < 16; q++) v[q] = make_double2(t + q, t - q);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int q = 0; q < 16; q++) buf[q * 64 + t] = v[q];
    asm volatile("bar.sync 1, 64;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 16; q++) v[q] = buf[((q * 5 + 1) & 15) * 64 + (t ^ 1)];
    asm volatile("bar.sync 1, 64;" ::: "memory");
  }
  const long long t1 = clock64();
  double2 s = make_double2(0, 0);
#pragma unroll
  for (int q = 0; q < 16; q++) { s.x += v[q].x; s.y += v[q].y; }
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = (unsigned long long)(t1 - t0);
}

template <int C> void run(int warps_per_smsp, double *out, unsigned long long *cyc) {
  const int iters = 2000;
  unsigned long long h;
  for (int rep = 0; rep < 2; rep++) dfma_kernel<C><<<1, 128 * warps_per_smsp>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (iters * 8.0 * C);
  printf("  %d warp(s)/sub-partition, %2d independent chains: %6.2f cycles per DFMA per warp (%5.2f per SMSP-issued DFMA)\n",
         warps_per_smsp, C, per, per / warps_per_smsp);
}

int main() {
  double *out; unsigned long long *cyc;
  cudaMalloc(&out, 8 * 1024 * 2); cudaMalloc(&cyc, 8);
  printf("DFMA (C = 1 is the dependent-issue latency):\n");
  for (int w = 1; w <= 4; w *= 2) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); run<16>(w, out, cyc); }
  unsigned long long h;
  for (int rep = 0; rep < 2; rep++) smem_kernel<<<1, 64>>>((double2 *)out, cyc, 1000);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("exchange round (16 STS.128, bar.sync 64, 16 LDS.128, bar.sync 64), 2 warps alone on the SM: %.0f cycles\n", (double)h / 1000);
  return 0;
}
