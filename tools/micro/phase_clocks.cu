// tools/micro/phase_clocks.cu -- where do the cycles of ONE CMUX step go?
// Replays the blind-rotation loop of pbs_n2048_k1_l1_v3_kernel (same phase
// functions, same order) with a clock64() read between phases, for 1..2 CTAs
// per SM, and prints the average cycles per phase of warp 0.  Measurement
// tool only (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../tfhe-rs_b200/csrc -o phase_clocks phase_clocks.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "pbs_n2048.cuh"
using namespace b200;

#define NPH 16
__device__ __forceinline__ long long clk() {
  long long c;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory");
  return c;
}

__global__ void __launch_bounds__(128, 2)
timed_kernel(const cplx *__restrict__ bsk, const Fft1024Tables *__restrict__ tables,
             uint32_t n, uint32_t base_log, unsigned long long *out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  P22SmemV3 &sm = *reinterpret_cast<P22SmemV3 *>(smem_raw);
  const int tid = threadIdx.x, g = tid >> 6, t = tid & 63;
  for (uint32_t i = tid; i < n; i += 128)
    sm.a_hat[i] = (uint16_t)(1 + ((i * 2654435761u + blockIdx.x * 40503u) % 4095u));
  for (uint32_t j = tid; j < 2 * P22_N; j += 128)
    sm.acc[j >> 11][j & (P22_N - 1)] = j * 2654435761u + blockIdx.x;
  cplx tw2[3], tw3[15];
#pragma unroll
  for (int e = 0; e < 3; e++) tw2[e] = tables->pass2[t >> 2][e];
#pragma unroll
  for (int e = 0; e < 15; e++) tw3[e] = tables->pass3[t][e];
  __syncthreads();
  uint32_t *acc_g = sm.acc[g];
  cplx *xa_g = sm.xa[g], *xb_g = sm.xb[g];
  const cplx *xa_other = sm.xa[1 - g];
  const cplx *bsk_own = bsk + (size_t)g * (2 * P22_M) + (size_t)g * P22_M + t;
  const cplx *bsk_oth = bsk + (size_t)g * (2 * P22_M) + (size_t)(1 - g) * P22_M;
  long long acc[NPH];
  for (int p = 0; p < NPH; p++) acc[p] = 0;
  long long t0, t1;
#define LAP(p) t1 = clk(); acc[p] += t1 - t0; t0 = t1;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t a = sm.a_hat[i];
    const size_t step = (size_t)i * (4 * P22_M);
    cplx v[16], b_own[16];
    t0 = clk();
    p22v3_load_digits(acc_g, t, a, base_log, v);        LAP(0)
    radix16_fwd(v, c_fft1024_pass1);                    LAP(1)
    x1_store_p1(xa_g, t, v); group_barrier(g); x1_load_p2(xa_g, t, v); LAP(2)
    pass2_fwd(v, tw2);                                  LAP(3)
    x2_store_p2(xb_g, t, v); group_barrier(g); x2_load_p3(xb_g, t, v); LAP(4)
    radix16_fwd(v, tw3);                                LAP(5)
#pragma unroll
    for (int b = 0; b < 16; b++) b_own[b] = ldcg_cplx(bsk_own + step + b * 64);
    spec_store(xa_g, t, v); __syncthreads();            LAP(6)
    p22v3_mac(v, b_own, xa_other, bsk_oth + step, t, LdcgLoader()); LAP(7)
    __syncthreads();                                    LAP(8)
    radix16_inv(v, tw3);                                LAP(9)
    x2_store_p3(xb_g, t, v); group_barrier(g); x2_load_p2(xb_g, t, v); LAP(10)
    pass2_inv(v, tw2);                                  LAP(11)
    x1_store_p2(xa_g, t, v); group_barrier(g); x1_load_p1(xa_g, t, v); LAP(12)
    radix16_inv(v, c_fft1024_pass1);                    LAP(13)
    p22v2_acc_update(acc_g, t, v); group_barrier(g);    LAP(14)
  }
  if (tid == 0 && blockIdx.x == 0)
    for (int p = 0; p < NPH; p++) out[p] = (unsigned long long)acc[p];
  if (sm.acc[0][tid] == 0x12345678u) out[NPH] = 1; // keep results live
}

// the same replay for the v6 kernel: exchange 2 through tensor memory, key
// block of the step from the one-slot TMA ring (pbs_n2048_k1_l1_v6_kernel<0>)
__global__ void __launch_bounds__(128, 2)
timed_kernel_v6(const cplx *__restrict__ bsk, const Fft1024Tables *__restrict__ tables,
                uint32_t n, uint32_t base_log, unsigned long long *out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  P22SmemV6<4, false> &sm = *reinterpret_cast<P22SmemV6<4, false> *>(smem_raw);
  const int tid = threadIdx.x, g = tid >> 6, t = tid & 63;
  if (tid < 32) tmem_alloc(&sm.tmem_base, 64);
  if (tid == 0) { mbar_init(&sm.bar, 1); mbar_fence_init(); }
  for (uint32_t j = tid; j < 2 * P22_N; j += 128)
    sm.acc[j >> 11][j & (P22_N - 1)] = j * 2654435761u + blockIdx.x;
  cplx tw2[3], tw3[15];
#pragma unroll
  for (int e = 0; e < 3; e++) tw2[e] = tables->pass2[x1t_q(t)][e];
#pragma unroll
  for (int e = 0; e < 15; e++) tw3[e] = tables->pass3[t][e];
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tmw = sm.tmem_base + ((uint32_t)((tid >> 5) * 32) << 16);
  auto tma_issue = [&](uint32_t i) {
    mbar_arrive_expect_tx(&sm.bar, 4u * P22_M * (uint32_t)sizeof(cplx));
#pragma unroll
    for (int q = 0; q < 4; q++)
      tma_bulk_g2s(&sm.ring[q][0], bsk + (size_t)i * (4 * P22_M) + (size_t)q * P22_M,
                   P22_M * (uint32_t)sizeof(cplx), &sm.bar);
  };
  if (tid == 0) tma_issue(0);
  uint32_t *acc_g = sm.acc[g];
  cplx *xa_g = sm.xa[g];
  const cplx *xa_other = sm.xa[1 - g];
  const cplx *k_own = &sm.ring[2 * g + g][0];
  const cplx *k_oth = &sm.ring[2 * g + (1 - g)][0];
  uint32_t own[32];
  p22v4_own_init(acc_g, t, own);
  long long acc[NPH];
  for (int p = 0; p < NPH; p++) acc[p] = 0;
  long long t0, t1;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t a = 1 + ((i * 2654435761u + blockIdx.x * 40503u) % 4095u);
    cplx v[16];
    t0 = clk();
    p22v4_load_digits(acc_g, t, a, base_log, own, v);   LAP(0)
    radix16_fwd(v, c_fft1024_pass1);                    LAP(1)
    x1t_store_p1(xa_g, t, v); group_barrier(g); x1t_load_p2(xa_g, t, v); LAP(2)
    pass2_fwd(v, tw2); spec_guard_arrive(g, t);         LAP(3)
    x2t_store_p2(tmw, v); x2t_load_p3(tmw, v);          LAP(4)
    radix16_fwd(v, tw3);                                LAP(5)
    spec_guard_wait(g, t); spec_store(xa_g, t, v); __syncthreads(); LAP(6)
    mbar_wait_parity(&sm.bar, i & 1u);
#pragma unroll
    for (int b = 0; b < 16; b++)
      v[b] = cfma(xa_other[b * 64 + t], k_oth[b * 64 + t], cmul(v[b], k_own[b * 64 + t]));
    LAP(7)
    __syncthreads();
    if (tid == 0 && i + 1 < n) tma_issue(i + 1);
    LAP(8)
    radix16_inv(v, tw3);                                LAP(9)
    x2t_store_p3(tmw, v); x2t_load_p2(tmw, v);          LAP(10)
    pass2_inv(v, tw2);                                  LAP(11)
    x1t_store_p2(xa_g, t, v); group_barrier(g); x1t_load_p1(xa_g, t, v); LAP(12)
    radix16_inv(v, c_fft1024_pass1);                    LAP(13)
    p22v4_acc_update(acc_g, t, v, own); group_barrier(g); LAP(14)
  }
  if (tid == 0 && blockIdx.x == 0)
    for (int p = 0; p < NPH; p++) out[p] = (unsigned long long)acc[p];
  if (sm.acc[0][tid] == 0x12345678u) out[NPH] = 1;
  tmem_fence_before_sync();
  __syncthreads();
  if (tid < 32) { tmem_fence_after_sync(); tmem_dealloc(sm.tmem_base, 64); }
}

int main(int argc, char **argv) {
  const uint32_t n = 918;
  const int ctas = argc > 1 ? atoi(argv[1]) : 1;
  Fft1024Tables *host = new Fft1024Tables;
  b200_fill_fft1024_tables(host);
  Fft1024Tables *dt;
  cudaMalloc(&dt, sizeof(*host));
  cudaMemcpy(dt, host, sizeof(*host), cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(c_fft1024_pass1, host->pass1, sizeof(host->pass1));
  const size_t words = (size_t)n * 4 * P22_M;
  std::vector<cplx> h(words);
  for (size_t i = 0; i < words; i++) { h[i].re = 1e-9 * (double)((i * 2654435761u) % 1000) ; h[i].im = -h[i].re; }
  cplx *bsk; cudaMalloc(&bsk, words * sizeof(cplx));
  cudaMemcpy(bsk, h.data(), words * sizeof(cplx), cudaMemcpyHostToDevice);
  unsigned long long *out; cudaMalloc(&out, (NPH + 1) * 8); cudaMemset(out, 0, (NPH + 1) * 8);
  const bool v6 = argc > 2 && atoi(argv[2]) == 6;
  cudaFuncSetAttribute(timed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P22SmemV3));
  cudaFuncSetAttribute(timed_kernel_v6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P22SmemV6<4, false>));
  for (int rep = 0; rep < 2; rep++) {
    if (v6) timed_kernel_v6<<<ctas, 128, sizeof(P22SmemV6<4, false>)>>>(bsk, dt, n, 23, out);
    else timed_kernel<<<ctas, 128, sizeof(P22SmemV3)>>>(bsk, dt, n, 23, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  }
  unsigned long long hout[NPH + 1];
  cudaMemcpy(hout, out, sizeof(hout), cudaMemcpyDeviceToHost);
  const char *names[NPH] = {"load_digits", "fwd pass1 (radix16)", "exchange 1", "fwd pass2 (radix4)", "exchange 2",
    "fwd pass3 (radix16)", "key prefetch issue + spectrum share + syncthreads", "MAC", "syncthreads", "inv pass3",
    "exchange 2'", "inv pass2", "exchange 1'", "inv pass1", "acc update + barrier", ""};
  unsigned long long tot = 0;
  for (int p = 0; p < 15; p++) tot += hout[p];
  printf("%s, CTAs %d: %.0f cycles per CMUX step (warp 0 of CTA 0)\n", v6 ? "v6 (tensor-memory exchange 2, TMA key ring)" : "v3", ctas, (double)tot / n);
  for (int p = 0; p < 15; p++) printf("  %-52s %7.0f  %5.1f %%\n", names[p], (double)hout[p] / n, 100.0 * hout[p] / tot);
  return 0;
}
