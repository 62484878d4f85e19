// tools/micro/pipes.cu -- which B200 pipes can carry part of the blind rotation
// next to the fp64 FMA pipe and the shared-memory crossbar (round 2 questions):
//   (a) DMMA (mma.sync m8n8k4 f64) alone, DFMA alone, both interleaved in one
//       warp and in different warps of one sub-partition: distinct pipes?
//   (b) TMEM as a per-thread scratch: tcgen05.st / tcgen05.ld 32x32b round trip,
//       1 / 4 / 8 warps, and concurrently with an LDS/STS exchange stream;
//   (c) SHFL.BFLY throughput against LDS/STS.128 for the quad-local exchange;
//   (d) integer ALU (LOP3 / IADD3 / SHF / IMAD) issue rate per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pipes pipes.cu
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

// ---------------------------------------------------------------- (a) DMMA
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode 0: DFMA only (8 chains); 1: DMMA only (4 chains); 2: both interleaved in
// every warp; 3: even warps DFMA, odd warps DMMA
__global__ void dmma_kernel(double *out, unsigned long long *cyc, int iters, int mode) {
  double f[8], c0[4], c1[4];
#pragma unroll
  for (int i = 0; i < 8; i++) f[i] = 1.0 + threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 4; i++) { c0[i] = i; c1[i] = -i; }
  const double a = 1.0000001, b = 1e-9;
  const int w = threadIdx.x >> 5;
  const bool do_f = mode == 0 || mode == 2 || (mode == 3 && (w & 4) == 0);
  const bool do_m = mode == 1 || mode == 2 || (mode == 3 && (w & 4) != 0);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (do_f) {
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] = fma(f[k], a, b);
      }
      if (do_m) {
#pragma unroll
        for (int k = 0; k < 4; k++) dmma(c0[k], c1[k], a, b);
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += f[i];
#pragma unroll
  for (int i = 0; i < 4; i++) s += c0[i] + c1[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = (unsigned long long)(t1 - t0);
}

// ---------------------------------------------------------------- (b) TMEM
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define TM_ST16(taddr, r)                                                                          \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")
#define TM_LD16(taddr, r)                                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),      \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
               : "r"(taddr) : "memory")

// mode 0: 4 x (st16) + wait + 4 x (ld16) + wait per iteration = 64 words / thread each way (the size of one exchange)
// mode 1: loads only (4 x ld16 + wait);  mode 2: stores only
// mode 3: TMEM round trip in warps 0-3, LDS/STS.128 exchange in warps 4-7
// mode 4: LDS/STS exchange only in warps 4-7 (baseline for mode 3)
// mode 5: correctness of the 32x32b round trip (lane-private scratch)
__global__ void tmem_kernel(uint32_t *out, unsigned long long *cyc, int iters, int mode, int ncols) {
  __shared__ uint32_t tbase_s;
  __shared__ double2 buf[2048];
  const int w = threadIdx.x >> 5;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tbase_s)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tbase_s;
  const uint32_t taddr = tbase + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)((w >> 2) * 64);
  uint32_t r[4][16];
#pragma unroll
  for (int q = 0; q < 4; q++)
#pragma unroll
    for (int i = 0; i < 16; i++) r[q][i] = threadIdx.x * 64 + q * 16 + i;
  double2 v[16];
#pragma unroll
  for (int q = 0; q < 16; q++) v[q] = make_double2(threadIdx.x + q, threadIdx.x - q);
  const bool tm = mode <= 2 || mode == 5 || (mode == 3 && w < 4);
  const bool ex = (mode == 3 || mode == 4) && w >= 4;
  const int t = threadIdx.x & 127;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (tm) {
      if (mode != 1) {
#pragma unroll
        for (int q = 0; q < 4; q++) TM_ST16(taddr + q * 16, r[q]);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      if (mode != 2) {
#pragma unroll
        for (int q = 0; q < 4; q++) TM_LD16(taddr + q * 16, r[q]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
          for (int i = 0; i < 16; i++) r[q][i] += (mode == 5 ? 0u : 1u);
      }
    }
    if (ex) {
#pragma unroll
      for (int q = 0; q < 16; q++) buf[q * 128 + t] = v[q];
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 16; q++) v[q] = buf[((q * 5 + 1) & 15) * 128 + (t ^ 1)];
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
  if (mode == 5) {
    // every word must read back as written
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int i = 0; i < 16; i++) s += (r[q][i] != threadIdx.x * 64 + q * 16 + i);
  } else {
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int i = 0; i < 16; i++) s += r[q][i];
    double acc = 0;
#pragma unroll
    for (int q = 0; q < 16; q++) acc += v[q].x + v[q].y;
    s += (uint32_t)acc;
  }
  out[threadIdx.x] = s;
  if ((threadIdx.x & 127) == 0) cyc[threadIdx.x >> 7] = (unsigned long long)(t1 - t0);
  __syncthreads();
  if (w == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(ncols) : "memory");
}

// ---------------------------------------------------------------- (c) SHFL
__global__ void shfl_kernel(uint32_t *out, unsigned long long *cyc, int iters) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = threadIdx.x * 16 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = __shfl_xor_sync(0xffffffffu, r[i], 1 + (i & 1));
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += r[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = (unsigned long long)(t1 - t0);
}

// ---------------------------------------------------------------- (d) integer ALU
template <int OP> __global__ void int_kernel(uint32_t *out, unsigned long long *cyc, int iters, uint32_t k) {
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = threadIdx.x * 8 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (OP == 0) r[i] = (r[i] ^ k) & (r[(i + 1) & 7] | 0x55u);          // LOP3
        if (OP == 1) r[i] = r[i] + k + r[(i + 3) & 7];                       // IADD3
        if (OP == 2) r[i] = __funnelshift_r(r[i], r[(i + 1) & 7], k & 31);   // SHF
        if (OP == 3) r[i] = r[i] * k + r[(i + 5) & 7];                       // IMAD
        if (OP == 4) r[i] = (uint32_t)((int32_t)r[i] >> (k & 31));           // SHF.R.S32
        if (OP == 5) r[i] = ((int32_t)r[i] < (int32_t)k) ? r[(i + 1) & 7] : r[i]; // ISETP + SEL
      }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += r[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = (unsigned long long)(t1 - t0);
}

// I2F.F64.S32 and the magic-add alternative
template <int OP> __global__ void i2f_kernel(double *out, unsigned long long *cyc, int iters) {
  int32_t r[8];
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { r[i] = threadIdx.x * 8 + i; acc[i] = 0; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        double d;
        if (OP == 0) d = (double)r[i];
        else d = __hiloint2double(0x43300000, r[i] ^ 0x80000000) ; // 2^52 + 2^31 + r, the subtract folds into the consumer
        acc[i] += d;
        r[i] += it;
      }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = (unsigned long long)(t1 - t0);
}

int main() {
  double *out; unsigned long long *cyc; unsigned long long h[4];
  CK(cudaMalloc(&out, 8 * 4096)); CK(cudaMalloc(&cyc, 64));
  const int iters = 2000;
  printf("(a) fp64: per iteration a warp issues 32 DFMA (mode 0), 16 DMMA m8n8k4 (mode 1), both (2), or split by warp (3)\n");
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 4; mode++) {
      if (mode == 3 && warps < 8) continue;
      for (int rep = 0; rep < 2; rep++) dmma_kernel<<<1, 32 * warps>>>(out, cyc, iters, mode);
      CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
      printf("  %2d warps, mode %d: %8.1f cycles / iteration\n", warps, mode, (double)h[0] / iters);
    }
  }
  printf("(b) TMEM 32x32b scratch, 64 words per thread each way per iteration\n");
  for (int rep = 0; rep < 2; rep++) tmem_kernel<<<1, 128>>>((uint32_t *)out, cyc, 1, 5, 64);
  {
    CK(cudaDeviceSynchronize());
    uint32_t bad[128]; CK(cudaMemcpy(bad, out, sizeof bad, cudaMemcpyDeviceToHost));
    uint32_t nb = 0; for (int i = 0; i < 128; i++) nb += bad[i];
    printf("  round-trip check: %u mismatching words\n", nb);
  }
  const char *names[] = {"st+wait+ld+wait", "ld+wait only", "st+wait only", "TMEM round trip (warps 0-3) next to an LDS/STS.128 exchange (warps 4-7)", "LDS/STS.128 exchange alone (warps 4-7)"};
  for (int mode = 0; mode < 5; mode++) {
    for (int threads : {32, 128, 256}) {
      if (mode >= 3 && threads != 256) continue;
      for (int rep = 0; rep < 2; rep++) tmem_kernel<<<1, threads>>>((uint32_t *)out, cyc, iters, mode, 128);
      CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
      printf("  %3d threads, %-70s: %8.1f cycles / iteration (second 128-thread group: %8.1f)\n", threads, names[mode], (double)h[0] / iters, threads > 128 ? (double)h[1] / iters : 0.0);
    }
  }
  printf("(c) SHFL.BFLY, 16 per thread per iteration\n");
  for (int warps : {1, 4, 8}) {
    for (int rep = 0; rep < 2; rep++) shfl_kernel<<<1, 32 * warps>>>((uint32_t *)out, cyc, iters);
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("  %d warps: %6.2f cycles per SHFL per warp, %6.2f per SM-issued SHFL\n", warps, (double)h[0] / (iters * 16.0), (double)h[0] / (iters * 16.0 * warps));
  }
  printf("(d) integer ALU, 32 ops per thread per iteration, cycles per warp instruction per sub-partition\n");
  const char *ops[] = {"LOP3 x2", "IADD3", "SHF (funnel)", "IMAD", "SHF.R.S32", "ISETP+SEL"};
  for (int warps : {4, 8}) {
    for (int op = 0; op < 6; op++) {
      for (int rep = 0; rep < 2; rep++) {
        if (op == 0) int_kernel<0><<<1, 32 * warps>>>((uint32_t *)out, cyc, iters, 3);
        if (op == 1) int_kernel<1><<<1, 32 * warps>>>((uint32_t *)out, cyc, iters, 3);
        if (op == 2) int_kernel<2><<<1, 32 * warps>>>((uint32_t *)out, cyc, iters, 3);
        if (op == 3) int_kernel<3><<<1, 32 * warps>>>((uint32_t *)out, cyc, iters, 3);
        if (op == 4) int_kernel<4><<<1, 32 * warps>>>((uint32_t *)out, cyc, iters, 3);
        if (op == 5) int_kernel<5><<<1, 32 * warps>>>((uint32_t *)out, cyc, iters, 3);
      }
      CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
      printf("  %d warps/SM, %-14s: %6.2f cycles per source-level op per SMSP\n", warps, ops[op], (double)h[0] / (iters * 32.0 * (warps / 4)));
    }
  }
  for (int op = 0; op < 2; op++) {
    for (int rep = 0; rep < 2; rep++) {
      if (op == 0) i2f_kernel<0><<<1, 128>>>(out, cyc, iters);
      else i2f_kernel<1><<<1, 128>>>(out, cyc, iters);
    }
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("  int32 -> double + DADD, %s: %6.2f cycles per conversion per SMSP (1 warp)\n", op == 0 ? "I2F.F64.S32" : "exponent splice (LOP + DADD)", (double)h[0] / (iters * 32.0));
  }
  return 0;
}
