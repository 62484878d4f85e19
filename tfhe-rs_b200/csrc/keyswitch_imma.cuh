// keyswitch_imma.cuh -- batched LWE keyswitch on the int8 tensor cores.
//
//   out[s][o] = [o == n_out] * b_in[s] - sum_{i < n_in} sum_{j < l}
//               digit_j(a_in[s][i]) * KSK[i][j][o]            (wrapping u64)
//
// (keyswitch_lwe_ciphertext, tfhe/src/core_crypto/algorithms/
// lwe_keyswitch.rs:137-232; replaces the reference's `tgemm_all_levels*`
// keyswitch GEMM, backends/tfhe-cuda-backend/cuda/src/crypto/keyswitch.cuh,
// which runs u64 multiplies on the CUDA cores.)
//
// This IS a GEMM -- digits[S x K] times KSK[K x (n_out+1)], K = n_in * l -- and
// it is exact on 8-bit integer tensor cores without touching the key:
//   * a KSK word is sum_b byte_b * 2^(8b), so the verbatim host-layout key
//     KSK[K][n_out+1] of u64 *is* a row-major u8 matrix Kb[K][8*(n_out+1)]
//     (little endian): no key conversion, no copy;
//   * the decomposition digits are in [-B/2, B/2] and fit s8 for base_log <= 7;
//   * P[s][8o + b] = sum_k digit[s][k] * Kb[k][8o + b] accumulates exactly in
//     s32 (|P| <= K * 255 * B/2 < 2^31, checked by the launcher), and
//     out[s][o] = [..] - sum_b P[s][8o + b] * 2^(8b)  (mod 2^64).
// For P22 that is 4096 x 8192 x 7352 int8 MACs instead of 3.1e10 u64 MACs on
// the fp64/integer pipes.
//
// Kernel: mma.sync.m16n8k32 (s8 x u8 -> s32), CTA tile 128 samples x 256 key
// bytes (32 output words) x 64 k, 8 warps of 64 x 64, 4-stage cp.async ring.
// The key tile arrives n-contiguous (as it lies in memory) while the MMA wants
// 4 consecutive k per register: each lane loads 4 k-rows x 4 bytes and
// transposes 4x4 bytes with PRMT, which yields its B fragments of FOUR n8
// tiles whose columns are the interleave n = 4*g + q (q = tile).  With that
// column order the 8 byte planes of one output word land in ONE lane (tile q:
// bytes q and 4+q), so the epilogue recombines without shuffles.
#pragma once
#include "pbs_generic_phases.cuh"

#include <cuda_runtime.h>

namespace b200 {

constexpr int KI_BM = 128;    // samples per CTA
// key bytes per CTA = 128 * NB (NB = 32-byte column blocks per warp, 1 or 2)
constexpr int KI_BK = 64;     // k per stage
constexpr int KI_STAGES = 4;
constexpr int KI_THREADS = 256;
constexpr int KI_A_STAGE = KI_BM * KI_BK; // 8 KiB
template <int NB> struct KiCfg {
  static constexpr int BN = 128 * NB;
  static constexpr int B_STAGE = KI_BK * BN; // 8 / 16 KiB
  static constexpr int SMEM = KI_STAGES * (KI_A_STAGE + B_STAGE);
};

// digit pre-pass: D[s][i * l + j] = digit_j(a_in[s][i]) as s8 (j = 0 is level
// l, the order of the KSK slots).  D is [rows_pad][k_pad], zero-initialised by
// the launcher (padding rows / columns stay 0).
__global__ void __launch_bounds__(256)
ks_digits_kernel(int8_t *__restrict__ D, const uint64_t *__restrict__ lwe_in,
                 const uint64_t *__restrict__ in_idx, uint32_t n_in,
                 uint32_t base_log, uint32_t l, uint32_t k_pad) {
  const uint32_t s = blockIdx.x; // samples on x: no 65,535 limit
  const uint32_t i = blockIdx.y * 256 + threadIdx.x;
  if (i >= n_in)
    return;
  uint64_t st = decomp_init_state(lwe_in[(in_idx ? in_idx[s] : (uint64_t)s) * (uint64_t)(n_in + 1) + i],
                                  base_log, l);
  int8_t *row = D + (size_t)s * k_pad + (size_t)i * l;
  for (uint32_t j = 0; j < l; j++)
    row[j] = (int8_t)decomp_next_digit(&st, base_log);
}

__device__ __forceinline__ void ki_cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void ki_cp_async8(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void ki_commit() {
  asm volatile("cp.async.commit_group;");
}
template <int N> __device__ __forceinline__ void ki_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ void ki_mma(int32_t (&c)[4], const uint32_t (&a)[4],
                                       uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 "
               "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// grid = (rows_pad / 128, ceil(8 * (n_out + 1) / BN)), block = 256,
// dynamic smem = KiCfg<NB>::SMEM.  NB = 2: 64 x 64 warp tiles, 128 accumulator
// registers, one CTA per SM; NB = 1: 64 x 32 warp tiles, two CTAs per SM.
template <int NB>
__global__ void __launch_bounds__(KI_THREADS, NB == 2 ? 1 : 2)
keyswitch_imma_kernel(uint64_t *__restrict__ lwe_out,
                      const uint64_t *__restrict__ out_idx,
                      const uint64_t *__restrict__ lwe_in,
                      const uint64_t *__restrict__ in_idx,
                      const uint8_t *__restrict__ ksk_bytes,
                      const int8_t *__restrict__ D, uint32_t n_in,
                      uint32_t n_out, uint32_t l, uint32_t k_pad,
                      uint32_t count) {
  extern __shared__ __align__(128) unsigned char ki_smem[];
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(ki_smem);
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + KI_STAGES * KI_A_STAGE;
  constexpr int KI_BN = KiCfg<NB>::BN;
  constexpr int KI_B_STAGE = KiCfg<NB>::B_STAGE;
  constexpr int WN = 32 * NB; // byte columns per warp

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3; // 2 x 4 warps of 64 x 64
  const uint32_t m0 = blockIdx.x * KI_BM;
  const uint32_t n0 = blockIdx.y * KI_BN;
  const uint32_t out_len = n_out + 1;
  const uint32_t n_bytes = out_len * 8; // key row length in bytes
  const uint32_t K = n_in * l;
  const uint32_t ksteps = k_pad / KI_BK;

  // ---- producers -------------------------------------------------------
  // A: 128 rows x 64 B = 512 16-byte chunks, 2 per thread
  // B: 64 rows x BN B = 1024 * NB 8-byte chunks, 4 * NB per thread
  auto load_stage = [&](uint32_t kstep, int stage) {
    const uint32_t k0 = kstep * KI_BK;
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const int idx = tid + c * KI_THREADS;
      const int row = idx >> 2, chunk = idx & 3;
      const uint32_t dst = a_base + stage * KI_A_STAGE + row * KI_BK +
                           ((chunk ^ ((row >> 1) & 3)) << 4);
      ki_cp_async16(dst, D + (size_t)(m0 + row) * k_pad + k0 + chunk * 16);
    }
#pragma unroll
    for (int c = 0; c < 4 * NB; c++) {
      const int idx = tid + c * KI_THREADS;
      const int row = idx / (KI_BN / 8), c8 = idx % (KI_BN / 8);
      const uint32_t col = (uint32_t)c8 * 8;
      const uint32_t dst = b_base + stage * KI_B_STAGE + row * KI_BN +
                           (col ^ (((row >> 2) & 3) << 5));
      const uint32_t k = k0 + row, n = n0 + col;
      if (k < K && n < n_bytes)
        ki_cp_async8(dst, ksk_bytes + (size_t)k * n_bytes + n);
      else
        asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(dst), "r"(0u));
    }
  };

  int32_t acc[4][4 * NB][4];
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int nt = 0; nt < 4 * NB; nt++)
#pragma unroll
      for (int c = 0; c < 4; c++)
        acc[mi][nt][c] = 0;

#pragma unroll
  for (int s = 0; s < KI_STAGES - 1; s++) {
    if ((uint32_t)s < ksteps)
      load_stage(s, s);
    ki_commit();
  }

  for (uint32_t kstep = 0; kstep < ksteps; kstep++) {
    ki_wait<KI_STAGES - 2>();
    __syncthreads();
    {
      const uint32_t nxt = kstep + KI_STAGES - 1;
      if (nxt < ksteps)
        load_stage(nxt, nxt % KI_STAGES);
      ki_commit();
    }
    const int stage = kstep % KI_STAGES;
    const uint32_t a_st = a_base + stage * KI_A_STAGE;
    const uint32_t b_st = b_base + stage * KI_B_STAGE;
#pragma unroll
    for (int ks = 0; ks < KI_BK / 32; ks++) {
      uint32_t a[4][4];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) {
        const int q = lane >> 3, r = lane & 7;
        const int row = wm * 64 + mi * 16 + r + (q & 1) * 8;
        const int chunk = ks * 2 + (q >> 1);
        const uint32_t addr =
            a_st + row * KI_BK + ((chunk ^ ((row >> 1) & 3)) << 4);
        asm volatile(
            "ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
            : "=r"(a[mi][0]), "=r"(a[mi][1]), "=r"(a[mi][2]), "=r"(a[mi][3])
            : "r"(addr));
      }
#pragma unroll
      for (int nb = 0; nb < NB; nb++) {
        uint32_t bf[2][4]; // [k half][tile q]
#pragma unroll
        for (int h = 0; h < 2; h++) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int row = ks * 32 + 16 * h + 4 * t + e;
            const uint32_t col = (uint32_t)(wn * WN + nb * 32 + 4 * g);
            const uint32_t addr = b_st + row * KI_BN + (col ^ ((uint32_t)t << 5));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[e]) : "r"(addr));
          }
          const uint32_t lo01 = __byte_perm(w[0], w[1], 0x5140);
          const uint32_t hi01 = __byte_perm(w[0], w[1], 0x7362);
          const uint32_t lo23 = __byte_perm(w[2], w[3], 0x5140);
          const uint32_t hi23 = __byte_perm(w[2], w[3], 0x7362);
          bf[h][0] = __byte_perm(lo01, lo23, 0x5410);
          bf[h][1] = __byte_perm(lo01, lo23, 0x7632);
          bf[h][2] = __byte_perm(hi01, hi23, 0x5410);
          bf[h][3] = __byte_perm(hi01, hi23, 0x7632);
        }
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
          for (int q = 0; q < 4; q++)
            ki_mma(acc[mi][nb * 4 + q], a[mi], bf[0][q], bf[1][q]);
      }
    }
  }
  ki_wait<0>();

  // ---- epilogue: recombine the 8 byte planes of each word ---------------
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const uint32_t s = m0 + wm * 64 + mi * 16 + g + half * 8;
      if (s >= count)
        continue;
      const uint64_t in_row = (in_idx ? in_idx[s] : (uint64_t)s) * (uint64_t)(n_in + 1);
      uint64_t *out_row = lwe_out + (out_idx ? out_idx[s] : (uint64_t)s) * (uint64_t)out_len;
#pragma unroll
      for (int nb = 0; nb < NB; nb++) {
        const uint32_t o = blockIdx.y * (KI_BN / 8) + wn * (WN / 8) + nb * 4 + t;
        if (o >= out_len)
          continue;
        uint64_t v = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          v += (uint64_t)(int64_t)acc[mi][nb * 4 + q][half * 2 + 0] << (8 * q);
          v += (uint64_t)(int64_t)acc[mi][nb * 4 + q][half * 2 + 1]
               << (8 * (q + 4));
        }
        const uint64_t body = (o == n_out) ? lwe_in[in_row + n_in] : 0;
        out_row[o] = body - v;
      }
    }
}

} // namespace b200
