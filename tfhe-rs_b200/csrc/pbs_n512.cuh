// pbs_n512.cuh -- sm_100a register-FFT classic PBS for (N = 512, l = 1,
// k <= 4): PARAM_MESSAGE_1_CARRY_1_KS_PBS (n = 879, k = 4, log B = 23;
// tfhe/src/shortint/parameters/v1_0/classic/tuniform/p_fail_2_minus_128/
// ks_pbs.rs:11-21) and its relatives.  Round 1 and the reference's CUDA backend
// both run this set on generic shared-memory FFT kernels (16.5-17.5 k PBS/s on a
// B200, profiles/round2.md); this is the (2048, 1, 1) design re-cut for a
// transform of 256 complex points.
//
// Replaces host_programmable_bootstrap / device_programmable_bootstrap_*
// (backends/tfhe-cuda-backend/cuda/src/pbs/programmable_bootstrap_classic.cuh)
// for these shapes; restates FourierLweBootstrapKeyView::bootstrap
// (fft64/crypto/bootstrap.rs:294-380,480-520).
//
//   * M = 256 = 16 x 16: 16 threads per polynomial, 16 complex values per thread
//     in registers, ONE exchange per transform -- a 16 x 16 transpose inside a
//     half-warp, ordered by __syncwarp (negacyclic_fft.cuh, xq_*);
//   * one LWE = (k+1) polynomials = 16 (k+1) threads; G LWEs share a CTA and run
//     in lock step, so the Fourier key block of a step (n.b. 100 KiB for k = 4)
//     is requested by G thread groups at once and the second request hits L1;
//   * the accumulator polynomial p is only ever touched by its own 16 threads
//     (rotation stays inside a polynomial): u32 words in shared memory, the
//     thread's own 32 words handed over in registers (p22v4 scheme), warp-level
//     synchronisation; two CTA barriers per step, around the Fourier MAC;
//   * Fourier key layout: [i][column c][row r][b < 16][q < 16] complex128,
//     value at slot pos = 16 q + b, pre-scaled by 2^-64 / 256 * 2^32; thread
//     (c, q) reads 16 contiguous bytes per (r, b), a half-warp 256 contiguous.
#pragma once
#include "pbs_n2048.cuh" // ldcg_cplx, phases

#include <cuda_runtime.h>

namespace b200 {

#define P512_N 512
#define P512_M 256

__constant__ cplx c_fft256_pass1[15];

// read-only, L1-cacheable (small tables shared by every thread group)
__device__ __forceinline__ cplx ldnc_cplx(const cplx *p) {
  const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
  return cmake(v.x, v.y);
}

// RING: the Fourier key block of a step ((k+1)^2 x 4 KiB, contiguous in the
// layout below) is brought into shared memory by the TMA unit a whole step ahead
// (one slot, cp.async.bulk + mbarrier, as in pbs_n2048_k1_l1_v6_kernel) and is
// shared by the G LWEs of the CTA; one CTA per SM.  !RING: every thread streams
// its key values from L2 through a 3-chunk register ring; two CTAs per SM.
template <int K, int G, bool RING = false>
struct N512Smem {
  static constexpr int P = K + 1;
  cplx xs[G][P][P512_M];       // exchange buffer, then shared spectrum  4 KiB each
  cplx ring[RING ? P * P : 1][RING ? P512_M : 1];
  // rows padded by 16 words: the two polynomials of a warp then sit 16 banks
  // apart and their 32-bit rotated reads / updates do not collide
  uint32_t acc[G][P][P512_N + 16];
  uint16_t a_hat[G][1024 + 8];
  uint32_t b_hat[G];
  unsigned long long red_half[G][4];
  long long red_dbl[G][4];
  unsigned long long bar;
};

// rotate + decompose for N = 512 (p22v4_load_digits with stride 16): thread u
// holds complex coefficients j = 16*j1 + u (re <- coefficient j, im <- j + 256)
__device__ __forceinline__ void n512_load_digits(const uint32_t *acc_p, int u,
                                                 uint32_t a, uint32_t base_log,
                                                 const uint32_t own[32],
                                                 cplx v[16]) {
  const uint32_t d = a & (P512_N - 1);
  const bool neg0 = (a >> 9) != 0u; // a >= N
  const uint32_t half = 1u << (31 - base_log);
  const uint32_t sh = 32 - base_log;
  const int32_t tie_below = (int32_t)(0x80000000u + half);
  const int32_t plus_half_base = (int32_t)(1u << (base_log - 1));
  const uint32_t base4 = ((uint32_t)u - d) * 4u; // 4 * (j - d) for j1 = 0
  const unsigned char *accb = reinterpret_cast<const unsigned char *>(acc_p);
#pragma unroll
  for (int j1 = 0; j1 < 16; j1++) {
    const uint32_t ub0 = base4 + 64u * j1;       // 4 * (j - d)
    const uint32_t ub1 = ub0 + 4u * P512_M;      // 4 * (j + M - d)
    const uint32_t ib0 = ub0 & (4u * P512_N - 4u);
    const uint32_t r0 = *reinterpret_cast<const uint32_t *>(accb + ib0);
    const uint32_t r1 =
        *reinterpret_cast<const uint32_t *>(accb + (ib0 ^ (4u * P512_M)));
    const bool n0 = ((int32_t)ub0 < 0) != neg0;
    const bool n1 = ((int32_t)ub1 < 0) != neg0;
    const uint32_t x0 = n0 ? (0u - r0) - own[j1] : r0 - own[j1];
    const uint32_t x1 = n1 ? (0u - r1) - own[16 + j1] : r1 - own[16 + j1];
    int32_t d0 = (int32_t)(x0 + half) >> sh;
    int32_t d1 = (int32_t)(x1 + half) >> sh;
    // balanced tie (decomposer.rs:163-188): x in [2^31, 2^31 + half) -> +B/2
    if ((int32_t)x0 < tie_below)
      d0 = plus_half_base;
    if ((int32_t)x1 < tie_below)
      d1 = plus_half_base;
    v[j1] = cmake(int_to_double(d0), int_to_double(d1));
  }
}

__device__ __forceinline__ void n512_acc_update(uint32_t *acc_p, int u,
                                                const cplx v[16],
                                                uint32_t own[32]) {
#pragma unroll
  for (int j1 = 0; j1 < 16; j1++) {
    const uint32_t j = 16u * j1 + (uint32_t)u;
    // own[] is the current accumulator word: no read of shared memory
    own[j1] += scaled_double_to_torus32(v[j1].re);
    own[16 + j1] += scaled_double_to_torus32(v[j1].im);
    acc_p[j] = own[j1];
    acc_p[j + P512_M] = own[16 + j1];
  }
}

// persistent grid (see below), block = 16 (K+1) G
template <int K, int G, bool RING = false>
__global__ void __launch_bounds__(16 * (K + 1) * G, RING ? 1 : 2)
pbs_n512_kernel(uint64_t *__restrict__ lwe_out,
                const uint64_t *__restrict__ out_idx,
                const uint64_t *__restrict__ luts,
                const uint64_t *__restrict__ lut_idx,
                const uint64_t *__restrict__ lwe_in,
                const uint64_t *__restrict__ in_idx,
                const cplx *__restrict__ bsk,
                const Fft256Tables *__restrict__ tables, uint32_t n,
                uint32_t base_log, uint32_t num_samples, uint32_t num_many_lut,
                uint32_t lut_stride, int centered_ms) {
  constexpr int P = K + 1;
  constexpr int TPL = 16 * P; // threads per LWE
  constexpr int NT = TPL * G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Smem = N512Smem<K, G, RING>;
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x;
  [[maybe_unused]] auto tma_issue = [&](uint32_t i) {
    if constexpr (RING) {
      constexpr uint32_t bytes = (uint32_t)(P * P * P512_M * sizeof(cplx));
      constexpr uint32_t piece = (uint32_t)(P * P512_M * sizeof(cplx)); // one column
      mbar_arrive_expect_tx(&sm.bar, bytes);
#pragma unroll
      for (int c = 0; c < P; c++)
        tma_bulk_g2s(&sm.ring[c * P][0],
                     bsk + (size_t)i * (P * P * P512_M) + (size_t)c * (P * P512_M),
                     piece, &sm.bar);
    }
  };
  if constexpr (RING) {
    if (tid == 0) {
      mbar_init(&sm.bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
  }
  [[maybe_unused]] uint32_t ring_phase = 0; // completed copies so far (parity of the wait)
  const int lw = tid / TPL;          // LWE of this CTA
  const int tl = tid - lw * TPL;     // thread inside the LWE
  const int p = tl >> 4;             // polynomial (= output column in the MAC)
  const int u = tl & 15;             // thread inside the polynomial
  const uint32_t log_mod = 10;       // log2(2N)
  // PERSISTENT grid: at most two CTAs per SM are launched and every CTA walks the
  // sample list with the grid as stride.  All resident CTAs then sit at (nearly)
  // the same CMUX step at any time, so the slice of the Fourier key in use -- a
  // few steps x 100 KiB for k = 4 -- stays in L2.  The whole key of this set
  // (90 MB) does not fit one L2 partition: with one CTA per sample pair and
  // waves starting at different times the kernel streamed it from HBM again
  // for every wave (23.6 k PBS/s, profiles/round2.md).
  for (uint32_t grp = blockIdx.x; grp * G < num_samples; grp += gridDim.x) {
  // the last group may hold fewer than G samples: the spare thread groups replay
  // the last sample (all barriers are CTA wide) and do not write
  const uint32_t s_raw = grp * G + lw;
  const bool live = s_raw < num_samples;
  const uint32_t s = live ? s_raw : num_samples - 1;

  // ---- prologue: modulus switch (standard or centered mean) ---------------
  const uint64_t *ct = lwe_in + in_idx[s] * (uint64_t)(n + 1);
  if (tl == 0) {
    sm.red_half[lw][0] = 0;
    sm.red_dbl[lw][0] = 0;
  }
  __syncthreads();
  {
    unsigned long long half_sum = 0;
    long long dbl_sum = 0;
    for (uint32_t i = tl; i < n; i += TPL) {
      const uint64_t a = ct[i];
      sm.a_hat[lw][i] = (uint16_t)modulus_switch_u64(a, log_mod);
      if (centered_ms) {
        int64_t dd;
        half_sum += (unsigned long long)centered_ms_half_error(a, log_mod, &dd);
        dbl_sum += dd;
      }
    }
    // per-LWE reduction through shared memory (TPL is not a warp multiple)
    if (centered_ms) {
      atomicAdd(&sm.red_half[lw][0], half_sum);
      atomicAdd(reinterpret_cast<unsigned long long *>(&sm.red_dbl[lw][0]),
                (unsigned long long)dbl_sum);
    }
  }
  __syncthreads();
  if (tl == 0) {
    uint64_t body = ct[n];
    if (centered_ms) {
      uint64_t hs = sm.red_half[lw][0];
      const int64_t ds = sm.red_dbl[lw][0];
      hs -= (uint64_t)(ds / 2);
      body += hs - ((uint64_t)1 << (63 - log_mod));
    }
    sm.b_hat[lw] = modulus_switch_u64(body, log_mod);
  }
  __syncthreads();
  // ---- prologue: acc = LUT * X^{-b_hat} ------------------------------------
  {
    const uint64_t *lut = luts + lut_idx[s] * (uint64_t)(P * P512_N);
    const uint32_t b_hat = sm.b_hat[lw];
    for (uint32_t j = tl; j < (uint32_t)(P * P512_N); j += TPL) {
      const uint32_t r = j >> 9, jj = j & (P512_N - 1);
      sm.acc[lw][r][jj] =
          torus64_to_32(rot_div_coeff(lut + r * P512_N, P512_N, jj, b_hat));
    }
  }
  __syncthreads();

  uint32_t *acc_p = sm.acc[lw][p];
  cplx *xs_p = sm.xs[lw][p];
  const cplx *xs_l = &sm.xs[lw][0][0];
  const cplx *tw2_src = &tables->pass2[u][0];
  uint32_t own[32];
#pragma unroll
  for (int j1 = 0; j1 < 16; j1++) {
    own[j1] = acc_p[16 * j1 + u];
    own[16 + j1] = acc_p[16 * j1 + u + P512_M];
  }
  // (GGSW i, column p) block = [row r][b][q]; this thread reads element q = u
  const cplx *key_col = bsk + (size_t)p * (P * P512_M) + u;
  const size_t key_step = (size_t)P * P * P512_M;

  if constexpr (RING) {
    if (tid == 0 && n > 0)
      tma_issue(0); // the slot is free: every reader of the previous group passed a barrier
  }
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t a = sm.a_hat[lw][i];
    cplx v[16];
    n512_load_digits(acc_p, u, a, base_log, own, v);
    radix16_fwd(v, c_fft256_pass1);
    xq_store_p1(xs_p, u, v);
    __syncwarp();
    xq_load_p2(xs_p, u, v);
    {
      cplx tw2[15];
#pragma unroll
      for (int e = 0; e < 15; e++)
        tw2[e] = ldnc_cplx(tw2_src + e);
      radix16_fwd(v, tw2);
    }
    __syncwarp(); // every lane has read its exchange values
    spec256_store(xs_p, u, v);
    if constexpr (RING) {
      __syncthreads();
      mbar_wait_parity(&sm.bar, ring_phase & 1u);
      ring_phase++;
      // ---- Fourier MAC from shared memory ------------------------------------
      // the thread's own spectrum is still in registers: start with row r = p,
      // then walk the other rows cyclically (16 fewer 128-bit loads per thread)
      const cplx *kcol = &sm.ring[p * P][0];
      {
        const cplx *kp = kcol + (size_t)p * P512_M + u;
#pragma unroll
        for (int b = 0; b < 16; b++)
          v[b] = cmul(v[b], kp[b * 16]);
      }
#pragma unroll
      for (int rr = 1; rr < P; rr++) {
        int r = p + rr;
        if (r >= P)
          r -= P;
        const cplx *fr = xs_l + (size_t)r * P512_M + u;
        const cplx *kr = kcol + (size_t)r * P512_M + u;
#pragma unroll
        for (int b = 0; b < 16; b++)
          v[b] = cfma(fr[b * 16], kr[b * 16], v[b]);
      }
    } else {
    // the Fourier key of the step streams through a ring of three 8-value chunks
    // (two chunks = 16 x 128-bit loads in flight per thread); the first two are
    // requested before the share barrier
    const cplx *krow = key_col + (size_t)i * key_step;
    cplx kbuf[3][8];
    auto key_chunk = [&](int ch, cplx (&dst)[8]) {
      const int r = ch >> 1, hb = (ch & 1) * 8;
#pragma unroll
      for (int b = 0; b < 8; b++)
        dst[b] = ldcg_cplx(krow + (size_t)r * P512_M + (hb + b) * 16);
    };
    key_chunk(0, kbuf[0]);
    key_chunk(1, kbuf[1]);
    __syncthreads();
    // ---- Fourier MAC: out[b] = sum_r spec_r[b][q] * key[i][p][r][b][q] -------
#pragma unroll
    for (int ch = 0; ch < 2 * P; ch++) {
      const int r = ch >> 1, hb = (ch & 1) * 8;
      if (ch + 2 < 2 * P)
        key_chunk(ch + 2, kbuf[(ch + 2) % 3]);
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const cplx f = xs_l[(size_t)r * P512_M + (hb + b) * 16 + u];
        v[hb + b] = r == 0 ? cmul(f, kbuf[ch % 3][b]) : cfma(f, kbuf[ch % 3][b], v[hb + b]);
      }
    }
    }
    __syncthreads(); // every read of the shared spectra (and of the ring) done
    if constexpr (RING) {
      if (tid == 0 && i + 1 < n)
        tma_issue(i + 1);
    }
    {
      cplx tw2[15];
#pragma unroll
      for (int e = 0; e < 15; e++)
        tw2[e] = ldnc_cplx(tw2_src + e);
      radix16_inv(v, tw2);
    }
    xq_store_p2(xs_p, u, v);
    __syncwarp();
    xq_load_p1(xs_p, u, v);
    radix16_inv(v, c_fft256_pass1);
    n512_acc_update(acc_p, u, v, own);
    __syncwarp(); // the polynomial's 16 threads are its only readers
  }
  __syncthreads();

  // ---- epilogue: sample extract (a16), optional many-LUT --------------------
  if (live) {
    const uint64_t out_len = (uint64_t)K * P512_N + 1;
    for (uint32_t m = 0; m < num_many_lut; m++) {
      const uint32_t nth = m * lut_stride;
      uint64_t *out =
          lwe_out + ((uint64_t)m * num_samples + out_idx[s]) * out_len;
      for (uint32_t w = tl; w < (uint32_t)(K * P512_N); w += TPL) {
        const uint32_t r = w >> 9, tt = w & (P512_N - 1);
        const uint32_t x = tt <= nth ? sm.acc[lw][r][nth - tt]
                                     : 0u - sm.acc[lw][r][P512_N + nth - tt];
        out[w] = (uint64_t)x << 32;
      }
      if (tl == 0)
        out[(size_t)K * P512_N] = (uint64_t)sm.acc[lw][K][nth] << 32;
    }
  }
  __syncthreads(); // the working set is reused by the next sample group
  } // persistent loop
  (void)NT;
}

// ---------------------------------------------------------------------------
// BSK conversion for this kernel: standard-domain u64 polynomial, source order
// [i][row r][column c][N] (level count 1) -> spectrum scaled by 2^-64 / 256 *
// 2^32, stored at [i][c][r][b][q].  grid = n (k+1)^2 / 2 (two polynomials per
// warp), block = 32.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
bsk_convert_n512_kernel(cplx *__restrict__ dst, const uint64_t *__restrict__ src,
                        const Fft256Tables *__restrict__ tables, uint32_t P,
                        uint32_t total_polys) {
  __shared__ cplx xb[2][P512_M];
  const int hw = threadIdx.x >> 4, u = threadIdx.x & 15;
  uint32_t poly = blockIdx.x * 2 + hw;
  const bool live = poly < total_polys;
  if (!live)
    poly = total_polys - 1;
  const uint32_t c = poly % P, r = (poly / P) % P, i = poly / (P * P);
  const uint64_t *src_p = src + (size_t)poly * P512_N;
  // 2^-64 / 256 * 2^32 = 2^-40
  const double scale = 9.094947017729282379150390625e-13;
  cplx v[16];
#pragma unroll
  for (int j1 = 0; j1 < 16; j1++) {
    const uint32_t j = 16u * j1 + u;
    v[j1] = cmake(ll_to_double((int64_t)src_p[j]) * scale,
                  ll_to_double((int64_t)src_p[j + P512_M]) * scale);
  }
  radix16_fwd(v, c_fft256_pass1);
  xq_store_p1(xb[hw], u, v);
  __syncwarp();
  xq_load_p2(xb[hw], u, v);
  radix16_fwd(v, tables->pass2[u]);
  if (live) {
    cplx *out = dst + (((size_t)i * P + c) * P + r) * P512_M;
    spec256_store(out, u, v);
  }
}

} // namespace b200
