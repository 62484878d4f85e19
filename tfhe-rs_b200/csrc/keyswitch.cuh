// keyswitch.cuh -- batched LWE keyswitch as a fused-decomposition u64 "GEMM".
//
//   out[s][o] = [o == n_out] * b_in[s] - sum_{i < n_in} sum_{j < l}
//               digit_j(a_in[s][i]) * KSK[i][j][o]            (wrapping u64)
//
// Restates keyswitch_lwe_ciphertext (tfhe/src/core_crypto/algorithms/
// lwe_keyswitch.rs:137-232); replaces the reference GPU kernels
// `keyswitch` and `tgemm_all_levels*` (backends/tfhe-cuda-backend/cuda/src/
// crypto/keyswitch.cuh:199-340, linearalgebra/multiplication.cuh:102-342).
//
// Not tensor-core work (u64 wrapping integers).  The u64 x small-signed-digit
// MAC is two IMADs: the key word is staged as (lo_s, hi_adj) with lo_s the low
// half read as SIGNED and hi_adj = hi + (lo >> 31), so that
//   d * k = d * lo_s  (mad.wide.s32, 64-bit accumulate)  +  (d * hi_adj) << 32
// (mad.lo into the accumulator's high word), all mod 2^64.
// CUDA-core tiling: a CTA owns
// a 64-sample x 64-output tile, walks the input dimension in chunks of 32/l mask
// elements (x l levels), decomposes those mask elements once into shared
// memory, stages the matching KSK rows with coalesced 128-bit loads, and every
// thread keeps a 4 x 4 register tile of u64 accumulators.
#pragma once
#include "pbs_generic_phases.cuh"

#include <cuda_runtime.h>

namespace b200 {

constexpr int KS_TS = 64;  // samples per CTA
constexpr int KS_TO = 64;  // outputs per CTA
constexpr int KS_KC = 32;  // (mask element, level) rows per chunk; CI = 32 / l

// acc += d * k (mod 2^64) with k pre-split as described above
__device__ __forceinline__ void ks_mac(uint64_t &acc, int32_t d, uint2 k) {
  asm("{\n\t"
      ".reg .b32 lo, hi;\n\t"
      "mad.wide.s32 %0, %1, %2, %0;\n\t"
      "mov.b64 {lo, hi}, %0;\n\t"
      "mad.lo.s32 hi, %1, %3, hi;\n\t"
      "mov.b64 %0, {lo, hi};\n\t"
      "}"
      : "+l"(acc)
      : "r"(d), "r"((int32_t)k.x), "r"((int32_t)k.y));
}

// grid = (ceil(count/64), ceil((n_out+1)/64)), block = 256
__global__ void __launch_bounds__(256)
keyswitch_kernel(uint64_t *__restrict__ lwe_out,
                 const uint64_t *__restrict__ out_idx,
                 const uint64_t *__restrict__ lwe_in,
                 const uint64_t *__restrict__ in_idx,
                 const uint64_t *__restrict__ ksk, uint32_t n_in,
                 uint32_t n_out, uint32_t base_log, uint32_t l,
                 uint32_t count) {
  __shared__ int32_t dig[KS_KC][KS_TS + 1]; // [kk][sample]
  __shared__ uint2 kt[KS_KC][KS_TO];        // [kk][output] = (lo_s, hi_adj)
  __shared__ uint64_t in_base[KS_TS];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4; // 16 x 16 threads, 4x4 each
  const uint32_t s0 = blockIdx.x * KS_TS, o0 = blockIdx.y * KS_TO;
  const uint32_t out_len = n_out + 1;

  if (tid < KS_TS) {
    const uint32_t s = s0 + tid;
    in_base[tid] = s < count ? (in_idx ? in_idx[s] : (uint64_t)s) * (uint64_t)(n_in + 1) : 0;
  }
  __syncthreads();

  uint64_t accu[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++)
      accu[a][b] = 0;

  const uint32_t CI = KS_KC / l; // launcher guarantees 1 <= l <= 32
  const uint32_t kchunk = CI * l;
  for (uint32_t i0 = 0; i0 < n_in; i0 += CI) {
    // decompose CI mask elements of the 64 samples
    for (uint32_t w = tid; w < KS_TS * CI; w += 256) {
      const uint32_t sl = w & (KS_TS - 1), ci = w >> 6;
      const uint32_t s = s0 + sl, i = i0 + ci;
      uint64_t st = 0;
      const bool live = s < count && i < n_in;
      if (live)
        st = decomp_init_state(lwe_in[in_base[sl] + i], base_log, l);
      for (uint32_t j = 0; j < l; j++)
        dig[ci * l + j][sl] = live ? (int32_t)decomp_next_digit(&st, base_log) : 0;
    }
    // stage KSK rows [i0 .. i0+CI) x l levels x 64 outputs
    for (uint32_t w = tid; w < kchunk * KS_TO; w += 256) {
      const uint32_t kk = w >> 6, oc = w & (KS_TO - 1);
      const uint32_t i = i0 + kk / l, o = o0 + oc;
      const uint64_t kw = (i < n_in && o < out_len)
                              ? ksk[((size_t)i * l + (kk % l)) * out_len + o]
                              : 0;
      const uint32_t lo = (uint32_t)kw;
      kt[kk][oc] = make_uint2(lo, (uint32_t)(kw >> 32) + (lo >> 31));
    }
    __syncthreads();
    for (uint32_t kk = 0; kk < kchunk; kk++) {
      uint2 kv[4];
      int32_t dv[4];
#pragma unroll
      for (int b = 0; b < 4; b++)
        kv[b] = kt[kk][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; a++)
        dv[a] = dig[kk][ty + 16 * a];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
          ks_mac(accu[a][b], dv[a], kv[b]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int a = 0; a < 4; a++) {
    const uint32_t sl = ty + 16 * a, s = s0 + sl;
    if (s >= count)
      continue;
    uint64_t *o_row = lwe_out + (out_idx ? out_idx[s] : (uint64_t)s) * (uint64_t)out_len;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const uint32_t o = o0 + tx + 16 * b;
      if (o >= out_len)
        continue;
      uint64_t v = (uint64_t)0 - accu[a][b];
      if (o == n_out)
        v += lwe_in[in_base[sl] + n_in];
      o_row[o] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// 64 -> 32 keyswitch (KS32 parameter sets): u64 input ciphertexts, u32 key,
// u32 output.  Restates keyswitch_lwe_ciphertext_with_scalar_change
// (algorithms/lwe_keyswitch.rs:331-455): the mask elements are decomposed with
// the 64-bit decomposer, the digits multiply u32 key words (wrapping), and the
// body is the input body rounded to its top 32 bits (closest representable of
// a 1-level, 32-bit decomposition, then >> 32).  Replaces the reference's
// <uint64_t, uint32_t> instances (crypto/keyswitch.cuh:88-117,199-340).
// Same tiling as keyswitch_kernel; one IMAD per MAC.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
keyswitch_64_32_kernel(uint32_t *__restrict__ lwe_out,
                       const uint64_t *__restrict__ out_idx,
                       const uint64_t *__restrict__ lwe_in,
                       const uint64_t *__restrict__ in_idx,
                       const uint32_t *__restrict__ ksk, uint32_t n_in,
                       uint32_t n_out, uint32_t base_log, uint32_t l,
                       uint32_t count) {
  __shared__ int32_t dig[KS_KC][KS_TS + 1];
  __shared__ uint32_t kt[KS_KC][KS_TO];
  __shared__ uint64_t in_base[KS_TS];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const uint32_t s0 = blockIdx.x * KS_TS, o0 = blockIdx.y * KS_TO;
  const uint32_t out_len = n_out + 1;
  if (tid < KS_TS) {
    const uint32_t s = s0 + tid;
    in_base[tid] = s < count ? (in_idx ? in_idx[s] : (uint64_t)s) * (uint64_t)(n_in + 1) : 0;
  }
  __syncthreads();
  uint32_t accu[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++)
      accu[a][b] = 0;
  const uint32_t CI = KS_KC / l;
  const uint32_t kchunk = CI * l;
  for (uint32_t i0 = 0; i0 < n_in; i0 += CI) {
    for (uint32_t w = tid; w < KS_TS * CI; w += 256) {
      const uint32_t sl = w & (KS_TS - 1), ci = w >> 6;
      const uint32_t s = s0 + sl, i = i0 + ci;
      uint64_t st = 0;
      const bool live = s < count && i < n_in;
      if (live)
        st = decomp_init_state(lwe_in[in_base[sl] + i], base_log, l);
      for (uint32_t j = 0; j < l; j++)
        dig[ci * l + j][sl] = live ? (int32_t)decomp_next_digit(&st, base_log) : 0;
    }
    for (uint32_t w = tid; w < kchunk * KS_TO; w += 256) {
      const uint32_t kk = w >> 6, oc = w & (KS_TO - 1);
      const uint32_t i = i0 + kk / l, o = o0 + oc;
      kt[kk][oc] = (i < n_in && o < out_len)
                       ? ksk[((size_t)i * l + (kk % l)) * out_len + o]
                       : 0u;
    }
    __syncthreads();
    for (uint32_t kk = 0; kk < kchunk; kk++) {
      uint32_t kv[4];
      int32_t dv[4];
#pragma unroll
      for (int b = 0; b < 4; b++)
        kv[b] = kt[kk][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; a++)
        dv[a] = dig[kk][ty + 16 * a];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
          accu[a][b] += (uint32_t)dv[a] * kv[b];
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const uint32_t sl = ty + 16 * a, s = s0 + sl;
    if (s >= count)
      continue;
    uint32_t *o_row = lwe_out + (out_idx ? out_idx[s] : (uint64_t)s) * (uint64_t)out_len;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const uint32_t o = o0 + tx + 16 * b;
      if (o >= out_len)
        continue;
      uint32_t v = 0u - accu[a][b];
      if (o == n_out) // closest representable on 32 bits, then downscale
        v += (uint32_t)((lwe_in[in_base[sl] + n_in] + 0x80000000ull) >> 32);
      o_row[o] = v;
    }
  }
}


// ---------------------------------------------------------------------------
// fp64-pipe keyswitch.  Digits are tiny (|d| <= B/2) and each 32-bit half of a
// key word times a digit, summed over all n_in * l terms, stays far below 2^53,
// so both halves are accumulated EXACTLY with DFMA:
//   lo_acc += d * lo_s      hi_acc += d * hi_adj      (lo_s signed, see above)
//   out = (i64)lo_acc + ((i64)hi_acc << 32)   (mod 2^64)
// Two DFMA per u64 MAC and no integer carry chains: the integer version spends
// 4-5 instructions per MAC (IMAD.WIDE + IADD3 + IADD3.X + IMAD).  Valid while
// (base_log - 1) + 32 + ceil(log2(n_in * l)) <= 53; the launcher falls back to
// keyswitch_kernel otherwise.  Results are bit-identical to the integer path.
// grid = (ceil(count/64), ceil((n_out+1)/64)), block = 256, dynamic smem.
// ---------------------------------------------------------------------------
constexpr int KSF_KC = 32;
struct KsfSmem {
  double2 kt[KSF_KC][KS_TO];      // (lo_s, hi_adj) as doubles   32 KiB
  double dig[KSF_KC][KS_TS + 2];  // digits as doubles            16.5 KiB
  uint64_t in_base[KS_TS];
};

__global__ void __launch_bounds__(256)
keyswitch_f64_kernel(uint64_t *__restrict__ lwe_out,
                     const uint64_t *__restrict__ out_idx,
                     const uint64_t *__restrict__ lwe_in,
                     const uint64_t *__restrict__ in_idx,
                     const uint64_t *__restrict__ ksk, uint32_t n_in,
                     uint32_t n_out, uint32_t base_log, uint32_t l,
                     uint32_t count) {
  extern __shared__ __align__(16) unsigned char ks_smem_raw[];
  KsfSmem &sm = *reinterpret_cast<KsfSmem *>(ks_smem_raw);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const uint32_t s0 = blockIdx.x * KS_TS, o0 = blockIdx.y * KS_TO;
  const uint32_t out_len = n_out + 1;

  if (tid < KS_TS) {
    const uint32_t s = s0 + tid;
    sm.in_base[tid] = s < count ? (in_idx ? in_idx[s] : (uint64_t)s) * (uint64_t)(n_in + 1) : 0;
  }
  __syncthreads();

  double alo[4][4], ahi[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++)
      alo[a][b] = ahi[a][b] = 0.0;

  const uint32_t CI = KSF_KC / l;
  const uint32_t kchunk = CI * l;
  for (uint32_t i0 = 0; i0 < n_in; i0 += CI) {
    for (uint32_t w = tid; w < KS_TS * CI; w += 256) {
      const uint32_t sl = w & (KS_TS - 1), ci = w >> 6;
      const uint32_t s = s0 + sl, i = i0 + ci;
      uint64_t st = 0;
      const bool live = s < count && i < n_in;
      if (live)
        st = decomp_init_state(lwe_in[sm.in_base[sl] + i], base_log, l);
      for (uint32_t j = 0; j < l; j++)
        sm.dig[ci * l + j][sl] =
            live ? (double)(int32_t)decomp_next_digit(&st, base_log) : 0.0;
    }
    for (uint32_t w = tid; w < kchunk * KS_TO; w += 256) {
      const uint32_t kk = w >> 6, oc = w & (KS_TO - 1);
      const uint32_t i = i0 + kk / l, o = o0 + oc;
      const uint64_t kw = (i < n_in && o < out_len)
                              ? ksk[((size_t)i * l + (kk % l)) * out_len + o]
                              : 0;
      const uint32_t lo = (uint32_t)kw;
      const uint32_t hi = (uint32_t)(kw >> 32) + (lo >> 31);
      sm.kt[kk][oc] = make_double2((double)(int32_t)lo, (double)(int32_t)hi);
    }
    __syncthreads();
    for (uint32_t kk = 0; kk < kchunk; kk++) {
      double2 kv[4];
      double dv[4];
#pragma unroll
      for (int b = 0; b < 4; b++)
        kv[b] = sm.kt[kk][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; a++)
        dv[a] = sm.dig[kk][ty + 16 * a];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          alo[a][b] = fma(dv[a], kv[b].x, alo[a][b]);
          ahi[a][b] = fma(dv[a], kv[b].y, ahi[a][b]);
        }
    }
    __syncthreads();
  }

#pragma unroll
  for (int a = 0; a < 4; a++) {
    const uint32_t sl = ty + 16 * a, s = s0 + sl;
    if (s >= count)
      continue;
    uint64_t *o_row = lwe_out + (out_idx ? out_idx[s] : (uint64_t)s) * (uint64_t)out_len;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const uint32_t o = o0 + tx + 16 * b;
      if (o >= out_len)
        continue;
      const uint64_t sum = (uint64_t)__double2ll_rn(alo[a][b]) +
                           ((uint64_t)__double2ll_rn(ahi[a][b]) << 32);
      uint64_t v = (uint64_t)0 - sum;
      if (o == n_out)
        v += lwe_in[sm.in_base[sl] + n_in];
      o_row[o] = v;
    }
  }
}

} // namespace b200
