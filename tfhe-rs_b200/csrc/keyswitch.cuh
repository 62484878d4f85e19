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
// Not tensor-core work (u64 wrapping integers).  CUDA-core tiling: a CTA owns
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
  __shared__ uint64_t kt[KS_KC][KS_TO];     // [kk][output]
  __shared__ uint64_t in_base[KS_TS];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4; // 16 x 16 threads, 4x4 each
  const uint32_t s0 = blockIdx.x * KS_TS, o0 = blockIdx.y * KS_TO;
  const uint32_t out_len = n_out + 1;

  if (tid < KS_TS) {
    const uint32_t s = s0 + tid;
    in_base[tid] = s < count ? in_idx[s] * (uint64_t)(n_in + 1) : 0;
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
      kt[kk][oc] = (i < n_in && o < out_len)
                       ? ksk[((size_t)i * l + (kk % l)) * out_len + o]
                       : 0;
    }
    __syncthreads();
    for (uint32_t kk = 0; kk < kchunk; kk++) {
      uint64_t kv[4];
      int64_t dv[4];
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
          accu[a][b] += (uint64_t)dv[a] * kv[b];
    }
    __syncthreads();
  }

#pragma unroll
  for (int a = 0; a < 4; a++) {
    const uint32_t sl = ty + 16 * a, s = s0 + sl;
    if (s >= count)
      continue;
    uint64_t *o_row = lwe_out + out_idx[s] * (uint64_t)out_len;
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

} // namespace b200
