// seeded_key.cuh -- on-device decompression of a SEEDED bootstrap key.
//
// tfhe-rs clients ship bootstrap keys "seeded": only the body polynomial of
// every GGSW row plus the CSPRNG seed the masks were drawn from
// (SeededLweBootstrapKey, tfhe/src/core_crypto/entities/seeded_lwe_bootstrap_key.rs;
// decompression tfhe/src/core_crypto/algorithms/
// seeded_lwe_bootstrap_key_decompression.rs:36-60 ->
// seeded_ggsw_ciphertext_list_decompression.rs).  The reference expands them on
// the host and uploads the full key.  Here only the bodies cross PCIe (half the
// bytes): the masks are regenerated on the GPU and fed straight into the
// Fourier conversion.
//
// The mask stream is the byte table of tfhe-csprng's AES-128 counter-mode
// generator (tfhe-csprng/src/generators/aes_ctr/generic.rs:178-193):
//   byte p = AES_K(le128(counter0 + (p >> 4)))[p & 15]
// consumed as little-endian u64 words (commons/math/random/uniform.rs:11-20)
// in container order [ggsw][level][row][mask poly][coefficient]; forked child
// generators own consecutive byte ranges (generic.rs:142-176), so the whole key
// is ONE contiguous run of the table.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

struct AesCtrKey {
  uint32_t rk[44]; // 11 round keys, 4 little-endian column words each
};

struct AesTables {
  uint32_t te0[256]; // bytes (2S, S, S, 3S) of column rows 0..3, little endian
  uint8_t sbox[256];
};

// ---- host side: FIPS-197 tables and key schedule ---------------------------
inline uint8_t aes_xtime(uint8_t a) {
  return (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1b : 0));
}
inline uint8_t aes_gf_mul(uint8_t a, uint8_t b) {
  uint8_t p = 0;
  for (int i = 0; i < 8; i++) {
    if (b & 1)
      p ^= a;
    a = aes_xtime(a);
    b >>= 1;
  }
  return p;
}
inline void aes_fill_tables(AesTables &t) {
  for (int x = 0; x < 256; x++) {
    uint8_t inv = 0;
    if (x) { // x^254 = x^-1 in GF(2^8)
      uint8_t acc = 1, base = (uint8_t)x;
      for (int e = 254; e; e >>= 1) {
        if (e & 1)
          acc = aes_gf_mul(acc, base);
        base = aes_gf_mul(base, base);
      }
      inv = acc;
    }
    auto rotl = [](uint8_t v, int s) { return (uint8_t)((v << s) | (v >> (8 - s))); };
    const uint8_t s =
        (uint8_t)(inv ^ rotl(inv, 1) ^ rotl(inv, 2) ^ rotl(inv, 3) ^ rotl(inv, 4) ^ 0x63);
    t.sbox[x] = s;
    const uint8_t s2 = aes_xtime(s), s3 = (uint8_t)(s2 ^ s);
    t.te0[x] = (uint32_t)s2 | ((uint32_t)s << 8) | ((uint32_t)s << 16) |
               ((uint32_t)s3 << 24);
  }
}
inline void aes_expand_key(const uint8_t key[16], const AesTables &t,
                           AesCtrKey &out) {
  uint8_t rk[11][16];
  for (int i = 0; i < 16; i++)
    rk[0][i] = key[i];
  uint8_t rcon = 1;
  for (int r = 1; r <= 10; r++) {
    const uint8_t *prev = rk[r - 1];
    uint8_t tmp[4] = {(uint8_t)(t.sbox[prev[13]] ^ rcon), t.sbox[prev[14]],
                      t.sbox[prev[15]], t.sbox[prev[12]]};
    rcon = aes_xtime(rcon);
    for (int c = 0; c < 4; c++)
      for (int b = 0; b < 4; b++)
        rk[r][4 * c + b] =
            (uint8_t)(prev[4 * c + b] ^ (c == 0 ? tmp[b] : rk[r][4 * (c - 1) + b]));
  }
  for (int r = 0; r < 11; r++)
    for (int c = 0; c < 4; c++)
      out.rk[4 * r + c] = (uint32_t)rk[r][4 * c] | ((uint32_t)rk[r][4 * c + 1] << 8) |
                          ((uint32_t)rk[r][4 * c + 2] << 16) |
                          ((uint32_t)rk[r][4 * c + 3] << 24);
}

// ---- device side -------------------------------------------------------------
__device__ __forceinline__ uint32_t aes_rotl(uint32_t v, int s) {
  return __funnelshift_l(v, v, s);
}

// one AES-128 block: in/out as 4 little-endian column words
__device__ __forceinline__ void aes128_encrypt(const uint32_t *__restrict__ te0,
                                               const uint8_t *__restrict__ sbox,
                                               const AesCtrKey &key, uint32_t s[4]) {
#pragma unroll
  for (int c = 0; c < 4; c++)
    s[c] ^= key.rk[c];
#pragma unroll 1
  for (int r = 1; r < 10; r++) {
    uint32_t n[4];
#pragma unroll
    for (int c = 0; c < 4; c++)
      n[c] = te0[s[c] & 0xff] ^ aes_rotl(te0[(s[(c + 1) & 3] >> 8) & 0xff], 8) ^
             aes_rotl(te0[(s[(c + 2) & 3] >> 16) & 0xff], 16) ^
             aes_rotl(te0[s[(c + 3) & 3] >> 24], 24) ^ key.rk[4 * r + c];
#pragma unroll
    for (int c = 0; c < 4; c++)
      s[c] = n[c];
  }
  uint32_t n[4];
#pragma unroll
  for (int c = 0; c < 4; c++)
    n[c] = ((uint32_t)sbox[s[c] & 0xff] |
            ((uint32_t)sbox[(s[(c + 1) & 3] >> 8) & 0xff] << 8) |
            ((uint32_t)sbox[(s[(c + 2) & 3] >> 16) & 0xff] << 16) |
            ((uint32_t)sbox[s[(c + 3) & 3] >> 24] << 24)) ^
           key.rk[40 + c];
#pragma unroll
  for (int c = 0; c < 4; c++)
    s[c] = n[c];
}

// Expands the seeded key into the standard-domain container
// [row = (ggsw, level, glwe row)][k + 1 polys][N]: mask polys from the AES-CTR
// table, body poly copied.  One thread per AES block (two mask words).
//   word w of the mask stream sits at table byte 8 * (w + word_shift),
//   word_shift in {0, 1} (a start index 8 bytes into a block).
__global__ void __launch_bounds__(256)
seeded_bsk_expand_kernel(uint64_t *__restrict__ standard,
                         const uint64_t *__restrict__ bodies,
                         const AesTables *__restrict__ tables, AesCtrKey key,
                         uint64_t ctr_lo, uint64_t ctr_hi, uint32_t word_shift,
                         uint64_t rows, uint32_t k, uint32_t N) {
  __shared__ uint32_t te0[256];
  __shared__ uint8_t sbox[256];
  te0[threadIdx.x] = tables->te0[threadIdx.x];
  sbox[threadIdx.x] = tables->sbox[threadIdx.x];
  __syncthreads();

  const uint64_t mask_words = rows * (uint64_t)k * N;
  const uint64_t blocks = (mask_words + word_shift + 1) / 2;
  const uint64_t words_per_row = (uint64_t)k * N;
  for (uint64_t blk = (uint64_t)blockIdx.x * 256 + threadIdx.x; blk < blocks;
       blk += (uint64_t)gridDim.x * 256) {
    // counter = ctr + blk (128-bit)
    const uint64_t lo = ctr_lo + blk;
    const uint64_t hi = ctr_hi + (lo < ctr_lo ? 1 : 0);
    uint32_t s[4] = {(uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi,
                     (uint32_t)(hi >> 32)};
    aes128_encrypt(te0, sbox, key, s);
    const uint64_t w0 = ((uint64_t)s[1] << 32) | s[0];
    const uint64_t w1 = ((uint64_t)s[3] << 32) | s[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint64_t q = 2 * blk + h;
      if (q < word_shift)
        continue;
      const uint64_t w = q - word_shift;
      if (w >= mask_words)
        continue;
      const uint64_t row = w / words_per_row, in_row = w % words_per_row;
      standard[row * (uint64_t)(k + 1) * N + in_row] = h ? w1 : w0;
    }
  }
  // bodies
  const uint64_t body_words = rows * N;
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < body_words;
       i += (uint64_t)gridDim.x * 256) {
    const uint64_t row = i / N, j = i % N;
    standard[row * (uint64_t)(k + 1) * N + (uint64_t)k * N + j] = bodies[i];
  }
}

} // namespace b200
