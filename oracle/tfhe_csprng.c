/* tfhe_csprng.c -- see tfhe_csprng.h.  TEST INFRASTRUCTURE ONLY.
 *
 * AES-128 is written from FIPS-197 (S-box computed from its definition, no
 * tables typed in); an AES-NI twin is used when the CPU has it.  Both are
 * checked against the FIPS-197 appendix vectors the reference's own block
 * cipher tests use (tfhe-csprng/src/generators/aes_ctr/block_cipher.rs:62-110).
 */
#include "tfhe_csprng.h"

#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#define CSPRNG_X86 1
#else
#define CSPRNG_X86 0
#endif

/* ---------------------------------------------------------------------- */
/* portable AES-128                                                        */
/* ---------------------------------------------------------------------- */
static uint8_t g_sbox[256];
static int g_sbox_ready = 0;

static uint8_t gf_mul(uint8_t a, uint8_t b) {
  uint8_t p = 0;
  for (int i = 0; i < 8; i++) {
    if (b & 1)
      p ^= a;
    const uint8_t hi = a & 0x80;
    a = (uint8_t)(a << 1);
    if (hi)
      a ^= 0x1b; /* x^8 + x^4 + x^3 + x + 1 */
    b >>= 1;
  }
  return p;
}

static uint8_t rotl8(uint8_t x, int s) {
  return (uint8_t)((x << s) | (x >> (8 - s)));
}

static void sbox_init(void) {
  if (g_sbox_ready)
    return;
  for (int x = 0; x < 256; x++) {
    /* multiplicative inverse in GF(2^8) (0 -> 0): x^254 */
    uint8_t inv = 0;
    if (x) {
      uint8_t acc = 1, base = (uint8_t)x;
      int e = 254;
      while (e) {
        if (e & 1)
          acc = gf_mul(acc, base);
        base = gf_mul(base, base);
        e >>= 1;
      }
      inv = acc;
    }
    /* affine transformation (FIPS-197 5.1.1) */
    g_sbox[x] = (uint8_t)(inv ^ rotl8(inv, 1) ^ rotl8(inv, 2) ^ rotl8(inv, 3) ^
                          rotl8(inv, 4) ^ 0x63);
  }
  g_sbox_ready = 1;
}

static void key_expand(const uint8_t key[16], uint8_t rk[11][16]) {
  sbox_init();
  memcpy(rk[0], key, 16);
  uint8_t rcon = 1;
  for (int r = 1; r <= 10; r++) {
    const uint8_t *prev = rk[r - 1];
    uint8_t t[4] = {g_sbox[prev[13]], g_sbox[prev[14]], g_sbox[prev[15]],
                    g_sbox[prev[12]]};
    t[0] ^= rcon;
    rcon = gf_mul(rcon, 2);
    for (int c = 0; c < 4; c++) {
      for (int b = 0; b < 4; b++) {
        const uint8_t left = (c == 0) ? t[b] : rk[r][(c - 1) * 4 + b];
        rk[r][c * 4 + b] = (uint8_t)(prev[c * 4 + b] ^ left);
      }
    }
  }
}

static void encrypt_block_portable(const uint8_t rk[11][16],
                                   const uint8_t in[16], uint8_t out[16]) {
  uint8_t s[16], t[16];
  for (int i = 0; i < 16; i++)
    s[i] = in[i] ^ rk[0][i];
  for (int r = 1; r <= 10; r++) {
    /* SubBytes + ShiftRows: state is column-major, byte (row, col) = s[4c+r] */
    for (int c = 0; c < 4; c++)
      for (int row = 0; row < 4; row++)
        t[4 * c + row] = g_sbox[s[4 * ((c + row) & 3) + row]];
    if (r < 10) {
      for (int c = 0; c < 4; c++) {
        const uint8_t a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2],
                      a3 = t[4 * c + 3];
        s[4 * c + 0] = (uint8_t)(gf_mul(a0, 2) ^ gf_mul(a1, 3) ^ a2 ^ a3);
        s[4 * c + 1] = (uint8_t)(a0 ^ gf_mul(a1, 2) ^ gf_mul(a2, 3) ^ a3);
        s[4 * c + 2] = (uint8_t)(a0 ^ a1 ^ gf_mul(a2, 2) ^ gf_mul(a3, 3));
        s[4 * c + 3] = (uint8_t)(gf_mul(a0, 3) ^ a1 ^ a2 ^ gf_mul(a3, 2));
      }
    } else {
      memcpy(s, t, 16);
    }
    for (int i = 0; i < 16; i++)
      s[i] ^= rk[r][i];
  }
  memcpy(out, s, 16);
}

/* ---------------------------------------------------------------------- */
/* AES-NI twin                                                             */
/* ---------------------------------------------------------------------- */
#if CSPRNG_X86
__attribute__((target("aes,sse2"))) static void
encrypt_blocks_aesni(const uint8_t rk[11][16], uint64_t first_index,
                     uint8_t *out, size_t blocks) {
  __m128i k[11];
  for (int r = 0; r < 11; r++)
    k[r] = _mm_loadu_si128((const __m128i *)rk[r]);
  for (size_t b = 0; b < blocks; b++) {
    /* counter = aes index as a little-endian u128 (high half 0 here) */
    __m128i x = _mm_set_epi64x(0, (long long)(first_index + b));
    x = _mm_xor_si128(x, k[0]);
    for (int r = 1; r < 10; r++)
      x = _mm_aesenc_si128(x, k[r]);
    x = _mm_aesenclast_si128(x, k[10]);
    _mm_storeu_si128((__m128i *)(out + 16 * b), x);
  }
}
#endif

int csprng_uses_aesni(void) {
#if CSPRNG_X86
  static int cached = -1;
  if (cached < 0) {
    __builtin_cpu_init();
    cached = __builtin_cpu_supports("aes") ? 1 : 0;
    if (getenv("ORACLE_CSPRNG_PORTABLE"))
      cached = 0;
  }
  return cached;
#else
  return 0;
#endif
}

static void encrypt_counter_blocks(const uint8_t rk[11][16],
                                   uint64_t first_index, uint8_t *out,
                                   size_t blocks) {
#if CSPRNG_X86
  if (csprng_uses_aesni()) {
    encrypt_blocks_aesni(rk, first_index, out, blocks);
    return;
  }
#endif
  for (size_t b = 0; b < blocks; b++) {
    uint8_t ctr[16] = {0};
    const uint64_t idx = first_index + b;
    for (int i = 0; i < 8; i++)
      ctr[i] = (uint8_t)(idx >> (8 * i));
    encrypt_block_portable(rk, ctr, out + 16 * b);
  }
}

void csprng_aes128_encrypt_block(const uint8_t key[16], const uint8_t in[16],
                                 uint8_t out[16]) {
  uint8_t rk[11][16];
  key_expand(key, rk);
  encrypt_block_portable(rk, in, out);
}

/* ---------------------------------------------------------------------- */
/* byte table generator                                                    */
/* ---------------------------------------------------------------------- */
void csprng_init(csprng_gen *g, uint64_t seed_lo, uint64_t seed_hi) {
  uint8_t key[16];
  for (int i = 0; i < 8; i++) {
    key[i] = (uint8_t)(seed_lo >> (8 * i));
    key[8 + i] = (uint8_t)(seed_hi >> (8 * i));
  }
  key_expand(key, g->round_keys);
  g->pos = 0;
}

void csprng_at(const csprng_gen *parent, uint64_t pos, csprng_gen *child) {
  memcpy(child->round_keys, parent->round_keys, sizeof(parent->round_keys));
  child->pos = pos;
}

void csprng_fill_bytes(csprng_gen *g, uint8_t *out, size_t count) {
  enum { CHUNK_BLOCKS = 256 };
  uint8_t buf[16 * CHUNK_BLOCKS];
  while (count) {
    const uint64_t first = g->pos >> 4;
    const size_t skip = (size_t)(g->pos & 15);
    size_t blocks = (skip + count + 15) / 16;
    if (blocks > CHUNK_BLOCKS)
      blocks = CHUNK_BLOCKS;
    encrypt_counter_blocks((const uint8_t(*)[16])g->round_keys, first, buf,
                           blocks);
    size_t take = blocks * 16 - skip;
    if (take > count)
      take = count;
    memcpy(out, buf + skip, take);
    out += take;
    count -= take;
    g->pos += take;
  }
}

uint64_t csprng_uniform_u64(csprng_gen *g) {
  uint8_t b[8];
  csprng_fill_bytes(g, b, 8);
  uint64_t v = 0;
  for (int i = 0; i < 8; i++)
    v |= (uint64_t)b[i] << (8 * i);
  return v;
}

void csprng_fill_uniform_u64(csprng_gen *g, uint64_t *out, size_t count) {
  /* little-endian host assumed (x86-64 / aarch64): bytes land in place */
  csprng_fill_bytes(g, (uint8_t *)out, count * 8);
}

void csprng_fill_binary_u64(csprng_gen *g, uint64_t *out, size_t count) {
  uint8_t *bytes = (uint8_t *)malloc(count ? count : 1);
  csprng_fill_bytes(g, bytes, count);
  for (size_t i = 0; i < count; i++)
    out[i] = bytes[i] & 1u;
  free(bytes);
}

uint32_t csprng_tuniform_bytes(uint32_t bound_log2) {
  return (bound_log2 + 2 + 7) / 8;
}

static int64_t tuniform_from_bytes(const uint8_t *b, uint32_t bound_log2) {
  const uint32_t bits = bound_log2 + 2;
  const uint32_t nbytes = (bits + 7) / 8;
  uint64_t v = 0;
  for (uint32_t i = 0; i < nbytes; i++)
    v |= (uint64_t)b[i] << (8 * i);
  v &= (bits >= 64) ? ~(uint64_t)0 : (((uint64_t)1 << bits) - 1);
  const uint64_t low = v & 1;
  v = (v >> 1) + low;
  return (int64_t)(v - ((uint64_t)1 << bound_log2));
}

int64_t csprng_tuniform(csprng_gen *g, uint32_t bound_log2) {
  uint8_t b[8];
  csprng_fill_bytes(g, b, csprng_tuniform_bytes(bound_log2));
  return tuniform_from_bytes(b, bound_log2);
}

/* ---------------------------------------------------------------------- */
/* tfhe-rs draw order                                                      */
/* ---------------------------------------------------------------------- */
void csprng_resources_init(csprng_resources *r, uint64_t seed_lo,
                           uint64_t seed_hi) {
  csprng_gen seeder;
  csprng_init(&seeder, seed_lo, seed_hi);
  uint64_t s[6];
  csprng_fill_uniform_u64(&seeder, s, 6); /* three u128, little-endian */
  csprng_init(&r->mask, s[0], s[1]);
  csprng_init(&r->noise, s[2], s[3]);
  csprng_init(&r->secret, s[4], s[5]);
}

void csprng_gen_binary_key(csprng_resources *r, uint64_t *key, size_t count) {
  csprng_fill_binary_u64(&r->secret, key, count);
}

/* acc += a * s, s binary, negacyclic */
static void mul_add_binary(uint64_t *restrict acc, const uint64_t *restrict a,
                           const uint64_t *restrict s, uint32_t N) {
  for (uint32_t i = 0; i < N; i++) {
    if (!s[i])
      continue;
    for (uint32_t j = 0; j < N - i; j++)
      acc[i + j] += a[j];
    for (uint32_t j = N - i; j < N; j++)
      acc[i + j - N] -= a[j];
  }
}

/* One constant GGSW whose generators start at the given absolute positions
 * (ggsw_encryption.rs:251-277 level order, :361-413 rows,
 * glwe_encryption.rs:99-120 mask then noise). */
static void ggsw_at(const csprng_resources *r, uint64_t mask_pos,
                    uint64_t noise_pos, const uint64_t *glwe_key, uint32_t k,
                    uint32_t N, uint32_t base_log, uint32_t level_count,
                    uint32_t noise_bound_log2, uint64_t m, uint64_t *ggsw_out) {
  csprng_gen mask, noise;
  csprng_at(&r->mask, mask_pos, &mask);
  csprng_at(&r->noise, noise_pos, &noise);
  const size_t row_len = (size_t)(k + 1) * N;
  const uint32_t nb = csprng_tuniform_bytes(noise_bound_log2);
  uint8_t *nbytes = (uint8_t *)malloc((size_t)N * nb);
  for (uint32_t t = 0; t < level_count; t++) {
    const uint32_t level = level_count - t;
    const uint64_t factor = ((uint64_t)0 - m) << (64 - base_log * level);
    for (uint32_t row = 0; row <= k; row++) {
      uint64_t *glwe = ggsw_out + ((size_t)t * (k + 1) + row) * row_len;
      uint64_t *body = glwe + (size_t)k * N;
      if (row < k) {
        for (uint32_t j = 0; j < N; j++)
          body[j] = glwe_key[(size_t)row * N + j] * factor;
      } else {
        memset(body, 0, (size_t)N * sizeof(uint64_t));
        body[0] = (uint64_t)0 - factor;
      }
      csprng_fill_uniform_u64(&mask, glwe, (size_t)k * N);
      csprng_fill_bytes(&noise, nbytes, (size_t)N * nb);
      for (uint32_t j = 0; j < N; j++)
        body[j] += (uint64_t)tuniform_from_bytes(nbytes + (size_t)j * nb,
                                                 noise_bound_log2);
      for (uint32_t p = 0; p < k; p++)
        mul_add_binary(body, glwe + (size_t)p * N, glwe_key + (size_t)p * N, N);
    }
  }
  free(nbytes);
}

static uint64_t ggsw_mask_bytes(uint32_t k, uint32_t N, uint32_t level_count) {
  return (uint64_t)level_count * (k + 1) * k * N * 8;
}
static uint64_t ggsw_noise_bytes(uint32_t k, uint32_t N, uint32_t level_count,
                                 uint32_t noise_bound_log2) {
  return (uint64_t)level_count * (k + 1) * N *
         csprng_tuniform_bytes(noise_bound_log2);
}

void csprng_gen_bsk(csprng_resources *r, const uint64_t *lwe_key, uint32_t n,
                    const uint64_t *glwe_key, uint32_t k, uint32_t N,
                    uint32_t base_log, uint32_t level_count,
                    uint32_t noise_bound_log2, uint64_t *bsk_out) {
  const size_t ggsw_len = (size_t)level_count * (k + 1) * (k + 1) * N;
  const uint64_t mb = ggsw_mask_bytes(k, N, level_count);
  const uint64_t nb = ggsw_noise_bytes(k, N, level_count, noise_bound_log2);
  const uint64_t m0 = r->mask.pos, n0 = r->noise.pos;
#pragma omp parallel for schedule(dynamic, 4)
  for (uint32_t i = 0; i < n; i++)
    ggsw_at(r, m0 + (uint64_t)i * mb, n0 + (uint64_t)i * nb, glwe_key, k, N,
            base_log, level_count, noise_bound_log2, lwe_key[i],
            bsk_out + (size_t)i * ggsw_len);
  r->mask.pos = m0 + (uint64_t)n * mb;
  r->noise.pos = n0 + (uint64_t)n * nb;
}

void csprng_gen_multi_bit_bsk(csprng_resources *r, const uint64_t *lwe_key,
                              uint32_t n, const uint64_t *glwe_key, uint32_t k,
                              uint32_t N, uint32_t base_log,
                              uint32_t level_count, uint32_t grouping_factor,
                              uint32_t noise_bound_log2, uint64_t *bsk_out) {
  const uint32_t g = grouping_factor;
  const uint32_t per_group = 1u << g;
  const uint32_t total = (n / g) * per_group;
  const size_t ggsw_len = (size_t)level_count * (k + 1) * (k + 1) * N;
  const uint64_t mb = ggsw_mask_bytes(k, N, level_count);
  const uint64_t nb = ggsw_noise_bytes(k, N, level_count, noise_bound_log2);
  const uint64_t m0 = r->mask.pos, n0 = r->noise.pos;
#pragma omp parallel for schedule(dynamic, 4)
  for (uint32_t idx = 0; idx < total; idx++) {
    const uint32_t grp = idx / per_group, sel = idx % per_group;
    /* combine_key_bits (lwe_multi_bit_bootstrap_key_generation.rs:504-529) */
    uint64_t m = 1;
    for (uint32_t u = 0; u < g; u++) {
      const uint32_t pos = g - (u + 1);
      const uint64_t inv = ((sel >> pos) & 1u) ^ 1u;
      m *= lwe_key[grp * g + u] ^ inv;
    }
    ggsw_at(r, m0 + (uint64_t)idx * mb, n0 + (uint64_t)idx * nb, glwe_key, k,
            N, base_log, level_count, noise_bound_log2, m,
            bsk_out + (size_t)idx * ggsw_len);
  }
  r->mask.pos = m0 + (uint64_t)total * mb;
  r->noise.pos = n0 + (uint64_t)total * nb;
}

void csprng_lwe_encrypt(csprng_resources *r, const uint64_t *key, uint32_t n,
                        uint64_t plaintext, uint32_t noise_bound_log2,
                        uint64_t *ct_out) {
  csprng_fill_uniform_u64(&r->mask, ct_out, n);
  uint64_t b = (uint64_t)csprng_tuniform(&r->noise, noise_bound_log2);
  for (uint32_t i = 0; i < n; i++)
    b += ct_out[i] * key[i];
  ct_out[n] = b + plaintext;
}
