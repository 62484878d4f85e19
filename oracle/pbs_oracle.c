/*
 * pbs_oracle.c -- CPU restatement of the tfhe-rs core_crypto PBS path.
 * TEST INFRASTRUCTURE ONLY (see pbs_oracle.h).  Pinned by the KATs and the
 * regenerated reference goldens listed in the header; PBS output words are
 * the one thing that cannot be pinned across implementations.
 *
 * Written from the algorithm description and the
 * reference's behaviour; no reference source is copied.  Citations are
 * relative to /root/reference.
 */
#include "pbs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define W 64

static void *xalloc(size_t bytes) {
  bytes = (bytes + 63) & ~(size_t)63;
  if (bytes == 0)
    bytes = 64;
  void *p = aligned_alloc(64, bytes);
  if (!p)
    abort();
  return p;
}

/* ====================================================================== */
/* PRNG                                                                    */
/* ====================================================================== */
static inline uint64_t rotl64(uint64_t x, int k) {
  return (x << k) | (x >> (64 - k));
}

void orc_rng_seed(orc_rng *rng, uint64_t seed) {
  /* splitmix64 expansion of the seed */
  for (int i = 0; i < 4; i++) {
    seed += 0x9E3779B97F4A7C15ull;
    uint64_t z = seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    rng->s[i] = z ^ (z >> 31);
  }
}

uint64_t orc_rng_next(orc_rng *rng) {
  uint64_t *s = rng->s;
  const uint64_t result = rotl64(s[1] * 5, 7) * 9;
  const uint64_t t = s[1] << 17;
  s[2] ^= s[0];
  s[3] ^= s[1];
  s[1] ^= s[2];
  s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl64(s[3], 45);
  return result;
}

void orc_fill_uniform(orc_rng *rng, uint64_t *out, size_t count) {
  for (size_t i = 0; i < count; i++)
    out[i] = orc_rng_next(rng);
}

void orc_fill_binary(orc_rng *rng, uint64_t *out, size_t count) {
  uint64_t bits = 0;
  for (size_t i = 0; i < count; i++) {
    if ((i & 63) == 0)
      bits = orc_rng_next(rng);
    out[i] = bits & 1;
    bits >>= 1;
  }
}

/* t_uniform.rs:85-112: b+2 random bits r; v = (r>>1) + (r&1) - 2^b */
int64_t orc_tuniform(orc_rng *rng, uint32_t bound_log2) {
  uint64_t r = orc_rng_next(rng) & ((1ull << (bound_log2 + 2)) - 1);
  return (int64_t)(r >> 1) + (int64_t)(r & 1) - ((int64_t)1 << bound_log2);
}

/* ====================================================================== */
/* small integer routines                                                  */
/* ====================================================================== */
uint64_t orc_modulus_switch(uint64_t x, uint32_t log_modulus) {
  if (log_modulus == W)
    return x;
  return (x + (1ull << (W - log_modulus - 1))) >> (W - log_modulus);
}

uint64_t orc_decomposer_init_state(uint64_t x, uint32_t base_log,
                                   uint32_t level_count) {
  const uint32_t R = base_log * level_count;
  uint64_t r = x >> (W - R - 1);
  const uint64_t rb = r & 1;
  r = (r + 1) >> 1;
  r &= (R == 64) ? ~0ull : ((1ull << R) - 1);
  const uint64_t bal = (((r - 1) | (rb << (R - 1))) & r) >> (R - 1);
  return r - (bal << R);
}

static inline int64_t decompose_one_level(uint64_t *state, uint32_t base_log) {
  const uint64_t mask = (1ull << base_log) - 1;
  const uint64_t res = *state & mask;
  uint64_t st = (uint64_t)((int64_t)*state >> base_log); /* arithmetic */
  const uint64_t carry = (((res - 1) | st) & res) >> (base_log - 1);
  st += carry;
  *state = st;
  return (int64_t)(res - (carry << base_log));
}

void orc_decompose(uint64_t x, uint32_t base_log, uint32_t level_count,
                   int64_t *digits_out) {
  uint64_t st = orc_decomposer_init_state(x, base_log, level_count);
  for (uint32_t t = 0; t < level_count; t++)
    digits_out[t] = decompose_one_level(&st, base_log);
}

uint64_t orc_closest_representable(uint64_t x, uint32_t base_log,
                                   uint32_t level_count) {
  const uint32_t R = base_log * level_count;
  return orc_decomposer_init_state(x, base_log, level_count) << (W - R);
}

void orc_monomial_div(uint64_t *out, const uint64_t *in, uint32_t N,
                      uint32_t d) {
  const uint32_t rem = d % N;
  const int odd = (d / N) & 1;
  /* out[j] = in[j + rem] for j < N - rem ; out[j] = -in[j + rem - N] else;
   * all negated on an odd number of full cycles. */
  for (uint32_t j = 0; j < N - rem; j++)
    out[j] = odd ? (uint64_t)0 - in[j + rem] : in[j + rem];
  for (uint32_t j = N - rem; j < N; j++)
    out[j] = odd ? in[j + rem - N] : (uint64_t)0 - in[j + rem - N];
}

void orc_monomial_mul_and_subtract(uint64_t *out, const uint64_t *in,
                                   uint32_t N, uint32_t d) {
  const uint32_t rem = d % N;
  const int odd = (d / N) & 1;
  /* (in * X^d)[j] = -in[j - rem + N] for j < rem, in[j - rem] otherwise */
  for (uint32_t j = 0; j < rem; j++) {
    uint64_t v = (uint64_t)0 - in[j + N - rem];
    if (odd)
      v = (uint64_t)0 - v;
    out[j] = v - in[j];
  }
  for (uint32_t j = rem; j < N; j++) {
    uint64_t v = in[j - rem];
    if (odd)
      v = (uint64_t)0 - v;
    out[j] = v - in[j];
  }
}

void orc_negacyclic_mul_add_exact(uint64_t *restrict out,
                                  const int64_t *restrict a,
                                  const uint64_t *restrict b, uint32_t N) {
  for (uint32_t i = 0; i < N; i++) {
    const uint64_t d = (uint64_t)a[i];
    if (d == 0)
      continue;
    uint64_t *o1 = out + i;
    for (uint32_t j = 0; j < N - i; j++)
      o1[j] += d * b[j];
    uint64_t *o2 = out;
    const uint64_t *b2 = b + (N - i);
    for (uint32_t j = 0; j < i; j++)
      o2[j] -= d * b2[j];
  }
}

/* ====================================================================== */
/* encryption / key generation                                             */
/* ====================================================================== */
void orc_lwe_encrypt(orc_rng *rng, const uint64_t *key, uint32_t n,
                     uint64_t plaintext, int32_t noise_log2,
                     uint64_t *ct_out) {
  uint64_t b = plaintext;
  for (uint32_t i = 0; i < n; i++) {
    const uint64_t a = orc_rng_next(rng);
    ct_out[i] = a;
    b += a * key[i];
  }
  if (noise_log2 >= 0)
    b += (uint64_t)orc_tuniform(rng, (uint32_t)noise_log2);
  ct_out[n] = b;
}

uint64_t orc_lwe_decrypt(const uint64_t *key, uint32_t n,
                         const uint64_t *ct) {
  uint64_t acc = ct[n];
  for (uint32_t i = 0; i < n; i++)
    acc -= ct[i] * key[i];
  return acc;
}

/* acc += a * s for a binary key polynomial s (negacyclic) */
static void negacyclic_mul_add_binary(uint64_t *restrict acc,
                                      const uint64_t *restrict a,
                                      const uint64_t *restrict s,
                                      uint32_t N) {
  for (uint32_t i = 0; i < N; i++) {
    if (!s[i])
      continue;
    for (uint32_t j = 0; j < N - i; j++)
      acc[i + j] += a[j];
    for (uint32_t j = N - i; j < N; j++)
      acc[i + j - N] -= a[j];
  }
}

void orc_glwe_encrypt_assign(orc_rng *rng, const uint64_t *glwe_key,
                             uint32_t k, uint32_t N, int32_t noise_log2,
                             uint64_t *mask_out, uint64_t *body_inout) {
  orc_fill_uniform(rng, mask_out, (size_t)k * N);
  for (uint32_t r = 0; r < k; r++)
    negacyclic_mul_add_binary(body_inout, mask_out + (size_t)r * N,
                              glwe_key + (size_t)r * N, N);
  if (noise_log2 >= 0)
    for (uint32_t j = 0; j < N; j++)
      body_inout[j] += (uint64_t)orc_tuniform(rng, (uint32_t)noise_log2);
}

/* one GGSW of the cleartext `m` (any u64; a key bit or a product of bits) */
static void encrypt_constant_ggsw(orc_rng *rng, const uint64_t *glwe_key,
                                  uint32_t k, uint32_t N, uint32_t base_log,
                                  uint32_t level_count, int32_t noise_log2,
                                  uint64_t m, uint64_t *ggsw_out) {
  const size_t row_len = (size_t)(k + 1) * N;
  for (uint32_t t = 0; t < level_count; t++) {
    const uint32_t level = level_count - t;
    /* factor = (-m) * q / B^level  (ggsw_encryption.rs:20-44) */
    const uint64_t factor = ((uint64_t)0 - m) << (W - base_log * level);
    for (uint32_t r = 0; r <= k; r++) {
      uint64_t *row = ggsw_out + ((size_t)t * (k + 1) + r) * row_len;
      uint64_t *body = row + (size_t)k * N;
      if (r < k) {
        for (uint32_t j = 0; j < N; j++)
          body[j] = glwe_key[(size_t)r * N + j] * factor;
      } else {
        memset(body, 0, (size_t)N * sizeof(uint64_t));
        body[0] = (uint64_t)0 - factor;
      }
      orc_glwe_encrypt_assign(rng, glwe_key, k, N, noise_log2, row, body);
    }
  }
}

void orc_gen_bsk(orc_rng *rng, const uint64_t *lwe_key, uint32_t n,
                 const uint64_t *glwe_key, uint32_t k, uint32_t N,
                 uint32_t base_log, uint32_t level_count, int32_t noise_log2,
                 uint64_t *bsk_out) {
  const size_t ggsw_len = (size_t)level_count * (k + 1) * (k + 1) * N;
  /* independent child generators so generation can run in parallel and
   * stays deterministic for a given parent state */
  uint64_t *seeds = (uint64_t *)malloc((size_t)n * sizeof(uint64_t));
  for (uint32_t i = 0; i < n; i++)
    seeds[i] = orc_rng_next(rng);
#pragma omp parallel for schedule(dynamic, 4)
  for (uint32_t i = 0; i < n; i++) {
    orc_rng child;
    orc_rng_seed(&child, seeds[i]);
    encrypt_constant_ggsw(&child, glwe_key, k, N, base_log, level_count,
                          noise_log2, lwe_key[i], bsk_out + i * ggsw_len);
  }
  free(seeds);
}

void orc_gen_multi_bit_bsk(orc_rng *rng, const uint64_t *lwe_key, uint32_t n,
                           const uint64_t *glwe_key, uint32_t k, uint32_t N,
                           uint32_t base_log, uint32_t level_count,
                           uint32_t grouping_factor, int32_t noise_log2,
                           uint64_t *bsk_out) {
  const uint32_t g = grouping_factor;
  const uint32_t per_group = 1u << g;
  const uint32_t groups = n / g;
  const size_t ggsw_len = (size_t)level_count * (k + 1) * (k + 1) * N;
  const uint32_t total = groups * per_group;
  uint64_t *seeds = (uint64_t *)malloc((size_t)total * sizeof(uint64_t));
  for (uint32_t i = 0; i < total; i++)
    seeds[i] = orc_rng_next(rng);
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
    orc_rng child;
    orc_rng_seed(&child, seeds[idx]);
    encrypt_constant_ggsw(&child, glwe_key, k, N, base_log, level_count,
                          noise_log2, m, bsk_out + idx * ggsw_len);
  }
  free(seeds);
}

void orc_gen_ksk(orc_rng *rng, const uint64_t *key_in, uint32_t dim_in,
                 const uint64_t *key_out, uint32_t dim_out, uint32_t base_log,
                 uint32_t level_count, int32_t noise_log2, uint64_t *ksk_out) {
  const size_t lwe_len = (size_t)dim_out + 1;
  for (uint32_t i = 0; i < dim_in; i++)
    for (uint32_t j = 0; j < level_count; j++) {
      const uint32_t level = level_count - j;
      const uint64_t pt = key_in[i] << (W - base_log * level);
      orc_lwe_encrypt(rng, key_out, dim_out, pt, noise_log2,
                      ksk_out + ((size_t)i * level_count + j) * lwe_len);
    }
}

/* ====================================================================== */
/* keyswitch                                                               */
/* ====================================================================== */
void orc_keyswitch(const uint64_t *ksk, uint32_t dim_in, uint32_t dim_out,
                   uint32_t base_log, uint32_t level_count,
                   const uint64_t *ct_in, uint64_t *restrict ct_out) {
  const size_t lwe_len = (size_t)dim_out + 1;
  memset(ct_out, 0, lwe_len * sizeof(uint64_t));
  ct_out[dim_out] = ct_in[dim_in];
  for (uint32_t i = 0; i < dim_in; i++) {
    uint64_t st = orc_decomposer_init_state(ct_in[i], base_log, level_count);
    for (uint32_t j = 0; j < level_count; j++) {
      const uint64_t d = (uint64_t)decompose_one_level(&st, base_log);
      if (d == 0)
        continue;
      const uint64_t *restrict row =
          ksk + ((size_t)i * level_count + j) * lwe_len;
      for (size_t o = 0; o < lwe_len; o++)
        ct_out[o] -= d * row[o];
    }
  }
}

void orc_keyswitch_batch(const uint64_t *ksk, uint32_t dim_in,
                         uint32_t dim_out, uint32_t base_log,
                         uint32_t level_count, const uint64_t *cts_in,
                         uint64_t *cts_out, uint32_t count,
                         uint32_t num_threads) {
  (void)num_threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads ? num_threads : orc_max_threads())
  for (uint32_t s = 0; s < count; s++)
    orc_keyswitch(ksk, dim_in, dim_out, base_log, level_count,
                  cts_in + (size_t)s * (dim_in + 1),
                  cts_out + (size_t)s * (dim_out + 1));
}

/* ====================================================================== */
/* modulus switch                                                          */
/* ====================================================================== */
uint64_t orc_centered_ms_body_correction(const uint64_t *ct, uint32_t n,
                                         uint32_t log_modulus) {
  uint64_t sum_half = 0;
  int64_t sum_dbl = 0;
  for (uint32_t i = 0; i < n; i++) {
    const uint64_t a = ct[i];
    const uint64_t rounded = orc_modulus_switch(a, log_modulus)
                             << (W - log_modulus);
    const int64_t err = (int64_t)(rounded - a);
    const int64_t half = err / 2; /* truncates toward zero, like Rust */
    sum_dbl += 2 * half - err;
    sum_half += (uint64_t)half;
  }
  const uint64_t sum_halving = (uint64_t)(sum_dbl / 2);
  sum_half -= sum_halving;
  return sum_half - (1ull << (W - log_modulus - 1));
}

void orc_lwe_modulus_switch(const uint64_t *ct, uint32_t n,
                            uint32_t log_modulus, int centered,
                            uint32_t *ms_out) {
  for (uint32_t i = 0; i < n; i++)
    ms_out[i] = (uint32_t)orc_modulus_switch(ct[i], log_modulus);
  uint64_t body = ct[n];
  if (centered)
    body += orc_centered_ms_body_correction(ct, n, log_modulus);
  ms_out[n] = (uint32_t)orc_modulus_switch(body, log_modulus);
}

/* ====================================================================== */
/* LUT, sample extract                                                     */
/* ====================================================================== */
void orc_make_lut(const uint64_t *f_values, uint32_t p, uint64_t delta,
                  uint32_t k, uint32_t N, uint64_t *glwe_out) {
  const uint32_t box = N / p, half = box / 2;
  uint64_t *tmp = (uint64_t *)malloc((size_t)N * sizeof(uint64_t));
  for (uint32_t i = 0; i < p; i++)
    for (uint32_t j = 0; j < box; j++)
      tmp[i * box + j] = f_values[i] * delta;
  for (uint32_t j = 0; j < half; j++)
    tmp[j] = (uint64_t)0 - tmp[j];
  memset(glwe_out, 0, (size_t)k * N * sizeof(uint64_t));
  uint64_t *body = glwe_out + (size_t)k * N;
  for (uint32_t j = 0; j < N; j++) /* rotate_left(half) */
    body[j] = tmp[(j + half) % N];
  free(tmp);
}

void orc_sample_extract(const uint64_t *glwe, uint32_t k, uint32_t N,
                        uint32_t nth, uint64_t *lwe_out) {
  lwe_out[(size_t)k * N] = glwe[(size_t)k * N + nth];
  for (uint32_t r = 0; r < k; r++) {
    const uint64_t *A = glwe + (size_t)r * N;
    uint64_t *o = lwe_out + (size_t)r * N;
    for (uint32_t t = 0; t <= nth; t++)
      o[t] = A[nth - t];
    for (uint32_t t = nth + 1; t < N; t++)
      o[t] = (uint64_t)0 - A[N + nth - t];
  }
}

/* ====================================================================== */
/* negacyclic FFT (twisted half-size complex transform)                    */
/* ====================================================================== */
struct orc_fft_plan {
  uint32_t N, M, logM;
  uint32_t A, B, logA, logB;   /* M = A * B, four-step split */
  double *twist_re, *twist_im; /* e^{i pi j / N}, j < M */
  double *twA_re, *twA_im;     /* length-A DIF twiddles e^{-2 pi i j/(2h)} at h-1+j */
  double *twB_re, *twB_im;     /* same for length B */
  double *T_re, *T_im;         /* [A][B] inter-step twiddles w_M^{j2 * k1(p)} */
  uint32_t *freq_of_slot;      /* private slot -> natural frequency k */
  uint32_t *slot_of_freq;      /* inverse permutation */
  double *root_re, *root_im;   /* e^{i pi e / N}, e < 2N (monomials) */
};

static uint32_t bitrev_u32(uint32_t i, uint32_t bits) {
  uint32_t r = 0;
  for (uint32_t b = 0; b < bits; b++)
    if (i & (1u << b))
      r |= 1u << (bits - 1 - b);
  return r;
}

static void fill_dif_twiddles(double *re, double *im, uint32_t len) {
  const long double pi = 3.14159265358979323846264338327950288L;
  for (uint32_t h = 1; h < len; h <<= 1)
    for (uint32_t j = 0; j < h; j++) {
      const long double a = -pi * (long double)j / (long double)h;
      re[h - 1 + j] = (double)cosl(a);
      im[h - 1 + j] = (double)sinl(a);
    }
}

orc_fft_plan *orc_fft_plan_new(uint32_t N) {
  orc_fft_plan *p = (orc_fft_plan *)calloc(1, sizeof(*p));
  const uint32_t M = N / 2;
  p->N = N;
  p->M = M;
  uint32_t lg = 0;
  while ((1u << lg) < M)
    lg++;
  p->logM = lg;
  p->logA = (lg + 1) / 2;
  p->logB = lg / 2;
  p->A = 1u << p->logA;
  p->B = 1u << p->logB;
  const uint32_t A = p->A, B = p->B;
  p->twist_re = (double *)xalloc(sizeof(double) * M);
  p->twist_im = (double *)xalloc(sizeof(double) * M);
  p->twA_re = (double *)xalloc(sizeof(double) * (A + 8));
  p->twA_im = (double *)xalloc(sizeof(double) * (A + 8));
  p->twB_re = (double *)xalloc(sizeof(double) * (B + 8));
  p->twB_im = (double *)xalloc(sizeof(double) * (B + 8));
  p->T_re = (double *)xalloc(sizeof(double) * M);
  p->T_im = (double *)xalloc(sizeof(double) * M);
  p->freq_of_slot = (uint32_t *)malloc(sizeof(uint32_t) * M);
  p->slot_of_freq = (uint32_t *)malloc(sizeof(uint32_t) * M);
  p->root_re = (double *)xalloc(sizeof(double) * 2 * N);
  p->root_im = (double *)xalloc(sizeof(double) * 2 * N);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (uint32_t j = 0; j < M; j++) {
    const long double a = pi * (long double)j / (long double)N;
    p->twist_re[j] = (double)cosl(a);
    p->twist_im[j] = (double)sinl(a);
  }
  fill_dif_twiddles(p->twA_re, p->twA_im, A);
  fill_dif_twiddles(p->twB_re, p->twB_im, B);
  /* after the length-A DIF along rows, row position pr holds k1 = bitrev_A(pr);
   * inter-step twiddle w_M^{j2 * k1} = e^{-2 pi i j2 k1 / M} */
  for (uint32_t pr = 0; pr < A; pr++) {
    const uint32_t k1 = bitrev_u32(pr, p->logA);
    for (uint32_t j2 = 0; j2 < B; j2++) {
      const long double a =
          -2.0L * pi * (long double)(((uint64_t)j2 * k1) % M) / (long double)M;
      p->T_re[pr * B + j2] = (double)cosl(a);
      p->T_im[pr * B + j2] = (double)sinl(a);
    }
  }
  /* final layout: Z[q][pr] (B rows, A columns) = X[k1(pr) + A * k2(q)] */
  for (uint32_t q = 0; q < B; q++)
    for (uint32_t pr = 0; pr < A; pr++) {
      const uint32_t k =
          bitrev_u32(pr, p->logA) + A * bitrev_u32(q, p->logB);
      p->freq_of_slot[q * A + pr] = k;
      p->slot_of_freq[k] = q * A + pr;
    }
  for (uint32_t e = 0; e < 2 * N; e++) {
    const long double a = pi * (long double)e / (long double)N;
    p->root_re[e] = (double)cosl(a);
    p->root_im[e] = (double)sinl(a);
  }
  return p;
}

void orc_fft_plan_free(orc_fft_plan *p) {
  if (!p)
    return;
  free(p->twist_re);
  free(p->twist_im);
  free(p->twA_re);
  free(p->twA_im);
  free(p->twB_re);
  free(p->twB_im);
  free(p->T_re);
  free(p->T_im);
  free(p->freq_of_slot);
  free(p->slot_of_freq);
  free(p->root_re);
  free(p->root_im);
  free(p);
}

/* radix-2 DIF along the ROW index of a rows x cols row-major matrix; every
 * butterfly loop runs over `cols` contiguous elements (vectorises).  Row
 * position pr ends up holding frequency bitrev(pr). */
static inline __attribute__((always_inline)) void
col_fft_dif(double *restrict re, double *restrict im, const uint32_t rows,
            const uint32_t cols, const double *twr, const double *twi) {
  for (uint32_t h = rows >> 1; h >= 1; h >>= 1)
    for (uint32_t b = 0; b < rows; b += 2 * h)
      for (uint32_t j = 0; j < h; j++) {
        const double wr = twr[h - 1 + j], wi = twi[h - 1 + j];
        double *restrict r0 = re + (size_t)(b + j) * cols;
        double *restrict i0 = im + (size_t)(b + j) * cols;
        double *restrict r1 = re + (size_t)(b + j + h) * cols;
        double *restrict i1 = im + (size_t)(b + j + h) * cols;
#pragma GCC ivdep
        for (uint32_t c = 0; c < cols; c++) {
          const double ur = r0[c], ui = i0[c], vr = r1[c], vi = i1[c];
          const double dr = ur - vr, di = ui - vi;
          r0[c] = ur + vr;
          i0[c] = ui + vi;
          r1[c] = dr * wr - di * wi;
          i1[c] = dr * wi + di * wr;
        }
      }
}

/* exact inverse (unnormalised) of col_fft_dif: DIT with conjugate twiddles */
static inline __attribute__((always_inline)) void
col_fft_dit_inv(double *restrict re, double *restrict im, const uint32_t rows,
                const uint32_t cols, const double *twr, const double *twi) {
  for (uint32_t h = 1; h < rows; h <<= 1)
    for (uint32_t b = 0; b < rows; b += 2 * h)
      for (uint32_t j = 0; j < h; j++) {
        const double wr = twr[h - 1 + j], wi = twi[h - 1 + j];
        double *restrict r0 = re + (size_t)(b + j) * cols;
        double *restrict i0 = im + (size_t)(b + j) * cols;
        double *restrict r1 = re + (size_t)(b + j + h) * cols;
        double *restrict i1 = im + (size_t)(b + j + h) * cols;
#pragma GCC ivdep
        for (uint32_t c = 0; c < cols; c++) {
          const double vr = r1[c] * wr + i1[c] * wi; /* v * conj(w) */
          const double vi = i1[c] * wr - r1[c] * wi;
          const double ur = r0[c], ui = i0[c];
          r0[c] = ur + vr;
          i0[c] = ui + vi;
          r1[c] = ur - vr;
          i1[c] = ui - vi;
        }
      }
}

#include <immintrin.h>
static void transpose(const double *restrict src, double *restrict dst,
                      uint32_t rows, uint32_t cols) {
  /* dst[c][r] = src[r][c]; 4x4 blocks in registers when the shape allows */
  if ((rows & 3) == 0 && (cols & 3) == 0) {
    for (uint32_t r0 = 0; r0 < rows; r0 += 4)
      for (uint32_t c0 = 0; c0 < cols; c0 += 4) {
        const __m256d a = _mm256_loadu_pd(src + (size_t)(r0 + 0) * cols + c0);
        const __m256d b = _mm256_loadu_pd(src + (size_t)(r0 + 1) * cols + c0);
        const __m256d c = _mm256_loadu_pd(src + (size_t)(r0 + 2) * cols + c0);
        const __m256d d = _mm256_loadu_pd(src + (size_t)(r0 + 3) * cols + c0);
        const __m256d t0 = _mm256_unpacklo_pd(a, b), t1 = _mm256_unpackhi_pd(a, b);
        const __m256d t2 = _mm256_unpacklo_pd(c, d), t3 = _mm256_unpackhi_pd(c, d);
        _mm256_storeu_pd(dst + (size_t)(c0 + 0) * rows + r0, _mm256_permute2f128_pd(t0, t2, 0x20));
        _mm256_storeu_pd(dst + (size_t)(c0 + 1) * rows + r0, _mm256_permute2f128_pd(t1, t3, 0x20));
        _mm256_storeu_pd(dst + (size_t)(c0 + 2) * rows + r0, _mm256_permute2f128_pd(t0, t2, 0x31));
        _mm256_storeu_pd(dst + (size_t)(c0 + 3) * rows + r0, _mm256_permute2f128_pd(t1, t3, 0x31));
      }
    return;
  }
  for (uint32_t r = 0; r < rows; r++)
    for (uint32_t c = 0; c < cols; c++)
      dst[(size_t)c * rows + r] = src[(size_t)r * cols + c];
}

/* in-place forward transform, kernel e^{-2 pi i jk/M}: natural order in,
 * private order out (slot s holds frequency freq_of_slot[s]).  Four-step:
 * length-A transforms down the columns of the A x B input, twiddle,
 * transpose, length-B transforms down the columns of the B x A matrix. */
static void fft_dif(const orc_fft_plan *p, double *restrict re,
                    double *restrict im) {
  const uint32_t M = p->M, A = p->A, B = p->B;
  double tre[M] __attribute__((aligned(64))), tim[M] __attribute__((aligned(64)));
  if (A == 32 && B == 32)
    col_fft_dif(re, im, 32, 32, p->twA_re, p->twA_im);
  else
    col_fft_dif(re, im, A, B, p->twA_re, p->twA_im);
  const double *restrict Tr = p->T_re, *restrict Ti = p->T_im;
#pragma GCC ivdep
  for (uint32_t i = 0; i < M; i++) {
    const double xr = re[i], xi = im[i];
    re[i] = xr * Tr[i] - xi * Ti[i];
    im[i] = xr * Ti[i] + xi * Tr[i];
  }
  transpose(re, tre, A, B);
  transpose(im, tim, A, B);
  if (A == 32 && B == 32)
    col_fft_dif(tre, tim, 32, 32, p->twB_re, p->twB_im);
  else
    col_fft_dif(tre, tim, B, A, p->twB_re, p->twB_im);
  memcpy(re, tre, sizeof(double) * M);
  memcpy(im, tim, sizeof(double) * M);
}

/* exact inverse of fft_dif up to the factor M: private order in, natural out */
static void fft_dit_inv(const orc_fft_plan *p, double *restrict re,
                        double *restrict im) {
  const uint32_t M = p->M, A = p->A, B = p->B;
  double tre[M] __attribute__((aligned(64))), tim[M] __attribute__((aligned(64)));
  if (A == 32 && B == 32)
    col_fft_dit_inv(re, im, 32, 32, p->twB_re, p->twB_im);
  else
    col_fft_dit_inv(re, im, B, A, p->twB_re, p->twB_im);
  transpose(re, tre, B, A);
  transpose(im, tim, B, A);
  const double *restrict Tr = p->T_re, *restrict Ti = p->T_im;
#pragma GCC ivdep
  for (uint32_t i = 0; i < M; i++) {
    const double xr = tre[i], xi = tim[i];
    re[i] = xr * Tr[i] + xi * Ti[i]; /* * conj(T) */
    im[i] = xi * Tr[i] - xr * Ti[i];
  }
  if (A == 32 && B == 32)
    col_fft_dit_inv(re, im, 32, 32, p->twA_re, p->twA_im);
  else
    col_fft_dit_inv(re, im, A, B, p->twA_re, p->twA_im);
}

/* private order forward transforms */
/* exact for |x| < 2^51: plant the integer in the mantissa of 1.5 * 2^52 */
static inline double i64_to_double_small(int64_t x) {
  union { int64_t i; double d; } u;
  u.i = x + 0x4338000000000000ll;
  return u.d - 6755399441055744.0;
}

static void fwd_integer_priv(const orc_fft_plan *p, const int64_t *poly,
                             double *restrict re, double *restrict im) {
  const uint32_t M = p->M;
#pragma GCC ivdep
  for (uint32_t j = 0; j < M; j++) {
    const double a = i64_to_double_small(poly[j]), b = i64_to_double_small(poly[j + M]);
    re[j] = a * p->twist_re[j] - b * p->twist_im[j];
    im[j] = a * p->twist_im[j] + b * p->twist_re[j];
  }
  fft_dif(p, re, im);
}

static void fwd_torus_priv(const orc_fft_plan *p, const uint64_t *poly,
                           double *restrict re, double *restrict im) {
  const uint32_t M = p->M;
  const double norm = 0x1p-64;
  for (uint32_t j = 0; j < M; j++) {
    const double a = (double)(int64_t)poly[j] * norm;
    const double b = (double)(int64_t)poly[j + M] * norm;
    re[j] = a * p->twist_re[j] - b * p->twist_im[j];
    im[j] = a * p->twist_im[j] + b * p->twist_re[j];
  }
  fft_dif(p, re, im);
}

/* torus/mod.rs:75-81.  The reference rounds half away from zero (Rust
 * f64::round); round-to-nearest-even is used here so the loop vectorises.
 * The two differ only on exact .5 ties of the input, where frac = +-0.5 in
 * both cases and the result is +-2^63 (same torus element). */
static inline uint64_t from_torus(double x) {
  const double magic = 6755399441055744.0; /* 1.5 * 2^52 */
  const double r = (x + magic) - magic;    /* rint(x), |x| < 2^51 */
  const double fract = (x - r) * 0x1p64;
  if (fract >= 0x1p63)
    return (uint64_t)INT64_MAX; /* Rust `as` saturates */
  if (fract <= -0x1p63)
    return (uint64_t)INT64_MIN;
  return (uint64_t)(int64_t)llrint(fract);
}

/* spectrum in private order, destroyed */
static void add_backward_torus_priv(const orc_fft_plan *p, double *restrict re,
                                    double *restrict im,
                                    uint64_t *poly_inout) {
  const uint32_t M = p->M;
  fft_dit_inv(p, re, im);
  const double norm = 1.0 / (double)M;
  const double magic = 6755399441055744.0; /* 1.5 * 2^52 */
  const double *restrict twr = p->twist_re, *restrict twi = p->twist_im;
  /* pass 1 (vectorises): untwist, scale, reduce mod 1, scale to 2^64 and
   * clamp into the i64 range (Rust `as` saturates, torus/mod.rs:75-81) */
#pragma GCC ivdep
  for (uint32_t j = 0; j < M; j++) {
    const double wr = twr[j] * norm, wi = -twi[j] * norm;
    const double tr = re[j] * wr - im[j] * wi;
    const double ti = re[j] * wi + im[j] * wr;
    double fr = (tr - ((tr + magic) - magic)) * 0x1p64;
    double fi = (ti - ((ti + magic) - magic)) * 0x1p64;
    fr = fr > 0x1.fffffffffffffp62 ? 0x1.fffffffffffffp62 : fr;
    fi = fi > 0x1.fffffffffffffp62 ? 0x1.fffffffffffffp62 : fi;
    re[j] = fr < -0x1p63 ? -0x1p63 : fr;
    im[j] = fi < -0x1p63 ? -0x1p63 : fi;
  }
  /* pass 2: exact conversions (the values are integers already); two plain
   * loops so that an AVX-512DQ build turns them into vcvttpd2qq */
  uint64_t *restrict lo = poly_inout, *restrict hi = poly_inout + M;
#pragma GCC ivdep
  for (uint32_t j = 0; j < M; j++)
    lo[j] += (uint64_t)(int64_t)re[j];
#pragma GCC ivdep
  for (uint32_t j = 0; j < M; j++)
    hi[j] += (uint64_t)(int64_t)im[j];
}

static void to_natural(const orc_fft_plan *p, const double *re,
                       const double *im, double *out_re, double *out_im) {
  for (uint32_t k = 0; k < p->M; k++) {
    out_re[k] = re[p->slot_of_freq[k]];
    out_im[k] = im[p->slot_of_freq[k]];
  }
}

void orc_fft_forward_integer(const orc_fft_plan *p, const int64_t *poly,
                             double *out_re, double *out_im) {
  double *re = (double *)malloc(sizeof(double) * p->M * 2), *im = re + p->M;
  fwd_integer_priv(p, poly, re, im);
  to_natural(p, re, im, out_re, out_im);
  free(re);
}

void orc_fft_forward_real(const orc_fft_plan *p, const double *poly,
                          double *out_re, double *out_im) {
  const uint32_t M = p->M;
  double *re = (double *)malloc(sizeof(double) * M * 2), *im = re + M;
  for (uint32_t j = 0; j < M; j++) {
    const double a = poly[j], b = poly[j + M];
    re[j] = a * p->twist_re[j] - b * p->twist_im[j];
    im[j] = a * p->twist_im[j] + b * p->twist_re[j];
  }
  fft_dif(p, re, im);
  to_natural(p, re, im, out_re, out_im);
  free(re);
}

void orc_fft_forward_torus(const orc_fft_plan *p, const uint64_t *poly,
                           double *out_re, double *out_im) {
  double *re = (double *)malloc(sizeof(double) * p->M * 2), *im = re + p->M;
  fwd_torus_priv(p, poly, re, im);
  to_natural(p, re, im, out_re, out_im);
  free(re);
}

void orc_fft_add_backward_torus(const orc_fft_plan *p, const double *in_re,
                                const double *in_im, uint64_t *poly_inout) {
  const uint32_t M = p->M;
  double *re = (double *)malloc(sizeof(double) * M * 2), *im = re + M;
  for (uint32_t k = 0; k < M; k++) { /* natural -> private */
    re[p->slot_of_freq[k]] = in_re[k];
    im[p->slot_of_freq[k]] = in_im[k];
  }
  add_backward_torus_priv(p, re, im, poly_inout);
  free(re);
}

void orc_bsk_to_fourier(const orc_fft_plan *p, const uint64_t *bsk,
                        size_t poly_count, double *out_re, double *out_im) {
  const uint32_t N = p->N, M = p->M;
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < poly_count; q++)
    fwd_torus_priv(p, bsk + q * N, out_re + q * M, out_im + q * M);
}

/* ====================================================================== */
/* external product / blind rotation                                       */
/* ====================================================================== */
typedef struct {
  int64_t *digits;  /* level_count * (k+1) * N */
  double *fre, *fim; /* level_count * (k+1) * M */
  double *ore, *oim; /* (k+1) * M */
  uint64_t *ct1;     /* (k+1) * N */
  uint64_t *tmp;     /* (k+1) * N */
  double *gre, *gim; /* multi-bit bundle: level*(k+1)^2*M */
  uint64_t *gstd;    /* multi-bit exact bundle: level*(k+1)^2*N */
} scratch_t;

static void scratch_init(scratch_t *s, uint32_t k, uint32_t N, uint32_t l,
                         int multibit, int exact) {
  const uint32_t M = N / 2;
  memset(s, 0, sizeof(*s));
  s->digits = (int64_t *)xalloc(sizeof(int64_t) * l * (k + 1) * N);
  s->fre = (double *)xalloc(sizeof(double) * l * (k + 1) * M);
  s->fim = (double *)xalloc(sizeof(double) * l * (k + 1) * M);
  s->ore = (double *)xalloc(sizeof(double) * (k + 1) * M);
  s->oim = (double *)xalloc(sizeof(double) * (k + 1) * M);
  s->ct1 = (uint64_t *)xalloc(sizeof(uint64_t) * (k + 1) * N);
  s->tmp = (uint64_t *)xalloc(sizeof(uint64_t) * (k + 1) * N);
  if (multibit && !exact) {
    const size_t g = (size_t)l * (k + 1) * (k + 1) * M;
    s->gre = (double *)xalloc(sizeof(double) * g);
    s->gim = (double *)xalloc(sizeof(double) * g);
  }
  if (multibit && exact)
    s->gstd = (uint64_t *)xalloc(sizeof(uint64_t) * (size_t)l * (k + 1) * (k + 1) * N);
}

static void scratch_free(scratch_t *s) {
  free(s->digits);
  free(s->fre);
  free(s->fim);
  free(s->ore);
  free(s->oim);
  free(s->ct1);
  free(s->tmp);
  free(s->gre);
  free(s->gim);
  free(s->gstd);
}

/* ggsw.rs:519-537 -- digits[t][r][j], t = 0 is level l */
static void decompose_glwe(const uint64_t *glwe, uint32_t k, uint32_t N,
                           uint32_t base_log, uint32_t l, int64_t *digits) {
  const size_t plane = (size_t)(k + 1) * N;
  if (l == 1) {
    /* single level: the digit is the balanced closest representable itself;
     * branch-free form of decomposer.rs:163-188 that the compiler vectorises */
    const uint32_t R = base_log, sh = W - R - 1;
    const uint64_t mask = (1ull << R) - 1;
#pragma GCC ivdep
    for (size_t idx = 0; idx < plane; idx++) {
      uint64_t r = glwe[idx] >> sh;
      const uint64_t rb = r & 1;
      r = ((r + 1) >> 1) & mask;
      const uint64_t bal = (((r - 1) | (rb << (R - 1))) & r) >> (R - 1);
      digits[idx] = (int64_t)(r - (bal << R));
    }
    return;
  }
  for (size_t idx = 0; idx < plane; idx++) {
    uint64_t st = orc_decomposer_init_state(glwe[idx], base_log, l);
    for (uint32_t t = 0; t < l; t++)
      digits[t * plane + idx] = decompose_one_level(&st, base_log);
  }
}

/* acc += ggsw (x) glwe_in, Fourier mode.  ggsw planes: [t][r][c][M] */
static void ext_product_fft(const orc_fft_plan *p, scratch_t *s, uint64_t *acc,
                            const double *restrict gre,
                            const double *restrict gim,
                            const uint64_t *glwe_in, uint32_t k, uint32_t N,
                            uint32_t base_log, uint32_t l) {
  const uint32_t M = N / 2;
  decompose_glwe(glwe_in, k, N, base_log, l, s->digits);
  for (uint32_t q = 0; q < l * (k + 1); q++)
    fwd_integer_priv(p, s->digits + (size_t)q * N, s->fre + (size_t)q * M,
                     s->fim + (size_t)q * M);
  for (uint32_t c = 0; c <= k; c++) {
    double *restrict ore = s->ore + (size_t)c * M;
    double *restrict oim = s->oim + (size_t)c * M;
    int first = 1;
    for (uint32_t t = 0; t < l; t++)
      for (uint32_t r = 0; r <= k; r++) {
        const double *restrict fr = s->fre + ((size_t)t * (k + 1) + r) * M;
        const double *restrict fi = s->fim + ((size_t)t * (k + 1) + r) * M;
        const size_t off = (((size_t)t * (k + 1) + r) * (k + 1) + c) * M;
        const double *restrict br = gre + off, *restrict bi = gim + off;
        if (first) {
          for (uint32_t j = 0; j < M; j++) {
            ore[j] = fr[j] * br[j] - fi[j] * bi[j];
            oim[j] = fr[j] * bi[j] + fi[j] * br[j];
          }
          first = 0;
        } else {
          for (uint32_t j = 0; j < M; j++) {
            ore[j] += fr[j] * br[j] - fi[j] * bi[j];
            oim[j] += fr[j] * bi[j] + fi[j] * br[j];
          }
        }
      }
    add_backward_torus_priv(p, ore, oim, acc + (size_t)c * N);
  }
}

/* exact twin; ggsw standard layout [t][r][c][N] */
static void ext_product_exact(scratch_t *s, uint64_t *acc,
                              const uint64_t *ggsw, const uint64_t *glwe_in,
                              uint32_t k, uint32_t N, uint32_t base_log,
                              uint32_t l) {
  decompose_glwe(glwe_in, k, N, base_log, l, s->digits);
  for (uint32_t t = 0; t < l; t++)
    for (uint32_t r = 0; r <= k; r++)
      for (uint32_t c = 0; c <= k; c++)
        orc_negacyclic_mul_add_exact(
            acc + (size_t)c * N, s->digits + ((size_t)t * (k + 1) + r) * N,
            ggsw + (((size_t)t * (k + 1) + r) * (k + 1) + c) * N, N);
}

void orc_add_external_product_fft(const orc_fft_plan *p, uint64_t *acc,
                                  const double *ggsw_re, const double *ggsw_im,
                                  const uint64_t *glwe_in, uint32_t k,
                                  uint32_t N, uint32_t base_log, uint32_t l) {
  scratch_t s;
  scratch_init(&s, k, N, l, 0, 0);
  ext_product_fft(p, &s, acc, ggsw_re, ggsw_im, glwe_in, k, N, base_log, l);
  scratch_free(&s);
}

void orc_add_external_product_exact(uint64_t *acc, const uint64_t *ggsw,
                                    const uint64_t *glwe_in, uint32_t k,
                                    uint32_t N, uint32_t base_log,
                                    uint32_t l) {
  scratch_t s;
  scratch_init(&s, k, N, l, 0, 1);
  ext_product_exact(&s, acc, ggsw, glwe_in, k, N, base_log, l);
  scratch_free(&s);
}

static void rotate_by_body(scratch_t *s, uint64_t *acc, uint32_t k, uint32_t N,
                           uint32_t b_hat) {
  for (uint32_t r = 0; r <= k; r++) {
    memcpy(s->tmp, acc + (size_t)r * N, sizeof(uint64_t) * N);
    orc_monomial_div(acc + (size_t)r * N, s->tmp, N, b_hat);
  }
}

static void blind_rotate_impl(const orc_fft_plan *p, scratch_t *s,
                              uint64_t *acc, const uint32_t *ms,
                              const uint64_t *bsk_std, const double *bsk_re,
                              const double *bsk_im, uint32_t n, uint32_t k,
                              uint32_t N, uint32_t base_log, uint32_t l,
                              int exact) {
  const uint32_t M = N / 2;
  const size_t ggsw_polys = (size_t)l * (k + 1) * (k + 1);
  rotate_by_body(s, acc, k, N, ms[n]);
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t a = ms[i];
    if (a == 0)
      continue;
    for (uint32_t r = 0; r <= k; r++)
      orc_monomial_mul_and_subtract(s->ct1 + (size_t)r * N,
                                    acc + (size_t)r * N, N, a);
    if (exact)
      ext_product_exact(s, acc, bsk_std + i * ggsw_polys * N, s->ct1, k, N,
                        base_log, l);
    else
      ext_product_fft(p, s, acc, bsk_re + i * ggsw_polys * M,
                      bsk_im + i * ggsw_polys * M, s->ct1, k, N, base_log, l);
  }
}

void orc_blind_rotate_fft(const orc_fft_plan *p, uint64_t *acc,
                          const uint32_t *ms, const double *bsk_re,
                          const double *bsk_im, uint32_t n, uint32_t k,
                          uint32_t N, uint32_t base_log, uint32_t l) {
  scratch_t s;
  scratch_init(&s, k, N, l, 0, 0);
  blind_rotate_impl(p, &s, acc, ms, NULL, bsk_re, bsk_im, n, k, N, base_log, l,
                    0);
  scratch_free(&s);
}

void orc_blind_rotate_exact(uint64_t *acc, const uint32_t *ms,
                            const uint64_t *bsk, uint32_t n, uint32_t k,
                            uint32_t N, uint32_t base_log, uint32_t l) {
  scratch_t s;
  scratch_init(&s, k, N, l, 0, 1);
  blind_rotate_impl(NULL, &s, acc, ms, bsk, NULL, NULL, n, k, N, base_log, l,
                    1);
  scratch_free(&s);
}

/* ------------------------------------------------------------ multi-bit -- */
/* degree of subset s (1 <= s < 2^g) of group `grp`
 * (lwe_multi_bit_programmable_bootstrapping.rs:30-65) */
static uint32_t multi_bit_degree(const uint64_t *ct_in, uint32_t grp,
                                 uint32_t g, uint32_t sel, uint32_t log_mod) {
  uint64_t sum = 0;
  for (uint32_t u = 0; u < g; u++) {
    const uint32_t pos = g - (u + 1);
    if ((sel >> pos) & 1u)
      sum += ct_in[grp * g + u];
  }
  return (uint32_t)orc_modulus_switch(sum, log_mod);
}

static void multi_bit_blind_rotate_impl(const orc_fft_plan *p, scratch_t *s,
                                        uint64_t *acc, const uint64_t *ct_in,
                                        const uint64_t *bsk_std,
                                        const double *bsk_re,
                                        const double *bsk_im, uint32_t n,
                                        uint32_t k, uint32_t N,
                                        uint32_t base_log, uint32_t l,
                                        uint32_t g, int exact) {
  const uint32_t M = N / 2;
  uint32_t log_mod = 1;
  while ((1u << log_mod) < 2 * N)
    log_mod++;
  const uint32_t per_group = 1u << g, groups = n / g;
  const size_t ggsw_polys = (size_t)l * (k + 1) * (k + 1);
  rotate_by_body(s, acc, k, N,
                 (uint32_t)orc_modulus_switch(ct_in[n], log_mod));
  for (uint32_t grp = 0; grp < groups; grp++) {
    const size_t base = (size_t)grp * per_group * ggsw_polys;
    memcpy(s->ct1, acc, sizeof(uint64_t) * (k + 1) * N);
    memset(acc, 0, sizeof(uint64_t) * (k + 1) * N);
    if (!exact) {
      /* bundle = GGSW_0 + sum_s GGSW_s * X^{deg_s} in the Fourier domain
       * (prepare_multi_bit_ggsw_mem_optimized, :116-156) */
      memcpy(s->gre, bsk_re + base * M, sizeof(double) * ggsw_polys * M);
      memcpy(s->gim, bsk_im + base * M, sizeof(double) * ggsw_polys * M);
      for (uint32_t sel = 1; sel < per_group; sel++) {
        const uint32_t deg = multi_bit_degree(ct_in, grp, g, sel, log_mod);
        const double *br = bsk_re + (base + sel * ggsw_polys) * M;
        const double *bi = bsk_im + (base + sel * ggsw_polys) * M;
        for (uint32_t pos = 0; pos < M; pos++) {
          /* spectrum slot `pos` is frequency kf = freq_of_slot[pos], root
           * e^{i pi (1 - 4 kf)/N}; X^deg there = e^{i pi (1-4kf) deg / N} */
          const uint32_t kf = p->freq_of_slot[pos];
          const uint32_t e =
              (uint32_t)(((int64_t)deg * (1 - 4 * (int64_t)kf)) &
                         (int64_t)(2 * N - 1));
          const double mr = p->root_re[e], mi = p->root_im[e];
          for (size_t q = 0; q < ggsw_polys; q++) {
            const double xr = br[q * M + pos], xi = bi[q * M + pos];
            s->gre[q * M + pos] += xr * mr - xi * mi;
            s->gim[q * M + pos] += xr * mi + xi * mr;
          }
        }
      }
      ext_product_fft(p, s, acc, s->gre, s->gim, s->ct1, k, N, base_log, l);
    } else {
      memcpy(s->gstd, bsk_std + base * N, sizeof(uint64_t) * ggsw_polys * N);
      for (uint32_t sel = 1; sel < per_group; sel++) {
        const uint32_t deg = multi_bit_degree(ct_in, grp, g, sel, log_mod);
        const uint64_t *src = bsk_std + (base + sel * ggsw_polys) * N;
        for (size_t q = 0; q < ggsw_polys; q++) {
          /* += src * X^deg : reuse monomial_div with 2N - deg */
          orc_monomial_div(s->tmp, src + q * N, N, (2 * N - deg) % (2 * N));
          for (uint32_t j = 0; j < N; j++)
            s->gstd[q * N + j] += s->tmp[j];
        }
      }
      ext_product_exact(s, acc, s->gstd, s->ct1, k, N, base_log, l);
    }
  }
}

void orc_multi_bit_blind_rotate_fft(const orc_fft_plan *p, uint64_t *acc,
                                    const uint64_t *ct_in,
                                    const double *bsk_re, const double *bsk_im,
                                    uint32_t n, uint32_t k, uint32_t N,
                                    uint32_t base_log, uint32_t l,
                                    uint32_t g) {
  scratch_t s;
  scratch_init(&s, k, N, l, 1, 0);
  multi_bit_blind_rotate_impl(p, &s, acc, ct_in, NULL, bsk_re, bsk_im, n, k, N,
                              base_log, l, g, 0);
  scratch_free(&s);
}

void orc_multi_bit_blind_rotate_exact(uint64_t *acc, const uint64_t *ct_in,
                                      const uint64_t *bsk, uint32_t n,
                                      uint32_t k, uint32_t N,
                                      uint32_t base_log, uint32_t l,
                                      uint32_t g) {
  scratch_t s;
  scratch_init(&s, k, N, l, 1, 1);
  multi_bit_blind_rotate_impl(NULL, &s, acc, ct_in, bsk, NULL, NULL, n, k, N,
                              base_log, l, g, 1);
  scratch_free(&s);
}

/* ====================================================================== */
/* batched PBS                                                             */
/* ====================================================================== */
uint32_t orc_max_threads(void) {
#ifdef _OPENMP
  return (uint32_t)omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_pbs_batch(const orc_fft_plan *p, const orc_pbs_params *prm,
                   const uint64_t *bsk_std, const double *bsk_re,
                   const double *bsk_im, const uint64_t *luts,
                   const uint64_t *lut_idx, const uint64_t *cts_in,
                   const uint64_t *in_idx, uint64_t *cts_out,
                   const uint64_t *out_idx, uint32_t count, int exact,
                   uint32_t num_threads) {
  const uint32_t n = prm->n, k = prm->k, N = prm->N, l = prm->level_count;
  const uint32_t g = prm->grouping_factor;
  const int multibit = g > 1;
  const size_t glwe_len = (size_t)(k + 1) * N, out_len = (size_t)k * N + 1;
  const uint32_t many = prm->num_many_lut ? prm->num_many_lut : 1;
  uint32_t log_mod = 1;
  while ((1u << log_mod) < 2 * N)
    log_mod++;
  if (!num_threads)
    num_threads = orc_max_threads();
#pragma omp parallel num_threads(num_threads)
  {
    scratch_t s;
    scratch_init(&s, k, N, l, multibit, exact);
    uint64_t *acc = (uint64_t *)xalloc(sizeof(uint64_t) * glwe_len);
    uint32_t *ms = (uint32_t *)malloc(sizeof(uint32_t) * (n + 1));
#pragma omp for schedule(dynamic, 1)
    for (uint32_t sidx = 0; sidx < count; sidx++) {
      const uint64_t ii = in_idx ? in_idx[sidx] : sidx;
      const uint64_t oi = out_idx ? out_idx[sidx] : sidx;
      const uint64_t li = lut_idx ? lut_idx[sidx] : 0;
      const uint64_t *ct = cts_in + ii * (n + 1);
      memcpy(acc, luts + li * glwe_len, sizeof(uint64_t) * glwe_len);
      if (multibit) {
        multi_bit_blind_rotate_impl(p, &s, acc, ct, bsk_std, bsk_re, bsk_im, n,
                                    k, N, prm->base_log, l, g, exact);
      } else {
        orc_lwe_modulus_switch(ct, n, log_mod, prm->centered_ms, ms);
        blind_rotate_impl(p, &s, acc, ms, bsk_std, bsk_re, bsk_im, n, k, N,
                          prm->base_log, l, exact);
      }
      for (uint32_t j = 0; j < many; j++)
        orc_sample_extract(acc, k, N, j * prm->lut_stride,
                           cts_out + ((size_t)j * count + oi) * out_len);
    }
    free(ms);
    free(acc);
    scratch_free(&s);
  }
}
