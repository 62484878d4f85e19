/*
 * pbs_oracle.h -- CPU restatement of the tfhe-rs core_crypto PBS path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * as the timed CPU baseline.  The shipped path (tfhe-rs_b200/) never links
 * or dlopens it and has no CPU fallback.
 *
 * PARITY STATUS: pinned against every golden the checkout holds; "unpinned"
 * only at the level of PBS output WORDS, which the reference itself declares
 * reproducible only on the GPU generation that produced them
 * (gpu/algorithms/test/pbs_golden/mod.rs:68-80).  The reference is Rust and
 * cannot be built in this image (no cargo/rustc) and its apps/test-vectors
 * .cbor files are Git-LFS pointers.  What IS pinned: (1) the forward
 * negacyclic transform against the reference's committed
 * fft16x4x16_golden_v1 vector (tolerance KAT), (2) every doc-test example the
 * reference carries for the integer routines restated here (decomposer,
 * monomial mul/div, sample extract), (3) the reference's own semantic
 * assertion decrypt(PBS(Enc(m))) == f(m), (4) FFT-mode vs exact-integer mode
 * agreement, (5) via tfhe_csprng.c (bit-exact on the tfhe-csprng KATs): the
 * keys of the reference's pbs_golden run are regenerated from its seed and
 * decrypt the committed H100 ciphertexts to f(m); the oracle bootstraps the
 * regenerated inputs with the regenerated BSK to the same messages.
 * See DESIGN.md section "Oracle".
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose algorithm it restates.  All torus arithmetic is u64 wrapping.
 */
#ifndef PBS_ORACLE_H
#define PBS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ----------------------------------------------------------------- PRNG -- */
/* Own generator (xoshiro256** seeded through splitmix64).  NOT tfhe-csprng:
 * keys/inputs are therefore not those of the reference's seeded tests.      */
typedef struct {
  uint64_t s[4];
} orc_rng;

void orc_rng_seed(orc_rng *rng, uint64_t seed);
uint64_t orc_rng_next(orc_rng *rng);
void orc_fill_uniform(orc_rng *rng, uint64_t *out, size_t count);
void orc_fill_binary(orc_rng *rng, uint64_t *out, size_t count);
/* TUniform(bound_log2): uniform on [-2^b, 2^b] with halved end weights
 * (tfhe/src/core_crypto/commons/math/random/t_uniform.rs:85-112).          */
int64_t orc_tuniform(orc_rng *rng, uint32_t bound_log2);

/* ------------------------------------------------------- small routines -- */
/* tfhe/src/core_crypto/fft_impl/common.rs:10-23                             */
uint64_t orc_modulus_switch(uint64_t x, uint32_t log_modulus);
/* decomposer.rs:163-188 (init state) and iter.rs:131-151 (one level).
 * digits_out[0] is level `level_count` (smallest weight), last is level 1.  */
uint64_t orc_decomposer_init_state(uint64_t x, uint32_t base_log,
                                   uint32_t level_count);
void orc_decompose(uint64_t x, uint32_t base_log, uint32_t level_count,
                   int64_t *digits_out);
/* decomposer.rs closest_representable */
uint64_t orc_closest_representable(uint64_t x, uint32_t base_log,
                                   uint32_t level_count);

/* polynomial_algorithms.rs:544-583 : out = in * X^{-d}  (d in [0,2N))       */
void orc_monomial_div(uint64_t *out, const uint64_t *in, uint32_t N,
                      uint32_t d);
/* polynomial_algorithms.rs:662-730 : out = in * X^{d} - in                   */
void orc_monomial_mul_and_subtract(uint64_t *out, const uint64_t *in,
                                   uint32_t N, uint32_t d);
/* exact negacyclic product out (+)= a_signed * b, schoolbook
 * (role of polynomial_karatsuba_wrapping_mul, polynomial_algorithms.rs:1106) */
void orc_negacyclic_mul_add_exact(uint64_t *out, const int64_t *a,
                                  const uint64_t *b, uint32_t N);

/* --------------------------------------------------------- encryption ---- */
/* lwe_encryption.rs:99-113.  noise_log2 < 0 => noiseless.                   */
void orc_lwe_encrypt(orc_rng *rng, const uint64_t *key, uint32_t n,
                     uint64_t plaintext, int32_t noise_log2, uint64_t *ct_out);
/* lwe_encryption.rs:560-561 : returns b - <a,s>                              */
uint64_t orc_lwe_decrypt(const uint64_t *key, uint32_t n, const uint64_t *ct);
/* glwe_encryption.rs:424-447 ; body_inout holds the plaintext polynomial on
 * entry, mask_out gets k uniform polynomials.                                */
void orc_glwe_encrypt_assign(orc_rng *rng, const uint64_t *glwe_key,
                             uint32_t k, uint32_t N, int32_t noise_log2,
                             uint64_t *mask_out, uint64_t *body_inout);

/* Standard-domain BSK, layout [i<n][level idx t][row r<=k][poly c<=k][N],
 * level idx t holds level (l - t)   (ggsw_encryption.rs:20-44,141-159,
 * 361-412; entities/lwe_bootstrap_key.rs:130-142).                          */
void orc_gen_bsk(orc_rng *rng, const uint64_t *lwe_key, uint32_t n,
                 const uint64_t *glwe_key, uint32_t k, uint32_t N,
                 uint32_t base_log, uint32_t level_count, int32_t noise_log2,
                 uint64_t *bsk_out);
/* Multi-bit BSK: n/g groups x 2^g GGSW, GGSW s of a group encrypts
 * prod_u (key bit u XOR inverted selector bit)
 * (lwe_multi_bit_bootstrap_key_generation.rs:21-78,504-529).                */
void orc_gen_multi_bit_bsk(orc_rng *rng, const uint64_t *lwe_key, uint32_t n,
                           const uint64_t *glwe_key, uint32_t k, uint32_t N,
                           uint32_t base_log, uint32_t level_count,
                           uint32_t grouping_factor, int32_t noise_log2,
                           uint64_t *bsk_out);
/* KSK layout [i<dim_in][slot j<l_ks][dim_out+1], slot j = level l_ks - j
 * (lwe_keyswitch_key_generation.rs:168-188).                                */
void orc_gen_ksk(orc_rng *rng, const uint64_t *key_in, uint32_t dim_in,
                 const uint64_t *key_out, uint32_t dim_out, uint32_t base_log,
                 uint32_t level_count, int32_t noise_log2, uint64_t *ksk_out);

/* -------------------------------------------------------- keyswitch ------ */
/* lwe_keyswitch.rs:137-232                                                   */
void orc_keyswitch(const uint64_t *ksk, uint32_t dim_in, uint32_t dim_out,
                   uint32_t base_log, uint32_t level_count,
                   const uint64_t *ct_in, uint64_t *ct_out);
void orc_keyswitch_batch(const uint64_t *ksk, uint32_t dim_in,
                         uint32_t dim_out, uint32_t base_log,
                         uint32_t level_count, const uint64_t *cts_in,
                         uint64_t *cts_out, uint32_t count,
                         uint32_t num_threads);

/* ------------------------------------------------------ modulus switch --- */
/* ms_out[0..n) = switched mask, ms_out[n] = switched body.  centered != 0
 * selects the centered-mean variant (modulus_switch.rs:35-100).             */
void orc_lwe_modulus_switch(const uint64_t *ct, uint32_t n,
                            uint32_t log_modulus, int centered,
                            uint32_t *ms_out);
/* modulus_switch.rs:55-100 body correction alone */
uint64_t orc_centered_ms_body_correction(const uint64_t *ct, uint32_t n,
                                         uint32_t log_modulus);

/* ------------------------------------------------------------- LUT ------- */
/* lwe_programmable_bootstrapping/mod.rs:26-83.  f_values[i] = f(i), i < p.   */
void orc_make_lut(const uint64_t *f_values, uint32_t p, uint64_t delta,
                  uint32_t k, uint32_t N, uint64_t *glwe_out);

/* glwe_sample_extraction.rs:119-165                                          */
void orc_sample_extract(const uint64_t *glwe, uint32_t k, uint32_t N,
                        uint32_t nth, uint64_t *lwe_out);

/* ------------------------------------------------------------- FFT ------- */
typedef struct orc_fft_plan orc_fft_plan;
orc_fft_plan *orc_fft_plan_new(uint32_t N);
void orc_fft_plan_free(orc_fft_plan *plan);

/* fft/mod.rs:224-265 (forward_as_integer), natural frequency order:
 * X[k] = sum_j (p[j] + i p[j+N/2]) e^{i pi j / N} e^{-2 pi i jk/(N/2)}.
 * out_re/out_im have N/2 entries.                                            */
void orc_fft_forward_integer(const orc_fft_plan *plan, const int64_t *poly,
                             double *out_re, double *out_im);
/* same transform of real (double) inputs -- used for the golden KAT          */
void orc_fft_forward_real(const orc_fft_plan *plan, const double *poly,
                          double *out_re, double *out_im);
/* fft/mod.rs:201-222 (forward_as_torus): inputs scaled by 2^-64              */
void orc_fft_forward_torus(const orc_fft_plan *plan, const uint64_t *poly,
                           double *out_re, double *out_im);
/* fft/mod.rs:289-330 (add_backward_as_torus) + torus/mod.rs:75-81            */
void orc_fft_add_backward_torus(const orc_fft_plan *plan, const double *in_re,
                                const double *in_im, uint64_t *poly_inout);

/* Fourier BSK in the oracle's private (bit-reversed) order, separate re/im
 * planes: `count` polynomials of N u64 -> count * N/2 re and im values
 * (lwe_bootstrap_key_conversion.rs:97).                                      */
void orc_bsk_to_fourier(const orc_fft_plan *plan, const uint64_t *bsk,
                        size_t poly_count, double *out_re, double *out_im);

/* -------------------------------------------------------- blind rotate --- */
/* bootstrap.rs:294-380 with ggsw.rs:483-602.  acc = trivial LUT GLWE,
 * modified in place.  ms = n mask values then the body value.               */
void orc_blind_rotate_fft(const orc_fft_plan *plan, uint64_t *acc,
                          const uint32_t *ms, const double *bsk_re,
                          const double *bsk_im, uint32_t n, uint32_t k,
                          uint32_t N, uint32_t base_log, uint32_t level_count);
/* exact-integer twin (karatsuba_pbs.rs:199-413) on the standard BSK          */
void orc_blind_rotate_exact(uint64_t *acc, const uint32_t *ms,
                            const uint64_t *bsk, uint32_t n, uint32_t k,
                            uint32_t N, uint32_t base_log,
                            uint32_t level_count);
/* one external product acc += ggsw (x) glwe_in, both modes (test entry)      */
void orc_add_external_product_fft(const orc_fft_plan *plan, uint64_t *acc,
                                  const double *ggsw_re, const double *ggsw_im,
                                  const uint64_t *glwe_in, uint32_t k,
                                  uint32_t N, uint32_t base_log,
                                  uint32_t level_count);
void orc_add_external_product_exact(uint64_t *acc, const uint64_t *ggsw,
                                    const uint64_t *glwe_in, uint32_t k,
                                    uint32_t N, uint32_t base_log,
                                    uint32_t level_count);

/* multi-bit (lwe_multi_bit_programmable_bootstrapping.rs:30-65,116-156,
 * 647-860): standard modulus switch only.                                   */
void orc_multi_bit_blind_rotate_fft(const orc_fft_plan *plan, uint64_t *acc,
                                    const uint64_t *ct_in,
                                    const double *bsk_re, const double *bsk_im,
                                    uint32_t n, uint32_t k, uint32_t N,
                                    uint32_t base_log, uint32_t level_count,
                                    uint32_t grouping_factor);
void orc_multi_bit_blind_rotate_exact(uint64_t *acc, const uint64_t *ct_in,
                                      const uint64_t *bsk, uint32_t n,
                                      uint32_t k, uint32_t N,
                                      uint32_t base_log, uint32_t level_count,
                                      uint32_t grouping_factor);

/* ------------------------------------------------------------- PBS ------- */
/* fft64_pbs.rs:924-1137 / bootstrap.rs:480-520: copy LUT, mod switch, blind
 * rotate, sample extract.  Batched with the C-ABI's index conventions:
 * sample s reads cts_in[in_idx[s]], LUT lut_idx[s], writes
 * cts_out[out_idx[s]] (+ j*count*(kN+1) for the j-th of num_many_lut outputs,
 * extracted at coefficient j*lut_stride).  Index arrays may be NULL
 * (= trivial).  exact != 0 uses the integer multiplier and `bsk_std`;
 * otherwise the Fourier planes.  grouping_factor > 1 selects multi-bit.     */
typedef struct {
  uint32_t n, k, N, base_log, level_count;
  uint32_t grouping_factor; /* 0 or 1: classic */
  int centered_ms;          /* classic only */
  uint32_t num_many_lut, lut_stride;
} orc_pbs_params;

void orc_pbs_batch(const orc_fft_plan *plan, const orc_pbs_params *prm,
                   const uint64_t *bsk_std, const double *bsk_re,
                   const double *bsk_im, const uint64_t *luts,
                   const uint64_t *lut_idx, const uint64_t *cts_in,
                   const uint64_t *in_idx, uint64_t *cts_out,
                   const uint64_t *out_idx, uint32_t count, int exact,
                   uint32_t num_threads);

uint32_t orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
