/*
 * tfhe_b200.h -- C ABI of libtfhe_cuda_backend_b200.so
 *
 * Drop-in boundary: every symbol below has the NAME and SIGNATURE of the
 * symbol of the same name exported by tfhe-rs' `backends/tfhe-cuda-backend`
 * (static lib libtfhe_cuda_backend.a, bound from Rust by bindgen in
 * backends/tfhe-cuda-backend/src/bindings.rs).  The citation after each
 * declaration is the reference header that declares it, relative to
 * /root/reference/backends/.  All `void *` ciphertext / key / index
 * arguments are DEVICE pointers unless stated; `stream` is a cudaStream_t.
 *
 * Conventions kept from the reference:
 *   - no return codes: any CUDA error or violated precondition prints to
 *     stderr and abort()s;
 *   - every call is asynchronous on `stream` and first selects `gpu_index`;
 *   - the caller owns all ciphertext / key / index buffers, the callee owns
 *     only the scratch object between scratch_* and cleanup_*;
 *   - the Fourier bootstrap key layout inside `dest` is private to the
 *     engine but fits the caller-allocated n*(k+1)^2*l*N f64 words.
 *
 * Symbols prefixed b200_ are additions (test / measurement entry points and
 * the multi-GPU key broadcast); the reference has no equivalent.
 */
#ifndef TFHE_B200_H
#define TFHE_B200_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* tfhe-cuda-backend/cuda/include/pbs/pbs_enums.h:4-6 */
typedef enum { MULTI_BIT = 0, CLASSICAL = 1 } PBS_TYPE;
typedef enum { DEFAULT = 0, CG = 1, TBC = 2 } PBS_VARIANT;
typedef enum { NO_REDUCTION = 0, CENTERED = 1 } PBS_MS_REDUCTION_T;

/* ---- device helpers: tfhe-cuda-common/cuda/include/device.h:59-88 ------- */
void *cuda_create_stream_ffi(uint32_t gpu_index);
void cuda_destroy_stream(void *stream, uint32_t gpu_index);
void cuda_synchronize_stream(void *stream, uint32_t gpu_index);
uint32_t cuda_is_available(void);
void *cuda_malloc(uint64_t size, uint32_t gpu_index);
void *cuda_malloc_async(uint64_t size, void *stream, uint32_t gpu_index);
bool cuda_check_valid_malloc(uint64_t size, uint32_t gpu_index);
uint64_t cuda_device_total_memory(uint32_t gpu_index);
void cuda_memcpy_async_to_gpu(void *dest, const void *src, uint64_t size,
                              void *stream, uint32_t gpu_index);
void cuda_memcpy_async_gpu_to_gpu(void *dest, void const *src, uint64_t size,
                                  void *stream, uint32_t gpu_index);
void cuda_memcpy_gpu_to_gpu(void *dest, void const *src, uint64_t size,
                            uint32_t gpu_index);
void cuda_memcpy_async_to_cpu(void *dest, const void *src, uint64_t size,
                              void *stream, uint32_t gpu_index);
void cuda_memset_async(void *dest, uint64_t val, uint64_t size, void *stream,
                       uint32_t gpu_index);
int cuda_get_number_of_gpus(void);
int cuda_get_number_of_sms(void);
void cuda_synchronize_device(uint32_t gpu_index);
void cuda_drop(void *ptr, uint32_t gpu_index);
void cuda_drop_async(void *ptr, void *stream, uint32_t gpu_index);
uint32_t cuda_get_max_shared_memory(uint32_t gpu_index);

/* ---- classic PBS: tfhe-cuda-backend/cuda/include/pbs/programmable_bootstrap.h */
/* :50-53  src = HOST pointer, standard-domain BSK [i][level][row][col][N] u64 */
void cuda_convert_lwe_programmable_bootstrap_key_64_async(
    void *stream, uint32_t gpu_index, void *dest, void const *src,
    uint32_t input_lwe_dim, uint32_t glwe_dim, uint32_t level_count,
    uint32_t polynomial_size);
/* :60-64  returns the scratch bytes; allocate_gpu_memory=false = size query */
uint64_t scratch_cuda_programmable_bootstrap_64_async(
    void *stream, uint32_t gpu_index, int8_t **buffer, uint32_t lwe_dimension,
    uint32_t glwe_dimension, uint32_t polynomial_size, uint32_t level_count,
    uint32_t input_lwe_ciphertext_count, bool allocate_gpu_memory,
    PBS_MS_REDUCTION_T noise_reduction_type);
/* :81-88 */
void cuda_programmable_bootstrap_64_async(
    void *stream, uint32_t gpu_index, void *lwe_array_out,
    void const *lwe_output_indexes, void const *lut_vector,
    void const *lut_vector_indexes, void const *lwe_array_in,
    void const *lwe_input_indexes, void const *bootstrapping_key,
    int8_t *buffer, uint32_t lwe_dimension, uint32_t glwe_dimension,
    uint32_t polynomial_size, uint32_t base_log, uint32_t level_count,
    uint32_t num_samples, uint32_t num_many_lut, uint32_t lut_stride);
/* :97-98 */
void cleanup_cuda_programmable_bootstrap_64(void *stream, uint32_t gpu_index,
                                            int8_t **pbs_buffer);

/* u32 torus, programmable_bootstrap.h:47-50,72-79: key words, ciphertexts,
 * accumulators and the three index vectors are u32.  The reference exposes no
 * scratch function for this variant; `buffer` is a scratch object made by
 * scratch_cuda_programmable_bootstrap_64_async with the same (k, N, l).  Words
 * are widened to the top half of a u64 inside the kernel and rounded back to 32
 * bits on output (the blind rotation itself is the 64-bit path). */
void cuda_convert_lwe_programmable_bootstrap_key_32_async(
    void *stream, uint32_t gpu_index, void *dest, void const *src,
    uint32_t input_lwe_dim, uint32_t glwe_dim, uint32_t level_count,
    uint32_t polynomial_size);
void cuda_programmable_bootstrap_lwe_ciphertext_vector_32_async(
    void *stream, uint32_t gpu_index, void *lwe_array_out,
    void const *lwe_output_indexes, void const *lut_vector,
    void const *lut_vector_indexes, void const *lwe_array_in,
    void const *lwe_input_indexes, void const *bootstrapping_key,
    int8_t *buffer, uint32_t lwe_dimension, uint32_t glwe_dimension,
    uint32_t polynomial_size, uint32_t base_log, uint32_t level_count,
    uint32_t num_samples, uint32_t num_many_lut, uint32_t lut_stride);

/* ---- multi-bit PBS: .../include/pbs/programmable_bootstrap_multibit.h:9-42 */
bool has_support_to_cuda_programmable_bootstrap_cg_multi_bit(
    uint32_t glwe_dimension, uint32_t polynomial_size, uint32_t level_count,
    uint32_t num_samples, uint32_t max_shared_memory);
void cuda_convert_lwe_multi_bit_programmable_bootstrap_key_64_async(
    void *stream, uint32_t gpu_index, void *dest, void const *src,
    uint32_t input_lwe_dim, uint32_t glwe_dim, uint32_t level_count,
    uint32_t polynomial_size, uint32_t grouping_factor);
uint64_t scratch_cuda_multi_bit_programmable_bootstrap_64_async(
    void *stream, uint32_t gpu_index, int8_t **pbs_buffer,
    uint32_t glwe_dimension, uint32_t polynomial_size, uint32_t level_count,
    uint32_t input_lwe_ciphertext_count, bool allocate_gpu_memory);
void cuda_multi_bit_programmable_bootstrap_64_async(
    void *stream, uint32_t gpu_index, void *lwe_array_out,
    void const *lwe_output_indexes, void const *lut_vector,
    void const *lut_vector_indexes, void const *lwe_array_in,
    void const *lwe_input_indexes, void const *bootstrapping_key,
    int8_t *buffer, uint32_t lwe_dimension, uint32_t glwe_dimension,
    uint32_t polynomial_size, uint32_t grouping_factor, uint32_t base_log,
    uint32_t level_count, uint32_t num_samples, uint32_t num_many_lut,
    uint32_t lut_stride);
void cleanup_cuda_multi_bit_programmable_bootstrap_64(void *stream,
                                                      uint32_t gpu_index,
                                                      int8_t **pbs_buffer);

/* ---- keyswitch: .../include/keyswitch/keyswitch.h:18-44 ----------------- */
void cuda_keyswitch_lwe_ciphertext_vector_64_64_async(
    void *stream, uint32_t gpu_index, void *lwe_array_out,
    void const *lwe_output_indexes, void const *lwe_array_in,
    void const *lwe_input_indexes, void const *ksk, uint32_t lwe_dimension_in,
    uint32_t lwe_dimension_out, uint32_t base_log, uint32_t level_count,
    uint32_t num_samples);
void cuda_keyswitch_gemm_64_64_async(
    void *stream, uint32_t gpu_index, void *lwe_array_out,
    void const *lwe_output_indexes, void const *lwe_array_in,
    void const *lwe_input_indexes, void const *ksk, uint32_t lwe_dimension_in,
    uint32_t lwe_dimension_out, uint32_t base_log, uint32_t level_count,
    uint32_t num_samples, bool uses_trivial_indexes);
/* KS32 variants, keyswitch.h:23-28,42-47: u64 input ciphertexts and indexes,
 * u32 keyswitch key (same [i][level slot][n_out+1] nesting), u32 output; body =
 * input body rounded to its top 32 bits (lwe_keyswitch.rs:331-455).
 * base_log * level_count <= 32. */
void cuda_keyswitch_lwe_ciphertext_vector_64_32_async(
    void *stream, uint32_t gpu_index, void *lwe_array_out,
    void const *lwe_output_indexes, void const *lwe_array_in,
    void const *lwe_input_indexes, void const *ksk, uint32_t lwe_dimension_in,
    uint32_t lwe_dimension_out, uint32_t base_log, uint32_t level_count,
    uint32_t num_samples);
void cuda_keyswitch_gemm_64_32_async(
    void *stream, uint32_t gpu_index, void *lwe_array_out,
    void const *lwe_output_indexes, void const *lwe_array_in,
    void const *lwe_input_indexes, void const *ksk, uint32_t lwe_dimension_in,
    uint32_t lwe_dimension_out, uint32_t base_log, uint32_t level_count,
    uint32_t num_samples, bool uses_trivial_indexes);

/* ---- stand-alone integer stages: .../include/ciphertext.h:15-32 ---------- */
/* LWE id = sample extraction of coefficient nth_array[id] % num_lwes_stored_per_glwe
 * of GLWE id / num_lwes_to_extract_per_glwe (nth_array on the device). */
void cuda_glwe_sample_extract_64_async(
    void *stream, uint32_t gpu_index, void *lwe_array_out,
    void const *glwe_array_in, uint32_t const *nth_array, uint32_t num_nths,
    uint32_t num_lwes_to_extract_per_glwe, uint32_t num_lwes_stored_per_glwe,
    uint32_t glwe_dimension, uint32_t polynomial_size);
/* x -> (x + 2^(63 - log_modulus)) >> (64 - log_modulus), element-wise */
void cuda_modulus_switch_inplace_64_async(void *stream, uint32_t gpu_index,
                                          void *lwe_array_out, uint32_t size,
                                          uint32_t log_modulus);
void cuda_modulus_switch_64_async(void *stream, uint32_t gpu_index,
                                  void *lwe_out, const void *lwe_in,
                                  uint32_t size, uint32_t log_modulus);
/* one LWE of lwe_dimension + 1 words: mask switched as above, body after the
 * centered-mean correction (algorithms/modulus_switch.rs:35-100) */
void cuda_centered_modulus_switch_64_async(void *stream, uint32_t gpu_index,
                                           void *lwe_out, const void *lwe_in,
                                           uint32_t lwe_dimension,
                                           uint32_t log_modulus);

/* ---- additions (no reference equivalent) -------------------------------- */
/* forward negacyclic transform of `total_polynomials` real polynomials given
 * as N/2 interleaved (re,im) = (p[j], p[j+N/2]) f64 pairs; output in NATURAL
 * frequency order with the reference's convention
 * X[k] = sum_j z_j e^{i pi j/N} e^{-2 pi i jk/(N/2)} (the contract of the
 * reference's test-only cuda_forward_fft16x4x16_async,
 * include/pbs/programmable_bootstrap.h:26-29).  polynomial_size must be 2048. */
void b200_forward_negacyclic_fft_async(void *stream, uint32_t gpu_index,
                                       void const *input, void *output,
                                       uint32_t polynomial_size,
                                       uint32_t total_polynomials);
/* Seeded (compressed) bootstrap key ingest: `seeded_bodies` = HOST pointer to
 * the body polynomials of a SeededLweBootstrapKey / SeededLweMultiBitBootstrapKey
 * ([ggsw][level][glwe row][N] u64, tfhe/src/core_crypto/entities/
 * seeded_lwe_bootstrap_key.rs); the masks are regenerated ON THE GPU from the
 * compression seed's AES-128 counter-mode table (tfhe-csprng generic.rs:178-193;
 * decompression order of seeded_lwe_bootstrap_key_decompression.rs:36-60) and
 * the key is converted exactly as cuda_convert_lwe_[multi_bit_]programmable_
 * bootstrap_key_64_async would convert the decompressed key.  `aes_key` = the
 * 16 key bytes (Seed(u128) little endian, or the XOF-derived key),
 * counter = first AES index + offset of the generator, first_byte_index in
 * {0, 8}.  grouping_factor 0/1 = classic key.  `dest` sized as for the
 * non-seeded conversion. */
void b200_convert_seeded_lwe_programmable_bootstrap_key_64_async(
    void *stream, uint32_t gpu_index, void *dest, void const *seeded_bodies,
    const uint8_t *aes_key, uint64_t counter_lo, uint64_t counter_hi,
    uint32_t first_byte_index, uint32_t input_lwe_dim, uint32_t glwe_dim,
    uint32_t level_count, uint32_t polynomial_size, uint32_t grouping_factor);
/* pin the keyswitch kernel: 0 automatic (default), 1 int8 tensor cores
 * (keyswitch_imma.cuh), 2 fp64 pipe, 3 integer pipe (keyswitch.cuh).  A path
 * whose exactness precondition does not hold for the given decomposition
 * falls through to the next one.  Also settable with B200_KS_PATH. */
void b200_set_keyswitch_path(int path);
/* pin the classic-PBS kernel of the (N = 2048, k = 1, l = 1) fast path for A/B
 * measurements and bit-identity tests: 0 automatic (default), 1 first kernel
 * (u64 accumulator), 3 round-1 register kernel, 5 lean rotate/decompose with
 * exchange 2 through shared memory, 4 / 6 other key-prefetch schedules, 7 TMA
 * key ring (one CTA per SM), 8 exchange 2 through tensor memory (tmem_x2.cuh),
 * 9 tensor-memory exchange + one-slot TMA key ring at two CTAs per SM (v6),
 * 14 the same with the switched mask staged in shared memory (one CTA per SM),
 * 15..19 v6 with other rotate/decompose instruction sequences, 20..22 v6
 * with the CTA barriers around the MAC replaced by arrive / wait pairs (19 and 20
 * are what the automatic dispatch launches for <= / > one CTA per SM),
 * 10 v6 with the register key prefetch, 11 v6 with the own key row in
 * registers and the other row through the ring, 12 v7 (pass-3 twiddles parked
 * in tensor memory, both key rows prefetched in registers, no ring), 13 v7
 * compiled for three CTAs per SM.
 * All variants compute the same function; 3..8 are bit-identical to each
 * other.  Also settable with B200_PBS_VARIANT. */
void b200_set_pbs_variant(int variant);
/* (N = 512, l = 1, k <= 4) register kernel (csrc/pbs_n512.cuh), e.g.
 * PARAM_MESSAGE_1_CARRY_1: 0 automatic (TMA key ring, one CTA per SM, 1 / 2 / 3
 * LWEs per CTA by launch size), 3 / 2 / 4 ring with one / two / three LWEs per
 * CTA, 1 register key ring with two CTAs per SM.  Same results in every mode.
 * Also B200_N512_MODE; B200_N512_GENERIC=1 keeps these shapes on the generic
 * kernel (read at library load: it also selects the key layout). */
void b200_set_n512_mode(int mode);
/* Which shapes run on their dedicated register kernels: bit 0 (N = 512, l = 1,
 * k <= 4; csrc/pbs_n512.cuh), bit 1 (N = 8192, k = 1, l = 2; csrc/pbs_n8192.cuh,
 * PARAM_MESSAGE_3_CARRY_3).  Default 3; a cleared bit keeps the shape on the
 * generic kernels (A/B measurements).  Bit 2 (value 4) selects the
 * first-generation N = 8192 kernel (per-thread key loads instead of the TMA
 * ring; same key layout, same results; B200_N8192_GEN1=1).  Bit 3 (value 8)
 * selects a debug instance of the default kernel for compute-sanitizer
 * racecheck (every thread arrives on the ring's `empty` barriers itself;
 * B200_N8192_RACECHECK=1).  The setting selects the device key
 * layout: convert a key and run its PBS under the same value.  Environment:
 * B200_N512_GENERIC=1 / B200_N8192_GENERIC=1 clear bit 0 / bit 1 at load. */
void b200_set_register_kernels(int mask);
/* DEVIATION FROM THE REFERENCE, switchable.  The multi-bit PBS kernels round
 * an exact tie of the bits dropped by the gadget decomposition to EVEN; the
 * reference (commons/math/decomposition/decomposer.rs:163-188) rounds it up.
 * Their accumulator is re-assigned from f64 every step (and is a 32-bit word
 * in the N = 2048 register kernels), so exact ties are common and always-up
 * biases every coefficient: measured output-noise variance 7.6x (g = 3) /
 * 1.10x (g = 4) of the reference formula with the reference rule against
 * 0.21x / 0.40x with the even rule (profiles/round1.md).  Outputs are valid
 * ciphertexts of the same plaintext either way.  reference_exact != 0 selects
 * the reference's rule bit for bit (also B200_MULTIBIT_TIES=reference).  The
 * same switch governs the classic (N = 8192, k = 1, l = 2) register kernel
 * (csrc/pbs_n8192.cuh): its 32-bit accumulator leaves two bits below the
 * gadget's 30, a tie is hit by one value in four, and always-up measured 10x
 * the oracle's output noise on PARAM_MESSAGE_3_CARRY_3 (profiles/round2.md,
 * section 8).  Every other classic PBS kernel and the keyswitch always use the
 * reference rule (their ties have probability <= 2^-9). */
void b200_set_multibit_tie_rule(int reference_exact);
/* multi-bit PBS (N = 2048, k = 1): launches of at most `max_samples` LWEs take
 * the low-latency path -- the per-sample key bundle of all n/g groups is built
 * by one GPU-wide kernel into a stream-ordered workspace, then one CTA per LWE
 * runs the n/g external products (the reference's keybundle + accumulate
 * split, programmable_bootstrap_multibit.cuh:30-430); larger launches use the
 * fused kernel that never materialises the bundle.  -1 = default (the SM
 * count), 0 = never.  Also settable with B200_MULTIBIT_LL_MAX. */
void b200_set_multibit_ll_max(int max_samples);
/* number of kernels this library has launched in the calling process */
uint64_t b200_kernel_launch_count(void);
/* 1 if (lwe_dim, glwe_dim, N, level_count) runs on the register-FFT kernel */
int b200_pbs_uses_fast_path(uint32_t lwe_dimension, uint32_t glwe_dimension,
                            uint32_t polynomial_size, uint32_t level_count);
const char *b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TFHE_B200_H */
