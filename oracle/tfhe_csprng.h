/* tfhe_csprng.h -- restatement of the reference's deterministic CSPRNG
 * (crate `tfhe-csprng`, in-tree at /root/reference/tfhe-csprng) and of the
 * key-generation / encryption draw order of tfhe-rs `core_crypto`, so that the
 * oracle can REGENERATE the keys and inputs behind the reference's committed
 * GPU golden outputs (tfhe/src/core_crypto/gpu/algorithms/test/pbs_golden).
 *
 * TEST INFRASTRUCTURE ONLY (same rule as pbs_oracle.h): used by tests/ to pin
 * parity, never by the product path.
 *
 * What is restated
 *  - the byte table: byte p of the stream of a generator with AES-128 key K is
 *    AES_K(le128(p >> 4))[p & 15]            (aes_ctr/generic.rs:178-193,
 *    aes_ctr/states.rs:40-50, implem/soft/block_cipher.rs:24-31,70-78);
 *    key = the 128-bit seed, little-endian    (aes_ctr/generic.rs:96-105);
 *    a fresh generator starts at byte 0       (aes_ctr/mod.rs:222-229);
 *  - forking: child i of (n children, b bytes each) owns bytes
 *    [p0 + i*b, p0 + (i+1)*b), the parent resumes at p0 + n*b
 *                                             (aes_ctr/generic.rs:142-176);
 *  - scalar sampling: uniform u64/u128 = little-endian bytes
 *    (math/random/uniform.rs:11-20); uniform binary = one byte & 1
 *    (uniform_binary.rs:11-20); TUniform(b) = ceil((b+2)/8) bytes, masked to
 *    b+2 bits, (v >> 1) + (v & 1) - 2^b       (t_uniform.rs:63-82).
 */
#ifndef TFHE_CSPRNG_ORACLE_H
#define TFHE_CSPRNG_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint8_t round_keys[11][16];
  uint64_t pos; /* byte position in the table (2^64 bytes is plenty here) */
} csprng_gen;

/* AES-128 single block (FIPS-197), exposed for the known-answer tests */
void csprng_aes128_encrypt_block(const uint8_t key[16], const uint8_t in[16],
                                 uint8_t out[16]);
/* 1 when the AES-NI path is compiled in and usable on this CPU */
int csprng_uses_aesni(void);

/* generator seeded with Seed(u128) = (seed_hi << 64) | seed_lo */
void csprng_init(csprng_gen *g, uint64_t seed_lo, uint64_t seed_hi);
void csprng_fill_bytes(csprng_gen *g, uint8_t *out, size_t count);
/* child generator starting at absolute byte `pos` (fork semantics) */
void csprng_at(const csprng_gen *parent, uint64_t pos, csprng_gen *child);

uint64_t csprng_uniform_u64(csprng_gen *g);
void csprng_fill_uniform_u64(csprng_gen *g, uint64_t *out, size_t count);
void csprng_fill_binary_u64(csprng_gen *g, uint64_t *out, size_t count);
int64_t csprng_tuniform(csprng_gen *g, uint32_t bound_log2);
uint32_t csprng_tuniform_bytes(uint32_t bound_log2);

/* The three generators of a tfhe-rs test / golden run
 * (pbs_golden/mod.rs:132-147, algorithms/test/mod.rs:36-47):
 *   seeder = Generator(seed); mask = Generator(seeder.u128());
 *   noise = Generator(seeder.u128()); secret = Generator(seeder.u128()). */
typedef struct {
  csprng_gen mask, noise, secret;
} csprng_resources;
void csprng_resources_init(csprng_resources *r, uint64_t seed_lo,
                           uint64_t seed_hi);

/* allocate_and_generate_new_binary_{lwe,glwe}_secret_key */
void csprng_gen_binary_key(csprng_resources *r, uint64_t *key, size_t count);

/* par_generate_lwe_bootstrap_key (lwe_bootstrap_key_generation.rs:250-314):
 * standard-domain BSK in the reference's container order
 * [n][level (l..1)][row k+1][poly k+1][N]. */
void csprng_gen_bsk(csprng_resources *r, const uint64_t *lwe_key, uint32_t n,
                    const uint64_t *glwe_key, uint32_t k, uint32_t N,
                    uint32_t base_log, uint32_t level_count,
                    uint32_t noise_bound_log2, uint64_t *bsk_out);

/* par_generate_lwe_multi_bit_bootstrap_key
 * (lwe_multi_bit_bootstrap_key_generation.rs:430-530):
 * [n/g][2^g ggsw][level][row][poly][N]. */
void csprng_gen_multi_bit_bsk(csprng_resources *r, const uint64_t *lwe_key,
                              uint32_t n, const uint64_t *glwe_key, uint32_t k,
                              uint32_t N, uint32_t base_log,
                              uint32_t level_count, uint32_t grouping_factor,
                              uint32_t noise_bound_log2, uint64_t *bsk_out);

/* encrypt_lwe_ciphertext (lwe_encryption.rs): n mask words then one noise
 * sample; ct_out has n+1 words. */
void csprng_lwe_encrypt(csprng_resources *r, const uint64_t *key, uint32_t n,
                        uint64_t plaintext, uint32_t noise_bound_log2,
                        uint64_t *ct_out);

#ifdef __cplusplus
}
#endif
#endif
