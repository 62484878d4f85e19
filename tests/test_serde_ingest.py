"""tfhe-rs_b200/serde_ingest.py: the serde / bincode wire format of the keys the GPU path ingests.
UNPINNED against real tfhe-rs output (the checkout has no serialized key and no Rust toolchain is available):
these tests pin the layout to the byte recipe written in the module docstring and check round trips."""
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def S():
    from tfhe_rs_b200 import serde_ingest

    return serde_ingest


def _u64s(*v):
    return b"".join(struct.pack("<Q", x) for x in v)


def test_plain_layout_of_a_tiny_bootstrap_key_byte_for_byte(S):
    # n = 1, glwe_size = 2, N = 2, l = 1: 1 * 1 * 2 * 2 * 2 = 8 words
    data = np.arange(1, 9, dtype=np.uint64)
    key = S.LweBootstrapKey(data, glwe_size=2, polynomial_size=2, decomp_base_log=23, decomp_level_count=1)
    want = (_u64s(8) + _u64s(*range(1, 9))  # Vec<u64>: length, elements
            + _u64s(2, 2, 23, 1)            # GlweSize, PolynomialSize, DecompositionBaseLog, DecompositionLevelCount
            + _u64s(0, 0) + _u64s(64))      # SerializableCiphertextModulus { modulus: u128 = 0 (native), scalar_bits }
    assert S.write_lwe_bootstrap_key(key) == want
    back = S.read_lwe_bootstrap_key(want)
    assert np.array_equal(back.data, data) and back.input_lwe_dimension == 1 and back.glwe_dimension == 1


def test_versioned_layout_adds_the_dispatch_variant_indexes(S):
    data = np.arange(8, dtype=np.uint64)
    key = S.LweBootstrapKey(data, 2, 2, 15, 1)
    tag = lambda v: struct.pack("<I", v)
    want = (tag(1) + tag(1)                       # LweBootstrapKeyVersions::V1, GgswCiphertextListVersions::V1
            + _u64s(8) + data.tobytes()           # Vec<u64> is not wrapped
            + tag(0) + _u64s(2) + tag(0) + _u64s(2) + tag(0) + _u64s(15) + tag(0) + _u64s(1)
            + tag(0) + _u64s(0, 0) + _u64s(64))   # SerializableCiphertextModulusVersions::V0
    assert S.write_lwe_bootstrap_key(key, versioned=True) == want
    assert np.array_equal(S.read_lwe_bootstrap_key(want, versioned=True).data, data)
    with pytest.raises(ValueError):               # a plain reader must not swallow the versioned form
        S.read_lwe_bootstrap_key(want, versioned=False)


@pytest.mark.parametrize("versioned", [False, True])
def test_round_trips(S, versioned):
    rng = np.random.default_rng(1)
    bsk = S.LweBootstrapKey(rng.integers(0, 1 << 64, size=5 * 2 * 3 * 3 * 256, dtype=np.uint64), 3, 256, 12, 2)
    b = S.read_lwe_bootstrap_key(S.write_lwe_bootstrap_key(bsk, versioned), versioned)
    assert np.array_equal(b.data, bsk.data) and (b.glwe_size, b.polynomial_size, b.decomp_base_log,
                                                 b.decomp_level_count, b.input_lwe_dimension) == (3, 256, 12, 2, 5)
    for cs in (S.CompressionSeed("ctr", seed=(1 << 127) + 12345, aes_index=7, byte_index=8),
               S.CompressionSeed("xof", xof_data=b"TFHE_GEN" + bytes(range(16)))):
        sk = S.SeededLweBootstrapKey(rng.integers(0, 1 << 64, size=5 * 2 * 3 * 256, dtype=np.uint64), 3, 256, 12, 2, cs)
        k = S.read_seeded_lwe_bootstrap_key(S.write_seeded_lwe_bootstrap_key(sk, versioned), versioned)
        assert np.array_equal(k.data, sk.data) and k.compression_seed == cs and k.input_lwe_dimension == 5
    ksk = S.LweKeyswitchKey(rng.integers(0, 1 << 64, size=512 * 4 * 25, dtype=np.uint64), 4, 4, 25)
    k = S.read_lwe_keyswitch_key(S.write_lwe_keyswitch_key(ksk, versioned), versioned)
    assert np.array_equal(k.data, ksk.data) and (k.input_lwe_dimension, k.output_lwe_dimension) == (512, 24)


def test_malformed_inputs_are_rejected(S):
    key = S.LweBootstrapKey(np.zeros(8, dtype=np.uint64), 2, 2, 23, 1)
    good = S.write_lwe_bootstrap_key(key)
    with pytest.raises(ValueError):
        S.read_lwe_bootstrap_key(good[:-3])                      # truncated
    with pytest.raises(ValueError):
        S.read_lwe_bootstrap_key(good + b"\0")                   # trailing bytes
    with pytest.raises(ValueError):
        S.read_lwe_bootstrap_key(good[:-8] + _u64s(32))          # a u32-torus key
    with pytest.raises(ValueError):
        S.read_lwe_bootstrap_key(good[:-24] + _u64s(1 << 32, 0) + _u64s(64))  # non-native modulus
    with pytest.raises(AssertionError):
        S.read_lwe_bootstrap_key(S.write_lwe_bootstrap_key(S.LweBootstrapKey(np.zeros(9, dtype=np.uint64), 2, 2, 23, 1)))


@pytest.mark.gpu
def test_serialized_keys_bootstrap_on_the_gpu(oracle, keyset):
    """bytes -> serde_ingest -> C ABI: a KS -> PBS with keys that only ever existed as serialized buffers decrypts
    like the oracle; the seeded form gives the same device key as the decompressed one."""
    import torch

    assert torch.cuda.is_available()
    from tfhe_rs_b200 import gpu, serde_ingest as S

    P = oracle.TOY_K1
    keys = keyset(P)
    streams = gpu.CudaStreams.new_single_gpu(0)
    for versioned in (False, True):
        bsk_bytes = S.write_lwe_bootstrap_key(S.LweBootstrapKey(keys.bsk, P.k + 1, P.N, P.pbs_base_log, P.pbs_level),
                                              versioned)
        ksk_bytes = S.write_lwe_keyswitch_key(S.LweKeyswitchKey(keys.ksk, P.ks_base_log, P.ks_level, P.n + 1), versioned)
        d_bsk = S.bootstrap_key_to_gpu(bsk_bytes, streams, "Centered" if P.centered_ms else None, versioned)
        d_ksk = S.keyswitch_key_to_gpu(ksk_bytes, streams, versioned)
        assert (d_bsk.input_lwe_dimension, d_ksk.input_key_lwe_dimension, d_ksk.output_key_lwe_dimension) == \
               (P.n, P.big_n, P.n)
        msgs = np.arange(8) % P.p
        big = oracle.lwe_encrypt_batch(oracle.Rng(3), keys.glwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta),
                                       P.lwe_noise_log2)
        f = [(5 * i + 2) % P.p for i in range(P.p)]
        lut = oracle.make_lut(P, f)
        d_big = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(big, streams)
        d_small = gpu.CudaLweCiphertextList.new(P.n, len(msgs), streams)
        idx = gpu.trivial_indexes(len(msgs), streams)
        gpu.cuda_keyswitch_lwe_ciphertext(d_ksk, d_big, d_small, idx, idx, True, streams)
        d_out = gpu.CudaLweCiphertextList.new(P.big_n, len(msgs), streams)
        d_lut = gpu.CudaGlweCiphertextList.from_glwe_ciphertext_list(lut, P.k, P.N, streams)
        lut_idx = gpu.CudaVec.new(len(msgs), streams)  # zeros: one shared LUT
        gpu.cuda_programmable_bootstrap_lwe_ciphertext(d_small, d_out, d_lut, lut_idx, idx, idx, d_bsk, streams)
        streams.synchronize()
        dec = oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, d_out.to_lwe_ciphertext_list(streams)), P.delta, P.p)
        assert np.array_equal(dec, np.array([f[m] for m in msgs]))
