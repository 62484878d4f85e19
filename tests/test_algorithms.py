"""Host-side algorithms of the package (numpy only) against the oracle."""
import numpy as np

from tfhe_rs_b200 import algorithms as A


def test_lut_generator_matches_oracle(oracle):
    for P, f in ((oracle.PARAM_MESSAGE_2_CARRY_2_KS_PBS, [(7 * i + 2) % 16 for i in range(16)]),
                 (oracle.TOY_K2_L2, [(3 * i) % 16 for i in range(16)])):
        got = A.generate_programmable_bootstrap_glwe_lut(P.N, P.k + 1, 16, P.delta, f)
        assert np.array_equal(got, oracle.make_lut(P, f))
        got2 = A.generate_programmable_bootstrap_glwe_lut(P.N, P.k + 1, 16, P.delta, lambda x: f[x])
        assert np.array_equal(got2, got)


def test_many_lut_accumulator_semantics(oracle, keyset):
    """Two functions of a 3-bit message in one accumulator, extracted with
    num_many_lut = 2 and the stride the generator returns
    (programmable_bootstrap_classic.cuh:491-495, 698-735)."""
    P = oracle.TOY_K1
    keys = keyset(P, seed=3, with_ksk=False)
    f0 = lambda x: (x * x) % 16
    f1 = lambda x: (15 - x) % 16
    acc, stride = A.generate_many_lut_accumulator(P.N, P.k + 1, 16, P.delta, [f0, f1])
    msgs = np.arange(8, dtype=np.uint64)
    cts = oracle.lwe_encrypt_batch(oracle.Rng(4), keys.lwe_sk, msgs * np.uint64(P.delta), P.lwe_noise_log2)
    out = oracle.pbs_batch(keys, acc, cts, num_many_lut=2, lut_stride=stride)
    dec = oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, out), P.delta, 16).reshape(2, -1)
    assert list(dec[0]) == [f0(int(m)) for m in msgs]
    assert list(dec[1]) == [f1(int(m)) for m in msgs]
