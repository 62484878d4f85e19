"""Radix multiplication (SURVEY 8 f1 / BASELINE configs[3]): the host-side
cascade of tfhe-rs_b200/integer.py, (a) on CPU with an oracle-backed engine
(host logic only), (b) on the GPU through the C ABI."""
import numpy as np
import pytest

from tfhe_rs_b200 import integer


class OracleEngine(integer.BlockEngine):
    """numpy blocks, LUT evaluation = oracle keyswitch + PBS (tests only)."""

    def __init__(self, oracle, keys):
        self.O, self.keys = oracle, keys
        P = keys.params
        self.luts = np.stack([oracle.make_lut(P, f) for f in integer.lut_functions()])
        self.big = P.big_n
        self.pbs_count = 0

    def zeros(self, count):
        return np.zeros((count, self.big + 1), dtype=np.uint64)

    def stack(self, rows):
        return np.stack(list(rows))

    def add(self, a, b):
        with np.errstate(over="ignore"):
            return a + b

    def scalar_mul(self, a, s):
        with np.errstate(over="ignore"):
            return a * np.uint64(s)

    def apply_luts(self, blocks, lut_ids):
        small = self.O.keyswitch_batch(self.keys, blocks)
        self.pbs_count += len(lut_ids)
        return self.O.pbs_batch(self.keys, self.luts, small, lut_idx=np.asarray(lut_ids, dtype=np.uint64))


def encrypt_radix(oracle, keys, rng, value, num_blocks):
    P = keys.params
    digits = [(value >> (2 * i)) & 3 for i in range(num_blocks)]
    return oracle.lwe_encrypt_batch(rng, keys.glwe_sk, np.array(digits, dtype=np.uint64) * np.uint64(P.delta),
                                    P.lwe_noise_log2)


def decrypt_radix(oracle, keys, blocks):
    P = keys.params
    d = oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, np.asarray(blocks)), P.delta, P.p)
    assert all(int(x) < 4 for x in d), d  # carries must be empty after full_propagate
    return sum(int(x) << (2 * i) for i, x in enumerate(d))


def test_lut_tables():
    lsb, msb, msg, carry = integer.lut_functions()[:4]
    for l in range(4):
        for r in range(4):
            assert lsb[4 * l + r] == (l * r) % 4 and msb[4 * l + r] == (l * r) // 4
    assert msg[13] == 1 and carry[13] == 3


def test_parallel_carry_propagation_matches_sequential(oracle, keyset):
    """Worst-case carry chains (0x3333 + 1 style) through both propagators."""
    from tfhe_rs_b200.integer import full_propagate, full_propagate_parallel

    P = oracle.Params("TOY_2_2_N512", n=32, k=1, N=512, pbs_base_log=23, pbs_level=1, ks_base_log=4, ks_level=4,
                      lwe_noise_log2=30, glwe_noise_log2=8)
    keys = keyset(P, seed=31)
    eng = OracleEngine(oracle, keys)
    rng = oracle.Rng(6)
    for digits in ([15, 3, 3, 3, 3, 3, 3, 3], [7, 12, 3, 15, 0, 3, 4, 9], [4, 3, 3, 0, 3, 3, 3, 15]):
        blocks = oracle.lwe_encrypt_batch(rng, keys.glwe_sk, np.array(digits, dtype=np.uint64) * np.uint64(P.delta),
                                          P.lwe_noise_log2)
        want = sum(d << (2 * i) for i, d in enumerate(digits)) % (1 << 16)
        assert decrypt_radix(oracle, keys, full_propagate(eng, blocks, 8)) == want
        assert decrypt_radix(oracle, keys, full_propagate_parallel(eng, blocks, 8)) == want


@pytest.mark.parametrize("num_blocks,a,b", [(4, 0xB7, 0x5D), (8, 0xFFFF, 0xFFFF), (8, 12345, 54321)])
def test_radix_mul_host_logic_on_oracle_engine(oracle, keyset, num_blocks, a, b):
    """integer/server_key/radix_parallel/tests_unsigned/test_mul.rs style:
    decrypt(mul(enc a, enc b)) == a*b mod 4^blocks; toy N=256 parameters keep
    the CPU cost low (the host cascade is parameter independent)."""
    P = oracle.Params("TOY_2_2_N512", n=32, k=1, N=512, pbs_base_log=23, pbs_level=1, ks_base_log=4, ks_level=4,
                      lwe_noise_log2=30, glwe_noise_log2=8)
    keys = keyset(P, seed=31)
    eng = OracleEngine(oracle, keys)
    rng = oracle.Rng(5)
    lhs, rhs = encrypt_radix(oracle, keys, rng, a, num_blocks), encrypt_radix(oracle, keys, rng, b, num_blocks)
    out = integer.unchecked_mul(eng, lhs, rhs)
    assert decrypt_radix(oracle, keys, out) == (a * b) % (1 << (2 * num_blocks))


@pytest.mark.gpu
def test_fheuint64_mul_gpu(oracle, keyset):
    """configs[3]: FheUint64 x FheUint64, 32 blocks, PARAM_MESSAGE_2_CARRY_2_KS_PBS,
    full KS+PBS cascade on one GPU through the C ABI."""
    import torch

    assert torch.cuda.is_available()
    from tfhe_rs_b200 import gpu, server_key

    P = oracle.PARAM_MESSAGE_2_CARRY_2_KS_PBS
    keys = keyset(P, seed=0xB2000001)
    streams = gpu.CudaStreams.new_single_gpu(0)
    skey = server_key.upload_server_key(keys.bsk, keys.ksk, n=P.n, k=P.k, N=P.N, pbs_base_log=P.pbs_base_log,
                                        pbs_level=P.pbs_level, ks_base_log=P.ks_base_log, ks_level=P.ks_level,
                                        centered_ms=True, streams=streams)
    rsk = integer.CudaRadixServerKey(skey, None, P.k, P.N)  # accumulators from the package's own generator
    rng = oracle.Rng(77)
    for a, b in [(0xDEADBEEFCAFEF00D, 0x123456789ABCDEF1), ((1 << 64) - 1, (1 << 64) - 1)]:
        lhs = integer.CudaUnsignedRadixCiphertext(rsk.engine.from_numpy(encrypt_radix(oracle, keys, rng, a, 32)))
        rhs = integer.CudaUnsignedRadixCiphertext(rsk.engine.from_numpy(encrypt_radix(oracle, keys, rng, b, 32)))
        out = rsk.unchecked_mul(lhs, rhs)
        got = decrypt_radix(oracle, keys, rsk.engine.to_numpy(out.blocks))
        assert got == (a * b) % (1 << 64)
    assert rsk.engine.pbs_count > 2000
