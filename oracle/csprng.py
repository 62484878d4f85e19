"""ctypes binding of oracle/tfhe_csprng.c: the reference's deterministic CSPRNG
and key-generation draw order, used to regenerate the keys / inputs behind the
reference's committed golden PBS outputs.

TEST INFRASTRUCTURE ONLY (same rule as oracle.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as _o


class _Gen(C.Structure):
    _fields_ = [("round_keys", (C.c_uint8 * 16) * 11), ("pos", C.c_uint64)]


class _Resources(C.Structure):
    _fields_ = [("mask", _Gen), ("noise", _Gen), ("secret", _Gen)]


_declared = False


def _lib():
    global _declared
    l = _o.lib()
    if not _declared:
        u32, u64 = C.c_uint32, C.c_uint64
        u8p = C.POINTER(C.c_uint8)
        gp, rp = C.POINTER(_Gen), C.POINTER(_Resources)
        sig = {
            "csprng_aes128_encrypt_block": (None, [u8p, u8p, u8p]),
            "csprng_uses_aesni": (C.c_int, []),
            "csprng_init": (None, [gp, u64, u64]),
            "csprng_fill_bytes": (None, [gp, u8p, C.c_size_t]),
            "csprng_at": (None, [gp, u64, gp]),
            "csprng_uniform_u64": (u64, [gp]),
            "csprng_tuniform": (C.c_int64, [gp, u32]),
            "csprng_resources_init": (None, [rp, u64, u64]),
            "csprng_gen_binary_key": (None, [rp, _o._u64p, C.c_size_t]),
            "csprng_gen_bsk": (None, [rp, _o._u64p, u32, _o._u64p, u32, u32, u32, u32, u32, _o._u64p]),
            "csprng_gen_multi_bit_bsk": (None, [rp, _o._u64p, u32, _o._u64p, u32, u32, u32, u32, u32, u32, _o._u64p]),
            "csprng_lwe_encrypt": (None, [rp, _o._u64p, u32, u64, u32, _o._u64p]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _declared = True
    return l


def aes128_encrypt_block(key: bytes, block: bytes) -> bytes:
    k = (C.c_uint8 * 16)(*key)
    i = (C.c_uint8 * 16)(*block)
    o = (C.c_uint8 * 16)()
    _lib().csprng_aes128_encrypt_block(k, i, o)
    return bytes(o)


def uses_aesni() -> bool:
    return bool(_lib().csprng_uses_aesni())


class Generator:
    """AesCtrGenerator seeded with Seed(u128) (tfhe-csprng aes_ctr/generic.rs)."""

    def __init__(self, seed: int):
        self._g = _Gen()
        _lib().csprng_init(C.byref(self._g), seed & (2**64 - 1), seed >> 64)

    def bytes(self, count: int) -> bytes:
        buf = (C.c_uint8 * count)()
        _lib().csprng_fill_bytes(C.byref(self._g), buf, count)
        return bytes(buf)

    def seek(self, pos: int):
        self._g.pos = pos

    def uniform_u64(self) -> int:
        return int(_lib().csprng_uniform_u64(C.byref(self._g)))

    def tuniform(self, bound_log2: int) -> int:
        return int(_lib().csprng_tuniform(C.byref(self._g), bound_log2))


class Resources:
    """DeterministicSeeder -> (mask, noise, secret) generators
    (pbs_golden/mod.rs:132-147)."""

    def __init__(self, seed: int):
        self._r = _Resources()
        _lib().csprng_resources_init(C.byref(self._r), seed & (2**64 - 1), seed >> 64)
        # the seeds the DeterministicSeeder handed out (mask, noise, secret)
        seeder = Generator(seed)
        raw = seeder.bytes(48)
        self.mask_seed, self.noise_seed, self.secret_seed = (
            int.from_bytes(raw[16 * i: 16 * i + 16], "little") for i in range(3))

    @property
    def mask_position(self) -> int:
        """byte position of the mask generator in its table"""
        return int(self._r.mask.pos)

    def binary_key(self, count: int) -> np.ndarray:
        out = np.zeros(count, dtype=np.uint64)
        _lib().csprng_gen_binary_key(C.byref(self._r), _o.u64p(out), count)
        return out

    def bsk(self, lwe_key, glwe_key, p: "_o.Params", noise_bound_log2: int) -> np.ndarray:
        out = np.zeros(p.n * p.pbs_level * (p.k + 1) ** 2 * p.N, dtype=np.uint64)
        _lib().csprng_gen_bsk(C.byref(self._r), _o.u64p(lwe_key), p.n, _o.u64p(glwe_key), p.k, p.N,
                              p.pbs_base_log, p.pbs_level, noise_bound_log2, _o.u64p(out))
        return out

    def multi_bit_bsk(self, lwe_key, glwe_key, p: "_o.Params", noise_bound_log2: int) -> np.ndarray:
        g = p.grouping_factor
        out = np.zeros((p.n // g) * (1 << g) * p.pbs_level * (p.k + 1) ** 2 * p.N, dtype=np.uint64)
        _lib().csprng_gen_multi_bit_bsk(C.byref(self._r), _o.u64p(lwe_key), p.n, _o.u64p(glwe_key), p.k, p.N,
                                        p.pbs_base_log, p.pbs_level, g, noise_bound_log2, _o.u64p(out))
        return out

    def lwe_encrypt(self, key, plaintext: int, noise_bound_log2: int) -> np.ndarray:
        n = len(key)
        out = np.zeros(n + 1, dtype=np.uint64)
        _lib().csprng_lwe_encrypt(C.byref(self._r), _o.u64p(key), n, plaintext, noise_bound_log2, _o.u64p(out))
        return out


# tfhe/src/shortint/parameters/v1_1/multi_bit/tuniform/p_fail_2_minus_128/ks_pbs_gpu.rs:205-225
PARAM_GPU_MULTI_BIT_GROUP_4_MESSAGE_2_CARRY_2_KS_PBS = _o.Params(
    "PARAM_GPU_MULTI_BIT_GROUP_4_MESSAGE_2_CARRY_2_KS_PBS_TUNIFORM_2M128",
    n=920, k=1, N=2048, pbs_base_log=22, pbs_level=1, ks_base_log=3, ks_level=5,
    lwe_noise_log2=45, glwe_noise_log2=17, grouping_factor=4, centered_ms=False,
)

GOLDEN_SEED = 0x0D1C_E5ED_9017_2048  # pbs_golden/mod.rs:83
GOLDEN_MESSAGES = (1, 7, 15)  # pbs_golden/mod.rs:103


def golden_lut(params: "_o.Params") -> np.ndarray:
    """f(x) = (2x - 1) mod p, pbs_golden/mod.rs:236."""
    return _o.make_lut(params, [(2 * x - 1) % params.p for x in range(params.p)])


def golden_keyset(params: "_o.Params", seed: int = GOLDEN_SEED, messages=GOLDEN_MESSAGES):
    """Replay run_{classical,multi_bit}_pbs_golden_batch (pbs_golden/mod.rs:215-285,
    331-395) up to the bootstrap call: secret keys, BSK, then the inputs, all
    from the three deterministic generators.  Returns (KeySet, inputs[len(messages)][n+1])."""
    r = Resources(seed)
    lwe_sk = r.binary_key(params.n)
    glwe_sk = r.binary_key(params.k * params.N)
    if params.grouping_factor > 1:
        bsk = r.multi_bit_bsk(lwe_sk, glwe_sk, params, params.glwe_noise_log2)
    else:
        bsk = r.bsk(lwe_sk, glwe_sk, params, params.glwe_noise_log2)
    inputs = np.stack([r.lwe_encrypt(lwe_sk, (m * params.delta) % 2**64, params.lwe_noise_log2) for m in messages])
    return _o.KeySet(params, lwe_sk, glwe_sk, bsk, None), inputs
