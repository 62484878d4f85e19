"""Key ingest from tfhe-rs's serde wire format.

tfhe-rs users hold their evaluation keys as bytes: `bincode::serialize(&key)`
(plain serde) or `bincode::serialize(&key.versionize())` (tfhe-versionable, what
`safe_serialize` wraps).  This module reads and writes both encodings for the
three standard-domain keys the GPU path ingests -- the same entities
`core_crypto::gpu` converts from:

  LweBootstrapKeyOwned<u64>        tfhe/src/core_crypto/entities/lwe_bootstrap_key.rs:101-110
    = GgswCiphertextList            .../ggsw_ciphertext_list.rs:14-26
  SeededLweBootstrapKeyOwned<u64>  .../seeded_lwe_bootstrap_key.rs:20-30
    = SeededGgswCiphertextList      .../seeded_ggsw_ciphertext_list.rs:20-33
  LweKeyswitchKeyOwned<u64>        .../lwe_keyswitch_key.rs:79-90

Byte layout (derived from the struct definitions above, serde's derive -- fields
in declaration order, newtype structs transparent -- and bincode 1.x's default
configuration: little endian, fixed-width integers, u64 sequence lengths, u32
enum variant indexes, usize as u64):

  Vec<u64>                      u64 len, len x u64
  GlweSize / PolynomialSize / DecompositionBaseLog / DecompositionLevelCount /
  LweSize (newtypes over usize) u64                      (commons/parameters.rs:61-227)
  CiphertextModulus<u64>        SerializableCiphertextModulus { modulus: u128,
                                scalar_bits: usize }; modulus 0 = native 2^64
                                (commons/ciphertext_modulus.rs:25-60,80-93)
  CompressionSeed               { inner: AesCtrParams { seed: SeedKind,
                                first_index: TableIndex { aes_index: AesIndex(u128),
                                byte_index: ByteIndex(usize) } } }
                                (commons/math/random/generator.rs:22-24,
                                tfhe-csprng/src/generators/aes_ctr/mod.rs:213-220,
                                index.rs:23-58)
  SeedKind                      u32 variant: 0 Ctr(Seed(u128)), 1 Xof(XofSeed { data: Vec<u8> })
                                (tfhe-csprng/src/seeders/mod.rs:9-26,98-110)

Versioned encoding: every type that derives `Versionize` is wrapped in its
dispatch enum, i.e. preceded by the u32 index of its current variant
(core_crypto/backward_compatibility/entities/*.rs, commons/*.rs,
tfhe-csprng/src/*/backward_compatibility/mod.rs): LweBootstrapKey V1,
GgswCiphertextList V1, SeededLweBootstrapKey V1, SeededGgswCiphertextList V1,
LweKeyswitchKey V2, CompressionSeed V1, everything else V0; primitives and
Vec<u64> are not wrapped.

PARITY STATUS: UNPINNED.  The reference checkout holds no serialized key (its
`*.bcode` / `*.cbor` files are Git-LFS pointers) and Rust is not available to
produce one, so these layouts are pinned only by the definitions cited above and
by round trips of this module's own writer (tests/test_serde_ingest.py).

The parsed arrays feed the existing C-ABI conversions unchanged:
`bootstrap_key_to_gpu` / `seeded_bootstrap_key_to_gpu` / `keyswitch_key_to_gpu`.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Optional

import numpy as np

# current variant index of each versions-dispatch enum (see the module docstring)
_V = {"LweBootstrapKey": 1, "GgswCiphertextList": 1, "SeededLweBootstrapKey": 1, "SeededGgswCiphertextList": 1,
      "LweKeyswitchKey": 2, "CompressionSeed": 1}


class _Reader:
    def __init__(self, buf, versioned: bool):
        self.b = memoryview(buf).cast("B")
        self.o = 0
        self.versioned = versioned

    def take(self, n: int) -> memoryview:
        if self.o + n > len(self.b):
            raise ValueError("truncated input: need %d bytes at offset %d of %d" % (n, self.o, len(self.b)))
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def u32(self) -> int:
        return struct.unpack("<I", self.take(4))[0]

    def u64(self) -> int:
        return struct.unpack("<Q", self.take(8))[0]

    def u128(self) -> int:
        lo, hi = struct.unpack("<QQ", self.take(16))
        return lo | (hi << 64)

    def tag(self, name: str):
        """dispatch-enum variant index in front of a Versionize type"""
        if self.versioned:
            want = _V.get(name, 0)
            got = self.u32()
            if got != want:
                raise ValueError("%s: version variant %d, this reader knows V%d" % (name, got, want))

    def size(self, name: str) -> int:
        self.tag(name)
        return self.u64()

    def vec_u64(self) -> np.ndarray:
        n = self.u64()
        return np.frombuffer(self.take(8 * n), dtype="<u8").astype(np.uint64, copy=True)

    def vec_u8(self) -> bytes:
        return bytes(self.take(self.u64()))

    def modulus(self):
        self.tag("SerializableCiphertextModulus")
        modulus, bits = self.u128(), self.u64()
        if bits != 64:
            raise ValueError("CiphertextModulus for %d-bit scalars, expected 64" % bits)
        return modulus  # 0 = native 2^64

    def end(self):
        if self.o != len(self.b):
            raise ValueError("%d trailing bytes" % (len(self.b) - self.o))


class _Writer:
    def __init__(self, versioned: bool):
        self.parts = []
        self.versioned = versioned

    def u32(self, v):
        self.parts.append(struct.pack("<I", v))

    def u64(self, v):
        self.parts.append(struct.pack("<Q", v))

    def u128(self, v):
        self.parts.append(struct.pack("<QQ", v & (2 ** 64 - 1), v >> 64))

    def tag(self, name):
        if self.versioned:
            self.u32(_V.get(name, 0))

    def size(self, name, v):
        self.tag(name)
        self.u64(v)

    def vec_u64(self, a):
        a = np.ascontiguousarray(a, dtype="<u8").reshape(-1)
        self.u64(a.size)
        self.parts.append(a.tobytes())

    def modulus(self, modulus=0):
        self.tag("SerializableCiphertextModulus")
        self.u128(modulus)
        self.u64(64)

    def bytes(self) -> bytes:
        return b"".join(self.parts)


@dataclass
class LweBootstrapKey:
    """standard-domain BSK, data layout [i][level][row][col][N] as in the reference"""
    data: np.ndarray
    glwe_size: int
    polynomial_size: int
    decomp_base_log: int
    decomp_level_count: int
    ciphertext_modulus: int = 0  # 0 = native

    @property
    def glwe_dimension(self) -> int:
        return self.glwe_size - 1

    @property
    def input_lwe_dimension(self) -> int:
        per_ggsw = self.decomp_level_count * self.glwe_size * self.glwe_size * self.polynomial_size
        assert self.data.size % per_ggsw == 0, "container length is not a multiple of the GGSW size"
        return self.data.size // per_ggsw


@dataclass
class CompressionSeed:
    kind: str              # "ctr" (Seed(u128) is the AES key) or "xof"
    seed: int = 0          # kind == "ctr"
    xof_data: bytes = b""  # kind == "xof": domain separator || seed bytes
    aes_index: int = 0
    byte_index: int = 0


@dataclass
class SeededLweBootstrapKey:
    """bodies only, layout [i][level][row][N]; masks come from the seed"""
    data: np.ndarray
    glwe_size: int
    polynomial_size: int
    decomp_base_log: int
    decomp_level_count: int
    compression_seed: CompressionSeed
    ciphertext_modulus: int = 0

    @property
    def glwe_dimension(self) -> int:
        return self.glwe_size - 1

    @property
    def input_lwe_dimension(self) -> int:
        per_ggsw = self.decomp_level_count * self.glwe_size * self.polynomial_size
        assert self.data.size % per_ggsw == 0
        return self.data.size // per_ggsw


@dataclass
class LweKeyswitchKey:
    """layout [i][level][output_lwe_size]"""
    data: np.ndarray
    decomp_base_log: int
    decomp_level_count: int
    output_lwe_size: int
    ciphertext_modulus: int = 0

    @property
    def output_lwe_dimension(self) -> int:
        return self.output_lwe_size - 1

    @property
    def input_lwe_dimension(self) -> int:
        per_elem = self.decomp_level_count * self.output_lwe_size
        assert self.data.size % per_elem == 0
        return self.data.size // per_elem


def _check_native(modulus: int, what: str):
    if modulus != 0:
        raise ValueError("%s: non-native ciphertext modulus %d; the GPU path handles the native 2^64 torus" %
                         (what, modulus))


# ---- LweBootstrapKey -----------------------------------------------------------
def read_lwe_bootstrap_key(buf, versioned: bool = False) -> LweBootstrapKey:
    r = _Reader(buf, versioned)
    r.tag("LweBootstrapKey")
    r.tag("GgswCiphertextList")
    data = r.vec_u64()
    key = LweBootstrapKey(data, r.size("GlweSize"), r.size("PolynomialSize"), r.size("DecompositionBaseLog"),
                          r.size("DecompositionLevelCount"), r.modulus())
    r.end()
    _check_native(key.ciphertext_modulus, "LweBootstrapKey")
    key.input_lwe_dimension  # container-length check (ggsw_ciphertext_list.rs: from_container asserts the same)
    return key


def write_lwe_bootstrap_key(key: LweBootstrapKey, versioned: bool = False) -> bytes:
    w = _Writer(versioned)
    w.tag("LweBootstrapKey")
    w.tag("GgswCiphertextList")
    w.vec_u64(key.data)
    w.size("GlweSize", key.glwe_size)
    w.size("PolynomialSize", key.polynomial_size)
    w.size("DecompositionBaseLog", key.decomp_base_log)
    w.size("DecompositionLevelCount", key.decomp_level_count)
    w.modulus(key.ciphertext_modulus)
    return w.bytes()


# ---- SeededLweBootstrapKey -----------------------------------------------------
def _read_compression_seed(r: _Reader) -> CompressionSeed:
    r.tag("CompressionSeed")
    r.tag("AesCtrParams")
    r.tag("SeedKind")
    variant = r.u32()
    if variant == 0:
        r.tag("Seed")
        cs = CompressionSeed("ctr", seed=r.u128())
    elif variant == 1:
        r.tag("XofSeed")
        cs = CompressionSeed("xof", xof_data=r.vec_u8())
    else:
        raise ValueError("SeedKind variant %d" % variant)
    r.tag("TableIndex")
    r.tag("AesIndex")
    cs.aes_index = r.u128()
    cs.byte_index = r.size("ByteIndex")
    return cs


def _write_compression_seed(w: _Writer, cs: CompressionSeed):
    w.tag("CompressionSeed")
    w.tag("AesCtrParams")
    w.tag("SeedKind")
    if cs.kind == "ctr":
        w.u32(0)
        w.tag("Seed")
        w.u128(cs.seed)
    else:
        w.u32(1)
        w.tag("XofSeed")
        w.u64(len(cs.xof_data))
        w.parts.append(bytes(cs.xof_data))
    w.tag("TableIndex")
    w.tag("AesIndex")
    w.u128(cs.aes_index)
    w.size("ByteIndex", cs.byte_index)


def read_seeded_lwe_bootstrap_key(buf, versioned: bool = False) -> SeededLweBootstrapKey:
    r = _Reader(buf, versioned)
    r.tag("SeededLweBootstrapKey")
    r.tag("SeededGgswCiphertextList")
    data = r.vec_u64()
    glwe_size, poly = r.size("GlweSize"), r.size("PolynomialSize")
    base_log, level = r.size("DecompositionBaseLog"), r.size("DecompositionLevelCount")
    seed = _read_compression_seed(r)
    key = SeededLweBootstrapKey(data, glwe_size, poly, base_log, level, seed, r.modulus())
    r.end()
    _check_native(key.ciphertext_modulus, "SeededLweBootstrapKey")
    key.input_lwe_dimension
    return key


def write_seeded_lwe_bootstrap_key(key: SeededLweBootstrapKey, versioned: bool = False) -> bytes:
    w = _Writer(versioned)
    w.tag("SeededLweBootstrapKey")
    w.tag("SeededGgswCiphertextList")
    w.vec_u64(key.data)
    w.size("GlweSize", key.glwe_size)
    w.size("PolynomialSize", key.polynomial_size)
    w.size("DecompositionBaseLog", key.decomp_base_log)
    w.size("DecompositionLevelCount", key.decomp_level_count)
    _write_compression_seed(w, key.compression_seed)
    w.modulus(key.ciphertext_modulus)
    return w.bytes()


# ---- LweKeyswitchKey -----------------------------------------------------------
def read_lwe_keyswitch_key(buf, versioned: bool = False) -> LweKeyswitchKey:
    r = _Reader(buf, versioned)
    r.tag("LweKeyswitchKey")
    data = r.vec_u64()
    key = LweKeyswitchKey(data, r.size("DecompositionBaseLog"), r.size("DecompositionLevelCount"), r.size("LweSize"),
                          r.modulus())
    r.end()
    _check_native(key.ciphertext_modulus, "LweKeyswitchKey")
    key.input_lwe_dimension
    return key


def write_lwe_keyswitch_key(key: LweKeyswitchKey, versioned: bool = False) -> bytes:
    w = _Writer(versioned)
    w.tag("LweKeyswitchKey")
    w.vec_u64(key.data)
    w.size("DecompositionBaseLog", key.decomp_base_log)
    w.size("DecompositionLevelCount", key.decomp_level_count)
    w.size("LweSize", key.output_lwe_size)
    w.modulus(key.ciphertext_modulus)
    return w.bytes()


# ---- to the GPU (through the C ABI, like core_crypto::gpu does from the entities) -----
def bootstrap_key_to_gpu(buf, streams, ms_noise_reduction: Optional[str] = None, versioned: bool = False):
    """bytes of an LweBootstrapKeyOwned<u64> -> CudaLweBootstrapKey
    (cuda_convert_lwe_programmable_bootstrap_key_64_async)"""
    from . import gpu

    k = read_lwe_bootstrap_key(buf, versioned)
    return gpu.CudaLweBootstrapKey.from_lwe_bootstrap_key(
        k.data, k.input_lwe_dimension, k.glwe_dimension, k.polynomial_size, k.decomp_base_log, k.decomp_level_count,
        ms_noise_reduction, streams)


def seeded_bootstrap_key_to_gpu(buf, streams, ms_noise_reduction: Optional[str] = None, versioned: bool = False):
    """bytes of a SeededLweBootstrapKeyOwned<u64> -> CudaLweBootstrapKey; the masks are regenerated on the GPU
    (b200_convert_seeded_lwe_programmable_bootstrap_key_64_async).  SeedKind::Ctr only: the XOF key derivation is
    host-side SHAKE work that this engine does not restate."""
    from . import gpu

    k = read_seeded_lwe_bootstrap_key(buf, versioned)
    cs = k.compression_seed
    if cs.kind != "ctr":
        raise NotImplementedError("SeedKind::Xof: derive (key, counter offset) with tfhe-csprng first")
    return gpu.CudaLweBootstrapKey.from_seeded_lwe_bootstrap_key(
        k.data, cs.seed, k.input_lwe_dimension, k.glwe_dimension, k.polynomial_size, k.decomp_base_log,
        k.decomp_level_count, ms_noise_reduction, streams, first_aes_index=cs.aes_index,
        first_byte_index=cs.byte_index)


def keyswitch_key_to_gpu(buf, streams, versioned: bool = False):
    """bytes of an LweKeyswitchKeyOwned<u64> -> CudaLweKeyswitchKey (verbatim copy: the engine reads the host layout)"""
    from . import gpu

    k = read_lwe_keyswitch_key(buf, versioned)
    return gpu.CudaLweKeyswitchKey.from_lwe_keyswitch_key(k.data, k.input_lwe_dimension, k.output_lwe_dimension,
                                                          k.decomp_base_log, k.decomp_level_count, streams)
