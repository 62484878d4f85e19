"""Host-side mirror of tfhe-rs ``core_crypto::gpu`` for the PBS path.

The reference's host language is Rust, which this image does not have, so the
caller side of the C ABI is mirrored in Python with the SAME names, argument
order and assertion behaviour as the Rust functions it stands in for:

  CudaStreams                      tfhe/src/core_crypto/gpu/mod.rs
  CudaVec                          tfhe/src/core_crypto/gpu/vec.rs
  CudaLweCiphertextList            .../gpu/entities/lwe_ciphertext_list.rs
  CudaGlweCiphertextList           .../gpu/entities/glwe_ciphertext_list.rs
  CudaLweBootstrapKey              .../gpu/entities/lwe_bootstrap_key.rs:60-110
  CudaLweMultiBitBootstrapKey      .../gpu/entities/lwe_multi_bit_bootstrap_key.rs
  CudaLweKeyswitchKey              .../gpu/entities/lwe_keyswitch_key.rs
  cuda_programmable_bootstrap_lwe_ciphertext
                                   .../gpu/algorithms/lwe_programmable_bootstrapping.rs:10
  cuda_multi_bit_programmable_bootstrap_lwe_ciphertext
                                   .../gpu/algorithms/lwe_multi_bit_programmable_bootstrapping.rs:10
  cuda_keyswitch_lwe_ciphertext    .../gpu/algorithms/lwe_keyswitch.rs:12
  programmable_bootstrap (ffi)     tfhe/src/core_crypto/gpu/ffi.rs:21-92

PyTorch is used for device memory and streams only; all arithmetic happens in
libtfhe_cuda_backend_b200.so.  Nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("tfhe_rs_b200: no CUDA device available; the PBS path has no CPU fallback")


def _np_to_i64(a: np.ndarray) -> torch.Tensor:
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    return torch.from_numpy(a)


class CudaStreams:
    """One stream per GPU (gpu/mod.rs `CudaStreams`)."""

    def __init__(self, gpu_indexes: Sequence[int]):
        _require_cuda()
        self.gpu_indexes = list(gpu_indexes)
        self.streams = [torch.cuda.Stream(device=i) for i in self.gpu_indexes]

    @classmethod
    def new_single_gpu(cls, gpu_index: int = 0) -> "CudaStreams":
        return cls([gpu_index])

    @classmethod
    def new_multi_gpu(cls) -> "CudaStreams":
        _require_cuda()
        return cls(range(torch.cuda.device_count()))

    def __len__(self):
        return len(self.streams)

    def ptr(self, i: int = 0) -> int:
        return self.streams[i].cuda_stream

    def device(self, i: int = 0) -> torch.device:
        return torch.device("cuda", self.gpu_indexes[i])

    def synchronize(self):
        for s in self.streams:
            s.synchronize()

    def synchronize_one(self, i: int):
        self.streams[i].synchronize()


class CudaVec:
    """Typed device buffer on one GPU (gpu/vec.rs `CudaVec`)."""

    def __init__(self, tensor: torch.Tensor, np_dtype):
        self.t = tensor
        self.np_dtype = np.dtype(np_dtype)

    @classmethod
    def new(cls, length: int, streams: CudaStreams, stream_index: int = 0, np_dtype=np.uint64) -> "CudaVec":
        tdt = torch.float64 if np.dtype(np_dtype) == np.float64 else torch.int64
        with torch.cuda.stream(streams.streams[stream_index]):
            t = torch.zeros(length, dtype=tdt, device=streams.device(stream_index))
        return cls(t, np_dtype)

    @classmethod
    def from_cpu_async(cls, src: np.ndarray, streams: CudaStreams, stream_index: int = 0) -> "CudaVec":
        src = np.ascontiguousarray(src)
        host = _np_to_i64(src.reshape(-1)) if src.dtype != np.float64 else torch.from_numpy(src.reshape(-1))
        with torch.cuda.stream(streams.streams[stream_index]):
            t = host.to(streams.device(stream_index), non_blocking=True)
        return cls(t, src.dtype)

    def copy_from_cpu_async(self, src: np.ndarray, streams: CudaStreams, stream_index: int = 0):
        host = _np_to_i64(np.ascontiguousarray(src).reshape(-1))
        with torch.cuda.stream(streams.streams[stream_index]):
            self.t.copy_(host, non_blocking=True)

    def to_cpu(self, streams: Optional[CudaStreams] = None, stream_index: int = 0) -> np.ndarray:
        if streams is not None:
            with torch.cuda.stream(streams.streams[stream_index]):
                h = self.t.cpu()
        else:
            h = self.t.cpu()
        a = h.numpy()
        return a.view(self.np_dtype) if self.np_dtype == np.uint64 else a

    def __len__(self):
        return self.t.numel()

    def as_c_ptr(self) -> int:
        return self.t.data_ptr()

    @property
    def gpu_index(self) -> int:
        return self.t.device.index


def trivial_indexes(count: int, streams: CudaStreams, stream_index: int = 0) -> CudaVec:
    with torch.cuda.stream(streams.streams[stream_index]):
        t = torch.arange(count, dtype=torch.int64, device=streams.device(stream_index))
    return CudaVec(t, np.uint64)


@dataclass
class CudaLweCiphertextList:
    d_vec: CudaVec
    lwe_dimension: int
    lwe_ciphertext_count: int

    @classmethod
    def new(cls, lwe_dimension: int, count: int, streams: CudaStreams, stream_index: int = 0):
        return cls(CudaVec.new(count * (lwe_dimension + 1), streams, stream_index), lwe_dimension, count)

    @classmethod
    def from_lwe_ciphertext_list(cls, h_cts: np.ndarray, streams: CudaStreams, stream_index: int = 0):
        h_cts = np.ascontiguousarray(h_cts, dtype=np.uint64)
        assert h_cts.ndim == 2
        return cls(CudaVec.from_cpu_async(h_cts, streams, stream_index), h_cts.shape[1] - 1, h_cts.shape[0])

    def to_lwe_ciphertext_list(self, streams: Optional[CudaStreams] = None) -> np.ndarray:
        return self.d_vec.to_cpu(streams).reshape(self.lwe_ciphertext_count, self.lwe_dimension + 1)


@dataclass
class CudaGlweCiphertextList:
    d_vec: CudaVec
    glwe_dimension: int
    polynomial_size: int
    glwe_ciphertext_count: int

    @classmethod
    def from_glwe_ciphertext_list(cls, h: np.ndarray, glwe_dimension: int, polynomial_size: int,
                                  streams: CudaStreams, stream_index: int = 0):
        h = np.ascontiguousarray(h, dtype=np.uint64).reshape(-1, (glwe_dimension + 1) * polynomial_size)
        return cls(CudaVec.from_cpu_async(h, streams, stream_index), glwe_dimension, polynomial_size, h.shape[0])


class CudaModulusSwitchNoiseReductionConfiguration:
    """gpu/entities/lwe_bootstrap_key.rs:15 -- only `Centered` exists on GPU."""

    CENTERED = "Centered"


@dataclass
class CudaLweBootstrapKey:
    d_vec: CudaVec  # f64 words, engine-private Fourier layout
    input_lwe_dimension: int
    glwe_dimension: int
    polynomial_size: int
    decomp_base_log: int
    decomp_level_count: int
    ms_noise_reduction_configuration: Optional[str] = None

    @property
    def output_lwe_dimension(self) -> int:
        return self.glwe_dimension * self.polynomial_size

    @classmethod
    def from_lwe_bootstrap_key(cls, h_bsk: np.ndarray, input_lwe_dimension: int, glwe_dimension: int,
                               polynomial_size: int, decomp_base_log: int, decomp_level_count: int,
                               ms_noise_reduction_configuration: Optional[str], streams: CudaStreams,
                               stream_index: int = 0) -> "CudaLweBootstrapKey":
        """Upload + Fourier conversion of a standard-domain BSK (host u64,
        layout [i][level][row][col][N]); the device buffer is allocated by
        the caller side exactly as lwe_bootstrap_key.rs:77-85 does."""
        h_bsk = np.ascontiguousarray(h_bsk, dtype=np.uint64).reshape(-1)
        k1 = glwe_dimension + 1
        words = input_lwe_dimension * k1 * k1 * decomp_level_count * polynomial_size
        assert h_bsk.size == words, "bootstrap key size does not match its parameters"
        d_vec = CudaVec.new(words, streams, stream_index, np_dtype=np.float64)
        gi = streams.gpu_indexes[stream_index]
        _lib.lib().cuda_convert_lwe_programmable_bootstrap_key_64_async(
            streams.ptr(stream_index), gi, d_vec.as_c_ptr(), h_bsk.ctypes.data, input_lwe_dimension,
            glwe_dimension, decomp_level_count, polynomial_size)
        streams.synchronize_one(stream_index)  # h_bsk may be freed by the caller
        return cls(d_vec, input_lwe_dimension, glwe_dimension, polynomial_size, decomp_base_log,
                   decomp_level_count, ms_noise_reduction_configuration)

    @classmethod
    def from_seeded_lwe_bootstrap_key(cls, h_bodies: np.ndarray, compression_seed: int, input_lwe_dimension: int,
                                      glwe_dimension: int, polynomial_size: int, decomp_base_log: int,
                                      decomp_level_count: int, ms_noise_reduction_configuration: Optional[str],
                                      streams: CudaStreams, stream_index: int = 0,
                                      first_aes_index: int = 0, first_byte_index: int = 0) -> "CudaLweBootstrapKey":
        """Seeded key ingest (SeededLweBootstrapKey: bodies [i][level][row][N] +
        CompressionSeed(Seed(u128))): only the bodies are uploaded, the masks
        are regenerated on the GPU from the seed's AES-CTR table."""
        h_bodies = np.ascontiguousarray(h_bodies, dtype=np.uint64).reshape(-1)
        k1 = glwe_dimension + 1
        assert h_bodies.size == input_lwe_dimension * decomp_level_count * k1 * polynomial_size
        words = input_lwe_dimension * k1 * k1 * decomp_level_count * polynomial_size
        d_vec = CudaVec.new(words, streams, stream_index, np_dtype=np.float64)
        key = (compression_seed & (2 ** 128 - 1)).to_bytes(16, "little")
        gi = streams.gpu_indexes[stream_index]
        _lib.lib().b200_convert_seeded_lwe_programmable_bootstrap_key_64_async(
            streams.ptr(stream_index), gi, d_vec.as_c_ptr(), h_bodies.ctypes.data, key,
            first_aes_index & (2 ** 64 - 1), first_aes_index >> 64, first_byte_index, input_lwe_dimension,
            glwe_dimension, decomp_level_count, polynomial_size, 1)
        streams.synchronize_one(stream_index)
        return cls(d_vec, input_lwe_dimension, glwe_dimension, polynomial_size, decomp_base_log,
                   decomp_level_count, ms_noise_reduction_configuration)


@dataclass
class CudaLweMultiBitBootstrapKey:
    d_vec: CudaVec
    input_lwe_dimension: int
    glwe_dimension: int
    polynomial_size: int
    decomp_base_log: int
    decomp_level_count: int
    grouping_factor: int

    @property
    def output_lwe_dimension(self) -> int:
        return self.glwe_dimension * self.polynomial_size

    @classmethod
    def from_lwe_multi_bit_bootstrap_key(cls, h_bsk: np.ndarray, input_lwe_dimension: int, glwe_dimension: int,
                                         polynomial_size: int, decomp_base_log: int, decomp_level_count: int,
                                         grouping_factor: int, streams: CudaStreams, stream_index: int = 0):
        h_bsk = np.ascontiguousarray(h_bsk, dtype=np.uint64).reshape(-1)
        k1 = glwe_dimension + 1
        num_ggsw = (input_lwe_dimension // grouping_factor) << grouping_factor
        words = num_ggsw * k1 * k1 * decomp_level_count * polynomial_size
        assert h_bsk.size == words, "multi-bit bootstrap key size does not match its parameters"
        d_vec = CudaVec.new(words, streams, stream_index, np_dtype=np.float64)
        gi = streams.gpu_indexes[stream_index]
        _lib.lib().cuda_convert_lwe_multi_bit_programmable_bootstrap_key_64_async(
            streams.ptr(stream_index), gi, d_vec.as_c_ptr(), h_bsk.ctypes.data, input_lwe_dimension,
            glwe_dimension, decomp_level_count, polynomial_size, grouping_factor)
        streams.synchronize_one(stream_index)
        return cls(d_vec, input_lwe_dimension, glwe_dimension, polynomial_size, decomp_base_log,
                   decomp_level_count, grouping_factor)

    @classmethod
    def from_seeded_lwe_multi_bit_bootstrap_key(cls, h_bodies: np.ndarray, compression_seed: int,
                                                input_lwe_dimension: int, glwe_dimension: int, polynomial_size: int,
                                                decomp_base_log: int, decomp_level_count: int, grouping_factor: int,
                                                streams: CudaStreams, stream_index: int = 0):
        """Seeded multi-bit key ingest (bodies [ggsw][level][row][N] + seed)."""
        h_bodies = np.ascontiguousarray(h_bodies, dtype=np.uint64).reshape(-1)
        k1 = glwe_dimension + 1
        num_ggsw = (input_lwe_dimension // grouping_factor) << grouping_factor
        assert h_bodies.size == num_ggsw * decomp_level_count * k1 * polynomial_size
        d_vec = CudaVec.new(num_ggsw * k1 * k1 * decomp_level_count * polynomial_size, streams, stream_index,
                            np_dtype=np.float64)
        key = (compression_seed & (2 ** 128 - 1)).to_bytes(16, "little")
        gi = streams.gpu_indexes[stream_index]
        _lib.lib().b200_convert_seeded_lwe_programmable_bootstrap_key_64_async(
            streams.ptr(stream_index), gi, d_vec.as_c_ptr(), h_bodies.ctypes.data, key, 0, 0, 0,
            input_lwe_dimension, glwe_dimension, decomp_level_count, polynomial_size, grouping_factor)
        streams.synchronize_one(stream_index)
        return cls(d_vec, input_lwe_dimension, glwe_dimension, polynomial_size, decomp_base_log,
                   decomp_level_count, grouping_factor)


@dataclass
class CudaLweKeyswitchKey:
    d_vec: CudaVec  # verbatim copy of the host KSK [i][level][n_out+1]
    input_key_lwe_dimension: int
    output_key_lwe_dimension: int
    decomp_base_log: int
    decomp_level_count: int

    @classmethod
    def from_lwe_keyswitch_key(cls, h_ksk: np.ndarray, input_dim: int, output_dim: int, base_log: int,
                               level_count: int, streams: CudaStreams, stream_index: int = 0):
        h_ksk = np.ascontiguousarray(h_ksk, dtype=np.uint64).reshape(-1)
        assert h_ksk.size == input_dim * level_count * (output_dim + 1)
        return cls(CudaVec.from_cpu_async(h_ksk, streams, stream_index), input_dim, output_dim, base_log,
                   level_count)


# ---------------------------------------------------------------------------
# ffi.rs level: scratch -> run -> cleanup on every call
# ---------------------------------------------------------------------------
def programmable_bootstrap(streams: CudaStreams, lwe_array_out: CudaVec, lwe_out_indexes: CudaVec,
                           test_vector: CudaVec, test_vector_indexes: CudaVec, lwe_array_in: CudaVec,
                           lwe_in_indexes: CudaVec, bootstrapping_key: CudaVec, lwe_dimension: int,
                           glwe_dimension: int, polynomial_size: int, base_log: int, level: int,
                           num_samples: int, ms_noise_reduction_configuration: Optional[str],
                           num_many_lut: int = 1, lut_stride: int = 0, stream_index: int = 0):
    """core_crypto/gpu/ffi.rs:21-92."""
    L = _lib.lib()
    gi = streams.gpu_indexes[stream_index]
    sp = streams.ptr(stream_index)
    buf = C.POINTER(C.c_int8)()
    L.scratch_cuda_programmable_bootstrap_64_async(
        sp, gi, C.byref(buf), lwe_dimension, glwe_dimension, polynomial_size, level, num_samples, True,
        1 if ms_noise_reduction_configuration else 0)
    L.cuda_programmable_bootstrap_64_async(
        sp, gi, lwe_array_out.as_c_ptr(), lwe_out_indexes.as_c_ptr(), test_vector.as_c_ptr(),
        test_vector_indexes.as_c_ptr(), lwe_array_in.as_c_ptr(), lwe_in_indexes.as_c_ptr(),
        bootstrapping_key.as_c_ptr(), buf, lwe_dimension, glwe_dimension, polynomial_size, base_log, level,
        num_samples, num_many_lut, lut_stride)
    L.cleanup_cuda_programmable_bootstrap_64(sp, gi, C.byref(buf))


def programmable_bootstrap_multi_bit(streams: CudaStreams, lwe_array_out: CudaVec, lwe_out_indexes: CudaVec,
                                     test_vector: CudaVec, test_vector_indexes: CudaVec, lwe_array_in: CudaVec,
                                     lwe_in_indexes: CudaVec, bootstrapping_key: CudaVec, lwe_dimension: int,
                                     glwe_dimension: int, polynomial_size: int, base_log: int, level: int,
                                     grouping_factor: int, num_samples: int, num_many_lut: int = 1,
                                     lut_stride: int = 0, stream_index: int = 0):
    """core_crypto/gpu/ffi.rs:208-... (multi-bit twin)."""
    L = _lib.lib()
    gi = streams.gpu_indexes[stream_index]
    sp = streams.ptr(stream_index)
    buf = C.POINTER(C.c_int8)()
    L.scratch_cuda_multi_bit_programmable_bootstrap_64_async(
        sp, gi, C.byref(buf), glwe_dimension, polynomial_size, level, num_samples, True)
    L.cuda_multi_bit_programmable_bootstrap_64_async(
        sp, gi, lwe_array_out.as_c_ptr(), lwe_out_indexes.as_c_ptr(), test_vector.as_c_ptr(),
        test_vector_indexes.as_c_ptr(), lwe_array_in.as_c_ptr(), lwe_in_indexes.as_c_ptr(),
        bootstrapping_key.as_c_ptr(), buf, lwe_dimension, glwe_dimension, polynomial_size, grouping_factor,
        base_log, level, num_samples, num_many_lut, lut_stride)
    L.cleanup_cuda_multi_bit_programmable_bootstrap_64(sp, gi, C.byref(buf))


class PbsScratch:
    """Keeps the scratch object alive across calls (what the integer layer's
    int_radix_lut does) instead of scratch/cleanup per call."""

    def __init__(self, streams: CudaStreams, lwe_dimension: int, glwe_dimension: int, polynomial_size: int,
                 level: int, num_samples: int, centered: bool, multi_bit: bool = False, stream_index: int = 0):
        self.streams, self.stream_index, self.multi_bit = streams, stream_index, multi_bit
        self.buf = C.POINTER(C.c_int8)()
        L = _lib.lib()
        gi, sp = streams.gpu_indexes[stream_index], streams.ptr(stream_index)
        if multi_bit:
            L.scratch_cuda_multi_bit_programmable_bootstrap_64_async(
                sp, gi, C.byref(self.buf), glwe_dimension, polynomial_size, level, num_samples, True)
        else:
            L.scratch_cuda_programmable_bootstrap_64_async(
                sp, gi, C.byref(self.buf), lwe_dimension, glwe_dimension, polynomial_size, level, num_samples,
                True, 1 if centered else 0)

    def close(self):
        if self.buf:
            L = _lib.lib()
            gi, sp = self.streams.gpu_indexes[self.stream_index], self.streams.ptr(self.stream_index)
            if self.multi_bit:
                L.cleanup_cuda_multi_bit_programmable_bootstrap_64(sp, gi, C.byref(self.buf))
            else:
                L.cleanup_cuda_programmable_bootstrap_64(sp, gi, C.byref(self.buf))
            self.buf = C.POINTER(C.c_int8)()


# ---------------------------------------------------------------------------
# algorithms level
# ---------------------------------------------------------------------------
def _assert_same_gpu(streams: CudaStreams, *vecs: CudaVec):
    gi = streams.gpu_indexes[0]
    for v in vecs:
        assert v.gpu_index == gi, f"GPU error: all data should reside on GPU {gi} (found {v.gpu_index})"


def cuda_programmable_bootstrap_lwe_ciphertext(input: CudaLweCiphertextList, output: CudaLweCiphertextList,
                                               accumulator: CudaGlweCiphertextList, lut_indexes: CudaVec,
                                               output_indexes: CudaVec, input_indexes: CudaVec,
                                               bsk: CudaLweBootstrapKey, streams: CudaStreams):
    """gpu/algorithms/lwe_programmable_bootstrapping.rs:10-140."""
    assert input.lwe_dimension == bsk.input_lwe_dimension, "Mismatched input LweDimension"
    assert output.lwe_dimension == bsk.output_lwe_dimension, "Mismatched output LweDimension"
    assert accumulator.glwe_dimension == bsk.glwe_dimension, "Mismatched GlweSize"
    assert accumulator.polynomial_size == bsk.polynomial_size, "Mismatched PolynomialSize"
    _assert_same_gpu(streams, input.d_vec, output.d_vec, accumulator.d_vec, lut_indexes, output_indexes,
                     input_indexes, bsk.d_vec)
    num_samples = input.lwe_ciphertext_count
    assert len(lut_indexes) >= num_samples and len(output_indexes) >= num_samples and len(input_indexes) >= num_samples
    programmable_bootstrap(streams, output.d_vec, output_indexes, accumulator.d_vec, lut_indexes, input.d_vec,
                           input_indexes, bsk.d_vec, input.lwe_dimension, bsk.glwe_dimension,
                           bsk.polynomial_size, bsk.decomp_base_log, bsk.decomp_level_count, num_samples,
                           bsk.ms_noise_reduction_configuration)


def cuda_multi_bit_programmable_bootstrap_lwe_ciphertext(input: CudaLweCiphertextList,
                                                         output: CudaLweCiphertextList,
                                                         accumulator: CudaGlweCiphertextList,
                                                         lut_indexes: CudaVec, output_indexes: CudaVec,
                                                         input_indexes: CudaVec,
                                                         multi_bit_bsk: CudaLweMultiBitBootstrapKey,
                                                         streams: CudaStreams):
    """gpu/algorithms/lwe_multi_bit_programmable_bootstrapping.rs:10-145."""
    b = multi_bit_bsk
    assert input.lwe_dimension == b.input_lwe_dimension, "Mismatched input LweDimension"
    assert output.lwe_dimension == b.output_lwe_dimension, "Mismatched output LweDimension"
    assert accumulator.glwe_dimension == b.glwe_dimension, "Mismatched GlweSize"
    assert accumulator.polynomial_size == b.polynomial_size, "Mismatched PolynomialSize"
    _assert_same_gpu(streams, input.d_vec, output.d_vec, accumulator.d_vec, lut_indexes, output_indexes,
                     input_indexes, b.d_vec)
    programmable_bootstrap_multi_bit(streams, output.d_vec, output_indexes, accumulator.d_vec, lut_indexes,
                                     input.d_vec, input_indexes, b.d_vec, input.lwe_dimension, b.glwe_dimension,
                                     b.polynomial_size, b.decomp_base_log, b.decomp_level_count,
                                     b.grouping_factor, input.lwe_ciphertext_count)


def cuda_keyswitch_lwe_ciphertext(lwe_keyswitch_key: CudaLweKeyswitchKey,
                                  input_lwe_ciphertext: CudaLweCiphertextList,
                                  output_lwe_ciphertext: CudaLweCiphertextList, input_indexes: CudaVec,
                                  output_indexes: CudaVec, uses_trivial_indices: bool, streams: CudaStreams,
                                  use_gemm_ks: bool = True, stream_index: int = 0):
    """gpu/algorithms/lwe_keyswitch.rs:12-140."""
    k = lwe_keyswitch_key
    assert k.input_key_lwe_dimension == input_lwe_ciphertext.lwe_dimension, "Mismatched input LweDimension"
    assert k.output_key_lwe_dimension == output_lwe_ciphertext.lwe_dimension, "Mismatched output LweDimension"
    _assert_same_gpu(streams, k.d_vec, input_lwe_ciphertext.d_vec, output_lwe_ciphertext.d_vec, input_indexes,
                     output_indexes)
    L = _lib.lib()
    gi, sp = streams.gpu_indexes[stream_index], streams.ptr(stream_index)
    args = (sp, gi, output_lwe_ciphertext.d_vec.as_c_ptr(), output_indexes.as_c_ptr(),
            input_lwe_ciphertext.d_vec.as_c_ptr(), input_indexes.as_c_ptr(), k.d_vec.as_c_ptr(),
            k.input_key_lwe_dimension, k.output_key_lwe_dimension, k.decomp_base_log, k.decomp_level_count,
            input_lwe_ciphertext.lwe_ciphertext_count)
    if use_gemm_ks:
        L.cuda_keyswitch_gemm_64_64_async(*args, uses_trivial_indices)
    else:
        L.cuda_keyswitch_lwe_ciphertext_vector_64_64_async(*args)


def forward_negacyclic_fft(input: CudaVec, output: CudaVec, polynomial_size: int, total_polynomials: int,
                           streams: CudaStreams, stream_index: int = 0):
    """Test entry (role of gpu::forward_fft16x4x16_async, test/fft/mod.rs:76-100)."""
    _lib.lib().b200_forward_negacyclic_fft_async(streams.ptr(stream_index), streams.gpu_indexes[stream_index],
                                                input.as_c_ptr(), output.as_c_ptr(), polynomial_size,
                                                total_polynomials)


def cuda_extract_lwe_samples_from_glwe_ciphertext_list(glwes: "CudaGlweCiphertextList", nths: Sequence[int],
                                                       lwes_per_glwe: int, streams: CudaStreams,
                                                       stream_index: int = 0) -> "CudaLweCiphertextList":
    """core_crypto::gpu::cuda_extract_lwe_samples_from_glwe_ciphertext_list
    (gpu/algorithms/glwe_sample_extraction.rs): LWE i = coefficient nths[i] of
    GLWE i // lwes_per_glwe."""
    nth = np.ascontiguousarray(nths, dtype=np.uint32)
    with torch.cuda.stream(streams.streams[stream_index]):
        d_nth = torch.from_numpy(nth.astype(np.int32)).to(streams.device(stream_index))
    k, N = glwes.glwe_dimension, glwes.polynomial_size
    out = CudaLweCiphertextList.new(k * N, len(nth), streams, stream_index)
    _lib.lib().cuda_glwe_sample_extract_64_async(
        streams.ptr(stream_index), streams.gpu_indexes[stream_index], out.d_vec.as_c_ptr(), glwes.d_vec.as_c_ptr(),
        d_nth.data_ptr(), len(nth), lwes_per_glwe, N, k, N)
    streams.synchronize_one(stream_index)  # d_nth may go out of scope
    return out


def cuda_modulus_switch_ciphertext(ct: CudaVec, log_modulus: int, streams: CudaStreams, stream_index: int = 0):
    """In place, element-wise (gpu/algorithms/modulus_switch.rs; ciphertext.h:21-23)."""
    _lib.lib().cuda_modulus_switch_inplace_64_async(streams.ptr(stream_index), streams.gpu_indexes[stream_index],
                                                   ct.as_c_ptr(), len(ct), log_modulus)


def cuda_centered_modulus_switch_ciphertext(ct_in: CudaVec, lwe_dimension: int, log_modulus: int,
                                            streams: CudaStreams, stream_index: int = 0) -> CudaVec:
    """One LWE, centered-mean noise reduction (ciphertext.h:29-32)."""
    out = CudaVec.new(lwe_dimension + 1, streams, stream_index)
    _lib.lib().cuda_centered_modulus_switch_64_async(streams.ptr(stream_index), streams.gpu_indexes[stream_index],
                                                    out.as_c_ptr(), ct_in.as_c_ptr(), lwe_dimension, log_modulus)
    return out
