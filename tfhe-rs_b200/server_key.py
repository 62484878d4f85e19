"""Device-side KS -> PBS atomic pattern for batches of shortint blocks.

Mirrors StandardAtomicPatternServerKey::apply_lookup_table_assign for the
KS_PBS order (tfhe/src/shortint/atomic_pattern/standard.rs:162-199):
keyswitch big-key LWEs to the small key, then bootstrap with per-sample LUT
indexes.  Also the GPU twin `execute_keyswitch_async` + `execute_pbs_async`
(backends/tfhe-cuda-backend/cuda/src/integer/integer.cuh:869-1000).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import gpu


@dataclass
class CudaServerKey:
    """Evaluation keys of one parameter set resident on one GPU."""

    ksk: gpu.CudaLweKeyswitchKey
    bsk: object  # CudaLweBootstrapKey | CudaLweMultiBitBootstrapKey
    streams: gpu.CudaStreams

    @property
    def multi_bit(self) -> bool:
        return isinstance(self.bsk, gpu.CudaLweMultiBitBootstrapKey)

    @property
    def small_dim(self) -> int:
        return self.bsk.input_lwe_dimension

    @property
    def big_dim(self) -> int:
        return self.bsk.output_lwe_dimension

    def keyswitch(self, cts_big: gpu.CudaLweCiphertextList, out: Optional[gpu.CudaLweCiphertextList] = None,
                  in_idx: Optional[gpu.CudaVec] = None, out_idx: Optional[gpu.CudaVec] = None):
        n = cts_big.lwe_ciphertext_count
        if out is None:
            out = gpu.CudaLweCiphertextList.new(self.small_dim, n, self.streams)
        triv = in_idx is None and out_idx is None
        if in_idx is None:
            in_idx = gpu.trivial_indexes(n, self.streams)
        if out_idx is None:
            out_idx = in_idx
        gpu.cuda_keyswitch_lwe_ciphertext(self.ksk, cts_big, out, in_idx, out_idx, triv, self.streams)
        return out

    def bootstrap(self, cts_small: gpu.CudaLweCiphertextList, luts: gpu.CudaGlweCiphertextList,
                  lut_idx: Optional[gpu.CudaVec] = None, out: Optional[gpu.CudaLweCiphertextList] = None,
                  in_idx: Optional[gpu.CudaVec] = None, out_idx: Optional[gpu.CudaVec] = None):
        n = cts_small.lwe_ciphertext_count
        if out is None:
            out = gpu.CudaLweCiphertextList.new(self.big_dim, n, self.streams)
        if in_idx is None:
            in_idx = gpu.trivial_indexes(n, self.streams)
        if out_idx is None:
            out_idx = in_idx
        if lut_idx is None:
            lut_idx = gpu.CudaVec.new(n, self.streams)  # all zeros: one shared LUT
        if self.multi_bit:
            gpu.cuda_multi_bit_programmable_bootstrap_lwe_ciphertext(cts_small, out, luts, lut_idx, out_idx,
                                                                     in_idx, self.bsk, self.streams)
        else:
            gpu.cuda_programmable_bootstrap_lwe_ciphertext(cts_small, out, luts, lut_idx, out_idx, in_idx,
                                                           self.bsk, self.streams)
        return out

    def apply_lookup_table(self, cts_big: gpu.CudaLweCiphertextList, luts: gpu.CudaGlweCiphertextList,
                           lut_idx: Optional[gpu.CudaVec] = None) -> gpu.CudaLweCiphertextList:
        """KS then PBS (EncryptionKeyChoice::Big), standard.rs:162-182."""
        return self.bootstrap(self.keyswitch(cts_big), luts, lut_idx)


def upload_server_key(h_bsk: np.ndarray, h_ksk: np.ndarray, *, n: int, k: int, N: int, pbs_base_log: int,
                      pbs_level: int, ks_base_log: int, ks_level: int, grouping_factor: int = 1,
                      centered_ms: bool = True, streams: Optional[gpu.CudaStreams] = None) -> CudaServerKey:
    streams = streams or gpu.CudaStreams.new_single_gpu(0)
    if grouping_factor > 1:
        bsk = gpu.CudaLweMultiBitBootstrapKey.from_lwe_multi_bit_bootstrap_key(
            h_bsk, n, k, N, pbs_base_log, pbs_level, grouping_factor, streams)
    else:
        bsk = gpu.CudaLweBootstrapKey.from_lwe_bootstrap_key(
            h_bsk, n, k, N, pbs_base_log, pbs_level,
            gpu.CudaModulusSwitchNoiseReductionConfiguration.CENTERED if centered_ms else None, streams)
    ksk = gpu.CudaLweKeyswitchKey.from_lwe_keyswitch_key(h_ksk, k * N, n, ks_base_log, ks_level, streams)
    return CudaServerKey(ksk, bsk, streams)
