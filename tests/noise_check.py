#!/usr/bin/env python
"""One-off measurement (not part of the bench contract): output-noise variance
of the GPU PBS on real keys against the reference's formula
`pbs_variance_132_bits_security_tuniform_fft_mul`
(commons/noise_formulas/lwe_programmable_bootstrap.rs:86-150), on fresh
encryptions under the small key (the protocol of
gpu/algorithms/test/noise_distribution/lwe_programmable_bootstrapping_noise.rs:
PBS only, no keyswitch).  Keys, inputs and decryption come from the oracle
(test infrastructure); the PBS runs through the C ABI.  Prints one JSON line
per parameter set."""
import argparse
import dataclasses
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/: it uses the oracle)
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=8192)
    ap.add_argument("--oracle-samples", type=int, default=0, help="also run the CPU oracle (FFT mode) on this many")
    args = ap.parse_args()
    from oracle import csprng, oracle as O
    from tests.noise_formula import multi_bit_pbs_variance_tuniform_fft, pbs_variance_tuniform_fft
    from tfhe_rs_b200 import algorithms, gpu, server_key

    streams = gpu.CudaStreams.new_single_gpu(0)
    sets = [("classic P22 (standard modulus switch)", dataclasses.replace(O.PARAM_MESSAGE_2_CARRY_2_KS_PBS, centered_ms=False)),
            ("classic P22 (centered-mean modulus switch)", O.PARAM_MESSAGE_2_CARRY_2_KS_PBS),
            ("multi-bit g=3", O.PARAM_MULTI_BIT_GROUP_3_MESSAGE_2_CARRY_2_KS_PBS),
            ("multi-bit g=4", csprng.PARAM_GPU_MULTI_BIT_GROUP_4_MESSAGE_2_CARRY_2_KS_PBS)]
    for name, P in sets:
        keys = O.keygen(P, 0xB2000001, with_ksk=False)
        skey = server_key.upload_server_key(
            keys.bsk, np.zeros(P.big_n * P.ks_level * (P.n + 1), dtype=np.uint64), n=P.n, k=P.k, N=P.N,
            pbs_base_log=P.pbs_base_log, pbs_level=P.pbs_level, ks_base_log=P.ks_base_log, ks_level=P.ks_level,
            grouping_factor=P.grouping_factor, centered_ms=P.centered_ms, streams=streams)
        msgs = np.arange(args.samples) % 16
        cts = O.lwe_encrypt_batch(O.Rng(5), keys.lwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta), P.lwe_noise_log2)
        lut = algorithms.generate_programmable_bootstrap_glwe_lut(P.N, P.k + 1, 16, P.delta, lambda x: x)
        d_in = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(cts, streams)
        d_lut = gpu.CudaGlweCiphertextList.from_glwe_ciphertext_list(lut, P.k, P.N, streams)
        out = skey.bootstrap(d_in, d_lut).to_lwe_ciphertext_list(streams)
        ph = O.lwe_decrypt_batch(keys.glwe_sk, out)
        dec = O.decode(ph, P.delta, 16)
        with np.errstate(over="ignore"):
            err = (ph - msgs.astype(np.uint64) * np.uint64(P.delta)).astype(np.int64) / 2.0 ** 64
        if P.grouping_factor > 1:
            bound = multi_bit_pbs_variance_tuniform_fft(P.n, P.k, P.N, P.pbs_base_log, P.pbs_level, P.grouping_factor)
        else:
            bound = pbs_variance_tuniform_fft(P.n, P.k, P.N, P.pbs_base_log, P.pbs_level)
        extra = {}
        if args.oracle_samples:
            m = args.oracle_samples
            ref = O.pbs_batch(keys, lut, cts[:m])
            with np.errstate(over="ignore"):
                rerr = (O.lwe_decrypt_batch(keys.glwe_sk, ref) - msgs[:m].astype(np.uint64) * np.uint64(P.delta)
                        ).astype(np.int64) / 2.0 ** 64
            extra = {"oracle_samples": m, "oracle_variance": float(rerr.var()),
                     "oracle_ratio": float(rerr.var() / bound)}
        print(json.dumps({"set": name, "samples": args.samples, "decode_errors": int((dec != msgs).sum()),
                          "variance": float(err.var()), "formula": bound, "ratio": float(err.var() / bound),
                          "mean_over_std": float(err.mean() / err.std()),
                          "max_abs_over_std": float(np.abs(err).max() / err.std()), **extra}))


if __name__ == "__main__":
    main()
