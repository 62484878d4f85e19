#!/usr/bin/env python
"""Side measurements for the non-headline BASELINE configs (synthetic data,
CUDA events on the launch stream):
  configs[2]  multi-bit PBS batch=4096, PARAM_MULTI_BIT_GROUP_3 (N=2048, l=2)
  keyswitch   batch=4096, kN=2048 -> n=918, 4 levels
Prints one JSON line per measurement.  Not part of the bench contract."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--what", default="multibit,ks")
    ap.add_argument("--grouping", type=int, default=3, help="3: PARAM_MULTI_BIT_GROUP_3 (l=2, logB=15); "
                    "4: PARAM_GPU_MULTI_BIT_GROUP_4 (n=920, l=1, logB=22)")
    args = ap.parse_args()
    import torch

    import tfhe_rs_b200
    from tfhe_rs_b200 import gpu

    L = tfhe_rs_b200.lib()
    streams = gpu.CudaStreams.new_single_gpu(0)
    stream = streams.streams[0]
    rng = np.random.default_rng(3)
    batch = args.batch

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        with torch.cuda.stream(stream):
            fn()
            for s, e in evs:
                s.record(stream)
                fn()
                e.record(stream)
        streams.synchronize()
        return float(np.mean([s.elapsed_time(e) for s, e in evs]))

    if "multibit" in args.what:
        n, k, N, bl, lv, g = (918, 1, 2048, 15, 2, 3) if args.grouping == 3 else (920, 1, 2048, 22, 1, 4)
        num_ggsw = (n // g) << g
        words = num_ggsw * lv * 4 * N
        h = rng.integers(0, 1 << 64, size=words, dtype=np.uint64)
        bsk = gpu.CudaLweMultiBitBootstrapKey.from_lwe_multi_bit_bootstrap_key(h, n, k, N, bl, lv, g, streams)
        del h
        d_in = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(
            rng.integers(0, 1 << 64, size=(batch, n + 1), dtype=np.uint64), streams)
        d_out = gpu.CudaLweCiphertextList.new(k * N, batch, streams)
        lut = np.zeros(2 * N, dtype=np.uint64)
        lut[N:] = np.repeat(np.arange(16, dtype=np.uint64) << np.uint64(59), N // 16)
        d_lut = gpu.CudaGlweCiphertextList.from_glwe_ciphertext_list(lut, k, N, streams)
        idx = gpu.trivial_indexes(batch, streams)
        lidx = gpu.CudaVec.new(batch, streams)
        sc = gpu.PbsScratch(streams, n, k, N, lv, batch, centered=False, multi_bit=True)

        def run():
            L.cuda_multi_bit_programmable_bootstrap_64_async(
                streams.ptr(0), 0, d_out.d_vec.as_c_ptr(), idx.as_c_ptr(), d_lut.d_vec.as_c_ptr(), lidx.as_c_ptr(),
                d_in.d_vec.as_c_ptr(), idx.as_c_ptr(), bsk.d_vec.as_c_ptr(), sc.buf, n, k, N, g, bl, lv, batch, 1, 0)

        ms = timed(run, args.steps)
        sc.close()
        bytes_per_pbs = num_ggsw * lv * 4 * (N // 2) * 16
        print(json.dumps({"what": f"multi-bit PBS g={g}, N=2048, l={lv}" + (" (generic kernel)" if os.environ.get("B200_MULTIBIT_GENERIC") else " (register-FFT kernel)"), "batch": batch, "ms": ms,
                          "pbs_per_s": batch / ms * 1e3, "algorithmic_GBps": bytes_per_pbs * batch / ms / 1e6}))
    if "classic" in args.what:
        # other classic sets through whatever kernel the dispatcher picks (synthetic key):
        # 1_1: n=879,k=4,N=512,l=1,logB=23; 2_2 (headline) for comparison
        for name, (n, k, N, bl, lv) in (("PARAM_MESSAGE_1_CARRY_1", (879, 4, 512, 23, 1)),
                                         ("PARAM_MESSAGE_2_CARRY_2", (918, 1, 2048, 23, 1))):
            words = n * lv * (k + 1) * (k + 1) * N
            h = rng.integers(0, 1 << 64, size=words, dtype=np.uint64)
            bsk = gpu.CudaLweBootstrapKey.from_lwe_bootstrap_key(h, n, k, N, bl, lv, None, streams)
            del h
            d_in = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(
                rng.integers(0, 1 << 64, size=(batch, n + 1), dtype=np.uint64), streams)
            d_out = gpu.CudaLweCiphertextList.new(k * N, batch, streams)
            lut = np.zeros((k + 1) * N, dtype=np.uint64)
            lut[k * N:] = np.repeat(np.arange(16, dtype=np.uint64) << np.uint64(59), N // 16)
            d_lut = gpu.CudaGlweCiphertextList.from_glwe_ciphertext_list(lut, k, N, streams)
            idx = gpu.trivial_indexes(batch, streams)
            lidx = gpu.CudaVec.new(batch, streams)
            sc = gpu.PbsScratch(streams, n, k, N, lv, batch, centered=False, multi_bit=False)

            def run():
                L.cuda_programmable_bootstrap_64_async(
                    streams.ptr(0), 0, d_out.d_vec.as_c_ptr(), idx.as_c_ptr(), d_lut.d_vec.as_c_ptr(),
                    lidx.as_c_ptr(), d_in.d_vec.as_c_ptr(), idx.as_c_ptr(), bsk.d_vec.as_c_ptr(), sc.buf, n, k, N,
                    bl, lv, batch, 1, 0)

            ms = timed(run, args.steps)
            sc.close()
            print(json.dumps({"what": f"classic PBS {name} (n={n}, k={k}, N={N}, l={lv})",
                              "register_kernel": bool(L.b200_pbs_uses_fast_path(n, k, N, lv)), "batch": batch,
                              "ms": ms, "pbs_per_s": batch / ms * 1e3}))
    if "ks" in args.what:
        nin, nout, bl, lv = 2048, 918, 4, 4
        ksk = gpu.CudaLweKeyswitchKey.from_lwe_keyswitch_key(
            rng.integers(0, 1 << 64, size=nin * lv * (nout + 1), dtype=np.uint64), nin, nout, bl, lv, streams)
        d_in = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(
            rng.integers(0, 1 << 64, size=(batch, nin + 1), dtype=np.uint64), streams)
        d_out = gpu.CudaLweCiphertextList.new(nout, batch, streams)
        idx = gpu.trivial_indexes(batch, streams)

        def run():
            gpu.cuda_keyswitch_lwe_ciphertext(ksk, d_in, d_out, idx, idx, True, streams)

        macs = batch * nin * lv * (nout + 1)
        for path, name in ((1, "int8 tensor cores"), (2, "fp64 pipe"), (3, "integer pipe")):
            L.b200_set_keyswitch_path(path)
            ms = timed(run, max(args.steps, 3))
            print(json.dumps({"what": f"keyswitch 2048->918 l=4 ({name})", "batch": batch, "ms": ms,
                              "ks_per_s": batch / ms * 1e3, "u64_GMAC_per_s": macs / ms / 1e6,
                              "int8_TOPS": 2 * 8 * macs / ms / 1e9 if path == 1 else None}))
        L.b200_set_keyswitch_path(0)


if __name__ == "__main__":
    main()
