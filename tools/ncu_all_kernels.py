#!/usr/bin/env python
"""Launches every CUDA kernel of libtfhe_cuda_backend_b200.so once or twice, with
synthetic keys of the real sizes and modest batches, so that ONE
  ncu --set full --clock-control none -k regex:b200 -o <rep> python tools/ncu_all_kernels.py
captures all of them (profiles/r2_all_kernels.txt is the committed summary,
made with `python profiles/summarize.py table <rep>`).  Measurement tool only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import tfhe_rs_b200
    from tfhe_rs_b200 import gpu

    L = tfhe_rs_b200.lib()
    streams = gpu.CudaStreams.new_single_gpu(0)
    rng = np.random.default_rng(11)
    u64 = lambda *shape: rng.integers(0, 1 << 64, size=shape, dtype=np.uint64)

    def lut_for(k, N):
        lut = np.zeros((k + 1) * N, dtype=np.uint64)
        lut[k * N:] = np.repeat(np.arange(16, dtype=np.uint64) << np.uint64(59), N // 16)
        return gpu.CudaGlweCiphertextList.from_glwe_ciphertext_list(lut, k, N, streams)

    def classic(bsk, n, k, N, bl, lv, batch):
        d_in = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(u64(batch, n + 1), streams)
        d_out = gpu.CudaLweCiphertextList.new(k * N, batch, streams)
        idx, lidx = gpu.trivial_indexes(batch, streams), gpu.CudaVec.new(batch, streams)
        d_lut = lut_for(k, N)
        sc = gpu.PbsScratch(streams, n, k, N, lv, batch, centered=True)
        L.cuda_programmable_bootstrap_64_async(
            streams.ptr(0), 0, d_out.d_vec.as_c_ptr(), idx.as_c_ptr(), d_lut.d_vec.as_c_ptr(), lidx.as_c_ptr(),
            d_in.d_vec.as_c_ptr(), idx.as_c_ptr(), bsk.d_vec.as_c_ptr(), sc.buf, n, k, N, bl, lv, batch, 1, 0)
        streams.synchronize()
        sc.close()

    def multibit(mbsk, n, k, N, bl, lv, g, batch):
        d_in = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(u64(batch, n + 1), streams)
        d_out = gpu.CudaLweCiphertextList.new(k * N, batch, streams)
        idx, lidx = gpu.trivial_indexes(batch, streams), gpu.CudaVec.new(batch, streams)
        d_lut = lut_for(k, N)
        sc = gpu.PbsScratch(streams, n, k, N, lv, batch, centered=False, multi_bit=True)
        L.cuda_multi_bit_programmable_bootstrap_64_async(
            streams.ptr(0), 0, d_out.d_vec.as_c_ptr(), idx.as_c_ptr(), d_lut.d_vec.as_c_ptr(), lidx.as_c_ptr(),
            d_in.d_vec.as_c_ptr(), idx.as_c_ptr(), mbsk.d_vec.as_c_ptr(), sc.buf, n, k, N, g, bl, lv, batch, 1, 0)
        streams.synchronize()
        sc.close()

    # ---- classic fast path (N = 2048, k = 1, l = 1): conversion, shipped kernels, earlier generations
    n, k, N, bl, lv = 918, 1, 2048, 23, 1
    bsk = gpu.CudaLweBootstrapKey.from_lwe_bootstrap_key(u64(n * 4 * N), n, k, N, bl, lv, "Centered", streams)
    classic(bsk, n, k, N, bl, lv, 148)     # v6 <0, true>: TMA ring, one CTA per SM
    classic(bsk, n, k, N, bl, lv, 592)     # v6 <2, true>: hybrid, two CTAs per SM
    for variant in (12, 13, 5, 1):         # v7, v7 x3, round-2 start (v3), first kernel
        L.b200_set_pbs_variant(variant)
        classic(bsk, n, k, N, bl, lv, 592)
    L.b200_set_pbs_variant(0)
    # seeded ingest: AES-CTR mask expansion + the same conversion kernel
    gpu.CudaLweBootstrapKey.from_seeded_lwe_bootstrap_key(u64(n * lv * 2 * N), 0x1234, n, k, N, bl, lv, "Centered",
                                                          streams)
    del bsk
    # ---- N = 512 register kernel (1_1: k = 4; TMA key ring with 3 / 1 LWEs per CTA, register ring), generic kernels:
    # (k = 2, N = 1024, l = 2) in shared memory, 3_3 (N = 8192, l = 2) over the global workspace and on the
    # tensor-memory register kernels (csrc/pbs_n8192.cuh)
    sb = gpu.CudaLweBootstrapKey.from_lwe_bootstrap_key(u64(879 * 25 * 512), 879, 4, 512, 23, 1, "Centered", streams)
    classic(sb, 879, 4, 512, 23, 1, 444)
    classic(sb, 879, 4, 512, 23, 1, 148)
    L.b200_set_n512_mode(1)
    classic(sb, 879, 4, 512, 23, 1, 592)
    L.b200_set_n512_mode(0)
    del sb
    # the 3_3 shape three ways: workspace kernel (mask 1), first-generation tensor-memory kernel (mask 7), default
    for (sn, sk, sN, sbl, slv, batch, mask) in ((600, 2, 1024, 12, 2, 296, 3), (1077, 1, 8192, 15, 2, 148, 1),
                                                (1077, 1, 8192, 15, 2, 148, 7), (1077, 1, 8192, 15, 2, 148, 3)):
        L.b200_set_register_kernels(mask)
        sb = gpu.CudaLweBootstrapKey.from_lwe_bootstrap_key(u64(sn * slv * (sk + 1) * (sk + 1) * sN), sn, sk, sN, sbl,
                                                            slv, "Centered", streams)
        classic(sb, sn, sk, sN, sbl, slv, batch)
        del sb
    L.b200_set_register_kernels(3)
    # ---- multi-bit: conversion, fused kernel, low-latency pair (bundle + sequential, TMA ring for l = 1)
    for (mn, mbl, mlv, g) in ((920, 22, 1, 4), (918, 15, 2, 3)):
        mb = gpu.CudaLweMultiBitBootstrapKey.from_lwe_multi_bit_bootstrap_key(
            u64(((mn // g) << g) * mlv * 4 * N), mn, 1, N, mbl, mlv, g, streams)
        multibit(mb, mn, 1, N, mbl, mlv, g, 592)   # fused
        multibit(mb, mn, 1, N, mbl, mlv, g, 32)    # low latency
        del mb
    # ---- keyswitch 2048 -> 918, 4 levels: int8 tensor cores (+ digit kernel), fp64 pipe, integer pipe, u32 output
    nin, nout, kbl, klv = 2048, 918, 4, 4
    ksk = gpu.CudaLweKeyswitchKey.from_lwe_keyswitch_key(u64(nin * klv * (nout + 1)), nin, nout, kbl, klv, streams)
    d_big = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(u64(4096, nin + 1), streams)
    d_small = gpu.CudaLweCiphertextList.new(nout, 4096, streams)
    idx = gpu.trivial_indexes(4096, streams)
    for path in (1, 2, 3):
        L.b200_set_keyswitch_path(path)
        gpu.cuda_keyswitch_lwe_ciphertext(ksk, d_big, d_small, idx, idx, True, streams)
        streams.synchronize()
    L.b200_set_keyswitch_path(0)
    d_small32 = gpu.CudaVec.new(4096 * (nout + 1) // 2 + 1, streams)
    L.cuda_keyswitch_lwe_ciphertext_vector_64_32_async(
        streams.ptr(0), 0, d_small32.as_c_ptr(), idx.as_c_ptr(), d_big.d_vec.as_c_ptr(), idx.as_c_ptr(),
        ksk.d_vec.as_c_ptr(), nin, nout, kbl, klv, 4096)
    streams.synchronize()
    # ---- stand-alone stages
    glwes = gpu.CudaGlweCiphertextList.from_glwe_ciphertext_list(u64(64 * 2 * N), 1, N, streams)
    gpu.cuda_extract_lwe_samples_from_glwe_ciphertext_list(glwes, list(range(64)), 1, streams)
    ct = gpu.CudaVec.from_cpu_async(u64(4096 * 919), streams)
    gpu.cuda_modulus_switch_ciphertext(ct, 12, streams)
    gpu.cuda_centered_modulus_switch_ciphertext(gpu.CudaVec.from_cpu_async(u64(919), streams), 918, 12, streams)
    z = gpu.CudaVec.from_cpu_async(rng.standard_normal(2048 * 64), streams)
    o = gpu.CudaVec.new(2048 * 64, streams, np_dtype=np.float64)
    gpu.forward_negacyclic_fft(z, o, 2048, 64, streams)
    streams.synchronize()
    print("launched", L.b200_kernel_launch_count(), "kernels")


if __name__ == "__main__":
    main()
