#!/usr/bin/env python
"""Same-box A/B of two libraries that export the tfhe-cuda-backend C ABI:
this engine (default) and the reference's own CUDA backend built for sm_100
by oracle/build_ref_cuda.sh (`--lib ref`, i.e. B200_LIB_PATH=oracle/_ref/...).

Every measurement goes through the SAME harness (tfhe-rs_b200/gpu.py mirrors,
same synthetic keys and inputs, CUDA events on the launch stream, warm-up
first).  Workloads: classic PBS P22 at batch 1 / 148 / 296 / 4096, keyswitch
2048 -> 918 (4 levels), KS + PBS, multi-bit PBS g=3 (l=2) and g=4 (l=1) at
batch 1 / 148 / 4096.  Prints one JSON line per measurement and, with --out,
writes the list to a file.  Measurement infrastructure, not part of the
bench.py contract (bench.py calls it for its `reference_gpu` block).

  python tools/ab_bench.py --lib ours --out gpurun_out/ab_ours.json
  python tools/ab_bench.py --lib ref  --out gpurun_out/ab_ref.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", choices=["ours", "ref"], default="ours")
    ap.add_argument("--what", default="classic,ks,kspbs,multibit3,multibit4")
    ap.add_argument("--batches", default="1,148,296,4096")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libtfhe_cuda_backend_ref.so")
    if args.lib == "ref":
        if not os.path.exists(ref_path):
            print(json.dumps({"lib": "ref", "unavailable": "oracle/_ref/libtfhe_cuda_backend_ref.so not built "
                              "(run oracle/build_ref_cuda.sh where /root/reference exists)"}))
            return
        os.environ["B200_LIB_PATH"] = ref_path
    else:
        os.environ.pop("B200_LIB_PATH", None)

    import numpy as np
    import torch

    import tfhe_rs_b200
    from tfhe_rs_b200 import gpu

    L = tfhe_rs_b200.lib()
    streams = gpu.CudaStreams.new_single_gpu(0)
    stream = streams.streams[0]
    rng = np.random.default_rng(3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    what = set(args.what.split(","))
    batches = [int(b) for b in args.batches.split(",")]
    results = []

    def emit(d):
        d = {"lib": args.lib, **d}
        results.append(d)
        print(json.dumps(d), flush=True)

    def timed(fn, steps, flush_l2):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        with torch.cuda.stream(stream):
            fn()
            fn()
            for s, e in evs:
                if flush_l2:
                    flush.zero_()
                s.record(stream)
                fn()
                e.record(stream)
        streams.synchronize()
        ts = [s.elapsed_time(e) for s, e in evs]
        return float(np.median(ts)), float(min(ts))

    def lut_for(k, N):
        lut = np.zeros((k + 1) * N, dtype=np.uint64)
        lut[k * N:] = np.repeat(np.arange(16, dtype=np.uint64) << np.uint64(59), N // 16)
        return gpu.CudaGlweCiphertextList.from_glwe_ciphertext_list(lut, k, N, streams)

    # ---- classic PBS, P22 ---------------------------------------------------
    n, k, N, bl, lv = 918, 1, 2048, 23, 1
    bsk = None
    if what & {"classic", "kspbs"}:
        h = rng.integers(0, 1 << 64, size=n * lv * 4 * N, dtype=np.uint64)
        bsk = gpu.CudaLweBootstrapKey.from_lwe_bootstrap_key(h, n, k, N, bl, lv, "Centered", streams)
        del h
    if "classic" in what:
        d_lut = lut_for(k, N)
        for batch in batches:
            d_in = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(
                rng.integers(0, 1 << 64, size=(batch, n + 1), dtype=np.uint64), streams)
            d_out = gpu.CudaLweCiphertextList.new(k * N, batch, streams)
            idx = gpu.trivial_indexes(batch, streams)
            lidx = gpu.CudaVec.new(batch, streams)
            sc = gpu.PbsScratch(streams, n, k, N, lv, batch, centered=True)

            def run():
                L.cuda_programmable_bootstrap_64_async(
                    streams.ptr(0), 0, d_out.d_vec.as_c_ptr(), idx.as_c_ptr(), d_lut.d_vec.as_c_ptr(),
                    lidx.as_c_ptr(), d_in.d_vec.as_c_ptr(), idx.as_c_ptr(), bsk.d_vec.as_c_ptr(), sc.buf, n, k, N,
                    bl, lv, batch, 1, 0)

            ms, best = timed(run, args.steps, batch >= 1024)
            streams.synchronize()
            sc.close()
            emit({"what": "classic PBS P22 (n=918,k=1,N=2048,l=1), centered MS", "batch": batch, "ms": ms,
                  "ms_best": best, "pbs_per_s": batch / ms * 1e3})

    # ---- the other classic shortint sets (generic kernels on both sides) -------
    # V1_x_PARAM_MESSAGE_1_CARRY_1 / 3_CARRY_3 _KS_PBS_TUNIFORM_2M128
    # (tfhe/src/shortint/parameters/v1_0/classic/tuniform/p_fail_2_minus_128/ks_pbs.rs:11-21,67-77);
    # 4_4 has N = 65536, outside the CUDA backend's [256..16384]
    for tag, (sn, sk, sN, sbl, slv) in (("set11", (879, 4, 512, 23, 1)), ("set33", (1077, 1, 8192, 15, 2))):
        if tag not in what:
            continue
        h = rng.integers(0, 1 << 64, size=sn * slv * (sk + 1) * (sk + 1) * sN, dtype=np.uint64)
        sbsk = gpu.CudaLweBootstrapKey.from_lwe_bootstrap_key(h, sn, sk, sN, sbl, slv, "Centered", streams)
        del h
        d_lut = lut_for(sk, sN)
        for batch in batches:
            d_in = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(
                rng.integers(0, 1 << 64, size=(batch, sn + 1), dtype=np.uint64), streams)
            d_out = gpu.CudaLweCiphertextList.new(sk * sN, batch, streams)
            idx = gpu.trivial_indexes(batch, streams)
            lidx = gpu.CudaVec.new(batch, streams)
            sc = gpu.PbsScratch(streams, sn, sk, sN, slv, batch, centered=True)

            def run_set():
                L.cuda_programmable_bootstrap_64_async(
                    streams.ptr(0), 0, d_out.d_vec.as_c_ptr(), idx.as_c_ptr(), d_lut.d_vec.as_c_ptr(),
                    lidx.as_c_ptr(), d_in.d_vec.as_c_ptr(), idx.as_c_ptr(), sbsk.d_vec.as_c_ptr(), sc.buf, sn, sk, sN,
                    sbl, slv, batch, 1, 0)

            ms, best = timed(run_set, max(2, args.steps - 1), batch >= 1024)
            streams.synchronize()
            sc.close()
            emit({"what": "classic PBS %s (n=%d,k=%d,N=%d,l=%d,logB=%d), centered MS" %
                  ({"set11": "1_1", "set33": "3_3"}[tag], sn, sk, sN, slv, sbl), "batch": batch, "ms": ms,
                  "ms_best": best, "pbs_per_s": batch / ms * 1e3})
        del sbsk

    # ---- keyswitch and KS + PBS ----------------------------------------------
    if what & {"ks", "kspbs"}:
        nin, nout, kbl, klv = 2048, 918, 4, 4
        ksk = gpu.CudaLweKeyswitchKey.from_lwe_keyswitch_key(
            rng.integers(0, 1 << 64, size=nin * klv * (nout + 1), dtype=np.uint64), nin, nout, kbl, klv, streams)
        for batch in [b for b in batches if b in (64, 148, 4096)] or [4096]:
            d_big = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(
                rng.integers(0, 1 << 64, size=(batch, nin + 1), dtype=np.uint64), streams)
            d_small = gpu.CudaLweCiphertextList.new(nout, batch, streams)
            idx = gpu.trivial_indexes(batch, streams)

            def run_ks():
                gpu.cuda_keyswitch_lwe_ciphertext(ksk, d_big, d_small, idx, idx, True, streams)

            if "ks" in what:
                ms, best = timed(run_ks, max(args.steps, 3), batch >= 1024)
                emit({"what": "keyswitch gemm 2048->918, 4 levels x 2^4", "batch": batch, "ms": ms, "ms_best": best,
                      "ks_per_s": batch / ms * 1e3})
            if "kspbs" in what:
                d_lut = lut_for(k, N)
                d_out = gpu.CudaLweCiphertextList.new(k * N, batch, streams)
                lidx = gpu.CudaVec.new(batch, streams)
                sc = gpu.PbsScratch(streams, n, k, N, lv, batch, centered=True)

                def run_kspbs():
                    run_ks()
                    L.cuda_programmable_bootstrap_64_async(
                        streams.ptr(0), 0, d_out.d_vec.as_c_ptr(), idx.as_c_ptr(), d_lut.d_vec.as_c_ptr(),
                        lidx.as_c_ptr(), d_small.d_vec.as_c_ptr(), idx.as_c_ptr(), bsk.d_vec.as_c_ptr(), sc.buf,
                        n, k, N, bl, lv, batch, 1, 0)

                ms, best = timed(run_kspbs, args.steps, batch >= 1024)
                streams.synchronize()
                sc.close()
                emit({"what": "KS + PBS P22", "batch": batch, "ms": ms, "ms_best": best,
                      "pbs_per_s": batch / ms * 1e3})
    del bsk

    # ---- multi-bit PBS --------------------------------------------------------
    for tag, (n, k, N, bl, lv, g) in (("multibit3", (918, 1, 2048, 15, 2, 3)), ("multibit4", (920, 1, 2048, 22, 1, 4))):
        if tag not in what:
            continue
        num_ggsw = (n // g) << g
        h = rng.integers(0, 1 << 64, size=num_ggsw * lv * 4 * N, dtype=np.uint64)
        mbsk = gpu.CudaLweMultiBitBootstrapKey.from_lwe_multi_bit_bootstrap_key(h, n, k, N, bl, lv, g, streams)
        del h
        d_lut = lut_for(k, N)
        for batch in [b for b in batches if b != 296]:
            d_in = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(
                rng.integers(0, 1 << 64, size=(batch, n + 1), dtype=np.uint64), streams)
            d_out = gpu.CudaLweCiphertextList.new(k * N, batch, streams)
            idx = gpu.trivial_indexes(batch, streams)
            lidx = gpu.CudaVec.new(batch, streams)
            sc = gpu.PbsScratch(streams, n, k, N, lv, batch, centered=False, multi_bit=True)

            def run_mb():
                L.cuda_multi_bit_programmable_bootstrap_64_async(
                    streams.ptr(0), 0, d_out.d_vec.as_c_ptr(), idx.as_c_ptr(), d_lut.d_vec.as_c_ptr(),
                    lidx.as_c_ptr(), d_in.d_vec.as_c_ptr(), idx.as_c_ptr(), mbsk.d_vec.as_c_ptr(), sc.buf, n, k, N,
                    g, bl, lv, batch, 1, 0)

            ms, best = timed(run_mb, args.steps if batch < 1024 else max(2, args.steps - 1), batch >= 1024)
            streams.synchronize()
            sc.close()
            emit({"what": f"multi-bit PBS g={g} (n={n},k=1,N=2048,l={lv},logB={bl})", "batch": batch, "ms": ms,
                  "ms_best": best, "pbs_per_s": batch / ms * 1e3})
        del mbsk

    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
