#!/usr/bin/env python
"""Helper of tests/test_ref_cuda_parity.py (test infrastructure): runs KS -> PBS
on the keys / inputs of an .npz through ONE library exporting the
tfhe-cuda-backend C ABI -- this engine (`--lib ours`) or the reference's own
CUDA backend built by oracle/build_ref_cuda.sh (`--lib ref`) -- in its own
process, and writes the keyswitched and bootstrapped ciphertexts back."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", choices=["ours", "ref"], required=True)
    ap.add_argument("--inp", required=True)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    if a.lib == "ref":
        os.environ["B200_LIB_PATH"] = os.path.join(ROOT, "oracle", "_ref", "libtfhe_cuda_backend_ref.so")
    else:
        os.environ.pop("B200_LIB_PATH", None)
    import numpy as np

    from tfhe_rs_b200 import gpu, server_key

    d = np.load(a.inp)
    n, k, N, bl, lv, kbl, klv, g, centered = (int(x) for x in d["params"])
    streams = gpu.CudaStreams.new_single_gpu(0)
    skey = server_key.upload_server_key(d["bsk"], d["ksk"], n=n, k=k, N=N, pbs_base_log=bl, pbs_level=lv,
                                        ks_base_log=kbl, ks_level=klv, grouping_factor=g,
                                        centered_ms=bool(centered), streams=streams)
    d_big = gpu.CudaLweCiphertextList.from_lwe_ciphertext_list(d["big"], streams)
    d_lut = gpu.CudaGlweCiphertextList.from_glwe_ciphertext_list(d["lut"], k, N, streams)
    d_small = skey.keyswitch(d_big)
    out = skey.bootstrap(d_small, d_lut)
    streams.synchronize()
    np.savez(a.out, small=d_small.to_lwe_ciphertext_list(streams), out=out.to_lwe_ciphertext_list(streams))


if __name__ == "__main__":
    main()
