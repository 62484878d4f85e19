#!/usr/bin/env python
"""Side measurement for BASELINE configs[3]: latency of one FheUint64 x FheUint64
(32 blocks, PARAM_MESSAGE_2_CARRY_2_KS_PBS, full KS+PBS cascade) on one GPU.
Synthetic (random) keys and ciphertexts: the launch sequence and the PBS count
do not depend on the data.  Correctness of the cascade is covered by
tests/test_integer_mul.py.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from tfhe_rs_b200 import gpu, integer, server_key

    multi_bit = "--multi-bit" in sys.argv  # PARAM_GPU_MULTI_BIT_GROUP_4_MESSAGE_2_CARRY_2
    k, N = 1, 2048
    rng = np.random.default_rng(5)
    streams = gpu.CudaStreams.new_single_gpu(0)
    if multi_bit:
        n, g = 920, 4
        h_bsk = rng.integers(0, 1 << 64, size=((n // g) << g) * 4 * N, dtype=np.uint64)
        h_ksk = rng.integers(0, 1 << 64, size=k * N * 5 * (n + 1), dtype=np.uint64)
        skey = server_key.upload_server_key(h_bsk, h_ksk, n=n, k=k, N=N, pbs_base_log=22, pbs_level=1,
                                            ks_base_log=3, ks_level=5, grouping_factor=g, centered_ms=False,
                                            streams=streams)
    else:
        n = 918
        h_bsk = rng.integers(0, 1 << 64, size=n * 4 * N, dtype=np.uint64)
        h_ksk = rng.integers(0, 1 << 64, size=k * N * 4 * (n + 1), dtype=np.uint64)
        skey = server_key.upload_server_key(h_bsk, h_ksk, n=n, k=k, N=N, pbs_base_log=23, pbs_level=1,
                                            ks_base_log=4, ks_level=4, centered_ms=True, streams=streams)
    luts = rng.integers(0, 1 << 64, size=(4, 2 * N), dtype=np.uint64)
    rsk = integer.CudaRadixServerKey(skey, luts, k, N)
    mk = lambda: integer.CudaUnsignedRadixCiphertext(
        rsk.engine.from_numpy(rng.integers(0, 1 << 64, size=(32, k * N + 1), dtype=np.uint64)))
    a, b = mk(), mk()
    rsk.unchecked_mul(a, b)
    streams.synchronize()
    reps = 3
    rsk.engine.pbs_count = 0
    t0 = time.perf_counter()
    for _ in range(reps):
        rsk.unchecked_mul(a, b)
    streams.synchronize()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"what": "FheUint64 x FheUint64 unchecked_mul, 32 blocks, 1 GPU" +
                      (", multi-bit g=4" if multi_bit else ", classic P22"), "latency_ms": dt * 1e3,
                      "pbs_per_mul": rsk.engine.pbs_count // reps, "ops_per_s": 1.0 / dt}))


if __name__ == "__main__":
    main()
