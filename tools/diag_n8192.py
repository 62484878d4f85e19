#!/usr/bin/env python
"""Diagnostic for csrc/pbs_n8192.cuh: are the two generations of the register kernel deterministic run to run, and
where do their outputs differ?  (development tool; needs a GPU)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle  # noqa: E402
import test_gpu_parity as T  # noqa: E402

G = T.G.__wrapped__() if hasattr(T.G, "__wrapped__") else None
if G is None:
    import torch
    import tfhe_rs_b200
    from tfhe_rs_b200 import gpu, server_key
    G = type("G", (), dict(gpu=gpu, sk=server_key, lib=tfhe_rs_b200.lib(), streams=gpu.CudaStreams.new_single_gpu(0),
                           torch=torch))

for n in (1, 2, 40):
    P = oracle.Params("DIAG_N8192_n%d" % n, n=n, k=1, N=8192, pbs_base_log=15, pbs_level=2, ks_base_log=4, ks_level=5,
                      lwe_noise_log2=40, glwe_noise_log2=3, message_bits=3, carry_bits=3)
    keys = oracle.keygen(P, 77 + n, with_ksk=False)
    count = 5
    msgs = (np.arange(count) * 5 + 2) % P.p
    small = oracle.lwe_encrypt_batch(oracle.Rng(3), keys.lwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta),
                                     P.lwe_noise_log2)
    lut = oracle.make_lut(P, [(5 * i + 3) % P.p for i in range(P.p)])
    outs = {}
    for name, mask in (("g2a", 3), ("g2b", 3), ("g1a", 7), ("g1b", 7), ("ws", 1)):
        G.lib.b200_set_register_kernels(mask)
        outs[name] = T._gpu_pbs(G, T._upload(G, keys), lut, small)
    G.lib.b200_set_register_kernels(3)
    ph = {k: oracle.lwe_decrypt_batch(keys.glwe_sk, o) for k, o in outs.items()}
    ref = oracle.lwe_decrypt_batch(keys.glwe_sk, oracle.pbs_batch(keys, lut, small))
    print("n =", n, "a_hat-free facts:")
    for a, b in (("g2a", "g2b"), ("g1a", "g1b"), ("g2a", "g1a")):
        ne = outs[a] != outs[b]
        print("   %s vs %s: identical=%s differing words=%d of %d; rows %s; first cols %s" % (
            a, b, not ne.any(), int(ne.sum()), ne.size, np.unique(np.nonzero(ne)[0])[:8], np.nonzero(ne)[1][:12]))
        if ne.any():
            d = (outs[a] - outs[b]).astype(np.int64)
            print("      word diff log2 max %.1f" % np.log2(np.abs(d).max() + 1.0))
    for k in ("g2a", "g1a", "ws"):
        d = (ph[k] - ref).astype(np.int64).astype(np.float64)
        print("   phase %s - oracle: log2 max |diff| = %.1f" % (k, np.log2(np.abs(d).max() + 1.0)))
