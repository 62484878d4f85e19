"""Extract the reference's committed FFT16x4x16 golden spectrum into a small
fixture (tests/golden/fft16x4x16_golden_v1.npz).

Source (data only, no code): /root/reference/tfhe/src/core_crypto/gpu/algorithms/
test/fft/fft_data/fft16x4x16_golden_v1.rs -- EXPECTED_RE / EXPECTED_IM, the f64
bit patterns an H100 produced for the closed-form input of
test/fft/mod.rs:51-71 (forward negacyclic FFT, N = 2048, natural frequency
order).  Run in the build container only: /root/reference is absent on the
GPU box.
"""
import os
import re
import sys

import numpy as np

SRC = "/root/reference/tfhe/src/core_crypto/gpu/algorithms/test/fft/fft_data/fft16x4x16_golden_v1.rs"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fft16x4x16_golden_v1.npz")


def main():
    text = open(SRC).read()
    blocks = re.findall(r"pub const (EXPECTED_RE|EXPECTED_IM)[^=]*=\s*\[(.*?)\];", text, re.S)
    arrays = {}
    for name, body in blocks:
        vals = [int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]+", body)]
        arrays[name] = np.array(vals, dtype=np.uint64)
    assert arrays["EXPECTED_RE"].size == 1024 and arrays["EXPECTED_IM"].size == 1024
    np.savez_compressed(DST, expected_re_bits=arrays["EXPECTED_RE"], expected_im_bits=arrays["EXPECTED_IM"])
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    sys.exit(main())
