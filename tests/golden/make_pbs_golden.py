"""Extract the reference's committed GPU PBS golden ciphertexts into a fixture.

Source (read-only, only present in the build container):
  /root/reference/tfhe/src/core_crypto/gpu/algorithms/test/pbs_golden/
      pbs_golden_data/pbs_golden_v1.rs
which holds, for the seed / messages of pbs_golden/mod.rs:83,103, the H100
output LWE ciphertext (2048 mask words + body) of
  * CLASSICAL_EXPECTED          PARAM_MESSAGE_2_CARRY_2_KS_PBS_TUNIFORM_2M128
  * MULTI_BIT_GROUP_4_EXPECTED  PARAM_GPU_MULTI_BIT_GROUP_4_MESSAGE_2_CARRY_2_...
for messages 1, 7, 15 under the LUT f(x) = (2x - 1) mod 16.

Only the numeric arrays are extracted (data, not code).  Run:
    python tests/golden/make_pbs_golden.py
"""
import os
import re

import numpy as np

CSPRNG_SRC = "/root/reference/tfhe-csprng/src/generators/mod.rs"  # test_vectors, :250-277 (Seed(1), 256 bytes)
SRC = ("/root/reference/tfhe/src/core_crypto/gpu/algorithms/test/pbs_golden/"
       "pbs_golden_data/pbs_golden_v1.rs")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pbs_golden_v1.npz")


def parse(text, name):
    start = text.index(f"pub const {name}")
    end = text.find("pub const", start + 10)
    block = text[start: end if end > 0 else len(text)]
    cts = []
    for m in re.finditer(r"// msg = (\d+)\s*&\[(.*?)\]", block, re.S):
        words = [int(w, 16) for w in re.findall(r"0x([0-9a-fA-F]+)", m.group(2))]
        cts.append((int(m.group(1)), np.array(words, dtype=np.uint64)))
    return cts


def main():
    text = open(SRC).read()
    out = {}
    for name, key in (("CLASSICAL_EXPECTED", "classical"), ("MULTI_BIT_GROUP_4_EXPECTED", "multi_bit_g4")):
        cts = parse(text, name)
        assert [m for m, _ in cts] == [1, 7, 15], [m for m, _ in cts]
        arr = np.stack([c for _, c in cts])
        assert arr.shape == (3, 2049), arr.shape
        out[key] = arr
    out["messages"] = np.array([1, 7, 15], dtype=np.uint64)
    out["seed"] = np.array([0x0D1CE5ED90172048, 0], dtype=np.uint64)  # (lo, hi) of GOLDEN_SEED
    ctext = open(CSPRNG_SRC).read()
    blk = ctext[ctext.index("pub fn test_vectors<"):]
    blk = blk[blk.index("EXPECTED_BYTE: [u8; N_BYTES] = [") + len("EXPECTED_BYTE: [u8; N_BYTES] = ["):]
    blk = blk[: blk.index("];")]
    kat = np.array([int(x) for x in re.findall(r"\d+", blk)], dtype=np.uint8)
    assert kat.size == 256, kat.size
    out["csprng_seed1_bytes"] = kat
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
