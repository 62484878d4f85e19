"""Host-side helpers of the PBS path that the reference also keeps on the host
(`core_crypto::algorithms`): accumulator (LUT) generation.  numpy only; the
results are uploaded with `gpu.CudaGlweCiphertextList.from_glwe_ciphertext_list`.
"""
from __future__ import annotations

from typing import Callable, Sequence, Union

import numpy as np


def generate_programmable_bootstrap_glwe_lut(polynomial_size: int, glwe_size: int, message_modulus: int,
                                             delta: int, f: Union[Callable[[int], int], Sequence[int]]) -> np.ndarray:
    """Trivial GLWE accumulator encoding f over `message_modulus` boxes
    (tfhe/src/core_crypto/algorithms/lwe_programmable_bootstrapping/mod.rs:26-83):
    box = N / message_modulus; body[i*box .. (i+1)*box) = f(i) * delta; the
    first half box is negated and the body rotated left by half a box; the
    glwe_size - 1 mask polynomials are zero.  Returns glwe_size * N u64 words."""
    N = polynomial_size
    assert N % message_modulus == 0, "polynomial_size must be a multiple of message_modulus"
    box = N // message_modulus
    values = [int(f(i)) if callable(f) else int(f[i]) for i in range(message_modulus)]
    body = np.repeat(np.array([(v * delta) & (2 ** 64 - 1) for v in values], dtype=np.uint64), box)
    half = box // 2
    with np.errstate(over="ignore"):
        body[:half] = np.uint64(0) - body[:half]
    body = np.roll(body, -half)
    out = np.zeros(glwe_size * N, dtype=np.uint64)
    out[(glwe_size - 1) * N:] = body
    return out


def generate_many_lut_accumulator(polynomial_size: int, glwe_size: int, message_modulus: int, delta: int,
                                  fs: Sequence[Union[Callable[[int], int], Sequence[int]]]):
    """Several functions of the same (smaller) message packed in ONE accumulator
    for `num_many_lut` extraction (shortint/server_key/mod.rs `generate_many_lookup_table`,
    consumed by the PBS with lut_stride: programmable_bootstrap_classic.cuh:491-495):
    function j fills the boxes [j * message_modulus / len(fs) ...).  Returns
    (accumulator, lut_stride)."""
    m = len(fs)
    assert message_modulus % m == 0
    sub = message_modulus // m  # inputs each function may take
    table = []
    for fj in fs:
        table.extend((int(fj(i)) if callable(fj) else int(fj[i])) for i in range(sub))
    acc = generate_programmable_bootstrap_glwe_lut(polynomial_size, glwe_size, message_modulus, delta, table)
    return acc, sub * (polynomial_size // message_modulus)
