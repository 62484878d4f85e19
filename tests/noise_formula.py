"""Restatement of the reference's PBS output-noise formula
`pbs_variance_132_bits_security_tuniform_fft_mul`
(tfhe/src/core_crypto/commons/noise_formulas/lwe_programmable_bootstrap.rs:86-150)
with PBS_FFT_64_MANTISSA_SIZE = 53 (noise_simulation/mod.rs:29).  Returns the
variance as a fraction of the torus squared."""
import math


def pbs_variance_tuniform_fft(n, k, N, base_log, level, mantissa=53.0, modulus=2.0 ** 64):
    B, l, ln, L2E = 2.0 ** base_log, float(level), math.log, math.log2(math.e)
    t0 = 0.0 if (mantissa - L2E * ln(modulus) >= 0) else (-mantissa + L2E * ln(modulus))
    fft = (0.00705 * 2.0 ** (2.0 * t0 + 2.88539008177793 * ln(B) - 2.88539008177793 * ln(modulus))
           * l ** 1.01827 * k ** 1.22003 * N ** 2.22003 * (k + 1.0) ** 1.01827)
    key = (l * N * (2.0 ** (4.44 - 2.88539008177793 * ln(modulus))
                    + (1 / 3.0) * modulus ** -2.0
                    * (2.0 ** (2.0 * math.ceil(-0.025167785 * k * N + L2E * ln(modulus) + 4.10067100000001)) + 0.5))
           * ((1 / 12.0) * B ** 2 + 0.166666666666667) * (k + 1.0))
    rnd = (-1 / 24.0 * modulus ** -2.0
           + 0.5 * k * N * (0.0208333333333333 * modulus ** -2.0 + 0.0416666666666667 * B ** (-2.0 * l))
           + (1 / 24.0) * B ** (-2.0 * l))
    return n * (fft + key + rnd)
