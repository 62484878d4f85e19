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


def multi_bit_pbs_variance_tuniform_fft(n, k, N, base_log, level, grouping_factor, mantissa=53.0, modulus=2.0 ** 64):
    """`multi_bit_pbs_variance_132_bits_security_tuniform_gf_{2,3,4}_fft_mul`
    (commons/noise_formulas/lwe_multi_bit_programmable_bootstrap.rs:258-462):
    the three grouping factors differ by the FFT-term constants and the number
    of GGSWs summed into a bundle (2^g)."""
    coef, e_l, e_k, e_n = {2: (0.0022, 1.04148, 1.94548, 2.94548),
                           3: (0.00492, 1.0111, 1.90722, 2.90722),
                           4: (0.00855, 1.00715, 1.90759, 2.90759)}[grouping_factor]
    B, l, ln, L2E = 2.0 ** base_log, float(level), math.log, math.log2(math.e)
    t0 = 0.0 if (mantissa - L2E * ln(modulus) >= 0) else (-mantissa + L2E * ln(modulus))
    fft = (coef * 2.0 ** (2.0 * t0 + 2.88539008177793 * ln(B) - 2.88539008177793 * ln(modulus))
           * l ** e_l * k ** e_k * N ** e_n * (k + 1.0) ** e_l)
    key = ((2.0 ** grouping_factor) * l * N
           * (2.0 ** (4.44 - 2.88539008177793 * ln(modulus))
              + (1 / 3.0) * modulus ** -2.0
              * (2.0 ** (2.0 * math.ceil(-0.025167785 * k * N + L2E * ln(modulus) + 4.10067100000001)) + 0.5))
           * ((1 / 12.0) * B ** 2 + 0.166666666666667) * (k + 1.0))
    rnd = (-1 / 12.0 * modulus ** -2.0
           + k * N * (0.0208333333333333 * modulus ** -2.0 + 0.0416666666666667 * B ** (-2.0 * l))
           + (1 / 12.0) * B ** (-2.0 * l))
    return (1.0 / grouping_factor) * n * (fft + key + rnd)
