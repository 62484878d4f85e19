// pbs_n8192_phases.cuh -- the parts of the (N = 8192, k = 1, l = 2) register
// kernel (pbs_n8192.cuh) that compile for host and device: sizes, the Fourier key
// layout and the rotate + decompose phase.  The CTA emulator (tests/emu) replays
// a whole PBS with these and the transform pieces of negacyclic_fft.cuh.
#pragma once
#include "pbs_multibit_n2048_phases.cuh" // digits_u32, digits2_u32

#define P8K_N 8192
#define P8K_M 4096

// Fourier key: [i][level slot][row r][column c][b < 16][t3 < 256] complex128;
// slot pos = 16 t3 + b of the (i, level slot, r, c) polynomial's spectrum
B200_HD size_t n8192_key_offset(uint32_t i, uint32_t lvl, uint32_t r, uint32_t c,
                                uint32_t b, uint32_t t3) {
  return (((((size_t)i * 2 + lvl) * 2 + r) * 2 + c) * 16 + b) * 256 + t3;
}
// pre-scale of the key spectrum: 2^-64 / 4096 * 2^32 (the inverse transform is
// unnormalised and the accumulator keeps the top 32 bits)
#define P8K_KEY_SCALE 5.684341886080801486968994140625e-14

// both levels of ct1 = acc * X^a - acc for polynomial `acc_p`: level slot 0 as
// doubles, level slot 1 packed (low half: coefficient j, high half: j + 4096)
B200_HD void n8192_load_digits2(const uint32_t *acc_p, int t,
                                                   uint32_t a, uint32_t base_log,
                                                   bool ties_even, cplx v[16],
                                                   uint32_t packed[16]) {
  const uint32_t d = a & (P8K_N - 1);
  const bool neg0 = (a >> 13) != 0u;
  const uint32_t base4 = ((uint32_t)t - d) * 4u;
  const unsigned char *accb = reinterpret_cast<const unsigned char *>(acc_p);
#pragma unroll
  for (int j1 = 0; j1 < 16; j1++) {
    const uint32_t j = 256u * j1 + (uint32_t)t;
    const uint32_t ub0 = base4 + 1024u * j1;
    const uint32_t ub1 = ub0 + 4u * P8K_M;
    const uint32_t ib0 = ub0 & (4u * P8K_N - 4u);
    const uint32_t r0 = *reinterpret_cast<const uint32_t *>(accb + ib0);
    const uint32_t r1 =
        *reinterpret_cast<const uint32_t *>(accb + (ib0 ^ (4u * P8K_M)));
    const bool n0 = ((int32_t)ub0 < 0) != neg0;
    const bool n1 = ((int32_t)ub1 < 0) != neg0;
    const uint32_t x0 = (n0 ? 0u - r0 : r0) - acc_p[j];
    const uint32_t x1 = (n1 ? 0u - r1 : r1) - acc_p[j + P8K_M];
    int32_t d0[2], d1[2];
    digits2_u32(x0, base_log, d0, ties_even);
    digits2_u32(x1, base_log, d1, ties_even);
    v[j1] = cmake(int_to_double(d0[0]), int_to_double(d1[0]));
    packed[j1] = ((uint32_t)d0[1] & 0xFFFFu) | ((uint32_t)d1[1] << 16); // |digit| <= 2^14
  }
}
B200_HD void n8192_unpack_digits(const uint32_t packed[16], cplx v[16]) {
#pragma unroll
  for (int j1 = 0; j1 < 16; j1++)
    v[j1] = cmake(int_to_double((int32_t)(packed[j1] << 16) >> 16),
                  int_to_double((int32_t)packed[j1] >> 16));
}

