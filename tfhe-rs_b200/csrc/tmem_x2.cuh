// tmem_x2.cuh -- exchange 2 of the 16 x 4 x 16 transform through TENSOR MEMORY.
//
// Exchange 2 (pass-2 layout <-> pass-3 layout, negacyclic_fft.cuh) only moves
// data between the four threads that share a pass-2 sub-problem: a 4 x 4 block
// transpose of 16-byte values inside a warp.  Through shared memory it costs
// 16 STS.128 + 16 LDS.128 per thread = 128 wavefronts per warp and direction,
// 1,024 of the ~3,000 shared-memory wavefronts of a CMUX step -- and the
// shared-memory pipe (128 B/clk/SM) is the tightest resource of the blind
// rotation.  Blackwell's tensor memory is a second on-chip
// store with its own data path: 128 lanes x 512 columns x 32 bit per SM, warp w
// of a CTA owns lanes 32 (w % 4) .. +31.  Two access shapes make it a transpose
// unit (PTX tcgen05.ld / tcgen05.st, layouts as in CUTLASS'
// cute/atom/copy_traits_sm100.hpp, SM100_TMEM_{LOAD,STORE}_*):
//   .32x32b      thread t <-> lane t, consecutive columns          (lane-private)
//   .16x256b.xK  thread t <-> lanes t/4 and t/4 + 8 (relative to a lane base
//                that is a multiple of 16), columns 8k + 2 (t%4) + {0, 1} of
//                each of the K 8-column groups: the mma accumulator fragment
// Writing with one shape and reading with the other moves each 64-bit chunk
// from lane p + 8m to thread 4p + a (or back), which is exactly exchange 2 once
// the pass-2 work is assigned as   lane p + 8m  <->  (sub-problem 8w + p,
// quarter m)   instead of round 1's  lane 4p + m.  No shared memory, no barrier
// (the instructions are warp-collective; tcgen05.wait orders st -> ld).
// Measured (tools/micro/pipes.cu, profiles/r2e_pipes.txt): a 64-word round trip
// takes one warp 346 cycles alone, the four lane quarters are independent, and a
// saturating LDS/STS stream next to it is not slowed down.
//
// Column map of one thread's 64 words (pass-2 side, lane-private):
//   value v[4 bl + a], component h (0 re, 1 im), 32-bit word e (0 lo, 1 hi)
//   -> column 16 bl + 8 h + 2 a + e
// Register map on the pass-3 side (thread 4p + a), lane base L0 in {0, 16}:
//   register 4 k + 2 s + e of the .x8 access = lane L0 + p + 8 s, column
//   8 k + 2 a + e;  with k = 2 bl + h and m = L0 / 8 + s it is value v[4 m + bl].
//
// The CPU emulator (tests/emu) compiles the same entry points against a plain
// array model of these layouts; the GPU tests (forward-transform golden through
// the C ABI, bit-identity against the shared-memory variant) pin the model to
// the hardware.
#pragma once
#include "hd.cuh"

namespace b200 {

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t tm_smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// one warp allocates `ncols` columns (power of two >= 32) for the CTA and
// publishes the base address through shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               ::"r"(tm_smem_u32(slot_in_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

#define B200_TM_R16(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
                       "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
#define B200_TM_W16(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])

// 16 consecutive columns of the thread's own lane
__device__ __forceinline__ void tm_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), B200_TM_R16(r) : "memory");
}
__device__ __forceinline__ void tm_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : B200_TM_W16(r) : "r"(taddr) : "memory");
}
// four 8-column groups (32 columns) of 16 lanes in the fragment layout
__device__ __forceinline__ void tm_st_16x256b_x4(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), B200_TM_R16(r) : "memory");
}
__device__ __forceinline__ void tm_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : B200_TM_W16(r) : "r"(taddr) : "memory");
}

__device__ __forceinline__ uint32_t dlo(double x) { return (uint32_t)__double2loint(x); }
__device__ __forceinline__ uint32_t dhi(double x) { return (uint32_t)__double2hiint(x); }
__device__ __forceinline__ double mkd(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

// `tw` = TMEM address of (first lane of this warp's quarter, first column of the
// CTA's 64-column block).

// pass-2 side, forward: v[4 bl + a] -> columns 16 bl + 8 h + 2 a + e of the own lane
__device__ __forceinline__ void x2t_store_p2(uint32_t tw, const cplx v[16]) {
#pragma unroll
  for (int bl = 0; bl < 4; bl++) {
    uint32_t r[16];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      r[2 * a] = dlo(v[4 * bl + a].re);
      r[2 * a + 1] = dhi(v[4 * bl + a].re);
      r[8 + 2 * a] = dlo(v[4 * bl + a].im);
      r[8 + 2 * a + 1] = dhi(v[4 * bl + a].im);
    }
    tm_st_32x32b_x16(tw + 16 * bl, r);
  }
  tmem_wait_st();
}
// pass-2 side, inverse: the same columns back into v[4 bl + a]
__device__ __forceinline__ void x2t_load_p2(uint32_t tw, cplx v[16]) {
  uint32_t r[4][16];
#pragma unroll
  for (int bl = 0; bl < 4; bl++)
    tm_ld_32x32b_x16(tw + 16 * bl, r[bl]);
  tmem_wait_ld();
#pragma unroll
  for (int bl = 0; bl < 4; bl++)
#pragma unroll
    for (int a = 0; a < 4; a++)
      v[4 * bl + a] = cmake(mkd(r[bl][2 * a], r[bl][2 * a + 1]),
                            mkd(r[bl][8 + 2 * a], r[bl][8 + 2 * a + 1]));
}
// pass-3 side, forward: thread 4p + a collects v[4 m + bl] from lanes p + 8 m.
// Split into issue / unpack so that a caller can put other tensor-memory loads
// under the same tcgen05.wait::ld.
__device__ __forceinline__ void x2t_load_p3_issue(uint32_t tw, uint32_t (&r)[2][2][16]) {
#pragma unroll
  for (int lh = 0; lh < 2; lh++)
#pragma unroll
    for (int ch = 0; ch < 2; ch++)
      tm_ld_16x256b_x4(tw + ((uint32_t)(16 * lh) << 16) + 32 * ch, r[lh][ch]);
}
__device__ __forceinline__ void x2t_load_p3_unpack(const uint32_t (&r)[2][2][16], cplx v[16]) {
#pragma unroll
  for (int lh = 0; lh < 2; lh++)
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
      for (int bl = 0; bl < 4; bl++) {
        // k = 2 bl + h; column half ch = k / 4, register 4 (k % 4) + 2 s + e
        const int kre = 2 * bl, kim = 2 * bl + 1;
        const uint32_t *qre = r[lh][kre >> 2], *qim = r[lh][kim >> 2];
        v[4 * (2 * lh + s) + bl] =
            cmake(mkd(qre[4 * (kre & 3) + 2 * s], qre[4 * (kre & 3) + 2 * s + 1]),
                  mkd(qim[4 * (kim & 3) + 2 * s], qim[4 * (kim & 3) + 2 * s + 1]));
      }
}
__device__ __forceinline__ void x2t_load_p3(uint32_t tw, cplx v[16]) {
  uint32_t r[2][2][16]; // [lane half L0/16][column half][register]
  x2t_load_p3_issue(tw, r);
  tmem_wait_ld();
  x2t_load_p3_unpack(r, v);
}
// 15 loop-invariant complex values (the pass-3 twiddles: 60 registers) parked in
// 64 lane-private columns and fetched back around the two passes that use them
__device__ __forceinline__ void tm_park15(uint32_t taddr, const cplx x[15]) {
#pragma unroll
  for (int q = 0; q < 4; q++) {
    uint32_t r[16];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int idx = 4 * q + e;
      const cplx c = idx < 15 ? x[idx] : cmake(0.0, 0.0);
      r[4 * e] = dlo(c.re);
      r[4 * e + 1] = dhi(c.re);
      r[4 * e + 2] = dlo(c.im);
      r[4 * e + 3] = dhi(c.im);
    }
    tm_st_32x32b_x16(taddr + 16 * q, r);
  }
  tmem_wait_st();
}
__device__ __forceinline__ void tm_fetch15_issue(uint32_t taddr, uint32_t (&r)[4][16]) {
#pragma unroll
  for (int q = 0; q < 4; q++)
    tm_ld_32x32b_x16(taddr + 16 * q, r[q]);
}
__device__ __forceinline__ void tm_fetch15_unpack(const uint32_t (&r)[4][16], cplx x[15]) {
#pragma unroll
  for (int idx = 0; idx < 15; idx++) {
    const int q = idx >> 2, e = idx & 3;
    x[idx] = cmake(mkd(r[q][4 * e], r[q][4 * e + 1]), mkd(r[q][4 * e + 2], r[q][4 * e + 3]));
  }
}
// pass-3 side, inverse: scatter v[4 m + bl] to lanes p + 8 m
__device__ __forceinline__ void x2t_store_p3(uint32_t tw, const cplx v[16]) {
#pragma unroll
  for (int lh = 0; lh < 2; lh++)
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
      uint32_t r[16];
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const int k = 4 * ch + kk, bl = k >> 1, h = k & 1;
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const cplx x = v[4 * (2 * lh + s) + bl];
          const double c = h ? x.im : x.re;
          r[4 * kk + 2 * s] = dlo(c);
          r[4 * kk + 2 * s + 1] = dhi(c);
        }
      }
      tm_st_16x256b_x4(tw + ((uint32_t)(16 * lh) << 16) + 32 * ch, r);
    }
  tmem_wait_st();
}

#endif // __CUDACC__

// ---------------------------------------------------------------------------
// Host model of the same four entry points for the CPU CTA emulator: `tm` is a
// [32 lanes][64 columns] word array per warp, `lane` the thread's lane.
// ---------------------------------------------------------------------------
#if !defined(__CUDACC__)
struct TmemWarpModel {
  uint32_t w[32][64];
};
static inline void tmm_split(double x, uint32_t &lo, uint32_t &hi) {
  uint64_t b;
  __builtin_memcpy(&b, &x, 8);
  lo = (uint32_t)b;
  hi = (uint32_t)(b >> 32);
}
static inline double tmm_join(uint32_t lo, uint32_t hi) {
  const uint64_t b = ((uint64_t)hi << 32) | lo;
  double x;
  __builtin_memcpy(&x, &b, 8);
  return x;
}
static inline void x2t_store_p2(TmemWarpModel &tm, int lane, const cplx v[16]) {
  for (int bl = 0; bl < 4; bl++)
    for (int a = 0; a < 4; a++) {
      tmm_split(v[4 * bl + a].re, tm.w[lane][16 * bl + 2 * a], tm.w[lane][16 * bl + 2 * a + 1]);
      tmm_split(v[4 * bl + a].im, tm.w[lane][16 * bl + 8 + 2 * a], tm.w[lane][16 * bl + 8 + 2 * a + 1]);
    }
}
static inline void x2t_load_p2(const TmemWarpModel &tm, int lane, cplx v[16]) {
  for (int bl = 0; bl < 4; bl++)
    for (int a = 0; a < 4; a++)
      v[4 * bl + a] = cmake(tmm_join(tm.w[lane][16 * bl + 2 * a], tm.w[lane][16 * bl + 2 * a + 1]),
                            tmm_join(tm.w[lane][16 * bl + 8 + 2 * a], tm.w[lane][16 * bl + 8 + 2 * a + 1]));
}
// .16x256b: thread `lane` <-> lanes L0 + lane/4 (+8), columns 8 k + 2 (lane%4) + e
static inline void x2t_load_p3(const TmemWarpModel &tm, int lane, cplx v[16]) {
  const int p = lane >> 2, a = lane & 3;
  for (int m = 0; m < 4; m++)
    for (int bl = 0; bl < 4; bl++) {
      const int src = p + 8 * m;
      v[4 * m + bl] = cmake(tmm_join(tm.w[src][8 * (2 * bl) + 2 * a], tm.w[src][8 * (2 * bl) + 2 * a + 1]),
                            tmm_join(tm.w[src][8 * (2 * bl + 1) + 2 * a], tm.w[src][8 * (2 * bl + 1) + 2 * a + 1]));
    }
}
static inline void x2t_store_p3(TmemWarpModel &tm, int lane, const cplx v[16]) {
  const int p = lane >> 2, a = lane & 3;
  for (int m = 0; m < 4; m++)
    for (int bl = 0; bl < 4; bl++) {
      const int dst = p + 8 * m;
      tmm_split(v[4 * m + bl].re, tm.w[dst][8 * (2 * bl) + 2 * a], tm.w[dst][8 * (2 * bl) + 2 * a + 1]);
      tmm_split(v[4 * m + bl].im, tm.w[dst][8 * (2 * bl + 1) + 2 * a], tm.w[dst][8 * (2 * bl + 1) + 2 * a + 1]);
    }
}
#endif

} // namespace b200
