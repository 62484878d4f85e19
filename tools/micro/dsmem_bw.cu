// tools/micro/dsmem_bw.cu -- how fast can two CTAs of a cluster exchange a
// 16 KiB spectrum through distributed shared memory on B200?  Three ways:
//   (a) every thread stores 128-bit words to the peer (st.shared::cluster)
//   (b) every thread loads 128-bit words from the peer (ld.shared::cluster)
//   (c) one thread issues a bulk async copy smem -> peer smem
//       (cp.async.bulk.shared::cluster.shared::cta, mbarrier complete_tx)
// both CTAs act at the same time (bidirectional), REPS rounds, cycles per round
// from clock64.  Measurement tool only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dsmem_bw dsmem_bw.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;

constexpr int BYTES = 16384;
constexpr int REPS = 200;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
k(unsigned long long *out, int mode) {
  __shared__ __align__(128) unsigned char src[BYTES];
  __shared__ __align__(128) unsigned char dst[BYTES];
  __shared__ __align__(8) unsigned long long bar;
  cg::cluster_group cl = cg::this_cluster();
  const uint32_t rank = cl.block_rank(), peer = rank ^ 1;
  const int tid = threadIdx.x;
  for (int i = tid; i < BYTES / 4; i += 128) ((uint32_t *)src)[i] = i * 7 + rank;
  if (tid == 0) mbar_init(smem_u32(&bar), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  cl.sync();
  const uint32_t peer_dst = mapa(smem_u32(dst), peer);
  const uint32_t peer_src = mapa(smem_u32(src), peer);
  const uint32_t peer_bar = mapa(smem_u32(&bar), peer);
  long long t0 = clock64();
  uint4 sink = make_uint4(0, 0, 0, 0);
  for (int r = 0; r < REPS; r++) {
    if (mode == 0) {
      for (int i = tid; i < BYTES / 16; i += 128) {
        const uint4 v = ((const uint4 *)src)[i];
        asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(peer_dst + i * 16), "r"(v.x),
                     "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
      cl.sync();
    } else if (mode == 1) {
      for (int i = tid; i < BYTES / 16; i += 128) {
        uint4 v;
        asm volatile("ld.shared::cluster.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(peer_src + i * 16) : "memory");
        sink.x ^= v.x; sink.y ^= v.y; sink.z ^= v.z; sink.w ^= v.w;
      }
      cl.sync();
    } else {
      if (tid == 0) {
        mbar_expect_tx(smem_u32(&bar), BYTES); // my own barrier: the peer's copy lands here
      }
      cl.sync(); // both barriers armed (cost included; mode 3 measures it alone)
      if (tid == 0) {
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(peer_dst), "r"(smem_u32(src)), "r"(BYTES), "r"(peer_bar) : "memory");
      }
      mbar_wait(smem_u32(&bar), r & 1);
    }
  }
  long long t1 = clock64();
  if (mode == 3) { // cluster barrier alone
    t0 = clock64();
    for (int r = 0; r < REPS; r++) cl.sync();
    t1 = clock64();
  }
  if (tid == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  if (sink.x == 0x12345u) out[1] = ((uint32_t *)dst)[tid];
  cl.sync();
}

int main() {
  unsigned long long *out, h[2];
  cudaMalloc(&out, 16);
  const char *names[4] = {"st.shared::cluster 128-bit, all threads (+cluster.sync)",
                          "ld.shared::cluster 128-bit, all threads (+cluster.sync)",
                          "cp.async.bulk smem->peer smem (+cluster.sync to arm, mbarrier wait)",
                          "cluster.sync alone"};
  for (int mode = 0; mode < 4; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      k<<<2, 128>>>(out, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    const double cyc = (double)h[0] / REPS;
    printf("%-72s %8.0f cycles/round  %6.1f B/cycle per direction\n", names[mode], cyc, mode < 3 ? BYTES / cyc : 0.0);
  }
  return 0;
}
