// umma_rate.cu -- microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128) on B200 as a function of N,
// operand major-ness and the number of independent TMEM accumulators the issuing thread rotates over.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu ; run: ./umma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(int N, int a_mn, int b_mn, int nacc, int reps, long long* out, int b_off_rows = 0,
                                            int b_lbo16 = 0) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s2u(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s2u(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
                           ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo = ((s2u(smem) >> 4) & 0x3FFF) | ((a_mn ? 512u : 1u) << 16);
    // b_off_rows: start the B window that many 128-byte rows into the swizzled tile (a shifted slab window);
    // b_lbo16: override the MN-group stride (in 16-byte units), e.g. 8 = groups one row apart (merged taps)
    const uint32_t b_lo = (((s2u(smem + 32768) >> 4) + (uint32_t)b_off_rows * 8) & 0x3FFF) |
                          ((b_lbo16 ? (uint32_t)b_lbo16 : (b_mn ? 512u : 1u)) << 16);
    long long t0 = clock64();
#pragma unroll 8
    for (int r = 0; r < reps; ++r) {
      const uint32_t acc = tmem + (uint32_t)(r & (nacc - 1)) * N;
      const uint32_t al = a_lo + (r & 3) * (a_mn ? 128 : 2), bl = b_lo + (r & 3) * (b_mn ? 128 : 2);
      asm volatile(
          "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(acc),
          "r"(al), "r"(bl), "r"(idesc), "r"(1u), "r"(hi)
          : "memory");
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(&bar)) : "memory");
    asm volatile(
        "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s2u(&bar))
        : "memory");
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 2048;
  printf("%4s %5s %5s %5s %12s %12s\n", "N", "a_mn", "b_mn", "nacc", "issue cyc/mma", "total cyc/mma");
  int Ns[] = {32, 64, 128, 256};
  for (int N : Ns)
    for (int mn = 0; mn < 2; ++mn)
      for (int nacc : {1, 2, 4}) {
        if (nacc * N > 512) continue;
        k<<<1, 128, 100 * 1024>>>(N, mn, mn, nacc, reps, d);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("%4d %5d %5d %5d %12.1f %12.1f\n", N, mn, mn, nacc, (double)h[0] / reps, (double)h[1] / reps);
      }
  printf("\nMN-major B windows that start off the 8-row swizzle atom (slab wgrad), a_mn = b_mn = 1\n");
  printf("%4s %9s %7s %12s\n", "N", "off_rows", "lbo16", "total cyc/mma");
  int N2[] = {64, 128, 192};
  for (int N : N2)
    for (int off : {0, 1, 5, 8, 21})
      for (int lbo : {0, 8}) {
        if (lbo == 8 && N == 64) continue;
        k<<<1, 128, 100 * 1024>>>(N, 1, 1, 1, reps, d, off, lbo);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("%4d %9d %7d %12.1f\n", N, off, lbo, (double)h[1] / reps);
      }
  printf("\nK-major A windows off the atom (slab forward), b K-major\n");
  return 0;
}
