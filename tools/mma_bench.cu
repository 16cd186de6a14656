// Microbenchmark: issue rate of tcgen05.mma kind::tf32 for several shapes / operand sources.
// One CTA per SM, one thread issues `iters` MMAs on a fixed smem tile, commits, waits.
#include <cstdio>
#include <cuda_runtime.h>
#include "../tailored_avsr_b200/csrc/ptx.cuh"
using namespace tavsr;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS tf32; 1: TS tf32; 2: SS bf16
__global__ void __launch_bounds__(128, 1) bench(int mode, int N, int iters, int group, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar_final;
  __shared__ uint32_t s_tmem;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar_final, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t fmt = mode == 2 ? UMMA_FMT_BF16 : UMMA_FMT_TF32;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t(N) >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a = umma_desc_kmajor_sw128(smem_u32(smem));
    const uint64_t b = umma_desc_kmajor_sw128(smem_u32(smem + 16384));
    long long t0 = clock64();
    for (int it = 0; it < iters; it += group) {
      for (int g = 0; g < group; ++g) {
        const int k = g & 3;
        if (mode == 0) umma_ss<true>(tb, a + 2 * k, b + 2 * k, idesc, 1);
        else if (mode == 1) mma_ts(tb, tb + 256 + 8 * k, b + 2 * k, idesc, 1);
        else mma_f16(tb, a + 2 * k, b + 2 * k, idesc, 1);
      }
      umma_commit(&bar);  // intermediate commits (as the ring would do); never waited on
    }
    umma_commit(&bar_final);
    mbar_wait(&bar_final, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

// two issuing threads (warps 0 and 1): modeA for warp 0, modeB for warp 1, separate accumulators
__global__ void __launch_bounds__(128, 1) bench2(int modeA, int modeB, int N, int iters, int group, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint64_t bar_final[2];
  __shared__ uint32_t s_tmem;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 2; ++i) { mbar_init(&bar[i], 1); mbar_init(&bar_final[i], 1); } fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = s_tmem;
  const int w = threadIdx.x >> 5;
  long long t0 = clock64();
  if ((threadIdx.x & 31) == 0 && w < 2) {
    const int mode = w == 0 ? modeA : modeB;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t(N) >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a = umma_desc_kmajor_sw128(smem_u32(smem));
    const uint64_t b = umma_desc_kmajor_sw128(smem_u32(smem + 16384));
    const uint32_t d = tb + w * 128;  // N <= 128 here
    for (int it = 0; it < iters; it += group) {
      for (int g = 0; g < group; ++g) {
        const int k = g & 3;
        if (mode == 0) umma_ss<true>(d, a + 2 * k, b + 2 * k, idesc, 1);
        else mma_ts(d, tb + 256 + 8 * k, b + 2 * k, idesc, 1);
      }
      umma_commit(&bar[w]);
    }
    umma_commit(&bar_final[w]);
    mbar_wait(&bar_final[w], 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

// CTA pair (cta_group::2): leader issues M=256 MMAs
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench_pair(int mode, int N, int iters, int group, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar_final;
  __shared__ uint32_t s_tmem;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar_final, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc_2sm(&s_tmem, 512); tmem_relinquish_2sm(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tb = s_tmem;
  long long t0 = clock64();
  if (threadIdx.x == 0 && cluster_ctarank() == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t(N) >> 3) << 17) | ((256u >> 4) << 24);
    const uint64_t a = umma_desc_kmajor_sw128(smem_u32(smem));
    const uint64_t b = umma_desc_kmajor_sw128(smem_u32(smem + 16384));
    for (int it = 0; it < iters; it += group) {
      for (int g = 0; g < group; ++g) {
        const int k = g & 3;
        if (mode == 0) umma_ss_2sm<true>(tb, a + 2 * k, b + 2 * k, idesc, 1);
        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tb), "r"(tb + 256 + 8 * k), "l"(b + 2 * k),
               "r"(idesc), "r"(1u) : "memory");
      }
      umma_commit_2sm(&bar);
    }
    umma_commit_2sm(&bar_final);
  }
  if (threadIdx.x == 0) { mbar_wait(&bar_final, 0); out[blockIdx.x] = clock64() - t0; }
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc_2sm(tb, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 256 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4096;
  const char* names[3] = {"SS tf32", "TS tf32", "SS bf16"};
  for (int grid : {1, 148})
    for (int mode = 0; mode < 3; ++mode)
      for (int N : {64, 128, 256})
        for (int group : {4, 16}) {
          bench<<<grid, 128, 64 * 1024>>>(mode, N, iters, group, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[256];
          cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("grid %3d %-8s M=128 N=%3d commit-every=%2d : %7.1f cycles/MMA (err %d)\n", grid, names[mode], N,
                 group, double(mx) / iters, int(e));
        }
  cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(bench_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {64, 128})
    for (int mb : {0, 1})
      for (int group : {4, 16}) {
        bench2<<<148, 128, 64 * 1024>>>(0, mb, N, iters, group, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[256];
        cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("two issuers: warp0 SS + warp1 %s  N=%3d commit-every=%2d : %7.1f cycles per MMA-pair slot (each warp %d MMAs) (err %d)\n",
               mb ? "TS" : "SS", N, group, double(mx) / iters, iters, int(e));
      }
  for (int mode : {0, 1})
    for (int N : {128, 256})
      for (int group : {4, 16}) {
        bench_pair<<<148, 128, 64 * 1024>>>(mode, N, iters, group, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[256];
        cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("pair (cta_group::2) %s tf32 M=256 N=%3d commit-every=%2d : %7.1f cycles/MMA (err %d)\n",
               mode ? "TS" : "SS", N, group, double(mx) / iters, int(e));
      }
  return 0;
}
