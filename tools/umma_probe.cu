// Probe of the kind::f16 (bf16) operand layouts the bf16 mode relies on, checked against a CPU
// product on one CTA.  Each variant prints max |err| so that ONE GPU run settles the encodings:
//   v0  SS, A K-major SW128, B K-major SW128                     (baseline, known good)
//   v1  SS, B MN-major SW128: B tile stored [k rows][64 n] (a TMA {64, rows} box), N = 64
//   v2  SS, B MN-major SW128, N = 128: two MN atoms (two boxes), LBO = box stride
//   v3  TS, A packed bf16x2 in TMEM (lane = row, column j = (A[2j] lo, A[2j+1] hi)), B K-major
//   v4  TS packed A + MN-major B (the P.V product of the attention kernel)
//   v5  as v2 but LBO / SBO swapped                              (diagnostic)
//   v6  as v3 but (hi, lo) packing order swapped                 (diagnostic)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../tailored_avsr_b200/csrc/ptx.cuh"
using namespace tavsr;

constexpr int M = 128, K = 64;

__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b),
               "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// A: [M][K] bf16 row-major (global).  Bkn: [K][N] bf16 row-major (global).  out: [M][N] fp32.
__global__ void __launch_bounds__(128, 1) probe(int variant, int N, const __nv_bfloat16* A,
                                                const __nv_bfloat16* Bkn, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = align_smem_1024(raw);
  uint8_t* sA = smem;            // 128 rows x 128 B (K = 64 bf16), SW128 K-major
  uint8_t* sB = smem + 16384;    // K-major: N rows x 128 B; MN-major: per 64-n box, K rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x;
  const bool b_mn = variant == 1 || variant == 2 || variant == 4 || variant == 5;
  const bool a_tmem = variant == 3 || variant == 4 || variant == 6;
  // A tile, K-major SW128: element (r, c) at r*128 + ((c/8) ^ (r&7))*16 + (c%8)*2
  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, c = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2) = A[i];
  }
  if (!b_mn) {
    // B K-major: row n holds B[k][n] over k (i.e. the transposed matrix), SW128
    for (int i = tid; i < N * K; i += 128) {
      const int n = i / K, k = i % K;
      *reinterpret_cast<__nv_bfloat16*>(sB + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) =
          Bkn[k * N + n];
    }
  } else {
    // B MN-major: box b = n / 64 at offset b * (K * 128); inside, row k holds n%64 contiguous, SW128
    for (int i = tid; i < N * K; i += 128) {
      const int k = i / N, n = i % N;
      const int b = n >> 6, nn = n & 63;
      *reinterpret_cast<__nv_bfloat16*>(sB + b * (K * 128) + k * 128 + (((nn >> 3) ^ (k & 7)) << 4) +
                                        (nn & 7) * 2) = Bkn[i];
    }
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (tid < 32) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = s_tmem;
  const uint32_t lane_off = static_cast<uint32_t>((tid >> 5) * 32) << 16;
  if (a_tmem) {
    // thread = row; columns [256, 256 + K/2): packed pairs
    uint32_t r[32];
    for (int j = 0; j < 32; ++j) {
      const __nv_bfloat16 lo = A[tid * K + 2 * j], hi = A[tid * K + 2 * j + 1];
      const uint32_t l = __bfloat16_as_ushort(lo), h = __bfloat16_as_ushort(hi);
      r[j] = variant == 6 ? (h | (l << 16)) : (l | (h << 16));
    }
    tmem_st32(tb + lane_off + 256, r);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid == 0) {
    uint32_t idesc = umma_idesc(UMMA_FMT_BF16, 128, N);
    if (b_mn) idesc |= kUmmaBMajorMN;
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t bd;
      if (!b_mn) {
        bd = umma_desc_kmajor_sw128(smem_u32(sB)) + 2 * ks;  // +32 B inside the swizzle row
      } else {
        const uint32_t lbo = K * 128, sbo = 1024;
        const uint32_t addr = smem_u32(sB) + ks * 2048;  // 16 k rows = two 8-row atoms
        bd = variant == 5 ? desc_mn_sw128(addr, sbo, lbo) : desc_mn_sw128(addr, lbo, sbo);
      }
      if (a_tmem) mma_f16_ts(tb, tb + 256 + 8 * ks, bd, idesc, ks ? 1u : 0u);
      else mma_f16_ss(tb, umma_desc_kmajor_sw128(smem_u32(sA)) + 2 * ks, bd, idesc, ks ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int c = 0; c < N / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(tb + lane_off + c * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[tid * N + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tb, 512);
}

int main() {
  std::vector<__nv_bfloat16> hA(M * K), hB(K * 256);
  std::vector<float> fA(M * K), fB(K * 256);
  srand(1);
  for (int i = 0; i < M * K; ++i) {
    hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    fA[i] = __bfloat162float(hA[i]);
  }
  __nv_bfloat16 *dA, *dB;
  float* dO;
  cudaMalloc(&dA, M * K * 2);
  cudaMalloc(&dB, K * 256 * 2);
  cudaMalloc(&dO, M * 256 * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int variants[7] = {0, 1, 2, 3, 4, 5, 6};
  const int Ns[7] = {64, 64, 128, 64, 64, 128, 64};
  for (int v = 0; v < 7; ++v) {
    const int N = Ns[v];
    for (int i = 0; i < K * N; ++i) {
      hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
      fB[i] = __bfloat162float(hB[i]);
    }
    cudaMemcpy(dB, hB.data(), K * N * 2, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0, M * 256 * 4);
    probe<<<1, 128, 100 * 1024>>>(variants[v], N, dA, dB, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("variant %d: CUDA error %s\n", variants[v], cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> o(M * N);
    cudaMemcpy(o.data(), dO, M * N * 4, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        for (int k = 0; k < K; ++k) acc += double(fA[m * K + k]) * fB[k * N + n];
        mx = fmax(mx, fabs(acc - o[m * N + n]));
      }
    printf("variant %d (N=%d): max |err| = %.3e  %s\n", variants[v], N, mx, mx < 1e-3 ? "PASS" : "FAIL");
  }
  return 0;
}
