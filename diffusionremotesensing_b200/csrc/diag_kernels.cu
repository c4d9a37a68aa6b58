// Micro-benchmarks behind the design decisions of conv_gemm2.cu (debug entry points only, not on the product path).
//   mma_rate2: `issuers` warps of one CTA per SM each issue `iters` K-blocks of nk tcgen05.mma (M = 128, N = n, K = 16)
//   into their own accumulator, on operands resident in shared memory.
//     mode 0: descriptors held in registers (hardware rate of the shape / swizzle mode)
//     mode 1: one 32-byte KB3 record per K-block read from a __grid_constant__ table (the production loop up to r1d)
//     mode 2: one packed 64-bit record per K-block (a offset | b offset | column | init)
#include <string.h>

#include "conv_gemm2.cuh"
#include "ptx.cuh"

namespace drs {

struct Rate2Table {
  KB3 kb[32];
  uint64_t packed[32];
};

__global__ void __launch_bounds__(192) mma_rate2_kernel(int n, int nk, int layout, int sbo16, int issuers, int iters,
                                                        int mode, long long* out,
                                                        const __grid_constant__ Rate2Table tab) {
  extern __shared__ uint8_t dyn_smem[];
  __shared__ __align__(8) uint64_t s_done[4];
  __shared__ uint32_t s_tmem_base;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t dyn_u32 = smem_u32(dyn_smem);
  uint8_t* const base = dyn_smem + ((1024u - (dyn_u32 & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&s_done[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&s_tmem_base, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  if (warp >= 1 && warp <= issuers) {
    const int w = warp - 1;
    const uint32_t a16 = smem_u32(base + w * 32768) >> 4;       // each issuer reads its own A region
    const uint32_t b16 = smem_u32(base + 131072) >> 4;           // shared weights
    const uint32_t hi_a = static_cast<uint32_t>(sbo16) | (1u << 14) | (static_cast<uint32_t>(layout) << 29);
    const uint32_t row_bytes = (layout == 2) ? 128u : (layout == 4 ? 64u : 32u);
    const uint32_t hi_b = ((8u * row_bytes) >> 4) | (1u << 14) | (static_cast<uint32_t>(layout) << 29);
    const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(n));
    const uint32_t acc = tmem + static_cast<uint32_t>(w * 128);
    const long long t0 = clock64();
    if (elect_one()) {
      if (mode == 0) {
        uint32_t tap = 0;
        for (int i = 0; i < iters; ++i) {
          const uint32_t ao = a16 + tap * (row_bytes >> 4), bo = b16 + tap * static_cast<uint32_t>(n) * (row_bytes >> 4);
          for (int k = 0; k < nk; ++k)
            umma_bf16_split(acc, ((ao + 2u * k) & 0x3FFFu) | 0x10000u, hi_a, ((bo + 2u * k) & 0x3FFFu) | 0x10000u, hi_b,
                            idesc, 1u);
          tap = (tap == 8u) ? 0u : tap + 1u;
        }
      } else if (mode == 1) {
        for (int i = 0; i < iters; i += 9) {
          for (int t = 0; t < 9; ++t) {
            const KB3 K = tab.kb[t];
            const uint32_t a_lo = K.a_lo + a16, b_lo = K.b_lo + b16;
            const uint32_t d = acc + K.col;
            umma_bf16_split(d, a_lo, K.a_hi, b_lo, K.b_hi, K.idesc, (K.flags & KB2_INIT) ? 0u : 1u);
            for (int k = 1; k < nk; ++k) umma_bf16_split(d, a_lo + 2u * k, K.a_hi, b_lo + 2u * k, K.b_hi, K.idesc, 1u);
          }
        }
      } else {
        for (int i = 0; i < iters; i += 9) {
          for (int t = 0; t < 9; ++t) {
            const uint64_t P = tab.packed[t];
            const uint32_t lo = static_cast<uint32_t>(P), hi = static_cast<uint32_t>(P >> 32);
            const uint32_t a_lo = ((lo & 0x3FFFu) + a16) | 0x10000u;
            const uint32_t b_lo = ((hi & 0x3FFFFu) + b16) | 0x10000u;
            const uint32_t d = acc + ((lo >> 14) & 0x1FFu);
            const uint32_t accf = (lo >> 23) & 1u;
            umma_bf16_split(d, a_lo, hi_a, b_lo, hi_b, idesc, accf);
            for (int k = 1; k < nk; ++k) umma_bf16_split(d, a_lo + 2u * k, hi_a, b_lo + 2u * k, hi_b, idesc, 1u);
          }
        }
      }
      umma_commit(&s_done[w]);
    }
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(&s_done[w], 0, nullptr, 0);
    const long long t2 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0 && w == 0) {
      out[0] = t1 - t0;  // issue time
      out[1] = t2 - t0;  // completion time
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int mma_rate2(int n, int nk, int layout, int sbo16, int issuers, int iters, int mode, long long* out_host) {
  long long* d = nullptr;
  cudaError_t e = cudaMalloc(&d, 16);
  if (e != cudaSuccess) return static_cast<int>(e);
  cudaMemset(d, 0, 16);
  cudaFuncSetAttribute(mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t row_bytes = (layout == 2) ? 128u : (layout == 4 ? 64u : 32u);
  Rate2Table tab;
  memset(&tab, 0, sizeof(tab));
  for (uint32_t t = 0; t < 9; ++t) {
    const uint32_t a_off = ((t / 3u) * 10u + (t % 3u)) * (row_bytes >> 4);
    const uint32_t b_off = t * static_cast<uint32_t>(n) * (row_bytes >> 4);
    tab.kb[t].a_lo = a_off | 0x10000u;
    tab.kb[t].b_lo = b_off | 0x10000u;
    tab.kb[t].a_hi = static_cast<uint32_t>(sbo16) | (1u << 14) | (static_cast<uint32_t>(layout) << 29);
    tab.kb[t].b_hi = ((8u * row_bytes) >> 4) | (1u << 14) | (static_cast<uint32_t>(layout) << 29);
    tab.kb[t].idesc = umma_idesc_host(128, n);
    tab.kb[t].nk = static_cast<uint8_t>(nk);
    tab.kb[t].flags = 0;
    tab.packed[t] = static_cast<uint64_t>(a_off | (1u << 23)) | (static_cast<uint64_t>(b_off) << 32);
  }
  if (issuers < 1) issuers = 1;
  if (issuers > 4) issuers = 4;
  mma_rate2_kernel<<<sms, 192, 182 * 1024>>>(n, nk, layout, sbo16, issuers, iters, mode, d, tab);
  e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out_host, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return static_cast<int>(e);
}

}  // namespace drs
