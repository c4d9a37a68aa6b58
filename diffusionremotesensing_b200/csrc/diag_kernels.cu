// libdrs_b200_diag.so: micro-benchmarks behind the design decisions of conv_gemm2.cu. Built as its OWN shared library
// (include/drs_b200_diag.h); nothing here is linked into libdrs_b200.so or reachable from the product path.
//   mma_rate2: `issuers` warps of one CTA per SM each issue `iters` K-blocks of nk tcgen05.mma (M = 128, N = n, K = 16)
//   into their own accumulator, on operands resident in shared memory.
//     mode 0: descriptors held in registers (hardware rate of the shape / swizzle mode)
//     mode 1: one 32-byte KB3 record per K-block read from a __grid_constant__ table (the production loop up to r1d)
//     mode 2: one packed 64-bit record per K-block (a offset | b offset | column | init)
#include <string.h>

#include <algorithm>

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
    const int b_tiles = max(1, min(9, static_cast<int>((48u * 1024u) / (static_cast<uint32_t>(n) * row_bytes))));
    // one accumulator region per issuer: 512 columns shared evenly (N = 256 therefore needs issuers <= 2)
    const uint32_t acc = tmem + static_cast<uint32_t>(w * (issuers > 2 ? 128 : 256));
    const long long t0 = clock64();
    if (elect_one()) {
      if (mode == 0) {
        uint32_t tap = 0;
        for (int i = 0; i < iters; ++i) {
          // weight tiles wrap inside the 48 KiB region behind b16 (the r1 version walked past the allocation for
          // N >= 64, which is the illegal access logged in round 1)
          const uint32_t ao = a16 + tap * (row_bytes >> 4);
          const uint32_t bo = b16 + (tap % static_cast<uint32_t>(b_tiles)) * static_cast<uint32_t>(n) * (row_bytes >> 4);
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
    const uint32_t b_tiles = std::max(1u, std::min(9u, (48u * 1024u) / (static_cast<uint32_t>(n) * row_bytes)));
    const uint32_t b_off = (t % b_tiles) * static_cast<uint32_t>(n) * (row_bytes >> 4);
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
  if (n > 256 || n < 16 || n % 16 || (n > 128 && issuers > 2)) return static_cast<int>(cudaErrorInvalidValue);
  mma_rate2_kernel<<<sms, 192, 182 * 1024>>>(n, nk, layout, sbo16, issuers, iters, mode, d, tab);
  e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out_host, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return static_cast<int>(e);
}


// ------------------------------------------------------------------------------------------------
// first form (drs_debug_mma_rate): cycles per tcgen05.mma M=128 x N x K=16 issued back to back by one elected thread
// ------------------------------------------------------------------------------------------------


struct RateTable { KB3 kb[16]; };

__global__ void __launch_bounds__(128) mma_rate_kernel(int n, int iters, int unroll4, long long* out,
                                                       const __grid_constant__ RateTable tab) {
  extern __shared__ uint8_t dyn_smem[];
  __shared__ __align__(8) uint64_t s_done;
  __shared__ uint32_t s_tmem_base;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t dyn_u32 = smem_u32(dyn_smem);
  uint8_t* const base = dyn_smem + ((1024u - (dyn_u32 & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 180 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&s_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&s_tmem_base, 256);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  if (warp == 1) {
    // unroll4 bit 0: four MMAs (K slices) per iteration; bits 8..23: A group stride in 16-byte units (default 64 =
    // 1024 B); bits 24..31: A start offset in 128-byte rows
    const uint32_t sbo16 = ((unroll4 >> 8) & 0xFFFF) ? ((unroll4 >> 8) & 0xFFFF) : 64u;
    const uint32_t row0 = (unroll4 >> 24) & 0xFF;
    const int vary = (unroll4 >> 1) & 1;  // bit 1: walk A over 9 tap offsets and B over 9 weight tiles
    unroll4 &= 1;
    const uint32_t a16 = (smem_u32(base) >> 4) + (row0 & 0x7F) * 8u, b16 = smem_u32(base + 32768) >> 4;
    const uint32_t b16b = smem_u32(base + 24576) >> 4;
    const uint32_t hi_a = sbo16 | (1u << 14) | (2u << 29);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((static_cast<uint32_t>(n) >> 3) << 17) | (8u << 24);
    const long long t0 = clock64();
    if (elect_one()) {
      uint32_t tap = 0;
      for (int i = 0; i < iters; ++i) {
        if (vary && (row0 & 0x80) && (row0 & 0x70)) {
          // lab: the convolution kernel's issue loop, one feature at a time
          //   0x10: nk-dependent branches   0x20: accumulate flag from the record   0x40: LAST-flag loop exit
          const int feat = row0 & 0x70;
          uint32_t kbi = 0;
          for (;;) {
            const KB3 K = tab.kb[kbi];
            const uint32_t ao = a16 + (K.a_lo & 0xFFFFu), bo = b16b + (K.b_lo & 0xFFFFu);
            const uint32_t d = tmem + K.col;
            const uint32_t accf = (feat & 0x20) ? ((K.flags & KB2_INIT) ? 0u : 1u) : 1u;
            umma_bf16_split(d, (ao & 0x3FFFu) | 0x10000u, K.a_hi, (bo & 0x3FFFu) | 0x10000u, K.b_hi, K.idesc, accf);
            if (feat & 0x10) {
              if (K.nk >= 2)
                umma_bf16_split(d, ((ao + 2u) & 0x3FFFu) | 0x10000u, K.a_hi, ((bo + 2u) & 0x3FFFu) | 0x10000u, K.b_hi, K.idesc, 1u);
              if (K.nk == 4) {
                umma_bf16_split(d, ((ao + 4u) & 0x3FFFu) | 0x10000u, K.a_hi, ((bo + 4u) & 0x3FFFu) | 0x10000u, K.b_hi, K.idesc, 1u);
                umma_bf16_split(d, ((ao + 6u) & 0x3FFFu) | 0x10000u, K.a_hi, ((bo + 6u) & 0x3FFFu) | 0x10000u, K.b_hi, K.idesc, 1u);
              }
            } else {
#pragma unroll
              for (int k = 1; k < 4; ++k)
                umma_bf16_split(d, ((ao + 2u * k) & 0x3FFFu) | 0x10000u, K.a_hi, ((bo + 2u * k) & 0x3FFFu) | 0x10000u,
                                K.b_hi, K.idesc, 1u);
            }
            ++kbi;
            if (feat & 0x40) {
              if (K.flags & KB2_LAST) break;
            } else if (kbi == 9u) {
              break;
            }
          }
        } else if (vary && (row0 & 0x80)) {
          // descriptors fetched from the kernel-parameter table with a dynamic index, like the convolution kernel
          const KB3 K = tab.kb[tap];
          const uint32_t ao = a16 + (K.a_lo & 0xFFFFu), bo = b16b + (K.b_lo & 0xFFFFu);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_split(tmem + K.col, ((ao + 2u * k) & 0x3FFFu) | 0x10000u, K.a_hi, ((bo + 2u * k) & 0x3FFFu) | 0x10000u,
                            K.b_hi, K.idesc, 1u);
          tap = (tap == 8u) ? 0u : tap + 1u;
        } else if (vary) {
          // same access pattern as the convolution: tap (ky, kx) of a 10-pixel-wide halo tile, its own weight tile
          const uint32_t ao = a16 + ((tap / 3u) * 10u + (tap % 3u)) * 8u, bo = b16b + tap * (static_cast<uint32_t>(n) * 8u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_split(tmem, ((ao + 2u * k) & 0x3FFFu) | 0x10000u, hi_a, ((bo + 2u * k) & 0x3FFFu) | 0x10000u, hi,
                            idesc, 1u);
          tap = (tap == 8u) ? 0u : tap + 1u;
        } else if (unroll4) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_split(tmem, ((a16 + 2u * k) & 0x3FFFu) | 0x10000u, hi_a, ((b16 + 2u * k) & 0x3FFFu) | 0x10000u, hi,
                            idesc, 1u);
        } else {
          umma_bf16_split(tmem, (a16 & 0x3FFFu) | 0x10000u, hi_a, (b16 & 0x3FFFu) | 0x10000u, hi, idesc, 1u);
        }
      }
      umma_commit(&s_done);
    }
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(&s_done, 0, nullptr, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) {
      out[0] = t1 - t0;  // issue time
      out[1] = t2 - t0;  // completion time
    }
  } else if (iters < 0) {
    (void)0;
  } else if ((n & 1) == 0 && (unroll4 & 0x4)) {
    // mode bit 2: the other three warps wait on the completion barrier exactly like epilogue warps do
    mbar_wait(&s_done, 0, nullptr, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

int mma_rate(int n, int iters, int unroll4, int ctas_per_sm, long long* out_host) {
  long long* d = nullptr;
  cudaError_t e = cudaMalloc(&d, 16);
  if (e != cudaSuccess) return static_cast<int>(e);
  cudaMemset(d, 0, 16);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  RateTable tab;
  memset(&tab, 0, sizeof(tab));
  for (uint32_t t = 0; t < 9; ++t) {
    tab.kb[t].a_lo = ((t / 3u) * 10u + (t % 3u)) * 8u;
    tab.kb[t].b_lo = t * static_cast<uint32_t>(n) * 8u;
    tab.kb[t].a_hi = 80u | (1u << 14) | (2u << 29);
    tab.kb[t].b_hi = 64u | (1u << 14) | (2u << 29);
    tab.kb[t].idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((static_cast<uint32_t>(n) >> 3) << 17) | (8u << 24);
    tab.kb[t].nk = 4;
    tab.kb[t].flags = static_cast<uint8_t>((t == 0 ? (KB2_INIT | KB2_FIRST) : 0) | (t == 8 ? KB2_LAST : 0));
  }
  mma_rate_kernel<<<sms * ctas_per_sm, 128, 182 * 1024>>>(n, iters, unroll4, d, tab);
  e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out_host, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return static_cast<int>(e);
}


}  // namespace drs

// ------------------------------------------------------------------------------------------------
// C ABI of the diagnostics library
// ------------------------------------------------------------------------------------------------
#include <stdio.h>

#include "../../include/drs_b200_diag.h"

static thread_local char g_diag_err[256] = "";

static int diag_fail(int cuda_error, const char* what) {
  snprintf(g_diag_err, sizeof(g_diag_err), "CUDA error %d (%s) in %s", cuda_error,
           cudaGetErrorString(static_cast<cudaError_t>(cuda_error)), what);
  cudaGetLastError();
  return -3;
}

extern "C" {

const char* drs_diag_last_error(void) { return g_diag_err; }

int drs_debug_mma_rate(int n, int iters, int unroll4, int ctas_per_sm, long long* out_host) {
  const int r = drs::mma_rate(n, iters, unroll4, ctas_per_sm, out_host);
  return r ? diag_fail(r, "mma_rate_kernel") : 0;
}

int drs_debug_mma_rate2(int n, int nk, int layout, int sbo16, int issuers, int iters, int mode, long long* out_host) {
  const int r = drs::mma_rate2(n, nk, layout, sbo16, issuers, iters, mode, out_host);
  return r ? diag_fail(r, "mma_rate2_kernel") : 0;
}

}  // extern "C"
