// CUDA-core kernels around the tensor-core convolutions: condition encoder, conv0, time tables, DDPM update,
// aggregation blend. Reference lines cited per kernel are relative to the reference checkout (see DESIGN.md for the map).
#include <stdlib.h>
#include <string.h>
#include "small_kernels.cuh"

#include <cuda_bf16.h>
#include <math.h>

namespace drs {

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args... args);

static inline int cdiv(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// Tiny-channel 3x3 conv (RRDB condition encoder, UNet_model_superres.py:230-260,345,353)
// ------------------------------------------------------------------------------------------------
__global__ void conv3x3_small_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                     const float* __restrict__ bias, const float* __restrict__ residual,
                                     float* __restrict__ out, int B, int Cin, int Cout, int H, int W, int relu,
                                     int nhwc_out) {
  __shared__ float sw[16 * 4 * 9];
  __shared__ float sb[16];
  for (int i = threadIdx.x; i < Cout * Cin * 9; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const long long total = static_cast<long long>(B) * H * W;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % W);
  const int y = static_cast<int>((idx / W) % H);
  const int b = static_cast<int>(idx / (static_cast<long long>(W) * H));
  float acc[16];
#pragma unroll
  for (int co = 0; co < 16; ++co) acc[co] = (co < Cout) ? sb[co] : 0.f;
  for (int ci = 0; ci < Cin; ++ci) {
    const float* ip = in + (static_cast<size_t>(b) * Cin + ci) * H * W;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const float v = __ldg(ip + static_cast<size_t>(yy) * W + xx);
#pragma unroll
        for (int co = 0; co < 16; ++co)
          if (co < Cout) acc[co] = fmaf(v, sw[(co * Cin + ci) * 9 + ky * 3 + kx], acc[co]);
      }
    }
  }
#pragma unroll
  for (int co = 0; co < 16; ++co) {
    if (co >= Cout) break;
    float v = acc[co];
    if (relu) v = fmaxf(v, 0.f);
    if (residual) v += residual[((static_cast<size_t>(b) * Cout + co) * H + y) * W + x];
    if (nhwc_out)
      out[((static_cast<size_t>(b) * H + y) * W + x) * Cout + co] = v;
    else
      out[((static_cast<size_t>(b) * Cout + co) * H + y) * W + x] = v;
  }
}

int launch_conv3x3_small(const float* in, const float* w, const float* bias, const float* residual, float* out, int B,
                         int Cin, int Cout, int H, int W, int relu, int nhwc_out, cudaStream_t s) {
  if (Cin > 4 || Cout > 16) return static_cast<int>(cudaErrorInvalidValue);
  const long long total = static_cast<long long>(B) * H * W;
  conv3x3_small_kernel<<<cdiv(total, 256), 256, 0, s>>>(in, w, bias, residual, out, B, Cin, Cout, H, W, relu,
                                                        nhwc_out);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Bicubic x k upsample, ATen semantics (UNet_model_superres.py:348-351: F.interpolate(..., mode='bicubic'))
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }
__device__ __forceinline__ void cubic_coeffs(float t, float* c) {
  const float A = -0.75f;
  c[0] = cubic2(t + 1.f, A);
  c[1] = cubic1(t, A);
  const float u = 1.f - t;
  c[2] = cubic1(u, A);
  c[3] = cubic2(u + 1.f, A);
}

__global__ void bicubic_up_kernel(const float* __restrict__ in, float* __restrict__ out, int BC, int H, int W, int OH,
                                  int OW, float scale) {
  const long long total = static_cast<long long>(BC) * OH * OW;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = static_cast<int>(idx % OW);
  const int oy = static_cast<int>((idx / OW) % OH);
  const int bc = static_cast<int>(idx / (static_cast<long long>(OW) * OH));
  const float rx = scale * (ox + 0.5f) - 0.5f;
  const float ry = scale * (oy + 0.5f) - 0.5f;
  const float fx = floorf(rx), fy = floorf(ry);
  const int ix = static_cast<int>(fx), iy = static_cast<int>(fy);
  float cx[4], cy[4];
  cubic_coeffs(rx - fx, cx);
  cubic_coeffs(ry - fy, cy);
  const float* ip = in + static_cast<size_t>(bc) * H * W;
  float r = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int yy = min(max(iy - 1 + j, 0), H - 1);
    float rowv = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int xx = min(max(ix - 1 + i, 0), W - 1);
      rowv += __ldg(ip + static_cast<size_t>(yy) * W + xx) * cx[i];
    }
    r += rowv * cy[j];
  }
  out[idx] = r;
}

int launch_bicubic_up(const float* in, float* out, int B, int C, int H, int W, int k, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * C * H * k * W * k;
  const float scale = static_cast<float>(1.0 / static_cast<double>(k));
  bicubic_up_kernel<<<cdiv(total, 256), 256, 0, s>>>(in, out, B * C, H, W, H * k, W * k, scale);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// conv0 + condition add -> bf16 NHWC (UNet_model_superres.py:342,355)
// ------------------------------------------------------------------------------------------------
// The 16 x Cx x 3 x 3 weights and the bias travel as a kernel parameter (constant bank), transposed to
// [ci][ky][kx][co]: the FMAs take them as constant operands, so the load/store unit only sees the activations.
// One warp = 128 consecutive pixels, lane l owns pixels l, l + 32, l + 64, l + 96 (every global access of the warp
// is lane-contiguous); 64 accumulators per thread, updated with packed fp32x2 FMAs (two independent IEEE fp32 FMAs
// per instruction: same rounding as fmaf, taps summed in the (ci, ky, kx) order of a direct convolution on top of
// bias + condition feature).
// The condition feature is channel-planar fp32 [ncond, 16, S, S].
struct Conv0Weights {
  float w[4 * 9 * 16];
  float b[16];
};

__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long v, unsigned long long w) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(v), "l"(w));
}
__device__ __forceinline__ unsigned long long pack2f(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2f(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}

template <int CX>
__global__ void __launch_bounds__(128, 5)
conv0_kernel(const float* __restrict__ x, const __grid_constant__ Conv0Weights cw, const float* __restrict__ cond,
             __nv_bfloat16* __restrict__ out, int nb, int nx, int ncond, int S) {
  // the warp's 128 pixels x 16 channels of bf16 output (4 KiB, contiguous in global memory) are transposed through
  // shared memory so that every global store instruction writes 512 contiguous bytes
  __shared__ __align__(16) uint4 s_out[4][256];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int S4 = S >> 2;
  const long long total = static_cast<long long>(nb) * S * S4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wbase = (static_cast<long long>(blockIdx.x) * 4 + warp) * 32;  // first pixel quad of this warp
  const long long idx = wbase + lane;
  const bool ok = idx < total;
  const long long i2 = ok ? idx : 0;
  const int px = static_cast<int>(i2 % S4) * 4;
  const int py = static_cast<int>((i2 / S4) % S);
  const int b = static_cast<int>(i2 / (static_cast<long long>(S4) * S));
  const size_t plane = static_cast<size_t>(S) * S;
  // The accumulators start from bias + condition feature, so the 16 condition loads are in flight together with the
  // activation loads (one memory-latency phase per thread instead of two; the sum order is (b + cond) + taps).
  unsigned long long acc[4][8];  // [pixel][channel pair]
  if (cond && ok) {
    const float* cp = cond + static_cast<size_t>(b % ncond) * 16 * plane + static_cast<size_t>(py) * S + px;
#pragma unroll
    for (int c2 = 0; c2 < 8; ++c2) {
      const float4 lo = __ldg(reinterpret_cast<const float4*>(cp + (2 * c2) * plane));
      const float4 hi = __ldg(reinterpret_cast<const float4*>(cp + (2 * c2 + 1) * plane));
      const float b0 = cw.b[2 * c2], b1 = cw.b[2 * c2 + 1];
      acc[0][c2] = pack2f(b0 + lo.x, b1 + hi.x);
      acc[1][c2] = pack2f(b0 + lo.y, b1 + hi.y);
      acc[2][c2] = pack2f(b0 + lo.z, b1 + hi.z);
      acc[3][c2] = pack2f(b0 + lo.w, b1 + hi.w);
    }
  } else {
#pragma unroll
    for (int c2 = 0; c2 < 8; ++c2) {
      const unsigned long long b2 = pack2f(cw.b[2 * c2], cw.b[2 * c2 + 1]);
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[p][c2] = b2;
    }
  }
  const float* xin = x + static_cast<size_t>(b % nx) * CX * plane;
#pragma unroll
  for (int ci = 0; ci < CX; ++ci) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = py + ky - 1;
      float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (ok && yy >= 0 && yy < S) {
        const float* row = xin + ci * plane + static_cast<size_t>(yy) * S + px;
        const float4 m = __ldg(reinterpret_cast<const float4*>(row));
        v[1] = m.x; v[2] = m.y; v[3] = m.z; v[4] = m.w;
        if (px > 0) v[0] = __ldg(row - 1);
        if (px + 4 < S) v[5] = __ldg(row + 4);
      }
      unsigned long long vv[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) vv[i] = pack2f(v[i], v[i]);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* wt = cw.w + ((ci * 3 + ky) * 3 + kx) * 16;
#pragma unroll
        for (int c2 = 0; c2 < 8; ++c2) {
          const unsigned long long w2 = pack2f(wt[2 * c2], wt[2 * c2 + 1]);
#pragma unroll
          for (int p = 0; p < 4; ++p) ffma2(acc[p][c2], vv[p + kx], w2);
        }
      }
    }
  }
  float a[4][16];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c2 = 0; c2 < 8; ++c2) unpack2f(acc[p][c2], a[p][2 * c2], a[p][2 * c2 + 1]);
  // thread t owns bytes [128 t, 128 t + 128) of the warp's output: eight 16-byte chunks, XOR-swizzled by t & 7
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(a[p][2 * i], a[p][2 * i + 1]);
      pk[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    s_out[warp][lane * 8 + ((2 * p) ^ (lane & 7))] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    s_out[warp][lane * 8 + ((2 * p + 1) ^ (lane & 7))] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
  __syncwarp();
  uint4* o = reinterpret_cast<uint4*>(out) + wbase * 8;  // 8 chunks per pixel quad
  const long long chunks_left = (total - wbase) * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int t = 4 * k + (lane >> 3), c = lane & 7;
    if (k * 32 + lane < chunks_left) o[k * 32 + lane] = s_out[warp][t * 8 + (c ^ (t & 7))];
  }
}

// w_host / bias_host: the fp32 parameters on the HOST ([16][Cx][3][3], [16]); cond: planar [ncond, 16, S, S] or null
int launch_conv0(const float* x, const float* w_host, const float* bias_host, const float* cond, void* out, int nb,
                 int nx, int ncond, int Cx, int S, cudaStream_t s) {
  if (Cx < 1 || Cx > 4) return static_cast<int>(cudaErrorInvalidValue);
  Conv0Weights cw;
  memset(&cw, 0, sizeof(cw));
  for (int co = 0; co < 16; ++co) {
    for (int r = 0; r < Cx * 9; ++r) cw.w[r * 16 + co] = w_host[co * Cx * 9 + r];
    cw.b[co] = bias_host[co];
  }
  if (S % 4) return static_cast<int>(cudaErrorInvalidValue);
  const long long total = static_cast<long long>(nb) * S * S;
  const unsigned grid = static_cast<unsigned>(cdiv(total, 512));
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  cudaError_t e;
  switch (Cx) {
    case 1: e = launch_pdl(conv0_kernel<1>, dim3(grid), dim3(128), s, x, cw, cond, o, nb, nx, ncond, S); break;
    case 2: e = launch_pdl(conv0_kernel<2>, dim3(grid), dim3(128), s, x, cw, cond, o, nb, nx, ncond, S); break;
    case 3: e = launch_pdl(conv0_kernel<3>, dim3(grid), dim3(128), s, x, cw, cond, o, nb, nx, ncond, S); break;
    default: e = launch_pdl(conv0_kernel<4>, dim3(grid), dim3(128), s, x, cw, cond, o, nb, nx, ncond, S); break;
  }
  if (e != cudaSuccess) return static_cast<int>(e);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Sinusoidal encoding [sin(t f_j) | cos(t f_j)], j < 50 (UNet_model_superres.py:328-335) + label embedding add
// (generate_new_imgs/UNet_model_generation.py:300-301)
// ------------------------------------------------------------------------------------------------
__global__ void pos_encoding_kernel(const float* __restrict__ t, const int* __restrict__ label,
                                    const float* __restrict__ emb, const float* __restrict__ inv_freq,
                                    float* __restrict__ out, int R) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * 100) return;
  const int r = idx / 100, j = idx % 100;
  const float a = t[r] * inv_freq[j % 50];
  float v = (j < 50) ? sinf(a) : cosf(a);
  if (label && emb) {
    const int l = label[r];
    if (l >= 0) v += emb[l * 100 + j];
  }
  out[idx] = v;
}

int launch_pos_encoding(const float* t, const int* label, const float* emb, const float* inv_freq, float* out, int R,
                        cudaStream_t s) {
  pos_encoding_kernel<<<cdiv(static_cast<long long>(R) * 100, 256), 256, 0, s>>>(t, label, emb, inv_freq, out, R);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Small fp32 GEMM for the time MLPs (Linear -> SiLU -> Linear -> ReLU, UNet_model_superres.py:143-151,161)
// ------------------------------------------------------------------------------------------------
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;
__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N, int K, int act) {
  __shared__ float sA[SG_BK][SG_BM + 1];
  __shared__ float sB[SG_BK][SG_BN + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    for (int i = threadIdx.x; i < SG_BM * SG_BK; i += 256) {
      const int r = i / SG_BK, k = i % SG_BK;
      const int gm = m0 + r, gk = k0 + k;
      sA[k][r] = (gm < M && gk < K) ? A[static_cast<size_t>(gm) * lda + gk] : 0.f;
      const int gn = n0 + r;
      sB[k][r] = (gn < N && gk < K) ? B[static_cast<size_t>(gn) * ldb + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j] + (bias ? bias[gn] : 0.f);
      if (act == 1) v = v / (1.f + expf(-v));
      if (act == 2) v = fmaxf(v, 0.f);
      C[static_cast<size_t>(gm) * ldc + gn] = v;
    }
  }
}

int launch_sgemm_nt(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M,
                    int N, int K, int act, cudaStream_t s) {
  dim3 grid(cdiv(N, SG_BN), cdiv(M, SG_BM));
  sgemm_nt_kernel<<<grid, 256, 0, s>>>(A, lda, B, ldb, bias, C, ldc, M, N, K, act);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// DDPM posterior update (train_diffusion_superres.py:240-249) and CFG lerp
// (generate_new_imgs/train_diffusion_generation.py:239-242)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ddpm_one(float x, float eps, float z, float c1, float c2, float c3, bool has_z) {
  // 1/sqrt(a) * (x - ((1-a)/sqrt(1-ah)) * eps) + sqrt(b) * z  -- each op rounded like the reference (no FMA)
  float r = __fmul_rn(c1, __fsub_rn(x, __fmul_rn(c2, eps)));
  if (has_z) r = __fadd_rn(r, __fmul_rn(c3, z));
  return r;
}
__device__ __forceinline__ float lerp_aten(float u, float c, float w) {
  // at::lerp(start=u, end=c, weight=w): w < 0.5 ? u + w*(c-u) : c - (c-u)*(1-w)
  const float d = __fsub_rn(c, u);
  return (fabsf(w) < 0.5f) ? __fadd_rn(u, __fmul_rn(w, d)) : __fsub_rn(c, __fmul_rn(d, __fsub_rn(1.f, w)));
}

// trow / counter (optional): sampler bookkeeping folded into this launch. Every block reads *step before it does
// anything else and bumps `counter` when it is done, so the block that brings the count to gridDim.x knows that nobody
// still needs the old value: it decrements the step index and the time-table rows and re-arms the counter.
// Each thread owns kUpdU float4 quads a block-stride apart and issues all of its loads before the first use, so one
// round trip to HBM covers the whole kernel (the r1 form ran four dependent round trips per block: step, coefficient,
// data, fence + atomic). `step` was written at least one whole launch earlier, so it is read BEFORE
// griddepcontrol.wait and overlaps the tail of the preceding convolution.
constexpr int kUpdU = 4;
template <bool CFG>
__global__ void __launch_bounds__(256, CFG ? 3 : 4)
ddpm_update_kernel(float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                   const float* __restrict__ coef, int* step, size_t n4, float cfg_scale, int* trow,
                   int n_rows, int row_dec, int* counter) {
  constexpr bool cfg = CFG;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int step_now = *reinterpret_cast<volatile int*>(step);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const bool has_z = (noise != nullptr);
  const size_t base = static_cast<size_t>(blockIdx.x) * (256 * kUpdU) + threadIdx.x;
  float4 xv[kUpdU], ev[kUpdU], uv[CFG ? kUpdU : 1], zv[kUpdU];
#pragma unroll
  for (int u = 0; u < kUpdU; ++u) {
    const size_t i = base + static_cast<size_t>(u) * 256;
    xv[u] = ev[u] = zv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cfg) uv[CFG ? u : 0] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
      xv[u] = reinterpret_cast<const float4*>(x)[i];
      ev[u] = __ldcs(reinterpret_cast<const float4*>(eps) + i);
      if (cfg) uv[CFG ? u : 0] = __ldcs(reinterpret_cast<const float4*>(eps) + n4 + i);
      if (has_z) zv[u] = __ldcs(reinterpret_cast<const float4*>(noise) + i);
    }
  }
  // the coefficient row is fetched behind the data loads (its latency hides under theirs)
  const float4 cf = __ldg(reinterpret_cast<const float4*>(coef) + step_now);
#pragma unroll
  for (int u = 0; u < kUpdU; ++u) {
    const size_t i = base + static_cast<size_t>(u) * 256;
    if (i >= n4) continue;
    float4 e4 = ev[u];
    if (cfg) {
      const float4 u4 = uv[CFG ? u : 0];
      e4.x = lerp_aten(u4.x, e4.x, cfg_scale);
      e4.y = lerp_aten(u4.y, e4.y, cfg_scale);
      e4.z = lerp_aten(u4.z, e4.z, cfg_scale);
      e4.w = lerp_aten(u4.w, e4.w, cfg_scale);
    }
    float4 o;
    o.x = ddpm_one(xv[u].x, e4.x, zv[u].x, cf.x, cf.y, cf.z, has_z);
    o.y = ddpm_one(xv[u].y, e4.y, zv[u].y, cf.x, cf.y, cf.z, has_z);
    o.z = ddpm_one(xv[u].z, e4.z, zv[u].z, cf.x, cf.y, cf.z, has_z);
    o.w = ddpm_one(xv[u].w, e4.w, zv[u].w, cf.x, cf.y, cf.z, has_z);
    reinterpret_cast<float4*>(x)[i] = o;
  }
  if (trow) {
    // No fence: the only cross-block hazard is a block still reading the OLD *step after the last block wrote the new
    // one, and every block's read completed (its value fed the coefficient load consumed above) before its atomic.
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      const int done = atomicAdd(counter, 1);
      s_last = (done == static_cast<int>(gridDim.x) - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      for (int i = threadIdx.x; i < n_rows; i += blockDim.x) trow[i] -= row_dec;
      if (threadIdx.x == 0) {
        *step -= 1;
        *counter = 0;
      }
    }
  }
}

// Launch helper: programmatic dependent launch (the kernel's griddepcontrol.wait orders it after its predecessor).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args... args) {
  static const bool no_pdl = (getenv("DRS_V2_NO_PDL") != nullptr);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int launch_ddpm_update(float* x, const float* eps, const float* noise, const float* coef, int* step, size_t numel,
                       int cfg, float cfg_scale, int* trow, int n_rows, int row_dec, int* counter, cudaStream_t s) {
  if (numel % 4) return static_cast<int>(cudaErrorInvalidValue);
  const size_t n4 = numel / 4;
  const int blocks = cdiv(static_cast<long long>(n4), 256 * kUpdU);
  const cudaError_t e =
      cfg ? launch_pdl(ddpm_update_kernel<true>, dim3(blocks), dim3(256), s, x, eps, noise, coef, step, n4, cfg_scale,
                       trow, n_rows, row_dec, counter)
          : launch_pdl(ddpm_update_kernel<false>, dim3(blocks), dim3(256), s, x, eps, noise, coef, step, n4, cfg_scale,
                       trow, n_rows, row_dec, counter);
  if (e != cudaSuccess) return static_cast<int>(e);
  return static_cast<int>(cudaGetLastError());
}

// Forward noising (train_diffusion_superres.py:183-190): x_t = sqrt(ah[t_b]) * x + sqrt(1 - ah[t_b]) * eps with the
// per-sample factors given as arrays; multiply and add rounded separately like the reference expression.
__global__ void noise_images_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                    const float* __restrict__ sa, const float* __restrict__ sb, float* __restrict__ out,
                                    size_t per_sample4, size_t n4) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const size_t b = i / per_sample4;
    const float a = __ldg(sa + b), c = __ldg(sb + b);
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 ev = __ldg(reinterpret_cast<const float4*>(eps) + i);
    float4 o;
    o.x = __fadd_rn(__fmul_rn(a, xv.x), __fmul_rn(c, ev.x));
    o.y = __fadd_rn(__fmul_rn(a, xv.y), __fmul_rn(c, ev.y));
    o.z = __fadd_rn(__fmul_rn(a, xv.z), __fmul_rn(c, ev.z));
    o.w = __fadd_rn(__fmul_rn(a, xv.w), __fmul_rn(c, ev.w));
    reinterpret_cast<float4*>(out)[i] = o;
  }
}
int launch_noise_images(const float* x, const float* eps, const float* sa, const float* sb, float* out, int n,
                        size_t per_sample, cudaStream_t s) {
  if (per_sample % 4) return static_cast<int>(cudaErrorInvalidValue);
  const size_t n4 = static_cast<size_t>(n) * per_sample / 4;
  int blocks = cdiv(static_cast<long long>(n4), 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  noise_images_kernel<<<blocks, 256, 0, s>>>(x, eps, sa, sb, out, per_sample / 4, n4);
  return static_cast<int>(cudaGetLastError());
}

__global__ void advance_kernel(int* trow, int n, int dec, int* step) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) trow[i] -= dec;
  if (threadIdx.x == 0) *step -= 1;
}
int launch_advance(int* trow, int n, int dec, int* step, cudaStream_t s) {
  advance_kernel<<<1, 256, 0, s>>>(trow, n, dec, step);
  return static_cast<int>(cudaGetLastError());
}
__global__ void set_rows_kernel(int* trow, const int* uniq, int n, int* step, int step_value, int mul) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) trow[i] = step_value * mul + (uniq ? uniq[i] : 0);
  if (threadIdx.x == 0) *step = step_value;
}
int launch_set_rows(int* trow, const int* uniq, int n, int* step, int step_value, int mul, cudaStream_t s) {
  set_rows_kernel<<<1, 256, 0, s>>>(trow, uniq, n, step, step_value, mul);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Aggregation blend (Aggregation_Sampling.py:91-110)
// ------------------------------------------------------------------------------------------------
__global__ void blend_gather_kernel(const float* __restrict__ patches, const int* __restrict__ ys, int ny,
                                    const int* __restrict__ xs, int nx, const float* __restrict__ weight,
                                    float* __restrict__ out, float* __restrict__ wsum_out, int C, int H, int W,
                                    int P, int do_clamp) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y;
  if (X >= W) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float cnt = 0.f;
  const size_t pp = static_cast<size_t>(P) * P;
  for (int iy = 0; iy < ny; ++iy) {
    const int ly = Y - __ldg(ys + iy);
    if (ly < 0 || ly >= P) continue;
    for (int ix = 0; ix < nx; ++ix) {
      const int lx = X - __ldg(xs + ix);
      if (lx < 0 || lx >= P) continue;
      const float w = __ldg(weight + static_cast<size_t>(ly) * P + lx);
      const float* pb = patches + static_cast<size_t>(iy * nx + ix) * C * pp + static_cast<size_t>(ly) * P + lx;
      for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(__ldg(pb + c * pp), w));
      cnt = __fadd_rn(cnt, w);
    }
  }
  for (int c = 0; c < C; ++c) {
    float v = __fdiv_rn(acc[c], cnt);
    if (do_clamp) v = fminf(fmaxf(v, 0.f), 1.f);
    out[(static_cast<size_t>(c) * H + Y) * W + X] = v;
  }
  if (wsum_out) wsum_out[static_cast<size_t>(Y) * W + X] = cnt;
}

int launch_blend_gather(const float* patches, const int* ys, int ny, const int* xs, int nx, const float* weight,
                        float* out, float* wsum_out, int C, int H, int W, int P, int do_clamp, cudaStream_t s) {
  if (C > 4) return static_cast<int>(cudaErrorInvalidValue);
  dim3 grid(cdiv(W, 256), H);
  blend_gather_kernel<<<grid, 256, 0, s>>>(patches, ys, ny, xs, nx, weight, out, wsum_out, C, H, W, P, do_clamp);
  return static_cast<int>(cudaGetLastError());
}

// Vector form (window starts, P and W multiples of 4): one thread = four consecutive output pixels of one row, every
// access a 128-bit one (a warp reads 512 contiguous bytes of each patch row it touches). The windows covering a row /
// a pixel quad come from two small range tables built on the host ((first, last + 1) indices into the sorted start
// lists), so nothing is scanned. Patches are visited in row-major patch order and every pixel is accumulated with a
// separately rounded multiply and add, exactly like the scalar kernel and the reference's `+=` sequence.
// 50 registers, five blocks per SM: 168 us = 6.1 TB/s (0.93 of the measured copy bandwidth) at the cfg-5 shape. (A
// variant that issued the loads of all of a pixel's patches before the first add needed 88 registers, two blocks per
// SM, and ran at 246 us.)
template <int C>
__device__ __forceinline__ void blend_accumulate4(float4 (&acc)[C], float4& cnt, const float4 (&v)[C], const float4& w) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    acc[c].x = __fadd_rn(acc[c].x, __fmul_rn(v[c].x, w.x));
    acc[c].y = __fadd_rn(acc[c].y, __fmul_rn(v[c].y, w.y));
    acc[c].z = __fadd_rn(acc[c].z, __fmul_rn(v[c].z, w.z));
    acc[c].w = __fadd_rn(acc[c].w, __fmul_rn(v[c].w, w.w));
  }
  cnt.x = __fadd_rn(cnt.x, w.x);
  cnt.y = __fadd_rn(cnt.y, w.y);
  cnt.z = __fadd_rn(cnt.z, w.z);
  cnt.w = __fadd_rn(cnt.w, w.w);
}

template <int C>
__global__ void __launch_bounds__(256)
blend_gather4_kernel(const float* __restrict__ patches, const int* __restrict__ ys, const int* __restrict__ xs, int nx,
                     const int2* __restrict__ row_rng, const int2* __restrict__ col_rng,
                     const float* __restrict__ weight, float* __restrict__ out, float* __restrict__ wsum_out, int H,
                     int W, int P, int do_clamp) {
  const int X4 = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y;
  if (4 * X4 >= W) return;
  const int X = 4 * X4;
  const int2 ry = __ldg(row_rng + Y);
  const int2 cx = __ldg(col_rng + X4);
  float4 acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 cnt = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t pp = static_cast<size_t>(P) * P;
  {
    for (int iy = ry.x; iy < ry.y; ++iy) {
      const int ly = Y - __ldg(ys + iy);
      for (int ix = cx.x; ix < cx.y; ++ix) {
        const int lx = X - __ldg(xs + ix);
        const size_t o = static_cast<size_t>(ly) * P + lx;
        const float4 w = __ldg(reinterpret_cast<const float4*>(weight + o));
        const float* pb = patches + static_cast<size_t>(iy * nx + ix) * C * pp + o;
        float4 v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldcs(reinterpret_cast<const float4*>(pb + c * pp));
        blend_accumulate4<C>(acc, cnt, v, w);
      }
    }
  }
  const size_t po = static_cast<size_t>(Y) * W + X;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float4 r;
    r.x = __fdiv_rn(acc[c].x, cnt.x);
    r.y = __fdiv_rn(acc[c].y, cnt.y);
    r.z = __fdiv_rn(acc[c].z, cnt.z);
    r.w = __fdiv_rn(acc[c].w, cnt.w);
    if (do_clamp) {
      r.x = fminf(fmaxf(r.x, 0.f), 1.f);
      r.y = fminf(fmaxf(r.y, 0.f), 1.f);
      r.z = fminf(fmaxf(r.z, 0.f), 1.f);
      r.w = fminf(fmaxf(r.w, 0.f), 1.f);
    }
    __stcs(reinterpret_cast<float4*>(out + static_cast<size_t>(c) * H * W + po), r);
  }
  if (wsum_out) __stcs(reinterpret_cast<float4*>(wsum_out + po), cnt);
}

// tables: [ny + nx] window starts, then int2 row_rng[H], int2 col_rng[W / 4] (device, 8-byte aligned ranges)
int launch_blend_gather4(const float* patches, const int* ys, int ny, const int* xs, int nx, const int2* row_rng,
                         const int2* col_rng, const float* weight, float* out, float* wsum_out, int C, int H, int W,
                         int P, int do_clamp, cudaStream_t s) {
  (void)ny;
  dim3 grid(cdiv(W / 4, 256), H);
  switch (C) {
    case 1: blend_gather4_kernel<1><<<grid, 256, 0, s>>>(patches, ys, xs, nx, row_rng, col_rng, weight, out, wsum_out, H, W, P, do_clamp); break;
    case 2: blend_gather4_kernel<2><<<grid, 256, 0, s>>>(patches, ys, xs, nx, row_rng, col_rng, weight, out, wsum_out, H, W, P, do_clamp); break;
    case 3: blend_gather4_kernel<3><<<grid, 256, 0, s>>>(patches, ys, xs, nx, row_rng, col_rng, weight, out, wsum_out, H, W, P, do_clamp); break;
    case 4: blend_gather4_kernel<4><<<grid, 256, 0, s>>>(patches, ys, xs, nx, row_rng, col_rng, weight, out, wsum_out, H, W, P, do_clamp); break;
    default: return static_cast<int>(cudaErrorInvalidValue);
  }
  return static_cast<int>(cudaGetLastError());
}

__global__ void blend_accumulate_kernel(const float* __restrict__ patch, const float* __restrict__ weight,
                                        float* __restrict__ acc, float* __restrict__ wsum, int C, int H, int W, int P,
                                        int y0, int x0) {
  const int lx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ly = blockIdx.y;
  if (lx >= P) return;
  const float w = weight[static_cast<size_t>(ly) * P + lx];
  const size_t o = static_cast<size_t>(y0 + ly) * W + (x0 + lx);
  for (int c = 0; c < C; ++c) {
    const size_t oi = static_cast<size_t>(c) * H * W + o;
    acc[oi] = __fadd_rn(acc[oi], __fmul_rn(patch[(static_cast<size_t>(c) * P + ly) * P + lx], w));
  }
  wsum[o] = __fadd_rn(wsum[o], w);
}
int launch_blend_accumulate(const float* patch, const float* weight, float* acc, float* wsum, int C, int H, int W,
                            int P, int y0, int x0, cudaStream_t s) {
  dim3 grid(cdiv(P, 128), P);
  blend_accumulate_kernel<<<grid, 128, 0, s>>>(patch, weight, acc, wsum, C, H, W, P, y0, x0);
  return static_cast<int>(cudaGetLastError());
}
__global__ void blend_finalize_kernel(float* __restrict__ acc, const float* __restrict__ wsum, int C, size_t HW,
                                      int do_clamp, int* zero_flag) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const float w = wsum[i];
  if (w == 0.f && zero_flag) atomicExch(zero_flag, 1);
  for (int c = 0; c < C; ++c) {
    float v = __fdiv_rn(acc[c * HW + i], w);
    if (do_clamp) v = fminf(fmaxf(v, 0.f), 1.f);
    acc[c * HW + i] = v;
  }
}
int launch_blend_finalize(float* acc, const float* wsum, int C, int H, int W, int do_clamp, int* zero_flag,
                          cudaStream_t s) {
  const size_t HW = static_cast<size_t>(H) * W;
  blend_finalize_kernel<<<cdiv(static_cast<long long>(HW), 256), 256, 0, s>>>(acc, wsum, C, HW, do_clamp, zero_flag);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// L2 flush for cold-cache timing: READS a buffer larger than L2 (a memset would leave the L2 full of dirty lines whose
// write-back then competes with the kernel being timed)
// ------------------------------------------------------------------------------------------------
__global__ void l2_flush_read_kernel(const uint4* __restrict__ buf, size_t n16, unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldg(buf + i);
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x9E3779B9u && sink) *sink = acc;  // practically never: keeps the loads alive
}
int launch_l2_flush_read(const void* buf, size_t bytes, cudaStream_t s) {
  l2_flush_read_kernel<<<148 * 8, 256, 0, s>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, nullptr);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Layout converters (layer-level debug entry points only)
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int C,
                                         int H, int W) {
  const long long total = static_cast<long long>(B) * C * H * W;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const long long p = idx / C;
  const int x = static_cast<int>(p % W);
  const int y = static_cast<int>((p / W) % H);
  const int b = static_cast<int>(p / (static_cast<long long>(W) * H));
  out[idx] = __float2bfloat16_rn(in[((static_cast<size_t>(b) * C + c) * H + y) * W + x]);
}
__global__ void nhwc_bf16_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int B, int C,
                                         int H, int W) {
  const long long total = static_cast<long long>(B) * C * H * W;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % W);
  const int y = static_cast<int>((idx / W) % H);
  const int c = static_cast<int>((idx / (static_cast<long long>(W) * H)) % C);
  const int b = static_cast<int>(idx / (static_cast<long long>(W) * H * C));
  out[idx] = __bfloat162float(in[((static_cast<size_t>(b) * H + y) * W + x) * C + c]);
}
int launch_nchw_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * C * H * W;
  nchw_to_nhwc_bf16_kernel<<<cdiv(total, 256), 256, 0, s>>>(in, reinterpret_cast<__nv_bfloat16*>(out), B, C, H, W);
  return static_cast<int>(cudaGetLastError());
}
int launch_nhwc_bf16_to_nchw(const void* in, float* out, int B, int C, int H, int W, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * C * H * W;
  nhwc_bf16_to_nchw_kernel<<<cdiv(total, 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(in), out, B, C, H,
                                                            W);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace drs
