// Epilogue shared by the tensor-core convolution kernels: one thread owns one accumulator row (= output pixel),
// reads its N fp32 columns from TMEM in chunks of 16 and applies the fused per-channel work.
#pragma once
#include <cuda_bf16.h>

#include "conv_gemm.cuh"
#include "ptx.cuh"

namespace drs {

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void ld16_shared(const float* src, float* dst) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = s4[q];
    dst[4 * q + 0] = t.x;
    dst[4 * q + 1] = t.y;
    dst[4 * q + 2] = t.z;
    dst[4 * q + 3] = t.w;
  }
}
__device__ __forceinline__ void ld16_global(const float* src, float* dst) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = __ldg(s4 + q);
    dst[4 * q + 0] = t.x;
    dst[4 * q + 1] = t.y;
    dst[4 * q + 2] = t.z;
    dst[4 * q + 3] = t.w;
  }
}

// taddr: TMEM address of this thread's lane quadrant at column 0 of the tile's accumulator.
// (x, y, b): output-grid pixel of this thread; valid: inside the grid; W, H: grid size; N: channels of this
// grid.y slice; oc_off: first output channel of the slice; s_par: [4][kMaxN] = scale, bias, scale2 | wvec, bias2.
template <int EPI>
__device__ __forceinline__ void conv_epilogue(const EpiArgs& e, uint32_t taddr, int x, int y, int b, bool valid, int W,
                                              int H, int N, int oc_off, const float (*s_par)[kMaxN]) {
  if (EPI == EPI_STD) {
    const int flags = e.flags;
    const float* te_row = nullptr;
    if (flags & (F_TE | F_PRE)) te_row = e.te + static_cast<size_t>(valid ? __ldg(e.trow + b) : 0) * e.te_stride;
    const float* te_post = te_row ? te_row + e.te_off + oc_off : nullptr;
    const float* te_pre = nullptr;
    if (flags & F_PRE) {
      const int ry = (y == 0) ? 0 : ((y == H - 1) ? 2 : 1);
      const int rx = (x == 0) ? 0 : ((x == W - 1) ? 2 : 1);
      te_pre = te_row + e.pre_off + (ry * 3 + rx) * e.OC + oc_off;
    }
    float rs = 1.0f;
    if ((flags & F_ROWSCALE) && valid)
      rs = __ldg(e.psi + (static_cast<size_t>(b) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1));

    for (int g = 0; g < e.n_groups; ++g) {
      const int oy = (e.oscale == 2) ? (2 * y + (g >> 1)) : y;
      const int ox = (e.oscale == 2) ? (2 * x + (g & 1)) : x;
      __nv_bfloat16* optr = reinterpret_cast<__nv_bfloat16*>(e.out) +
                            ((static_cast<size_t>(b) * e.OH + oy) * e.OW + ox) * e.OC + oc_off;
      const uint32_t colbase = static_cast<uint32_t>(g * N);
      for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16], w[16], p[16];
        tmem_ld16(taddr + colbase + c0, v);
        if (flags & (F_DUAL_PRE | F_DUAL_POST)) tmem_ld16(taddr + e.col2 + c0, w);
        // per-channel vectors as 128-bit loads, issued before the TMEM wait so their latency overlaps
        float sc[16], bi[16];
        ld16_shared(&s_par[0][c0], sc);
        ld16_shared(&s_par[1][c0], bi);
        if (flags & F_PRE) ld16_global(te_pre + c0, p);
        tmem_ld_wait();
        if (flags & F_ROWSCALE) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= rs;
        }
        if (flags & F_PRE) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += p[i];
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], sc[i], bi[i]);
        if (flags & F_DUAL_PRE) {
          ld16_shared(&s_par[2][c0], sc);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaf(w[i], sc[i], v[i]);
        }
        if (flags & F_RELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
        }
        if (flags & F_DUAL_POST) {
          ld16_shared(&s_par[3][c0], bi);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += w[i] + bi[i];
        }
        if (flags & F_TE) {
          ld16_global(te_post + c0, p);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += p[i];
        }
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
        if (valid) {
          uint4* o = reinterpret_cast<uint4*>(optr + c0);
          o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
    }
  } else if (EPI == EPI_PSI) {
    // psi = sigmoid(w . relu(acc + bias) + b): the thread holds every channel of its pixel
    float p = 0.0f;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = c0 + i;
        p = fmaf(s_par[2][c], fmaxf(fmaf(v[i], s_par[0][c], s_par[1][c]), 0.0f), p);
      }
    }
    p += __ldg(e.bvec);
    const float sg = 1.0f / (1.0f + __expf(-p));
    if (valid) reinterpret_cast<float*>(e.out)[(static_cast<size_t>(b) * H + y) * W + x] = sg;
  } else {
    // output 1x1 conv (N -> nvec <= 4) on the fp32 accumulator, fp32 NCHW result
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = c0 + i;
        const float t = fmaf(v[i], s_par[0][c], s_par[1][c]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < e.nvec) acc[k] = fmaf((&s_par[2][0])[k * N + c], t, acc[k]);
      }
    }
    if (valid) {
      const size_t plane = static_cast<size_t>(H) * W;
      const size_t pix = static_cast<size_t>(y) * W + x;
      float* o = reinterpret_cast<float*>(e.out);
      for (int k = 0; k < e.nvec; ++k) o[(static_cast<size_t>(b) * e.nvec + k) * plane + pix] = acc[k] + __ldg(e.bvec + k);
    }
  }
}

// Loads the per-channel epilogue parameters of one grid.y slice into shared memory (all threads of the CTA).
template <int EPI>
__device__ __forceinline__ void load_epilogue_params(const EpiArgs& e, int n_sub, int oc_off, float (*s_par)[kMaxN],
                                                     int tid, int nthreads) {
  for (int c = tid; c < n_sub; c += nthreads) {
    s_par[0][c] = e.scale ? __ldg(e.scale + oc_off + c) : 1.0f;
    s_par[1][c] = e.bias ? __ldg(e.bias + oc_off + c) : 0.0f;
    if (EPI == EPI_STD) {
      s_par[2][c] = e.scale2 ? __ldg(e.scale2 + oc_off + c) : 1.0f;
      s_par[3][c] = e.bias2 ? __ldg(e.bias2 + oc_off + c) : 0.0f;
    }
  }
  if (EPI != EPI_STD) {
    const int nw = (EPI == EPI_PSI) ? n_sub : e.nvec * n_sub;
    for (int c = tid; c < nw; c += nthreads) (&s_par[2][0])[c] = __ldg(e.wvec + c);
  }
}

}  // namespace drs
