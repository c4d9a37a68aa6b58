// Epilogue shared by the tensor-core convolution kernels: one thread owns one accumulator row (= output pixel),
// reads its N fp32 columns from TMEM in chunks of 16 and applies the fused per-channel work.
#pragma once
#include <cuda_bf16.h>

#include "conv_gemm.cuh"
#include "ptx.cuh"

namespace drs {

#ifdef DRS_EPI_TRACE
static __device__ long long g_epi_trace[64];
static __device__ int g_epi_trace_on;
#define EPI_TL(slot)                                                                    \
  do {                                                                                  \
    if (threadIdx.x == 0 && blockIdx.x == 0 && g_epi_trace_on && (slot) < 64) g_epi_trace[(slot)] = clock64();                 \
  } while (0)
#else
#define EPI_TL(slot) do {} while (0)
#endif

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

// Staged output path of the persistent kernel: every epilogue warp owns a 4 KiB shared-memory staging area split
// into nbuf sub-box buffers of 32 pixel rows x sbc channels (bf16). A thread writes its pixel's channels there
// (swizzled like the output tensor map, so the quarter-warp stores are bank-conflict free) and one lane hands each
// finished sub-box to the TMA unit, which writes full 32-byte sectors to global memory. The per-thread 16-byte
// global stores this replaces cost one L1 tag cycle per touched line and instruction (32 per warp store).
struct TmaStoreCtx {
  const CUtensorMap* map;  // output view (c, x, py, y, b): box = (sbc, 8, 1, 4, 1)
  uint8_t* stage;          // this warp's staging area, 1024-byte aligned
  int sbc;                 // channels per sub-box: min(64, N)
  int nbuf;                // buffers in the staging area: 4096 / (32 * sbc * 2), at most 4
  int buf;                 // next buffer (persists across tiles)
  int x0, y0, b;           // box origin of this warp: tile x, tile y + 4 * quadrant, image
};

// one 256-bit global store (a full 32-byte sector) instead of two 128-bit ones; p is 32-byte aligned
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// taddr: TMEM address of this thread's lane quadrant at column 0 of the tile's accumulator.
// (x, y, b): output-grid pixel of this thread; valid: inside the grid; W, H: grid size; N: channels of this
// grid.y slice; oc_off: first output channel of the slice; s_par: [4][kMaxN] = scale, bias, scale2 | wvec, bias2.
template <int EPI>
__device__ __forceinline__ void conv_epilogue(const EpiArgs& e, uint32_t taddr, int x, int y, int b, bool valid, int W,
                                              int H, int N, int oc_off, const float (*s_par)[kMaxN],
                                              TmaStoreCtx* ts = nullptr, int g_begin = 0, int g_end = -1) {
  if (g_end < 0) g_end = e.n_groups;
  if (EPI == EPI_STD) {
    const int flags = e.flags;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t st_row = 0, st_mask = 0, st_base = 0, st_bufbytes = 0;
    if (ts) {
      st_row = lane * static_cast<uint32_t>(ts->sbc) * 2u;         // byte offset of this thread's pixel row
      st_mask = (ts->sbc == 64) ? 7u : ((ts->sbc == 32) ? 3u : 1u);  // SWIZZLE_128B / 64B / 32B
      st_base = smem_u32(ts->stage);
      st_bufbytes = 64u * static_cast<uint32_t>(ts->sbc);
    }
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

    for (int g = g_begin; g < g_end; ++g) {
      const int oy = (e.oscale == 2) ? (2 * y + (g >> 1)) : y;
      const int ox = (e.oscale == 2) ? (2 * x + (g & 1)) : x;
      __nv_bfloat16* optr = reinterpret_cast<__nv_bfloat16*>(e.out) +
                            ((static_cast<size_t>(b) * e.OH + oy) * e.OW + ox) * e.OC + oc_off;
      const uint32_t colbase = static_cast<uint32_t>(g * N);
      for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16], w[16], p[16];
        EPI_TL(g * 16 + (c0 >> 4) * 4 + 0);
        tmem_ld16(taddr + colbase + c0, v);
        if (flags & (F_DUAL_PRE | F_DUAL_POST)) tmem_ld16(taddr + e.col2 + c0, w);
        // per-channel vectors as 128-bit loads, issued before the TMEM wait so their latency overlaps
        float sc[16], bi[16];
        ld16_shared(&s_par[0][c0], sc);
        ld16_shared(&s_par[1][c0], bi);
        if (flags & F_PRE) ld16_global(te_pre + c0, p);
        tmem_ld_wait();
        EPI_TL(g * 16 + (c0 >> 4) * 4 + 1);
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
        EPI_TL(g * 16 + (c0 >> 4) * 4 + 2);
        if (ts) {
          const int cs = c0 & (ts->sbc - 1);  // first channel of this chunk inside its sub-box
          if (cs == 0) {
            // the buffer about to be overwritten must have been read by the store issued nbuf sub-boxes ago
            if (lane == 0) {
              if (ts->nbuf == 1)
                bulk_wait_read<0>();
              else if (ts->nbuf == 2)
                bulk_wait_read<1>();
              else
                bulk_wait_read<3>();
            }
            __syncwarp();
          }
          const uint32_t buf = st_base + static_cast<uint32_t>(ts->buf) * st_bufbytes;
          const uint32_t off = st_row + static_cast<uint32_t>(cs) * 2u;
          const uint32_t sw = ((off >> 7) & st_mask) << 4;
          st_shared_v4(buf + (off ^ sw), pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(buf + ((off + 16u) ^ sw), pk[4], pk[5], pk[6], pk[7]);
          if (cs + 16 == ts->sbc) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              const int cc = ((e.oscale == 2) ? (g & 1) * e.OC : 0) + oc_off + c0 + 16 - ts->sbc;
              tma_store_5d(ts->map, reinterpret_cast<const void*>(ts->stage + static_cast<size_t>(ts->buf) * st_bufbytes),
                           cc, ts->x0, (e.oscale == 2) ? (g >> 1) : 0, ts->y0, ts->b);
              bulk_commit();
            }
            ts->buf = (ts->buf + 1 == ts->nbuf) ? 0 : ts->buf + 1;
          }
        } else if (valid) {
          st_global_v8(optr + c0, pk);
        }
        EPI_TL(g * 16 + (c0 >> 4) * 4 + 3);
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
      // per-channel vectors as 128-bit loads (shared-memory bandwidth is what bounds this kernel): scale, bias and
      // one float4 of the (up to four) output-conv weights per channel, zero beyond nvec
      float sc[16], bi[16];
      ld16_shared(&s_par[0][c0], sc);
      ld16_shared(&s_par[1][c0], bi);
      tmem_ld_wait();
      const float4* wq = reinterpret_cast<const float4*>(&s_par[2][0]) + c0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float t = fmaf(v[i], sc[i], bi[i]);
        const float4 w4 = wq[i];
        acc[0] = fmaf(w4.x, t, acc[0]);
        acc[1] = fmaf(w4.y, t, acc[1]);
        acc[2] = fmaf(w4.z, t, acc[2]);
        acc[3] = fmaf(w4.w, t, acc[3]);
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

// ------------------------------------------------------------------------------------------------
// EPI_STD with the flag word known at compile time (persistent kernel): no per-chunk flag tests, the time-embedding
// row comes from a per-warp shared-memory copy (s_te, filled before the accumulator wait), and the TMEM load of chunk
// i + 1 is in flight while chunk i is scaled, packed and stored.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld16_raw(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// The wait names the destination registers as read-write operands, so the compiler can neither read nor copy them
// before the asynchronous load has landed.
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&r)[16], uint32_t (&q)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]),
                 "+r"(q[8]), "+r"(q[9]), "+r"(q[10]), "+r"(q[11]), "+r"(q[12]), "+r"(q[13]), "+r"(q[14]), "+r"(q[15])
               :
               : "memory");
}

template <int FL>
struct StdEpilogue {
  static constexpr bool kDual = (FL & (F_DUAL_PRE | F_DUAL_POST)) != 0;
  const EpiArgs& e;
  const float (*s_par)[kMaxN];
  const float* s_te;      // this warp's staged time-embedding row (F_TE)
  const float* te_pre;    // this thread's border-class row (F_PRE)
  TmaStoreCtx* ts;
  __nv_bfloat16* orow;    // direct stores: this thread's output pixel of the current group
  uint32_t lane, st_row, st_mask, st_base, st_bufbytes;
  float rs;
  bool valid;
  int oc_off;
  int tl_slot = 0;  // DRS_EPI_TRACE: first trace slot of the chunk being processed

  // scale / pack / store one chunk of 16 channels starting at channel c0 of column group g
  __device__ __forceinline__ void chunk(const uint32_t (&rv)[16], const uint32_t (&rw)[16], int g, int c0) {
    uint32_t pk[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = c0 + 4 * q;
      // Parameter reads go through the shared-memory pipe the MMAs saturate: with none at all the step is 3 % faster
      // (timing experiment, profiles/epilogue_trace_r2.txt), so the scale vector is skipped where it is 1.
      const float4 sc = (FL & F_NOSCALE) ? make_float4(1.f, 1.f, 1.f, 1.f)
                                         : *reinterpret_cast<const float4*>(&s_par[0][c]);
      const float4 bi = *reinterpret_cast<const float4*>(&s_par[1][c]);
      float x0 = __uint_as_float(rv[4 * q + 0]), x1 = __uint_as_float(rv[4 * q + 1]);
      float x2 = __uint_as_float(rv[4 * q + 2]), x3 = __uint_as_float(rv[4 * q + 3]);
      if (FL & F_ROWSCALE) { x0 *= rs; x1 *= rs; x2 *= rs; x3 *= rs; }
      if (FL & F_PRE) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(te_pre + c));
        x0 += t.x; x1 += t.y; x2 += t.z; x3 += t.w;
      }
      if (FL & F_NOSCALE) { x0 += bi.x; x1 += bi.y; x2 += bi.z; x3 += bi.w; }   // == fmaf(x, 1, b) bit for bit
      else { x0 = fmaf(x0, sc.x, bi.x); x1 = fmaf(x1, sc.y, bi.y); x2 = fmaf(x2, sc.z, bi.z); x3 = fmaf(x3, sc.w, bi.w); }
      if (FL & F_DUAL_PRE) {
        const float4 s2 = *reinterpret_cast<const float4*>(&s_par[2][c]);
        x0 = fmaf(__uint_as_float(rw[4 * q + 0]), s2.x, x0); x1 = fmaf(__uint_as_float(rw[4 * q + 1]), s2.y, x1);
        x2 = fmaf(__uint_as_float(rw[4 * q + 2]), s2.z, x2); x3 = fmaf(__uint_as_float(rw[4 * q + 3]), s2.w, x3);
      }
      if (FL & F_RELU) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
      if (FL & F_DUAL_POST) {
        const float4 b2 = *reinterpret_cast<const float4*>(&s_par[3][c]);
        x0 += __uint_as_float(rw[4 * q + 0]) + b2.x; x1 += __uint_as_float(rw[4 * q + 1]) + b2.y;
        x2 += __uint_as_float(rw[4 * q + 2]) + b2.z; x3 += __uint_as_float(rw[4 * q + 3]) + b2.w;
      }
      if (FL & F_TE) {
        const float4 t = *reinterpret_cast<const float4*>(s_te + c);
        x0 += t.x; x1 += t.y; x2 += t.z; x3 += t.w;
      }
      pk[2 * q] = pack_bf16(x0, x1);
      pk[2 * q + 1] = pack_bf16(x2, x3);
    }
    EPI_TL(tl_slot + 2);
    if (ts) {
      const int cs = c0 & (ts->sbc - 1);  // first channel of this chunk inside its sub-box
      if (cs == 0) {
        // the buffer about to be overwritten must have been read by the store issued nbuf sub-boxes ago
        if (lane == 0) {
          if (ts->nbuf == 1)
            bulk_wait_read<0>();
          else if (ts->nbuf == 2)
            bulk_wait_read<1>();
          else
            bulk_wait_read<3>();
        }
        __syncwarp();
      }
      const uint32_t buf = st_base + static_cast<uint32_t>(ts->buf) * st_bufbytes;
      const uint32_t off = st_row + static_cast<uint32_t>(cs) * 2u;
      const uint32_t sw = ((off >> 7) & st_mask) << 4;
      st_shared_v4(buf + (off ^ sw), pk[0], pk[1], pk[2], pk[3]);
      st_shared_v4(buf + ((off + 16u) ^ sw), pk[4], pk[5], pk[6], pk[7]);
      if (cs + 16 == ts->sbc) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          const int cc = ((e.oscale == 2) ? (g & 1) * e.OC : 0) + oc_off + c0 + 16 - ts->sbc;
          tma_store_5d(ts->map, reinterpret_cast<const void*>(ts->stage + static_cast<size_t>(ts->buf) * st_bufbytes), cc,
                       ts->x0, (e.oscale == 2) ? (g >> 1) : 0, ts->y0, ts->b);
          bulk_commit();
        }
        ts->buf = (ts->buf + 1 == ts->nbuf) ? 0 : ts->buf + 1;
      }
    } else if (valid) {
      st_global_v8(orow + c0, pk);
    }
  }
};

template <int FL>
__device__ __forceinline__ void conv_epilogue_std_ct(const EpiArgs& e, uint32_t taddr, int x, int y, int b, bool valid,
                                                     int W, int H, int N, int oc_off, const float (*s_par)[kMaxN],
                                                     const float* s_te, TmaStoreCtx* ts, int g_begin, int g_end) {
  StdEpilogue<FL> E{e, s_par, s_te, nullptr, ts, nullptr, threadIdx.x & 31u, 0, 0, 0, 0, 1.0f, valid, oc_off};
  if (ts) {
    E.st_row = E.lane * static_cast<uint32_t>(ts->sbc) * 2u;
    E.st_mask = (ts->sbc == 64) ? 7u : ((ts->sbc == 32) ? 3u : 1u);
    E.st_base = smem_u32(ts->stage);
    E.st_bufbytes = 64u * static_cast<uint32_t>(ts->sbc);
  }
  if (FL & F_PRE) {
    const int ry = (y == 0) ? 0 : ((y == H - 1) ? 2 : 1);
    const int rx = (x == 0) ? 0 : ((x == W - 1) ? 2 : 1);
    E.te_pre = e.te + static_cast<size_t>(valid ? __ldg(e.trow + b) : 0) * e.te_stride + e.pre_off +
               (ry * 3 + rx) * e.OC + oc_off;
  }
  if ((FL & F_ROWSCALE) && !(FL & F_GATE) && valid)
    E.rs = __ldg(e.psi + (static_cast<size_t>(b) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1));
  if (FL & F_GATE) {
    // Fused attention gate (UNet_model_superres.py:101-106): accumulator columns [0, nvec) hold W_g g + W_x x of this
    // thread's (gate-resolution) pixel; psi = sigmoid(w_psi . relu(. + b_g + b_x) + b_psi) stays in a register and
    // scales the four parity groups behind them (same expression, same order as EPI_PSI: both paths agree bitwise).
    float p = 0.0f;
    for (int c0 = 0; c0 < e.nvec; c0 += 16) {
      float v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = c0 + i;
        p = fmaf(s_par[3][c], fmaxf(fmaf(v[i], 1.0f, s_par[2][c]), 0.0f), p);
      }
    }
    p += __ldg(e.bvec);
    E.rs = 1.0f / (1.0f + __expf(-p));
    if (e.psi_out && oc_off == 0 && g_begin == 0 && valid)
      e.psi_out[(static_cast<size_t>(b) * H + y) * W + x] = E.rs;
    taddr += static_cast<uint32_t>(e.nvec);  // the groups follow the gate columns
  }

  if (StdEpilogue<FL>::kDual) {
    // two accumulators per chunk: single-buffered (32 registers), which keeps the kernel at two CTAs per SM
    uint32_t v[16], w[16];
    E.orow = reinterpret_cast<__nv_bfloat16*>(e.out) + ((static_cast<size_t>(b) * e.OH + y) * e.OW + x) * e.OC + oc_off;
    for (int c0 = 0; c0 < N; c0 += 16) {
      tmem_ld16_raw(taddr + c0, v);
      tmem_ld16_raw(taddr + e.col2 + c0, w);
      tmem_ld_wait32(v, w);
      E.chunk(v, w, 0, c0);
    }
    return;
  }
  uint32_t va[16], vb[16], wa[16], wb[16];
  tmem_ld16_raw(taddr + static_cast<uint32_t>(g_begin * N), va);
  for (int g = g_begin; g < g_end; ++g) {
    const int oy = (e.oscale == 2) ? (2 * y + (g >> 1)) : y;
    const int ox = (e.oscale == 2) ? (2 * x + (g & 1)) : x;
    E.orow = reinterpret_cast<__nv_bfloat16*>(e.out) + ((static_cast<size_t>(b) * e.OH + oy) * e.OW + ox) * e.OC + oc_off;
    const uint32_t colbase = taddr + static_cast<uint32_t>(g * N);
    const bool more_groups = (g + 1 < g_end);
    for (int c0 = 0; c0 < N; c0 += 32) {
      // chunk A = [c0, c0 + 16): its load is in flight; start chunk B, work on A, then the same with roles swapped
      const bool has_b = (c0 + 16 < N);
      EPI_TL(((g - g_begin) * (N >> 4) + (c0 >> 4)) * 4 + 0);
      if (StdEpilogue<FL>::kDual) tmem_ld_wait32(va, wa); else tmem_ld_wait16(va);
      EPI_TL(((g - g_begin) * (N >> 4) + (c0 >> 4)) * 4 + 1);
      if (has_b) {
        tmem_ld16_raw(colbase + c0 + 16, vb);
        if (StdEpilogue<FL>::kDual) tmem_ld16_raw(taddr + e.col2 + c0 + 16, wb);
      }
      E.tl_slot = ((g - g_begin) * (N >> 4) + (c0 >> 4)) * 4;
      E.chunk(va, wa, g, c0);
      EPI_TL(((g - g_begin) * (N >> 4) + (c0 >> 4)) * 4 + 3);
      if (has_b) {
        EPI_TL(((g - g_begin) * (N >> 4) + (c0 >> 4) + 1) * 4 + 0);
        if (StdEpilogue<FL>::kDual) tmem_ld_wait32(vb, wb); else tmem_ld_wait16(vb);
        EPI_TL(((g - g_begin) * (N >> 4) + (c0 >> 4) + 1) * 4 + 1);
        if (c0 + 32 < N) {
          tmem_ld16_raw(colbase + c0 + 32, va);
          if (StdEpilogue<FL>::kDual) tmem_ld16_raw(taddr + e.col2 + c0 + 32, wa);
        } else if (more_groups) {
          tmem_ld16_raw(colbase + N, va);
          if (StdEpilogue<FL>::kDual) tmem_ld16_raw(taddr + e.col2, wa);
        }
        E.tl_slot = ((g - g_begin) * (N >> 4) + (c0 >> 4) + 1) * 4;
        E.chunk(vb, wb, g, c0 + 16);
        EPI_TL(((g - g_begin) * (N >> 4) + (c0 >> 4) + 1) * 4 + 3);
      } else if (more_groups) {
        tmem_ld16_raw(colbase + N, va);  // N == 16: no overlap (does not occur in the three UNets)
        if (StdEpilogue<FL>::kDual) tmem_ld16_raw(taddr + e.col2, wa);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Transposed convolution, 64 output channels per CTA, bias only (F_NOSCALE | F_TR64), TMA stores. The chunked
// epilogue above spends ~700 cycles per 16 channels on these layers (4 x 64 columns per tile and nothing to hide
// them behind: profiles/epilogue_trace_r2.txt): two rounds of parameter reads behind the shared-memory traffic of the
// MMAs, runtime sub-box bookkeeping on the uniform datapath and a branch every few instructions. Here the 64 biases
// live in registers for the whole kernel (they are the same for the four phases), a phase group is two 32-column
// TMEM loads (the second in flight under the math of the first), and the loop body is straight-line code: add, pack,
// eight 16-byte staging stores, one TMA store of the warp's 32 pixels x 64 channels. Same arithmetic (fp32 add,
// round-to-nearest-even pack), so the results are bit-identical to the chunked path.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32_raw(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_x32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                 "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                 "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// half: 0 / 1 = channels [0, 32) / [32, 64) of the group; `row` = this lane's 128-byte staging row, `sw` its swizzle
__device__ __forceinline__ void tr64_half(const uint32_t (&v)[32], const float (&bias)[64], int half, uint32_t row,
                                          uint32_t sw) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = 8 * j + 2 * i;
      pk[i] = pack_bf16(__uint_as_float(v[c]) + bias[32 * half + c], __uint_as_float(v[c + 1]) + bias[32 * half + c + 1]);
    }
    st_shared_v4(row + ((static_cast<uint32_t>(half * 4 + j) << 4) ^ sw), pk[0], pk[1], pk[2], pk[3]);
  }
}

// Staged stores only: per-thread 32-byte global stores of the phase scatter were measured slower than the chunked
// path (ups.1.transform 43.4 against 41.9 us), so launches without a staging area keep the chunked epilogue.
__device__ __forceinline__ void conv_epilogue_tr64(const EpiArgs& e, uint32_t taddr, int oc_off, const float (&bias)[64],
                                                   TmaStoreCtx* ts, int g_begin, int g_end) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t row = smem_u32(ts->stage) + lane * 128u;
  const uint32_t sw = (lane & 7u) << 4;
  uint32_t va[32], vb[32];
  tmem_ld32_raw(taddr + static_cast<uint32_t>(g_begin * 64), va);
  for (int g = g_begin; g < g_end; ++g) {
    const uint32_t col = taddr + static_cast<uint32_t>(g * 64);
    tmem_ld_wait_x32(va);
    tmem_ld32_raw(col + 32u, vb);
    // the staging rows are reused by every group: the TMA unit must have read the previous one
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
    tr64_half(va, bias, 0, row, sw);
    tmem_ld_wait_x32(vb);
    if (g + 1 < g_end) tmem_ld32_raw(col + 64u, va);
    tr64_half(vb, bias, 1, row, sw);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_5d(ts->map, reinterpret_cast<const void*>(ts->stage), (g & 1) * e.OC + oc_off, ts->x0, g >> 1, ts->y0,
                   ts->b);
      bulk_commit();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused attention gate with 32 channels (F_GATE | F_ROWSCALE, the full-resolution gate of the UNets), TMA stores:
// the same idea as conv_epilogue_tr64. The folded BatchNorm of the result conv (32 scales, 32 biases - the same for
// the four parity groups) stays in registers for the whole kernel, psi comes from one 32-column TMEM load, every
// parity group is one 32-column load (the next in flight under the math), four 16-byte staging stores and one TMA
// store of 32 pixels x 32 channels into one of the warp's two 2 KiB buffers. Expressions and their order are those
// of conv_epilogue_std_ct<F_GATE | F_ROWSCALE>, so both paths agree bitwise.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gate32_group(const uint32_t (&v)[32], float rs, const float (&sc)[32],
                                             const float (&bi)[32], const EpiArgs& e, TmaStoreCtx* ts, uint32_t row,
                                             uint32_t sw, int g, uint32_t lane) {
  // the buffer about to be overwritten must have been read by the store issued two groups ago
  if (lane == 0) bulk_wait_read<1>();
  __syncwarp();
  const uint32_t buf = row + static_cast<uint32_t>(ts->buf) * 2048u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = 8 * j + 2 * i;
      const float x0 = fmaf(__uint_as_float(v[c]) * rs, sc[c], bi[c]);
      const float x1 = fmaf(__uint_as_float(v[c + 1]) * rs, sc[c + 1], bi[c + 1]);
      pk[i] = pack_bf16(x0, x1);
    }
    st_shared_v4(buf + ((static_cast<uint32_t>(j) << 4) ^ sw), pk[0], pk[1], pk[2], pk[3]);
  }
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    tma_store_5d(ts->map, reinterpret_cast<const void*>(ts->stage + static_cast<size_t>(ts->buf) * 2048u),
                 (g & 1) * e.OC, ts->x0, g >> 1, ts->y0, ts->b);
    bulk_commit();
  }
  ts->buf ^= 1;
}

// (x, y, b): this thread's pixel of the gate-resolution grid; [g_begin, g_end): an even number of parity groups
__device__ __forceinline__ void conv_epilogue_gate32(const EpiArgs& e, uint32_t taddr, int x, int y, int b, bool valid,
                                                     int W, int H, const float (*s_par)[kMaxN], const float (&sc)[32],
                                                     const float (&bi)[32], TmaStoreCtx* ts, int g_begin, int g_end) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t va[32], vb[32];
  tmem_ld32_raw(taddr, va);
  tmem_ld_wait_x32(va);
  tmem_ld32_raw(taddr + 32u + static_cast<uint32_t>(g_begin * 32), vb);  // first group, under the gate math
  float p = 0.0f;
#pragma unroll
  for (int c = 0; c < 32; ++c)
    p = fmaf(s_par[3][c], fmaxf(fmaf(__uint_as_float(va[c]), 1.0f, s_par[2][c]), 0.0f), p);
  p += __ldg(e.bvec);
  const float rs = 1.0f / (1.0f + __expf(-p));
  if (e.psi_out && g_begin == 0 && valid) e.psi_out[(static_cast<size_t>(b) * H + y) * W + x] = rs;
  taddr += 32u;  // the groups follow the gate columns
  const uint32_t row = smem_u32(ts->stage) + lane * 64u;
  const uint32_t sw = ((lane >> 1) & 3u) << 4;
  for (int g = g_begin; g < g_end; g += 2) {
    tmem_ld_wait_x32(vb);
    tmem_ld32_raw(taddr + static_cast<uint32_t>((g + 1) * 32), va);
    gate32_group(vb, rs, sc, bi, e, ts, row, sw, g, lane);
    tmem_ld_wait_x32(va);
    if (g + 2 < g_end) tmem_ld32_raw(taddr + static_cast<uint32_t>((g + 2) * 32), vb);
    gate32_group(va, rs, sc, bi, e, ts, row, sw, g + 1, lane);
  }
}

// Loads the per-channel epilogue parameters of one grid.y slice into shared memory (all threads of the CTA).
template <int EPI>
__device__ __forceinline__ void load_epilogue_params(const EpiArgs& e, int n_sub, int oc_off, float (*s_par)[kMaxN],
                                                     int tid, int nthreads) {
  const bool gate = (EPI == EPI_STD) && (e.flags & F_GATE);
  for (int c = tid; c < n_sub; c += nthreads) {
    s_par[0][c] = e.scale ? __ldg(e.scale + oc_off + c) : 1.0f;
    s_par[1][c] = e.bias ? __ldg(e.bias + oc_off + c) : 0.0f;
    if (EPI == EPI_STD && !gate) {
      s_par[2][c] = e.scale2 ? __ldg(e.scale2 + oc_off + c) : 1.0f;
      s_par[3][c] = e.bias2 ? __ldg(e.bias2 + oc_off + c) : 0.0f;
    }
  }
  if (gate) {
    // every split needs the whole gate: bias (b_g + b_x) in scale2, w_psi in wvec, nvec channels
    for (int c = tid; c < e.nvec; c += nthreads) {
      s_par[2][c] = __ldg(e.scale2 + c);
      s_par[3][c] = __ldg(e.wvec + c);
    }
  }
  if (EPI == EPI_PSI) {
    for (int c = tid; c < n_sub; c += nthreads) (&s_par[2][0])[c] = __ldg(e.wvec + c);
  } else if (EPI == EPI_OUT) {
    // output-conv weights interleaved per channel: [c][k], k < 4, zero for k >= nvec (n_sub <= 128)
    for (int i = tid; i < 4 * n_sub; i += nthreads) {
      const int c = i >> 2, k = i & 3;
      (&s_par[2][0])[i] = (k < e.nvec) ? __ldg(e.wvec + k * n_sub + c) : 0.0f;
    }
  }
}

}  // namespace drs
