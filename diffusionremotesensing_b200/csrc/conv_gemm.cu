// tcgen05 / TMEM / TMA implicit-GEMM convolution kernel for sm_100a (see conv_gemm.cuh for the model).
//
// CTA = 192 threads:
//   warp 0 (one lane)  TMA producer: per K-block one 5-D tensor load (activations) + one bulk copy (weights)
//   warp 1             owns the TMEM allocation; one lane issues tcgen05.mma and the commits
//   warps 2..5         epilogue: tcgen05.ld (thread = pixel row), fused affine / ReLU / adds, NHWC bf16 store
// A multi-stage smem ring (full/empty mbarriers) decouples TMA from the tensor core; a final commit on
// `tmem_full` hands the accumulator to the epilogue warps. Several CTAs are resident per SM (small stages,
// <= 512 TMEM columns each), so one CTA's epilogue overlaps its neighbours' main loops.
#include "conv_gemm.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>

namespace drs {

constexpr int kMaxStages = 8;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                 const ConvArgs a) {
  extern __shared__ uint8_t dyn_smem[];
  __shared__ __align__(16) KBlock s_kb[kMaxKBlocks];
  __shared__ __align__(8) uint64_t s_full[kMaxStages];
  __shared__ __align__(8) uint64_t s_empty[kMaxStages];
  __shared__ __align__(8) uint64_t s_tmem_full;
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_par[4][kMaxN];  // scale, bias, scale2 (or wvec), bias2

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const EpiArgs& e = a.epi;

  const uint32_t dyn_u32 = smem_u32(dyn_smem);
  uint8_t* const stage_base = dyn_smem + ((1024u - (dyn_u32 & 1023u)) & 1023u);

  const int tile = blockIdx.x;
  const int tx = tile % a.tiles_x;
  const int ty = (tile / a.tiles_x) % a.tiles_y;
  const int tb = tile / (a.tiles_x * a.tiles_y);
  const int x0 = tx * a.tw, y0 = ty * a.th, b0 = tb * a.tb;
  const int oc_off = blockIdx.y * a.n_sub;
  const int nkb = a.nkb;

  // ---- one-time setup --------------------------------------------------------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.kblocks + static_cast<size_t>(blockIdx.y) * nkb);
    uint4* dst = reinterpret_cast<uint4*>(s_kb);
    for (int i = threadIdx.x; i < nkb * 2; i += kGemmThreads) dst[i] = __ldg(src + i);
    for (int c = threadIdx.x; c < a.n_sub; c += kGemmThreads) {
      s_par[0][c] = e.scale ? __ldg(e.scale + oc_off + c) : 1.0f;
      s_par[1][c] = e.bias ? __ldg(e.bias + oc_off + c) : 0.0f;
      if (EPI == EPI_STD) {
        s_par[2][c] = e.scale2 ? __ldg(e.scale2 + oc_off + c) : 1.0f;
        s_par[3][c] = e.bias2 ? __ldg(e.bias2 + oc_off + c) : 0.0f;
      }
    }
    if (EPI != EPI_STD) {
      const int nw = (EPI == EPI_PSI) ? a.n_sub : e.nvec * a.n_sub;
      for (int c = threadIdx.x; c < nw; c += kGemmThreads) (&s_par[2][0])[c] = __ldg(e.wvec + c);
    }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map0);
    tma_prefetch_desc(&map1);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    mbar_init(&s_tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&s_tmem_base, static_cast<uint32_t>(a.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;

  if (warp == 0) {
    // ---- TMA producer --------------------------------------------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const KBlock K = s_kb[kb];
        mbar_wait(&s_empty[s], ph ^ 1u, a.err, 1);
        uint8_t* sa = stage_base + static_cast<size_t>(s) * a.stage_bytes;
        uint8_t* sb = sa + a.a_bytes;
        mbar_expect_tx(&s_full[s], static_cast<uint32_t>(kTileM * K.ck * 2) + K.b_bytes);
        tma_load_5d(sa, K.src ? &map1 : &map0, &s_full[s], K.c, x0 + K.dx, K.py, y0 + K.dy, b0);
        bulk_load(sb, a.wpack + K.b_off, K.b_bytes, &s_full[s]);
        if (++s == a.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----------------------------------------------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const KBlock K = s_kb[kb];
        mbar_wait(&s_full[s], ph, a.err, 2);
        tc_fence_after();
        const uint32_t sa = smem_u32(stage_base + static_cast<size_t>(s) * a.stage_bytes);
        const uint32_t sb = sa + a.a_bytes;
        const uint32_t row_bytes = K.ck * 2u;
        const uint64_t da = umma_desc_kmajor(sa, row_bytes);
        const uint64_t db = umma_desc_kmajor(sb, row_bytes);
        const uint32_t idesc = umma_idesc_bf16(kTileM, K.n);
        const int nk = K.ck >> 4;
        for (int k = 0; k < nk; ++k) {
          // +32 B per K=16 slice inside the swizzle span: start-address field is in 16 B units
          umma_bf16(tmem + K.col, da + 2u * k, db + 2u * k, idesc, (K.init && k == 0) ? 0u : 1u);
        }
        umma_commit(&s_empty[s]);  // frees the smem stage once these MMAs have drained
        if (++s == a.stages) { s = 0; ph ^= 1u; }
      }
      umma_commit(&s_tmem_full);
    }
  } else {
    // ---- epilogue ------------------------------------------------------------------------------
    mbar_wait(&s_tmem_full, 0, a.err, 3);
    tc_fence_after();
    const int q = warp & 3;             // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;      // accumulator row == pixel index inside the tile
    const int lx = row % a.tw;
    const int ly = (row / a.tw) % a.th;
    const int lb = row / (a.tw * a.th);
    const int x = x0 + lx, y = y0 + ly, b = b0 + lb;
    const bool valid = (x < a.W) && (y < a.H) && (b < a.B);
    const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const int N = a.n_sub;

    if (EPI == EPI_STD) {
      const int flags = e.flags;
      const float* te_row = nullptr;
      if (flags & (F_TE | F_PRE)) te_row = e.te + static_cast<size_t>(valid ? __ldg(e.trow + b) : 0) * e.te_stride;
      const float* te_post = te_row ? te_row + e.te_off + oc_off : nullptr;
      const float* te_pre = nullptr;
      if (flags & F_PRE) {
        const int ry = (y == 0) ? 0 : ((y == a.H - 1) ? 2 : 1);
        const int rx = (x == 0) ? 0 : ((x == a.W - 1) ? 2 : 1);
        te_pre = te_row + e.pre_off + (ry * 3 + rx) * e.OC + oc_off;
      }
      float rs = 1.0f;
      if ((flags & F_ROWSCALE) && valid)
        rs = __ldg(e.psi + (static_cast<size_t>(b) * (a.H >> 1) + (y >> 1)) * (a.W >> 1) + (x >> 1));

      for (int g = 0; g < e.n_groups; ++g) {
        const int oy = (e.oscale == 2) ? (2 * y + (g >> 1)) : y;
        const int ox = (e.oscale == 2) ? (2 * x + (g & 1)) : x;
        __nv_bfloat16* optr = reinterpret_cast<__nv_bfloat16*>(e.out) +
                              ((static_cast<size_t>(b) * e.OH + oy) * e.OW + ox) * e.OC + oc_off;
        const uint32_t colbase = static_cast<uint32_t>(g * N);
        for (int c0 = 0; c0 < N; c0 += 16) {
          float v[16], w[16];
          tmem_ld16(taddr + colbase + c0, v);
          if (flags & (F_DUAL_PRE | F_DUAL_POST)) tmem_ld16(taddr + e.col2 + c0, w);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            float t = v[i];
            if (flags & F_ROWSCALE) t *= rs;
            if (flags & F_PRE) t += valid ? __ldg(te_pre + c) : 0.0f;
            t = fmaf(t, s_par[0][c], s_par[1][c]);
            if (flags & F_DUAL_PRE) t = fmaf(w[i], s_par[2][c], t);
            if (flags & F_RELU) t = fmaxf(t, 0.0f);
            if (flags & F_DUAL_POST) t += w[i] + s_par[3][c];
            if (flags & F_TE) t += valid ? __ldg(te_post + c) : 0.0f;
            v[i] = t;
          }
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
      // psi = sigmoid(w . relu(acc + bias) + b): thread holds every channel of its pixel
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
      if (valid) reinterpret_cast<float*>(e.out)[(static_cast<size_t>(b) * a.H + y) * a.W + x] = sg;
    } else {
      // output 1x1 conv (N -> nvec <= 4) on the fp32 accumulator, fp32 NCHW result (optionally the DDPM update)
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
        const size_t plane = static_cast<size_t>(a.H) * a.W;
        const size_t pix = static_cast<size_t>(y) * a.W + x;
        if (e.flags & F_UPDATE) {
          const float4 cf = __ldg(reinterpret_cast<const float4*>(e.coef) + __ldg(e.step));
          for (int k = 0; k < e.nvec; ++k) {
            const size_t idx = (static_cast<size_t>(b) * e.nvec + k) * plane + pix;
            const float eps = acc[k] + __ldg(e.bvec + k);
            // same rounding order as the reference expression (no FMA contraction)
            float r = __fmul_rn(cf.x, __fsub_rn(e.x[idx], __fmul_rn(cf.y, eps)));
            if (e.noise) r = __fadd_rn(r, __fmul_rn(cf.z, __ldg(e.noise + idx)));
            e.x[idx] = r;
          }
        } else {
          float* o = reinterpret_cast<float*>(e.out);
          for (int k = 0; k < e.nvec; ++k)
            o[(static_cast<size_t>(b) * e.nvec + k) * plane + pix] = acc[k] + __ldg(e.bvec + k);
        }
      }
    }
  }

  // ---- teardown ----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, static_cast<uint32_t>(a.tmem_cols));
  }
}

static constexpr int kMaxDynSmem = 200 * 1024;

int conv_gemm_set_smem_limits() {
  cudaError_t e;
  e = cudaFuncSetAttribute(conv_gemm_kernel<EPI_STD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(conv_gemm_kernel<EPI_PSI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(conv_gemm_kernel<EPI_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  return static_cast<int>(e);
}

int launch_conv_gemm(int epi_kind, const CUtensorMap& map0, const CUtensorMap& map1, const ConvArgs& args,
                     int n_tiles, int nsplit, size_t smem_bytes, cudaStream_t stream) {
  dim3 grid(static_cast<unsigned>(n_tiles), static_cast<unsigned>(nsplit), 1);
  dim3 block(kGemmThreads, 1, 1);
  switch (epi_kind) {
    case EPI_STD: conv_gemm_kernel<EPI_STD><<<grid, block, smem_bytes, stream>>>(map0, map1, args); break;
    case EPI_PSI: conv_gemm_kernel<EPI_PSI><<<grid, block, smem_bytes, stream>>>(map0, map1, args); break;
    case EPI_OUT: conv_gemm_kernel<EPI_OUT><<<grid, block, smem_bytes, stream>>>(map0, map1, args); break;
    default: return static_cast<int>(cudaErrorInvalidValue);
  }
  return static_cast<int>(cudaGetLastError());
}

}  // namespace drs
