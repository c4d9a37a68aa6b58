// tcgen05 / TMEM / TMA implicit-GEMM convolution kernel for sm_100a (see conv_gemm.cuh for the model).
//
// CTA = 192 threads:
//   warp 0 (one lane)  TMA producer: per K-block one 5-D tensor load (activations) + one bulk copy (weights)
//   warp 1             owns the TMEM allocation; one lane issues tcgen05.mma and the commits
//   warps 2..5         epilogue: tcgen05.ld (thread = pixel row), fused affine / ReLU / adds, NHWC bf16 store
// A multi-stage smem ring (full/empty mbarriers) decouples TMA from the tensor core; a final commit on
// `tmem_full` hands the accumulator to the epilogue warps. Several CTAs are resident per SM (small stages,
// <= 512 TMEM columns each), so one CTA's epilogue overlaps its neighbours' main loops.
#include "conv_epilogue.cuh"
#include "conv_gemm.cuh"
#include "ptx.cuh"

namespace drs {

constexpr int kMaxStages = 8;

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
  __shared__ __align__(16) float s_par[4][kMaxN];  // scale, bias, scale2 (or wvec), bias2

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
    load_epilogue_params<EPI>(e, a.n_sub, oc_off, s_par, threadIdx.x, kGemmThreads);
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
    conv_epilogue<EPI>(e, taddr, x, y, b, valid, a.W, a.H, a.n_sub, oc_off, s_par);
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
