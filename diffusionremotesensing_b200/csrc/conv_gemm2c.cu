// CTA-pair variant of the persistent halo-tile convolution kernel (conv_gemm2.cu): tcgen05.mma.cta_group::2.
//
// Two CTAs of a cluster (one TPC) run every MMA together: M = 256 = the 128-pixel tile of CTA 0 stacked on the
// 128-pixel tile of CTA 1, and each CTA keeps only HALF of every weight tile (rows [r n/2, (r + 1) n/2) of the N x K
// tile in CTA r). Per SM this halves the weight bytes that have to be held in / streamed into shared memory and the
// B-operand bytes the tensor core reads per MMA, which is what bounds the layers with large K * N (weights that do
// not fit one SM, 64 <= N <= 256): a weight image of up to ~2 x 110 KiB becomes resident, and a streamed one costs
// half the L2 -> SM traffic per tile.
//
// Roles per CTA are those of conv_gemm2.cu (8 epilogue warps, 2 MMA issuer warps, 1 producer warp); differences:
//   * a "unit" is a pair of neighbouring tiles (2u, 2u + 1): CTA r loads / drains tile 2u + r; a cluster keeps two
//     units in flight (pipelines 0 / 1), i.e. four tiles;
//   * only CTA 0's issuer lanes issue MMAs; they wait on CTA 0's full barriers, which collect the TMA bytes of BOTH
//     CTAs (the loads of CTA 1 name CTA 0's barrier: cp.async.bulk.tensor ... .cta_group::2), and release slots,
//     weight stages and accumulators in both CTAs with multicast commits;
//   * the accumulator-empty barriers live in CTA 0 and count the epilogue threads of both CTAs (remote arrives);
//   * weights travel through a 2-D tensor map over the packed weight blob (rows of 128 bytes), because only tensor
//     loads can signal the peer's barrier.
// Requirements checked on the host: an even number of tiles, one weight-tile size per launch, n * ck * 2 a multiple
// of 2 KiB (so that half a tile is still a whole number of swizzle atoms).
#include <stdlib.h>
#include <string.h>

#include "conv_epilogue.cuh"
#include "conv_gemm2.cuh"
#include "ptx.cuh"

namespace drs {

namespace {

struct Ring {
  int idx;
  uint32_t phase;
  __device__ __forceinline__ void advance(int n) {
    if (++idx == n) {
      idx = 0;
      phase ^= 1u;
    }
  }
};

#define TLC(pair_no, slot)                                                                            \
  do {                                                                                                \
    if ((a.timeline & 1) && blockIdx.x == 0 && (pair_no) < 62) a.timeline_buf[(pair_no) * 8 + (slot)] = clock64(); \
  } while (0)

constexpr int kIssuer0 = 8;
constexpr int kProducer = 10;

// K-blocks [kb, kb + cnt) of one sub-tile: nk MMAs each. b_lo of the record is the offset of the FULL tile; the half
// tile of this CTA pair member sits at half that offset.
template <int NK>
__device__ __forceinline__ void issue_pair(const Conv2Prog& prog, int kb, int cnt, uint32_t acc, uint32_t slot16,
                                           uint32_t b_base16) {
  const uint32_t a_hi = prog.kb[kb].a_hi, b_hi = prog.kb[kb].b_hi;
  for (int i = 0; i < cnt; ++i) {
    const uint4 q = *reinterpret_cast<const uint4*>(&prog.kb[kb + i]);  // a_lo, b_lo, col | nk | flags, idesc
    const uint32_t a_lo = q.x + slot16;
    const uint32_t b_lo = ((q.y & 0xFFFFu) >> 1) + 0x10000u + b_base16;
    const uint32_t d = acc + (q.z & 0xFFFFu);
    const uint32_t idesc = (q.w & ~(0x1Fu << 24)) | (16u << 24);  // M = 256
    umma_bf16_split_2cta(d, a_lo, a_hi, b_lo, b_hi, idesc, ((q.z >> 24) & KB2_INIT) ? 0u : 1u);
#pragma unroll
    for (int k = 1; k < NK; ++k) umma_bf16_split_2cta(d, a_lo + 2u * k, a_hi, b_lo + 2u * k, b_hi, idesc, 1u);
  }
}

}  // namespace

template <int FL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemm2Threads)
conv_gemm2c_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                   const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ Conv2Args a, const __grid_constant__ Conv2Prog prog) {
  extern __shared__ uint8_t dyn_smem[];
  __shared__ __align__(8) uint64_t s_afull[kMaxASlots], s_aempty[kMaxASlots];
  __shared__ __align__(8) uint64_t s_bfull[kMaxBStages], s_bempty[kMaxBStages];
  __shared__ __align__(8) uint64_t s_tfull[4], s_tempty[4];  // [pipeline][accumulator buffer]
  __shared__ __align__(8) uint64_t s_wready;
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_par[4][kMaxN];
  constexpr bool kStageTe = (FL & F_TE) != 0;
  __shared__ __align__(16) float s_te[kStageTe ? 8 : 1][kStageTe ? kMaxN : 4];

  if ((a.timeline & 4) && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMin(&a.span_buf[(a.launch_id & 63) * 2], t);
  }
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const EpiArgs& e = a.epi;

  const uint32_t dyn_u32 = smem_u32(dyn_smem);
  uint8_t* const a_base = dyn_smem + ((1024u - (dyn_u32 & 1023u)) & 1023u);
  uint8_t* const b_base = a_base + static_cast<size_t>(a.a_slots) * a.a_slot_bytes;

  const int cid = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int split = cid % a.nsplit;
  const int first_unit = cid / a.nsplit;
  const int unit_step = n_clusters / a.nsplit;
  const int n_units = a.n_tiles >> 1;
  const int oc_off = split * a.n_sub;
  const int nkb = a.nkb;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  // weights: row index (128-byte rows of the blob) of this split's image; this CTA's half of a tile starts
  // rank * half_rows rows into the tile
  const int half_rows = a.cg2_half_tile_bytes >> 7;
  const int w_row0 = static_cast<int>((a.w_split_off + static_cast<size_t>(split) * a.w_split_bytes) >> 7) +
                     static_cast<int>(rank) * half_rows;
  const uint32_t half_image = a.w_split_bytes >> 1;

  // ---- one-time setup --------------------------------------------------------------------------
  load_epilogue_params<EPI_STD>(e, a.n_sub, oc_off, s_par, threadIdx.x, kGemm2Threads);
  if (warp == kProducer && lane == 0) {
    tma_prefetch_desc(&map0);
    tma_prefetch_desc(&map1);
    tma_prefetch_desc(&map_w);
    if (a.store_sbc) tma_prefetch_desc(&map_out);
    for (int s = 0; s < a.a_slots; ++s) {
      mbar_init(&s_afull[s], 1);
      mbar_init(&s_aempty[s], 1);
    }
    for (int s = 0; s < a.b_stages; ++s) {
      mbar_init(&s_bfull[s], 1);
      mbar_init(&s_bempty[s], 2);  // released by both issuers of CTA 0 (multicast commits)
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&s_tfull[s], 1);
      mbar_init(&s_tempty[s], a.solo ? 512 : 256);  // epilogue threads of both CTAs
    }
    mbar_init(&s_wready, 1);
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs are initialised before anybody signals the peer
  if (warp == kIssuer0) {
    tmem_alloc_2cta(&s_tmem_base, static_cast<uint32_t>(a.tmem_cols));
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  griddep_launch();

  if (warp == kProducer) {
    // ---- producer: this CTA's tiles and weight halves; completion goes to CTA 0's full barriers ------------
    const uint32_t wready_l = mapa_u32(smem_u32(&s_wready), 0);
    if (a.resident && elect_one()) {
      if (leader) mbar_expect_tx(&s_wready, a.w_split_bytes);  // both halves
      for (int kb = 0; kb < nkb; ++kb) {
        const KB3 K = prog.kb[kb];
        tma_load_2d_2cta(b_base + (K.b_off >> 1), &map_w, wready_l, 0, w_row0 + static_cast<int>(K.b_off >> 7));
      }
    }
    __syncwarp();
    griddep_wait();  // activations of the previous layer
    Ring ar{0, 0}, br{0, 0};
    for (int it = 0;; ++it) {
      const int u0 = first_unit + 2 * it * unit_step;
      if (u0 >= n_units) break;
      int x0[2], y0[2], bb[2];
      bool valid[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int unit = u0 + p * unit_step;
        valid[p] = unit < n_units;
        const int tile = 2 * unit + static_cast<int>(rank);
        bb[p] = tile / tiles_per_img;
        const int t2 = tile - bb[p] * tiles_per_img;
        y0[p] = (t2 / a.tiles_x) * kTile2H;
        x0[p] = (t2 % a.tiles_x) * kTile2W;
      }
      int st = 0, sub_first = 0, sub_cnt = 0;
      TLC(it, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const KB3 K = prog.kb[kb];
        if (K.flags & KB2_FIRST) {
          const SubTile T = prog.st[st++];
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            if (valid[p]) {
              mbar_wait(&s_aempty[ar.idx], ar.phase ^ 1u, a.err, 1);
              if (elect_one()) {
                if (leader) mbar_expect_tx(&s_afull[ar.idx], 2u * T.bytes);
                tma_load_5d_2cta(a_base + static_cast<size_t>(ar.idx) * a.a_slot_bytes, T.src ? &map1 : &map0,
                                 mapa_u32(smem_u32(&s_afull[ar.idx]), 0), T.c, x0[p] + T.dx0, 0, y0[p] + T.dy0, bb[p]);
              }
              __syncwarp();
            }
            ar.advance(a.a_slots);
          }
        }
        if (!a.resident) {
          if (K.flags & KB2_FIRST) {
            sub_first = kb;
            sub_cnt = static_cast<int>(K.b_bytes >> 24);
          }
          if ((kb - sub_first) % a.b_unit == 0) {
            const int nu = min(a.b_unit, sub_first + sub_cnt - kb);
            mbar_wait(&s_bempty[br.idx], br.phase ^ 1u, a.err, 1);
            if (elect_one()) {
              if (leader) mbar_expect_tx(&s_bfull[br.idx], 2u * static_cast<uint32_t>(nu) * a.cg2_half_tile_bytes);
              uint8_t* dst = b_base + static_cast<size_t>(br.idx) * a.b_stage_bytes;
              const uint32_t bar = mapa_u32(smem_u32(&s_bfull[br.idx]), 0);
              for (int j = 0; j < nu; ++j) {
                const KB3 Kj = prog.kb[kb + j];
                tma_load_2d_2cta(dst + static_cast<size_t>(j) * a.cg2_half_tile_bytes, &map_w, bar, 0,
                                 w_row0 + static_cast<int>(Kj.b_off >> 7));
              }
            }
            __syncwarp();
            br.advance(a.b_stages);
          }
        }
      }
    }
  } else if (warp == kIssuer0 || warp == kIssuer0 + 1) {
    // ---- MMA issuer of pipeline p (CTA 0 only; CTA 1's warps have nothing to do) ------------------------------
    const int p = warp - kIssuer0;
    if (leader) {
      if (a.resident) mbar_wait(&s_wready, 0, a.err, 2);
      Ring ar{0, 0}, br{0, 0};
      if (p) ar.advance(a.a_slots);
      const uint32_t b_base16 = smem_u32(b_base) >> 4;
      Ring tr{0, 0};
      for (int it = 0;; ++it) {
        const int u0 = first_unit + 2 * it * unit_step;
        if (u0 >= n_units) break;
        const bool valid = (u0 + p * unit_step) < n_units;
        const int tb = p * 2 + tr.idx;
        const uint32_t acc = tmem + static_cast<uint32_t>((p * a.acc_bufs + tr.idx) * a.acc_cols);
        if (valid) {
          mbar_wait(&s_tempty[tb], tr.phase ^ 1u, a.err, 2);
          tc_fence_after();
        }
        if (p == 0) TLC(it, 2);
        int kb = 0;
        while (kb < nkb) {
          if (valid) {
            mbar_wait(&s_afull[ar.idx], ar.phase, a.err, 2);
            tc_fence_after();
          }
          if (p == 0 && kb == 0) TLC(it, 3);
          const uint32_t slot16 = smem_u32(a_base + static_cast<size_t>(ar.idx) * a.a_slot_bytes) >> 4;
          const KB3 K0 = prog.kb[kb];
          const int cnt = static_cast<int>(K0.b_bytes >> 24);
          const bool lead_lane = elect_one();
          const int n_stage_units = a.resident ? 0 : (cnt + a.b_unit - 1) / a.b_unit;
          if (lead_lane) {
            if (a.resident) {
              if (valid) {
                if (K0.nk == 4)
                  issue_pair<4>(prog, kb, cnt, acc, slot16, b_base16);
                else if (K0.nk == 2)
                  issue_pair<2>(prog, kb, cnt, acc, slot16, b_base16);
                else
                  issue_pair<1>(prog, kb, cnt, acc, slot16, b_base16);
              }
            } else {
              for (int u = 0; u < cnt; u += a.b_unit) {
                const int nu = min(a.b_unit, cnt - u);
                mbar_wait(&s_bfull[br.idx], br.phase, a.err, 2);
                tc_fence_after();
                if (valid) {
                  const uint32_t bofs = b_base16 + static_cast<uint32_t>((br.idx * a.b_stage_bytes) >> 4);
                  if (K0.nk == 4)
                    issue_pair<4>(prog, kb + u, nu, acc, slot16, bofs);
                  else if (K0.nk == 2)
                    issue_pair<2>(prog, kb + u, nu, acc, slot16, bofs);
                  else
                    issue_pair<1>(prog, kb + u, nu, acc, slot16, bofs);
                  umma_commit_2cta(&s_bempty[br.idx]);
                } else {
                  // no second unit in the last round: release the stage unused, in both CTAs
                  mbar_arrive(&s_bempty[br.idx]);
                  mbar_arrive_cluster(mapa_u32(smem_u32(&s_bempty[br.idx]), 1));
                }
                br.advance(a.b_stages);
              }
            }
            if (valid) umma_commit_2cta(&s_aempty[ar.idx]);
          }
          const int kb_end = kb + cnt;
          __syncwarp();
          if (!lead_lane)
            for (int i = 0; i < n_stage_units; ++i) br.advance(a.b_stages);
          kb = kb_end;
          ar.advance(a.a_slots);
          ar.advance(a.a_slots);
        }
        if (valid && elect_one()) umma_commit_2cta(&s_tfull[tb]);
        __syncwarp();
        TLC(it, p ? 7 : 4);
        tr.advance(a.acc_bufs);
      }
    }
  } else {
    // ---- epilogue group p: this CTA's tile of every unit of pipeline p (or of both pipelines in solo mode) -------
    griddep_wait();
    const int p = warp >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int lx = row & (kTile2W - 1);
    const int ly = row >> 3;
    TmaStoreCtx ts;
    ts.map = &map_out;
    ts.sbc = a.store_sbc;
    ts.nbuf = a.store_sbc ? min(4, kStageBytesPerWarp / (64 * a.store_sbc)) : 1;
    ts.buf = 0;
    ts.stage = b_base + (a.resident ? half_image : static_cast<uint32_t>(a.b_stages * a.b_stage_bytes)) +
               static_cast<size_t>(warp) * kStageBytesPerWarp;
    const int n_jobs = a.solo ? 2 : 1;
    const int g_half = e.n_groups >> 1;
    for (int it = 0;; ++it) {
      const int u0 = first_unit + 2 * it * unit_step;
      if (u0 >= n_units) break;
      for (int job = 0; job < n_jobs; ++job) {
        const int pp = a.solo ? job : p;
        const int unit = u0 + pp * unit_step;
        if (unit >= n_units) continue;
        const int tile = 2 * unit + static_cast<int>(rank);
        const int bi = (a.acc_bufs == 2) ? (it & 1) : 0;
        const uint32_t ph = static_cast<uint32_t>((a.acc_bufs == 2) ? (it >> 1) : it) & 1u;
        const int tb = pp * 2 + bi;
        const int g_begin = a.solo ? p * g_half : 0;
        const int g_end = a.solo ? g_begin + g_half : e.n_groups;
        const int b = tile / tiles_per_img;
        const int t2 = tile - b * tiles_per_img;
        const int y = (t2 / a.tiles_x) * kTile2H + ly;
        const int x = (t2 % a.tiles_x) * kTile2W + lx;
        const bool valid = (x < a.W) && (y < a.H);
        if (kStageTe) {
          __syncwarp();
          const float* src = e.te + static_cast<size_t>(__ldg(e.trow + b)) * e.te_stride + e.te_off + oc_off;
          for (int c = lane * 4; c < a.n_sub; c += 128)
            *reinterpret_cast<float4*>(&s_te[warp][c]) = __ldg(reinterpret_cast<const float4*>(src + c));
          __syncwarp();
        }
        mbar_wait(&s_tfull[tb], ph, a.err, 3);
        tc_fence_after();
        if (threadIdx.x == 0) TLC(it, 5);
        const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>((pp * a.acc_bufs + bi) * a.acc_cols);
        ts.x0 = (t2 % a.tiles_x) * kTile2W;
        ts.y0 = (t2 / a.tiles_x) * kTile2H + q * 4;
        ts.b = b;
        conv_epilogue_std_ct<FL>(e, taddr, x, y, b, valid, a.W, a.H, a.n_sub, oc_off, s_par, s_te[kStageTe ? warp : 0],
                                 a.store_sbc ? &ts : nullptr, g_begin, g_end);
        tc_fence_before();
        mbar_arrive_cluster(mapa_u32(smem_u32(&s_tempty[tb]), 0));  // the issuers of CTA 0 reuse the accumulator
        if (threadIdx.x == 0) TLC(it, 6);
      }
    }
    if (a.store_sbc && lane == 0) bulk_wait_read<0>();
  }

  // ---- teardown: neither CTA may leave while its peer can still signal it or run MMAs on its memory ------
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kIssuer0) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem, static_cast<uint32_t>(a.tmem_cols));
  }
  // The pair-wide deallocation is a collective of one warp of each CTA. Neither CTA may retire before both have
  // executed it: the next kernel's CTA on that SM could otherwise enter the collective in its place (observed as an
  // illegal-address fault whenever one CTA-pair launch directly followed another).
  cluster_sync_all();
  if ((a.timeline & 4) && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(&a.span_buf[(a.launch_id & 63) * 2 + 1], t);
  }
}

#define DRS_GEMM2C_VARIANTS(X) \
  X(0)                         \
  X(F_NOSCALE)                 \
  X(F_RELU)                    \
  X(F_RELU | F_TE)             \
  X(F_RELU | F_DUAL_PRE)       \
  X(F_ROWSCALE)                \
  X(F_RELU | F_PRE)

static constexpr int kMaxDynSmem2c = 214 * 1024;

int conv_gemm2c_set_smem_limits() {
  cudaError_t e = cudaSuccess;
#define X(FL)            \
  if (e == cudaSuccess) \
    e = cudaFuncSetAttribute(conv_gemm2c_kernel<(FL)>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem2c);
  DRS_GEMM2C_VARIANTS(X)
#undef X
  return static_cast<int>(e);
}

bool conv_gemm2c_supports(int epi_kind, int flags) {
  if (epi_kind != EPI_STD) return false;
#define X(FL) \
  if (flags == (FL)) return true;
  DRS_GEMM2C_VARIANTS(X)
#undef X
  return false;
}

// Largest number of co-resident CTA pairs for this launch shape (0 on failure).
int conv_gemm2c_max_clusters(int flags, size_t smem_bytes) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2, 1, 1);
  cfg.blockDim = dim3(kGemm2Threads, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaErrorInvalidValue;
#define X(FL) \
  if (flags == (FL)) e = cudaOccupancyMaxActiveClusters(&n, conv_gemm2c_kernel<(FL)>, &cfg);
  DRS_GEMM2C_VARIANTS(X)
#undef X
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int launch_conv_gemm2c(const CUtensorMap& map0, const CUtensorMap& map1, const CUtensorMap& map_out,
                       const CUtensorMap& map_w, const Conv2Args& args, const Conv2Prog& prog, int grid,
                       size_t smem_bytes, cudaStream_t stream) {
  static const bool no_pdl = (getenv("DRS_V2_NO_PDL") != nullptr);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(kGemm2Threads, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  cudaError_t err = cudaErrorInvalidValue;
#define X(FL) \
  if (args.epi.flags == (FL)) err = cudaLaunchKernelEx(&cfg, conv_gemm2c_kernel<(FL)>, map0, map1, map_out, map_w, args, prog);
  DRS_GEMM2C_VARIANTS(X)
#undef X
  if (err != cudaSuccess) return static_cast<int>(err);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace drs
