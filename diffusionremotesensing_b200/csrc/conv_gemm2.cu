// Persistent halo-tile tcgen05 convolution kernel (see conv_gemm2.cuh for the model).
//
// CTA = 352 threads, persistent, looping over PAIRS of pixel tiles. One thread needs ~130-160 cycles of bookkeeping and
// issue per K-block while the tensor pipe needs 40-64 cycles per small-N MMA of this network, so one CTA runs two MMA
// issuers and two epilogue groups, one per tile of the pair and per TMEM accumulator buffer; both issuers walk the
// same K-block program in step, so a streamed weight unit is fetched once per pair.
//   warps 0..3   epilogue of tile 0 of the pair (tcgen05.ld, fused affine / ReLU / adds, staged TMA stores)
//   warps 4..7   epilogue of tile 1 (solo mode: both groups drain every tile, half of the column groups each)
//   warp 8, 9    tcgen05.mma issuers of tile 0 / tile 1 (one elected lane each); warp 8 owns the TMEM allocation
//   warp 10      producer: resident weights once; per tile one 5-D TMA box per A sub-tile (halo tile of one channel
//                block) and, when the weights are streamed, one bulk copy per UNIT of up to b_unit K-blocks and pair
// The single-lane roles have the highest warp ids: the sub-partition arbiter favours the highest id among eligible
// warps, so ALU-heavy epilogue warps never starve the issuer they wait for. Waits park the thread in hardware
// (mbarrier.try_wait with a suspend-time hint) and are bounded: a pipeline bug sets the error word instead of hanging.
// Rings: A slots (full/empty, filled in the order (sub-tile, tile-of-pair)), B stages (full / empty-by-both-issuers,
// streamed mode), TMEM buffer p (full/empty) for tile p of the pair.
// The kernel is launched as a programmatic dependent of its predecessor: everything up to griddep_wait() (barrier
// init, TMEM allocation, parameter staging, the resident weight image) only touches data that never changes during
// sampling and overlaps the previous layer's tail.
#include <stdlib.h>
#include <string.h>

#include "conv_epilogue.cuh"
#include "conv_gemm2.cuh"
#include "ptx.cuh"

namespace drs {

// Debug timeline (DRS_V2_TIMELINE=1): CTA 0 records SM-clock stamps of its first tile pairs, 8 slots per pair:
// 0 producer pair start, 1 producer last issue, 2 MMA(0) after tmem-empty wait, 3 MMA(0) after first A-full wait,
// 4 MMA(0) after last issue, 5 epilogue(0) after tmem-full wait, 6 epilogue(0) done, 7 MMA(1) after last issue.
__device__ long long g_timeline[64 * 8];
// DRS_V2_TIMELINE bit 2: per-launch [first CTA entry, last CTA exit] in globaltimer nanoseconds, by launch id
__device__ unsigned long long g_span[64][2];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
long long* conv_gemm2_timeline_dev() {
  long long* p = nullptr;
  cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_timeline);
  return p;
}
#define TL(tile_no, slot)                                                                       \
  do {                                                                                          \
    if ((a.timeline & 1) && blockIdx.x == 0 && (tile_no) < 64) g_timeline[(tile_no) * 8 + (slot)] = clock64(); \
  } while (0)

struct RingPos {
  int idx;
  uint32_t phase;
  __device__ __forceinline__ void advance(int n) {
    if (++idx == n) {
      idx = 0;
      phase ^= 1u;
    }
  }
};

// All K-blocks of one A sub-tile with resident weights: counted loop, NK slices per K-block, no branch but the
// back edge (every data-dependent uniform branch costs the single issuing lane ~20 cycles).
template <int NK>
__device__ __forceinline__ void issue_resident(const Conv2Prog& prog, int kb, int cnt, uint32_t acc, uint32_t slot16,
                                               uint32_t b_base16) {
  // the descriptor upper halves are constant over a sub-tile; the per-K-block part is one 128-bit load
  const uint32_t a_hi = prog.kb[kb].a_hi, b_hi = prog.kb[kb].b_hi;
  for (int i = 0; i < cnt; ++i) {
    const uint4 q = *reinterpret_cast<const uint4*>(&prog.kb[kb + i]);  // a_lo, b_lo, col | nk | flags, idesc
    const uint32_t a_lo = q.x + slot16;
    const uint32_t b_lo = q.y + b_base16;
    const uint32_t d = acc + (q.z & 0xFFFFu);
    umma_bf16_split(d, a_lo, a_hi, b_lo, b_hi, q.w, ((q.z >> 24) & KB2_INIT) ? 0u : 1u);
#pragma unroll
    for (int k = 1; k < NK; ++k) umma_bf16_split(d, a_lo + 2u * k, a_hi, b_lo + 2u * k, b_hi, q.w, 1u);
  }
}

constexpr int kMmaWarp0 = 8;      // issuer of tile 0 of a pair (tile 1: kMmaWarp0 + 1)
constexpr int kProducerWarp = 10;

// FL: EPI_STD flag word known at compile time (the combinations the UNets use: DRS_GEMM2_VARIANTS), or -1 for the
// generic epilogue.
template <int EPI, int FL>
__global__ void __launch_bounds__(kGemm2Threads)
conv_gemm2_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                  const __grid_constant__ CUtensorMap map_out, const __grid_constant__ Conv2Args a,
                  const __grid_constant__ Conv2Prog prog) {
  extern __shared__ uint8_t dyn_smem[];
  if ((a.timeline & 1) && blockIdx.x == 0 && threadIdx.x == 0) g_timeline[504] = clock64();
  if ((a.timeline & 4) && threadIdx.x == 0) atomicMin(&g_span[a.launch_id & 63][0], globaltimer_ns());
  __shared__ __align__(8) uint64_t s_afull[kMaxASlots], s_aempty[kMaxASlots];
  __shared__ __align__(8) uint64_t s_bfull[kMaxBStages], s_bempty[kMaxBStages];
  __shared__ __align__(8) uint64_t s_tfull[4], s_tempty[4];  // [tile of pair][accumulator buffer]
  __shared__ __align__(8) uint64_t s_wready;
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_par[4][kMaxN];
  constexpr bool kStageTe = (EPI == EPI_STD && FL >= 0 && (FL & F_TE));
  __shared__ __align__(16) float s_te[kStageTe ? 8 : 1][kStageTe ? kMaxN : 4];  // per epilogue warp

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform role index
  const int lane = threadIdx.x & 31;
  const EpiArgs& e = a.epi;

  const uint32_t dyn_u32 = smem_u32(dyn_smem);
  uint8_t* const a_base = dyn_smem + ((1024u - (dyn_u32 & 1023u)) & 1023u);
  uint8_t* const b_base = a_base + static_cast<size_t>(a.a_slots) * a.a_slot_bytes;

  const int split = blockIdx.x % a.nsplit;
  const int first_tile = blockIdx.x / a.nsplit;
  const int tile_step = gridDim.x / a.nsplit;
  const int oc_off = split * a.n_sub;
  const int nkb = a.nkb;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const uint8_t* const w_image = a.wpack + a.w_split_off + static_cast<size_t>(split) * a.w_split_bytes;

  // ---- one-time setup --------------------------------------------------------------------------
  load_epilogue_params<EPI>(e, a.n_sub, oc_off, s_par, threadIdx.x, kGemm2Threads);
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&map0);
    tma_prefetch_desc(&map1);
    if (a.store_sbc) tma_prefetch_desc(&map_out);
    for (int s = 0; s < a.a_slots; ++s) {
      mbar_init(&s_afull[s], 1);
      mbar_init(&s_aempty[s], 1);
    }
    for (int s = 0; s < a.b_stages; ++s) {
      mbar_init(&s_bfull[s], 1);
      mbar_init(&s_bempty[s], 2);  // released by both issuers
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&s_tfull[s], 1);
      mbar_init(&s_tempty[s], a.solo ? 256 : 128);
    }
    mbar_init(&s_wready, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp0) {
    tmem_alloc(&s_tmem_base, static_cast<uint32_t>(a.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  if ((a.timeline & 1) && blockIdx.x == 0 && threadIdx.x == 0) g_timeline[505] = clock64();
  // Everything above (and the resident weight image below) only touches parameters that never change during
  // sampling, so with a programmatic dependent launch it overlaps the tail of the previous layer.
  griddep_launch();

  if (warp == kProducerWarp) {
    // ---- producer (whole warp walks the program, one elected lane issues) -------------------------
    if (a.resident && elect_one()) {
      mbar_expect_tx(&s_wready, a.w_split_bytes);
      for (uint32_t off = 0; off < a.w_split_bytes; off += 16384u) {
        const uint32_t n = min(16384u, a.w_split_bytes - off);
        bulk_load(b_base + off, w_image + off, n, &s_wready);
      }
    }
    __syncwarp();
    griddep_wait();  // activations of the previous layer
    if ((a.timeline & 1) && blockIdx.x == 0 && lane == 0) g_timeline[506] = clock64();
    RingPos ar{0, 0}, br{0, 0};
    int pno = 0;
    for (int tile0 = first_tile; tile0 < a.n_tiles; tile0 += 2 * tile_step, ++pno) {
      int x0[2], y0[2], bb[2];
      bool valid[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int tile = tile0 + p * tile_step;
        valid[p] = tile < a.n_tiles;
        bb[p] = tile / tiles_per_img;
        const int t2 = tile - bb[p] * tiles_per_img;
        y0[p] = (t2 / a.tiles_x) * kTile2H;
        x0[p] = (t2 % a.tiles_x) * kTile2W;
      }
      int st = 0, sub_first = 0, sub_cnt = 0;
      TL(pno, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const KB3 K = prog.kb[kb];
        if (K.flags & KB2_FIRST) {
          const SubTile T = prog.st[st++];
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            if (valid[p]) {
              mbar_wait(&s_aempty[ar.idx], ar.phase ^ 1u, a.err, 1);
              if (elect_one()) {
                mbar_expect_tx(&s_afull[ar.idx], T.bytes);
                tma_load_5d(a_base + static_cast<size_t>(ar.idx) * a.a_slot_bytes, T.src ? &map1 : &map0,
                            &s_afull[ar.idx], T.c, x0[p] + T.dx0, 0, y0[p] + T.dy0, bb[p]);
              }
              __syncwarp();
            }
            ar.advance(a.a_slots);  // the slot sequence is (sub-tile, tile-of-pair) whether or not tile 1 exists
          }
        }
        if (!a.resident) {
          if (K.flags & KB2_FIRST) {
            sub_first = kb;
            sub_cnt = static_cast<int>(K.b_bytes >> 24);
          }
          if ((kb - sub_first) % a.b_unit == 0) {
            // one ring stage = the next (up to) b_unit weight tiles of this sub-tile, contiguous in the image
            const uint32_t tile_pad = ((K.b_bytes & 0xFFFFFFu) + 1023u) & ~1023u;
            const uint32_t bytes = static_cast<uint32_t>(min(a.b_unit, sub_first + sub_cnt - kb)) * tile_pad;
            mbar_wait(&s_bempty[br.idx], br.phase ^ 1u, a.err, 1);
            if (elect_one()) {
              mbar_expect_tx(&s_bfull[br.idx], bytes);
              uint8_t* dst = b_base + static_cast<size_t>(br.idx) * a.b_stage_bytes;
              for (uint32_t off = 0; off < bytes; off += 16384u)
                bulk_load(dst + off, w_image + K.b_off + off, min(16384u, bytes - off), &s_bfull[br.idx]);
            }
            __syncwarp();
            br.advance(a.b_stages);
          }
        }
      }
      TL(pno, 1);
    }
  } else if (warp == kMmaWarp0 || warp == kMmaWarp0 + 1) {
    // ---- MMA issuer of tile p of every pair ------------------------------------------------------------
    const int p = warp - kMmaWarp0;
    if (a.resident) mbar_wait(&s_wready, 0, a.err, 2);
    RingPos ar{0, 0}, br{0, 0};
    if (p) ar.advance(a.a_slots);
    const uint32_t b_base16 = smem_u32(b_base) >> 4;
    RingPos tr{0, 0};  // accumulator buffer of this pipeline (a.acc_bufs per tile of the pair)
    int pno = 0;
    for (int tile0 = first_tile; tile0 < a.n_tiles; tile0 += 2 * tile_step, ++pno) {
      const bool valid = (tile0 + p * tile_step) < a.n_tiles;
      const int tb = p * 2 + tr.idx;
      const uint32_t acc = tmem + static_cast<uint32_t>((p * a.acc_bufs + tr.idx) * a.acc_cols);
      if (valid) {
        mbar_wait(&s_tempty[tb], tr.phase ^ 1u, a.err, 2);
        tc_fence_after();
      }
      if (p == 0) TL(pno, 2);
      // One warp-level step per A sub-tile: the warp waits for its halo tile, then one elected lane issues every
      // MMA of every tap that reads it back to back (streamed weights are waited for by the issuing lane itself).
      int kb = 0;
#ifdef DRS_EPI_TRACE
      int itl_n = 0;  // traced build: streamed weight units of pair 1 of CTA 0 stamped so far
#endif
      while (kb < nkb) {
        if (valid) {
          mbar_wait(&s_afull[ar.idx], ar.phase, a.err, 2);
          tc_fence_after();
        }
        if (p == 0 && kb == 0) TL(pno, 3);
        const uint32_t slot16 = smem_u32(a_base + static_cast<size_t>(ar.idx) * a.a_slot_bytes) >> 4;
        const KB3 K0 = prog.kb[kb];
        const int cnt = static_cast<int>(K0.b_bytes >> 24);  // K-blocks that read this sub-tile
        const bool leader = elect_one();
        const int n_units = a.resident ? 0 : (cnt + a.b_unit - 1) / a.b_unit;
        if (leader) {
          if (a.resident) {
            // hot path: counted loop, K = 16 slice count fixed per sub-tile, no data-dependent branch
            if (valid) {
              if (K0.nk == 4)
                issue_resident<4>(prog, kb, cnt, acc, slot16, b_base16);
              else if (K0.nk == 2)
                issue_resident<2>(prog, kb, cnt, acc, slot16, b_base16);
              else
                issue_resident<1>(prog, kb, cnt, acc, slot16, b_base16);
            }
          } else {
            // streamed weights: one ring stage per unit of up to b_unit K-blocks (b_lo = offset inside the unit)
            for (int u0 = 0; u0 < cnt; u0 += a.b_unit) {
              const int nu = min(a.b_unit, cnt - u0);
#ifdef DRS_EPI_TRACE
              const bool itl = (a.timeline & 1) && blockIdx.x == 0 && p == 0 && pno == 1 && itl_n < 10;
              if (itl) g_epi_trace[32 + itl_n * 3] = clock64();
#endif
              mbar_wait(&s_bfull[br.idx], br.phase, a.err, 2);
              tc_fence_after();
#ifdef DRS_EPI_TRACE
              if (itl) g_epi_trace[32 + itl_n * 3 + 1] = clock64();
#endif
              if (valid) {
                const uint32_t bofs = b_base16 + static_cast<uint32_t>((br.idx * a.b_stage_bytes) >> 4);
                if (K0.nk == 4)
                  issue_resident<4>(prog, kb + u0, nu, acc, slot16, bofs);
                else if (K0.nk == 2)
                  issue_resident<2>(prog, kb + u0, nu, acc, slot16, bofs);
                else
                  issue_resident<1>(prog, kb + u0, nu, acc, slot16, bofs);
                umma_commit(&s_bempty[br.idx]);
#ifdef DRS_EPI_TRACE
                if (itl) g_epi_trace[32 + itl_n++ * 3 + 2] = clock64();
#endif
              } else {
                mbar_arrive(&s_bempty[br.idx]);  // no tile 1 in the last pair: release the stage unused
              }
              br.advance(a.b_stages);
            }
          }
          if (valid) umma_commit(&s_aempty[ar.idx]);
        }
        const int kb_end = kb + cnt;
        __syncwarp();
        if (!leader)
          for (int i = 0; i < n_units; ++i) br.advance(a.b_stages);
        kb = kb_end;
        ar.advance(a.a_slots);
        ar.advance(a.a_slots);
      }
      if (valid && elect_one()) umma_commit(&s_tfull[tb]);
      __syncwarp();
      TL(pno, p ? 7 : 4);
      tr.advance(a.acc_bufs);
    }
  } else {
    // ---- epilogue group p: tile p of every pair ----------------------------------------------------------
    griddep_wait();  // gate maps / state read by the epilogue, and write-after-read on the output tensor
    const int p = warp >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int lx = row & (kTile2W - 1);
    const int ly = row >> 3;
    int pno = 0;
    TmaStoreCtx ts;
    ts.map = &map_out;
    ts.sbc = a.store_sbc;
    ts.nbuf = a.store_sbc ? min(4, kStageBytesPerWarp / (64 * a.store_sbc)) : 1;
    ts.buf = 0;
    // staging areas follow the weight region (both 1024-byte aligned)
    ts.stage = b_base + (a.resident ? a.w_split_bytes : static_cast<uint32_t>(a.b_stages * a.b_stage_bytes)) +
               static_cast<size_t>(warp) * kStageBytesPerWarp;
    // Normal mode: group p drains pipeline p (all column groups). Solo mode (one accumulator buffer per pipeline and
    // several column groups, i.e. the transposed convolutions): both groups drain EVERY tile, half of the column
    // groups each, so the epilogue of tile k runs twice as fast and overlaps the MMAs of tile k + 1 of the other
    // pipeline.
    const int n_jobs = a.solo ? 2 : 1;
    const int g_half = e.n_groups >> 1;
    // F_TR64: the CTA's 64 biases stay in registers (conv_epilogue_tr64); any other geometry takes the chunked path
    constexpr bool kTr64 = (EPI == EPI_STD && FL >= 0 && (FL & F_TR64) != 0);
    const bool tr64 = kTr64 && a.store_sbc == 64 && a.n_sub == 64 && e.oscale == 2;
    float tr_bias[kTr64 ? 64 : 1];
    if constexpr (kTr64) {
#pragma unroll
      for (int i = 0; i < 64; ++i) tr_bias[i] = s_par[1][i];
    }
    // fused 32-channel attention gate: the result conv's folded BatchNorm in registers (conv_epilogue_gate32)
    constexpr bool kGate = (EPI == EPI_STD && FL >= 0 && (FL & F_GATE) != 0);
    const bool gate32 = kGate && a.store_sbc == 32 && a.n_sub == 32 && a.nsplit == 1 && e.nvec == 32 && e.oscale == 2 &&
                        ((g_half & 1) == 0 || !a.solo);
    float gate_sc[kGate ? 32 : 1], gate_bi[kGate ? 32 : 1];
    if constexpr (kGate) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        gate_sc[i] = s_par[0][i];
        gate_bi[i] = s_par[1][i];
      }
    }
    for (int tile0 = first_tile; tile0 < a.n_tiles; tile0 += 2 * tile_step, ++pno) {
      for (int job = 0; job < n_jobs; ++job) {
        const int pp = a.solo ? job : p;  // pipeline whose tile is drained
        const int tile = tile0 + pp * tile_step;
        if (tile >= a.n_tiles) continue;
        // pipeline pp uses buffer (pno mod acc_bufs) for its pno-th tile
        const int bi = (a.acc_bufs == 2) ? (pno & 1) : 0;
        const uint32_t ph = static_cast<uint32_t>((a.acc_bufs == 2) ? (pno >> 1) : pno) & 1u;
        const int tb = pp * 2 + bi;
        const int g_begin = a.solo ? p * g_half : 0;
        const int g_end = a.solo ? g_begin + g_half : e.n_groups;
        const int b = tile / tiles_per_img;
        const int t2 = tile - b * tiles_per_img;
        const int y = (t2 / a.tiles_x) * kTile2H + ly;
        const int x = (t2 % a.tiles_x) * kTile2W + lx;
        const bool valid = (x < a.W) && (y < a.H);
        if (kStageTe) {
          // this tile's time-embedding row -> the warp's shared copy (latency hidden behind the accumulator wait)
          __syncwarp();
          const float* src = e.te + static_cast<size_t>(__ldg(e.trow + b)) * e.te_stride + e.te_off + oc_off;
          for (int c = lane * 4; c < a.n_sub; c += 128)
            *reinterpret_cast<float4*>(&s_te[warp][c]) = __ldg(reinterpret_cast<const float4*>(src + c));
          __syncwarp();
        }
        mbar_wait(&s_tfull[tb], ph, a.err, 3);
        tc_fence_after();
        if (threadIdx.x == 0) TL(pno, 5);
        const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>((pp * a.acc_bufs + bi) * a.acc_cols);
        ts.x0 = (t2 % a.tiles_x) * kTile2W;
        ts.y0 = (t2 / a.tiles_x) * kTile2H + q * 4;
        ts.b = b;
#ifdef DRS_EPI_TRACE
        if (threadIdx.x == 0 && blockIdx.x == 0) g_epi_trace_on = ((a.timeline & 1) && pno == 3) ? 1 : 0;
#endif
        if (!(a.timeline & 2)) {
          if constexpr (EPI == EPI_STD && FL >= 0 && (FL & F_TR64) != 0) {
            if (tr64)
              conv_epilogue_tr64(e, taddr, oc_off, tr_bias, &ts, g_begin, g_end);
            else
              conv_epilogue_std_ct<(FL & ~F_TR64)>(e, taddr, x, y, b, valid, a.W, a.H, a.n_sub, oc_off, s_par,
                                                  s_te[kStageTe ? warp : 0], a.store_sbc ? &ts : nullptr, g_begin, g_end);
          } else if constexpr (kGate) {
            if (gate32)
              conv_epilogue_gate32(e, taddr, x, y, b, valid, a.W, a.H, s_par, gate_sc, gate_bi, &ts, g_begin, g_end);
            else
              conv_epilogue_std_ct<FL>(e, taddr, x, y, b, valid, a.W, a.H, a.n_sub, oc_off, s_par,
                                       s_te[kStageTe ? warp : 0], a.store_sbc ? &ts : nullptr, g_begin, g_end);
          } else if constexpr (EPI == EPI_STD && FL >= 0)
            conv_epilogue_std_ct<FL>(e, taddr, x, y, b, valid, a.W, a.H, a.n_sub, oc_off, s_par,
                                     s_te[kStageTe ? warp : 0], a.store_sbc ? &ts : nullptr, g_begin, g_end);
          else
            conv_epilogue<EPI>(e, taddr, x, y, b, valid, a.W, a.H, a.n_sub, oc_off, s_par,
                               (EPI == EPI_STD && a.store_sbc) ? &ts : nullptr, g_begin, g_end);
        }
        tc_fence_before();
        mbar_arrive(&s_tempty[tb]);
        if (threadIdx.x == 0) TL(pno, 6);
      }
    }
    // the staging areas must outlive the TMA unit's reads
    if (EPI == EPI_STD && a.store_sbc && lane == 0) bulk_wait_read<0>();
  }

  // ---- teardown ----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp0) {
    tc_fence_after();
    tmem_dealloc(tmem, static_cast<uint32_t>(a.tmem_cols));
    if ((a.timeline & 1) && blockIdx.x == 0 && lane == 0) g_timeline[507] = clock64();
    if ((a.timeline & 4) && lane == 0) atomicMax(&g_span[a.launch_id & 63][1], globaltimer_ns());
    // bit 3: exit time of every CTA (ns since the first CTA of this launch entered; needs bit 2 as well)
    if ((a.timeline & 8) && lane == 0 && blockIdx.x < 512)
      g_timeline[blockIdx.x] = static_cast<long long>(globaltimer_ns() - g_span[a.launch_id & 63][0]);
  }
}

static constexpr int kMaxDynSmem2 = 214 * 1024;
// the transposed-convolution variant (4.6 KiB static) may take what is left of the 227 KiB of a CTA
static constexpr int kMaxDynSmem2Tr = 222 * 1024;

// every instantiation: (EPI, FL)
#define DRS_GEMM2_VARIANTS(X)                     \
  X(EPI_STD, -1)                                  \
  X(EPI_STD, 0)                                   \
  X(EPI_STD, F_NOSCALE)                           \
  X(EPI_STD, F_NOSCALE | F_TR64)                  \
  X(EPI_STD, F_RELU)                              \
  X(EPI_STD, F_RELU | F_TE)                       \
  X(EPI_STD, F_RELU | F_TE | F_DUAL_POST)         \
  X(EPI_STD, F_RELU | F_DUAL_PRE)                 \
  X(EPI_STD, F_ROWSCALE)                          \
  X(EPI_STD, F_GATE | F_ROWSCALE)                 \
  X(EPI_STD, F_RELU | F_PRE)                      \
  X(EPI_PSI, -1)                                  \
  X(EPI_OUT, -1)

int conv_gemm2_set_smem_limits() {
  cudaError_t e = cudaSuccess;
#define X(EPI, FL)                                                                                              \
  if (e == cudaSuccess)                                                                                         \
    e = cudaFuncSetAttribute(conv_gemm2_kernel<EPI, (FL)>, cudaFuncAttributeMaxDynamicSharedMemorySize,                   \
                             ((FL) >= 0 && ((FL) & F_TR64)) ? kMaxDynSmem2Tr : kMaxDynSmem2);
  DRS_GEMM2_VARIANTS(X)
#undef X
  return static_cast<int>(e);
}

unsigned long long* conv_gemm2_span_dev() {
  unsigned long long* p = nullptr;
  cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_span);
  return p;
}

int conv_gemm2_spans(unsigned long long* host, int reset) {
  if (reset) {
    unsigned long long init[64][2];
    for (auto& r : init) {
      r[0] = ~0ull;
      r[1] = 0ull;
    }
    return static_cast<int>(cudaMemcpyToSymbol(g_span, init, sizeof(init)));
  }
  return static_cast<int>(cudaMemcpyFromSymbol(host, g_span, sizeof(unsigned long long) * 128));
}

int conv_gemm2_read_timeline(long long* host, int n) {
  if (n > 64 * 8) n = 64 * 8;
#ifdef DRS_EPI_TRACE
  // debug builds: entries 256.. carry the per-chunk epilogue stamps of thread 0 of CTA 0 in its fourth tile pair
  cudaError_t e = cudaMemcpyFromSymbol(host, g_timeline, n * sizeof(long long));
  if (e == cudaSuccess && n >= 320) e = cudaMemcpyFromSymbol(host + 256, g_epi_trace, 64 * sizeof(long long));
  return static_cast<int>(e);
#else
  return static_cast<int>(cudaMemcpyFromSymbol(host, g_timeline, n * sizeof(long long)));
#endif
}

int launch_conv_gemm2(int epi_kind, const CUtensorMap& map0, const CUtensorMap& map1, const CUtensorMap& map_out,
                      const Conv2Args& args, const Conv2Prog& prog, int grid, size_t smem_bytes, cudaStream_t stream) {
  dim3 g(static_cast<unsigned>(grid), 1, 1);
  dim3 block(kGemm2Threads, 1, 1);
  static const bool generic = (getenv("DRS_V2_GENERIC_EPILOGUE") != nullptr);
  const int fl = (epi_kind == EPI_STD && !generic) ? args.epi.flags : -1;
  bool done = false;
  static const bool no_pdl = (getenv("DRS_V2_NO_PDL") != nullptr);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = g;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  cudaError_t err = cudaSuccess;
#define X(EPI, FL)                                                                                      \
  if (!done && epi_kind == EPI && fl == (FL)) {                                                         \
    err = cudaLaunchKernelEx(&cfg, conv_gemm2_kernel<EPI, (FL)>, map0, map1, map_out, args, prog);      \
    done = true;                                                                                        \
  }
  DRS_GEMM2_VARIANTS(X)
#undef X
  if (!done) {
    if (epi_kind != EPI_STD) return static_cast<int>(cudaErrorInvalidValue);
    err = cudaLaunchKernelEx(&cfg, conv_gemm2_kernel<EPI_STD, -1>, map0, map1, map_out, args, prog);
  }
  if (err != cudaSuccess) return static_cast<int>(err);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace drs
