// Persistent halo-tile implicit-GEMM convolution (second-generation kernel), shared host/device declarations.
//
// conv_gemm.cu re-reads the 128-pixel activation tile from L2 once per filter tap, which makes the full-resolution
// layers L2->SM bandwidth bound. Here a CTA loads, per 64/32/16-channel block, ONE halo tile
// ((8 + halo) x (16 + halo) pixels, pixel-major, hardware swizzled) and every filter tap is an MMA whose A
// descriptor starts at a different pixel of that tile:
//     row m = 8 g + r of the MMA  <->  pixel (x0 + r, y0 + g);  address = start + g * SBO + r * pixel_bytes
// with SBO = halo_width * pixel_bytes, so a tap shift (dx, dy) is just start += (dy * halo_width + dx) * pixel_bytes.
// CTAs are persistent (static round-robin over pixel tiles), keep the layer's weights resident in shared memory when
// they fit (else stream them through a ring per tile) and double-buffer the TMEM accumulator so the epilogue of tile
// i overlaps the MMAs of tile i + 1. The K-block program travels as a __grid_constant__ kernel parameter, so the
// single issuing thread reads it with uniform constant loads (no shared-memory round trip per MMA).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_gemm.cuh"

namespace drs {

constexpr int kTile2W = 8;    // pixels per accumulator row group (fixed by the UMMA 8-row core-matrix group)
constexpr int kTile2H = 16;   // row groups per M = 128 tile
constexpr int kGemm2Threads = 352;  // 8 epilogue warps + 2 MMA issuers + 1 producer
constexpr int kMaxSubTiles = 16;
constexpr int kMaxASlots = 16;
constexpr int kMaxBStages = 16;

struct __align__(16) KB3 {  // one K-block = one filter tap of one channel block: nk MMAs of K = 16
  // first 16 bytes: everything the issuing lane needs per K-block (one 128-bit uniform load)
  uint32_t a_lo;      // lower half of the A descriptor relative to the A slot: (tap offset / 16) | LBO field
  uint32_t b_lo;      // lower half of the B descriptor relative to the weight region / ring stage | LBO field
  uint16_t col;       // accumulator column offset inside one TMEM buffer
  uint8_t nk;         // K = 16 slices (ck / 16)
  uint8_t flags;      // KB2_*
  uint32_t idesc;     // tcgen05 instruction descriptor (M = 128, N, bf16 x bf16 -> fp32)
  // second 16 bytes: constant over a sub-tile (descriptor upper halves) or used by the producer only
  uint32_t a_hi;      // upper half of the A descriptor (SBO = halo row pitch, version, swizzle mode)
  uint32_t b_hi;      // upper half of the B descriptor
  uint32_t b_off;     // byte offset of the weight tile inside one split's image (tiles are 1 KiB aligned)
  uint32_t b_bytes;   // bits 0..23: weight tile bytes (n * ck * 2); bits 24..31 (FIRST record): K-blocks of the sub-tile
};
static_assert(sizeof(KB3) == 32, "KB3 must be 32 bytes");

enum : uint8_t {
  KB2_INIT = 1,   // first K-block writing these accumulator columns: overwrite
  KB2_FIRST = 2,  // first K-block of an A sub-tile: acquire the next A slot
  KB2_LAST = 4,   // last K-block of an A sub-tile: release the slot
};

struct SubTile {
  int32_t c;        // coordinate 0 of the TMA box (channel block, plus px * C for stride-2 views)
  int16_t dx0, dy0; // halo origin relative to the tile origin (coordinates 1 and 3)
  uint32_t bytes;   // box bytes = ck * 2 * halo_w * npy * halo_h
  uint8_t src;      // tensor map 0 / 1
  uint8_t pad[3];
};
static_assert(sizeof(SubTile) == 16, "SubTile must be 16 bytes");

struct Conv2Prog {
  KB3 kb[kMaxKBlocks];
  SubTile st[kMaxSubTiles];
};

// Host-side encoders of the descriptor halves the kernel does not need to recompute per K-block.
inline uint32_t umma_desc_hi(uint32_t row_bytes, uint32_t sbo_bytes) {
  const uint32_t layout = (row_bytes == 128) ? 2u : (row_bytes == 64) ? 4u : 6u;  // SWIZZLE_128B / 64B / 32B
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
}
inline uint32_t umma_idesc_host(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

struct Conv2Args {
  const uint8_t* wpack;     // weight blob
  uint32_t w_split_off;     // byte offset of split 0's weight image inside wpack
  uint32_t w_split_bytes;   // bytes of one split's image (multiple of 1024)
  int nkb, n_sub_tiles;
  int resident;             // 1: the split's image is loaded once per CTA, 0: tiles stream through the B ring
  int W, H, B;              // output pixel grid
  int tiles_x, tiles_y, n_tiles;
  int a_slots, a_slot_bytes;
  int b_stages, b_stage_bytes;
  int b_unit;               // streamed weights: K-blocks per ring stage (units never straddle sub-tiles)
  int tmem_cols;            // allocation (power of two)
  int acc_cols;             // columns one tile's accumulators use
  int acc_bufs;             // 1 or 2
  int n_sub, nsplit;
  int cg2_half_tile_bytes;  // CTA-pair kernel: bytes of half a (padded) weight tile
  int solo;                 // 1: both epilogue groups drain every tile, half of the column groups each
  int store_sbc;            // EPI_STD: channels per TMA-store sub-box (0: per-thread global stores, no staging)
  unsigned long long* span_buf;  // debug: device address of the span table (CTA-pair kernel)
  long long* timeline_buf;  // debug: device buffer of the stamps (CTA-pair kernel)
  int launch_id;            // index of the launch inside the plan (debug spans)
  int timeline;             // debug: CTA 0 records clock stamps (DRS_V2_TIMELINE)
  int* err;
  EpiArgs epi;
};

constexpr int kStageBytesPerWarp = 4096;  // staging area of one epilogue warp (TMA-store path)
constexpr int kStageBytes = 8 * kStageBytesPerWarp;

int launch_conv_gemm2(int epi_kind, const CUtensorMap& map0, const CUtensorMap& map1, const CUtensorMap& map_out,
                      const Conv2Args& args,
                      const Conv2Prog& prog, int grid, size_t smem_bytes, cudaStream_t stream);
int conv_gemm2_set_smem_limits();
// CTA-pair variant (conv_gemm2c.cu)
int launch_conv_gemm2c(const CUtensorMap& map0, const CUtensorMap& map1, const CUtensorMap& map_out,
                       const CUtensorMap& map_w, const Conv2Args& args, const Conv2Prog& prog, int grid,
                       size_t smem_bytes, cudaStream_t stream);
int conv_gemm2c_set_smem_limits();
bool conv_gemm2c_supports(int epi_kind, int flags);
int conv_gemm2c_max_clusters(int flags, size_t smem_bytes);
int conv_gemm2_read_timeline(long long* host, int n);
int conv_gemm2_spans(unsigned long long* host, int reset);  // debug: [launch id][entry, exit] globaltimer ns
unsigned long long* conv_gemm2_span_dev();
long long* conv_gemm2_timeline_dev();  // device address of the debug timeline buffer (shared with conv_gemm2c.cu)

}  // namespace drs
