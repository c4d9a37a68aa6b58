// Row-streaming tcgen05 convolution kernel for the wide, low-channel layers (third-generation kernel), shared
// host/device declarations.
//
// conv_gemm2 issues one MMA per filter tap; every tcgen05.mma M=128 x N x K=16 reads its 4 KiB A operand from shared
// memory whatever N is, so a layer with N = 32 output channels runs at (4096 + 32 N) / 128 = 40 cycles per MMA against
// a 16-cycle tensor floor (profiles/mma_rate_r2.txt). The activation tile is the same for the three taps of a filter
// column; only the OUTPUT row they feed differs. Here
//   * a tile is 128 consecutive pixels of ONE image row, accumulator row m <-> pixel x0 + m, so the three vertical taps
//     of an input row feed the accumulators of three different output rows AT THE SAME TMEM LANES;
//   * the accumulators of consecutive output rows are adjacent column ranges of a TMEM ring (slot = running output-row
//     count mod S), so ONE MMA with the three vertical taps' weights stacked along N (dy = +1 | 0 | -1) adds an input
//     row into the accumulators of output rows y-1, y, y+1: N = 96 (192) instead of three N = 32 (64) MMAs, one A read
//     instead of three -> 56 (80) cycles per three taps instead of 120 (144);
//   * a pipeline streams down a column strip: every input row is loaded once (one TMA box of 130 pixels per channel block;
//     the horizontal taps are A-descriptor start offsets of 0 / 1 / 2 pixels), an output row is complete once the
//     input row below it has been issued, and the epilogue drains it while the issuer is several rows ahead.
// The flattened (image, column strip, row) index space is cut into equal contiguous ranges, one per pipeline; a range
// that starts / ends inside a strip re-reads one halo row on each side. One issuing thread needs ~1100 cycles of
// barrier round trips per row on top of ~45 cycles per MMA, more than the tensor pipe needs for the row, so a CTA runs
// TWO pipelines (own range, A ring, TMEM ring, issuer and producer warp, epilogue group): one issuer's bubbles are
// the other's MMAs. The MMA list of a row is a table in the kernel parameters (one 128-bit record per MMA).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_gemm.cuh"

namespace drs {

constexpr int kRowTile = 128;          // pixels per tile = TMEM lanes
constexpr int kRowPipes = 2;           // independent pipelines per CTA (own tile range, A ring, TMEM ring, issuer)
constexpr int kRowEpiWarps = 8;        // two groups of four: one per pipeline, or alternating rows of a single pipeline
constexpr int kRowThreads = (kRowEpiWarps + 2 * kRowPipes) * 32;  // + one issuer and one producer warp per pipeline
constexpr int kRowMaxSub = 8;
constexpr int kRowMaxMma = 24;         // tap records per input row (sub-tiles x horizontal taps)
constexpr int kRowMaxASlots = 8;       // row slots per pipeline
constexpr int kRowMaxRing = 16;        // per pipeline

enum : uint32_t {
  ROWTAP_3ROWS = 1,      // weight tile holds the three vertical taps (dy = +1 | 0 | -1); else one 1x1 tap (dy = 0)
  ROWTAP_RING_FIRST = 2  // first MMA of its accumulator ring in an input row: the one that initialises new output rows
};

struct __align__(16) RowMma {  // one horizontal tap of a sub-tile: nk tcgen05.mma (K = 16 slices) per run of output rows
  uint32_t a_lo;         // A descriptor lower half relative to the slot: (dx * pixel_bytes) >> 4 | LBO field
  uint32_t b_lo;         // weight tile offset inside the launch's weight image >> 4
  uint32_t grp16;        // bytes of one vertical tap's rows of the weight tile >> 4 (aw * pixel_bytes / 16)
  uint32_t flags;        // ROWTAP_*
};

struct RowSub {          // one A sub-tile = (source tensor, channel block) of one input row
  int32_t c;             // first channel (coordinate 0 of the TMA box)
  uint32_t bytes;        // box bytes = 130 * pixel_bytes
  uint32_t a_hi, b_hi;   // descriptor upper halves (SBO = 8 rows, version, swizzle mode)
  uint8_t ring;          // accumulator ring of this term (0 / 1)
  uint8_t aw;            // accumulator columns per output row in that ring (<= 64)
  uint8_t src;           // tensor map 0 / 1
  uint8_t rows3;         // 1: 3x3 term (reads halo rows), 0: 1x1 term (centre rows only)
  uint8_t first_mma;     // index into RowProg::mma
  uint8_t n_mma;         // records of this sub-tile: its horizontal taps (3 or 1)
  uint8_t nk;            // K = 16 slices per tap
  uint8_t off_kib;       // offset of this sub-tile inside a row slot, KiB
};
static_assert(sizeof(RowSub) == 24, "RowSub layout");

struct RowProg {
  RowSub sub[kRowMaxSub];
  RowMma mma[kRowMaxMma];
};

struct RowArgs {
  const uint8_t* wimage;   // this launch's weight image (1 KiB aligned), resident in shared memory
  uint32_t w_bytes;
  int n_sub;
  int W, H, B;             // pixel grid (W a multiple of 128)
  int tiles_x;             // W / 128
  int n_pipes;             // 1 or 2 pipelines per CTA
  int a_slots, a_slot_bytes;   // row slots per pipeline (one slot = every sub-tile of one input row), bytes of one
  uint32_t row_bytes3, row_bytes_all;  // TMA bytes of a halo row (3x3 terms only) / of a centre row
  int ring_slots;          // S: output rows in flight per ring and pipeline (a power of two >= 4)
  int ring_aw[2];          // accumulator columns per output row of ring 0 / 1 (0: no second ring)
  int tmem_cols;           // allocation (power of two >= ring columns of all pipelines)
  int n_acc;               // channels the epilogue produces per pixel (N of the epilogue)
  int store_sbc;           // EPI_STD: channels per TMA-store sub-box (0: per-thread stores)
  long long* timeline;     // debug (DRS_V2_TIMELINE & 1): pipeline 0 of CTA 0 stamps 16 clock values per row, else null
  long long* cta_times;    // debug (DRS_V2_TIMELINE & 8): [2 i] / [2 i + 1] = globaltimer at entry / exit of CTA i
  int* err;
  EpiArgs epi;
};

int launch_conv_row(int epi_kind, const CUtensorMap& map0, const CUtensorMap& map1, const CUtensorMap& map_out,
                    const RowArgs& args, const RowProg& prog, int grid, size_t smem_bytes, cudaStream_t stream);
int conv_row_set_smem_limits();
bool conv_row_supports(int epi_kind, int flags);

}  // namespace drs
