// Implicit-GEMM convolution on tcgen05 tensor cores: shared host/device declarations.
//
// One launch = one "GEMM program": a list of K-blocks. Every K-block names
//   * an activation tile A: 128 output pixels (tb x th x tw) x ck channels of one NHWC bf16 source, fetched
//     with a 5-D TMA box whose start is shifted by the filter tap (out-of-range rows/cols are zero-filled by
//     the TMA unit, which is exactly the conv zero padding),
//   * a weight tile B: n rows x ck, stored pre-swizzled in HBM so a 1-D bulk copy lands it in UMMA layout,
//   * the TMEM accumulator column it feeds.
// 3x3 / 1x1 / 2x2-stride-2 / 3x3-stride-2 convolutions, the four sub-pixel phases of ConvTranspose2d(3,2,1,1),
// concatenated inputs (split-K over two sources) and fused shortcut convolutions are all just different
// K-block lists for the same kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace drs {

constexpr int kTileM = 128;       // pixels per CTA tile = TMEM lanes
constexpr int kGemmThreads = 192; // warp0: TMA producer, warp1: MMA issuer + TMEM owner, warps2-5: epilogue
constexpr int kMaxKBlocks = 64;
constexpr int kMaxN = 256;

struct __align__(16) KBlock {
  int32_t c;        // coordinate 0 of the TMA box (channel, may include px*C for stride-2 views)
  int16_t dx, dy;   // added to the tile origin in coordinates 1 (x) and 3 (y)
  int16_t py;       // coordinate 2 (row parity for stride-2 views, else 0)
  uint8_t src;      // which tensor map (0 / 1)
  uint8_t ck;       // channels in this K-block: 16 / 32 / 64 (row bytes 32 / 64 / 128 = swizzle span)
  uint16_t n;       // MMA N (rows of the weight tile)
  uint16_t col;     // accumulator column offset in TMEM
  uint16_t init;    // 1: overwrite the accumulator (first K-block of this column group)
  uint16_t pad;
  uint32_t b_off;   // byte offset of the weight tile image inside the packed weight blob
  uint32_t b_bytes; // n * ck * 2
};
static_assert(sizeof(KBlock) == 32, "KBlock must be 32 bytes");

enum EpiKind : int { EPI_STD = 0, EPI_PSI = 1, EPI_OUT = 2 };

enum EpiFlags : int {
  F_RELU = 1,        // max(v, 0) after the affine
  F_DUAL_PRE = 2,    // v += acc2 * scale2 before the ReLU (fused shortcut conv with its own BN scale)
  F_DUAL_POST = 4,   // v += acc2 + bias2 after the ReLU (block-0 skip conv)
  F_TE = 8,          // v += te[trow[b]][te_off + c] after the ReLU (time embedding add)
  F_PRE = 16,        // v += te[trow[b]][pre_off + cls(y,x) * OC + c] before the affine (conv of a constant
                     // time vector under zero padding: 9 border classes)
  F_ROWSCALE = 32,   // acc *= psi[b, y/2, x/2] (attention gate commutes with the 1x1 conv)
  F_UPDATE = 64,     // EPI_OUT only: apply the DDPM posterior update in place instead of writing eps
  F_NOSCALE = 256,   // no per-channel scale (plain conv bias, no BatchNorm): the epilogue adds the bias and skips the
                     // scale vector's shared-memory reads (set by the plan, compile-time epilogues only)
  F_TR64 = 512,      // with F_NOSCALE: transposed convolution (four phase groups), 64 channels per CTA, bias only ->
                     // the register-resident epilogue conv_epilogue_tr64 (set by the model builder)
  F_GATE = 128,      // fused attention gate: the first `nvec` accumulator columns hold W_g g + W_x x; the thread turns
                     // them into psi = sigmoid(w . relu(. + bias) + b) and uses it as the row scale of the groups
};

struct EpiArgs {
  void* out;            // EPI_STD: bf16 NHWC [B, OH, OW, OC]; EPI_PSI: fp32 [B, H, W]; EPI_OUT: fp32 NCHW
  int OH, OW, OC;       // output tensor geometry
  int oscale;           // 1, or 2 for the ConvTranspose scatter (group g -> phase (g>>1, g&1))
  int n_groups;         // accumulator column groups that become separate outputs
  int group_n;          // channels per group handled by this CTA
  int col2;             // column offset of the second accumulator (F_DUAL_*)
  int flags;
  const float* scale;   // [OC] (nullptr -> 1)
  const float* bias;    // [OC]
  const float* scale2;  // [OC] F_DUAL_PRE
  const float* bias2;   // [OC] F_DUAL_POST
  const float* te;      // time table base
  const int* trow;      // [B] row of the time table per sample
  int te_stride;        // floats per row
  int te_off;           // offset of this layer's post-add vector inside a row
  int pre_off;          // offset of this layer's 9-class pre-add block inside a row
  const float* psi;     // F_ROWSCALE: [B, H/2, W/2] fp32
  float* psi_out;       // F_GATE: optional copy of the gate map [B, H, W] fp32 (grid resolution), written by split 0
  const float* wvec;    // EPI_PSI: [N]; EPI_OUT: [nvec, N]
  const float* bvec;    // EPI_PSI: [1]; EPI_OUT: [nvec]
  int nvec;
  // EPI_OUT + F_UPDATE (x <- c1 * (x - c2 * eps) + c3 * z), coefficients looked up by *step
  float* x;             // fp32 NCHW state
  const float* noise;   // fp32 NCHW or nullptr
  const float* coef;    // [steps][4] = c1, c2, c3, 0
  const int* step;
};

struct ConvArgs {
  const KBlock* kblocks;   // [nsplit][nkb]
  const uint8_t* wpack;    // packed weight blob
  int nkb;
  int tw, th, tb;          // tile geometry, tw*th*tb == 128
  int W, H, B;             // pixel grid the tiles walk over
  int tiles_x, tiles_y;
  int stages, stage_bytes, a_bytes;   // pipeline ring: each stage = [A | B], both 1024-aligned
  int tmem_cols;           // power of two >= 32
  int n_sub;               // output channels per blockIdx.y slice
  int* err;                // device error word
  EpiArgs epi;
};

// Launches the kernel; grid = (tiles, nsplit). Returns cudaError_t as int.
int launch_conv_gemm(int epi_kind, const CUtensorMap& map0, const CUtensorMap& map1, const ConvArgs& args,
                     int n_tiles, int nsplit, size_t smem_bytes, cudaStream_t stream);
int conv_gemm_set_smem_limits();

}  // namespace drs
