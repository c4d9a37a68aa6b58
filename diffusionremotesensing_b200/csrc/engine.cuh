// Host-side engine behind the C ABI (include/drs_b200.h): weight packing, K-block programs, plans, sampler.
//
//   DrsModel  shape-independent: parsed state_dict, bf16 swizzled weight tiles, folded eval-BatchNorm
//             scale/shift vectors, one GemmSpec per tensor-core launch of the UNet.
//   DrsPlan   shape-dependent: activation workspace (bf16 NHWC), TMA descriptors, the launch list of one UNet
//             evaluation, time-embedding table, sampler state and the captured CUDA graph of one reverse step.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/drs_b200.h"
#include "conv_gemm.cuh"
#include "conv_gemm2.cuh"
#include "conv_row.cuh"

namespace drs {

// convolution flavours a K-block list can express
enum ConvKind : int { CONV_3x3 = 0, CONV_3x3_S2 = 1, CONV_1x1 = 2, CONV_2x2_S2 = 3, CONV_T3x3_S2 = 4 };

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define DRS_CUDA(expr)                                             \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return ::drs::cuda_fail(_e, #expr);     \
  } while (0)
#define DRS_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != DRS_OK) return _r; \
  } while (0)

struct DevMem {
  void* p = nullptr;
  size_t bytes = 0;
  DevMem() = default;
  DevMem(const DevMem&) = delete;
  DevMem& operator=(const DevMem&) = delete;
  ~DevMem() { release(); }
  int alloc(size_t n);
  int upload(const void* host, size_t n);  // alloc + synchronous copy
  void release();
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// One weight tensor feeding an accumulator column range.
struct WeightRef {
  const float* w = nullptr;  // PyTorch layout: [OC, CinTotal, kh, kw]; ConvTranspose: [CinTotal, OC, 3, 3]
  int oc = 0;                // rows of this weight
  int cin_total = 0;         // Cin of the weight tensor
  int ci_off = 0;            // first input channel used (concatenated inputs)
};

// One convolution term of a launch: `stack` weights share the activation tile and are stacked along MMA N.
struct ConvTerm {
  int src = 0;         // 0 / 1: which source tensor of the launch
  int C = 0;           // channels read from that source
  int kind = CONV_3x3;
  std::vector<WeightRef> stack;
  int col_slot = 0;    // accumulator slot: TMEM column = col_slot * n_sub (ConvTranspose adds the phase)
};

struct GemmSpec {
  std::string name;
  std::string src_name[2];  // activation tensors read (plan names)
  std::string out_name;     // activation tensor written ("" for EPI_OUT: the caller's eps buffer)
  std::string psi_name;     // fused gate: fp32 gate map the epilogue also writes (activation taps)
  double macs_per_px = 0;   // set by builders that do not fill `kblocks` (drs_plan_launch_info)
  double weight_bytes = 0;
  int epi_kind = EPI_STD;
  int flags = 0;
  int OC = 0;        // channels of the output tensor (EPI_STD) / accumulator width (EPI_PSI, EPI_OUT)
  int n_sub = 0;     // channels per grid.y slice
  int nsplit = 1;
  int n_groups = 1;  // 4 for ConvTranspose
  int oscale = 1;
  int col2 = 0;
  int tmem_cols = 32;
  int nkb = 0;
  int src_C[2] = {0, 0};
  int src_ck[2] = {0, 0};
  int src_stride2[2] = {0, 0};
  int n_src = 1;
  int max_b_bytes = 0;
  int max_a_bytes = 0;
  std::vector<KBlock> kblocks;  // [nsplit][nkb]
  // parameter vectors (offsets in floats into the model's fp32 parameter blob, -1 = absent)
  long scale = -1, bias = -1, scale2 = -1, bias2 = -1, wvec = -1, bvec = -1;
  int nvec = 0;
  int te_off = 0, pre_off = 0;
  size_t kb_dev_off = 0;  // offset (in KBlocks) into the model's device K-block array

  // second-generation (persistent halo-tile) program of the same launch, see conv_gemm2.cuh
  struct V2 {
    bool usable = false;
    bool resident = false;            // weights of one split fit in shared memory
    int nkb = 0, n_sub_tiles = 0;
    Conv2Prog prog;                   // K-block / sub-tile program, passed to the kernel by value
    int halo_w[2] = {0, 0}, halo_h[2] = {0, 0}, npy[2] = {1, 1};   // TMA box geometry per source
    int a_slot_bytes = 0, b_stage_bytes = 0;
    int b_unit = 1;                   // streamed weights: K-blocks per ring stage
    int acc_cols = 0;
    uint32_t w_split_off = 0, w_split_bytes = 0;
  } v2;

  // row-streaming program of the same launch (conv_row.cuh): wide low-channel layers, resident weights
  struct Row {
    bool usable = false;
    RowProg prog;
    int n_sub = 0;
    uint32_t w_off = 0, w_bytes = 0;   // weight image inside the model's weight blob
    int a_slot_bytes = 0;              // one row slot: every sub-tile of an input row
    uint32_t row_bytes3 = 0, row_bytes_all = 0;
    int n_pipes = 1;                   // pipelines per CTA
    int ring_slots = 0;                // S per pipeline
    int ring_aw[2] = {0, 0};           // accumulator columns per output row of ring 0 / 1
    int col2 = 0;                      // column offset of the second accumulator seen by the epilogue
  } row;
};

struct TimeMlp {
  int C = 0;
  long w1 = -1, b1 = -1, w2 = -1, b2 = -1;  // fp32 blob offsets
  int te_off = 0;                           // offset of relu(mlp(t)) inside a table row
  long wtap = -1;                           // ups only: [9*C, C] tap-major copy of the 3x3 conv weight
  int pre_off = -1;                         // ups only: offset of the [9][C] border-class block
};

struct SmallConv {
  long w = -1, b = -1;
  int cin = 0, cout = 0;
};

}  // namespace drs

struct DrsModel {
  DrsModelDesc desc{};
  int device = 0;
  std::map<std::string, std::vector<float>> sd;  // host copy of the state_dict
  std::vector<float> fblob;                      // fp32 parameters (host staging)
  std::vector<uint8_t> wblob;                    // bf16 swizzled weight tiles (host staging)
  std::vector<drs::KBlock> kb_all;
  drs::DevMem d_fblob, d_wblob, d_kblocks;
  std::vector<drs::GemmSpec> gemms;              // in execution order
  drs::TimeMlp mlps[7];                          // conv_blocks.0-2, bottle_neck, ups.0-2
  int te_stride = 0;
  long inv_freq = -1, label_emb = -1;
  int n_layers = 0;            // gemms[0 .. n_layers) are the layers, the rest their narrow (32-channel) variants
  std::vector<int> alt;        // alt[i]: index of layer i's narrow variant in gemms, or -1
  std::vector<int> gate_alt;   // gate_alt[i]: for a psi launch i, index of the fused gate (psi + result) in gemms, or -1
  drs::SmallConv conv0, enc[7], cond_conv;       // enc: blocks.{0,1,2}.conv{1,2}, conv_out
  bool has_cond = false;

  const float* f(long off) const { return off < 0 ? nullptr : d_fblob.as<float>() + off; }
};

namespace drs {

struct ActTensor {
  std::string name;
  int C = 0, H = 0, W = 0;
  bool fp32_map = false;  // psi maps: fp32 [B, H, W]
  size_t offset = 0;      // bytes into the workspace
  size_t bytes = 0;
};

struct Launch {
  int spec = -1;  // index into model->gemms, or -1 for the conv0 kernel
  CUtensorMap map0, map1, map_out;  // map_out: output view of the TMA-store epilogue (second-generation kernel)
  ConvArgs args;
  int n_tiles = 0;
  size_t smem = 0;
  bool use_v2 = false;
  Conv2Args args2;
  const Conv2Prog* prog2 = nullptr;  // points into the model's GemmSpec
  int grid2 = 0;
  // CTA-pair kernel (conv_gemm2c.cu): its own argument block, weight-offset program, weight tensor map, grid, smem
  bool use_cg2 = false;
  Conv2Args args_c;
  Conv2Prog prog_c;
  CUtensorMap map_w;
  int grid_c = 0;
  size_t smem_c = 0;
  // row-streaming kernel (conv_row.cu)
  bool use_row = false;
  RowArgs args_r;
  const RowProg* prog_r = nullptr;
  int grid_r = 0;
  size_t smem_r = 0;
};

}  // namespace drs

struct DrsPlan {
  DrsModel* m = nullptr;
  int nb = 0, nx = 0, ncond = 0, S = 0, mag = 1;
  drs::DevMem workspace, cond_feat, cond_tmp, table, scratch, small, coef;
  std::map<std::string, drs::ActTensor> acts;
  std::vector<drs::Launch> launches;
  // device-side bookkeeping (inside `small`): trow[nb], uniq[nb], step, err
  int* d_trow = nullptr;
  int* d_uniq = nullptr;
  int* d_step = nullptr;
  int* d_err = nullptr;
  float* d_coef = nullptr;  // [noise_steps][4]
  float* d_tvals = nullptr; // scratch for time values
  int* d_labels = nullptr;
  int table_rows = 0;
  // sampler state
  int noise_steps = 0, n_uniq = 1, cur_step = 0;
  float cfg_scale = 0.f;
  float* x = nullptr;
  float* noise = nullptr;
  float* eps = nullptr;
  bool prepared = false, begun = false;
  // what the resident time table / coefficient table were built from (drs_sampler_prepare is a no-op on a match)
  std::vector<float> prep_coef;
  std::vector<int> prep_uniq, prep_idx;
  cudaGraphExec_t graph_noise = nullptr, graph_last = nullptr;
  cudaStream_t capture_stream = nullptr;
  cudaStream_t side_stream = nullptr;             // attention-gate branch of the decoder stages
  cudaEvent_t ev_fork[3] = {nullptr, nullptr, nullptr}, ev_join[3] = {nullptr, nullptr, nullptr};
  const float* last_x = nullptr;
  float* last_eps = nullptr;
};
