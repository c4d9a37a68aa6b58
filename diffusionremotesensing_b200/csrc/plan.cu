// DrsPlan: activation workspace, TMA descriptors, launch list, time tables, sampler, CUDA-graph replay.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>

#include "engine.cuh"
#include "small_kernels.cuh"

namespace drs {

// ------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 NHWC [B, H, W, C] seen as the 5-D tensor (c, x, py, y, b).
//   plain view:    (C,  W,   1, H,   B)
//   stride-2 view: (2C, W/2, 2, H/2, B) -- c = px * C + channel, so a 2x2 / 3x3 stride-2 tap is a plain box
static int make_map(CUtensorMap* map, const void* base, int B, int H, int W, int C, bool stride2, int ck, int tw,
                    int th, int tb, int tpy = 1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return DRS_E_CUDA;
  }
  const cuuint64_t e = 2;  // bytes per bf16
  cuuint64_t dims[5], strides[4];
  if (!stride2) {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = B;
    strides[0] = C * e;
    strides[1] = static_cast<cuuint64_t>(W) * C * e;
    strides[2] = static_cast<cuuint64_t>(W) * C * e;
    strides[3] = static_cast<cuuint64_t>(H) * W * C * e;
  } else {
    dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = B;
    strides[0] = 2 * C * e;
    strides[1] = static_cast<cuuint64_t>(W) * C * e;
    strides[2] = static_cast<cuuint64_t>(2) * W * C * e;
    strides[3] = static_cast<cuuint64_t>(H) * W * C * e;
  }
  const cuuint32_t box[5] = {static_cast<cuuint32_t>(ck), static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(tpy),
                             static_cast<cuuint32_t>(th), static_cast<cuuint32_t>(tb)};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw =
      (ck == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : (ck == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] stride2=%d ck=%d box=(%d,%d,%d)", static_cast<int>(r),
              B, H, W, C, static_cast<int>(stride2), ck, tw, th, tb);
    return DRS_E_CUDA;
  }
  return DRS_OK;
}

// ------------------------------------------------------------------------------------------------
// plan construction
// ------------------------------------------------------------------------------------------------
static void add_act(DrsPlan* p, size_t& cursor, const std::string& name, int C, int H, int W, bool fp32_map = false) {
  ActTensor t;
  t.name = name;
  t.C = C;
  t.H = H;
  t.W = W;
  t.fp32_map = fp32_map;
  t.bytes = static_cast<size_t>(p->nb) * H * W * (fp32_map ? 4 : C * 2);
  t.bytes = (t.bytes + 1023) & ~static_cast<size_t>(1023);
  t.offset = cursor;
  cursor += t.bytes;
  p->acts[name] = t;
}

static void tile_geometry(int W, int H, int* tw, int* th, int* tb) {
  int w = 16;
  while (w > 1 && w / 2 >= W) w /= 2;  // smallest power of two >= W, capped at 16
  int h = kTileM / w;
  int hh = 1;
  while (hh < H && hh < h) hh *= 2;     // smallest power of two >= H, capped at 128 / w
  *tw = w;
  *th = hh;
  *tb = kTileM / (w * hh);
}

static int bind_launch(DrsPlan* p, int spec_idx, const void* src0, const void* src1, int gridW, int gridH, int srcH[2],
                       int srcW[2], void* out, int OH, int OW, Launch* L) {
  const DrsModel* m = p->m;
  const GemmSpec& g = m->gemms[spec_idx];
  memset(L, 0, sizeof(*L));
  L->spec = spec_idx;
  ConvArgs& a = L->args;
  tile_geometry(gridW, gridH, &a.tw, &a.th, &a.tb);
  a.W = gridW;
  a.H = gridH;
  a.B = p->nb;
  a.tiles_x = (gridW + a.tw - 1) / a.tw;
  a.tiles_y = (gridH + a.th - 1) / a.th;
  const int tiles_b = (p->nb + a.tb - 1) / a.tb;
  L->n_tiles = a.tiles_x * a.tiles_y * tiles_b;
  const void* srcs[2] = {src0, src1};
  CUtensorMap* maps[2] = {&L->map0, &L->map1};
  for (int s = 0; s < g.n_src; ++s)
    DRS_TRY(make_map(maps[s], srcs[s], p->nb, srcH[s], srcW[s], g.src_C[s], g.src_stride2[s] != 0, g.src_ck[s], a.tw,
                     a.th, a.tb));
  if (g.n_src == 1) L->map1 = L->map0;
  a.kblocks = m->d_kblocks.as<KBlock>() + g.kb_dev_off;
  a.wpack = m->d_wblob.as<uint8_t>();
  a.nkb = g.nkb;
  a.a_bytes = g.max_a_bytes;
  a.stage_bytes = (g.max_a_bytes + g.max_b_bytes + 1023) & ~1023;
  a.tmem_cols = g.tmem_cols;
  a.n_sub = g.n_sub;
  a.err = p->d_err;
  // pipeline depth: as many CTAs per SM as TMEM allows (at most 3), each with up to 8 stages
  int ctas = std::min(512 / g.tmem_cols, 3);
  int stages = 2;
  for (; ctas >= 1; --ctas) {
    const int budget = (222 * 1024) / ctas - 8 * 1024;
    stages = std::min({budget / a.stage_bytes, 8, g.nkb});
    if (stages >= 3 || stages >= g.nkb || ctas == 1) break;
  }
  if (g.kblocks.empty()) stages = 1;  // second-generation-only program (fused gate): no first-generation pipeline
  if (stages < 1) {
    set_error("%s: stage of %d bytes does not fit shared memory", g.name.c_str(), a.stage_bytes);
    return DRS_E_INVALID;
  }
  a.stages = stages;
  L->smem = static_cast<size_t>(stages) * a.stage_bytes + 1024;
  EpiArgs& e = a.epi;
  e.out = out;
  e.OH = OH;
  e.OW = OW;
  e.OC = g.OC;
  e.oscale = g.oscale;
  e.n_groups = g.n_groups;
  e.group_n = g.n_sub;
  e.col2 = g.col2;
  e.flags = g.flags;
  e.scale = m->f(g.scale);
  e.bias = m->f(g.bias);
  e.scale2 = m->f(g.scale2);
  e.bias2 = m->f(g.bias2);
  e.te = p->table.as<float>();
  e.trow = p->d_trow;
  e.te_stride = m->te_stride;
  e.te_off = g.te_off;
  e.pre_off = g.pre_off;
  e.wvec = m->f(g.wvec);
  e.bvec = m->f(g.bvec);
  e.nvec = g.nvec;
  e.psi_out = nullptr;
  return DRS_OK;
}

// Second-generation binding (conv_gemm2.cuh): used when the output grid holds at least one 8 x 16 tile.
static int sm_count(int device) {
  static int n = 0;
  if (!n) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  return n > 0 ? n : 148;
}

static int bind_launch_v2(DrsPlan* p, const void* src0, const void* src1, int gridW, int gridH, int srcH[2],
                          int srcW[2], Launch* L) {
  const DrsModel* m = p->m;
  const GemmSpec& g = m->gemms[L->spec];
  const GemmSpec::V2& v = g.v2;
  L->use_v2 = false;
  static const bool disabled = (getenv("DRS_DISABLE_V2") != nullptr);
  if (disabled || !v.usable || gridH < kTile2H || gridW < kTile2W) return DRS_OK;
  Conv2Args& a = L->args2;
  memset(&a, 0, sizeof(a));
  const void* srcs[2] = {src0, src1};
  CUtensorMap* maps[2] = {&L->map0, &L->map1};
  CUtensorMap saved0 = L->map0, saved1 = L->map1;
  for (int s = 0; s < g.n_src; ++s) {
    const int r = make_map(maps[s], srcs[s], p->nb, srcH[s], srcW[s], g.src_C[s], g.src_stride2[s] != 0, g.src_ck[s],
                           v.halo_w[s], v.halo_h[s], 1, v.npy[s]);
    if (r != DRS_OK) {
      L->map0 = saved0;
      L->map1 = saved1;
      return r;
    }
  }
  if (g.n_src == 1) L->map1 = L->map0;
  L->prog2 = &v.prog;
  a.wpack = m->d_wblob.as<uint8_t>();
  a.w_split_off = v.w_split_off;
  a.w_split_bytes = v.w_split_bytes;
  a.nkb = v.nkb;
  a.n_sub_tiles = v.n_sub_tiles;
  a.resident = v.resident ? 1 : 0;
  a.W = gridW;
  a.H = gridH;
  a.B = p->nb;
  a.tiles_x = (gridW + kTile2W - 1) / kTile2W;
  a.tiles_y = (gridH + kTile2H - 1) / kTile2H;
  a.n_tiles = a.tiles_x * a.tiles_y * p->nb;
  a.a_slot_bytes = v.a_slot_bytes;
  a.b_stage_bytes = v.b_stage_bytes;
  a.b_unit = v.b_unit;
  a.acc_cols = v.acc_cols;
  if (2 * v.acc_cols > 512) return DRS_OK;  // the kernel keeps one accumulator per tile of a pair in TMEM
  // two accumulator buffers per tile of the pair when TMEM allows: the epilogue of pair i overlaps the MMAs of i + 1
  a.acc_bufs = (4 * v.acc_cols <= 512) ? 2 : 1;
  int alloc = 32;
  while (alloc < 2 * a.acc_bufs * v.acc_cols) alloc <<= 1;
  a.tmem_cols = alloc;
  a.n_sub = g.n_sub;
  a.nsplit = g.nsplit;
  // transposed convolutions (four column groups, one accumulator buffer per pipeline): both epilogue groups work on
  // every tile
  static const bool no_solo = (getenv("DRS_V2_NO_SOLO") != nullptr);
  a.solo = (!no_solo && g.epi_kind == EPI_STD && a.acc_bufs == 1 && g.n_groups >= 2 && g.n_groups % 2 == 0) ? 1 : 0;
  a.err = p->d_err;
  static const int timeline = getenv("DRS_V2_TIMELINE") ? atoi(getenv("DRS_V2_TIMELINE")) : 0;
  // bit 0: record stamps, bit 1: skip the epilogue body (timing experiments only); DRS_V2_TIMELINE_LAYER restricts
  // the recording to launches whose name contains the given substring
  static const char* const tl_layer = getenv("DRS_V2_TIMELINE_LAYER");
  a.timeline = (!tl_layer || g.name.find(tl_layer) != std::string::npos) ? timeline : 0;
  a.launch_id = static_cast<int>(p->launches.size());  // bound just before being appended
  a.epi = L->args.epi;
  // shared memory: weights (resident image or a ring) + A slots. A pair of tiles consumes slots in the order
  // (sub-tile, tile-of-pair) and every slot is released after its own taps, so two slots per sub-tile in flight plus
  // two of prefetch keep both issuers fed.
  const int spt = a.n_sub_tiles;
  a.b_stages = a.resident ? 1 : std::min(3, v.nkb);
  const int b_bytes = a.resident ? static_cast<int>(v.w_split_bytes) : a.b_stages * v.b_stage_bytes;
  const int want_slots = std::min(kMaxASlots, std::max(4, 2 * spt + 2));
  // Staged TMA-store epilogue (EPI_STD, bf16 NHWC output whose tiles never straddle two images): 4 KiB per
  // epilogue warp, taken when at least four A slots still fit.
  static const bool no_stage = (getenv("DRS_V2_NO_TMA_STORE") != nullptr);
  // Every width is staged (measured at cfg 2: staging everything 0.879 ms/step, >= 64 channels only 0.899).
  const bool can_stage = !no_stage && g.epi_kind == EPI_STD && (g.n_sub % 16) == 0 &&
                         (g.oscale == 1 || (a.epi.OW % 2 == 0 && a.epi.OH % 2 == 0));
  int stage_bytes = 0;
  // Each CTA already runs two MMA issuers and two epilogue groups; a second co-resident CTA is taken when TMEM and
  // shared memory allow it.
  int ctas = std::min(512 / alloc, 2);
  int slots = 0;
  for (; ctas >= 1; --ctas) {
    // (the register-resident transposed epilogue is worth its staging area even with two A slots: 6 KiB cover the
    // static shared memory, the alignment slack and the per-CTA system reserve when the CTA has the SM to itself)
    const bool tr64 = (g.flags & F_TR64) != 0 && ctas == 1;
    const int budget = (227 * 1024) / ctas - (tr64 ? 6 : 14) * 1024 - b_bytes;
    // An even ring gives every slot to exactly one of the two issuers: a consumer that shared a slot with the
    // other issuer would skip every second phase of its full barrier, and a parity wait cannot tell phase k from
    // phase k + 2.
    stage_bytes = 0;
    if (can_stage && (((budget - kStageBytes) / v.a_slot_bytes) & ~1) >= (tr64 ? 2 : 4)) stage_bytes = kStageBytes;
    slots = std::min(want_slots, (budget - stage_bytes) / v.a_slot_bytes) & ~1;
    if (slots >= 4 || ctas == 1) break;
  }
  if (slots < 2) return DRS_OK;  // does not fit: stay on the first-generation kernel
  a.a_slots = slots;
  a.store_sbc = 0;
  if (stage_bytes) {
    // One sub-box per tile (N <= 64, one column group) can use the whole 4 KiB staging area; with several sub-boxes
    // per tile a single buffer would make every sub-box wait for the TMA unit to read the previous one, so the area
    // is split into two 32-channel (or four 16-channel) buffers that rotate.
    a.store_sbc = std::min(64, g.n_sub);
    const int r = make_map(&L->map_out, a.epi.out, p->nb, a.epi.OH, a.epi.OW, a.epi.OC, g.oscale == 2, a.store_sbc,
                           kTile2W, 4, 1, 1);
    if (r != DRS_OK) return r;
  } else {
    L->map_out = L->map0;
  }
  L->smem = static_cast<size_t>(slots) * v.a_slot_bytes + b_bytes + stage_bytes + 1024;
  const int pairs = (a.n_tiles + 1) / 2;
  int grid = std::min(pairs * g.nsplit, sm_count(m->device) * ctas);
  grid -= grid % g.nsplit;
  if (grid < g.nsplit) grid = g.nsplit;
  L->grid2 = grid;
  L->use_v2 = true;
  static const bool verbose = (getenv("DRS_V2_VERBOSE") != nullptr);
  if (verbose)
    fprintf(stderr, "[drs] %s: conv_gemm2, grid %d (%d per SM), %s weights (%d KiB), %d A slots of %d B, staging %d B, "
            "TMEM %d columns, flags %d\n", g.name.c_str(), grid, ctas, a.resident ? "resident" : "streamed",
            b_bytes / 1024, slots, v.a_slot_bytes, stage_bytes, alloc, g.flags);
  return DRS_OK;
}

// CTA-pair binding (conv_gemm2c.cu) on top of a successful second-generation binding. Taken for the launches named in
// DRS_CG2 (comma separated fragments, "all", or "streamed" = every launch whose weights are not resident otherwise).
static bool cg2_wanted(const std::string& name, bool resident_v2, bool resident_cg2, int nkb) {
  // Default (measured at cfg 2, profiles/layers_r1f.json): take the CTA pair where it turns streamed weights into
  // resident ones (up_convs.1: 73 -> 56 us) and for the long streamed K loops, whose half-size weight tiles let one
  // ring stage carry four K-blocks (up_convs.0: 58 -> 52 us; 36 K-blocks: break-even; shorter loops lose a few
  // percent to the cluster hand-shakes). DRS_CG2 = none | all | streamed | <name fragments> overrides.
  static const char* const env = getenv("DRS_CG2");
  if (!env) return !resident_v2 && (resident_cg2 || nkb >= 36);
  const std::string list = env;
  if (list == "none") return false;
  if (list == "all") return true;
  size_t pos = 0;
  while (pos <= list.size()) {
    const size_t e = list.find(',', pos);
    const std::string frag = list.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
    if (frag == "streamed" && !resident_v2) return true;
    if (!frag.empty() && frag != "streamed" && name.find(frag) != std::string::npos) return true;
    if (e == std::string::npos) break;
    pos = e + 1;
  }
  return false;
}

static int bind_launch_cg2(DrsPlan* p, Launch* L) {
  const DrsModel* m = p->m;
  const GemmSpec& g = m->gemms[L->spec];
  const GemmSpec::V2& v = g.v2;
  L->use_cg2 = false;
  if (!L->use_v2 || !conv_gemm2c_supports(g.epi_kind, g.flags)) return DRS_OK;
  const Conv2Args& a2 = L->args2;
  if (a2.n_tiles % 2) return DRS_OK;
  // one weight-tile size per launch, half of it a whole number of 1 KiB swizzle atoms
  const uint32_t tile_bytes = v.prog.kb[0].b_bytes & 0xFFFFFFu;
  if (tile_bytes % 2048) return DRS_OK;
  for (int i = 0; i < v.nkb; ++i)
    if ((v.prog.kb[i].b_bytes & 0xFFFFFFu) != tile_bytes) return DRS_OK;
  if (static_cast<uint64_t>(tile_bytes) * v.nkb != v.w_split_bytes) return DRS_OK;
  const int half = static_cast<int>(tile_bytes / 2);
  if ((half >> 7) > 256) return DRS_OK;  // TMA box rows

  Conv2Args a = a2;
  a.cg2_half_tile_bytes = half;
  a.timeline_buf = a.timeline ? conv_gemm2_timeline_dev() : nullptr;
  a.span_buf = a.timeline ? conv_gemm2_span_dev() : nullptr;
  const int half_image = static_cast<int>(v.w_split_bytes / 2);
  const int budget = 227 * 1024 - 14 * 1024;
  const int spt = a.n_sub_tiles;
  const int want_slots = std::min(kMaxASlots, std::max(4, 2 * spt + 2));
  // resident when this CTA's half image leaves room for four A slots
  a.resident = (half_image + 4 * v.a_slot_bytes <= budget) ? 1 : 0;
  if (!cg2_wanted(g.name, v.resident, a.resident != 0, v.nkb)) return DRS_OK;
  // half-size tiles: twice as many K-blocks per ring stage for the same bytes (fewer full / empty round trips)
  const int unit_c = std::max(1, std::min(8, (32 * 1024) / half));
  a.b_unit = unit_c;
  a.b_stage_bytes = unit_c * half;
  a.b_stages = a.resident ? 1 : std::min(3, (v.nkb + unit_c - 1) / unit_c);
  const int b_bytes = a.resident ? half_image : a.b_stages * a.b_stage_bytes;
  int stage_bytes = a.store_sbc ? kStageBytes : 0;
  if (stage_bytes && (((budget - b_bytes - stage_bytes) / v.a_slot_bytes) & ~1) < 4) {
    stage_bytes = 0;
    a.store_sbc = 0;
  }
  const int slots = std::min(want_slots, (budget - b_bytes - stage_bytes) / v.a_slot_bytes) & ~1;
  if (slots < 2) return DRS_OK;
  a.a_slots = slots;
  L->smem_c = static_cast<size_t>(slots) * v.a_slot_bytes + b_bytes + stage_bytes + 1024;

  // weight offsets: absolute inside the (half) image when resident, else relative to the K-block's ring unit
  L->prog_c = v.prog;
  for (int i = 0; i < v.nkb;) {
    const int cnt = static_cast<int>(v.prog.kb[i].b_bytes >> 24);
    for (int j = 0; j < cnt; ++j) {
      const int first = i + (j / unit_c) * unit_c;
      const uint32_t off = a.resident ? v.prog.kb[i + j].b_off : (v.prog.kb[i + j].b_off - v.prog.kb[first].b_off);
      L->prog_c.kb[i + j].b_lo = (off / 16) | 0x10000u;
    }
    i += cnt;
  }
  // the packed weight blob as rows of 128 bytes; one box = this CTA's half of one tile
  {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return DRS_OK;
    const cuuint64_t rows = m->d_wblob.bytes / 128;
    const cuuint64_t dims[2] = {64, rows};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(half >> 7)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(&L->map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(m->d_wblob.p), dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return DRS_OK;
  }
  // one CTA pair per TPC: pad the request so that no second CTA fits beside it (two co-resident pairs interleave
  // their TMEM pair allocations)
  const size_t min_smem = 116 * 1024;
  if (L->smem_c < min_smem) L->smem_c = min_smem;
  const int max_clusters = conv_gemm2c_max_clusters(g.flags, L->smem_c);
  if (max_clusters < g.nsplit) return DRS_OK;
  const int n_units = a.n_tiles / 2;
  int clusters = std::min(((n_units + 1) / 2) * g.nsplit, max_clusters);
  clusters -= clusters % g.nsplit;
  if (clusters < g.nsplit) return DRS_OK;
  L->grid_c = 2 * clusters;
  L->args_c = a;
  L->use_cg2 = true;
  static const bool verbose = (getenv("DRS_V2_VERBOSE") != nullptr);
  if (verbose)
    fprintf(stderr, "[drs] %s: CTA-pair kernel, %d clusters, %s weights (%d KiB per CTA), %d A slots, smem %zu\n",
            g.name.c_str(), clusters, a.resident ? "resident" : "streamed", b_bytes / 1024, slots, L->smem_c);
  return DRS_OK;
}

// Row-streaming binding (conv_row.cu): full-width stride-1 layers with at most 64 output channels whose pixel grid is
// a multiple of 128 wide and tall enough to give every SM a dozen rows; takes precedence over the second-generation
// kernel. DRS_ROW=0 disables it, DRS_ROW=force drops the size threshold (layer-level tests).
static int bind_launch_row(DrsPlan* p, const void* src0, const void* src1, int gridW, int gridH, int srcH[2],
                           int srcW[2], Launch* L) {
  const DrsModel* m = p->m;
  const GemmSpec& g = m->gemms[L->spec];
  const GemmSpec::Row& r = g.row;
  L->use_row = false;
  static const char* const env = getenv("DRS_ROW");
  const bool disabled = env && env[0] == '0';
  const bool force = env && env[0] == 'f';
  if (disabled || !r.usable || gridW % kRowTile || gridH < 4) return DRS_OK;
  for (int s = 0; s < g.n_src; ++s)
    if (g.src_stride2[s] || srcH[s] != gridH || srcW[s] != gridW) return DRS_OK;
  const int sms = sm_count(m->device);
  const long long tiles = static_cast<long long>(p->nb) * (gridW / kRowTile) * gridH;
  if (!force && tiles < 12LL * sms) return DRS_OK;  // short ranges pay two halo rows each: the tile kernel wins
  RowArgs& a = L->args_r;
  memset(&a, 0, sizeof(a));
  a.wimage = m->d_wblob.as<uint8_t>() + r.w_off;
  a.w_bytes = r.w_bytes;
  a.n_sub = r.n_sub;
  a.W = gridW;
  a.H = gridH;
  a.B = p->nb;
  a.tiles_x = gridW / kRowTile;
  a.a_slot_bytes = r.a_slot_bytes;
  a.row_bytes3 = r.row_bytes3;
  a.row_bytes_all = r.row_bytes_all;
  // two pipelines only when each still gets a dozen rows (every range re-reads two halo rows)
  static const char* const pipes_env = getenv("DRS_ROW_PIPES");
  int n_pipes = pipes_env ? std::max(1, std::min(r.n_pipes, atoi(pipes_env)))
                          : ((tiles >= 24LL * sms) ? r.n_pipes : 1);
  a.ring_slots = r.ring_slots;
  a.ring_aw[0] = r.ring_aw[0];
  a.ring_aw[1] = r.ring_aw[1];
  a.tmem_cols = 512;
  a.n_acc = g.n_sub;
  a.err = p->d_err;
  {
    static const int timeline = getenv("DRS_V2_TIMELINE") ? atoi(getenv("DRS_V2_TIMELINE")) : 0;
    static const char* const tl_layer = getenv("DRS_V2_TIMELINE_LAYER");
    const bool mine = (!tl_layer || g.name.find(tl_layer) != std::string::npos);
    a.timeline = ((timeline & 1) && mine) ? conv_gemm2_timeline_dev() : nullptr;
    a.cta_times = ((timeline & 8) && mine) ? conv_gemm2_timeline_dev() : nullptr;
  }
  a.epi = L->args.epi;
  a.epi.col2 = r.col2;
  a.epi.n_groups = 1;
  const bool stage = (g.epi_kind == EPI_STD);
  const int stage_bytes = stage ? kStageBytes : 0;
  const int w_pad = static_cast<int>((r.w_bytes + 1023u) & ~1023u);
  const int budget = 227 * 1024 - 16 * 1024 - w_pad - stage_bytes;
  int slots = std::min(kRowMaxASlots, budget / r.a_slot_bytes / n_pipes);  // per pipeline
  if (n_pipes == 2 && slots < 3) {
    n_pipes = 1;  // shared memory holds the A ring of one pipeline only
    slots = std::min(kRowMaxASlots, budget / r.a_slot_bytes);
  }
  if (slots < 3) return DRS_OK;
  a.n_pipes = n_pipes;
  a.a_slots = slots;
  const void* srcs[2] = {src0, src1};
  CUtensorMap maps[2];
  for (int s = 0; s < g.n_src; ++s)
    DRS_TRY(make_map(&maps[s], srcs[s], p->nb, srcH[s], srcW[s], g.src_C[s], false, g.src_ck[s], kRowTile + 2, 1, 1));
  CUtensorMap out_map = maps[0];
  a.store_sbc = 0;
  if (stage) {
    a.store_sbc = std::min(64, g.n_sub);
    DRS_TRY(make_map(&out_map, a.epi.out, p->nb, a.epi.OH, a.epi.OW, a.epi.OC, false, a.store_sbc, 32, 1, 1));
  }
  L->map0 = maps[0];
  L->map1 = (g.n_src == 2) ? maps[1] : maps[0];
  L->map_out = out_map;
  L->prog_r = &r.prog;
  L->smem_r = static_cast<size_t>(slots) * n_pipes * r.a_slot_bytes + w_pad + stage_bytes + 1024;
  L->grid_r = static_cast<int>(std::min<long long>(sms, std::max<long long>(1, tiles / (4 * n_pipes))));
  L->use_row = true;
  L->use_v2 = false;
  L->use_cg2 = false;
  static const bool verbose = (getenv("DRS_V2_VERBOSE") != nullptr);
  if (verbose)
    fprintf(stderr, "[drs] %s: row kernel, %d CTAs x %d pipelines, %d A slots of %d B each, weights %u B, ring %d slots, smem %zu\n",
            g.name.c_str(), L->grid_r, n_pipes, slots, r.a_slot_bytes, r.w_bytes, r.ring_slots, L->smem_r);
  return DRS_OK;
}

static int launch_one(const DrsPlan* p, const Launch& L, float* eps, cudaStream_t st) {
  const GemmSpec& g = p->m->gemms[L.spec];
  if (L.use_row) {
    RowArgs a = L.args_r;
    if (g.epi_kind == EPI_OUT && eps) a.epi.out = eps;
    return launch_conv_row(g.epi_kind, L.map0, L.map1, L.map_out, a, *L.prog_r, L.grid_r, L.smem_r, st);
  }
  if (L.use_cg2)
    return launch_conv_gemm2c(L.map0, L.map1, L.map_out, L.map_w, L.args_c, L.prog_c, L.grid_c, L.smem_c, st);
  if (L.use_v2) {
    Conv2Args a = L.args2;
    if (g.epi_kind == EPI_OUT && eps) a.epi.out = eps;
    return launch_conv_gemm2(g.epi_kind, L.map0, L.map1, L.map_out, a, *L.prog2, L.grid2, L.smem, st);
  }
  ConvArgs a = L.args;
  if (g.epi_kind == EPI_OUT && eps) a.epi.out = eps;
  return launch_conv_gemm(g.epi_kind, L.map0, L.map1, a, L.n_tiles, g.nsplit, L.smem, st);
}

static int alloc_small(DrsPlan* p) {
  const size_t n = static_cast<size_t>(p->nb);
  DRS_TRY(p->small.alloc((3 * n + 16) * sizeof(int) + n * sizeof(float)));
  DRS_CUDA(cudaMemset(p->small.p, 0, p->small.bytes));
  int* base = p->small.as<int>();
  p->d_trow = base;
  p->d_uniq = base + n;
  p->d_labels = base + 2 * n;
  p->d_step = base + 3 * n;
  p->d_err = base + 3 * n + 4;
  p->d_tvals = reinterpret_cast<float*>(base + 3 * n + 16);
  return DRS_OK;
}

int plan_create(DrsModel* m, int nb, int nx, int ncond, int S, int mag, DrsPlan** out) {
  if (!m || !out) {
    set_error("drs_plan_create: null argument");
    return DRS_E_INVALID;
  }
  if (nb < 1 || nx < 1 || nb % nx || S < 16 || S % 8 || mag < 1) {
    set_error("drs_plan_create: bad shape nb=%d nx=%d S=%d mag=%d (S must be a multiple of 8, >= 16)", nb, nx, S, mag);
    return DRS_E_INVALID;
  }
  if (m->has_cond && ncond != 1 && ncond != nb) {
    set_error("drs_plan_create: ncond must be 1 or nb");
    return DRS_E_INVALID;
  }
  if (m->has_cond && m->desc.kind == DRS_MODEL_SUPERRES && S % mag) {
    set_error("drs_plan_create: S=%d is not a multiple of the magnification %d", S, mag);
    return DRS_E_INVALID;
  }
  DRS_CUDA(cudaSetDevice(m->device));
  std::unique_ptr<DrsPlan> p(new DrsPlan());
  p->m = m;
  p->nb = nb;
  p->nx = nx;
  p->ncond = m->has_cond ? ncond : 0;
  p->S = S;
  p->mag = (m->desc.kind == DRS_MODEL_SUPERRES) ? mag : 1;
  DRS_TRY(alloc_small(p.get()));

  // activations (bf16 NHWC unless noted)
  size_t cur = 0;
  const int S1 = S, S2 = S / 2, S4 = S / 4, S8 = S / 8;
  add_act(p.get(), cur, "h0", 16, S1, S1);
  add_act(p.get(), cur, "b0.h", 32, S1, S1);
  add_act(p.get(), cur, "b0.out", 32, S1, S1);
  add_act(p.get(), cur, "d0", 32, S2, S2);
  add_act(p.get(), cur, "b1.h", 64, S2, S2);
  add_act(p.get(), cur, "b1.out", 64, S2, S2);
  add_act(p.get(), cur, "d1", 64, S4, S4);
  add_act(p.get(), cur, "b2.h", 128, S4, S4);
  add_act(p.get(), cur, "b2.out", 128, S4, S4);
  add_act(p.get(), cur, "d2", 128, S8, S8);
  add_act(p.get(), cur, "bn.h", 256, S8, S8);
  add_act(p.get(), cur, "bn.out", 256, S8, S8);
  const int upC[4] = {256, 128, 64, 32};
  const int upS[3] = {S8, S4, S2};
  for (int i = 0; i < 3; ++i) {
    const std::string si = std::to_string(i);
    add_act(p.get(), cur, "g" + si, upC[i + 1], upS[i], upS[i]);
    add_act(p.get(), cur, "psi" + si, 1, upS[i], upS[i], true);
    add_act(p.get(), cur, "att" + si, upC[i + 1], 2 * upS[i], 2 * upS[i]);
    add_act(p.get(), cur, "uc" + si, upC[i], upS[i], upS[i]);
    add_act(p.get(), cur, "ut" + si, upC[i], 2 * upS[i], 2 * upS[i]);
    if (i < 2) add_act(p.get(), cur, "x" + si, upC[i + 1], 2 * upS[i], 2 * upS[i]);
  }
  DRS_TRY(p->workspace.alloc(cur));
  DRS_CUDA(cudaMemset(p->workspace.p, 0, cur));

  // condition features: fp32 NHWC [ncond, S, S, 16] + scratch for the encoder (3 NCHW planes sets at LR and SR size)
  if (m->has_cond) {
    DRS_TRY(p->cond_feat.alloc(static_cast<size_t>(ncond) * S * S * 16 * sizeof(float)));
    DRS_TRY(p->cond_tmp.alloc(static_cast<size_t>(ncond) * 4 * S * S * sizeof(float) * 3));
  }
  // a default one-row-per-sample time table so that drs_time_embed works without a sampler
  DRS_TRY(p->table.alloc(static_cast<size_t>(nb) * m->te_stride * sizeof(float)));
  p->table_rows = nb;

  // launch list
  uint8_t* ws = p->workspace.as<uint8_t>();
  const size_t n_layers = m->n_layers > 0 ? static_cast<size_t>(m->n_layers) : m->gemms.size();
  bool skip_next_result = false;
  for (size_t li = 0; li < n_layers; ++li) {
    // Small grids: when the wide form of a launch would occupy less than half of the SMs, its narrow variant (32 output
    // channels per CTA) spreads the K loop and the weight streaming over four or more times as many CTAs.
    size_t gi = li;
    {
      const GemmSpec& w = m->gemms[li];
      const ActTensor& s0w = p->acts.at(w.src_name[0]);
      const int gh = w.src_stride2[0] ? s0w.H / 2 : s0w.H, gw = w.src_stride2[0] ? s0w.W / 2 : s0w.W;
      const long tiles = static_cast<long>(p->nb) * ((gh + kTile2H - 1) / kTile2H) * ((gw + kTile2W - 1) / kTile2W);
      const int narrow_below = 74;  // half of the SMs (thresholds 148 / 300 lose 5 / 13 % at cfg 2)
      const long ctas_wide = ((tiles + 1) / 2) * w.nsplit;
      if (li < m->alt.size() && m->alt[li] >= 0 && gh >= kTile2H && gw >= kTile2W && ctas_wide < narrow_below)
        gi = static_cast<size_t>(m->alt[li]);
    }
    // Attention gate: the fused program (psi + result in one launch, conv_gemm2 only) replaces the pair when the
    // model built one for this block, the grid holds an 8 x 16 tile and neither DRS_NO_GATE_FUSION nor the generic
    // epilogue is requested; the result launch that follows is then skipped.
    {
      static const bool no_fuse = getenv("DRS_NO_GATE_FUSION") || getenv("DRS_V2_GENERIC_EPILOGUE") ||
                                  getenv("DRS_DISABLE_V2");
      if (skip_next_result) {
        skip_next_result = false;
        continue;
      }
      if (!no_fuse && li < m->gate_alt.size() && m->gate_alt[li] >= 0) {
        const GemmSpec& f = m->gemms[m->gate_alt[li]];
        const ActTensor& gs = p->acts.at(f.src_name[0]);
        if (gs.H >= kTile2H && gs.W >= kTile2W) {
          gi = static_cast<size_t>(m->gate_alt[li]);
          skip_next_result = true;
        }
      }
    }
    const GemmSpec& g = m->gemms[gi];
    const ActTensor& s0 = p->acts.at(g.src_name[0]);
    const void* src[2] = {ws + s0.offset, nullptr};
    int sH[2] = {s0.H, 0}, sW[2] = {s0.W, 0};
    if (g.n_src == 2) {
      const ActTensor& s1 = p->acts.at(g.src_name[1]);
      src[1] = ws + s1.offset;
      sH[1] = s1.H;
      sW[1] = s1.W;
    }
    const int gridH = g.src_stride2[0] ? s0.H / 2 : s0.H;
    const int gridW = g.src_stride2[0] ? s0.W / 2 : s0.W;
    if (g.n_src == 2) {
      const int h1 = g.src_stride2[1] ? sH[1] / 2 : sH[1];
      const int w1 = g.src_stride2[1] ? sW[1] / 2 : sW[1];
      if (h1 != gridH || w1 != gridW) {
        set_error("%s: source grids differ (%dx%d vs %dx%d)", g.name.c_str(), gridH, gridW, h1, w1);
        return DRS_E_INVALID;
      }
    }
    void* outp = nullptr;
    int OH = gridH * g.oscale, OW = gridW * g.oscale;
    if (!g.out_name.empty()) {
      const ActTensor& o = p->acts.at(g.out_name);
      outp = ws + o.offset;
      if (o.H != OH || o.W != OW) {
        set_error("%s: output grid mismatch", g.name.c_str());
        return DRS_E_INVALID;
      }
    }
    Launch L;
    DRS_TRY(bind_launch(p.get(), static_cast<int>(gi), src[0], src[1], gridW, gridH, sH, sW, outp, OH, OW, &L));
    if ((g.flags & F_ROWSCALE) && !(g.flags & F_GATE))
      L.args.epi.psi = reinterpret_cast<const float*>(ws + p->acts.at(g.src_name[1]).offset);
    if (g.flags & F_GATE) L.args.epi.psi_out = reinterpret_cast<float*>(ws + p->acts.at(g.psi_name).offset);
    DRS_TRY(bind_launch_v2(p.get(), src[0], src[1], gridW, gridH, sH, sW, &L));
    if ((g.flags & F_GATE) && !L.use_v2) {
      set_error("%s: the fused gate program could not be bound", g.name.c_str());
      return DRS_E_INVALID;
    }
    DRS_TRY(bind_launch_cg2(p.get(), &L));
    if (!(g.flags & F_ROWSCALE)) DRS_TRY(bind_launch_row(p.get(), src[0], src[1], gridW, gridH, sH, sW, &L));
    p->launches.push_back(L);
  }
  *out = p.release();
  return DRS_OK;
}

void plan_destroy(DrsPlan* p) {
  if (!p) return;
  if (p->graph_noise) cudaGraphExecDestroy(p->graph_noise);
  if (p->graph_last) cudaGraphExecDestroy(p->graph_last);
  if (p->capture_stream) cudaStreamDestroy(p->capture_stream);
  if (p->side_stream) {
    cudaStreamDestroy(p->side_stream);
    for (int i = 0; i < 3; ++i) {
      cudaEventDestroy(p->ev_fork[i]);
      cudaEventDestroy(p->ev_join[i]);
    }
  }
  delete p;
}

// ------------------------------------------------------------------------------------------------
// condition encoder (time-invariant)
// ------------------------------------------------------------------------------------------------
int cond_encode(DrsPlan* p, const float* cond, cudaStream_t st) {
  const DrsModel* m = p->m;
  if (!m->has_cond) return DRS_OK;
  if (!cond) {
    set_error("drs_cond_encode: null condition image");
    return DRS_E_INVALID;
  }
  const int Cc = m->desc.cond_channels;
  const int h = p->S / p->mag, w = h;
  const int n = p->ncond;
  const size_t plane = static_cast<size_t>(n) * Cc * p->S * p->S;  // big enough for LR and SR resolution
  float* t0 = p->cond_tmp.as<float>();
  float* t1 = t0 + plane;
  float* t2 = t1 + plane;
  auto conv = [&](const SmallConv& c, const float* in, const float* res, float* out, int H, int W, int relu,
                  int nhwc) {
    return launch_conv3x3_small(in, m->f(c.w), m->f(c.b), res, out, n, c.cin, c.cout, H, W, relu, nhwc, st);
  };
  // RRDB: three ResidualBlocks (conv-relu-conv + x), conv_out, + input   (UNet_model_superres.py:230-260)
  const float* cur = cond;
  float* bufs[2] = {t1, t2};
  for (int i = 0; i < 3; ++i) {
    DRS_CUDA(static_cast<cudaError_t>(conv(m->enc[2 * i], cur, nullptr, t0, h, w, 1, 0)));
    float* o = bufs[i & 1];
    DRS_CUDA(static_cast<cudaError_t>(conv(m->enc[2 * i + 1], t0, cur, o, h, w, 0, 0)));
    cur = o;
  }
  DRS_CUDA(static_cast<cudaError_t>(conv(m->enc[6], cur, cond, t0, h, w, 0, 0)));  // t0 = encoded condition
  const float* enc = t0;
  if (m->desc.kind == DRS_MODEL_SUPERRES && p->mag > 1) {
    DRS_CUDA(static_cast<cudaError_t>(launch_bicubic_up(t0, t1, n, Cc, h, w, p->mag, st)));
    enc = t1;
  }
  DRS_CUDA(static_cast<cudaError_t>(conv(m->cond_conv, enc, nullptr, p->cond_feat.as<float>(), p->S, p->S, 0, 0)));
  return DRS_OK;
}

// ------------------------------------------------------------------------------------------------
// time tables: rows of [relu(mlp_i(enc)) for the 7 MLPs | 9-class border sums for the three UpConvBlocks]
// ------------------------------------------------------------------------------------------------
__global__ void border_class_kernel(const float* __restrict__ taps, float* __restrict__ table, int R, int C,
                                    int te_stride, int pre_off) {
  // taps: [R][9][C] per-tap sums  S_k[oc] = sum_ci W[oc,ci,k] * te[ci];  class (ry, rx): 0 = first, 1 = inner, 2 = last
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(R) * 9 * C;
  if (idx >= total) return;
  const int oc = static_cast<int>(idx % C);
  const int cls = static_cast<int>((idx / C) % 9);
  const int r = static_cast<int>(idx / (9LL * C));
  const int ry = cls / 3, rx = cls % 3;
  const float* t = taps + static_cast<size_t>(r) * 9 * C;
  float acc = 0.f;
  for (int ky = 0; ky < 3; ++ky) {
    if ((ry == 0 && ky == 0) || (ry == 2 && ky == 2)) continue;
    for (int kx = 0; kx < 3; ++kx) {
      if ((rx == 0 && kx == 0) || (rx == 2 && kx == 2)) continue;
      acc += t[(ky * 3 + kx) * C + oc];
    }
  }
  table[static_cast<size_t>(r) * te_stride + pre_off + cls * C + oc] = acc;
}

// rows [r0, r0 + R) of the table from device arrays tvals[R] / labels[R]
static int fill_table_rows(DrsPlan* p, float* table, const float* tvals, const int* labels, int R, cudaStream_t st) {
  const DrsModel* m = p->m;
  const size_t need = static_cast<size_t>(R) * (100 + 256 + 9 * 256) * sizeof(float);
  if (p->scratch.bytes < need) {
    DRS_CUDA(cudaStreamSynchronize(st));
    DRS_TRY(p->scratch.alloc(need));
  }
  float* enc = p->scratch.as<float>();
  float* hid = enc + static_cast<size_t>(R) * 100;
  float* taps = hid + static_cast<size_t>(R) * 256;
  DRS_CUDA(static_cast<cudaError_t>(
      launch_pos_encoding(tvals, labels, m->f(m->label_emb), m->f(m->inv_freq), enc, R, st)));
  for (int i = 0; i < 7; ++i) {
    const TimeMlp& t = m->mlps[i];
    DRS_CUDA(static_cast<cudaError_t>(
        launch_sgemm_nt(enc, 100, m->f(t.w1), 100, m->f(t.b1), hid, t.C, R, t.C, 100, 1, st)));
    DRS_CUDA(static_cast<cudaError_t>(
        launch_sgemm_nt(hid, t.C, m->f(t.w2), t.C, m->f(t.b2), table + t.te_off, m->te_stride, R, t.C, t.C, 2, st)));
    if (t.wtap >= 0) {
      DRS_CUDA(static_cast<cudaError_t>(launch_sgemm_nt(table + t.te_off, m->te_stride, m->f(t.wtap), t.C, nullptr,
                                                        taps, 9 * t.C, R, 9 * t.C, t.C, 0, st)));
      const long long total = static_cast<long long>(R) * 9 * t.C;
      border_class_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(taps, table, R, t.C,
                                                                                      m->te_stride, t.pre_off);
      DRS_CUDA(cudaGetLastError());
    }
  }
  return DRS_OK;
}

static void rebind_table(DrsPlan* p) {
  for (Launch& L : p->launches) {
    L.args.epi.te = p->table.as<float>();
    L.args2.epi.te = p->table.as<float>();
    L.args_c.epi.te = p->table.as<float>();
    L.args_r.epi.te = p->table.as<float>();
  }
}

static void drop_graphs(DrsPlan* p) {
  if (p->graph_noise) cudaGraphExecDestroy(p->graph_noise);
  if (p->graph_last) cudaGraphExecDestroy(p->graph_last);
  p->graph_noise = p->graph_last = nullptr;
}

__global__ void iota_rows_kernel(int* trow, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) trow[i] = i;
}

int time_embed(DrsPlan* p, const float* t_dev, const int* label_dev, cudaStream_t st) {
  if (!t_dev) {
    set_error("drs_time_embed: null timestep array");
    return DRS_E_INVALID;
  }
  if (label_dev && p->m->label_emb < 0) {
    set_error("drs_time_embed: labels given but the model has no label_emb");
    return DRS_E_INVALID;
  }
  if (p->table_rows < p->nb) {
    set_error("drs_time_embed: table too small");
    return DRS_E_STATE;
  }
  DRS_TRY(fill_table_rows(p, p->table.as<float>(), t_dev, label_dev, p->nb, st));
  iota_rows_kernel<<<1, 256, 0, st>>>(p->d_trow, p->nb);
  DRS_CUDA(cudaGetLastError());
  p->prepared = false;  // the table no longer holds the sampler's rows
  return DRS_OK;
}

// ------------------------------------------------------------------------------------------------
// UNet forward
// ------------------------------------------------------------------------------------------------
static int enqueue_forward(DrsPlan* p, const float* x, float* eps, cudaStream_t st, cudaEvent_t* tev = nullptr) {
  const DrsModel* m = p->m;
  const ActTensor& h0 = p->acts.at("h0");
  if (tev) DRS_CUDA(cudaEventRecord(tev[0], st));
  DRS_CUDA(static_cast<cudaError_t>(launch_conv0(x, m->fblob.data() + m->conv0.w, m->fblob.data() + m->conv0.b,
                                                 m->has_cond ? p->cond_feat.as<float>() : nullptr,
                                                 p->workspace.as<uint8_t>() + h0.offset, p->nb, p->nx,
                                                 m->has_cond ? p->ncond : 1, m->desc.x_channels, p->S, st)));
  if (tev) DRS_CUDA(cudaEventRecord(tev[1], st));
  // Decoder stage i: [gating -> psi -> attention result] only meets [UpConvBlock conv -> transposed conv] at
  // up_convs.i, so the gate branch runs on a side stream (forked / joined with events; inside a CUDA-graph capture
  // this becomes two parallel branches of the graph). The gate kernels are small and latency-bound.
  static const bool no_fork = (getenv("DRS_NO_FORK") != nullptr);
  if (!no_fork && !p->side_stream) {
    DRS_CUDA(cudaStreamCreateWithFlags(&p->side_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 3; ++i) {
      DRS_CUDA(cudaEventCreateWithFlags(&p->ev_fork[i], cudaEventDisableTiming));
      DRS_CUDA(cudaEventCreateWithFlags(&p->ev_join[i], cudaEventDisableTiming));
    }
  }
  int stage = -1;
  bool forked = false;
  for (Launch& L : p->launches) {
    const GemmSpec& g = m->gemms[L.spec];
    cudaStream_t use = st;
    if (!no_fork) {
      const bool gate = g.name.rfind("gating_signals.", 0) == 0 || g.name.rfind("attention_blocks.", 0) == 0;
      if (g.name.rfind("gating_signals.", 0) == 0) {
        ++stage;
        DRS_CUDA(cudaEventRecord(p->ev_fork[stage], st));
        DRS_CUDA(cudaStreamWaitEvent(p->side_stream, p->ev_fork[stage], 0));
        forked = true;
      }
      if (gate && forked) use = p->side_stream;
      if (forked && g.name.rfind("up_convs.", 0) == 0) {
        DRS_CUDA(cudaEventRecord(p->ev_join[stage], p->side_stream));
        DRS_CUDA(cudaStreamWaitEvent(st, p->ev_join[stage], 0));
        forked = false;
      }
    }
    const int r = launch_one(p, L, eps, use);
    if (r != 0) {
      set_error("launch of %s failed: %s", g.name.c_str(), cudaGetErrorString(static_cast<cudaError_t>(r)));
      return DRS_E_CUDA;
    }
  }
  if (tev) DRS_CUDA(cudaEventRecord(tev[2], st));
  return DRS_OK;
}

// Live timing of the forward as the sampler runs it (same launch sequence, side-stream gate branch included):
// ms_out[0] = conv0, ms_out[1] = the chain of tensor-core launches (first launch to completion of the last one, no
// events in between), averaged over `iters` forwards.
int plan_time_forward(DrsPlan* p, const float* x, float* eps, int iters, float* ms_out, cudaStream_t st) {
  if (!x || !eps || !ms_out || iters < 1) {
    set_error("drs_plan_time_forward: bad arguments");
    return DRS_E_INVALID;
  }
  cudaEvent_t ev[3];
  for (auto& e : ev) DRS_CUDA(cudaEventCreate(&e));
  double acc0 = 0.0, acc1 = 0.0;
  int rc = DRS_OK;
  for (int it = 0; it < iters && rc == DRS_OK; ++it) {
    rc = enqueue_forward(p, x, eps, st, ev);
    if (rc != DRS_OK) break;
    const cudaError_t se = cudaStreamSynchronize(st);
    if (se != cudaSuccess) {
      rc = cuda_fail(se, "cudaStreamSynchronize(time_forward)");
      break;
    }
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, ev[0], ev[1]);
    cudaEventElapsedTime(&b, ev[1], ev[2]);
    acc0 += a;
    acc1 += b;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  ms_out[0] = static_cast<float>(acc0 / iters);
  ms_out[1] = static_cast<float>(acc1 / iters);
  return rc;
}

int check_pipeline_error(DrsPlan* p, cudaStream_t st) {
  int e = 0;
  DRS_CUDA(cudaMemcpyAsync(&e, p->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  if (e != 0) {
    set_error("tensor-core pipeline timed out (role %d: 1 = TMA producer, 2 = MMA issuer, 3 = epilogue)", e);
    cudaMemsetAsync(p->d_err, 0, sizeof(int), st);
    return DRS_E_PIPELINE;
  }
  return DRS_OK;
}

int unet_forward(DrsPlan* p, const float* x, float* eps, cudaStream_t st) {
  if (!x || !eps) {
    set_error("drs_unet_forward: null buffer");
    return DRS_E_INVALID;
  }
  return enqueue_forward(p, x, eps, st);
}

// ------------------------------------------------------------------------------------------------
// sampler
// ------------------------------------------------------------------------------------------------
int sampler_prepare(DrsPlan* p, int noise_steps, const float* c1, const float* c2, const float* c3,
                    const int* labels_host, float cfg_scale, cudaStream_t st) {
  const DrsModel* m = p->m;
  if (noise_steps < 2 || !c1 || !c2 || !c3) {
    set_error("drs_sampler_prepare: bad arguments");
    return DRS_E_INVALID;
  }
  if (labels_host && m->label_emb < 0) {
    set_error("drs_sampler_prepare: labels given but the model has no label_emb");
    return DRS_E_INVALID;
  }
  // distinct labels -> table columns
  std::vector<int> uniq, idx(p->nb, 0);
  if (labels_host) {
    for (int b = 0; b < p->nb; ++b) {
      const int l = labels_host[b];
      if (l < -1 || l >= m->desc.num_classes) {
        set_error("drs_sampler_prepare: label %d out of range", l);
        return DRS_E_INVALID;
      }
      auto it = std::find(uniq.begin(), uniq.end(), l);
      if (it == uniq.end()) {
        uniq.push_back(l);
        idx[b] = static_cast<int>(uniq.size()) - 1;
      } else {
        idx[b] = static_cast<int>(it - uniq.begin());
      }
    }
  } else {
    uniq.push_back(-1);
  }
  std::vector<float> hc(static_cast<size_t>(noise_steps) * 4);
  for (int i = 0; i < noise_steps; ++i) {
    hc[4 * i + 0] = c1[i];
    hc[4 * i + 1] = c2[i];
    hc[4 * i + 2] = c3[i];
    hc[4 * i + 3] = 0.f;
  }
  // Repeated sample() calls (aggregation sampling runs one per patch batch): when the schedule, the label set and
  // the batch's label indices are the ones the resident tables were built from, nothing is rebuilt, the stream is
  // not synchronised and both captured graphs stay valid. cfg_scale is a kernel argument inside the graphs.
  if (p->prepared && p->noise_steps == noise_steps && p->prep_coef == hc && p->prep_uniq == uniq &&
      p->prep_idx == idx) {
    if (p->cfg_scale != cfg_scale) {
      DRS_CUDA(cudaStreamSynchronize(st));
      drop_graphs(p);
      p->cfg_scale = cfg_scale;
    }
    p->begun = false;
    return DRS_OK;
  }
  p->prepared = false;
  p->n_uniq = static_cast<int>(uniq.size());
  p->noise_steps = noise_steps;
  p->cfg_scale = cfg_scale;
  const int rows = noise_steps * p->n_uniq;
  DRS_CUDA(cudaStreamSynchronize(st));
  drop_graphs(p);
  if (p->table_rows < rows) {
    DRS_TRY(p->table.alloc(static_cast<size_t>(rows) * m->te_stride * sizeof(float)));
    p->table_rows = rows;
    rebind_table(p);
  }
  DevMem tv, lb;
  DRS_TRY(p->coef.upload(hc.data(), hc.size() * sizeof(float)));
  p->d_coef = p->coef.as<float>();
  DRS_CUDA(cudaMemcpy(p->d_uniq, idx.data(), p->nb * sizeof(int), cudaMemcpyHostToDevice));
  // rows are ordered (step, distinct label); filled in chunks to bound the scratch
  const int chunk = 4096;
  DRS_TRY(tv.alloc(static_cast<size_t>(chunk) * sizeof(float)));
  DRS_TRY(lb.alloc(static_cast<size_t>(chunk) * sizeof(int)));
  std::vector<float> htv(chunk);
  std::vector<int> hlb(chunk);
  for (int r0 = 0; r0 < rows; r0 += chunk) {
    const int R = std::min(chunk, rows - r0);
    for (int r = 0; r < R; ++r) {
      htv[r] = static_cast<float>((r0 + r) / p->n_uniq);
      hlb[r] = uniq[(r0 + r) % p->n_uniq];
    }
    DRS_CUDA(cudaMemcpyAsync(tv.p, htv.data(), R * sizeof(float), cudaMemcpyHostToDevice, st));
    DRS_CUDA(cudaMemcpyAsync(lb.p, hlb.data(), R * sizeof(int), cudaMemcpyHostToDevice, st));
    DRS_TRY(fill_table_rows(p, p->table.as<float>() + static_cast<size_t>(r0) * m->te_stride, tv.as<float>(),
                            m->label_emb >= 0 ? lb.as<int>() : nullptr, R, st));
    DRS_CUDA(cudaStreamSynchronize(st));
  }
  p->prep_coef = std::move(hc);
  p->prep_uniq = uniq;
  p->prep_idx = idx;
  p->prepared = true;
  p->begun = false;
  return DRS_OK;
}

int sampler_begin(DrsPlan* p, float* x, float* noise, float* eps, int start_step, cudaStream_t st) {
  if (!p->prepared) {
    set_error("drs_sampler_begin: call drs_sampler_prepare first");
    return DRS_E_STATE;
  }
  if (!x || !eps || start_step < 1 || start_step >= p->noise_steps) {
    set_error("drs_sampler_begin: bad arguments (start_step=%d, noise_steps=%d)", start_step, p->noise_steps);
    return DRS_E_INVALID;
  }
  if (x != p->x || noise != p->noise || eps != p->eps) drop_graphs(p);
  p->x = x;
  p->noise = noise;
  p->eps = eps;
  p->cur_step = start_step;
  DRS_CUDA(static_cast<cudaError_t>(
      launch_set_rows(p->d_trow, p->d_uniq, p->nb, p->d_step, start_step, p->n_uniq, st)));
  p->begun = true;
  return DRS_OK;
}

static int enqueue_step(DrsPlan* p, bool with_noise, cudaStream_t st) {
  const DrsModel* m = p->m;
  DRS_TRY(enqueue_forward(p, p->x, p->eps, st));
  const size_t numel = static_cast<size_t>(p->nx) * m->desc.out_channels * p->S * p->S;
  const int cfg = (p->nb == 2 * p->nx) ? 1 : 0;
  // posterior update + sampler bookkeeping (step index and time-table rows) in one launch
  DRS_CUDA(static_cast<cudaError_t>(launch_ddpm_update(p->x, p->eps, with_noise ? p->noise : nullptr, p->d_coef,
                                                       p->d_step, numel, cfg, p->cfg_scale, p->d_trow, p->nb,
                                                       p->n_uniq, p->d_step + 1, st)));
  return DRS_OK;
}

int sampler_step(DrsPlan* p, int use_graph, cudaStream_t st) {
  if (!p->begun) {
    set_error("drs_sampler_step: call drs_sampler_begin first");
    return DRS_E_STATE;
  }
  if (p->cur_step < 1) {
    set_error("drs_sampler_step: the sampler already reached step 0");
    return DRS_E_STATE;
  }
  if (p->m->desc.x_channels != p->m->desc.out_channels) {
    set_error("drs_sampler_step: x_channels != out_channels");
    return DRS_E_INVALID;
  }
  // the reference injects noise for i > 1 and zeros at i == 1 (train_diffusion_superres.py:243-248)
  const bool with_noise = (p->cur_step > 1) && p->noise != nullptr;
  if (!use_graph) {
    DRS_TRY(enqueue_step(p, with_noise, st));
  } else {
    cudaGraphExec_t& ge = with_noise ? p->graph_noise : p->graph_last;
    if (!ge) {
      // Capture on a private stream: the caller's stream may be the legacy default stream, which cannot be
      // captured. Nothing executes during capture; the instantiated graph is replayed on the caller's stream.
      if (!p->capture_stream) DRS_CUDA(cudaStreamCreateWithFlags(&p->capture_stream, cudaStreamNonBlocking));
      cudaGraph_t graph = nullptr;
      DRS_CUDA(cudaStreamBeginCapture(p->capture_stream, cudaStreamCaptureModeThreadLocal));
      const int r = enqueue_step(p, with_noise, p->capture_stream);
      const cudaError_t ce = cudaStreamEndCapture(p->capture_stream, &graph);
      if (r != DRS_OK) {
        if (graph) cudaGraphDestroy(graph);
        return r;
      }
      DRS_CUDA(ce);
      const cudaError_t ie = cudaGraphInstantiate(&ge, graph, 0);
      cudaGraphDestroy(graph);
      DRS_CUDA(ie);
    }
    DRS_CUDA(cudaGraphLaunch(ge, st));
  }
  p->cur_step -= 1;
  return DRS_OK;
}

// Event-timed durations of the two CUDA-core kernels of a reverse step, each launch on a cold L2 (the caller's flush
// buffer, larger than L2, is read before every timed launch, which leaves the L2 cold and clean): ms_out[0] = conv0, ms_out[1] = posterior update
// with its bookkeeping tail. The sampler's step index is restored afterwards.
int sampler_time_hbm_kernels(DrsPlan* p, void* flush_dev, size_t flush_bytes, int iters, float* ms_out,
                             cudaStream_t st) {
  if (!p->begun || !ms_out || iters < 1 || iters >= p->cur_step) {
    set_error("drs_sampler_time_hbm_kernels: needs a begun sampler with more than `iters` steps left");
    return DRS_E_STATE;
  }
  const DrsModel* m = p->m;
  const ActTensor& h0 = p->acts.at("h0");
  const size_t numel = static_cast<size_t>(p->nx) * m->desc.out_channels * p->S * p->S;
  const int cfg = (p->nb == 2 * p->nx) ? 1 : 0;
  cudaEvent_t ev[4];
  for (auto& e : ev) DRS_CUDA(cudaEventCreate(&e));
  double acc[2] = {0.0, 0.0};
  int rc = DRS_OK;
  for (int it = 0; it < iters && rc == DRS_OK; ++it) {
    if (flush_dev) launch_l2_flush_read(flush_dev, flush_bytes, st);
    cudaEventRecord(ev[0], st);
    int r = launch_conv0(p->x, m->fblob.data() + m->conv0.w, m->fblob.data() + m->conv0.b,
                         m->has_cond ? p->cond_feat.as<float>() : nullptr, p->workspace.as<uint8_t>() + h0.offset,
                         p->nb, p->nx, m->has_cond ? p->ncond : 1, m->desc.x_channels, p->S, st);
    cudaEventRecord(ev[1], st);
    if (r != 0) rc = cuda_fail(static_cast<cudaError_t>(r), "conv0 (timing)");
    if (flush_dev) launch_l2_flush_read(flush_dev, flush_bytes, st);
    cudaEventRecord(ev[2], st);
    r = launch_ddpm_update(p->x, p->eps, p->noise, p->d_coef, p->d_step, numel, cfg, p->cfg_scale, p->d_trow, p->nb,
                           p->n_uniq, p->d_step + 1, st);
    cudaEventRecord(ev[3], st);
    if (r != 0) rc = cuda_fail(static_cast<cudaError_t>(r), "ddpm_update (timing)");
    const cudaError_t se = cudaStreamSynchronize(st);
    if (se != cudaSuccess) rc = cuda_fail(se, "cudaStreamSynchronize(time_hbm_kernels)");
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, ev[0], ev[1]);
    cudaEventElapsedTime(&b, ev[2], ev[3]);
    acc[0] += a;
    acc[1] += b;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  ms_out[0] = static_cast<float>(acc[0] / iters);
  ms_out[1] = static_cast<float>(acc[1] / iters);
  DRS_CUDA(static_cast<cudaError_t>(
      launch_set_rows(p->d_trow, p->d_uniq, p->nb, p->d_step, p->cur_step, p->n_uniq, st)));
  return rc;
}

// drs_debug_conv2d: binds gemms[0] of a single-layer model to caller buffers and runs it once.
int debug_bind_and_run(DrsPlan* p, const void* in, int gridW, int gridH, int srcH, int srcW, void* out, int OH, int OW,
                       cudaStream_t st) {
  int sH[2] = {srcH, 0}, sW[2] = {srcW, 0};
  Launch L;
  DRS_TRY(bind_launch(p, 0, in, nullptr, gridW, gridH, sH, sW, out, OH, OW, &L));
  DRS_TRY(bind_launch_v2(p, in, nullptr, gridW, gridH, sH, sW, &L));
  DRS_TRY(bind_launch_cg2(p, &L));
  DRS_TRY(bind_launch_row(p, in, nullptr, gridW, gridH, sH, sW, &L));
  const int r = launch_one(p, L, nullptr, st);
  if (r != 0) return cuda_fail(static_cast<cudaError_t>(r), "launch_conv_gemm(debug)");
  return DRS_OK;
}

int launches_per_step(const DrsPlan* p) { return static_cast<int>(p->launches.size()) + 2; }

// ------------------------------------------------------------------------------------------------
// per-launch accounting (bench / profiles): algorithmic work and live CUDA-event timing
// ------------------------------------------------------------------------------------------------
int launch_count(const DrsPlan* p) { return static_cast<int>(p->launches.size()) + 1; }

// launch 0 = conv0 (CUDA cores), launch i >= 1 = tensor-core launch i-1
int launch_info(const DrsPlan* p, int i, char* name, int name_cap, double* flops, double* bytes, int* ctas,
                int* smem_bytes) {
  const DrsModel* m = p->m;
  if (i < 0 || i >= launch_count(p)) {
    set_error("launch index %d out of range", i);
    return DRS_E_INVALID;
  }
  const double px = static_cast<double>(p->nb) * p->S * p->S;
  if (i == 0) {
    if (name) snprintf(name, name_cap, "conv0");
    if (flops) *flops = 2.0 * px * 16 * 9 * m->desc.x_channels;
    if (bytes) *bytes = static_cast<double>(p->nx) * m->desc.x_channels * p->S * p->S * 4 + px * 16 * 2 +
                        (m->has_cond ? static_cast<double>(p->ncond) * p->S * p->S * 16 * 4 : 0.0);
    if (ctas) *ctas = static_cast<int>((px + 127) / 128);
    if (smem_bytes) *smem_bytes = 0;
    return DRS_OK;
  }
  const Launch& L = p->launches[i - 1];
  const GemmSpec& g = m->gemms[L.spec];
  const double grid_px = static_cast<double>(p->nb) * L.args.H * L.args.W;
  double macs = 0, wbytes = 0;
  for (const KBlock& kb : g.kblocks) {
    macs += static_cast<double>(kb.n) * kb.ck;
    wbytes += kb.b_bytes;
  }
  if (g.kblocks.empty()) {
    macs = g.macs_per_px;
    wbytes = g.weight_bytes;
  }
  if (name) snprintf(name, name_cap, "%s", g.name.c_str());
  if (flops) {
    *flops = 2.0 * grid_px * macs;
    if (g.epi_kind == EPI_OUT) *flops += 2.0 * grid_px * g.n_sub * g.nvec;  // fused 1x1 output conv
    if (g.epi_kind == EPI_PSI) *flops += 2.0 * grid_px * g.n_sub;
  }
  if (bytes) {
    double b = wbytes;
    for (int s = 0; s < g.n_src; ++s) {
      const ActTensor& t = p->acts.at(g.src_name[s]);
      b += static_cast<double>(p->nb) * t.H * t.W * t.C * 2;
    }
    if ((g.flags & F_ROWSCALE) && !(g.flags & F_GATE)) b += grid_px;  // psi map, fp32 at quarter resolution
    if (g.epi_kind == EPI_OUT)
      b += grid_px * g.nvec * 4;
    else if (g.epi_kind == EPI_PSI)
      b += grid_px * 4;
    else
      b += grid_px * g.oscale * g.oscale * g.OC * 2;
    *bytes = b;
  }
  // negative: persistent grid (row kernel: -10000 - grid)
  if (ctas) *ctas = L.use_row ? -10000 - L.grid_r : (L.use_cg2 ? -L.grid_c : (L.use_v2 ? -L.grid2 : L.n_tiles * g.nsplit));
  if (smem_bytes) *smem_bytes = static_cast<int>(L.use_row ? L.smem_r : (L.use_cg2 ? L.smem_c : L.smem));
  return DRS_OK;
}

// Runs `iters` forwards with a CUDA event between consecutive launches (on `st`, the stream the kernels are
// launched on) and returns the mean duration of every launch in milliseconds. Synchronises the stream.
int plan_profile(DrsPlan* p, const float* x, float* eps, int iters, float* ms_out, cudaStream_t st) {
  if (!x || !eps || !ms_out || iters < 1) {
    set_error("drs_plan_profile: bad arguments");
    return DRS_E_INVALID;
  }
  const DrsModel* m = p->m;
  const int n = launch_count(p);
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) DRS_CUDA(cudaEventCreate(&e));
  std::vector<double> acc(n, 0.0);
  int rc = DRS_OK;
  for (int it = 0; it < iters && rc == DRS_OK; ++it) {
    const ActTensor& h0 = p->acts.at("h0");
    cudaEventRecord(ev[0], st);
    launch_conv0(x, m->fblob.data() + m->conv0.w, m->fblob.data() + m->conv0.b, m->has_cond ? p->cond_feat.as<float>() : nullptr,
                 p->workspace.as<uint8_t>() + h0.offset, p->nb, p->nx, m->has_cond ? p->ncond : 1,
                 m->desc.x_channels, p->S, st);
    cudaEventRecord(ev[1], st);
    for (size_t i = 0; i < p->launches.size(); ++i) {
      Launch& L = p->launches[i];
      const GemmSpec& g = m->gemms[L.spec];
      const int r = launch_one(p, L, eps, st);
      if (r != 0) rc = cuda_fail(static_cast<cudaError_t>(r), g.name.c_str());
      cudaEventRecord(ev[i + 2], st);
    }
    const cudaError_t se = cudaStreamSynchronize(st);
    if (se != cudaSuccess) rc = cuda_fail(se, "cudaStreamSynchronize(profile)");
    for (int i = 0; i < n && rc == DRS_OK; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      acc[i] += ms;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  for (int i = 0; i < n; ++i) ms_out[i] = static_cast<float>(acc[i] / iters);
  return rc;
}

}  // namespace drs
