// DrsModel: state_dict -> bf16 UMMA weight tiles + K-block programs + folded eval-BatchNorm vectors.
// Reference modules restated here (shapes and key names): UNet_model_superres.py:57-379,
// UNet_model_SAR_TO_NDVI.py:263-370, generate_new_imgs/UNet_model_generation.py:226-329.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <set>

#include "engine.cuh"

namespace drs {

// ------------------------------------------------------------------------------------------------
// errors / device memory
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return DRS_E_CUDA;
}

int DevMem::alloc(size_t n) {
  release();
  if (n == 0) n = 16;
  DRS_CUDA(cudaMalloc(&p, n));
  bytes = n;
  return DRS_OK;
}
int DevMem::upload(const void* host, size_t n) {
  DRS_TRY(alloc(n));
  if (n) DRS_CUDA(cudaMemcpy(p, host, n, cudaMemcpyHostToDevice));
  return DRS_OK;
}
void DevMem::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
static inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);                                                       // RNE
  return static_cast<uint16_t>(u >> 16);
}

static int pow2_at_least(int v, int lo) {
  int p = lo;
  while (p < v) p <<= 1;
  return p;
}

struct Tap {
  int ky, kx;       // weight tap
  int dx, dy;       // tile-origin shift (coordinates 1 and 3 of the TMA box)
  int py, px;       // parity selectors of a stride-2 view
  int group;        // ConvTranspose output phase (a << 1 | b), else 0
};

static std::vector<Tap> taps_of(int kind) {
  std::vector<Tap> t;
  switch (kind) {
    case CONV_3x3:
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) t.push_back({ky, kx, kx - 1, ky - 1, 0, 0, 0});
      break;
    case CONV_3x3_S2: {
      // input row 2*oy + ky - 1 seen through the view [H/2][2]: ky=0 -> (oy-1, 1), ky=1 -> (oy, 0), ky=2 -> (oy, 1)
      const int d[3] = {-1, 0, 0}, p[3] = {1, 0, 1};
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) t.push_back({ky, kx, d[kx], d[ky], p[ky], p[kx], 0});
      break;
    }
    case CONV_1x1:
      t.push_back({0, 0, 0, 0, 0, 0, 0});
      break;
    case CONV_2x2_S2:
      for (int ky = 0; ky < 2; ++ky)
        for (int kx = 0; kx < 2; ++kx) t.push_back({ky, kx, 0, 0, ky, kx, 0});
      break;
    case CONV_T3x3_S2: {
      // ConvTranspose2d(k=3, s=2, p=1, op=1): out[2i+a] takes (ky=1, i) for a=0 and (ky=2, i), (ky=0, i+1) for a=1
      struct P { int k, d; };
      const std::vector<P> ph[2] = {{{1, 0}}, {{2, 0}, {0, 1}}};
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
          for (const P& y : ph[a])
            for (const P& x : ph[b]) t.push_back({y.k, x.k, x.d, y.d, 0, 0, (a << 1) | b});
      break;
    }
  }
  return t;
}

// One K-block of a halo-tile program: the taps that read the SAME activation offset and feed CONSECUTIVE accumulator
// column groups, stacked along N (one MMA instead of one per tap). Only the transposed convolution has such taps: of
// its nine, the four at shift (0, 0) feed phases 0..3 (N = 4 n_sub), the two at (0, +1 row) phases 2, 3, the rest
// stand alone - five MMAs per K slice instead of nine, the same accumulation order per output element.
struct KRec {
  int dx, dy, py, px;
  std::vector<Tap> parts;  // ascending, consecutive groups
};

static std::vector<KRec> records_of(int kind, bool stack_phases) {
  std::vector<KRec> out;
  const std::vector<Tap> taps = taps_of(kind);
  if (kind != CONV_T3x3_S2 || !stack_phases) {
    for (const Tap& t : taps) out.push_back({t.dx, t.dy, t.py, t.px, {t}});
    return out;
  }
  const int shifts[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};  // (dx, dy) in the order every phase meets its taps
  for (const auto& sh : shifts) {
    std::vector<Tap> here;
    for (const Tap& t : taps)
      if (t.dx == sh[0] && t.dy == sh[1]) here.push_back(t);
    std::sort(here.begin(), here.end(), [](const Tap& a, const Tap& b) { return a.group < b.group; });
    for (size_t i = 0; i < here.size();) {
      size_t j = i + 1;
      while (j < here.size() && here[j].group == here[j - 1].group + 1) ++j;
      out.push_back({sh[0], sh[1], 0, 0, std::vector<Tap>(here.begin() + i, here.begin() + j)});
      i = j;
    }
  }
  return out;
}

struct Builder {
  DrsModel* m;
  explicit Builder(DrsModel* mm) : m(mm) {}

  const std::vector<float>* get(const std::string& key, size_t numel) {
    auto it = m->sd.find(key);
    if (it == m->sd.end()) {
      set_error("state_dict entry '%s' is missing", key.c_str());
      return nullptr;
    }
    if (it->second.size() != numel) {
      set_error("state_dict entry '%s' has %zu elements, expected %zu", key.c_str(), it->second.size(), numel);
      return nullptr;
    }
    return &it->second;
  }

  long push(const float* v, size_t n) {
    while (m->fblob.size() % 4) m->fblob.push_back(0.f);
    const long off = static_cast<long>(m->fblob.size());
    m->fblob.insert(m->fblob.end(), v, v + n);
    return off;
  }
  long push(const std::vector<float>& v) { return push(v.data(), v.size()); }

  // eval BatchNorm folded behind a conv bias: y = acc * s + (cb * s + beta - mean * s)
  bool bn_fold(const std::string& bn, const std::vector<float>* cb, int C, std::vector<float>& s,
               std::vector<float>& b) {
    const auto* w = get(bn + ".weight", C);
    const auto* be = get(bn + ".bias", C);
    const auto* mu = get(bn + ".running_mean", C);
    const auto* var = get(bn + ".running_var", C);
    if (!w || !be || !mu || !var) return false;
    s.resize(C);
    b.resize(C);
    for (int c = 0; c < C; ++c) {
      const float sc = (*w)[c] / sqrtf((*var)[c] + 1e-5f);
      s[c] = sc;
      b[c] = (cb ? (*cb)[c] : 0.f) * sc + ((*be)[c] - (*mu)[c] * sc);
    }
    return true;
  }

  // Narrow variants (32 output channels per CTA) of the launches whose wide form leaves most SMs idle on small grids:
  // same terms, same epilogue, four (or more) times as many CTAs, each streaming a quarter of the weights. The plan
  // picks the variant per launch (plan.cu); a variant is found through its name.
  std::vector<GemmSpec> narrow;
  std::vector<GemmSpec> fused_gates;  // "<attention block>.gate": replaces that block's psi + result launches

  bool build(GemmSpec& g, const std::vector<ConvTerm>& terms) {
    if (!build_one(g, terms)) return false;
    static const bool no_narrow = (getenv("DRS_NO_NARROW") != nullptr);
    if (!no_narrow && g.epi_kind == EPI_STD && g.n_sub >= 64 && g.OC % 32 == 0 && !(g.flags & F_DUAL_POST)) {
      GemmSpec n = g;
      n.kblocks.clear();
      n.v2 = GemmSpec::V2();
      n.max_a_bytes = n.max_b_bytes = 0;
      n.n_sub = 32;
      if (n.flags & F_DUAL_PRE) n.col2 = n.n_sub;
      if (build_one(n, terms)) narrow.push_back(std::move(n));
      else return false;
    }
    return true;
  }

  // Appends the K-blocks of `terms` for every N-split and packs the weight tiles.
  bool build_one(GemmSpec& g, const std::vector<ConvTerm>& terms) {
    // plain conv bias without BatchNorm (downs, transposed convs, up_convs): the epilogue skips the scale vector
    if (g.epi_kind == EPI_STD && g.scale < 0 && (g.flags & ~(F_NOSCALE | F_TR64)) == 0) {
      g.flags = F_NOSCALE;
      // transposed convolution with 64 channels per CTA: biases in registers (conv_epilogue_tr64, conv_gemm2 only)
      if (g.oscale == 2 && g.n_groups == 4 && g.n_sub == 64) g.flags |= F_TR64;
    }
    g.nsplit = g.OC / g.n_sub;
    if (g.nsplit * g.n_sub != g.OC || g.n_sub % 16) {
      set_error("%s: bad split OC=%d n_sub=%d", g.name.c_str(), g.OC, g.n_sub);
      return false;
    }
    int max_col = 0;
    g.kblocks.clear();
    for (int s = 0; s < g.nsplit; ++s) {
      std::set<int> seen;
      int count = 0;
      for (const ConvTerm& t : terms) {
        const int ck = t.C >= 64 ? 64 : t.C;
        if (ck != 16 && ck != 32 && ck != 64) {
          set_error("%s: unsupported source channel count %d", g.name.c_str(), t.C);
          return false;
        }
        if (t.C % ck) {
          set_error("%s: channels %d not a multiple of %d", g.name.c_str(), t.C, ck);
          return false;
        }
        g.src_C[t.src] = t.C;
        g.src_ck[t.src] = ck;
        g.src_stride2[t.src] = (t.kind == CONV_3x3_S2 || t.kind == CONV_2x2_S2) ? 1 : 0;
        g.n_src = std::max(g.n_src, t.src + 1);
        const int n = g.n_sub * static_cast<int>(t.stack.size());
        if (n > 256) {
          set_error("%s: MMA N=%d exceeds 256", g.name.c_str(), n);
          return false;
        }
        const int row_bytes = ck * 2;
        const uint32_t mask = static_cast<uint32_t>(row_bytes / 16 - 1);
        const bool transposed = (t.kind == CONV_T3x3_S2);
        const int kh = (t.kind == CONV_1x1) ? 1 : (t.kind == CONV_2x2_S2 ? 2 : 3);
        for (const Tap& tp : taps_of(t.kind)) {
          for (int c0 = 0; c0 < t.C; c0 += ck) {
            KBlock kb{};
            kb.c = tp.px * t.C + c0;
            kb.dx = static_cast<int16_t>(tp.dx);
            kb.dy = static_cast<int16_t>(tp.dy);
            kb.py = static_cast<int16_t>(tp.py);
            kb.src = static_cast<uint8_t>(t.src);
            kb.ck = static_cast<uint8_t>(ck);
            kb.n = static_cast<uint16_t>(n);
            const int col = (t.col_slot + tp.group) * g.n_sub;
            kb.col = static_cast<uint16_t>(col);
            kb.init = seen.insert(col).second ? 1 : 0;
            max_col = std::max(max_col, col + n);
            kb.b_bytes = static_cast<uint32_t>(n * row_bytes);
            while (m->wblob.size() % 128) m->wblob.push_back(0);
            kb.b_off = static_cast<uint32_t>(m->wblob.size());
            m->wblob.resize(m->wblob.size() + kb.b_bytes);
            uint8_t* tile = m->wblob.data() + kb.b_off;
            for (int r = 0; r < n; ++r) {
              const WeightRef& wr = t.stack[r / g.n_sub];
              const int oc = s * g.n_sub + (r % g.n_sub);
              for (int k = 0; k < ck; ++k) {
                const int ci = wr.ci_off + c0 + k;
                float v;
                if (transposed)
                  v = wr.w[((static_cast<size_t>(ci) * wr.oc + oc) * 3 + tp.ky) * 3 + tp.kx];
                else
                  v = wr.w[((static_cast<size_t>(oc) * wr.cin_total + ci) * kh + tp.ky) * kh + tp.kx];
                uint32_t o = static_cast<uint32_t>(r * row_bytes + k * 2);
                o ^= ((o >> 7) & mask) << 4;  // Swizzle<log2(row_bytes/16), 4, 3>
                const uint16_t h = f32_to_bf16(v);
                memcpy(tile + o, &h, 2);
              }
            }
            g.max_b_bytes = std::max(g.max_b_bytes, static_cast<int>(kb.b_bytes));
            g.max_a_bytes = std::max(g.max_a_bytes, kTileM * row_bytes);
            g.kblocks.push_back(kb);
            ++count;
          }
        }
      }
      if (s == 0) g.nkb = count;
    }
    if (g.nkb > kMaxKBlocks) {
      set_error("%s: %d K-blocks exceed the limit %d", g.name.c_str(), g.nkb, kMaxKBlocks);
      return false;
    }
    if (max_col > 512) {
      set_error("%s: %d accumulator columns exceed TMEM", g.name.c_str(), max_col);
      return false;
    }
    g.tmem_cols = pow2_at_least(max_col, 32);
    g.kb_dev_off = m->kb_all.size();
    m->kb_all.insert(m->kb_all.end(), g.kblocks.begin(), g.kblocks.end());
    build_v2(g, terms);
    build_row(g, terms);
    return true;
  }

  // Fused attention gate (UNet_model_superres.py:101-107) as ONE second-generation program: per tile of 8 x 16 gate-
  // resolution pixels the K-blocks
  //     W_g . g            (1x1, source 0)                       -> accumulator columns [0, Ch)
  //     W_x[py][px] . x    (2x2 stride 2, source 1 seen through its stride-2 view)  -> columns [0, Ch)
  //     W_r . x[py][px]    (the 1x1 `result` conv on the SAME resident parity plane) -> columns Ch + (2 py + px) n_sub
  // share the four parity planes of the skip tensor, which the unfused pair of launches reads twice; the epilogue
  // (F_GATE) turns the first Ch columns into psi and uses it as the row scale of the four parity groups, so the gate
  // map never travels through memory. Every N split (32 result channels) recomputes the whole gate. Only v2-resident
  // programs are built (Ch <= 64: decoder stages 1 and 2); the caller keeps the separate launches otherwise.
  bool build_gate(GemmSpec& g, int Ch, const std::vector<float>& wg, const std::vector<float>& wx,
                  const std::vector<float>& wr) {
    GemmSpec::V2& v = g.v2;
    v.usable = false;
    // 32-channel blocks: a skip sub-tile (8 x 16 gate pixels x both row parities) is 16 KiB, so six to eight A slots
    // fit beside the resident weights and the store staging (64-channel blocks would leave two)
    const int ck = 32;
    // Only the single-split case (Ch = 32: the full-resolution decoder stage, the most expensive gate) is taken: with
    // Ch = 64 two splits recompute the gate and reload every A tile, and the pair-of-tiles pipeline runs out of A slots
    // (measured in-graph at cfg 2: 40 us fused against 34.5 us for the two launches; Ch = 32: 46 against 63).
    if (Ch != ck) return true;
    const int pix = ck * 2, nblk = Ch / ck;
    g.n_sub = 32;
    g.nsplit = Ch / g.n_sub;
    g.n_src = 2;
    g.src_C[0] = g.src_C[1] = Ch;
    g.src_ck[0] = g.src_ck[1] = ck;
    g.src_stride2[0] = 0;
    g.src_stride2[1] = 1;
    g.nkb = 0;
    g.kblocks.clear();
    g.max_a_bytes = kTileM * pix;
    g.max_b_bytes = Ch * pix;
    const int n_kb = nblk * (1 + 2 * 2 * 2), n_st = nblk * 3;
    if (n_kb > kMaxKBlocks || n_st > kMaxSubTiles) return true;
    const int acc_cols = Ch + 4 * g.n_sub;
    if (2 * acc_cols > 512) return true;
    g.tmem_cols = pow2_at_least(acc_cols, 32);
    v.halo_w[0] = kTile2W; v.halo_h[0] = kTile2H; v.npy[0] = 1;
    v.halo_w[1] = kTile2W; v.halo_h[1] = kTile2H; v.npy[1] = 2;
    v.a_slot_bytes = (pix * kTile2W * 2 * kTile2H + 1023) & ~1023;
    const size_t tile_g = (static_cast<size_t>(Ch) * pix + 1023) & ~static_cast<size_t>(1023);
    const size_t tile_r = (static_cast<size_t>(g.n_sub) * pix + 1023) & ~static_cast<size_t>(1023);
    const size_t w_image = nblk * tile_g + static_cast<size_t>(nblk) * 4 * (tile_g + tile_r);
    if (w_image + 2 * static_cast<size_t>(v.a_slot_bytes) > 200 * 1024) return true;  // resident programs only
    v.resident = true;
    v.b_unit = 1;
    v.b_stage_bytes = static_cast<int>(tile_g);
    v.w_split_bytes = static_cast<uint32_t>(w_image);
    while (m->wblob.size() % 1024) m->wblob.push_back(0);
    v.w_split_off = static_cast<uint32_t>(m->wblob.size());
    m->wblob.resize(m->wblob.size() + w_image * g.nsplit);
    Conv2Prog& P = v.prog;
    memset(&P, 0, sizeof(P));
    const uint32_t mask = static_cast<uint32_t>(pix / 16 - 1);
    auto pack = [&](uint8_t* tile, int rows, auto weight_of) {
      for (int r = 0; r < rows; ++r)
        for (int k = 0; k < ck; ++k) {
          uint32_t o = static_cast<uint32_t>(r * pix + k * 2);
          o ^= ((o >> 7) & mask) << 4;
          const uint16_t h = f32_to_bf16(weight_of(r, k));
          memcpy(tile + o, &h, 2);
        }
    };
    for (int s = 0; s < g.nsplit; ++s) {
      uint8_t* image = m->wblob.data() + v.w_split_off + static_cast<size_t>(s) * w_image;
      size_t off = 0;
      int kb = 0, st = 0;
      bool col_seen[5] = {false, false, false, false, false};  // gate columns, four parity groups
      auto emit = [&](int a_off, uint32_t sbo_bytes, int n, int col, int seen_idx, bool first, bool last, size_t tile_bytes) {
        if (s == 0) {
          KB3& K = P.kb[kb];
          K.a_lo = static_cast<uint32_t>(a_off / 16) | 0x10000u;
          K.a_hi = umma_desc_hi(pix, sbo_bytes);
          K.b_lo = static_cast<uint32_t>(off / 16) | 0x10000u;
          K.b_hi = umma_desc_hi(pix, 8 * pix);
          K.idesc = umma_idesc_host(kTileM, n);
          K.col = static_cast<uint16_t>(col);
          K.nk = static_cast<uint8_t>(ck / 16);
          K.flags = static_cast<uint8_t>((col_seen[seen_idx] ? 0 : KB2_INIT) | (first ? KB2_FIRST : 0) | (last ? KB2_LAST : 0));
          K.b_off = static_cast<uint32_t>(off);
          K.b_bytes = static_cast<uint32_t>(n * pix);
        }
        col_seen[seen_idx] = true;
        off += tile_bytes;
        ++kb;
      };
      // source 0: the gating signal, one K-block per channel block
      for (int c0 = 0; c0 < Ch; c0 += ck) {
        if (s == 0) {
          SubTile& T = P.st[st];
          T.c = c0; T.dx0 = 0; T.dy0 = 0; T.src = 0;
          T.bytes = static_cast<uint32_t>(pix * kTile2W * kTile2H);
        }
        ++st;
        pack(image + off, Ch, [&](int r, int k) { return wg[static_cast<size_t>(r) * Ch + c0 + k]; });
        emit(0, static_cast<uint32_t>(kTile2W * pix), Ch, 0, 0, true, true, tile_g);
      }
      // source 1: the skip tensor through its stride-2 view; sub-tile = (column parity, channel block), both row
      // parities inside; per row parity one gate K-block and one result K-block on the same A rows
      for (int px = 0; px < 2; ++px) {
        for (int c0 = 0; c0 < Ch; c0 += ck) {
          if (s == 0) {
            SubTile& T = P.st[st];
            T.c = px * Ch + c0; T.dx0 = 0; T.dy0 = 0; T.src = 1;
            T.bytes = static_cast<uint32_t>(pix * kTile2W * 2 * kTile2H);
          }
          ++st;
          for (int py = 0; py < 2; ++py) {
            const int a_off = kTile2W * py * pix;
            const uint32_t sbo = static_cast<uint32_t>(kTile2W * 2 * pix);
            pack(image + off, Ch, [&](int r, int k) {
              return wx[((static_cast<size_t>(r) * Ch + c0 + k) * 2 + py) * 2 + px];
            });
            emit(a_off, sbo, Ch, 0, 0, py == 0, false, tile_g);
            pack(image + off, g.n_sub, [&](int r, int k) {
              return wr[static_cast<size_t>(s * g.n_sub + r) * Ch + c0 + k];
            });
            const int grp = (py << 1) | px;
            emit(a_off, sbo, g.n_sub, Ch + grp * g.n_sub, 1 + grp, false, py == 1, tile_r);
          }
        }
      }
      if (s == 0) {
        v.nkb = kb;
        v.n_sub_tiles = st;
      }
    }
    for (int i = 0; i < v.nkb; ++i) {
      if (!(P.kb[i].flags & KB2_FIRST)) continue;
      int cnt = 1;
      while (!(P.kb[i + cnt - 1].flags & KB2_LAST)) ++cnt;
      P.kb[i].b_bytes |= static_cast<uint32_t>(cnt) << 24;
    }
    v.acc_cols = acc_cols;
    v.usable = true;
    // algorithmic work per gate-resolution pixel (drs_plan_launch_info): gate 5 Ch^2 MACs, result 4 Ch^2
    g.macs_per_px = 9.0 * Ch * Ch;
    g.weight_bytes = static_cast<double>(w_image) * g.nsplit;
    return true;
  }

  // Row-streaming program (conv_row.cuh): 3x3 / 1x1 stride-1 terms, one N split, at most 64 channels per output row
  // and ring, weights resident. Weight tiles: one per (source channel block, horizontal tap), rows = the three
  // vertical taps' output channels stacked in the order of the output rows they feed (row above | same row | row
  // below, i.e. ky = 2, 1, 0), K-major, swizzled like the activation rows.
  void build_row(GemmSpec& g, const std::vector<ConvTerm>& terms) {
    GemmSpec::Row& r = g.row;
    r.usable = false;
    if (g.nsplit != 1 || g.n_groups != 1 || g.oscale != 1 || g.n_sub > 64) return;
    if (!conv_row_supports(g.epi_kind, g.flags)) return;
    int ring_aw[2] = {0, 0};
    for (const ConvTerm& t : terms) {
      if (t.kind != CONV_3x3 && t.kind != CONV_1x1) return;
      if (t.col_slot < 0 || t.col_slot > 1) return;
      if (t.kind == CONV_3x3 && t.col_slot != 0) return;  // the second ring only takes 1x1 terms (fused shortcut)
      const int aw = g.n_sub * static_cast<int>(t.stack.size());
      if (ring_aw[t.col_slot] && ring_aw[t.col_slot] != aw) return;
      ring_aw[t.col_slot] = aw;
      if (3 * aw > 256) return;
    }
    if (!ring_aw[0] || terms.empty() || terms[0].col_slot != 0) return;
    if (terms[0].kind != CONV_3x3) return;  // pointless for pure 1x1 layers: nothing to stack
    if (ring_aw[1] && ring_aw[1] != ring_aw[0]) return;  // the epilogue addresses both rings with one slot stride
    // ring slots per pipeline: a power of two (the kernel indexes the ring with a mask), at least 4 (an output row's
    // slot is needed again three input rows later). Two pipelines when TMEM holds two such rings.
    const int cols_per_slot = ring_aw[0] + ring_aw[1];
    if (4 * cols_per_slot > 512) return;
    const int n_pipes = (8 * cols_per_slot <= 512) ? 2 : 1;
    int S = 4;
    while (2 * S <= kRowMaxRing && 2 * S * cols_per_slot * n_pipes <= 512) S *= 2;
    memset(&r.prog, 0, sizeof(r.prog));
    int n_sub = 0, n_mma = 0, max_pix = 0;
    uint32_t row_slot = 0, bytes3 = 0, bytes1 = 0;   // row-slot layout: every sub-tile of one input row, 1 KiB aligned
    size_t image = 0;
    bool ring_seen[2] = {false, false};
    // pass 1: counts
    for (const ConvTerm& t : terms) {
      const int ck = g.src_ck[t.src];
      const int taps = (t.kind == CONV_3x3) ? 3 : 1;
      n_sub += t.C / ck;
      n_mma += taps * (t.C / ck);
    }
    if (n_sub > kRowMaxSub || n_mma > kRowMaxMma) return;
    n_sub = n_mma = 0;
    while (m->wblob.size() % 1024) m->wblob.push_back(0);
    const size_t w_off = m->wblob.size();
    for (const ConvTerm& t : terms) {
      const int ck = g.src_ck[t.src];
      const int pix = ck * 2;
      const bool three = (t.kind == CONV_3x3);
      const int kh = three ? 3 : 1;
      const int aw = ring_aw[t.col_slot];
      const int n_grp = three ? 3 : 1;
      const uint32_t mask = static_cast<uint32_t>(pix / 16 - 1);
      max_pix = std::max(max_pix, pix);
      for (int c0 = 0; c0 < t.C; c0 += ck) {
        RowSub& sub = r.prog.sub[n_sub++];
        sub.c = c0;
        sub.bytes = static_cast<uint32_t>((kRowTile + 2) * pix);
        sub.a_hi = umma_desc_hi(pix, 8 * pix);
        sub.b_hi = umma_desc_hi(pix, 8 * pix);
        sub.ring = static_cast<uint8_t>(t.col_slot);
        sub.aw = static_cast<uint8_t>(aw);
        sub.off_kib = static_cast<uint8_t>(row_slot >> 10);
        row_slot += (sub.bytes + 1023u) & ~1023u;
        (three ? bytes3 : bytes1) += sub.bytes;
        sub.src = static_cast<uint8_t>(t.src);
        sub.rows3 = three ? 1 : 0;
        sub.first_mma = static_cast<uint8_t>(n_mma);
        sub.n_mma = static_cast<uint8_t>(three ? 3 : 1);
        sub.nk = static_cast<uint8_t>(ck / 16);
        for (int dx = 0; dx < (three ? 3 : 1); ++dx) {
          // the box starts one pixel left of x0: horizontal tap dx reads from pixel dx on (a 1x1 term from pixel 1)
          const uint32_t a_off16 = static_cast<uint32_t>(((three ? dx : 1) * pix) >> 4);
          {
            RowMma& mm = r.prog.mma[n_mma++];
            mm.a_lo = 0x10000u | a_off16;
            mm.b_lo = static_cast<uint32_t>(image >> 4);
            mm.grp16 = static_cast<uint32_t>((aw * pix) >> 4);
            mm.flags = (three ? ROWTAP_3ROWS : 0u) | (ring_seen[t.col_slot] ? 0u : ROWTAP_RING_FIRST);
          }
          ring_seen[t.col_slot] = true;
          const size_t tile_bytes = (static_cast<size_t>(n_grp) * aw * pix + 1023) & ~static_cast<size_t>(1023);
          m->wblob.resize(w_off + image + tile_bytes);
          uint8_t* tile = m->wblob.data() + w_off + image;
          for (int gi = 0; gi < n_grp; ++gi) {
            const int ky = three ? 2 - gi : 0, kx = three ? dx : 0;
            for (int rr = 0; rr < aw; ++rr) {
              const WeightRef& wr = t.stack[rr / g.n_sub];
              const int oc = rr % g.n_sub;
              const int row = gi * aw + rr;
              for (int k = 0; k < ck; ++k) {
                const int ci = wr.ci_off + c0 + k;
                const float val = wr.w[((static_cast<size_t>(oc) * wr.cin_total + ci) * kh + ky) * kh + kx];
                uint32_t o = static_cast<uint32_t>(row * pix + k * 2);
                o ^= ((o >> 7) & mask) << 4;
                const uint16_t h = f32_to_bf16(val);
                memcpy(tile + o, &h, 2);
              }
            }
          }
          image += tile_bytes;
        }
      }
    }
    r.n_sub = n_sub;
    r.w_off = static_cast<uint32_t>(w_off);
    r.w_bytes = static_cast<uint32_t>(image);
    r.a_slot_bytes = static_cast<int>(row_slot);
    r.row_bytes3 = bytes3;
    r.row_bytes_all = bytes3 + bytes1;
    r.ring_slots = S;
    r.n_pipes = n_pipes;
    r.ring_aw[0] = ring_aw[0];
    r.ring_aw[1] = ring_aw[1];
    r.col2 = ring_aw[1] ? S * ring_aw[0] : g.col2;
    // resident weights + at least three row slots (one pipeline; the plan takes two when they fit) + the store
    // staging must fit
    const size_t need = image + static_cast<size_t>(3) * r.a_slot_bytes +
                        (g.epi_kind == EPI_STD ? kStageBytes : 0);
    if (need > 200 * 1024) {
      m->wblob.resize(w_off);
      return;
    }
    r.usable = true;
  }

  // Halo-tile program of the same launch (conv_gemm2.cuh). K-blocks are ordered sub-tile major: all taps that read
  // one (source, channel block[, column parity]) halo tile are consecutive, so the tile is loaded once.
  void build_v2(GemmSpec& g, const std::vector<ConvTerm>& terms) {
    GemmSpec::V2& v = g.v2;
    v.usable = false;
    bool src_seen[2] = {false, false};
    for (const ConvTerm& t : terms) {
      if (src_seen[t.src]) return;  // one convolution per source tensor only
      src_seen[t.src] = true;
    }
    // halo geometry per kind
    struct Geo { int dx_min, dy_min, hw, hh, npy; };
    auto geo_of = [](int kind) -> Geo {
      switch (kind) {
        case CONV_3x3: return {-1, -1, kTile2W + 2, kTile2H + 2, 1};
        case CONV_3x3_S2: return {-1, -1, kTile2W + 1, kTile2H + 1, 2};
        case CONV_1x1: return {0, 0, kTile2W, kTile2H, 1};
        case CONV_2x2_S2: return {0, 0, kTile2W, kTile2H, 2};
        default: return {0, 0, kTile2W + 1, kTile2H + 1, 1};  // CONV_T3x3_S2: taps at (0 / +1)
      }
    };
    Conv2Prog& P = v.prog;
    memset(&P, 0, sizeof(P));
    // phase stacking of the transposed convolution needs its four groups in one MMA (N = 4 n_sub <= 256)
    static const bool no_stack = (getenv("DRS_V2_NO_PHASE_STACK") != nullptr);
    bool stack_phases = !no_stack && g.n_groups == 4 && 4 * g.n_sub <= 256 && terms.size() == 1 &&
                        terms[0].stack.size() == 1 && terms[0].kind == CONV_T3x3_S2;
    if (stack_phases) {
      // resident weights only: streamed, a stacked tile fills a ring stage on its own - one full / empty handshake per
      // record instead of per four (ups.0.transform at cfg 2: 37.8 -> 38.9 us)
      const ConvTerm& t = terms[0];
      const int pix = g.src_ck[t.src] * 2;
      const size_t image = static_cast<size_t>(9) * g.n_sub * pix * (t.C / g.src_ck[t.src]);
      const size_t slot = (static_cast<size_t>(pix) * (kTile2W + 1) * (kTile2H + 1) + 1023) & ~static_cast<size_t>(1023);
      if (image + 2 * slot > 200 * 1024) stack_phases = false;
    }
    // pass 1: sizes and per-source constants (identical for every split)
    size_t w_image = 0;
    int max_a = 0, max_b = 0, n_kb = 0, n_st = 0;
    for (const ConvTerm& t : terms) {
      const int ck = g.src_ck[t.src];
      const int pix = ck * 2;
      const Geo G = geo_of(t.kind);
      v.halo_w[t.src] = G.hw;
      v.halo_h[t.src] = G.hh;
      v.npy[t.src] = G.npy;
      max_a = std::max(max_a, pix * G.hw * G.hh * G.npy);
      const int n1 = g.n_sub * static_cast<int>(t.stack.size());
      const int n_px = (t.kind == CONV_3x3_S2 || t.kind == CONV_2x2_S2) ? 2 : 1;
      for (const KRec& r : records_of(t.kind, stack_phases)) {
        const int n = n1 * static_cast<int>(r.parts.size());
        const size_t tile_bytes = (static_cast<size_t>(n) * pix + 1023) & ~static_cast<size_t>(1023);
        w_image += tile_bytes * (t.C / ck);
        n_kb += t.C / ck;
        max_b = std::max(max_b, n * pix);
      }
      n_st += n_px * (t.C / ck);
    }
    if (n_kb > kMaxKBlocks || n_st > kMaxSubTiles) return;
    v.a_slot_bytes = (max_a + 1023) & ~1023;
    // resident when the split's image plus two A slots leaves a CTA within ~200 KB
    v.resident = (w_image + 2 * static_cast<size_t>(v.a_slot_bytes) <= 200 * 1024);
    // Streamed weights travel in units of up to b_unit consecutive K-blocks of one sub-tile (<= 32 KiB): one bulk
    // copy, one full / empty barrier round trip and one tcgen05.commit per unit instead of per K-block.
    const int tile_pad = (max_b + 1023) & ~1023;
    v.b_unit = std::max(1, std::min(4, (32 * 1024) / tile_pad));
    v.b_stage_bytes = v.resident ? tile_pad : v.b_unit * tile_pad;
    v.w_split_bytes = static_cast<uint32_t>(w_image);
    while (m->wblob.size() % 1024) m->wblob.push_back(0);
    v.w_split_off = static_cast<uint32_t>(m->wblob.size());
    m->wblob.resize(m->wblob.size() + w_image * g.nsplit);

    int max_col = 0;
    for (int s = 0; s < g.nsplit; ++s) {
      std::set<int> seen;
      int count = 0, st_count = 0;
      size_t image_off = 0;  // offset inside this split's image (same for every split)
      for (const ConvTerm& t : terms) {
        const int ck = g.src_ck[t.src];
        const int pix = ck * 2;
        const Geo G = geo_of(t.kind);
        const uint32_t mask = static_cast<uint32_t>(pix / 16 - 1);
        const bool transposed = (t.kind == CONV_T3x3_S2);
        const int kh = (t.kind == CONV_1x1) ? 1 : (t.kind == CONV_2x2_S2 ? 2 : 3);
        const std::vector<KRec> recs = records_of(t.kind, stack_phases);
        const int n1 = g.n_sub * static_cast<int>(t.stack.size());
        const int n_px = (t.kind == CONV_3x3_S2 || t.kind == CONV_2x2_S2) ? 2 : 1;
        for (int px = 0; px < n_px; ++px) {
          for (int c0 = 0; c0 < t.C; c0 += ck) {
            std::vector<const KRec*> mine;
            for (const KRec& r : recs)
              if (r.px == px) mine.push_back(&r);
            if (mine.empty()) continue;
            if (s == 0) {
              SubTile& st = P.st[st_count++];
              st.c = px * t.C + c0;
              st.dx0 = static_cast<int16_t>(G.dx_min);
              st.dy0 = static_cast<int16_t>(G.dy_min);
              st.bytes = static_cast<uint32_t>(pix * G.hw * G.hh * G.npy);
              st.src = static_cast<uint8_t>(t.src);
            }
            for (size_t ti = 0; ti < mine.size(); ++ti) {
              const KRec& tp = *mine[ti];
              const int n = n1 * static_cast<int>(tp.parts.size());
              const int a_off = ((tp.dx - G.dx_min) + G.hw * (tp.py + G.npy * (tp.dy - G.dy_min))) * pix;
              const int col = (t.col_slot + tp.parts[0].group) * g.n_sub;
              uint8_t flags = 0;
              // (a stacked record meets all of its groups for the first time together: the shift (0, 0) one)
              bool first = false;
              for (const Tap& part : tp.parts) first = seen.insert((t.col_slot + part.group) * g.n_sub).second || first;
              if (first) flags |= KB2_INIT;
              if (ti == 0) flags |= KB2_FIRST;
              if (ti + 1 == mine.size()) flags |= KB2_LAST;
              max_col = std::max(max_col, col + n);
              if (s == 0) {
                KB3& kb = P.kb[count];
                kb.a_lo = static_cast<uint32_t>(a_off / 16) | 0x10000u;
                kb.a_hi = umma_desc_hi(pix, G.hw * G.npy * pix);
                kb.b_lo = static_cast<uint32_t>(image_off / 16) | 0x10000u;
                kb.b_hi = umma_desc_hi(pix, 8 * pix);
                kb.idesc = umma_idesc_host(kTileM, n);
                kb.col = static_cast<uint16_t>(col);
                kb.nk = static_cast<uint8_t>(ck / 16);
                kb.flags = flags;
                kb.b_off = static_cast<uint32_t>(image_off);
                kb.b_bytes = static_cast<uint32_t>(n * pix);
              }
              uint8_t* tile = m->wblob.data() + v.w_split_off + static_cast<size_t>(s) * w_image + image_off;
              image_off += (static_cast<size_t>(n) * pix + 1023) & ~static_cast<size_t>(1023);
              for (int r = 0; r < n; ++r) {
                const Tap& part = tp.parts[r / n1];
                const WeightRef& wr = t.stack[(r % n1) / g.n_sub];
                const int oc = s * g.n_sub + (r % g.n_sub);
                for (int k = 0; k < ck; ++k) {
                  const int ci = wr.ci_off + c0 + k;
                  float val;
                  if (transposed)
                    val = wr.w[((static_cast<size_t>(ci) * wr.oc + oc) * 3 + part.ky) * 3 + part.kx];
                  else
                    val = wr.w[((static_cast<size_t>(oc) * wr.cin_total + ci) * kh + part.ky) * kh + part.kx];
                  uint32_t o = static_cast<uint32_t>(r * pix + k * 2);
                  o ^= ((o >> 7) & mask) << 4;
                  const uint16_t h = f32_to_bf16(val);
                  memcpy(tile + o, &h, 2);
                }
              }
              ++count;
            }
          }
        }
      }
      if (s == 0) {
        v.nkb = count;
        v.n_sub_tiles = st_count;
      }
    }
    if (max_col > 512) return;
    // number of K-blocks of each sub-tile in the top byte of b_bytes of its FIRST record: the issuing lane's loop
    // is counted, with no data-dependent exit
    for (int i = 0; i < v.nkb; ++i) {
      if (!(P.kb[i].flags & KB2_FIRST)) continue;
      int cnt = 1;
      while (!(P.kb[i + cnt - 1].flags & KB2_LAST)) ++cnt;
      P.kb[i].b_bytes |= static_cast<uint32_t>(cnt) << 24;
    }
    if (!v.resident) {
      // streamed mode: b_lo = offset of the tile inside its unit (the issuing lane adds the ring stage)
      for (int i = 0; i < v.nkb;) {
        const int cnt = static_cast<int>(P.kb[i].b_bytes >> 24);
        for (int j = 0; j < cnt; ++j) {
          const int first = i + (j / v.b_unit) * v.b_unit;
          P.kb[i + j].b_lo = static_cast<uint32_t>((P.kb[i + j].b_off - P.kb[first].b_off) / 16) | 0x10000u;
        }
        i += cnt;
      }
    }
    v.acc_cols = max_col;
    v.usable = true;
  }
};

static WeightRef wref(const std::vector<float>* w, int oc, int cin_total, int ci_off = 0) {
  WeightRef r;
  r.w = w->data();
  r.oc = oc;
  r.cin_total = cin_total;
  r.ci_off = ci_off;
  return r;
}

static ConvTerm term(int src, int C, int kind, std::vector<WeightRef> stack, int slot = 0) {
  ConvTerm t;
  t.src = src;
  t.C = C;
  t.kind = kind;
  t.stack = std::move(stack);
  t.col_slot = slot;
  return t;
}

static int n_sub_for(int OC) { return OC > 128 ? 128 : OC; }

// ResConvBlock (UNet_model_superres.py:110-172) as two launches.
static const char* skip_conv_name(int kind) {
  // ResConvBlock's skip conv: UNet_model_superres.py:129, UNet_model_SAR_TO_NDVI.py:126, UNet_model_generation.py:127
  return kind == DRS_MODEL_SUPERRES ? ".conv_upsampled_lr_img" : (kind == DRS_MODEL_SAR_TO_NDVI ? ".conv_SAR_img" : ".conv_skip");
}

static bool build_res_block(Builder& B, const std::string& p, int cin, int cout, bool has_skip, const std::string& in,
                            const std::string& mid, const std::string& out, int te_off) {
  DrsModel* m = B.m;
  // launch 1: h = relu(bn1(conv1(x))) [+ conv_skip(x)] + relu(time_mlp(t))
  {
    GemmSpec g;
    g.name = p + ".conv1";
    g.src_name[0] = in;
    g.out_name = mid;
    g.OC = cout;
    g.n_sub = has_skip ? std::min(cout, 128) : n_sub_for(cout);
    g.flags = F_RELU | F_TE;
    g.te_off = te_off;
    const auto* w1 = B.get(p + ".conv1.0.weight", static_cast<size_t>(cout) * cin * 9);
    const auto* b1 = B.get(p + ".conv1.0.bias", cout);
    if (!w1 || !b1) return false;
    std::vector<float> s, b;
    if (!B.bn_fold(p + ".batch_norm1", b1, cout, s, b)) return false;
    g.scale = B.push(s);
    g.bias = B.push(b);
    std::vector<WeightRef> stack{wref(w1, cout, cin)};
    if (has_skip) {
      const auto* ws = B.get(p + skip_conv_name(m->desc.kind) + ".weight", static_cast<size_t>(cout) * cin * 9);
      const auto* bs = B.get(p + skip_conv_name(m->desc.kind) + ".bias", cout);
      if (!ws || !bs) return false;
      stack.push_back(wref(ws, cout, cin));
      g.flags |= F_DUAL_POST;
      g.bias2 = B.push(*bs);
      g.col2 = g.n_sub;
    }
    if (!B.build(g, {term(0, cin, CONV_3x3, stack)})) return false;
    m->gemms.push_back(std::move(g));
  }
  // launch 2: out = relu(bn2(conv2(h)) + bn3(shortcut1x1(x)))
  {
    GemmSpec g;
    g.name = p + ".conv2";
    g.src_name[0] = mid;
    g.src_name[1] = in;
    g.out_name = out;
    g.OC = cout;
    g.n_sub = n_sub_for(cout);
    g.flags = F_RELU | F_DUAL_PRE;
    g.col2 = g.n_sub;
    const auto* w2 = B.get(p + ".conv2.0.weight", static_cast<size_t>(cout) * cout * 9);
    const auto* b2 = B.get(p + ".conv2.0.bias", cout);
    const auto* w3 = B.get(p + ".shortcut_conv.0.weight", static_cast<size_t>(cout) * cin);
    const auto* b3 = B.get(p + ".shortcut_conv.0.bias", cout);
    if (!w2 || !b2 || !w3 || !b3) return false;
    std::vector<float> s2, bb2, s3, bb3;
    if (!B.bn_fold(p + ".batch_norm2", b2, cout, s2, bb2)) return false;
    if (!B.bn_fold(p + ".shortcut_batch_norm", b3, cout, s3, bb3)) return false;
    for (int c = 0; c < cout; ++c) bb2[c] += bb3[c];
    g.scale = B.push(s2);
    g.bias = B.push(bb2);
    g.scale2 = B.push(s3);
    if (!B.build(g, {term(0, cout, CONV_3x3, {wref(w2, cout, cout)}, 0),
                     term(1, cin, CONV_1x1, {wref(w3, cout, cin)}, 1)}))
      return false;
    m->gemms.push_back(std::move(g));
  }
  return true;
}

static bool build_time_mlp(Builder& B, const std::string& p, int C, TimeMlp& t) {
  const auto* w1 = B.get(p + ".time_mlp.0.weight", static_cast<size_t>(C) * 100);
  const auto* b1 = B.get(p + ".time_mlp.0.bias", C);
  const auto* w2 = B.get(p + ".time_mlp.2.weight", static_cast<size_t>(C) * C);
  const auto* b2 = B.get(p + ".time_mlp.2.bias", C);
  if (!w1 || !b1 || !w2 || !b2) return false;
  t.C = C;
  t.w1 = B.push(*w1);
  t.b1 = B.push(*b1);
  t.w2 = B.push(*w2);
  t.b2 = B.push(*b2);
  return true;
}

static bool build_small(Builder& B, const std::string& p, int cout, int cin, SmallConv& c) {
  const auto* w = B.get(p + ".weight", static_cast<size_t>(cout) * cin * 9);
  const auto* b = B.get(p + ".bias", cout);
  if (!w || !b) return false;
  c.w = B.push(*w);
  c.b = B.push(*b);
  c.cin = cin;
  c.cout = cout;
  return true;
}

// Appends the narrow variants after the layers and links each layer to its variant (by name).
static void append_narrow(DrsModel* m, Builder& B) {
  m->n_layers = static_cast<int>(m->gemms.size());
  m->alt.assign(m->gemms.size(), -1);
  for (GemmSpec& n : B.narrow) {
    for (int i = 0; i < m->n_layers; ++i)
      if (m->gemms[i].name == n.name) m->alt[i] = static_cast<int>(m->gemms.size());
    m->gemms.push_back(std::move(n));
  }
  B.narrow.clear();
  // fused gates: linked from their psi launch
  m->gate_alt.assign(m->gemms.size(), -1);
  for (GemmSpec& f : B.fused_gates) {
    const std::string psi_name = f.name.substr(0, f.name.size() - 4) + "psi";
    for (int i = 0; i < m->n_layers; ++i)
      if (m->gemms[i].name == psi_name) m->gate_alt[i] = static_cast<int>(m->gemms.size());
    m->gemms.push_back(std::move(f));
  }
  B.fused_gates.clear();
}

static bool build_model(DrsModel* m) {
  Builder B(m);
  const DrsModelDesc& d = m->desc;
  const int down[5] = {16, 32, 64, 128, 256};

  // ---- time-table row layout -------------------------------------------------------------------
  int off = 0;
  const int mlp_C[7] = {32, 64, 128, 256, 256, 128, 64};
  for (int i = 0; i < 7; ++i) {
    m->mlps[i].te_off = off;
    off += mlp_C[i];
  }
  for (int i = 4; i < 7; ++i) {
    m->mlps[i].pre_off = off;
    off += 9 * mlp_C[i];
  }
  m->te_stride = off;

  // ---- CUDA-core pieces ------------------------------------------------------------------------
  if (!build_small(B, "conv0", 16, d.x_channels, m->conv0)) return false;
  m->has_cond = (d.kind != DRS_MODEL_GENERATION);
  if (m->has_cond) {
    const std::string enc = (d.kind == DRS_MODEL_SUPERRES) ? "LR_encoder" : "SAR_encoder";
    const std::string cc = (d.kind == DRS_MODEL_SUPERRES) ? "conv_upsampled_lr_img" : "conv_SAR_img";
    const int Cc = d.cond_channels;
    if (Cc < 1 || Cc > 4) {
      set_error("cond_channels %d not in [1,4]", Cc);
      return false;
    }
    for (int i = 0; i < 3; ++i) {
      if (!build_small(B, enc + ".blocks." + std::to_string(i) + ".conv1", Cc, Cc, m->enc[2 * i])) return false;
      if (!build_small(B, enc + ".blocks." + std::to_string(i) + ".conv2", Cc, Cc, m->enc[2 * i + 1])) return false;
    }
    if (!build_small(B, enc + ".conv_out", Cc, Cc, m->enc[6])) return false;
    if (!build_small(B, cc, 16, Cc, m->cond_conv)) return false;
  }
  {
    std::vector<float> f(50);
    auto it = m->sd.find("pos_encoding.inv_freq");
    if (it != m->sd.end() && it->second.size() == 50) {
      f = it->second;
    } else {
      // 1 / 10000^(2j/100), the fp32 evaluation order of UNet_model_superres.py:329-331
      for (int j = 0; j < 50; ++j) f[j] = 1.0f / powf(10000.0f, static_cast<float>(2 * j) / 100.0f);
    }
    m->inv_freq = B.push(f);
  }
  if (d.kind == DRS_MODEL_GENERATION && d.num_classes > 0) {
    const auto* e = B.get("label_emb.weight", static_cast<size_t>(d.num_classes) * 100);
    if (!e) return false;
    m->label_emb = B.push(*e);
  }

  // ---- time MLPs -------------------------------------------------------------------------------
  for (int i = 0; i < 3; ++i)
    if (!build_time_mlp(B, "conv_blocks." + std::to_string(i), mlp_C[i], m->mlps[i])) return false;
  if (!build_time_mlp(B, "bottle_neck", 256, m->mlps[3])) return false;
  for (int i = 0; i < 3; ++i) {
    TimeMlp& t = m->mlps[4 + i];
    const int C = mlp_C[4 + i];
    if (!build_time_mlp(B, "ups." + std::to_string(i), C, t)) return false;
    // tap-major fp32 copy of ups.i.conv.weight: Wt[(ky*3+kx)*C + oc][ci]
    const auto* w = B.get("ups." + std::to_string(i) + ".conv.weight", static_cast<size_t>(C) * C * 9);
    if (!w) return false;
    std::vector<float> wt(static_cast<size_t>(9) * C * C);
    for (int oc = 0; oc < C; ++oc)
      for (int ci = 0; ci < C; ++ci)
        for (int k = 0; k < 9; ++k)
          wt[(static_cast<size_t>(k) * C + oc) * C + ci] = (*w)[(static_cast<size_t>(oc) * C + ci) * 9 + k];
    t.wtap = B.push(wt);
  }

  // ---- encoder path ----------------------------------------------------------------------------
  const char* lvl_in[4] = {"h0", "d0", "d1", "d2"};
  const char* lvl_mid[4] = {"b0.h", "b1.h", "b2.h", "bn.h"};
  const char* lvl_out[4] = {"b0.out", "b1.out", "b2.out", "bn.out"};
  for (int i = 0; i < 3; ++i) {
    const std::string p = "conv_blocks." + std::to_string(i);
    if (!build_res_block(B, p, down[i], down[i + 1], i == 0, lvl_in[i], lvl_mid[i], lvl_out[i], m->mlps[i].te_off))
      return false;
    GemmSpec g;
    g.name = "downs." + std::to_string(i);
    g.src_name[0] = lvl_out[i];
    g.out_name = lvl_in[i + 1];
    const int C = down[i + 1];
    g.OC = C;
    g.n_sub = n_sub_for(C);
    const auto* w = B.get(g.name + ".weight", static_cast<size_t>(C) * C * 9);
    const auto* b = B.get(g.name + ".bias", C);
    if (!w || !b) return false;
    g.bias = B.push(*b);
    if (!B.build(g, {term(0, C, CONV_3x3_S2, {wref(w, C, C)})})) return false;
    m->gemms.push_back(std::move(g));
  }
  if (!build_res_block(B, "bottle_neck", 128, 256, false, "d2", "bn.h", "bn.out", m->mlps[3].te_off)) return false;

  // ---- decoder path ----------------------------------------------------------------------------
  const int up[4] = {256, 128, 64, 32};
  std::string xin = "bn.out";
  for (int i = 0; i < 3; ++i) {
    const int C = up[i], Ch = up[i + 1];
    const std::string si = std::to_string(i);
    const std::string skip = lvl_out[2 - i];
    // gating_signal: relu(bn(conv1x1(x)))   (UNet_model_superres.py:209-225)
    {
      GemmSpec g;
      g.name = "gating_signals." + si;
      g.src_name[0] = xin;
      g.out_name = "g" + si;
      g.OC = Ch;
      g.n_sub = n_sub_for(Ch);
      g.flags = F_RELU;
      const auto* w = B.get(g.name + ".conv.weight", static_cast<size_t>(Ch) * C);
      const auto* b = B.get(g.name + ".conv.bias", Ch);
      if (!w || !b) return false;
      std::vector<float> s, bb;
      if (!B.bn_fold(g.name + ".batch_norm", b, Ch, s, bb)) return false;
      g.scale = B.push(s);
      g.bias = B.push(bb);
      if (!B.build(g, {term(0, C, CONV_1x1, {wref(w, Ch, C)})})) return false;
      m->gemms.push_back(std::move(g));
    }
    // attention gate map: psi = sigmoid(w_psi . relu(W_g g + W_x x) + b_psi)   (:101-104)
    const std::string ab = "attention_blocks." + si;
    {
      GemmSpec g;
      g.name = ab + ".psi";
      g.src_name[0] = "g" + si;
      g.src_name[1] = skip;
      g.out_name = "psi" + si;
      g.epi_kind = EPI_PSI;
      g.OC = Ch;
      g.n_sub = Ch;
      const auto* wg = B.get(ab + ".w_g.0.weight", static_cast<size_t>(Ch) * Ch);
      const auto* bg = B.get(ab + ".w_g.0.bias", Ch);
      const auto* wx = B.get(ab + ".w_x.0.weight", static_cast<size_t>(Ch) * Ch * 4);
      const auto* bx = B.get(ab + ".w_x.0.bias", Ch);
      const auto* wp = B.get(ab + ".psi.0.weight", Ch);
      const auto* bp = B.get(ab + ".psi.0.bias", 1);
      if (!wg || !bg || !wx || !bx || !wp || !bp) return false;
      std::vector<float> bsum(Ch);
      for (int c = 0; c < Ch; ++c) bsum[c] = (*bg)[c] + (*bx)[c];
      g.bias = B.push(bsum);
      g.wvec = B.push(*wp);
      g.bvec = B.push(*bp);
      g.nvec = 1;
      if (!B.build(g, {term(0, Ch, CONV_1x1, {wref(wg, Ch, Ch)}, 0), term(1, Ch, CONV_2x2_S2, {wref(wx, Ch, Ch)}, 0)}))
        return false;
      m->gemms.push_back(std::move(g));
    }
    // attention result: bn(conv1x1(psi_up * x)) = bn(psi_up * (W x) + b)   (:105-107)
    {
      GemmSpec g;
      g.name = ab + ".result";
      g.src_name[0] = skip;
      g.src_name[1] = "psi" + si;  // not a TMA source: row scale
      g.out_name = "att" + si;
      g.OC = Ch;
      g.n_sub = n_sub_for(Ch);
      g.flags = F_ROWSCALE;
      const auto* w = B.get(ab + ".result.0.weight", static_cast<size_t>(Ch) * Ch);
      const auto* b = B.get(ab + ".result.0.bias", Ch);
      if (!w || !b) return false;
      std::vector<float> s, bb;
      if (!B.bn_fold(ab + ".result.1", b, Ch, s, bb)) return false;
      g.scale = B.push(s);
      g.bias = B.push(bb);
      if (!B.build(g, {term(0, Ch, CONV_1x1, {wref(w, Ch, Ch)})})) return false;
      g.n_src = 1;
      m->gemms.push_back(std::move(g));
      // fused form of the two launches above (taken by the plan where its program is resident and the grid holds a
      // tile): parameters = result's folded BatchNorm (per output channel) + the whole gate (bias sum, w_psi, b_psi)
      GemmSpec f;
      f.name = ab + ".gate";
      f.src_name[0] = "g" + si;
      f.src_name[1] = skip;
      f.out_name = "att" + si;
      f.psi_name = "psi" + si;
      f.OC = Ch;
      f.n_groups = 4;
      f.oscale = 2;
      f.flags = F_GATE | F_ROWSCALE;
      f.scale = m->gemms.back().scale;
      f.bias = m->gemms.back().bias;
      const GemmSpec& psi_spec = m->gemms[m->gemms.size() - 2];
      f.scale2 = psi_spec.bias;   // b_g + b_x
      f.wvec = psi_spec.wvec;
      f.bvec = psi_spec.bvec;
      f.nvec = Ch;
      const auto* wg2 = B.get(ab + ".w_g.0.weight", static_cast<size_t>(Ch) * Ch);
      const auto* wx2 = B.get(ab + ".w_x.0.weight", static_cast<size_t>(Ch) * Ch * 4);
      if (!wg2 || !wx2 || !B.build_gate(f, Ch, *wg2, *wx2, *w)) return false;
      if (f.v2.usable) B.fused_gates.push_back(std::move(f));
    }
    // UpConvBlock conv: relu(bn(conv3x3(x + relu(time_mlp(t)))))   (:197-205)
    const std::string ub = "ups." + si;
    {
      GemmSpec g;
      g.name = ub + ".conv";
      g.src_name[0] = xin;
      g.out_name = "uc" + si;
      g.OC = C;
      g.n_sub = n_sub_for(C);
      g.flags = F_RELU | F_PRE;
      g.pre_off = m->mlps[4 + i].pre_off;
      const auto* w = B.get(ub + ".conv.weight", static_cast<size_t>(C) * C * 9);
      const auto* b = B.get(ub + ".conv.bias", C);
      if (!w || !b) return false;
      std::vector<float> s, bb;
      if (!B.bn_fold(ub + ".batch_norm", b, C, s, bb)) return false;
      g.scale = B.push(s);
      g.bias = B.push(bb);
      if (!B.build(g, {term(0, C, CONV_3x3, {wref(w, C, C)})})) return false;
      m->gemms.push_back(std::move(g));
    }
    // UpConvBlock transform: ConvTranspose2d(3, s2, p1, op1) as four sub-pixel phases   (:206)
    {
      GemmSpec g;
      g.name = ub + ".transform";
      g.src_name[0] = "uc" + si;
      g.out_name = "ut" + si;
      g.OC = C;
      g.n_sub = std::min(C, 64);
      g.n_groups = 4;
      g.oscale = 2;
      const auto* w = B.get(ub + ".transform.weight", static_cast<size_t>(C) * C * 9);
      const auto* b = B.get(ub + ".transform.bias", C);
      if (!w || !b) return false;
      g.bias = B.push(*b);
      if (!B.build(g, {term(0, C, CONV_T3x3_S2, {wref(w, C, C)})})) return false;
      m->gemms.push_back(std::move(g));
    }
    // up_convs: conv3x3(cat[x_up, att]) with the concatenation as split-K over two sources   (:376-377)
    {
      GemmSpec g;
      g.name = "up_convs." + si;
      g.src_name[0] = "ut" + si;
      g.src_name[1] = "att" + si;
      const int Cin = C + Ch;
      const auto* w = B.get(g.name + ".weight", static_cast<size_t>(Ch) * Cin * 9);
      const auto* b = B.get(g.name + ".bias", Ch);
      if (!w || !b) return false;
      g.OC = Ch;
      g.bias = B.push(*b);
      if (i < 2) {
        g.out_name = "x" + si;
        g.n_sub = n_sub_for(Ch);
      } else {
        // last stage: the 1x1 output conv (UNet_model_superres.py:379) runs on the fp32 accumulator
        g.epi_kind = EPI_OUT;
        g.n_sub = Ch;
        const auto* wo = B.get("output.weight", static_cast<size_t>(d.out_channels) * Ch);
        const auto* bo = B.get("output.bias", d.out_channels);
        if (!wo || !bo) return false;
        if (d.out_channels > 4) {
          set_error("out_channels %d > 4", d.out_channels);
          return false;
        }
        g.wvec = B.push(*wo);
        g.bvec = B.push(*bo);
        g.nvec = d.out_channels;
      }
      if (!B.build(g, {term(0, C, CONV_3x3, {wref(w, Ch, Cin, 0)}, 0), term(1, Ch, CONV_3x3, {wref(w, Ch, Cin, C)}, 0)}))
        return false;
      m->gemms.push_back(std::move(g));
    }
    xin = "x" + si;
  }
  append_narrow(m, B);
  return true;
}

int model_create(const DrsModelDesc* desc, const DrsTensor* tensors, int n_tensors, int device, DrsModel** out) {
  if (!desc || !out || (!tensors && n_tensors > 0)) {
    set_error("drs_model_create: null argument");
    return DRS_E_INVALID;
  }
  if (desc->kind < 0 || desc->kind > 2 || desc->x_channels < 1 || desc->x_channels > 4 || desc->out_channels < 1) {
    set_error("drs_model_create: bad descriptor");
    return DRS_E_INVALID;
  }
  DRS_CUDA(cudaSetDevice(device));
  int major = 0;
  DRS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) {
    set_error("device %d has compute capability %d.x; this library is sm_100a only", device, major);
    return DRS_E_INVALID;
  }
  std::unique_ptr<DrsModel> m(new DrsModel());
  m->desc = *desc;
  m->device = device;
  for (int i = 0; i < n_tensors; ++i) {
    if (!tensors[i].name || (!tensors[i].data && tensors[i].numel > 0)) {
      set_error("drs_model_create: tensor %d is null", i);
      return DRS_E_INVALID;
    }
    m->sd[tensors[i].name].assign(tensors[i].data, tensors[i].data + tensors[i].numel);
  }
  if (!build_model(m.get())) return DRS_E_MISSING;
  DRS_TRY(m->d_fblob.upload(m->fblob.data(), m->fblob.size() * sizeof(float)));
  DRS_TRY(m->d_wblob.upload(m->wblob.data(), m->wblob.size()));
  DRS_TRY(m->d_kblocks.upload(m->kb_all.data(), m->kb_all.size() * sizeof(KBlock)));
  int r = conv_gemm_set_smem_limits();
  if (r == 0) r = conv_gemm2_set_smem_limits();
  if (r == 0) r = conv_gemm2c_set_smem_limits();
  if (r == 0) r = conv_row_set_smem_limits();
  if (r != 0) return cuda_fail(static_cast<cudaError_t>(r), "cudaFuncSetAttribute(conv_gemm_kernel)");
  // the host copy of the state_dict is no longer needed
  m->sd.clear();
  std::vector<uint8_t>().swap(m->wblob);
  *out = m.release();
  return DRS_OK;
}

// Single-layer model for drs_debug_conv2d.
int build_debug_conv(DrsModel* m, const float* w, const float* bias, const float* scale, const float* shift, int Cin,
                     int Cout, int kind, int relu) {
  Builder B(m);
  GemmSpec g;
  g.name = "debug_conv";
  g.src_name[0] = "in";
  g.out_name = "out";
  g.OC = Cout;
  g.flags = relu ? F_RELU : 0;
  std::vector<float> s(Cout, 1.f), b(Cout, 0.f);
  for (int c = 0; c < Cout; ++c) {
    const float sc = scale ? scale[c] : 1.f;
    s[c] = sc;
    b[c] = (bias ? bias[c] : 0.f) * sc + (shift ? shift[c] : 0.f);
  }
  g.scale = B.push(s);
  g.bias = B.push(b);
  WeightRef wr;
  wr.w = w;
  wr.oc = Cout;
  wr.cin_total = Cin;
  if (kind == CONV_T3x3_S2) {
    g.n_sub = std::min(Cout, 64);
    g.n_groups = 4;
    g.oscale = 2;
  } else {
    g.n_sub = n_sub_for(Cout);
  }
  if (!B.build(g, {term(0, Cin, kind, {wr})})) return DRS_E_INVALID;
  m->gemms.push_back(std::move(g));
  append_narrow(m, B);
  DRS_TRY(m->d_fblob.upload(m->fblob.data(), m->fblob.size() * sizeof(float)));
  DRS_TRY(m->d_wblob.upload(m->wblob.data(), m->wblob.size()));
  DRS_TRY(m->d_kblocks.upload(m->kb_all.data(), m->kb_all.size() * sizeof(KBlock)));
  int r = conv_gemm_set_smem_limits();
  if (r == 0) r = conv_gemm2_set_smem_limits();
  if (r == 0) r = conv_gemm2c_set_smem_limits();
  if (r == 0) r = conv_row_set_smem_limits();
  if (r != 0) return cuda_fail(static_cast<cudaError_t>(r), "cudaFuncSetAttribute(conv_gemm_kernel)");
  return DRS_OK;
}

}  // namespace drs
