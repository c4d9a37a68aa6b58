// Row-streaming tcgen05 convolution kernel (see conv_row.cuh for the model).
//
// CTA = 384 threads, persistent, two pipelines with one contiguous range of (image, column strip, row) tiles each:
//   warps 0..3   epilogue of pipeline 0 (tcgen05.ld, fused affine / ReLU / adds, staged TMA stores or the fused 1x1
//                output convolution); with a single pipeline: its output rows n = 0, 2, 4, ...
//   warps 4..7   epilogue of pipeline 1; with a single pipeline: rows n = 1, 3, 5, ...
//   warps 8, 9   tcgen05.mma issuers of pipeline 0 / 1 (one elected lane each); warp 8 owns the TMEM allocation
//   warps 10, 11 producers: the weight image once (warp 10); per input row one TMA box (130 pixels x channel block)
//                per sub-tile
// Barriers: A slots (full / empty), accumulator-ring slots (full: the output row is complete, empty: 128 arrivals of
// the draining group). Waits are bounded like in conv_gemm2 (a pipeline bug sets the error word instead of hanging).
// Launched as a programmatic dependent: everything before griddep_wait() touches only sampling-invariant data.
#include <stdlib.h>
#include <string.h>

#include "conv_epilogue.cuh"
#include "conv_row.cuh"
#include "conv_gemm2.cuh"
#include "ptx.cuh"

namespace drs {

// Barrier wait of this kernel: the poll inline, the bounded back-off loop out of line. The issuer's per-row path has to
// stay small -- with ptx.cuh's fully inlined wait (sixteen unrolled try / sleep rounds at each of its eight call sites)
// one row walked through several KiB of code and ran at ~18 cycles per instruction (instruction-cache misses).
__device__ __noinline__ void row_wait_slow(uint64_t* bar, uint32_t parity, int* err, int code) {
  for (int tries = 0; tries < (1 << 15); ++tries) {
    if (mbar_try_wait_hint(bar, parity, 20000u)) return;
    if (err && (tries & 15) == 15 && *reinterpret_cast<volatile int*>(err) != 0) return;
  }
  if (err) atomicCAS(err, 0, code);
}
__device__ __forceinline__ void row_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  row_wait_slow(bar, parity, err, code);
}

struct RowRing {
  int idx;
  uint32_t phase;
  __device__ __forceinline__ void advance(int n) {
    if (++idx == n) {
      idx = 0;
      phase ^= 1u;
    }
  }
};

struct RowSeg {
  int b, x0, r0, r1;  // image, first pixel of the strip, output rows [r0, r1)
};

// The CTA's tile range [L, L1) of the flattened index ((b * tiles_x + tx) * H + y), cut at strip ends.
struct RowWalk {
  long long L, L1;
  int H, tiles_x;
  __device__ __forceinline__ bool next(RowSeg& s) {
    if (L >= L1) return false;
    const long long strip = L / H;
    s.r0 = static_cast<int>(L - strip * H);
    const long long rem = L1 - L;
    s.r1 = (rem < static_cast<long long>(H - s.r0)) ? s.r0 + static_cast<int>(rem) : H;
    s.b = static_cast<int>(strip / tiles_x);
    s.x0 = static_cast<int>(strip % tiles_x) * kRowTile;
    L += s.r1 - s.r0;
    return true;
  }
};

// Debug timeline (DRS_V2_TIMELINE=1): pipeline 0 of CTA 0, first 32 rows it touches, 16 slots per row --
// 0 issuer at row start, 1 after the ring-slot (empty) waits, 2 after the first A-full wait, 3 after the last MMA issue,
// 4 producer after the A-empty wait of the row's first sub-tile, 5 epilogue before / 6 after the accumulator-full wait,
// 7 epilogue done; issuer detail: 8 after the run construction (before the ring-slot waits), 9 / 10 / 11 first
// sub-tile after elect / after its MMAs / after its commit, 12 after the output-row commits.
// Issuer / producer rows count input rows, epilogue rows count output rows.
#define RTL(row_no, slot)                                                                                      \
  do {                                                                                                         \
    if (a.timeline && blockIdx.x == 0 && pipe == 0 && (row_no) < 32) a.timeline[(row_no) * 16 + (slot)] = clock64(); \
  } while (0)

// The three horizontal taps of one 3x3 sub-tile against one run of output rows: 3 * NK MMAs whose descriptors were all
// formed before the first one is issued (a record fetched, added and moved to the uniform datapath per MMA costs the
// issuing thread ~250 cycles of dependent latency; formed up front the chains overlap).
template <int NK>
__device__ __forceinline__ void row_issue3(uint32_t d, uint32_t idesc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b0,
                                           uint32_t b1, uint32_t b2, uint32_t a_hi, uint32_t b_hi, uint32_t skip_first) {
  umma_bf16_split_if(skip_first == 0u, d, a0, a_hi, b0, b_hi, idesc, 1u);
#pragma unroll
  for (int k = 1; k < NK; ++k) umma_bf16_split(d, a0 + 2u * k, a_hi, b0 + 2u * k, b_hi, idesc, 1u);
#pragma unroll
  for (int k = 0; k < NK; ++k) umma_bf16_split(d, a1 + 2u * k, a_hi, b1 + 2u * k, b_hi, idesc, 1u);
#pragma unroll
  for (int k = 0; k < NK; ++k) umma_bf16_split(d, a2 + 2u * k, a_hi, b2 + 2u * k, b_hi, idesc, 1u);
}

// What one input row does to the accumulator ring, worked out once per row: at most three "runs" of consecutive ring
// slots for the ordinary MMAs (ro: split only where the ring wraps) and for the row's initialising MMA (rf: split also
// where overwrite / accumulate changes), with their TMEM addresses and instruction descriptors.
struct RowPlan {
  uint32_t od[3], oi[3], ro_g0[3];
  uint32_t fd[3], fi[3], rf_g0[3], rf_acc[3];
  int n_ro, n_rf;
  uint32_t slot0, slot_c;     // ring slots of the output row above this input row / of its own (centre) row
  int n_init;                 // running index of the first output row this input row initialises
  uint32_t init_par;          // parity of that slot's "drained" barrier
  bool has_init, init_two;    // initialises one row / two rows (input row 0 of an image)
  bool centre, commit_prev, commit_own;
};

__device__ __forceinline__ void row_plan(RowPlan& P, int yi, const RowSeg& sg, int n_base, int H, int smask, int sshift,
                                         uint32_t aw0, uint32_t d_ring0, uint32_t idesc0) {
  P.centre = (yi >= sg.r0) && (yi < sg.r1);
  // output rows this input row feeds through the three vertical taps: group g <-> output row yi - 1 + g
  const int gl = max(sg.r0 - (yi - 1), 0), gh = min(sg.r1 - 1 - (yi - 1), 2);
  const int n_g0 = n_base + (yi - 1 - sg.r0);  // running index of group 0's output row
  const int slot0 = n_g0 & smask;
  P.slot0 = static_cast<uint32_t>(slot0);
  P.slot_c = static_cast<uint32_t>((n_g0 + 1) & smask);
  // An output row is initialised by its first input row: yi + 1 always is, row 0 also by input row 0.
  P.n_init = n_g0 + ((yi == 0 && gl <= 1) ? 1 : 2);
  P.has_init = (gh == 2) || (yi == 0 && gh >= 1);
  P.init_two = (yi == 0 && gl <= 1 && gh == 2);
  P.init_par = (static_cast<uint32_t>(P.n_init >> sshift) & 1u) ^ 1u;
  P.commit_prev = (yi - 1 >= sg.r0);
  P.commit_own = (yi == H - 1 && sg.r1 == H);
  uint32_t ro_slot[3], ro_ng[3], rf_slot[3], rf_ng[3];
  if (gl == 0 && gh == 2 && slot0 + 2 <= smask && yi != 0) {
    // the common row: three output rows, no ring wrap -> one MMA per tap and K slice; the initialising one is split
    // into [rows above and same: accumulate] and [row below: overwrite]
    P.n_ro = 1;
    ro_slot[0] = static_cast<uint32_t>(slot0); ro_ng[0] = 3; P.ro_g0[0] = 0;
    ro_slot[1] = ro_slot[2] = 0; ro_ng[1] = ro_ng[2] = 0; P.ro_g0[1] = P.ro_g0[2] = 0;
    P.n_rf = 2;
    rf_slot[0] = static_cast<uint32_t>(slot0); rf_ng[0] = 2; P.rf_g0[0] = 0; P.rf_acc[0] = 1;
    rf_slot[1] = static_cast<uint32_t>(slot0 + 2); rf_ng[1] = 1; P.rf_g0[1] = 2; P.rf_acc[1] = 0;
    rf_slot[2] = 0; rf_ng[2] = 0; P.rf_g0[2] = 0; P.rf_acc[2] = 1;
  } else {
    // link(g): groups g and g + 1 may share an MMA -- their slots are consecutive (no ring wrap between them) and,
    // for the initialising MMA, they agree on overwrite / accumulate
    const bool init1 = (yi == 0);                              // group 1 is initialised only by input row 0
    const bool lo0 = (slot0 != smask), lo1 = (((n_g0 + 1) & smask) != smask);
    const bool lf0 = lo0 && !init1, lf1 = lo1 && init1;        // inits: g0 never, g1 iff row 0, g2 always
    P.n_ro = P.n_rf = 0;
    int g = gl;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int cnt = 0;
      if (g <= gh) {
        cnt = 1;
        if (g + 1 <= gh && (g == 0 ? lo0 : lo1)) {
          cnt = 2;
          if (g == 0 && gh == 2 && lo1) cnt = 3;
        }
        P.n_ro = r + 1;
      }
      P.ro_g0[r] = static_cast<uint32_t>(g);
      ro_ng[r] = static_cast<uint32_t>(cnt);
      ro_slot[r] = static_cast<uint32_t>((n_g0 + g) & smask);
      g += (cnt > 0) ? cnt : 1;
    }
    g = gl;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int cnt = 0;
      if (g <= gh) {
        cnt = 1;
        if (g + 1 <= gh && (g == 0 ? lf0 : lf1)) {
          cnt = 2;
          if (g == 0 && gh == 2 && lf1) cnt = 3;
        }
        P.n_rf = r + 1;
      }
      P.rf_g0[r] = static_cast<uint32_t>(g);
      rf_ng[r] = static_cast<uint32_t>(cnt);
      rf_slot[r] = static_cast<uint32_t>((n_g0 + g) & smask);
      P.rf_acc[r] = (g == 2 || (g == 1 && init1)) ? 0u : 1u;
      g += (cnt > 0) ? cnt : 1;
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    P.od[r] = d_ring0 + ro_slot[r] * aw0;
    P.oi[r] = idesc0 | (((ro_ng[r] * aw0) >> 3) << 17);
    P.fd[r] = d_ring0 + rf_slot[r] * aw0;
    P.fi[r] = idesc0 | (((rf_ng[r] * aw0) >> 3) << 17);
  }
}

// Register-resident program of the issuer: the first two sub-tile records and, when they are 3x3 terms, their tap
// records (weight offsets already rebased to the shared-memory image).
struct RowRegs {
  RowSub T0, T1;
  uint32_t qa0[3], qb0[3], qa1[3], qb1[3];
  uint32_t grp0, grp1;
  bool fast1;
  uint32_t wb, idesc0, aw0, aw1, d_ring0, d_ring1;
};

// All MMAs of one input row (one elected lane). slot16: the row slot's shared-memory address >> 4.
__device__ __forceinline__ void row_issue(const RowProg& prog, const RowRegs& R, const RowPlan& P, uint32_t slot16,
                                          int n_sub) {
  int s_begin = 0;
  {
    // sub-tiles 0 and 1 from registers, one pass over their taps per run of output rows (one run unless the ring
    // wraps inside this row's three output rows). Sub-tile 0 is always the first 3x3 term of ring 0: it carries the
    // initialising MMAs, which cover every output row of this input row.
    {
      const uint32_t sub16 = slot16 + (static_cast<uint32_t>(R.T0.off_kib) << 6);
      const uint32_t a0 = R.qa0[0] + sub16, a1 = R.qa0[1] + sub16, a2 = R.qa0[2] + sub16;
#pragma unroll
      for (int r = 0; r < 3; ++r)
        if (r < P.n_rf)
          umma_bf16_split(P.fd[r], a0, R.T0.a_hi, R.qb0[0] + P.rf_g0[r] * R.grp0, R.T0.b_hi, P.fi[r], P.rf_acc[r]);
      for (int r = 0; r < P.n_ro; ++r) {  // a runtime loop: the MMA sequences below exist once in the code
        const uint32_t d = (r == 0) ? P.od[0] : ((r == 1) ? P.od[1] : P.od[2]);
        const uint32_t id = (r == 0) ? P.oi[0] : ((r == 1) ? P.oi[1] : P.oi[2]);
        const uint32_t boff = ((r == 0) ? P.ro_g0[0] : ((r == 1) ? P.ro_g0[1] : P.ro_g0[2])) * R.grp0;
        if (R.T0.nk == 4)
          row_issue3<4>(d, id, a0, a1, a2, R.qb0[0] + boff, R.qb0[1] + boff, R.qb0[2] + boff, R.T0.a_hi, R.T0.b_hi, 1u);
        else if (R.T0.nk == 2)
          row_issue3<2>(d, id, a0, a1, a2, R.qb0[0] + boff, R.qb0[1] + boff, R.qb0[2] + boff, R.T0.a_hi, R.T0.b_hi, 1u);
        else
          row_issue3<1>(d, id, a0, a1, a2, R.qb0[0] + boff, R.qb0[1] + boff, R.qb0[2] + boff, R.T0.a_hi, R.T0.b_hi, 1u);
      }
      s_begin = 1;
    }
    if (R.fast1) {
      const uint32_t sub16 = slot16 + (static_cast<uint32_t>(R.T1.off_kib) << 6);
      const uint32_t a0 = R.qa1[0] + sub16, a1 = R.qa1[1] + sub16, a2 = R.qa1[2] + sub16;
      for (int r = 0; r < P.n_ro; ++r) {
        const uint32_t d = (r == 0) ? P.od[0] : ((r == 1) ? P.od[1] : P.od[2]);
        const uint32_t id = (r == 0) ? P.oi[0] : ((r == 1) ? P.oi[1] : P.oi[2]);
        const uint32_t boff = ((r == 0) ? P.ro_g0[0] : ((r == 1) ? P.ro_g0[1] : P.ro_g0[2])) * R.grp1;
        if (R.T1.nk == 4)
          row_issue3<4>(d, id, a0, a1, a2, R.qb1[0] + boff, R.qb1[1] + boff, R.qb1[2] + boff, R.T1.a_hi, R.T1.b_hi, 0u);
        else if (R.T1.nk == 2)
          row_issue3<2>(d, id, a0, a1, a2, R.qb1[0] + boff, R.qb1[1] + boff, R.qb1[2] + boff, R.T1.a_hi, R.T1.b_hi, 0u);
        else
          row_issue3<1>(d, id, a0, a1, a2, R.qb1[0] + boff, R.qb1[1] + boff, R.qb1[2] + boff, R.T1.a_hi, R.T1.b_hi, 0u);
      }
      s_begin = 2;
    }
  }
  for (int s = s_begin; s < n_sub; ++s) {
    const RowSub T = (s == 1) ? R.T1 : prog.sub[s];
    if (!T.rows3 && !P.centre) continue;
    const uint32_t sub16 = slot16 + (static_cast<uint32_t>(T.off_kib) << 6);
    const uint32_t a_hi = T.a_hi, b_hi = T.b_hi;
    const int i_end = T.first_mma + T.n_mma;
    const int nk = T.nk;
    if (T.rows3) {
      for (int i = T.first_mma; i < i_end; ++i) {
        // one record per horizontal tap: nk MMAs (K = 16 slices) per run
        const uint4 q = *reinterpret_cast<const uint4*>(&prog.mma[i]);
        const uint32_t a_lo = q.x + sub16, b_lo = q.y + R.wb;
        for (int kk = 0; kk < nk; ++kk) {
#pragma unroll
          for (int r = 0; r < 3; ++r)
            if (r < P.n_ro)
              umma_bf16_split(P.od[r], a_lo + 2u * kk, a_hi, b_lo + P.ro_g0[r] * q.z + 2u * kk, b_hi, P.oi[r], 1u);
        }
      }
    } else {
      // 1x1 term: only the centre row; the first K slice of the ring's first tap overwrites
      const uint32_t dc = (T.ring ? R.d_ring1 + P.slot_c * R.aw1 : R.d_ring0 + P.slot_c * R.aw0);
      const uint32_t ic = R.idesc0 | ((static_cast<uint32_t>(T.aw) >> 3) << 17);
      for (int i = T.first_mma; i < i_end; ++i) {
        const uint4 q = *reinterpret_cast<const uint4*>(&prog.mma[i]);
        for (int kk = 0; kk < nk; ++kk)
          umma_bf16_split(dc, q.x + sub16 + 2u * kk, a_hi, q.y + R.wb + 2u * kk, b_hi, ic,
                          ((q.w & ROWTAP_RING_FIRST) && kk == 0) ? 0u : 1u);
      }
    }
  }
}

constexpr int kRowIssuerWarp0 = kRowEpiWarps;                 // issuer of pipeline p: kRowIssuerWarp0 + p
constexpr int kRowProducerWarp0 = kRowEpiWarps + kRowPipes;   // producer of pipeline p
constexpr int kRowMaxNacc = 64;  // channels per pixel the epilogue of this kernel handles

template <int EPI, int FL>
__global__ void __launch_bounds__(kRowThreads, 1)
conv_row_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                const __grid_constant__ CUtensorMap map_out, const __grid_constant__ RowArgs a,
                const __grid_constant__ RowProg prog) {
  extern __shared__ uint8_t dyn_smem[];
  __shared__ __align__(8) uint64_t s_afull[kRowPipes][kRowMaxASlots], s_aempty[kRowPipes][kRowMaxASlots];
  __shared__ __align__(8) uint64_t s_tfull[kRowPipes][kRowMaxRing], s_tempty[kRowPipes][kRowMaxRing];
  __shared__ __align__(8) uint64_t s_wready;
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_par[4][kMaxN];
  constexpr bool kStageTe = (EPI == EPI_STD && FL >= 0 && (FL & F_TE));
  __shared__ __align__(16) float s_te[kStageTe ? kRowEpiWarps : 1][kStageTe ? kRowMaxNacc : 4];

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const EpiArgs& e = a.epi;
  const int n_pipes = a.n_pipes;
  // debug (DRS_V2_TIMELINE=8): globaltimer at entry / exit of every CTA
  if (a.cta_times && threadIdx.x == 0 && blockIdx.x < 256) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.cta_times[2 * blockIdx.x] = static_cast<long long>(t);
  }
  // role -> pipeline: epilogue warps 0..3 / 4..7 serve pipeline 0 / 1 (or alternate rows of the only pipeline)
  int pipe = 0;
  if (warp >= kRowProducerWarp0) pipe = warp - kRowProducerWarp0;
  else if (warp >= kRowIssuerWarp0) pipe = warp - kRowIssuerWarp0;
  else if (n_pipes == 2) pipe = warp >> 2;

  const uint32_t dyn_u32 = smem_u32(dyn_smem);
  uint8_t* const smem0 = dyn_smem + ((1024u - (dyn_u32 & 1023u)) & 1023u);
  uint8_t* const a_base = smem0 + static_cast<size_t>(pipe) * a.a_slots * a.a_slot_bytes;
  uint8_t* const w_base = smem0 + static_cast<size_t>(n_pipes) * a.a_slots * a.a_slot_bytes;
  uint8_t* const stage_base = w_base + ((a.w_bytes + 1023u) & ~1023u);

  // this pipeline's tile range: ranges are numbered like virtual CTAs, 2 * blockIdx.x + pipe of 2 * gridDim.x
  const long long l_total = static_cast<long long>(a.B) * a.tiles_x * a.H;
  const long long v_id = static_cast<long long>(blockIdx.x) * n_pipes + pipe;
  const long long v_n = static_cast<long long>(gridDim.x) * n_pipes;
  RowWalk walk{l_total * v_id / v_n, l_total * (v_id + 1) / v_n, a.H, a.tiles_x};
  const int S = a.ring_slots;
  const int smask = S - 1;  // S is a power of two
  const int sshift = 31 - __clz(S);
  // TMEM columns: per pipeline [ring 0: S x aw0 | ring 1: S x aw1]
  const uint32_t ring0_col = static_cast<uint32_t>(pipe * S * (a.ring_aw[0] + a.ring_aw[1]));
  const uint32_t ring1_col = ring0_col + static_cast<uint32_t>(S * a.ring_aw[0]);
  uint64_t* const afull = s_afull[pipe];
  uint64_t* const aempty = s_aempty[pipe];
  uint64_t* const tfull = s_tfull[pipe];
  uint64_t* const tempty = s_tempty[pipe];

  // ---- one-time setup --------------------------------------------------------------------------
  load_epilogue_params<EPI>(e, a.n_acc, 0, s_par, threadIdx.x, kRowThreads);
  if (warp == kRowProducerWarp0 && lane == 0) {
    tma_prefetch_desc(&map0);
    tma_prefetch_desc(&map1);
    if (a.store_sbc) tma_prefetch_desc(&map_out);
    for (int p = 0; p < kRowPipes; ++p) {
      for (int s = 0; s < a.a_slots; ++s) {
        mbar_init(&s_afull[p][s], 1);
        mbar_init(&s_aempty[p][s], 1);
      }
      for (int s = 0; s < S; ++s) {
        mbar_init(&s_tfull[p][s], 1);
        mbar_init(&s_tempty[p][s], 128);
      }
    }
    mbar_init(&s_wready, 1);
    fence_mbar_init();
  }
  if (warp == kRowIssuerWarp0) {
    tmem_alloc(&s_tmem_base, static_cast<uint32_t>(a.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  griddep_launch();

  if (warp >= kRowProducerWarp0) {
    // ---- producer of pipeline `pipe` -----------------------------------------------------------------
    if (pipe == 0 && elect_one()) {
      mbar_expect_tx(&s_wready, a.w_bytes);
      for (uint32_t off = 0; off < a.w_bytes; off += 16384u)
        bulk_load(w_base + off, a.wimage + off, min(16384u, a.w_bytes - off), &s_wready);
    }
    __syncwarp();
    griddep_wait();  // activations of the previous layer
    if (pipe < n_pipes) {
      RowRing ar{0, 0};
      RowSeg sg;
      int row_no = 0;
      while (walk.next(sg)) {
        const int yi_lo = max(sg.r0 - 1, 0), yi_hi = min(sg.r1, a.H - 1);
        for (int yi = yi_lo; yi <= yi_hi; ++yi, ++row_no) {
          const bool centre = (yi >= sg.r0) && (yi < sg.r1);
          // one row slot = every sub-tile of this input row, one barrier (a 1x1 term has no use for the halo rows)
          row_wait(&aempty[ar.idx], ar.phase ^ 1u, a.err, 1);
          if (lane == 0) RTL(row_no, 4);
          if (elect_one()) {
            uint8_t* const slot = a_base + static_cast<size_t>(ar.idx) * a.a_slot_bytes;
            mbar_expect_tx(&afull[ar.idx], centre ? a.row_bytes_all : a.row_bytes3);
            for (int s = 0; s < a.n_sub; ++s) {
              const RowSub T = prog.sub[s];
              if (!T.rows3 && !centre) continue;
              tma_load_5d(slot + (static_cast<uint32_t>(T.off_kib) << 10), T.src ? &map1 : &map0, &afull[ar.idx], T.c,
                          sg.x0 - 1, 0, yi, sg.b);
            }
          }
          __syncwarp();
          ar.advance(a.a_slots);
        }
      }
    }
  } else if (warp >= kRowIssuerWarp0) {
    // ---- MMA issuer of pipeline `pipe` -------------------------------------------------------------------
    // Everything that depends on the row (which output rows it feeds, their ring slots, where the ring wraps, which
    // of them this row initialises) is worked out ONCE per input row by the whole warp, as at most three "runs" of
    // consecutive ring slots; per MMA the issuing lane reads one 128-bit record and does three adds. (A first
    // version recomputed slots with integer divisions inside the MMA loop: ~1000 cycles per MMA on one thread.)
    if (pipe < n_pipes) {
      row_wait(&s_wready, 0, a.err, 2);
      RowRing ar{0, 0};
      RowRegs R;
      R.wb = 0x10000u | (smem_u32(w_base) >> 4);
      R.idesc0 = umma_idesc_bf16(kRowTile, 0);
      R.aw0 = static_cast<uint32_t>(a.ring_aw[0]);
      R.aw1 = static_cast<uint32_t>(a.ring_aw[1]);
      R.d_ring0 = tmem + ring0_col;
      R.d_ring1 = tmem + ring1_col;
      R.T0 = prog.sub[0];
      R.T1 = prog.sub[1];
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const RowMma m0 = prog.mma[R.T0.first_mma + t];
        const RowMma m1 = prog.mma[R.T1.first_mma + t];
        R.qa0[t] = m0.a_lo; R.qb0[t] = m0.b_lo + R.wb;
        R.qa1[t] = m1.a_lo; R.qb1[t] = m1.b_lo + R.wb;
      }
      R.grp0 = prog.mma[R.T0.first_mma].grp16;
      R.grp1 = prog.mma[R.T1.first_mma].grp16;
      R.fast1 = (a.n_sub >= 2) && R.T1.rows3 && (R.T1.ring == 0);
      int n_base = 0;  // output rows of this pipeline before the current segment
      int row_no = 0;
      RowSeg sg;
      while (walk.next(sg)) {
        const int yi_lo = max(sg.r0 - 1, 0), yi_hi = min(sg.r1, a.H - 1);
        // TWO input rows per iteration: the barrier polls, the fence, the election and the warp re-convergence cost the
        // issuing warp ~1300 cycles per iteration whatever it issues, more than the MMAs of a low-channel row
        for (int yi = yi_lo; yi <= yi_hi; yi += 2, row_no += 2) {
          const bool two = (yi + 1 <= yi_hi);
          if (lane == 0) RTL(row_no, 0);
          RowRing arB = ar;
          arB.advance(a.a_slots);
          RowPlan PA, PB;
          row_plan(PA, yi, sg, n_base, a.H, smask, sshift, R.aw0, R.d_ring0, R.idesc0);
          // both barrier polls of the first row (the ring slot it initialises, its A row slot) are issued before the
          // second row is planned so that their latencies overlap
          const uint32_t ok_ring = PA.has_init ? mbar_try_wait(&tempty[PA.n_init & smask], PA.init_par) : 1u;
          const uint32_t ok_a = mbar_try_wait(&afull[ar.idx], ar.phase);
          if (two) row_plan(PB, yi + 1, sg, n_base, a.H, smask, sshift, R.aw0, R.d_ring0, R.idesc0);
          if (lane == 0) RTL(row_no, 8);
          if (!ok_ring) row_wait(&tempty[PA.n_init & smask], PA.init_par, a.err, 2);
          // (row 0 of an image initialises two output rows: the second slot was drained after the first, in order)
          if (PA.init_two) {
            const int n2 = PA.n_init + 1;
            row_wait(&tempty[n2 & smask], (static_cast<uint32_t>(n2 >> sshift) & 1u) ^ 1u, a.err, 2);
          }
          tc_fence_after();
          if (lane == 0) RTL(row_no, 1);
          if (!ok_a) row_wait(&afull[ar.idx], ar.phase, a.err, 2);
          if (lane == 0) RTL(row_no, 2);
          if (elect_one()) {
            RTL(row_no, 9);
            row_issue(prog, R, PA, smem_u32(a_base + static_cast<size_t>(ar.idx) * a.a_slot_bytes) >> 4, a.n_sub);
            RTL(row_no, 10);
            // one commit frees the row slot, one (two at the bottom of an image) publishes the completed output rows:
            // the row above this input row, at the bottom of the image also its own
            umma_commit(&aempty[ar.idx]);
            if (PA.commit_prev) umma_commit(&tfull[PA.slot0]);
            if (PA.commit_own) umma_commit(&tfull[PA.slot_c]);
            RTL(row_no, 11);
            if (two) {
              // second row: its waits are taken by the issuing lane alone (nobody else touches what they guard)
              if (PB.has_init) {
                row_wait(&tempty[PB.n_init & smask], PB.init_par, a.err, 2);
                tc_fence_after();
              }
              row_wait(&afull[arB.idx], arB.phase, a.err, 2);
              row_issue(prog, R, PB, smem_u32(a_base + static_cast<size_t>(arB.idx) * a.a_slot_bytes) >> 4, a.n_sub);
              umma_commit(&aempty[arB.idx]);
              if (PB.commit_prev) umma_commit(&tfull[PB.slot0]);
              if (PB.commit_own) umma_commit(&tfull[PB.slot_c]);
            }
          }
          __syncwarp();
          ar.advance(a.a_slots);
          if (two) ar.advance(a.a_slots);
          if (lane == 0) RTL(row_no, 3);
        }
        n_base += sg.r1 - sg.r0;
      }
    }
  } else {
    // ---- epilogue: with two pipelines group (warp / 4) drains every output row of its pipeline, with one pipeline
    // the two groups take alternate rows ------------------------------------------------------------------------
    griddep_wait();  // write-after-read on the output tensor, state read by the epilogue
    const int eg = warp >> 2;
    const int q = warp & 3;
    const int n_step = (n_pipes == 2) ? 1 : 2;
    TmaStoreCtx ts;
    ts.map = &map_out;
    ts.sbc = a.store_sbc;
    ts.nbuf = a.store_sbc ? min(4, kStageBytesPerWarp / (64 * a.store_sbc)) : 1;
    ts.buf = 0;
    ts.stage = stage_base + static_cast<size_t>(warp) * kStageBytesPerWarp;
    int n = 0;
    RowSeg sg;
    while (walk.next(sg)) {
      if (kStageTe) {
        // this image's time-embedding row -> the warp's shared copy
        __syncwarp();
        const float* src = e.te + static_cast<size_t>(__ldg(e.trow + sg.b)) * e.te_stride + e.te_off;
        for (int c = lane * 4; c < a.n_acc; c += 128)
          *reinterpret_cast<float4*>(&s_te[warp][c]) = __ldg(reinterpret_cast<const float4*>(src + c));
        __syncwarp();
      }
      const int x = sg.x0 + q * 32 + lane;
      for (int o = sg.r0; o < sg.r1; ++o, ++n) {
        if (n_step == 2 && (n & 1) != eg) continue;
        const int slot = n & smask;
        if (q == 0 && lane == 0) RTL(n, 5);
        row_wait(&tfull[slot], static_cast<uint32_t>(n >> sshift) & 1u, a.err, 3);
        tc_fence_after();
        if (q == 0 && lane == 0) RTL(n, 6);
        const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) + ring0_col +
                               static_cast<uint32_t>(slot * a.ring_aw[0]);
        ts.x0 = sg.x0 + q * 32;
        ts.y0 = o;
        ts.b = sg.b;
        if constexpr (EPI == EPI_STD)
          conv_epilogue_std_ct<FL>(e, taddr, x, o, sg.b, true, a.W, a.H, a.n_acc, 0, s_par, s_te[kStageTe ? warp : 0],
                                   a.store_sbc ? &ts : nullptr, 0, 1);
        else
          conv_epilogue<EPI>(e, taddr, x, o, sg.b, true, a.W, a.H, a.n_acc, 0, s_par, nullptr, 0, 1);
        tc_fence_before();
        mbar_arrive(&tempty[slot]);
        if (q == 0 && lane == 0) RTL(n, 7);
      }
    }
    if (EPI == EPI_STD && a.store_sbc && lane == 0) bulk_wait_read<0>();
  }

  // ---- teardown ----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kRowIssuerWarp0) {
    tc_fence_after();
    tmem_dealloc(tmem, static_cast<uint32_t>(a.tmem_cols));
  }
  if (a.cta_times && threadIdx.x == 0 && blockIdx.x < 256) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.cta_times[2 * blockIdx.x + 1] = static_cast<long long>(t);
  }
}

static constexpr int kMaxDynSmemRow = 214 * 1024;

// every instantiation: (EPI, FL) -- the flag words of the layers this kernel serves
#define DRS_ROW_VARIANTS(X)                       \
  X(EPI_STD, 0)                                   \
  X(EPI_STD, F_NOSCALE)                           \
  X(EPI_STD, F_RELU)                              \
  X(EPI_STD, F_RELU | F_TE)                       \
  X(EPI_STD, F_RELU | F_TE | F_DUAL_POST)         \
  X(EPI_STD, F_RELU | F_DUAL_PRE)                 \
  X(EPI_STD, F_RELU | F_PRE)                      \
  X(EPI_OUT, -1)

bool conv_row_supports(int epi_kind, int flags) {
  const int fl = (epi_kind == EPI_STD) ? flags : -1;
#define X(EPI, FL) \
  if (epi_kind == EPI && fl == (FL)) return true;
  DRS_ROW_VARIANTS(X)
#undef X
  return false;
}

int conv_row_set_smem_limits() {
  cudaError_t e = cudaSuccess;
#define X(EPI, FL)                                                                                             \
  if (e == cudaSuccess)                                                                                        \
    e = cudaFuncSetAttribute(conv_row_kernel<EPI, (FL)>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmemRow);
  DRS_ROW_VARIANTS(X)
#undef X
  return static_cast<int>(e);
}

int launch_conv_row(int epi_kind, const CUtensorMap& map0, const CUtensorMap& map1, const CUtensorMap& map_out,
                    const RowArgs& args, const RowProg& prog, int grid, size_t smem_bytes, cudaStream_t stream) {
  static const bool no_pdl = (getenv("DRS_V2_NO_PDL") != nullptr);
  const int fl = (epi_kind == EPI_STD) ? args.epi.flags : -1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(kRowThreads, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  cudaError_t err = cudaErrorInvalidValue;
  bool done = false;
#define X(EPI, FL)                                                                                 \
  if (!done && epi_kind == EPI && fl == (FL)) {                                                    \
    err = cudaLaunchKernelEx(&cfg, conv_row_kernel<EPI, (FL)>, map0, map1, map_out, args, prog);   \
    done = true;                                                                                   \
  }
  DRS_ROW_VARIANTS(X)
#undef X
  if (err != cudaSuccess) return static_cast<int>(err);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace drs
