// extern "C" surface declared in include/drs_b200.h (+ the test / instrumentation hooks of include/drs_b200_diag.h).
#include <string.h>

#include <algorithm>
#include <memory>
#include <vector>

#include "../../include/drs_b200_diag.h"
#include "engine.cuh"
#include "small_kernels.cuh"

namespace drs {
const char* last_error();
int model_create(const DrsModelDesc*, const DrsTensor*, int, int, DrsModel**);
int build_debug_conv(DrsModel*, const float*, const float*, const float*, const float*, int, int, int, int);
int plan_create(DrsModel*, int, int, int, int, int, DrsPlan**);
void plan_destroy(DrsPlan*);
int cond_encode(DrsPlan*, const float*, cudaStream_t);
int time_embed(DrsPlan*, const float*, const int*, cudaStream_t);
int unet_forward(DrsPlan*, const float*, float*, cudaStream_t);
int check_pipeline_error(DrsPlan*, cudaStream_t);
int sampler_prepare(DrsPlan*, int, const float*, const float*, const float*, const int*, float, cudaStream_t);
int sampler_begin(DrsPlan*, float*, float*, float*, int, cudaStream_t);
int sampler_step(DrsPlan*, int, cudaStream_t);
int launches_per_step(const DrsPlan*);
int launch_count(const DrsPlan*);
int launch_info(const DrsPlan*, int, char*, int, double*, double*, int*, int*);
int plan_profile(DrsPlan*, const float*, float*, int, float*, cudaStream_t);
int plan_time_forward(DrsPlan*, const float*, float*, int, float*, cudaStream_t);
int sampler_time_hbm_kernels(DrsPlan*, void*, size_t, int, float*, cudaStream_t);
int debug_bind_and_run(DrsPlan*, const void*, int, int, int, int, void*, int, int, cudaStream_t);
}  // namespace drs

using namespace drs;

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

const char* drs_last_error(void) { return last_error(); }
int drs_version(void) { return 100; }

int drs_model_create(const DrsModelDesc* desc, const DrsTensor* tensors, int n_tensors, int device, DrsModel** out) {
  return model_create(desc, tensors, n_tensors, device, out);
}
void drs_model_destroy(DrsModel* m) { delete m; }

int drs_plan_create(DrsModel* m, int nb, int nx, int ncond, int S, int magnification, DrsPlan** out) {
  return plan_create(m, nb, nx, ncond, S, magnification, out);
}
void drs_plan_destroy(DrsPlan* p) { plan_destroy(p); }
size_t drs_plan_workspace_bytes(const DrsPlan* p) {
  return p ? p->workspace.bytes + p->cond_feat.bytes + p->cond_tmp.bytes + p->table.bytes : 0;
}

int drs_cond_encode(DrsPlan* p, const float* cond_dev, void* stream) {
  if (!p) {
    set_error("null plan");
    return DRS_E_INVALID;
  }
  return cond_encode(p, cond_dev, as_stream(stream));
}
int drs_time_embed(DrsPlan* p, const float* t_dev, const int32_t* label_dev, void* stream) {
  if (!p) {
    set_error("null plan");
    return DRS_E_INVALID;
  }
  return time_embed(p, t_dev, label_dev, as_stream(stream));
}
int drs_unet_forward(DrsPlan* p, const float* x_dev, float* eps_dev, void* stream) {
  if (!p) {
    set_error("null plan");
    return DRS_E_INVALID;
  }
  return unet_forward(p, x_dev, eps_dev, as_stream(stream));
}
int drs_plan_check(DrsPlan* p, void* stream) {
  if (!p) {
    set_error("null plan");
    return DRS_E_INVALID;
  }
  return check_pipeline_error(p, as_stream(stream));
}

int drs_sampler_prepare(DrsPlan* p, int noise_steps, const float* c1, const float* c2, const float* c3,
                        const int32_t* labels_host, float cfg_scale, void* stream) {
  if (!p) {
    set_error("null plan");
    return DRS_E_INVALID;
  }
  return sampler_prepare(p, noise_steps, c1, c2, c3, labels_host, cfg_scale, as_stream(stream));
}
int drs_sampler_begin(DrsPlan* p, float* x_dev, float* noise_dev, float* eps_dev, int start_step, void* stream) {
  if (!p) {
    set_error("null plan");
    return DRS_E_INVALID;
  }
  return sampler_begin(p, x_dev, noise_dev, eps_dev, start_step, as_stream(stream));
}
int drs_sampler_step(DrsPlan* p, int use_graph, void* stream) {
  if (!p) {
    set_error("null plan");
    return DRS_E_INVALID;
  }
  return sampler_step(p, use_graph, as_stream(stream));
}
int drs_sampler_launches_per_step(const DrsPlan* p) { return p ? launches_per_step(p) : 0; }

int drs_plan_launch_count(const DrsPlan* p) { return p ? launch_count(p) : 0; }
int drs_plan_launch_info(const DrsPlan* p, int index, char* name, int name_capacity, double* flops, double* bytes,
                         int* ctas, int* smem_bytes) {
  if (!p) {
    set_error("null plan");
    return DRS_E_INVALID;
  }
  return launch_info(p, index, name, name_capacity, flops, bytes, ctas, smem_bytes);
}
int drs_plan_profile(DrsPlan* p, const float* x_dev, float* eps_dev, int iters, float* ms_out, void* stream) {
  if (!p) {
    set_error("null plan");
    return DRS_E_INVALID;
  }
  return plan_profile(p, x_dev, eps_dev, iters, ms_out, as_stream(stream));
}
int drs_plan_time_forward(DrsPlan* p, const float* x_dev, float* eps_dev, int iters, float* ms_out2, void* stream) {
  if (!p) {
    set_error("null plan");
    return DRS_E_INVALID;
  }
  return plan_time_forward(p, x_dev, eps_dev, iters, ms_out2, as_stream(stream));
}

int drs_sampler_time_hbm_kernels(DrsPlan* p, void* flush_dev, size_t flush_bytes, int iters, float* ms_out2,
                                 void* stream) {
  if (!p) {
    set_error("null plan");
    return DRS_E_INVALID;
  }
  return sampler_time_hbm_kernels(p, flush_dev, flush_bytes, iters, ms_out2, as_stream(stream));
}

int drs_ddpm_update(float* x_dev, const float* eps_dev, const float* noise_dev_or_null, float c1, float c2, float c3,
                    size_t numel, void* stream) {
  if (!x_dev || !eps_dev) {
    set_error("drs_ddpm_update: null buffer");
    return DRS_E_INVALID;
  }
  if (numel % 4) {
    set_error("drs_ddpm_update: numel must be a multiple of 4");
    return DRS_E_INVALID;
  }
  // scalars travel as a one-row coefficient table so the kernel is the one the sampler uses
  struct Tmp { float c[4]; int step; int pad[3]; };
  static thread_local void* dev = nullptr;
  if (!dev) DRS_CUDA(cudaMalloc(&dev, sizeof(Tmp)));
  Tmp h{{c1, c2, c3, 0.f}, 0, {0, 0, 0}};
  cudaStream_t st = as_stream(stream);
  DRS_CUDA(cudaMemcpyAsync(dev, &h, sizeof(Tmp), cudaMemcpyHostToDevice, st));
  const float* coef = reinterpret_cast<const float*>(dev);
  int* step = reinterpret_cast<int*>(reinterpret_cast<char*>(dev) + offsetof(Tmp, step));
  DRS_CUDA(static_cast<cudaError_t>(launch_ddpm_update(x_dev, eps_dev, noise_dev_or_null, coef, step, numel, 0, 0.f, nullptr, 0, 0, nullptr, st)));
  DRS_CUDA(cudaStreamSynchronize(st));  // `h` is a stack temporary
  return DRS_OK;
}

int drs_noise_images(const float* x_dev, const float* eps_dev, const float* sqrt_ah_dev, const float* sqrt_1m_ah_dev,
                     float* out_dev, int n, size_t per_sample, void* stream) {
  if (!x_dev || !eps_dev || !sqrt_ah_dev || !sqrt_1m_ah_dev || !out_dev || n < 1) {
    set_error("drs_noise_images: bad arguments");
    return DRS_E_INVALID;
  }
  if (per_sample % 4) {
    set_error("drs_noise_images: per-sample element count must be a multiple of 4");
    return DRS_E_INVALID;
  }
  DRS_CUDA(static_cast<cudaError_t>(launch_noise_images(x_dev, eps_dev, sqrt_ah_dev, sqrt_1m_ah_dev, out_dev, n,
                                                         per_sample, as_stream(stream))));
  return DRS_OK;
}

int drs_blend(const float* patches_dev, const int32_t* coords4_host, int n_patches, const float* weight_dev,
              float* out_dev, float* wsum_dev, int C, int H, int W, int P, int do_clamp, void* stream) {
  if (!patches_dev || !coords4_host || !weight_dev || !out_dev || !wsum_dev || n_patches < 1 || C < 1 || C > 4) {
    set_error("drs_blend: bad arguments");
    return DRS_E_INVALID;
  }
  cudaStream_t st = as_stream(stream);
  for (int i = 0; i < n_patches; ++i) {
    const int32_t* c = coords4_host + 4 * i;
    if (c[1] - c[0] != P || c[3] - c[2] != P || c[0] < 0 || c[2] < 0 || c[1] > H || c[3] > W) {
      set_error("drs_blend: patch %d window (%d,%d,%d,%d) does not fit [%d,%d] with P=%d", i, c[0], c[1], c[2], c[3],
                H, W, P);
      return DRS_E_INVALID;
    }
  }
  // Row-major grid of windows (what patchifier produces): gather form, one pass, patch-order accumulation.
  std::vector<int> ys, xs;
  for (int i = 0; i < n_patches; ++i) {
    const int y = coords4_host[4 * i], x = coords4_host[4 * i + 2];
    if (ys.empty() || y != ys.back()) {
      if (std::find(ys.begin(), ys.end(), y) == ys.end()) ys.push_back(y);
    }
    if (std::find(xs.begin(), xs.end(), x) == xs.end()) xs.push_back(x);
  }
  bool grid = (static_cast<int>(ys.size() * xs.size()) == n_patches) && std::is_sorted(ys.begin(), ys.end()) &&
              std::is_sorted(xs.begin(), xs.end());
  if (grid)
    for (int i = 0; i < n_patches && grid; ++i)
      grid = coords4_host[4 * i] == ys[i / xs.size()] && coords4_host[4 * i + 2] == xs[i % xs.size()];
  int* d_flag = nullptr;
  int h_flag = 0;
  if (grid) {
    // coverage check on the host: every row / column must fall in some window
    std::vector<char> cy(H, 0), cx(W, 0);
    for (int y : ys) std::fill(cy.begin() + y, cy.begin() + y + P, 1);
    for (int x : xs) std::fill(cx.begin() + x, cx.begin() + x + P, 1);
    if (std::find(cy.begin(), cy.end(), 0) != cy.end() || std::find(cx.begin(), cx.end(), 0) != cx.end()) {
      set_error("drs_blend: some output pixels are covered by no patch");
      return DRS_E_INVALID;
    }
    // device tables, cached per calling thread while the window grid stays the same (the aggregation CLI blends one
    // scene geometry over and over): [ys | xs | pad] then int2 row ranges [H] and int2 quad ranges [W / 4]
    const bool vec = (P % 4 == 0) && (W % 4 == 0) &&
                     std::all_of(xs.begin(), xs.end(), [](int x) { return x % 4 == 0; });
    struct Cache {
      std::vector<int> key;
      DevMem tables;
      int device = -1;
    };
    static thread_local Cache cache;
    std::vector<int> key{H, W, P, vec ? 1 : 0, static_cast<int>(ys.size())};
    key.insert(key.end(), ys.begin(), ys.end());
    key.insert(key.end(), xs.begin(), xs.end());
    int device = 0;
    DRS_CUDA(cudaGetDevice(&device));
    const size_t n_starts = (ys.size() + xs.size() + 1) & ~static_cast<size_t>(1);  // keeps the int2 tables aligned
    if (cache.key != key || cache.device != device || !cache.tables.p) {
      std::vector<int> host(n_starts, 0);
      std::copy(ys.begin(), ys.end(), host.begin());
      std::copy(xs.begin(), xs.end(), host.begin() + ys.size());
      if (vec) {
        auto ranges = [&](const std::vector<int>& starts, int extent, int step) {
          // starts are sorted: the windows covering [p, p + step) form one index interval
          for (int p0 = 0; p0 < extent; p0 += step) {
            int lo = 0, hi = 0;
            while (lo < static_cast<int>(starts.size()) && starts[lo] + P <= p0) ++lo;
            hi = lo;
            while (hi < static_cast<int>(starts.size()) && starts[hi] <= p0) ++hi;
            host.push_back(lo);
            host.push_back(hi);
          }
        };
        ranges(ys, H, 1);
        ranges(xs, W, 4);
      }
      DRS_CUDA(cudaStreamSynchronize(st));  // a previous blend on this stream may still read the old tables
      DRS_TRY(cache.tables.upload(host.data(), host.size() * sizeof(int)));
      cache.key = key;
      cache.device = device;
    }
    const int* d_ys = cache.tables.as<int>();
    const int* d_xs = d_ys + ys.size();
    if (vec) {
      const int2* row_rng = reinterpret_cast<const int2*>(d_ys + n_starts);
      const int2* col_rng = row_rng + H;
      DRS_CUDA(static_cast<cudaError_t>(launch_blend_gather4(patches_dev, d_ys, static_cast<int>(ys.size()), d_xs,
                                                             static_cast<int>(xs.size()), row_rng, col_rng, weight_dev,
                                                             out_dev, wsum_dev, C, H, W, P, do_clamp, st)));
    } else {
      DRS_CUDA(static_cast<cudaError_t>(launch_blend_gather(patches_dev, d_ys, static_cast<int>(ys.size()), d_xs,
                                                            static_cast<int>(xs.size()), weight_dev, out_dev, wsum_dev,
                                                            C, H, W, P, do_clamp, st)));
    }
    return DRS_OK;
  }
  // arbitrary window list: one scatter launch per patch keeps the reference's summation order
  DRS_CUDA(cudaMemsetAsync(out_dev, 0, static_cast<size_t>(C) * H * W * sizeof(float), st));
  DRS_CUDA(cudaMemsetAsync(wsum_dev, 0, static_cast<size_t>(H) * W * sizeof(float), st));
  const size_t pstride = static_cast<size_t>(C) * P * P;
  for (int i = 0; i < n_patches; ++i)
    DRS_CUDA(static_cast<cudaError_t>(launch_blend_accumulate(patches_dev + i * pstride, weight_dev, out_dev, wsum_dev,
                                                              C, H, W, P, coords4_host[4 * i],
                                                              coords4_host[4 * i + 2], st)));
  DRS_CUDA(cudaMalloc(&d_flag, sizeof(int)));
  cudaMemsetAsync(d_flag, 0, sizeof(int), st);
  const int r = launch_blend_finalize(out_dev, wsum_dev, C, H, W, do_clamp, d_flag, st);
  cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  const cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(d_flag);
  DRS_CUDA(static_cast<cudaError_t>(r));
  DRS_CUDA(se);
  if (h_flag) {
    set_error("drs_blend: some output pixels are covered by no patch");
    return DRS_E_INVALID;
  }
  return DRS_OK;
}

int drs_debug_conv2d(const float* x_dev, const float* w_host, const float* bias_host, const float* scale_host,
                     const float* shift_host, float* y_dev, int B, int Cin, int Cout, int H, int W, int kind, int relu,
                     int device, void* stream) {
  if (!x_dev || !w_host || !y_dev || kind < 0 || kind > 4) {
    set_error("drs_debug_conv2d: bad arguments");
    return DRS_E_INVALID;
  }
  cudaStream_t st = as_stream(stream);
  DRS_CUDA(cudaSetDevice(device));
  std::unique_ptr<DrsModel> m(new DrsModel());
  m->device = device;
  DRS_TRY(build_debug_conv(m.get(), w_host, bias_host, scale_host, shift_host, Cin, Cout, kind, relu));
  const GemmSpec& g = m->gemms[0];
  std::unique_ptr<DrsPlan> p(new DrsPlan());
  p->m = m.get();
  p->nb = B;
  const bool s2 = g.src_stride2[0] != 0;
  if (s2 && ((H | W) & 1)) {
    set_error("drs_debug_conv2d: stride-2 kinds need even H and W");
    return DRS_E_INVALID;
  }
  const int gh = s2 ? H / 2 : H, gw = s2 ? W / 2 : W;
  const int OH = gh * g.oscale, OW = gw * g.oscale;
  DevMem in_bf, out_bf, err;
  DRS_TRY(in_bf.alloc(static_cast<size_t>(B) * H * W * Cin * 2));
  DRS_TRY(out_bf.alloc(static_cast<size_t>(B) * OH * OW * Cout * 2));
  DRS_TRY(err.alloc(16));
  DRS_CUDA(cudaMemsetAsync(err.p, 0, 16, st));
  DRS_CUDA(cudaMemsetAsync(out_bf.p, 0, out_bf.bytes, st));
  p->d_err = err.as<int>();
  DRS_CUDA(static_cast<cudaError_t>(launch_nchw_to_nhwc_bf16(x_dev, in_bf.p, B, Cin, H, W, st)));
  DRS_TRY(debug_bind_and_run(p.get(), in_bf.p, gw, gh, H, W, out_bf.p, OH, OW, st));
  DRS_CUDA(static_cast<cudaError_t>(launch_nhwc_bf16_to_nchw(out_bf.p, y_dev, B, Cout, OH, OW, st)));
  DRS_TRY(check_pipeline_error(p.get(), st));
  return DRS_OK;
}

int drs_debug_l2_flush(const void* buf_dev, size_t bytes, void* stream) {
  if (!buf_dev || bytes < 16) {
    set_error("drs_debug_l2_flush: bad arguments");
    return DRS_E_INVALID;
  }
  DRS_CUDA(static_cast<cudaError_t>(launch_l2_flush_read(buf_dev, bytes, as_stream(stream))));
  return DRS_OK;
}

int drs_debug_spans(unsigned long long* out_host, int reset) {
  const int r = conv_gemm2_spans(out_host, reset);
  if (r != 0) return cuda_fail(static_cast<cudaError_t>(r), "conv_gemm2_spans");
  return DRS_OK;
}

int drs_debug_timeline(long long* out_host, int n) {
  const int r = conv_gemm2_read_timeline(out_host, n);
  if (r != 0) return cuda_fail(static_cast<cudaError_t>(r), "cudaMemcpyFromSymbol(g_timeline)");
  return DRS_OK;
}

int64_t drs_debug_fetch(DrsPlan* p, const char* name, float* out_dev, int64_t capacity, void* stream) {
  if (!p || !name || !out_dev) {
    set_error("drs_debug_fetch: null argument");
    return DRS_E_INVALID;
  }
  auto it = p->acts.find(name);
  if (it == p->acts.end()) {
    set_error("drs_debug_fetch: no activation named '%s'", name);
    return DRS_E_INVALID;
  }
  const ActTensor& t = it->second;
  const int64_t numel = static_cast<int64_t>(p->nb) * t.C * t.H * t.W;
  if (capacity < numel) {
    set_error("drs_debug_fetch: capacity %lld < %lld", static_cast<long long>(capacity), static_cast<long long>(numel));
    return DRS_E_INVALID;
  }
  cudaStream_t st = as_stream(stream);
  const uint8_t* src = p->workspace.as<uint8_t>() + t.offset;
  if (t.fp32_map) {
    DRS_CUDA(cudaMemcpyAsync(out_dev, src, numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    DRS_CUDA(static_cast<cudaError_t>(launch_nhwc_bf16_to_nchw(src, out_dev, p->nb, t.C, t.H, t.W, st)));
  }
  return numel;
}

}  // extern "C"
