// Non-GEMM kernels of the sampling path (all CUDA-core, HBM- or latency-bound). Host launch wrappers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace drs {

// fp32 NCHW 3x3 conv, stride 1, zero pad 1, tiny channel counts (Cin <= 4, Cout <= 16): condition encoder.
//   out = [relu](conv(in) + bias) [+ residual]; out layout NCHW fp32, or NHWC fp32 when nhwc_out != 0.
int launch_conv3x3_small(const float* in, const float* w, const float* bias, const float* residual, float* out, int B,
                         int Cin, int Cout, int H, int W, int relu, int nhwc_out, cudaStream_t s);

// PyTorch-compatible bicubic upsample (align_corners=False, A=-0.75, scale given as integer factor k), NCHW fp32.
int launch_bicubic_up(const float* in, float* out, int B, int C, int H, int W, int k, cudaStream_t s);

// conv0: h0[b,y,x,0:16] = bf16( conv3x3(x[b % nx])[0:16] + bias + cond[(b % ncond)] ), x fp32 NCHW [nx,Cx,S,S],
// cond fp32 channel-planar [ncond,16,S,S] or nullptr, out bf16 NHWC [nb,S,S,16]. w_host / bias_host are HOST pointers
// ([16][Cx][3][3], [16]): the parameters are passed to the kernel by value (constant bank).
int launch_conv0(const float* x, const float* w_host, const float* bias_host, const float* cond, void* out, int nb,
                 int nx, int ncond, int Cx, int S, cudaStream_t s);

// Sinusoidal time encoding (+ optional label embedding add): out[r, 0:100] for rows r < R.
//   t: fp32 [R]; label: int32 [R] (-1 = none) or nullptr; emb: [num_classes,100] or nullptr.
int launch_pos_encoding(const float* t, const int* label, const float* emb, const float* inv_freq, float* out, int R,
                        cudaStream_t s);

// C[M,N] (ldc) = act(A[M,K] (lda) * B[N,K]^T (ldb) + bias[N]); act: 0 none, 1 SiLU, 2 ReLU. fp32, CUDA cores.
int launch_sgemm_nt(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M,
                    int N, int K, int act, cudaStream_t s);

// x <- c1 * (x - c2 * eps) + c3 * z, coefficient row picked by *step; reference rounding order (no FMA).
//   cfg != 0: eps = lerp(eps_u, eps_c, cfg_scale) with eps_c = eps[0:numel], eps_u = eps[numel:2 numel] (ATen lerp form).
int launch_ddpm_update(float* x, const float* eps, const float* noise, const float* coef, int* step, size_t numel,
                       int cfg, float cfg_scale, int* trow, int n_rows, int row_dec, int* counter, cudaStream_t s);

// out = sa[b] * x + sb[b] * eps per sample b (forward noising), fp32, separately rounded multiply / add.
int launch_noise_images(const float* x, const float* eps, const float* sa, const float* sb, float* out, int n,
                        size_t per_sample, cudaStream_t s);

// trow[b] -= dec for b < n; *step -= 1   (one thread block; end-of-step bookkeeping kept on the device)
int launch_advance(int* trow, int n, int dec, int* step, cudaStream_t s);
int launch_set_rows(int* trow, const int* base_host_like_dev, int n, int* step, int step_value, int mul,
                    cudaStream_t s);

// Gaussian-weighted overlap blend, gather form, patch-order accumulation (bit-exact with the sequential scatter).
//   patches fp32 [ny*nx, C, P, P]; ys[ny], xs[nx] = unique window starts in the output; weight [P,P];
//   out fp32 [C, H, W] = clamp(sum_p patch*w / sum_p w, 0, 1); wsum_out optional [H, W].
int launch_blend_gather(const float* patches, const int* ys, int ny, const int* xs, int nx, const float* weight,
                        float* out, float* wsum_out, int C, int H, int W, int P, int do_clamp, cudaStream_t s);
// Vector form of the same (window starts, P and W multiples of 4): 128-bit accesses, four pixels per thread, window
// ranges per output row / pixel quad from host-built tables instead of a scan.
int launch_blend_gather4(const float* patches, const int* ys, int ny, const int* xs, int nx, const int2* row_rng,
                         const int2* col_rng, const float* weight, float* out, float* wsum_out, int C, int H, int W,
                         int P, int do_clamp, cudaStream_t s);
// Scatter form used for arbitrary window lists: one launch per patch keeps the reference's summation order.
int launch_blend_accumulate(const float* patch, const float* weight, float* acc, float* wsum, int C, int H, int W,
                            int P, int y0, int x0, cudaStream_t s);
int launch_blend_finalize(float* acc, const float* wsum, int C, int H, int W, int do_clamp, int* zero_flag,
                          cudaStream_t s);

// Reads `bytes` of a buffer (larger than L2) so that the next kernel starts on a cold, clean L2 (timing only).
int launch_l2_flush_read(const void* buf, size_t bytes, cudaStream_t s);

// fp32 NCHW <-> bf16 NHWC converters (debug / layer-level entry points)
int launch_nchw_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, cudaStream_t s);
int launch_nhwc_bf16_to_nchw(const void* in, float* out, int B, int C, int H, int W, cudaStream_t s);

}  // namespace drs
