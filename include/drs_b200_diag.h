/*
 * drs_b200_diag -- diagnostics that are NOT part of the drop-in boundary (include/drs_b200.h is).
 *
 *   1. libdrs_b200_diag.so (csrc/diag_kernels.cu): tcgen05.mma issue-rate micro-benchmarks. A separate shared
 *      library; the product library neither contains nor calls them (errors: see drs_diag_last_error below).
 *   2. Test / instrumentation hooks exported by libdrs_b200.so itself: the layer-level convolution entry point and
 *      activation taps the parity tests use, and the read-out of the clock stamps the production kernels record when
 *      the environment variable DRS_V2_TIMELINE is set (inert otherwise). Errors go through the product library's
 *      last-error string.
 */
#ifndef DRS_B200_DIAG_H
#define DRS_B200_DIAG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef DRS_B200_H
typedef struct DrsPlan DrsPlan;
#endif

/* ---- libdrs_b200_diag.so ------------------------------------------------------------------------------------ */
const char* drs_diag_last_error(void);

/* Debug micro-benchmark: one elected thread per CTA issues `iters` (x4 if unroll4) back-to-back tcgen05.mma
 * M=128 x n x K=16 (bf16) on shared-memory operands, `ctas_per_sm` CTAs per SM. out_host[0] = SM cycles until the
 * last instruction was issued, out_host[1] = cycles until all completed (CTA 0). Synchronises the device. */
int drs_debug_mma_rate(int n, int iters, int unroll4, int ctas_per_sm, long long* out_host);

/* Debug micro-benchmark, second form: `issuers` (1..4) warps of one CTA per SM each issue `iters` K-blocks of `nk`
 * MMAs (M=128 x n x K=16) with the given swizzle `layout` code (2 = 128 B, 4 = 64 B, 6 = 32 B rows) and A group stride
 * `sbo16` (16-byte units). mode 0: descriptors in registers; 1: 32-byte table record per K-block; 2: packed 64-bit
 * record. out_host as above. */
int drs_debug_mma_rate2(int n, int nk, int layout, int sbo16, int issuers, int iters, int mode, long long* out_host);

/* ---- hooks inside libdrs_b200.so ------------------------------------------------------------------------------ */
/* Layer-level entry point (tests / INTEGRATION): y = act((conv(x) + bias) * scale + shift) on the tensor-core path.
 * x_dev fp32 NCHW [B,Cin,H,W]; w_host fp32 PyTorch layout; kind: 0 = 3x3 s1 p1, 1 = 3x3 s2 p1, 2 = 1x1,
 * 3 = 2x2 s2 p0, 4 = ConvTranspose2d(3, s2, p1, op1) (weight [Cin,Cout,3,3]). y_dev fp32 NCHW.
 * scale_host / shift_host may be NULL. Synchronises the stream. */
int drs_debug_conv2d(const float* x_dev, const float* w_host, const float* bias_host, const float* scale_host,
                     const float* shift_host, float* y_dev, int B, int Cin, int Cout, int H, int W, int kind, int relu,
                     int device, void* stream);

/* Intermediate activations of the last forward, converted to fp32 NCHW (tests only). Returns numel written or <0.
 * name: "h0","b0.h","b0.out","d0","b1.out","d1","b2.out","d2","bn.out","g0","psi0","att0","uc0","ut0","x0",... */
int64_t drs_debug_fetch(DrsPlan* p, const char* name, float* out_dev, int64_t capacity, void* stream);

/* Timing aid: reads `bytes` (more than the 126 MB L2) of a device buffer on `stream`, so the next kernel starts on a
 * cold L2 that holds no dirty lines (a memset flush leaves write-backs that compete with the kernel being timed). */
int drs_debug_l2_flush(const void* buf_dev, size_t bytes, void* stream);

/* Debug: SM-clock stamps recorded by CTA 0 of the last second-generation convolution launch when the environment
 * variable DRS_V2_TIMELINE is set (8 values per pixel tile: producer start / last issue, MMA after TMEM-empty wait /
 * after first A-full wait / after last issue, epilogue after TMEM-full wait / done). n <= 512. */
int drs_debug_timeline(long long* out_host, int n);
/* Debug: with DRS_V2_TIMELINE bit 2 set every second-generation launch records [first CTA entry, last CTA exit] in
 * globaltimer nanoseconds under its launch index. reset != 0 re-arms the table, else it is copied to out_host[128]. */
int drs_debug_spans(unsigned long long* out_host, int reset);

#ifdef __cplusplus
}
#endif
#endif /* DRS_B200_DIAG_H */
