/*
 * drs_b200 — C ABI of the B200-native reverse-diffusion sampling path of AdrianoEttari/DiffusionRemoteSensing.
 *
 * The reference has no FFI layer: its boundary is the Python class surface. This header is the
 * boundary a binding would use instead of the PyTorch operator calls the reference makes on the hot path; every
 * entry point names the reference lines it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every function returns 0 on success, a negative DRS_E_* code otherwise; drs_last_error() returns a
 *     thread-local human readable message for the last failure on the calling thread.
 *   - all `dev` pointers are device pointers on the model's device, owned by the caller; the library only owns what
 *     drs_model_create / drs_plan_create allocated. Work is enqueued asynchronously on `stream` (a cudaStream_t
 *     passed as void*); the library never synchronises unless stated.
 *   - image tensors at the boundary are fp32 NCHW (what the Python surface exchanges); internal activations are
 *     bf16 NHWC, accumulation is fp32.
 *   - one DrsModel per device; calls on one handle must be serialised by the caller.
 */
#ifndef DRS_B200_H
#define DRS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRS_OK 0
#define DRS_E_INVALID (-1)   /* bad argument / unsupported shape */
#define DRS_E_MISSING (-2)   /* a required state_dict entry is absent or has the wrong size */
#define DRS_E_CUDA (-3)      /* CUDA runtime / driver error (message has the cudaError string) */
#define DRS_E_PIPELINE (-4)  /* a kernel reported an internal pipeline timeout */
#define DRS_E_STATE (-5)     /* call sequence error (e.g. sampler not prepared) */

typedef struct DrsModel DrsModel;
typedef struct DrsPlan DrsPlan;

/* Model families: UNet_model_superres.py:266, UNet_model_SAR_TO_NDVI.py:263,
 * generate_new_imgs/UNet_model_generation.py:226 */
enum { DRS_MODEL_SUPERRES = 0, DRS_MODEL_SAR_TO_NDVI = 1, DRS_MODEL_GENERATION = 2 };

typedef struct DrsModelDesc {
  int kind;          /* DRS_MODEL_* */
  int x_channels;    /* channels of the denoised image x (3 / 1 / 3) */
  int cond_channels; /* channels of the conditioning image (3 LR / 2 SAR / 0) */
  int out_channels;  /* channels of the predicted noise */
  int num_classes;   /* generation only: rows of label_emb.weight, 0 = none */
} DrsModelDesc;

/* One state_dict entry: host fp32, contiguous, PyTorch layout. */
typedef struct DrsTensor {
  const char* name;
  const float* data;
  int64_t numel;
} DrsTensor;

const char* drs_last_error(void);
int drs_version(void);

/* Packs a state_dict (same keys as the reference nn.Module) into bf16 UMMA-layout weight tiles, fp32 eval-BatchNorm
 * scale/shift vectors and K-block programs on `device`. Replaces: module construction + model.to(device)
 * (train_diffusion_superres.py:107) as far as the sampling path is concerned. */
int drs_model_create(const DrsModelDesc* desc, const DrsTensor* tensors, int n_tensors, int device, DrsModel** out);
void drs_model_destroy(DrsModel* m);

/* A plan fixes the forward batch `nb`, the image side `S` (multiple of 8), how many distinct x states feed it
 * (`nx`, nb % nx == 0; nb = 2*nx batches the conditional and unconditional passes of classifier-free guidance) and
 * how many conditioning images (`ncond`: 1 = broadcast like Diffusion.sample, or nb). It owns the activation
 * workspace, the TMA descriptors and the launch list. */
int drs_plan_create(DrsModel* m, int nb, int nx, int ncond, int S, int magnification, DrsPlan** out);
void drs_plan_destroy(DrsPlan* p);
size_t drs_plan_workspace_bytes(const DrsPlan* p);

/* Time-invariant condition branch: RRDB encoder, bicubic x magnification, 3x3 conv to 16 channels
 * (UNet_model_superres.py:345-353; UNet_model_SAR_TO_NDVI.py:341-343). cond_dev: fp32 [ncond, Cc, S/mag, S/mag].
 * Result stays inside the plan. No-op for the generation family. */
int drs_cond_encode(DrsPlan* p, const float* cond_dev, void* stream);

/* Time / class embedding rows for arbitrary per-sample timesteps (the UNet.forward drop-in):
 * pos_encoding + label_emb + the seven time MLPs (UNet_model_superres.py:328-339,143-151,161,199).
 * t_dev: fp32 [nb]; label_dev: int32 [nb] (-1 = no label) or NULL. */
int drs_time_embed(DrsPlan* p, const float* t_dev, const int32_t* label_dev, void* stream);

/* One UNet evaluation (UNet_model_superres.py:337-379 and the SAR / generation variants):
 * x_dev fp32 [nx, Cx, S, S] -> eps_dev fp32 [nb, Cout, S, S]. Uses the rows of the last drs_time_embed (or the
 * sampler's table) and the last drs_cond_encode. */
int drs_unet_forward(DrsPlan* p, const float* x_dev, float* eps_dev, void* stream);

/* Reads back the device error word the tensor-core kernels set when an mbarrier wait times out; synchronises the
 * stream. Returns DRS_E_PIPELINE if any launch since the last check reported a timeout. */
int drs_plan_check(DrsPlan* p, void* stream);

/* Sampler (Diffusion.sample, train_diffusion_superres.py:224-255 and the two siblings).
 * prepare: uploads per-step coefficients c1 = 1/sqrt(alpha), c2 = (1-alpha)/sqrt(1-alpha_hat), c3 = sqrt(beta)
 *          (host fp32 [noise_steps], computed by the caller with the reference's torch ops so they stay
 *          bit-identical) and precomputes the time-embedding table for every step and every distinct label.
 *          labels_host: int32 [nb] or NULL (-1 = unconditional row). cfg_scale is used when nb == 2*nx.
 * begin:   binds the caller-owned state / noise / eps buffers and positions the sampler at step `start_step`.
 * step:    one reverse step at the current step index: UNet forward, (CFG lerp,) posterior update, step -= 1.
 *          noise buffer content is consumed as z (caller fills it before every step; zeros at the last step).
 *          use_graph != 0 replays a CUDA graph captured on first use. */
int drs_sampler_prepare(DrsPlan* p, int noise_steps, const float* c1, const float* c2, const float* c3,
                        const int32_t* labels_host, float cfg_scale, void* stream);
int drs_sampler_begin(DrsPlan* p, float* x_dev, float* noise_dev, float* eps_dev, int start_step, void* stream);
int drs_sampler_step(DrsPlan* p, int use_graph, void* stream);
/* Number of kernel launches one sampler step issues (bench bookkeeping). */
int drs_sampler_launches_per_step(const DrsPlan* p);

/* Per-launch accounting of one UNet evaluation (bench.py / profiles). Launch 0 is the CUDA-core conv0 kernel, the
 * others are tensor-core launches in execution order. flops / bytes are ALGORITHMIC: 2 * MACs of the convolutions the
 * launch computes, and one read of every input + one write of the output + the weights.
 * drs_plan_profile runs `iters` evaluations with a CUDA event between consecutive launches on `stream` and writes
 * the mean duration (ms) of every launch to ms_out[drs_plan_launch_count()]; it synchronises the stream. */
int drs_plan_launch_count(const DrsPlan* p);
int drs_plan_launch_info(const DrsPlan* p, int index, char* name, int name_capacity, double* flops, double* bytes,
                         int* ctas, int* smem_bytes);
int drs_plan_profile(DrsPlan* p, const float* x_dev, float* eps_dev, int iters, float* ms_out, void* stream);
/* Times the forward exactly as the sampler enqueues it (gate branch on the side stream, no event between the
 * tensor-core launches): ms_out2[0] = conv0, ms_out2[1] = first tensor-core launch to completion of the last one,
 * averaged over `iters` evaluations. Synchronises `stream`. */
int drs_plan_time_forward(DrsPlan* p, const float* x_dev, float* eps_dev, int iters, float* ms_out2, void* stream);

/* Event-timed durations (ms) of the two CUDA-core kernels of a reverse step on a begun sampler, each launch preceded
 * by a read of the caller's flush buffer (larger than the 126 MB L2, leaves it cold and clean; NULL = no flush): ms_out2[0] = conv0
 * (UNet_model_superres.py:342,355), ms_out2[1] = the posterior update incl. its bookkeeping tail
 * (train_diffusion_superres.py:240-249). bench.py's HBM roofline entries. Needs more than `iters` steps left; the step
 * index is restored. Synchronises the stream. */
int drs_sampler_time_hbm_kernels(DrsPlan* p, void* flush_dev, size_t flush_bytes, int iters, float* ms_out2,
                                 void* stream);

/* Stand-alone posterior update (train_diffusion_superres.py:240-249), scalars given directly. */
int drs_ddpm_update(float* x_dev, const float* eps_dev, const float* noise_dev_or_null, float c1, float c2, float c3,
                    size_t numel, void* stream);

/* Forward noising, train_diffusion_superres.py:183-190: out[b] = sqrt_ah[b] * x[b] + sqrt_1m_ah[b] * eps[b] for n samples
 * of per_sample fp32 elements (multiple of 4); the per-sample factors are sqrt(alpha_hat[t_b]) and
 * sqrt(1 - alpha_hat[t_b]) gathered by the caller. Multiply and add are rounded separately like the reference. */
int drs_noise_images(const float* x_dev, const float* eps_dev, const float* sqrt_ah_dev, const float* sqrt_1m_ah_dev,
                     float* out_dev, int n, size_t per_sample, void* stream);

/* Aggregation_Sampling.py:91-110: Gaussian-weighted overlap blend of SR patches, summed in patch order.
 * patches_dev fp32 [n_patches, C, P, P]; coords4_host int32 [n_patches][4] = (y0, y1, x0, x1) in the output;
 * weight_dev fp32 [P, P]; out_dev fp32 [C, H, W]; wsum_dev fp32 [H, W] (scratch, also returned).
 * Returns DRS_E_INVALID if some output pixel is covered by no patch (the reference asserts). */
int drs_blend(const float* patches_dev, const int32_t* coords4_host, int n_patches, const float* weight_dev,
              float* out_dev, float* wsum_dev, int C, int H, int W, int P, int do_clamp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRS_B200_H */
