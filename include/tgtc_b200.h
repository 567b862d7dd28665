/*
 * tgtc_b200.h -- C ABI of the B200-native NeRF ray-render hot path of TGTC-Style.
 *
 * One shared library (libtgtc_b200.so, built by nvcc for sm_100a), plain
 * pointers and sizes, no torch types.  Every entry point that launches work
 * takes a CUDA stream (passed as void*), is asynchronous and stream-ordered,
 * and returns an int status; tgtc_last_error() gives the thread-local message
 * of the last non-zero status.  The library allocates nothing the caller sees:
 * inputs, outputs and workspaces are caller-owned device memory (the *_host
 * entry points own a private staging arena inside the context).
 *
 * The reference (/root/reference, pure Python/PyTorch) has no FFI; the seam this
 * ABI replaces is the set of four Python callables that the reference injects
 * into its render loops plus the ray generation that
 * feeds them.  Each entry point cites the reference interface it replaces.
 * The Python-side binding a maintainer adds is shown in INTEGRATION.md.
 */
#ifndef TGTC_B200_H
#define TGTC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGTC_ABI_VERSION 2

typedef struct tgtc_ctx tgtc_ctx;
typedef void* tgtc_stream; /* cudaStream_t */

enum {
  TGTC_OK = 0,
  TGTC_ERR_ARG = 1,         /* null / misaligned pointer, bad size, bad enum */
  TGTC_ERR_CUDA = 2,        /* a CUDA runtime call or launch failed */
  TGTC_ERR_STATE = 3,       /* weights not set, workspace too small, ... */
  TGTC_ERR_UNSUPPORTED = 4  /* shape outside what the kernel handles */
};

enum { TGTC_NET_COARSE = 0, TGTC_NET_FINE = 1 };

/* arithmetic of the MLP (K3+K4).  FP32: CUDA-core FFMA, reference-grade
 * (<=1e-3 end to end).  BF16: bf16 operands on tcgen05 tensor cores, fp32
 * accumulation in TMEM, fp32 sigma/rgb heads.  F16: the same kernel with fp16
 * operands (11-bit significand: 8x smaller operand rounding error than bf16 at
 * the same tensor-core rate; operand values saturate at +-65504) -- the mode
 * that meets the 1e-2 teacher-forced bound on every non-knife-edge ray
 * (DESIGN.md section 2); inference entry points only, training is bf16. */
enum { TGTC_MLP_FP32 = 0, TGTC_MLP_BF16 = 1, TGTC_MLP_F16 = 2 };

#define TGTC_NUM_PARAMS 24 /* 12 layers x (weight, bias) */

const char* tgtc_last_error(void);
int tgtc_abi_version(void);

/* One context per device (not shared between threads without a lock).
 * Owns the packed weights of both nets and small tables. */
int tgtc_create(int device, tgtc_ctx** out);
int tgtc_destroy(tgtc_ctx* ctx);

/* Network parameters.  Replaces: nn.Module state of models.StyleNerf
 * (models.py:182-223; MLP_style.__init__ models.py:63-93).
 * params[2*i], params[2*i+1] = weight [out,in] row-major fp32, bias [out] of
 * layer i in the order base_layers[0..7], sigma_layer, base_remap_layer,
 * rgb_layers[0], rgb_layers[1] (shapes: 256x63, 5x 256x256 with layer 5
 * 256x319, 1x256, 256x256, 128x283, 3x128).  Device pointers; packed on
 * `stream` into the fp32-transposed and bf16-swizzled images the kernels read.
 * Call again after every optimizer step / load_state_dict. */
int tgtc_set_weights(tgtc_ctx* ctx, int net, const float* const* params, tgtc_stream stream);

/* K1 -- ray generation + NDC warp.  Replaces dataset.get_rays_np
 * (dataset.py:33-42) and dataset.ndc_rays_np (dataset.py:44-61) followed by
 * the cast to fp32 rays.  fp64 arithmetic in the reference's operation order
 * (bit-exact), K row-major 3x3, c2w row-major 3x4 (host pointers, copied by
 * value).  Writes rays for pixels [pix_begin, pix_begin+n) of the H x W frame
 * (row-major, pixel p = row*W+col) to rays_o / rays_d ([n,3] fp32, device). */
int tgtc_raygen(tgtc_ctx* ctx, int H, int W, const double* K, const double* c2w, int ndc, double ndc_near,
                int pixel_alignment, int64_t pix_begin, int64_t n, float* rays_o, float* rays_d,
                tgtc_stream stream);

/* K2 -- stratified sample positions.  Replaces utils.sampling_pts_uniform
 * (utils.py:509-531).  harmony != 0: ts = 1/(1/near*(1-ts) + 1/far*ts) (utils.py:516; near, far non-zero) instead of the
 * linear row.  rand: NULL for perturb=False, else [n,S] uniforms replaying utils.py:518-524.  pts [n,S,3] may be NULL. ts [n,S]. */
int tgtc_sample_uniform(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, int64_t n, int n_samples,
                        double near, double far, int harmony, const float* rand, float* pts, float* ts, tgtc_stream stream);

/* K3+K4 -- positional encoding + MLP on explicit points.  Replaces
 * models.StyleNerf.forward (models.py:216-223) = Embedder x2 (models.py:46-60)
 * + MLP_style.forward (models.py:95-117), including utils.batchify
 * (utils.py:435-456; chunking is internal).  pts [n_rays*S,3]; dirs is [n_rays,3]
 * when dirs_per_ray (the reference passes a stride-0 expand, rendering.py:30)
 * else [n_rays*S,3].  Outputs: rgb [M,3], sigma [M]; optional (NULL to skip)
 * base_remap [M,256], pts_embed [M,63], dirs_embed [M,27] -- the other keys of
 * the reference's returned dict.  BF16 mode needs dirs_per_ray and
 * S in {32,64,128} or a multiple of 128, and no optional outputs. */
int tgtc_nerf_forward(tgtc_ctx* ctx, int net, int mode, const float* pts, const float* dirs, int dirs_per_ray,
                      int64_t n_rays, int S, float* rgb, float* sigma, float* base_remap, float* pts_embed,
                      float* dirs_embed, tgtc_stream stream);

/* K2+K3+K4 fused -- sample points are formed in-kernel from rays and t values
 * (pts = o + t*d never touches HBM).  ts: [n_rays,S] or NULL for the uniform
 * coarse positions linspace(0,1,S)*(far-near)+near.  Output rgbsigma
 * [n_rays,S,4] = (r,g,b,sigma). */
int tgtc_nerf_forward_rays(tgtc_ctx* ctx, int net, int mode, const float* rays_o, const float* rays_d,
                           const float* ts, int64_t n_rays, int S, double near, double far, float* rgbsigma,
                           tgtc_stream stream);

/* K5 -- alpha compositing.  Replaces utils.alpha_composition (utils.py:354-386).
 * Either (rgb [n,S,3], sigma [n,S]) or rgbsigma [n,S,4] (the other NULL).
 * ts [n,S] with ts_ray_stride = S, or one shared row with ts_ray_stride = 0
 * (the reference's expanded view).  noise: NULL or [n,S] = randn*sigma_noise_std.
 * Outputs (each may be NULL): rgb_out [n,3], depth_out [n], acc_out [n]
 * (the reference computes and drops it, utils.py:382), weights_out [n,S]. */
int tgtc_composite(tgtc_ctx* ctx, const float* rgb, const float* sigma, const float* rgbsigma, const float* ts,
                   int64_t ts_ray_stride, const float* noise, int white_bkgd, int64_t n, int S, float* rgb_out,
                   float* depth_out, float* acc_out, float* weights_out, tgtc_stream stream);

/* K5 backward -- gradient of alpha_composition for the training step (the autograd graph of
 * train_tgtcs.py:236-255 restricted to utils.py:354-386).  rgbsigma [n,S,4] and ts as in the forward; g_rgb [n,3] =
 * dL/d(rgb map); g_depth [n], g_acc [n] may be NULL.  Writes d_rgbsigma [n,S,4] = dL/d(r,g,b,sigma) per sample.
 * No gradient flows to ts (utils.py:576-579).  S <= 256. */
int tgtc_composite_backward(tgtc_ctx* ctx, const float* rgbsigma, const float* ts, int64_t ts_ray_stride, const float* noise,
                            int white_bkgd, int64_t n, int S, const float* g_rgb, const float* g_depth, const float* g_acc,
                            float* d_rgbsigma, tgtc_stream stream);

/* K6+K7 -- hierarchical inverse-CDF resampling + sorted union.  Replaces
 * utils.sampling_pts_fine_torch (utils.py:573-580) and utils.sample_pdf
 * (utils.py:583-609, det=True).  Bin selection is bit-exact with CPU torch
 * (sum in ATen's vector order, fp64 cdf).  ts [n,S] (ts_ray_stride S or 0),
 * weights [n,S].  Outputs: ts_out [n,S+n_fine]; optional pts_out
 * [n,S+n_fine,3] (needs rays_o/rays_d), inds_out [n,n_fine] int64 (the
 * searchsorted result), samples_out [n,n_fine] (unsorted new samples). */
int tgtc_sample_fine(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, const float* ts,
                     int64_t ts_ray_stride, const float* weights, int64_t n, int S, int n_fine, float* pts_out,
                     float* ts_out, int64_t* inds_out, float* samples_out, tgtc_stream stream);

/* The fused operator (north-star surface):
 *   render(rays_o, rays_d, near, far, chunk) -> {rgb, depth, acc, weights}
 * = the loop body of rendering.py:27-51 (cal_geometry) for one batch of rays:
 * coarse MLP -> compositing -> resampling -> fine MLP -> compositing
 * (tensor-core modes at 64 / 128 samples: three launches -- the compositing is
 * the last epilogue of each MLP kernel; bit-identical to tgtc_composite).
 * Any output pointer may be NULL.  All device memory. */
typedef struct tgtc_render_out {
  float* rgb;            /* [n,3]   fine */
  float* depth;          /* [n]     fine, NDC t units (t_exp) */
  float* acc;            /* [n]     fine */
  float* weights;        /* [n,S+F] fine */
  float* rgb_coarse;     /* [n,3] */
  float* depth_coarse;   /* [n] */
  float* acc_coarse;     /* [n] */
  float* weights_coarse; /* [n,S] */
  float* ts_fine;        /* [n,S+F] */
} tgtc_render_out;

/* bytes of device workspace tgtc_render needs for a call with these sizes
 * (chunk <= 0 means "all rays in one pass").  The first form is enough for every
 * mode; the _mode form returns what `mode` needs: in the tensor-core modes with
 * n_samples, n_samples+n_fine in {64,128} the compositing runs inside the MLP
 * kernel (a 128-sample tile is one fine ray / two coarse rays), the per-sample
 * (r,g,b,sigma) never reach HBM and the workspace shrinks from 48 B/sample to
 * ~4 B/sample (coarse weights + ts_fine). */
size_t tgtc_render_workspace_bytes(int64_t n_rays, int n_samples, int n_fine, int64_t chunk);
size_t tgtc_render_workspace_bytes_mode(int mode, int64_t n_rays, int n_samples, int n_fine, int64_t chunk);

int tgtc_render(tgtc_ctx* ctx, int mode, const float* rays_o, const float* rays_d, int64_t n_rays, double near,
                double far, int n_samples, int n_fine, int64_t chunk, int white_bkgd, const tgtc_render_out* out,
                void* workspace, size_t workspace_bytes, tgtc_stream stream);

/* Same operator with HOST buffers (pinned or pageable): copies rays H2D,
 * renders, copies the requested outputs D2H and synchronises `stream` before
 * returning.  Staging memory lives in the context.  This is the end-to-end
 * entry bench.py times as `e2e`. */
int tgtc_render_host(tgtc_ctx* ctx, int mode, const float* rays_o, const float* rays_d, int64_t n_rays, double near,
                     double far, int n_samples, int n_fine, int64_t chunk, int white_bkgd,
                     const tgtc_render_out* host_out, tgtc_stream stream);

/* Frame operator: K1 fused in front of tgtc_render for pixels
 * [pix_begin, pix_begin+n) of an H x W pinhole frame (replaces the
 * RaySampler precompute dataset.py:105-118 + the cal_geometry batch loop;
 * pixel_alignment as in tgtc_raygen / get_rays_np, dataset.py:33-36).
 * Device outputs; workspace as for tgtc_render plus 24*n bytes for the rays. */
size_t tgtc_render_frame_workspace_bytes(int64_t n_rays, int n_samples, int n_fine, int64_t chunk);
size_t tgtc_render_frame_workspace_bytes_mode(int mode, int64_t n_rays, int n_samples, int n_fine, int64_t chunk);
int tgtc_render_frame(tgtc_ctx* ctx, int mode, int H, int W, const double* K, const double* c2w, int ndc,
                      double ndc_near, int pixel_alignment, int64_t pix_begin, int64_t n, double near, double far, int n_samples,
                      int n_fine, int64_t chunk, int white_bkgd, const tgtc_render_out* out, void* workspace,
                      size_t workspace_bytes, tgtc_stream stream);

/* Training step (replaces the autograd graph of Origin_train, train_tgtcs.py:228-255, for one batch
 * of rays with perturb=0 / sigma_noise_std=0): forward of both nets with the activations stashed, the two MSE losses
 * against rgb_gt [n,3], and the full backward.  grads: flat fp32 buffer of 2 * tgtc_num_params() values -- coarse net
 * then fine net, each in the order of tgtc_set_weights (weight [out,in] row-major, bias) -- so one all-reduce covers
 * both nets; accumulate != 0 adds to it (ray chunks of one step).  The loss is
 * mean((rgb_coarse-gt)^2) + mean((rgb_fine-gt)^2) over n_rays_total*3 values (pass the whole step's ray count when
 * the step is split into chunks / ranks); loss_sums (device float[2], may be NULL) is incremented by the two
 * un-normalised squared-error sums.  rgb_coarse / rgb_fine [n,3] may be NULL.  bf16 tcgen05 path;
 * n_samples = n_fine = 64.  No gradient flows through the resampling (utils.py:576-579).
 * Stochastic options of the reference replay caller-drawn tensors (so a torch generator stream can be reproduced):
 * rand [n,64] uniforms = perturb=True (utils.py:518-524), noise_coarse [n,64] / noise_fine [n,128] =
 * randn * sigma_noise_std (utils.py:372-374); each may be NULL (= off). */
/* Overlap hook for data-parallel training: `cuda_event` (a caller-owned cudaEvent_t, NULL to clear) is recorded
 * on the step's stream by every following tgtc_train_step* call at the point where the COARSE net's half of `grads` is final,
 * i.e. before the fine net's forward / backward is enqueued -- the caller's gradient all-reduce of that half can then run on
 * another stream underneath the fine net's work (no gradient links the two nets: utils.py:576-579). */
int tgtc_train_set_coarse_event(tgtc_ctx* ctx, void* cuda_event);
size_t tgtc_train_workspace_bytes(tgtc_ctx* ctx, int64_t n_rays, int n_samples, int n_fine);
int64_t tgtc_num_params(void); /* 595 844 per net */
int tgtc_train_step(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, const float* rgb_gt, int64_t n_rays,
                    int64_t n_rays_total, double near, double far, int n_samples, int n_fine, const float* rand,
                    const float* noise_coarse, const float* noise_fine, float* grads, int accumulate, float* loss_sums,
                    float* rgb_coarse, float* rgb_fine, void* workspace, size_t workspace_bytes, tgtc_stream stream);

/* Stage-level training entries: the forward of ONE net with its activation stash, and the matching backward.  They replace
 * the autograd graph that PyTorch records for models.StyleNerf.forward (models.py:216-223, through utils.batchify,
 * utils.py:435-456) inside Origin_train (train_tgtcs.py:234, :246 -> loss.backward() :254), so that loop runs unchanged with the
 * model_forward drop-in (tgtc-style_b200/shims.py wraps these two calls in a torch.autograd.Function).
 *   forward : pts [n_rays*S,3], dirs [n_rays,3] (the reference passes a stride-0 expand of rays_d) -> rgbsigma [n_rays*S,4];
 *             the activation stash of the pass stays in `stash` (tgtc_nerf_stash_bytes, 1024-byte aligned, ~5.2 KB/sample)
 *   backward: dL/d(r,g,b,sigma) per sample [n_rays*S,4] + the forward's rgbsigma and stash -> grads, a flat fp32 buffer of
 *             tgtc_num_params() values in tgtc_set_weights order (accumulate != 0 adds); `scratch` (tgtc_nerf_backward_scratch_bytes)
 *             is free again when the call's work has run.  bf16 tcgen05 kernels; S in {64, 128}.  No gradient w.r.t. pts / dirs
 *             (they are data in the reference's graph too). */
size_t tgtc_nerf_stash_bytes(int64_t n_rays, int S);
size_t tgtc_nerf_backward_scratch_bytes(tgtc_ctx* ctx, int64_t n_rays, int S);
int tgtc_nerf_forward_stash(tgtc_ctx* ctx, int net, const float* pts, const float* dirs, int64_t n_rays, int S, float* rgbsigma,
                            void* stash, size_t stash_bytes, tgtc_stream stream);
int tgtc_nerf_backward(tgtc_ctx* ctx, int net, const float* dirs, int64_t n_rays, int S, const float* rgbsigma, const float* d_rgbsigma,
                       float* grads, int accumulate, void* stash, size_t stash_bytes, void* scratch, size_t scratch_bytes,
                       tgtc_stream stream);

/* The reference's optimizer step (torch.optim.Adam, betas (0.9, 0.999), eps 1e-8; train_tgtcs.py:39, :255) on flat fp32
 * device buffers of n values -- e.g. the 2 * tgtc_num_params() masters laid out like the gradient buffer of
 * tgtc_train_step.  step is the 1-based step count (bias corrections); the lr schedule (train_tgtcs.py:272-276) is the
 * caller's.  One kernel, 28 B per parameter. */
int tgtc_adam_step(tgtc_ctx* ctx, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                   double beta1, double beta2, double eps, int64_t step, tgtc_stream stream);

/* Per-ray style head.  Replaces models.StyleMLP_before_concat (models.py:120-147) and
 * models.StyleMLP_Wild_multilayers (models.py:149-180) as called by render_style / render_train_style
 * (rendering.py:118-178, :280-327).  params: 26 device pointers = module 1 layers.{0..4} (weight [out,in] row-major, bias),
 * then module 2 layers.{0..7}; shapes 256x95, 3x 256x288, 256x351 and 256x607, 3x 256x288, 256x351, 2x 256x288, 3x288
 * (style_D = 8, vae_latent = 32).  Everything needed later is packed / copied here; the tensors may be freed afterwards. */
int tgtc_set_style_weights(tgtc_ctx* ctx, const float* const* params, tgtc_stream stream);

/* The loop body of render_style for one batch of rays with perturb=False: NeRF trunk (base_remap, sigma) -> style
 * module 1 (concat_features) -> style module 2 (stylised rgb) -> compositing with the NeRF sigma -> resampling -> the
 * same on the fine net.  latent1 / latent2: device pointers to the 32 latent values of module 1 / module 2 for this
 * batch (one (style, frame) per call: latent1 = latents_model_1(...) row, rendering.py:125; latent2 = its mean over the
 * latent dim broadcast to 32, rendering.py:126,:139).  Outputs as tgtc_render.  tcgen05 path (mode = TGTC_MLP_BF16 or
 * TGTC_MLP_F16: operand format of all three kernels and of the feature tiles between them); 64 + 64 samples;
 * chunk <= 0 means passes of 32768 rays (128 KB of feature tiles per 128 samples live in the workspace).
 * A call of several passes (n_rays > chunk) alternates them between `stream` and a stream the context owns, forked from and
 * joined back to `stream` by events, each with its own half of the workspace: the call stays ordered on `stream` for the caller,
 * the results are bit-identical, and one pass's kernel prologues / tails overlap the other's kernels (measured: 4096-ray passes
 * over a 1008x756 frame 356.6 -> 345.0 ms).  With tgtc_profile_enable on, everything stays on `stream`. */
size_t tgtc_render_style_workspace_bytes(int64_t n_rays, int n_samples, int n_fine, int64_t chunk);
int tgtc_render_style(tgtc_ctx* ctx, int mode, const float* rays_o, const float* rays_d, int64_t n_rays, double near, double far,
                      int n_samples, int n_fine, int64_t chunk, const float* latent1, const float* latent2,
                      const tgtc_render_out* out, void* workspace, size_t workspace_bytes, tgtc_stream stream);

/* tgtc_render_style with PER-RAY latents inside one call: latents [n_rays,32] = what latents_model_1(style_ids, frame_ids)
 * returns for the batch (rendering.py:125), so a batch may mix (style, frame) pairs; module 2 sees every ray's mean(latent)
 * broadcast to 32 dims (rendering.py:126, :139).  The latent columns of every layer are folded into per-ray effective biases
 * (13 x 256 fp32 per ray, in the workspace) that the chain kernel's epilogue stages per half-tile. */
size_t tgtc_render_style_rays_workspace_bytes(int64_t n_rays, int n_samples, int n_fine, int64_t chunk);
int tgtc_render_style_rays(tgtc_ctx* ctx, int mode, const float* rays_o, const float* rays_d, int64_t n_rays, double near, double far,
                           int n_samples, int n_fine, int64_t chunk, const float* latents, const tgtc_render_out* out, void* workspace,
                           size_t workspace_bytes, tgtc_stream stream);

/* Stage entries of the per-ray style head on EXPLICIT features -- what the reference's injected callables compute
 * (train_tgtcs.py:46, :53: batchify-wrapped modules; called at rendering.py:129 and :140):
 *   tgtc_style_concat_forward = StyleMLP_before_concat.forward(x, latent)        -> concat_features   (models.py:137-147)
 *   tgtc_style_forward        = StyleMLP_Wild_multilayers.forward(x, concated, latent) -> rgb         (models.py:165-180)
 * x [M,63] = embedded points (the `pts` key of model_forward's dict), concated [M,512] = cat(base_remap, concat_features),
 * latent: 32 device floats -- ONE latent row per call (the reference expands one (style, frame) latent over the batch's
 * samples; a batch mixing latents is split by the caller).  concat_features [M,256], rgb [M,3] fp32.  The features are
 * converted to the operand format of `mode` (BF16 / F16) tile images in `workspace` (tgtc_style_stage_workspace_bytes(M),
 * 1024-byte aligned) and run through the same tcgen05 chain kernel as tgtc_render_style. */
size_t tgtc_style_stage_workspace_bytes(int64_t n_samples_total);
int tgtc_style_concat_forward(tgtc_ctx* ctx, int mode, const float* x, const float* latent, int64_t n_samples_total,
                              float* concat_features, void* workspace, size_t workspace_bytes, tgtc_stream stream);
int tgtc_style_forward(tgtc_ctx* ctx, int mode, const float* x, const float* concated, const float* latent, int64_t n_samples_total,
                       float* rgb, void* workspace, size_t workspace_bytes, tgtc_stream stream);

/* Per-kernel device timing of the MLP launches (the dominant kernel), for the roofline line of bench.py:
 * while enabled, every MLP launch is bracketed by cudaEvents on its stream (no host sync).
 * tgtc_profile_read synchronises those events and returns, since the last read/enable: the number of MLP
 * launches, their summed device time in ms, and their summed algorithmic FLOPs (1 186 816 per sample). */
int tgtc_profile_enable(tgtc_ctx* ctx, int on);
int tgtc_profile_read(tgtc_ctx* ctx, int64_t* launches, double* ms, double* flops);
/* same per kernel kind, without resetting: 0 inference MLP forward, 1 training MLP forward (activation stash),
 * 2 activation-gradient kernel, 3 weight-gradient kernel (+ partial reduction) */
int tgtc_profile_read_kind(tgtc_ctx* ctx, int kind, int64_t* launches, double* ms, double* flops);

/* number of kernel launches issued by this context so far (bench.py's
 * gpu_launches claim is counted here, not estimated) */
int64_t tgtc_launch_count(const tgtc_ctx* ctx);

/* ---- in-kernel random streams for the training step -----------------------------------------------------
 * tgtc_train_step_seeded = tgtc_train_step with the reference's stochastic options drawn inside the kernels instead of read
 * from caller tensors: perturb != 0 -> stratified jitter (utils.py:518-524) generated in the sampling kernel;
 * sigma_noise_std > 0 -> randn * std (utils.py:372-374) generated in the compositing forward and regenerated in its backward.
 * Generator: Philox4x32-10, counter = element index, key = seed (csrc/philox.cuh; CPU restatement oracle/philox_oracle.py).
 * Element e of a stream is a pure function of (seed, stream, e): use a different seed for every call of a step (ray chunks)
 * and every step.  The reference's own streams (torch's global generator) are not reproduced, only the distributions.
 * tgtc_philox_fill writes the same streams as tensors: stream_id 0 = jitter uniforms [n_rays*n_samples] (normal = 0),
 * 1 / 2 = coarse / fine sigma noise [n_rays*S] (normal = 1, std) -- feeding them to tgtc_train_step(rand, noise_*) gives
 * bit-identical results to the seeded call. */
int tgtc_train_step_seeded(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, const float* rgb_gt, int64_t n_rays,
                           int64_t n_rays_total, double near, double far, int n_samples, int n_fine, unsigned long long seed,
                           int perturb, double sigma_noise_std, float* grads, int accumulate, float* loss_sums, float* rgb_coarse,
                           float* rgb_fine, void* workspace, size_t workspace_bytes, tgtc_stream stream);
int tgtc_philox_fill(tgtc_ctx* ctx, unsigned long long seed, int stream_id, int normal, double std, int64_t n, float* out,
                     tgtc_stream stream);

/* ---- Style_train: the second training phase (train_tgtcs.py:311-495) ------------------------------------
 * The reference's step: perturbed coarse samples (train_tgtcs.py:362) -> frozen NeRF net -> style module 1 with the ray's
 * latent (latents_model_1(style_id, frame_id), :409) -> style module 2 with mean(latent) (:410-421) -> compositing (:424) ->
 * resampling -> the same on the fine net (:456-479) -> losses on the two rgb maps (:425, :480-483) -> backward into the two
 * style modules (style_optimizer, :54) and the latents (latents_model_1.optimize, :495).
 * tgtc_style_train_forward runs everything up to the rgb maps and leaves the activation stash of both passes in the
 * workspace; the caller evaluates its loss on (rgb_coarse, rgb_fine) and passes dL/d rgb_coarse, dL/d rgb_fine [n,3] to
 * tgtc_style_train_backward (same workspace, same n/lat1/noise arguments), which writes
 *   grads  flat fp32 [tgtc_style_num_params()] in tgtc_set_style_weights order (module 1 (W,b) x 5, module 2 (W,b) x 8)
 *   dlat1  [n,32]  dL/d lat1 (module 2 sees mean(lat1) broadcast to 32 dims; its share comes back as 1/32 per component)
 * lat1 [n,32]: per-ray latents.  rand [n,S] uniforms or NULL (perturb=False).  noise_*: randn*sigma_noise_std or NULL.
 * The NeRF nets are constants here (they are not in style_optimizer); no gradient flows through the resampling.
 * The context remembers which workspaces hold a forward stash and for which (n, samples, has_rand): a backward on any other
 * workspace, or with different arguments, fails with TGTC_ERR_STATE instead of reading garbage. */
size_t tgtc_style_train_workspace_bytes(tgtc_ctx* ctx, int64_t n_rays, int n_samples, int n_fine);
int64_t tgtc_style_num_params(void);
int tgtc_style_train_forward(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, int64_t n_rays, double near, double far,
                             int n_samples, int n_fine, const float* lat1, const float* rand, const float* noise_coarse,
                             const float* noise_fine, float* rgb_coarse, float* rgb_fine, void* workspace, size_t workspace_bytes,
                             tgtc_stream stream);
int tgtc_style_train_backward(tgtc_ctx* ctx, int64_t n_rays, int n_samples, int n_fine, const float* lat1, int has_rand,
                              const float* noise_coarse, const float* noise_fine, const float* d_rgb_coarse, const float* d_rgb_fine,
                              float* grads, int accumulate, float* dlat1, void* workspace, size_t workspace_bytes, tgtc_stream stream);
/* the same pair with perturb / sigma noise drawn inside the kernels from Philox4x32-10 (see tgtc_train_step_seeded): pass the
 * same (seed, perturb, sigma_noise_std) to the backward, which regenerates the noise instead of reading it */
int tgtc_style_train_forward_seeded(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, int64_t n_rays, double near, double far,
                                    int n_samples, int n_fine, const float* lat1, unsigned long long seed, int perturb,
                                    double sigma_noise_std, float* rgb_coarse, float* rgb_fine, void* workspace, size_t workspace_bytes,
                                    tgtc_stream stream);
int tgtc_style_train_backward_seeded(tgtc_ctx* ctx, int64_t n_rays, int n_samples, int n_fine, const float* lat1, unsigned long long seed,
                                     int perturb, double sigma_noise_std, const float* d_rgb_coarse, const float* d_rgb_fine, float* grads,
                                     int accumulate, float* dlat1, void* workspace, size_t workspace_bytes, tgtc_stream stream);

/* ---- the per-ray losses of Style_train on the composited maps (train_tgtcs.py:397-404, :425, :449-459, :480-484) --------------
 * tgtc_style_loss_sums:  sums[0] = sum (rgb_coarse - gt)^2, sums[1] = sum (rgb_fine - gt)^2 over [n,3];
 *   sums[2] = sum_i (cos(coh_coarse_i, prev_coarse_i) - cos(rgb_origin_i, prev_origin_i))^2,
 *   sums[3] = sum_i (cos(coh_fine_i, prev_fine_i) - cos(rgb_origin_i, rgb_origin_i))^2 over [n_coh,3]   (VGGNet.py:204-210;
 *   the fine term's reference similarity uses this batch's own originals: train_tgtcs.py:403 precedes :456).
 *   coh_coarse == NULL: no coherence term (sums[2] = sums[3] = 0).
 * tgtc_style_loss_grads: gradients of  scale_rgb * (sums[0] + sums[1]) + scale_coh * (sqrt(coh_ss[0] + 1e-8) + sqrt(coh_ss[1] + 1e-8))
 *   (utils.L2_norm, utils.py:459) w.r.t. the four maps; coh_ss = sums[2..3], summed over all ranks when the batch is sharded.
 * With scale_rgb = rgb_loss_lambda / (3 n world) and scale_coh = loss_coh_lambda these are d loss / d maps of the iteration. */
int tgtc_style_loss_sums(tgtc_ctx* ctx, const float* rgb_coarse, const float* rgb_fine, const float* rgb_gt, int64_t n,
                         const float* coh_coarse, const float* coh_fine, const float* prev_coarse, const float* prev_fine,
                         const float* rgb_origin, const float* prev_origin, int64_t n_coh, float* sums, tgtc_stream stream);
int tgtc_style_loss_grads(tgtc_ctx* ctx, const float* rgb_coarse, const float* rgb_fine, const float* rgb_gt, int64_t n,
                          const float* coh_coarse, const float* coh_fine, const float* prev_coarse, const float* prev_fine,
                          const float* rgb_origin, const float* prev_origin, int64_t n_coh, const float* coh_ss, double scale_rgb,
                          double scale_coh, float* d_rgb_coarse, float* d_rgb_fine, float* d_coh_coarse, float* d_coh_fine,
                          tgtc_stream stream);

/* ---- the latent model of Style_train (models.StyleLatents_variational, models.py:475-549) -----------------------------------
 * forward:  lat[i] = mu[style_id[i]] + sigma_scale * (table[(style_id[i] * frame_num + frame_id[i]) mod rows] - mu[style_id[i]])
 *           (models.py:490-506; the modulo is the reference's 7x tiling of the LLFF table, :496: table_tiles = 7 for
 *           dataset_type 'llff', 1 otherwise; a flat id at or beyond rows * table_tiles -- an IndexError in the reference -- traps), and
 *           logp_sum = sum_{i < n_logp} sum_k (lat_ik - mu_k)^2 / (exp(0.5 logvar_k) + 1e-3)   (minus_logp = logp_sum / n_logp, :531-537)
 * backward: table_grad[row] (+)= d/d table of  sum_i <dlat[i], lat[i]> + logp_scale * logp_sum  -- row by row, in ray order.
 * table [rows,32], mu / logvar [style_num,32], style_id / frame_id int64 [n], lat / dlat [n,32] fp32, all on the device. */
int tgtc_style_latents_forward(tgtc_ctx* ctx, const float* table, const float* mu, const float* logvar, const int64_t* style_id,
                               const int64_t* frame_id, int64_t n, int64_t n_logp, int rows, int frame_num, int table_tiles, double sigma_scale,
                               float* lat, float* logp_sum, tgtc_stream stream);
int tgtc_style_latents_backward(tgtc_ctx* ctx, const float* table, const float* mu, const float* logvar, const int64_t* style_id,
                                const int64_t* frame_id, int64_t n, int64_t n_logp, int rows, int frame_num, int table_tiles, double sigma_scale,
                                const float* dlat, double logp_scale, float* table_grad, int accumulate, tgtc_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* TGTC_B200_H */
