// C-ABI entry points of libtgtc_b200 (declared in include/tgtc_b200.h):
// argument validation, error plumbing, and the render pipeline that strings the
// kernels together in the order of the reference's loop body (rendering.py:27-51).
#include "common.cuh"

#include <string.h>

// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void tgtc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int tgtc_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  tgtc_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return TGTC_ERR_CUDA;
}

extern "C" const char* tgtc_last_error(void) { return g_err; }
extern "C" int tgtc_abi_version(void) { return TGTC_ABI_VERSION; }

#define CHECK_CTX(ctx)                                                  \
  TGTC_REQUIRE((ctx) != nullptr, TGTC_ERR_ARG, "null context");         \
  DeviceGuard _guard((ctx)->device);                                    \
  TGTC_REQUIRE(_guard.ok, TGTC_ERR_CUDA, "cannot select device %d", (ctx)->device)

#define CHECK_NET(ctx, which)                                                               \
  TGTC_REQUIRE((which) == TGTC_NET_COARSE || (which) == TGTC_NET_FINE, TGTC_ERR_ARG, "bad net id %d", (which)); \
  TGTC_REQUIRE((ctx)->net[(which)].set, TGTC_ERR_STATE, "weights of net %d not set (call tgtc_set_weights)", (which))

#define CHECK_MODE(mode) \
  TGTC_REQUIRE((mode) == TGTC_MLP_FP32 || (mode) == TGTC_MLP_BF16 || (mode) == TGTC_MLP_F16, TGTC_ERR_ARG, "bad MLP mode %d", (mode))

#define CHECK_PTR(p, name) TGTC_REQUIRE((p) != nullptr && aligned4(p), TGTC_ERR_ARG, "%s is null or misaligned", name)

// ---------------------------------------------------------------------------
extern "C" int tgtc_create(int device, tgtc_ctx** out) {
  TGTC_REQUIRE(out != nullptr, TGTC_ERR_ARG, "out is null");
  *out = nullptr;
  int count = 0;
  TGTC_CUDA(cudaGetDeviceCount(&count));
  TGTC_REQUIRE(device >= 0 && device < count, TGTC_ERR_ARG, "device %d out of range (%d visible)", device, count);
  cudaDeviceProp prop;
  TGTC_CUDA(cudaGetDeviceProperties(&prop, device));
  TGTC_REQUIRE(prop.major == 10, TGTC_ERR_UNSUPPORTED,
               "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
  tgtc_ctx* c = new tgtc_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  *out = c;
  return TGTC_OK;
}

extern "C" int tgtc_destroy(tgtc_ctx* ctx) {
  if (ctx == nullptr) return TGTC_OK;
  DeviceGuard g(ctx->device);
  for (int i = 0; i < 2; ++i) {
    if (ctx->net[i].f32_gemm) cudaFree(ctx->net[i].f32_gemm);
    if (ctx->net[i].smalls) cudaFree(ctx->net[i].smalls);
    if (ctx->net[i].tc_blob) cudaFree(ctx->net[i].tc_blob);
    if (ctx->net[i].tc_blob_h) cudaFree(ctx->net[i].tc_blob_h);
    if (ctx->net[i].tc_blobT) cudaFree(ctx->net[i].tc_blobT);
  }
  {
    StyleImage& si = ctx->style;
    void* bufs[] = {si.blob_c, si.blob_w, si.blob_c_h, si.blob_w_h, si.blob_T, si.head_w, si.bias_c, si.bias_w, si.head_b, si.latents, si.tables, si.wlat, si.wlatT};
    for (void* b : bufs) if (b) cudaFree(b);
  }
  if (ctx->arena) cudaFree(ctx->arena);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
  if (ctx->aux_join) cudaEventDestroy(ctx->aux_join);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  delete ctx;
  return TGTC_OK;
}

extern "C" int64_t tgtc_launch_count(const tgtc_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int tgtc_set_weights(tgtc_ctx* ctx, int net, const float* const* params, tgtc_stream stream) {
  CHECK_CTX(ctx);
  TGTC_REQUIRE(net == TGTC_NET_COARSE || net == TGTC_NET_FINE, TGTC_ERR_ARG, "bad net id %d", net);
  TGTC_REQUIRE(params != nullptr, TGTC_ERR_ARG, "params is null");
  for (int i = 0; i < TGTC_NUM_PARAMS; ++i)
    TGTC_REQUIRE(params[i] != nullptr && aligned4(params[i]), TGTC_ERR_ARG, "params[%d] is null or misaligned", i);
  return pack_weights(ctx, net, params, (cudaStream_t)stream);
}

extern "C" int tgtc_raygen(tgtc_ctx* ctx, int H, int W, const double* K, const double* c2w, int ndc, double ndc_near,
                           int pixel_alignment, int64_t pix_begin, int64_t n, float* rays_o, float* rays_d,
                           tgtc_stream stream) {
  CHECK_CTX(ctx);
  TGTC_REQUIRE(H > 0 && W > 0, TGTC_ERR_ARG, "bad frame size %dx%d", H, W);
  TGTC_REQUIRE(K != nullptr && c2w != nullptr, TGTC_ERR_ARG, "K / c2w is null");
  TGTC_REQUIRE(n >= 0 && pix_begin >= 0 && pix_begin + n <= (int64_t)H * W, TGTC_ERR_ARG,
               "pixel range [%lld,%lld) outside the %dx%d frame", (long long)pix_begin, (long long)(pix_begin + n), H, W);
  if (n == 0) return TGTC_OK;
  CHECK_PTR(rays_o, "rays_o");
  CHECK_PTR(rays_d, "rays_d");
  return launch_raygen(ctx, H, W, K, c2w, ndc, ndc_near, pixel_alignment, pix_begin, n, rays_o, rays_d, (cudaStream_t)stream);
}

extern "C" int tgtc_sample_uniform(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, int64_t n, int n_samples,
                                   double near, double far, int harmony, const float* rnd, float* pts, float* ts, tgtc_stream stream) {
  CHECK_CTX(ctx);
  TGTC_REQUIRE(n >= 0 && n_samples >= 1, TGTC_ERR_ARG, "bad sizes n=%lld S=%d", (long long)n, n_samples);
  if (n == 0) return TGTC_OK;
  CHECK_PTR(ts, "ts");
  if (pts != nullptr) { CHECK_PTR(rays_o, "rays_o"); CHECK_PTR(rays_d, "rays_d"); }
  return launch_sample_uniform(ctx, rays_o, rays_d, n, n_samples, near, far, rnd, pts, ts, (cudaStream_t)stream, nullptr, harmony);
}

// per-launch device timing: a pair of events of kind `kind` around the next launch(es) on `st`
static int prof_begin(tgtc_ctx* ctx, int kind, double flops, cudaStream_t st, cudaEvent_t* e1) {
  *e1 = nullptr;
  if (!ctx->profile) return TGTC_OK;
  if (ctx->ev_used + 2 > ctx->ev_pool.size()) {
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t e;
      TGTC_CUDA(cudaEventCreate(&e));
      ctx->ev_pool.push_back(e);
    }
  }
  if (ctx->ev_kind.size() < ctx->ev_pool.size() / 2) ctx->ev_kind.resize(ctx->ev_pool.size() / 2, 0);
  ctx->ev_kind[ctx->ev_used / 2] = kind;
  cudaEvent_t e0 = ctx->ev_pool[ctx->ev_used++];
  *e1 = ctx->ev_pool[ctx->ev_used++];
  if (kind == 0) ctx->prof_flops += flops;
  ctx->prof_work[kind & 3] += flops;
  TGTC_CUDA(cudaEventRecord(e0, st));
  return TGTC_OK;
}

static int run_mlp(tgtc_ctx* ctx, int net, int mode, const MlpIO& io, cudaStream_t st) {
  cudaEvent_t e1 = nullptr;
  int prc = prof_begin(ctx, 0, (double)io.n_rays * io.S * 1186816.0, st, &e1);
  if (prc) return prc;
  int rc;
  if (mode == TGTC_MLP_BF16 || mode == TGTC_MLP_F16) {
    TGTC_REQUIRE(mlp_tc_supports(io), TGTC_ERR_UNSUPPORTED,
                 "tensor-core MLP needs per-ray view dirs, S in {64,128} or a multiple of 128, and no feature outputs (S=%d)", io.S);
    rc = launch_mlp_tc(ctx, net, io, st, mode == TGTC_MLP_F16);
  } else {
    rc = launch_mlp_fp32(ctx, net, io, st);
  }
  if (rc) return rc;
  if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
  return TGTC_OK;
}

extern "C" int tgtc_profile_enable(tgtc_ctx* ctx, int on) {
  CHECK_CTX(ctx);
  ctx->profile = on != 0;
  ctx->ev_used = 0;
  ctx->prof_flops = 0.0;
  for (double& w : ctx->prof_work) w = 0.0;
  return TGTC_OK;
}

// kind: 0 inference MLP forward, 1 training MLP forward (stash), 2 activation-gradient kernel, 3 weight-gradient kernel.
// Does not reset (tgtc_profile_enable does).
extern "C" int tgtc_profile_read_kind(tgtc_ctx* ctx, int kind, int64_t* launches, double* ms, double* flops) {
  CHECK_CTX(ctx);
  TGTC_REQUIRE(kind >= 0 && kind < 4, TGTC_ERR_ARG, "bad profile kind %d", kind);
  double total = 0.0;
  int64_t cnt = 0;
  for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
    if (ctx->ev_kind[i / 2] != kind) continue;
    TGTC_CUDA(cudaEventSynchronize(ctx->ev_pool[i + 1]));
    float t = 0.f;
    TGTC_CUDA(cudaEventElapsedTime(&t, ctx->ev_pool[i], ctx->ev_pool[i + 1]));
    total += t;
    ++cnt;
  }
  if (launches) *launches = cnt;
  if (ms) *ms = total;
  if (flops) *flops = ctx->prof_work[kind];
  return TGTC_OK;
}

extern "C" int tgtc_profile_read(tgtc_ctx* ctx, int64_t* launches, double* ms, double* flops) {
  CHECK_CTX(ctx);
  double total = 0.0;
  int64_t cnt = 0;
  for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
    if (!ctx->ev_kind.empty() && ctx->ev_kind[i / 2] != 0) continue;
    TGTC_CUDA(cudaEventSynchronize(ctx->ev_pool[i + 1]));
    float t = 0.f;
    TGTC_CUDA(cudaEventElapsedTime(&t, ctx->ev_pool[i], ctx->ev_pool[i + 1]));
    total += t;
    ++cnt;
  }
  if (launches) *launches = cnt;
  if (ms) *ms = total;
  if (flops) *flops = ctx->prof_flops;
  ctx->ev_used = 0;
  ctx->prof_flops = 0.0;
  for (double& w : ctx->prof_work) w = 0.0;
  return TGTC_OK;
}

extern "C" int tgtc_nerf_forward(tgtc_ctx* ctx, int net, int mode, const float* pts, const float* dirs, int dirs_per_ray,
                                 int64_t n_rays, int S, float* rgb, float* sigma, float* base_remap, float* pts_embed,
                                 float* dirs_embed, tgtc_stream stream) {
  CHECK_CTX(ctx);
  CHECK_NET(ctx, net);
  CHECK_MODE(mode);
  TGTC_REQUIRE(n_rays >= 0 && S >= 1, TGTC_ERR_ARG, "bad sizes n_rays=%lld S=%d", (long long)n_rays, S);
  if (n_rays == 0) return TGTC_OK;
  CHECK_PTR(pts, "pts");
  CHECK_PTR(dirs, "dirs");
  CHECK_PTR(rgb, "rgb");
  CHECK_PTR(sigma, "sigma");
  MlpIO io;
  io.pts = pts; io.dirs = dirs; io.dirs_per_ray = dirs_per_ray ? 1 : 0;
  io.n_rays = n_rays; io.S = S;
  io.rgb = rgb; io.sigma = sigma; io.base_remap = base_remap; io.pts_embed = pts_embed; io.dirs_embed = dirs_embed;
  return run_mlp(ctx, net, mode, io, (cudaStream_t)stream);
}

extern "C" int tgtc_nerf_forward_rays(tgtc_ctx* ctx, int net, int mode, const float* rays_o, const float* rays_d,
                                      const float* ts, int64_t n_rays, int S, double near, double far, float* rgbsigma,
                                      tgtc_stream stream) {
  CHECK_CTX(ctx);
  CHECK_NET(ctx, net);
  CHECK_MODE(mode);
  TGTC_REQUIRE(n_rays >= 0 && S >= 1, TGTC_ERR_ARG, "bad sizes n_rays=%lld S=%d", (long long)n_rays, S);
  if (n_rays == 0) return TGTC_OK;
  CHECK_PTR(rays_o, "rays_o");
  CHECK_PTR(rays_d, "rays_d");
  TGTC_REQUIRE(rgbsigma != nullptr && aligned16(rgbsigma), TGTC_ERR_ARG, "rgbsigma is null or not 16-byte aligned");
  MlpIO io;
  io.rays_o = rays_o; io.rays_d = rays_d; io.ts = ts;
  io.t_scale = (float)(far - near); io.t_near = (float)near;
  io.n_rays = n_rays; io.S = S; io.rgbsigma = rgbsigma;
  return run_mlp(ctx, net, mode, io, (cudaStream_t)stream);
}

extern "C" int tgtc_composite(tgtc_ctx* ctx, const float* rgb, const float* sigma, const float* rgbsigma, const float* ts,
                              int64_t ts_ray_stride, const float* noise, int white_bkgd, int64_t n, int S, float* rgb_out,
                              float* depth_out, float* acc_out, float* weights_out, tgtc_stream stream) {
  CHECK_CTX(ctx);
  TGTC_REQUIRE(n >= 0 && S >= 1, TGTC_ERR_ARG, "bad sizes n=%lld S=%d", (long long)n, S);
  if (n == 0) return TGTC_OK;
  TGTC_REQUIRE((rgbsigma != nullptr) != (rgb != nullptr && sigma != nullptr), TGTC_ERR_ARG,
               "pass either rgbsigma or (rgb and sigma)");
  if (rgbsigma != nullptr) TGTC_REQUIRE(aligned16(rgbsigma), TGTC_ERR_ARG, "rgbsigma not 16-byte aligned");
  CHECK_PTR(ts, "ts");
  TGTC_REQUIRE(ts_ray_stride == 0 || ts_ray_stride >= S, TGTC_ERR_ARG, "bad ts_ray_stride %lld", (long long)ts_ray_stride);
  return launch_composite(ctx, rgb, sigma, rgbsigma, ts, ts_ray_stride, noise, white_bkgd, n, S, rgb_out, depth_out, acc_out,
                          weights_out, (cudaStream_t)stream);
}

extern "C" int tgtc_composite_backward(tgtc_ctx* ctx, const float* rgbsigma, const float* ts, int64_t ts_ray_stride,
                                       const float* noise, int white_bkgd, int64_t n, int S, const float* g_rgb,
                                       const float* g_depth, const float* g_acc, float* d_rgbsigma, tgtc_stream stream) {
  CHECK_CTX(ctx);
  TGTC_REQUIRE(n >= 0 && S >= 1, TGTC_ERR_ARG, "bad sizes n=%lld S=%d", (long long)n, S);
  if (n == 0) return TGTC_OK;
  CHECK_PTR(rgbsigma, "rgbsigma");
  CHECK_PTR(ts, "ts");
  CHECK_PTR(g_rgb, "g_rgb");
  CHECK_PTR(d_rgbsigma, "d_rgbsigma");
  TGTC_REQUIRE(aligned16(rgbsigma) && aligned16(d_rgbsigma), TGTC_ERR_ARG, "rgbsigma / d_rgbsigma not 16-byte aligned");
  TGTC_REQUIRE(ts_ray_stride == 0 || ts_ray_stride >= S, TGTC_ERR_ARG, "bad ts_ray_stride %lld", (long long)ts_ray_stride);
  return launch_composite_backward(ctx, rgbsigma, ts, ts_ray_stride, noise, white_bkgd, n, S, g_rgb, g_depth, g_acc, d_rgbsigma,
                                   (cudaStream_t)stream);
}

extern "C" int tgtc_sample_fine(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, const float* ts,
                                int64_t ts_ray_stride, const float* weights, int64_t n, int S, int n_fine, float* pts_out,
                                float* ts_out, int64_t* inds_out, float* samples_out, tgtc_stream stream) {
  CHECK_CTX(ctx);
  TGTC_REQUIRE(n >= 0, TGTC_ERR_ARG, "bad n=%lld", (long long)n);
  if (n == 0) return TGTC_OK;
  CHECK_PTR(ts, "ts");
  CHECK_PTR(weights, "weights");
  CHECK_PTR(ts_out, "ts_out");
  TGTC_REQUIRE(ts_ray_stride == 0 || ts_ray_stride >= S, TGTC_ERR_ARG, "bad ts_ray_stride %lld", (long long)ts_ray_stride);
  if (pts_out != nullptr) { CHECK_PTR(rays_o, "rays_o"); CHECK_PTR(rays_d, "rays_d"); }
  return launch_sample_fine(ctx, rays_o, rays_d, ts, ts_ray_stride, weights, n, S, n_fine, pts_out, ts_out, inds_out,
                            samples_out, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
// the fused operator

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct RenderWs {
  size_t off_rs_c, off_rs_f, off_w_c, off_ts_f, off_ts_c, total;
};

// test / A-B hook (not in the public header): 1 = keep the stand-alone compositing kernel in the tensor-core render path
static int g_no_fused_composite = 0;
extern "C" void tgtc_debug_no_fused_composite(int on) { g_no_fused_composite = on; }
static int g_no_fused_sample_fine = 0;     // 1 = keep the stand-alone resampling kernel between the coarse and the fine launch
extern "C" void tgtc_debug_no_fused_sample_fine(int on) { g_no_fused_sample_fine = on; }

// which passes composite inside the MLP kernel (fused K5): tensor-core modes with whole rays per 128-sample tile
static inline bool fused_pass(int mode, int samples) {
  return (mode == TGTC_MLP_BF16 || mode == TGTC_MLP_F16) && (samples == 64 || samples == 128) && !g_no_fused_composite;
}

static RenderWs render_ws_layout(int64_t pass_rays, int S, int F, int mode) {
  RenderWs w;
  size_t o = 0;
  // per-sample (r,g,b,sigma) only exists for passes that composite in a separate kernel; a fused pass keeps 12 B/ray of scratch
  // for an rgb map the caller did not ask for
  w.off_rs_c = o; o = align_up(o + (size_t)pass_rays * (fused_pass(mode, S) ? 12 : (size_t)S * 16), 256);
  w.off_rs_f = o; o = align_up(o + (size_t)pass_rays * (fused_pass(mode, S + F) ? 12 : (size_t)(S + F) * 16), 256);
  w.off_w_c = o;  o = align_up(o + (size_t)pass_rays * S * 4, 256);
  w.off_ts_f = o; o = align_up(o + (size_t)pass_rays * (S + F) * 4, 256);
  w.off_ts_c = o; o = align_up(o + (size_t)S * 4, 256);
  w.total = o;
  return w;
}

static inline int64_t pass_size(int64_t n, int64_t chunk) { return (chunk <= 0 || chunk > n) ? n : chunk; }

extern "C" size_t tgtc_render_workspace_bytes(int64_t n_rays, int n_samples, int n_fine, int64_t chunk) {
  if (n_rays <= 0 || n_samples <= 0 || n_fine < 0) return 0;
  return render_ws_layout(pass_size(n_rays, chunk), n_samples, n_fine, TGTC_MLP_FP32).total;   // enough for every mode
}
extern "C" size_t tgtc_render_workspace_bytes_mode(int mode, int64_t n_rays, int n_samples, int n_fine, int64_t chunk) {
  if (n_rays <= 0 || n_samples <= 0 || n_fine < 0) return 0;
  return render_ws_layout(pass_size(n_rays, chunk), n_samples, n_fine, mode).total;
}

static int check_render_args(tgtc_ctx* ctx, int mode, int64_t n, int S, int F) {
  CHECK_MODE(mode);
  CHECK_NET(ctx, TGTC_NET_COARSE);
  TGTC_REQUIRE(n >= 0, TGTC_ERR_ARG, "bad n_rays=%lld", (long long)n);
  TGTC_REQUIRE(S >= 10 && S <= 128, TGTC_ERR_UNSUPPORTED, "n_samples=%d outside [10,128]", S);
  TGTC_REQUIRE(F >= 0 && S + F <= 256, TGTC_ERR_UNSUPPORTED, "n_samples+n_fine=%d > 256", S + F);
  if (F > 0) { CHECK_NET(ctx, TGTC_NET_FINE); }
  return TGTC_OK;
}

// rendering.py:27-51 for rays [0,n): passes of `chunk` rays; tensor-core modes: coarse MLP(+compositing) -> resampling ->
// fine MLP(+compositing) = three launches per pass; fp32 mode: five
static int render_impl(tgtc_ctx* ctx, int mode, const float* rays_o, const float* rays_d, int64_t n, double near, double far,
                       int S, int F, int64_t chunk, int white_bkgd, const tgtc_render_out& out, void* workspace,
                       size_t workspace_bytes, cudaStream_t st) {
  if (n == 0) return TGTC_OK;
  const int64_t pass = pass_size(n, chunk);
  const RenderWs ws = render_ws_layout(pass, S, F, mode);
  TGTC_REQUIRE(workspace != nullptr && aligned16(workspace) && workspace_bytes >= ws.total, TGTC_ERR_STATE,
               "workspace too small or misaligned: need %zu bytes, got %zu", ws.total, workspace_bytes);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  float* rs_c = reinterpret_cast<float*>(base + ws.off_rs_c);
  float* rs_f = reinterpret_cast<float*>(base + ws.off_rs_f);
  float* ts_c = reinterpret_cast<float*>(base + ws.off_ts_c);
  const int T = S + F;
  // the shared coarse t row (utils.py:512-516; the reference expands it with stride 0 over rays): only the stand-alone
  // compositing / resampling kernels read it -- the fused passes form it in-kernel
  int rc = TGTC_OK;
  if (!(fused_pass(mode, S) && (F == 0 || (fused_pass(mode, T) && S == 64 && F == 64 && !g_no_fused_sample_fine)))) {
    rc = launch_sample_uniform(ctx, nullptr, nullptr, 1, S, near, far, nullptr, nullptr, ts_c, st);
    if (rc) return rc;
  }
  for (int64_t r0 = 0; r0 < n; r0 += pass) {
    const int64_t m = (n - r0 < pass) ? (n - r0) : pass;
    const float* o = rays_o + r0 * 3;
    const float* d = rays_d + r0 * 3;
    float* w_c = out.weights_coarse ? out.weights_coarse + r0 * S : reinterpret_cast<float*>(base + ws.off_w_c);
    float* ts_f = out.ts_fine ? out.ts_fine + r0 * T : reinterpret_cast<float*>(base + ws.off_ts_f);
    // coarse: sampling_pts_uniform + StyleNerf('coarse') fused (rendering.py:27-34)
    MlpIO ic;
    ic.rays_o = o; ic.rays_d = d; ic.ts = nullptr;
    ic.t_scale = (float)(far - near); ic.t_near = (float)near;
    ic.n_rays = m; ic.S = S; ic.rgbsigma = rs_c;
    const bool last = (F == 0);
    float* c_rgb = out.rgb_coarse ? out.rgb_coarse + r0 * 3 : (last && out.rgb ? out.rgb + r0 * 3 : nullptr);
    float* c_depth = out.depth_coarse ? out.depth_coarse + r0 : (last && out.depth ? out.depth + r0 : nullptr);
    float* c_acc = out.acc_coarse ? out.acc_coarse + r0 : (last && out.acc ? out.acc + r0 : nullptr);
    // tensor-core modes with whole rays per 128-sample tile: alpha_composition (rendering.py:36) runs inside the MLP kernel's
    // last epilogue (fused K5) -- the per-sample (r,g,b,sigma) never reach HBM
    const bool fuse_c = fused_pass(mode, S);
    if (fuse_c) {
      ic.rgbsigma = nullptr;
      ic.comp_rgb = c_rgb != nullptr ? c_rgb : reinterpret_cast<float*>(base + ws.off_rs_c);   // rgb is always produced: scratch when unwanted
      ic.comp_depth = c_depth; ic.comp_acc = c_acc; ic.comp_weights = w_c; ic.comp_white_bkgd = white_bkgd;
    }
    rc = run_mlp(ctx, TGTC_NET_COARSE, mode, ic, st);
    if (rc) return rc;
    if (!fuse_c) {
      // alpha_composition(coarse) (rendering.py:36)
      rc = launch_composite(ctx, nullptr, nullptr, rs_c, ts_c, 0, nullptr, white_bkgd, m, S, c_rgb, c_depth, c_acc, w_c, st);
      if (rc) return rc;
    }
    if (last) {
      // coarse-only render: the coarse result IS the result; mirror it into the fine slots when both were asked for
      if (out.rgb && out.rgb_coarse) TGTC_CUDA(cudaMemcpyAsync(out.rgb + r0 * 3, out.rgb_coarse + r0 * 3, (size_t)m * 12, cudaMemcpyDeviceToDevice, st));
      if (out.depth && out.depth_coarse) TGTC_CUDA(cudaMemcpyAsync(out.depth + r0, out.depth_coarse + r0, (size_t)m * 4, cudaMemcpyDeviceToDevice, st));
      if (out.acc && out.acc_coarse) TGTC_CUDA(cudaMemcpyAsync(out.acc + r0, out.acc_coarse + r0, (size_t)m * 4, cudaMemcpyDeviceToDevice, st));
      if (out.weights) TGTC_CUDA(cudaMemcpyAsync(out.weights + r0 * S, w_c, (size_t)m * S * 4, cudaMemcpyDeviceToDevice, st));
      if (out.ts_fine) { rc = launch_sample_uniform(ctx, nullptr, nullptr, m, S, near, far, nullptr, nullptr, out.ts_fine + r0 * S, st); if (rc) return rc; }
      continue;
    }
    // sampling_pts_fine_torch (rendering.py:43): its own kernel, or -- tensor-core modes at 64 + 64 samples -- run by the fine MLP
    // kernel's input-producer warps one iteration ahead of each tile (fused K6+K7, same code: sample_fine.cuh)
    const bool fuse_f = fused_pass(mode, T);
    const bool fuse_sf = fuse_f && S == 64 && F == 64 && !g_no_fused_sample_fine;
    if (!fuse_sf) {
      rc = launch_sample_fine(ctx, nullptr, nullptr, ts_c, 0, w_c, m, S, F, nullptr, ts_f, nullptr, nullptr, st);
      if (rc) return rc;
    }
    // fine MLP on the sorted union (rendering.py:44-46)
    MlpIO fi;
    fi.rays_o = o; fi.rays_d = d; fi.ts = ts_f;
    fi.n_rays = m; fi.S = T; fi.rgbsigma = rs_f;
    if (fuse_sf) {
      fi.fine_weights = w_c; fi.fine_ts = ts_f;
      fi.fine_t_scale = (float)(far - near); fi.fine_t_near = (float)near;
    }
    if (fuse_f) {
      fi.rgbsigma = nullptr;
      fi.comp_rgb = out.rgb ? out.rgb + r0 * 3 : reinterpret_cast<float*>(base + ws.off_rs_f);
      fi.comp_depth = out.depth ? out.depth + r0 : nullptr;
      fi.comp_acc = out.acc ? out.acc + r0 : nullptr;
      fi.comp_weights = out.weights ? out.weights + r0 * T : nullptr;
      fi.comp_white_bkgd = white_bkgd;
    }
    rc = run_mlp(ctx, TGTC_NET_FINE, mode, fi, st);
    if (rc) return rc;
    if (!fuse_f) {
      // alpha_composition(fine) (rendering.py:51)
      rc = launch_composite(ctx, nullptr, nullptr, rs_f, ts_f, T, nullptr, white_bkgd, m, T,
                            out.rgb ? out.rgb + r0 * 3 : nullptr, out.depth ? out.depth + r0 : nullptr,
                            out.acc ? out.acc + r0 : nullptr, out.weights ? out.weights + r0 * T : nullptr, st);
      if (rc) return rc;
    }
  }
  return TGTC_OK;
}

extern "C" int tgtc_render(tgtc_ctx* ctx, int mode, const float* rays_o, const float* rays_d, int64_t n_rays, double near,
                           double far, int n_samples, int n_fine, int64_t chunk, int white_bkgd, const tgtc_render_out* out,
                           void* workspace, size_t workspace_bytes, tgtc_stream stream) {
  CHECK_CTX(ctx);
  int rc = check_render_args(ctx, mode, n_rays, n_samples, n_fine);
  if (rc) return rc;
  TGTC_REQUIRE(out != nullptr, TGTC_ERR_ARG, "out is null");
  if (n_rays == 0) return TGTC_OK;
  CHECK_PTR(rays_o, "rays_o");
  CHECK_PTR(rays_d, "rays_d");
  return render_impl(ctx, mode, rays_o, rays_d, n_rays, near, far, n_samples, n_fine, chunk, white_bkgd, *out, workspace,
                     workspace_bytes, (cudaStream_t)stream);
}

static int ensure_arena(tgtc_ctx* ctx, size_t bytes, cudaStream_t st) {
  if (ctx->arena_bytes >= bytes) return TGTC_OK;
  if (ctx->arena) {
    TGTC_CUDA(cudaStreamSynchronize(st));
    TGTC_CUDA(cudaFree(ctx->arena));
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
  }
  TGTC_CUDA(cudaMalloc(&ctx->arena, bytes));
  ctx->arena_bytes = bytes;
  return TGTC_OK;
}

extern "C" int tgtc_render_host(tgtc_ctx* ctx, int mode, const float* rays_o, const float* rays_d, int64_t n_rays,
                                double near, double far, int n_samples, int n_fine, int64_t chunk, int white_bkgd,
                                const tgtc_render_out* host_out, tgtc_stream stream) {
  CHECK_CTX(ctx);
  int rc = check_render_args(ctx, mode, n_rays, n_samples, n_fine);
  if (rc) return rc;
  TGTC_REQUIRE(host_out != nullptr, TGTC_ERR_ARG, "host_out is null");
  if (n_rays == 0) return TGTC_OK;
  TGTC_REQUIRE(rays_o != nullptr && rays_d != nullptr, TGTC_ERR_ARG, "rays_o / rays_d is null");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = n_rays;
  const int S = n_samples, F = n_fine, T = S + F;
  // arena: rays | requested outputs | render workspace
  struct Slot { float* const* hostp; size_t bytes; size_t off; };
  const tgtc_render_out& h = *host_out;
  Slot slots[9] = {
      {&h.rgb, (size_t)n * 12, 0},           {&h.depth, (size_t)n * 4, 0},        {&h.acc, (size_t)n * 4, 0},
      {&h.weights, (size_t)n * T * 4, 0},    {&h.rgb_coarse, (size_t)n * 12, 0},  {&h.depth_coarse, (size_t)n * 4, 0},
      {&h.acc_coarse, (size_t)n * 4, 0},     {&h.weights_coarse, (size_t)n * S * 4, 0}, {&h.ts_fine, (size_t)n * T * 4, 0}};
  size_t o = 0;
  const size_t off_o = o; o = align_up(o + (size_t)n * 12, 256);
  const size_t off_d = o; o = align_up(o + (size_t)n * 12, 256);
  for (auto& s : slots) {
    if (*s.hostp != nullptr) { s.off = o; o = align_up(o + s.bytes, 256); }
  }
  const size_t off_ws = o;
  const size_t ws_bytes = tgtc_render_workspace_bytes_mode(mode, n, S, F, chunk);
  o += ws_bytes;
  rc = ensure_arena(ctx, o, st);
  if (rc) return rc;
  uint8_t* base = static_cast<uint8_t*>(ctx->arena);
  TGTC_CUDA(cudaMemcpyAsync(base + off_o, rays_o, (size_t)n * 12, cudaMemcpyHostToDevice, st));
  TGTC_CUDA(cudaMemcpyAsync(base + off_d, rays_d, (size_t)n * 12, cudaMemcpyHostToDevice, st));
  tgtc_render_out dev;
  memset(&dev, 0, sizeof(dev));
  float** devp[9] = {&dev.rgb, &dev.depth, &dev.acc, &dev.weights, &dev.rgb_coarse, &dev.depth_coarse,
                     &dev.acc_coarse, &dev.weights_coarse, &dev.ts_fine};
  for (int i = 0; i < 9; ++i)
    if (*slots[i].hostp != nullptr) *devp[i] = reinterpret_cast<float*>(base + slots[i].off);
  rc = render_impl(ctx, mode, reinterpret_cast<const float*>(base + off_o), reinterpret_cast<const float*>(base + off_d), n, near,
                   far, S, F, chunk, white_bkgd, dev, base + off_ws, ws_bytes, st);
  if (rc) return rc;
  for (int i = 0; i < 9; ++i)
    if (*slots[i].hostp != nullptr)
      TGTC_CUDA(cudaMemcpyAsync(*slots[i].hostp, *devp[i], slots[i].bytes, cudaMemcpyDeviceToHost, st));
  TGTC_CUDA(cudaStreamSynchronize(st));
  return TGTC_OK;
}

extern "C" size_t tgtc_render_frame_workspace_bytes(int64_t n_rays, int n_samples, int n_fine, int64_t chunk) {
  if (n_rays <= 0) return 0;
  return align_up((size_t)n_rays * 12, 256) * 2 + tgtc_render_workspace_bytes(n_rays, n_samples, n_fine, chunk);
}
extern "C" size_t tgtc_render_frame_workspace_bytes_mode(int mode, int64_t n_rays, int n_samples, int n_fine, int64_t chunk) {
  if (n_rays <= 0) return 0;
  return align_up((size_t)n_rays * 12, 256) * 2 + tgtc_render_workspace_bytes_mode(mode, n_rays, n_samples, n_fine, chunk);
}

extern "C" int tgtc_render_frame(tgtc_ctx* ctx, int mode, int H, int W, const double* K, const double* c2w, int ndc,
                                 double ndc_near, int pixel_alignment, int64_t pix_begin, int64_t n, double near, double far, int n_samples,
                                 int n_fine, int64_t chunk, int white_bkgd, const tgtc_render_out* out, void* workspace,
                                 size_t workspace_bytes, tgtc_stream stream) {
  CHECK_CTX(ctx);
  int rc = check_render_args(ctx, mode, n, n_samples, n_fine);
  if (rc) return rc;
  TGTC_REQUIRE(out != nullptr, TGTC_ERR_ARG, "out is null");
  TGTC_REQUIRE(H > 0 && W > 0 && K != nullptr && c2w != nullptr, TGTC_ERR_ARG, "bad camera");
  TGTC_REQUIRE(pix_begin >= 0 && pix_begin + n <= (int64_t)H * W, TGTC_ERR_ARG, "pixel range outside the frame");
  if (n == 0) return TGTC_OK;
  const size_t need = tgtc_render_frame_workspace_bytes_mode(mode, n, n_samples, n_fine, chunk);
  TGTC_REQUIRE(workspace != nullptr && aligned16(workspace) && workspace_bytes >= need, TGTC_ERR_STATE,
               "workspace too small or misaligned: need %zu bytes, got %zu", need, workspace_bytes);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  const size_t ray_bytes = align_up((size_t)n * 12, 256);
  float* ro = reinterpret_cast<float*>(base);
  float* rd = reinterpret_cast<float*>(base + ray_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_raygen(ctx, H, W, K, c2w, ndc, ndc_near, pixel_alignment, pix_begin, n, ro, rd, st);
  if (rc) return rc;
  return render_impl(ctx, mode, ro, rd, n, near, far, n_samples, n_fine, chunk, white_bkgd, *out, base + 2 * ray_bytes,
                     workspace_bytes - 2 * ray_bytes, st);
}

// ---------------------------------------------------------------------------
// training step: forward with activation stash + backward, both nets

struct TrainWs {
  size_t off_stash_h, off_stash_f, off_stash_pe, off_stash_mask, off_dz, off_dzf, off_dhead, off_rs, off_drs, off_w_c, off_ts_f, off_ts_c,
      off_rgb, off_g, off_partial, total;
};

static TrainWs train_ws_layout(tgtc_ctx* ctx, int64_t n, int S, int F) {
  TrainWs w;
  const int64_t Mf = n * (S + F), Mc = n * S;
  const int64_t Mmax = Mf > Mc ? Mf : Mc;
  const size_t tiles = (size_t)((Mmax + 127) / 128);
  size_t o = 0;
  w.off_stash_h = o;  o = align_up(o + tiles * kStashHBytesPerTile, 1024);
  w.off_stash_f = o;  o = align_up(o + tiles * kStashFBytesPerTile, 1024);
  w.off_stash_pe = o; o = align_up(o + tiles * kStashPeBytesPerTile, 1024);
  w.off_stash_mask = o; o = align_up(o + tiles * kStashMaskBytesPerTile, 1024);
  w.off_dz = o;       o = align_up(o + tiles * kStashHBytesPerTile, 1024);
  w.off_dzf = o;      o = align_up(o + tiles * kStashFBytesPerTile, 1024);
  w.off_dhead = o;    o = align_up(o + tiles * kStashPeBytesPerTile, 1024);
  w.off_rs = o;       o = align_up(o + (size_t)Mmax * 16, 256);
  w.off_drs = o;      o = align_up(o + (size_t)Mmax * 16, 256);
  w.off_w_c = o;      o = align_up(o + (size_t)n * S * 4, 256);
  w.off_ts_f = o;     o = align_up(o + (size_t)n * (S + F) * 4, 256);
  w.off_ts_c = o;     o = align_up(o + (size_t)n * S * 4, 256);
  w.off_rgb = o;      o = align_up(o + (size_t)n * 12, 256);
  w.off_g = o;        o = align_up(o + (size_t)n * 12, 256);
  w.off_partial = o;  o = align_up(o + (size_t)ctx->num_sms * bwd_partial_floats() * 4, 256);
  w.total = o;
  return w;
}

extern "C" size_t tgtc_train_workspace_bytes(tgtc_ctx* ctx, int64_t n_rays, int n_samples, int n_fine) {
  if (ctx == nullptr || n_rays <= 0 || n_samples <= 0 || n_fine < 0) return 0;
  return train_ws_layout(ctx, n_rays, n_samples, n_fine).total;
}

extern "C" int64_t tgtc_num_params(void) { return (int64_t)bwd_flat_floats(); }

// one network pass: forward (stash) -> compositing -> [caller hook: resampling] -> loss gradient -> compositing backward ->
// activation gradients -> weight gradients
static int train_pass(tgtc_ctx* ctx, int net, const float* rays_o, const float* rays_d, const float* ts, int64_t ts_stride, int64_t n,
                      int S, double near, double far, const float* noise, const PhiloxSrc* prng, const float* rgb_gt, float scale,
                      const TrainWs& ws, uint8_t* base, float* grads, int accumulate, float* sq_sum, float* rgb_out,
                      float* weights_out, cudaStream_t st) {
  TcStash stash;
  stash.h = base + ws.off_stash_h; stash.f = base + ws.off_stash_f; stash.pe = base + ws.off_stash_pe;
  stash.mask = reinterpret_cast<uint32_t*>(base + ws.off_stash_mask);
  TcDz dz;
  dz.dz = base + ws.off_dz; dz.dzf = base + ws.off_dzf; dz.dhead = base + ws.off_dhead;
  float* rs = reinterpret_cast<float*>(base + ws.off_rs);
  float* drs = reinterpret_cast<float*>(base + ws.off_drs);
  float* rgb = rgb_out != nullptr ? rgb_out : reinterpret_cast<float*>(base + ws.off_rgb);
  float* g = reinterpret_cast<float*>(base + ws.off_g);
  float* partial = reinterpret_cast<float*>(base + ws.off_partial);
  MlpIO io;
  io.rays_o = rays_o; io.rays_d = rays_d; io.ts = ts_stride == 0 ? nullptr : ts;
  io.t_scale = (float)(far - near); io.t_near = (float)near;
  io.n_rays = n; io.S = S; io.rgbsigma = rs;
  TGTC_REQUIRE(mlp_tc_supports(io), TGTC_ERR_UNSUPPORTED, "training needs n_samples in {64,128} (tcgen05 path)");
  const double samples = (double)n * S;
  cudaEvent_t e1 = nullptr;
  int rc = prof_begin(ctx, 1, samples * 1186816.0, st, &e1);
  if (rc) return rc;
  rc = launch_mlp_tc_train(ctx, net, io, stash, st);
  if (rc) return rc;
  if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
  rc = launch_composite(ctx, nullptr, nullptr, rs, ts, ts_stride, noise, 0, n, S, rgb, nullptr, nullptr, weights_out, st, prng);
  if (rc) return rc;
  rc = launch_mse_grad(ctx, rgb, rgb_gt, n, scale, g, sq_sum, st);
  if (rc) return rc;
  rc = launch_composite_backward(ctx, rs, ts, ts_stride, noise, 0, n, S, g, nullptr, nullptr, drs, st, prng);
  if (rc) return rc;
  // algorithmic FLOPs: dgrad skips the three slices into non-differentiable inputs (PE->L0, PE->L5, dirPE->rgb0)
  rc = prof_begin(ctx, 2, samples * 2.0 * (593408.0 - 16128.0 - 16128.0 - 3456.0), st, &e1);
  if (rc) return rc;
  rc = launch_mlp_dgrad(ctx, net, rs, drs, stash, dz, n * S, st);
  if (rc) return rc;
  if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
  rc = prof_begin(ctx, 3, samples * 1186816.0, st, &e1);
  if (rc) return rc;
  rc = launch_mlp_wgrad(ctx, stash, dz, rays_d, drs, n * S, S, partial, grads, accumulate, st);
  if (rc) return rc;
  if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
  return TGTC_OK;
}

static int train_step_impl(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, const float* rgb_gt, int64_t n_rays,
                           int64_t n_rays_total, double near, double far, int n_samples, int n_fine, const float* rand,
                           const float* noise_coarse, const float* noise_fine, const PhiloxSrc* jitter, const PhiloxSrc* prng_c,
                           const PhiloxSrc* prng_f, float* grads, int accumulate, float* loss_sums, float* rgb_coarse, float* rgb_fine,
                           void* workspace, size_t workspace_bytes, tgtc_stream stream) {
  CHECK_CTX(ctx);
  CHECK_NET(ctx, TGTC_NET_COARSE);
  CHECK_NET(ctx, TGTC_NET_FINE);
  TGTC_REQUIRE(n_rays >= 0 && n_rays_total >= n_rays, TGTC_ERR_ARG, "bad n_rays=%lld / n_rays_total=%lld", (long long)n_rays,
               (long long)n_rays_total);
  if (n_rays == 0) {
    // an empty shard (global batch < world size) still owns a gradient buffer that the caller all-reduces: it must read as zero
    if (!accumulate && grads != nullptr) TGTC_CUDA(cudaMemsetAsync(grads, 0, 2 * bwd_flat_floats() * sizeof(float), (cudaStream_t)stream));
    if (ctx->coarse_done != nullptr) TGTC_CUDA(cudaEventRecord(ctx->coarse_done, (cudaStream_t)stream));
    return TGTC_OK;
  }
  const int S = n_samples, F = n_fine;
  TGTC_REQUIRE(S == 64 && (S + F) == 128, TGTC_ERR_UNSUPPORTED,
               "training supports n_samples=64, n_fine=64 (configs/fern.txt:16-17); got %d+%d", S, F);
  CHECK_PTR(rays_o, "rays_o"); CHECK_PTR(rays_d, "rays_d"); CHECK_PTR(rgb_gt, "rgb_gt"); CHECK_PTR(grads, "grads");
  const TrainWs ws = train_ws_layout(ctx, n_rays, S, F);
  TGTC_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 1023) == 0 && workspace_bytes >= ws.total,
               TGTC_ERR_STATE, "training workspace too small or not 1024-byte aligned: need %zu bytes, got %zu", ws.total,
               workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  float* ts_c = reinterpret_cast<float*>(base + ws.off_ts_c);
  float* ts_f = reinterpret_cast<float*>(base + ws.off_ts_f);
  float* w_c = reinterpret_cast<float*>(base + ws.off_w_c);
  const size_t np = bwd_flat_floats();
  // loss = mse(rgb_gt, rgb_coarse) + mse(rgb_gt, rgb_fine), each a mean over n_rays_total*3 values (train_tgtcs.py:238-251)
  const float scale = 2.0f / (3.0f * (float)n_rays_total);
  // coarse sample positions: one shared row (perturb=False, utils.py:512-516) or per-ray stratified positions from the
  // caller's uniforms / the in-kernel Philox stream (perturb=True, utils.py:518-524)
  const bool per_ray = rand != nullptr || jitter != nullptr;
  const int64_t ts_c_stride = per_ray ? S : 0;
  int rc = launch_sample_uniform(ctx, nullptr, nullptr, per_ray ? n_rays : 1, S, near, far, rand, nullptr, ts_c, st, jitter);
  if (rc) return rc;
  rc = train_pass(ctx, TGTC_NET_COARSE, rays_o, rays_d, ts_c, ts_c_stride, n_rays, S, near, far, noise_coarse, prng_c, rgb_gt, scale, ws,
                  base, grads, accumulate, loss_sums, rgb_coarse, w_c, st);
  if (rc) return rc;
  // the coarse net's half of the gradient buffer is final here: a caller that all-reduces it on another stream waits for this
  // event while the fine net's forward / backward below still runs
  if (ctx->coarse_done != nullptr) TGTC_CUDA(cudaEventRecord(ctx->coarse_done, st));
  // no gradient flows through the resampling (utils.py:576-579): the two nets' backward passes are independent
  rc = launch_sample_fine(ctx, nullptr, nullptr, ts_c, ts_c_stride, w_c, n_rays, S, F, nullptr, ts_f, nullptr, nullptr, st);
  if (rc) return rc;
  return train_pass(ctx, TGTC_NET_FINE, rays_o, rays_d, ts_f, S + F, n_rays, S + F, near, far, noise_fine, prng_f, rgb_gt, scale, ws,
                    base, grads + np, accumulate, loss_sums != nullptr ? loss_sums + 1 : nullptr, rgb_fine, nullptr, st);
}

// ---------------------------------------------------------------------------
// stage-level training entries: the forward of ONE net with its activation stash, and the matching backward.  They back the
// torch.autograd.Function behind the model_forward drop-in (shims.py), so the reference's own Origin_train loop
// (train_tgtcs.py:228-255: model_forward -> alpha_composition -> loss.backward()) runs unchanged on the tensor-core kernels.
struct StageStash { size_t off_h, off_f, off_pe, off_mask, total; };
static StageStash stage_stash_layout(int64_t M) {
  StageStash w;
  const size_t tiles = (size_t)((M + 127) / 128);
  size_t o = 0;
  w.off_h = o;    o = align_up(o + tiles * kStashHBytesPerTile, 1024);
  w.off_f = o;    o = align_up(o + tiles * kStashFBytesPerTile, 1024);
  w.off_pe = o;   o = align_up(o + tiles * kStashPeBytesPerTile, 1024);
  w.off_mask = o; o = align_up(o + tiles * kStashMaskBytesPerTile, 1024);
  w.total = o;
  return w;
}
struct StageScratch { size_t off_dz, off_dzf, off_dhead, off_partial, total; };
static StageScratch stage_scratch_layout(tgtc_ctx* ctx, int64_t M) {
  StageScratch w;
  const size_t tiles = (size_t)((M + 127) / 128);
  size_t o = 0;
  w.off_dz = o;      o = align_up(o + tiles * kStashHBytesPerTile, 1024);
  w.off_dzf = o;     o = align_up(o + tiles * kStashFBytesPerTile, 1024);
  w.off_dhead = o;   o = align_up(o + tiles * kStashPeBytesPerTile, 1024);
  w.off_partial = o; o = align_up(o + (size_t)ctx->num_sms * bwd_partial_floats() * 4, 256);
  w.total = o;
  return w;
}
extern "C" size_t tgtc_nerf_stash_bytes(int64_t n_rays, int S) {
  if (n_rays <= 0 || S <= 0) return 0;
  return stage_stash_layout(n_rays * S).total;
}
extern "C" size_t tgtc_nerf_backward_scratch_bytes(tgtc_ctx* ctx, int64_t n_rays, int S) {
  if (ctx == nullptr || n_rays <= 0 || S <= 0) return 0;
  return stage_scratch_layout(ctx, n_rays * S).total;
}

static int stage_args(tgtc_ctx* ctx, int net, int64_t n_rays, int S, const void* stash, size_t stash_bytes) {
  CHECK_NET(ctx, net);
  TGTC_REQUIRE(n_rays >= 0, TGTC_ERR_ARG, "bad n_rays=%lld", (long long)n_rays);
  TGTC_REQUIRE(S == 64 || S == 128, TGTC_ERR_UNSUPPORTED, "the training kernels take S in {64,128} samples per ray; got %d", S);
  if (n_rays == 0) return TGTC_OK;
  const StageStash w = stage_stash_layout(n_rays * S);
  TGTC_REQUIRE(stash != nullptr && (reinterpret_cast<uintptr_t>(stash) & 1023) == 0 && stash_bytes >= w.total, TGTC_ERR_STATE,
               "stash too small or not 1024-byte aligned: need %zu bytes, got %zu", w.total, stash_bytes);
  return TGTC_OK;
}

extern "C" int tgtc_nerf_forward_stash(tgtc_ctx* ctx, int net, const float* pts, const float* dirs, int64_t n_rays, int S, float* rgbsigma,
                                       void* stash_p, size_t stash_bytes, tgtc_stream stream) {
  CHECK_CTX(ctx);
  int rc = stage_args(ctx, net, n_rays, S, stash_p, stash_bytes);
  if (rc) return rc;
  if (n_rays == 0) return TGTC_OK;
  CHECK_PTR(pts, "pts"); CHECK_PTR(dirs, "dirs"); CHECK_PTR(rgbsigma, "rgbsigma");
  TGTC_REQUIRE(aligned16(rgbsigma), TGTC_ERR_ARG, "rgbsigma not 16-byte aligned");
  const StageStash w = stage_stash_layout(n_rays * S);
  uint8_t* base = static_cast<uint8_t*>(stash_p);
  TcStash stash;
  stash.h = base + w.off_h; stash.f = base + w.off_f; stash.pe = base + w.off_pe;
  stash.mask = reinterpret_cast<uint32_t*>(base + w.off_mask);
  MlpIO io;
  io.pts = pts; io.dirs = dirs; io.dirs_per_ray = 1;
  io.n_rays = n_rays; io.S = S; io.rgbsigma = rgbsigma;
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e1 = nullptr;
  rc = prof_begin(ctx, 1, (double)n_rays * S * 1186816.0, st, &e1);
  if (rc) return rc;
  rc = launch_mlp_tc_train(ctx, net, io, stash, st);
  if (rc) return rc;
  if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
  return TGTC_OK;
}

extern "C" int tgtc_nerf_backward(tgtc_ctx* ctx, int net, const float* dirs, int64_t n_rays, int S, const float* rgbsigma,
                                  const float* d_rgbsigma, float* grads, int accumulate, void* stash_p, size_t stash_bytes, void* scratch_p,
                                  size_t scratch_bytes, tgtc_stream stream) {
  CHECK_CTX(ctx);
  int rc = stage_args(ctx, net, n_rays, S, stash_p, stash_bytes);
  if (rc) return rc;
  CHECK_PTR(grads, "grads");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rays == 0) {
    if (!accumulate) TGTC_CUDA(cudaMemsetAsync(grads, 0, bwd_flat_floats() * sizeof(float), st));
    return TGTC_OK;
  }
  CHECK_PTR(dirs, "dirs"); CHECK_PTR(rgbsigma, "rgbsigma"); CHECK_PTR(d_rgbsigma, "d_rgbsigma");
  TGTC_REQUIRE(aligned16(rgbsigma) && aligned16(d_rgbsigma), TGTC_ERR_ARG, "rgbsigma / d_rgbsigma not 16-byte aligned");
  const int64_t M = n_rays * S;
  const StageStash w = stage_stash_layout(M);
  const StageScratch sc = stage_scratch_layout(ctx, M);
  TGTC_REQUIRE(scratch_p != nullptr && (reinterpret_cast<uintptr_t>(scratch_p) & 1023) == 0 && scratch_bytes >= sc.total, TGTC_ERR_STATE,
               "backward scratch too small or not 1024-byte aligned: need %zu bytes, got %zu", sc.total, scratch_bytes);
  uint8_t* base = static_cast<uint8_t*>(stash_p);
  uint8_t* sb = static_cast<uint8_t*>(scratch_p);
  TcStash stash;
  stash.h = base + w.off_h; stash.f = base + w.off_f; stash.pe = base + w.off_pe;
  stash.mask = reinterpret_cast<uint32_t*>(base + w.off_mask);
  TcDz dz;
  dz.dz = sb + sc.off_dz; dz.dzf = sb + sc.off_dzf; dz.dhead = sb + sc.off_dhead;
  float* partial = reinterpret_cast<float*>(sb + sc.off_partial);
  const double samples = (double)M;
  cudaEvent_t e1 = nullptr;
  rc = prof_begin(ctx, 2, samples * 2.0 * (593408.0 - 16128.0 - 16128.0 - 3456.0), st, &e1);
  if (rc) return rc;
  rc = launch_mlp_dgrad(ctx, net, rgbsigma, d_rgbsigma, stash, dz, M, st);
  if (rc) return rc;
  if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
  rc = prof_begin(ctx, 3, samples * 1186816.0, st, &e1);
  if (rc) return rc;
  rc = launch_mlp_wgrad(ctx, stash, dz, dirs, d_rgbsigma, M, S, partial, grads, accumulate, st);
  if (rc) return rc;
  if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
  return TGTC_OK;
}

extern "C" int tgtc_train_set_coarse_event(tgtc_ctx* ctx, void* cuda_event) {
  TGTC_REQUIRE(ctx != nullptr, TGTC_ERR_ARG, "null context");
  ctx->coarse_done = (cudaEvent_t)cuda_event;
  return TGTC_OK;
}

extern "C" int tgtc_train_step(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, const float* rgb_gt, int64_t n_rays,
                               int64_t n_rays_total, double near, double far, int n_samples, int n_fine, const float* rand,
                               const float* noise_coarse, const float* noise_fine, float* grads, int accumulate,
                               float* loss_sums, float* rgb_coarse, float* rgb_fine, void* workspace, size_t workspace_bytes,
                               tgtc_stream stream) {
  return train_step_impl(ctx, rays_o, rays_d, rgb_gt, n_rays, n_rays_total, near, far, n_samples, n_fine, rand, noise_coarse, noise_fine,
                         nullptr, nullptr, nullptr, grads, accumulate, loss_sums, rgb_coarse, rgb_fine, workspace, workspace_bytes, stream);
}

// the same step with the reference's stochastic options drawn INSIDE the kernels from Philox4x32-10 (philox.cuh): stratified
// jitter in the sampling kernel, sigma noise in the compositing forward and (regenerated, not stored) backward
extern "C" int tgtc_train_step_seeded(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, const float* rgb_gt, int64_t n_rays,
                                      int64_t n_rays_total, double near, double far, int n_samples, int n_fine,
                                      unsigned long long seed, int perturb, double sigma_noise_std, float* grads, int accumulate,
                                      float* loss_sums, float* rgb_coarse, float* rgb_fine, void* workspace, size_t workspace_bytes,
                                      tgtc_stream stream) {
  const PhiloxSrc jit = {seed, kStreamJitter, 1.0f, 1};
  const PhiloxSrc nc = {seed, kStreamNoiseCoarse, (float)sigma_noise_std, 1};
  const PhiloxSrc nf = {seed, kStreamNoiseFine, (float)sigma_noise_std, 1};
  const bool noise = sigma_noise_std > 0.0;
  return train_step_impl(ctx, rays_o, rays_d, rgb_gt, n_rays, n_rays_total, near, far, n_samples, n_fine, nullptr, nullptr, nullptr,
                         perturb ? &jit : nullptr, noise ? &nc : nullptr, noise ? &nf : nullptr, grads, accumulate, loss_sums, rgb_coarse,
                         rgb_fine, workspace, workspace_bytes, stream);
}

// element e of stream `stream` (0 = jitter uniforms, 1 / 2 = coarse / fine sigma noise) under `seed`, as a tensor
extern "C" int tgtc_philox_fill(tgtc_ctx* ctx, unsigned long long seed, int stream_id, int normal, double std, int64_t n, float* out,
                                tgtc_stream stream) {
  CHECK_CTX(ctx);
  TGTC_REQUIRE(n >= 0 && stream_id >= 0, TGTC_ERR_ARG, "bad n=%lld / stream_id=%d", (long long)n, stream_id);
  if (n == 0) return TGTC_OK;
  CHECK_PTR(out, "out");
  DeviceGuard g(ctx->device);
  return launch_philox_fill(ctx, seed, (uint32_t)stream_id, normal, (float)std, n, out, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
// stylised render: the loop body of render_style (rendering.py:118-178) for one batch of rays

extern "C" int tgtc_set_style_weights(tgtc_ctx* ctx, const float* const* params, tgtc_stream stream) {
  CHECK_CTX(ctx);
  TGTC_REQUIRE(params != nullptr, TGTC_ERR_ARG, "params is null");
  for (int i = 0; i < 26; ++i) TGTC_REQUIRE(params[i] != nullptr && aligned4(params[i]), TGTC_ERR_ARG, "style param %d null or misaligned", i);
  DeviceGuard g(ctx->device);
  return style_set_weights(ctx, params, (cudaStream_t)stream);
}

struct StyleWs {
  size_t off_remap, off_cf, off_rs_c, off_rs_f, off_w_c, off_ts_f, off_ts_c, off_bias, total;
};
static StyleWs style_ws_layout(int64_t pass, int S, int F, bool per_ray = false) {
  StyleWs w;
  const size_t tiles = (size_t)((pass * (S + F) + 127) / 128);
  size_t o = 0;
  w.off_remap = o; o = align_up(o + tiles * 65536, 1024);
  w.off_cf = o;    o = align_up(o + tiles * 65536, 1024);
  w.off_rs_c = o;  o = align_up(o + (size_t)pass * S * 16, 256);
  w.off_rs_f = o;  o = align_up(o + (size_t)pass * (S + F) * 16, 256);
  w.off_w_c = o;   o = align_up(o + (size_t)pass * S * 4, 256);
  w.off_ts_f = o;  o = align_up(o + (size_t)pass * (S + F) * 4, 256);
  w.off_ts_c = o;  o = align_up(o + (size_t)S * 4, 256);
  w.off_bias = o;  if (per_ray) o = align_up(o + (size_t)pass * 13 * 256 * 4, 256);   // per-ray effective biases [pass][13][256] fp32
  w.total = o;
  return w;
}
static inline int64_t style_pass(int64_t n, int64_t chunk) { return pass_size(n, chunk <= 0 ? 32768 : chunk); }

// several passes: two workspace sets, even passes on the caller's stream, odd passes on the context's side stream
static inline size_t style_ws_sets(int64_t n_rays, int64_t pass) { return n_rays > pass ? 2 : 1; }
static inline size_t style_ws_set_bytes(const StyleWs& w) { return align_up(w.total, 1024); }

extern "C" size_t tgtc_render_style_workspace_bytes(int64_t n_rays, int n_samples, int n_fine, int64_t chunk) {
  if (n_rays <= 0 || n_samples <= 0 || n_fine < 0) return 0;
  const int64_t pass = style_pass(n_rays, chunk);
  return style_ws_set_bytes(style_ws_layout(pass, n_samples, n_fine)) * style_ws_sets(n_rays, pass);
}

static int render_style_impl(tgtc_ctx* ctx, int mode, const float* rays_o, const float* rays_d, int64_t n_rays, double near, double far,
                             int n_samples, int n_fine, int64_t chunk, const float* latent1, const float* latent2, const float* lat_rays,
                             const tgtc_render_out* out_p, void* workspace, size_t workspace_bytes, tgtc_stream stream) {
  CHECK_CTX(ctx);
  TGTC_REQUIRE(mode == TGTC_MLP_BF16 || mode == TGTC_MLP_F16, TGTC_ERR_ARG, "stylised render runs on the tensor-core path: mode must be BF16 or F16 (got %d)", mode);
  const bool f16 = mode == TGTC_MLP_F16;
  CHECK_NET(ctx, TGTC_NET_COARSE);
  CHECK_NET(ctx, TGTC_NET_FINE);
  TGTC_REQUIRE(ctx->style.set, TGTC_ERR_STATE, "style weights not set (tgtc_set_style_weights)");
  TGTC_REQUIRE(n_rays >= 0, TGTC_ERR_ARG, "bad n_rays=%lld", (long long)n_rays);
  if (n_rays == 0) return TGTC_OK;
  const int S = n_samples, F = n_fine;
  TGTC_REQUIRE(S == 64 && S + F == 128, TGTC_ERR_UNSUPPORTED, "stylised render supports n_samples=64, n_fine=64; got %d+%d", S, F);
  CHECK_PTR(rays_o, "rays_o"); CHECK_PTR(rays_d, "rays_d");
  if (lat_rays == nullptr) { CHECK_PTR(latent1, "latent1"); CHECK_PTR(latent2, "latent2"); }
  TGTC_REQUIRE(out_p != nullptr, TGTC_ERR_ARG, "out is null");
  const tgtc_render_out& out = *out_p;
  const int64_t pass = style_pass(n_rays, chunk);
  const StyleWs ws = style_ws_layout(pass, S, F, lat_rays != nullptr);
  const size_t nsets_ws = style_ws_sets(n_rays, pass), set_bytes = style_ws_set_bytes(ws);
  TGTC_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 1023) == 0 && workspace_bytes >= set_bytes * nsets_ws,
               TGTC_ERR_STATE, "style workspace too small or not 1024-byte aligned: need %zu bytes, got %zu", set_bytes * nsets_ws, workspace_bytes);
  // with the per-launch timers on (tgtc_profile_enable) everything stays on the caller's stream: overlapping launches of two
  // streams cannot be timed one by one
  const size_t nsets = ctx->profile ? 1 : nsets_ws;
  DeviceGuard g(ctx->device);
  cudaStream_t st0 = (cudaStream_t)stream;
  uint8_t* base0 = static_cast<uint8_t*>(workspace);
  float* ts_c = reinterpret_cast<float*>(base0 + ws.off_ts_c);   // the shared coarse positions: set 0's row, read by both streams
  const int T = S + F;
  int rc = TGTC_OK;
  if (nsets > 1 && ctx->aux_stream == nullptr) {
    TGTC_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    TGTC_CUDA(cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming));
    TGTC_CUDA(cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming));
  }
  if (lat_rays == nullptr) {
    rc = style_set_latents(ctx, latent1, latent2, st0);   // effective biases of both modules for this (style, frame)
    if (rc) return rc;
  }
  rc = launch_sample_uniform(ctx, nullptr, nullptr, 1, S, near, far, nullptr, nullptr, ts_c, st0);
  if (rc) return rc;
  if (nsets > 1) {   // fork: the side stream starts after everything queued so far on the caller's stream
    TGTC_CUDA(cudaEventRecord(ctx->aux_fork, st0));
    TGTC_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_fork, 0));
  }
  int64_t pass_idx = 0;
  for (int64_t r0 = 0; r0 < n_rays; r0 += pass, ++pass_idx) {
    // Passes are independent (different rays): even passes run on the caller's stream with workspace set 0, odd passes on the side
    // stream with set 1 -- one pass's kernel prologues / tails (TMEM allocation, barrier set-up, the last partial wave of tiles)
    // fill under the other pass's kernels instead of idling the SMs between ~10 dependent launches per pass.
    const int set = (nsets > 1) ? (int)(pass_idx & 1) : 0;
    cudaStream_t st = set ? ctx->aux_stream : st0;
    uint8_t* base = base0 + (size_t)set * set_bytes;
    uint8_t* remap = base + ws.off_remap;
    uint8_t* cf = base + ws.off_cf;
    float* rs_c = reinterpret_cast<float*>(base + ws.off_rs_c);
    float* rs_f = reinterpret_cast<float*>(base + ws.off_rs_f);
    float* bias_rays = lat_rays != nullptr ? reinterpret_cast<float*>(base + ws.off_bias) : nullptr;
    const int64_t m = (n_rays - r0 < pass) ? (n_rays - r0) : pass;
    float* w_c = out.weights_coarse ? out.weights_coarse + r0 * S : reinterpret_cast<float*>(base + ws.off_w_c);
    float* ts_f = out.ts_fine ? out.ts_fine + r0 * T : reinterpret_cast<float*>(base + ws.off_ts_f);
    if (bias_rays != nullptr) {
      // per-ray latents (rendering.py:125-127): every layer's latent columns folded into per-ray effective biases
      rc = launch_style_bias_rays(ctx, lat_rays + r0 * 32, m, bias_rays, st);
      if (rc) return rc;
    }
    for (int which = 0; which < 2; ++which) {
      MlpIO io;
      io.rays_o = rays_o + r0 * 3; io.rays_d = rays_d + r0 * 3;
      io.ts = which == 0 ? nullptr : ts_f;
      io.t_scale = (float)(far - near); io.t_near = (float)near;
      io.n_rays = m; io.S = which == 0 ? S : T;
      io.rgbsigma = which == 0 ? rs_c : rs_f;
      // NeRF trunk -> base_remap tile images + sigma (rendering.py:122-123 / :158-159)
      cudaEvent_t e1 = nullptr;
      rc = prof_begin(ctx, 0, (double)m * io.S * 2.0 * (593408.0 - 36224.0 - 384.0), st, &e1);
      if (rc) return rc;
      rc = launch_mlp_tc_trunk(ctx, which, io, remap, st, f16);
      if (rc) return rc;
      if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
      // style module 1 -> concat_features tile images (rendering.py:129-130)
      rc = prof_begin(ctx, 1, (double)m * io.S * 2.0 * 335360.0, st, &e1);
      if (rc) return rc;
      rc = launch_style_concat(ctx, io, cf, st, f16, bias_rays);
      if (rc) return rc;
      if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
      // style module 2 -> stylised rgb (rendering.py:132-142)
      rc = prof_begin(ctx, 2, (double)m * io.S * 2.0 * 614752.0, st, &e1);
      if (rc) return rc;
      rc = launch_style_wild(ctx, io, remap, cf, st, f16, bias_rays);
      if (rc) return rc;
      if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
      if (which == 0) {
        rc = launch_composite(ctx, nullptr, nullptr, rs_c, ts_c, 0, nullptr, 0, m, S, out.rgb_coarse ? out.rgb_coarse + r0 * 3 : nullptr,
                              out.depth_coarse ? out.depth_coarse + r0 : nullptr, out.acc_coarse ? out.acc_coarse + r0 : nullptr, w_c, st);
        if (rc) return rc;
        rc = launch_sample_fine(ctx, nullptr, nullptr, ts_c, 0, w_c, m, S, F, nullptr, ts_f, nullptr, nullptr, st);
        if (rc) return rc;
      } else {
        rc = launch_composite(ctx, nullptr, nullptr, rs_f, ts_f, T, nullptr, 0, m, T, out.rgb ? out.rgb + r0 * 3 : nullptr,
                              out.depth ? out.depth + r0 : nullptr, out.acc ? out.acc + r0 : nullptr,
                              out.weights ? out.weights + r0 * T : nullptr, st);
        if (rc) return rc;
      }
    }
  }
  if (nsets > 1) {   // join: the caller's stream continues after the side stream's passes
    TGTC_CUDA(cudaEventRecord(ctx->aux_join, ctx->aux_stream));
    TGTC_CUDA(cudaStreamWaitEvent(st0, ctx->aux_join, 0));
  }
  return TGTC_OK;
}

extern "C" int tgtc_render_style(tgtc_ctx* ctx, int mode, const float* rays_o, const float* rays_d, int64_t n_rays, double near, double far,
                                 int n_samples, int n_fine, int64_t chunk, const float* latent1, const float* latent2,
                                 const tgtc_render_out* out_p, void* workspace, size_t workspace_bytes, tgtc_stream stream) {
  return render_style_impl(ctx, mode, rays_o, rays_d, n_rays, near, far, n_samples, n_fine, chunk, latent1, latent2, nullptr, out_p, workspace,
                           workspace_bytes, stream);
}

// the same loop body with PER-RAY latents inside one call (rendering.py:125-127: latents_model_1 returns one row per ray)
extern "C" size_t tgtc_render_style_rays_workspace_bytes(int64_t n_rays, int n_samples, int n_fine, int64_t chunk) {
  if (n_rays <= 0 || n_samples <= 0 || n_fine < 0) return 0;
  const int64_t pass = style_pass(n_rays, chunk);
  return style_ws_set_bytes(style_ws_layout(pass, n_samples, n_fine, true)) * style_ws_sets(n_rays, pass);
}
extern "C" int tgtc_render_style_rays(tgtc_ctx* ctx, int mode, const float* rays_o, const float* rays_d, int64_t n_rays, double near, double far,
                                      int n_samples, int n_fine, int64_t chunk, const float* latents, const tgtc_render_out* out_p,
                                      void* workspace, size_t workspace_bytes, tgtc_stream stream) {
  if (n_rays > 0 && latents == nullptr) { tgtc_set_error("latents is null"); return TGTC_ERR_ARG; }
  return render_style_impl(ctx, mode, rays_o, rays_d, n_rays, near, far, n_samples, n_fine, chunk, nullptr, nullptr, latents, out_p, workspace,
                           workspace_bytes, stream);
}

// ---------------------------------------------------------------------------
// stage entries of the per-ray style head on explicit features: what the reference's injected callables concat_style_forward /
// style_forward compute (rendering.py:129-140), for one (style, frame) latent per call
extern "C" size_t tgtc_style_stage_workspace_bytes(int64_t n_samples_total) {
  return n_samples_total > 0 ? style_stage_workspace_bytes(n_samples_total) : 0;
}

static int style_stage_prologue(tgtc_ctx* ctx, int mode, int64_t M, const float* x, const float* latent, void* ws, size_t ws_bytes) {
  TGTC_REQUIRE(ctx->style.set, TGTC_ERR_STATE, "style weights not set (tgtc_set_style_weights)");
  TGTC_REQUIRE(mode == TGTC_MLP_BF16 || mode == TGTC_MLP_F16, TGTC_ERR_ARG, "style stages run on the tensor-core path: mode must be BF16 or F16 (got %d)", mode);
  TGTC_REQUIRE(M >= 0, TGTC_ERR_ARG, "bad sample count %lld", (long long)M);
  if (M == 0) return TGTC_OK;
  CHECK_PTR(x, "x"); CHECK_PTR(latent, "latent");
  TGTC_REQUIRE(ws != nullptr && (reinterpret_cast<uintptr_t>(ws) & 1023) == 0 && ws_bytes >= style_stage_workspace_bytes(M), TGTC_ERR_STATE,
               "style stage workspace too small or not 1024-byte aligned: need %zu bytes, got %zu", style_stage_workspace_bytes(M), ws_bytes);
  return TGTC_OK;
}

extern "C" int tgtc_style_concat_forward(tgtc_ctx* ctx, int mode, const float* x, const float* latent, int64_t n_samples_total,
                                         float* concat_features, void* workspace, size_t workspace_bytes, tgtc_stream stream) {
  CHECK_CTX(ctx);
  int rc = style_stage_prologue(ctx, mode, n_samples_total, x, latent, workspace, workspace_bytes);
  if (rc || n_samples_total == 0) return rc;
  CHECK_PTR(concat_features, "concat_features");
  cudaStream_t st = (cudaStream_t)stream;
  rc = style_set_latents(ctx, latent, latent, st);     // module 1's effective biases (module 2's are not used by this call)
  if (rc) return rc;
  return launch_style_concat_explicit(ctx, x, n_samples_total, concat_features, static_cast<uint8_t*>(workspace), mode == TGTC_MLP_F16, st);
}

extern "C" int tgtc_style_forward(tgtc_ctx* ctx, int mode, const float* x, const float* concated, const float* latent,
                                  int64_t n_samples_total, float* rgb, void* workspace, size_t workspace_bytes, tgtc_stream stream) {
  CHECK_CTX(ctx);
  int rc = style_stage_prologue(ctx, mode, n_samples_total, x, latent, workspace, workspace_bytes);
  if (rc || n_samples_total == 0) return rc;
  CHECK_PTR(concated, "concated"); CHECK_PTR(rgb, "rgb");
  cudaStream_t st = (cudaStream_t)stream;
  rc = style_set_latents(ctx, latent, latent, st);     // module 2's effective biases from ITS latent argument
  if (rc) return rc;
  return launch_style_wild_explicit(ctx, x, concated, n_samples_total, rgb, static_cast<uint8_t*>(workspace), mode == TGTC_MLP_F16, st);
}

// ---------------------------------------------------------------------------
// Style_train (train_tgtcs.py:311-495): forward with per-ray latents + stash, then the backward into the
// two style modules and the latents.  The loss lives between the two calls (it needs both rgb maps and, for the coherence
// term, the previous batch), so the stash of both passes stays in the caller's workspace.

struct StyleTrainWs {
  size_t off_c[2], off_w[2], off_pe[2], off_mask[2], off_remap[2], off_rs[2], off_ts[2];
  size_t off_w_c, off_bias, off_dz, off_dhead, off_drs, off_g, off_R, off_wlat, off_partial, total;
};
static StyleTrainWs style_train_ws_layout(tgtc_ctx* ctx, int64_t n, int S, int F) {
  StyleTrainWs w;
  size_t o = 0;
  const int Sp[2] = {S, S + F};
  for (int p = 0; p < 2; ++p) {
    const size_t tiles = (size_t)((n * Sp[p] + 127) / 128);
    w.off_c[p] = o;     o = align_up(o + tiles * 5 * 65536, 1024);
    w.off_w[p] = o;     o = align_up(o + tiles * 7 * 65536, 1024);
    w.off_pe[p] = o;    o = align_up(o + tiles * 16384, 1024);
    w.off_mask[p] = o;  o = align_up(o + tiles * kStyleMaskLayers * 8 * 128 * 4, 1024);
    w.off_remap[p] = o; o = align_up(o + tiles * 65536, 1024);
    w.off_rs[p] = o;    o = align_up(o + (size_t)n * Sp[p] * 16, 256);
    w.off_ts[p] = o;    o = align_up(o + (size_t)n * Sp[p] * 4, 256);
  }
  const size_t tiles = (size_t)((n * (S + F) + 127) / 128);
  w.off_w_c = o;     o = align_up(o + (size_t)n * S * 4, 256);
  w.off_bias = o;    o = align_up(o + (size_t)n * 13 * 256 * 4, 256);
  w.off_dz = o;      o = align_up(o + tiles * 12 * 65536, 1024);
  w.off_dhead = o;   o = align_up(o + tiles * 16384, 1024);
  w.off_drs = o;     o = align_up(o + (size_t)n * (S + F) * 16, 256);
  w.off_g = o;       o = align_up(o + (size_t)n * 12, 256);
  w.off_R = o;       o = align_up(o + (size_t)13 * 2 * tiles * 256 * 4, 256);
  w.off_wlat = o;    o = align_up(o + style_wlat_part_floats() * 4, 256);
  w.off_partial = o; o = align_up(o + (size_t)ctx->num_sms * style_partial_floats() * 4, 256);
  w.total = o;
  return w;
}

extern "C" size_t tgtc_style_train_workspace_bytes(tgtc_ctx* ctx, int64_t n_rays, int n_samples, int n_fine) {
  if (ctx == nullptr || n_rays <= 0 || n_samples <= 0 || n_fine < 0) return 0;
  return style_train_ws_layout(ctx, n_rays, n_samples, n_fine).total;
}
extern "C" int64_t tgtc_style_num_params(void) { return (int64_t)style_flat_floats(); }

static StyleStash style_stash_of(const StyleTrainWs& ws, uint8_t* base, int p) {
  StyleStash s;
  s.c = base + ws.off_c[p]; s.w = base + ws.off_w[p]; s.pe = base + ws.off_pe[p];
  s.mask = reinterpret_cast<uint32_t*>(base + ws.off_mask[p]);
  return s;
}

#define STYLE_TRAIN_PROLOGUE()                                                                                                   \
  CHECK_CTX(ctx);                                                                                                               \
  CHECK_NET(ctx, TGTC_NET_COARSE);                                                                                              \
  CHECK_NET(ctx, TGTC_NET_FINE);                                                                                                \
  TGTC_REQUIRE(ctx->style.set, TGTC_ERR_STATE, "style weights not set (tgtc_set_style_weights)");                               \
  TGTC_REQUIRE(n_rays >= 0, TGTC_ERR_ARG, "bad n_rays=%lld", (long long)n_rays);                                                \
  if (n_rays == 0) return TGTC_OK;                                                                                              \
  const int S = n_samples, F = n_fine, T = n_samples + n_fine;                                                                  \
  TGTC_REQUIRE(S == 64 && T == 128, TGTC_ERR_UNSUPPORTED, "style training supports n_samples=64, n_fine=64; got %d+%d", S, F);  \
  const StyleTrainWs ws = style_train_ws_layout(ctx, n_rays, S, F);                                                             \
  TGTC_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 1023) == 0 && workspace_bytes >= ws.total,     \
               TGTC_ERR_STATE, "style training workspace too small or not 1024-byte aligned: need %zu bytes, got %zu", ws.total, \
               workspace_bytes);                                                                                                \
  DeviceGuard g(ctx->device);                                                                                                   \
  cudaStream_t st = (cudaStream_t)stream;                                                                                       \
  uint8_t* base = static_cast<uint8_t*>(workspace)

static int style_train_forward_impl(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, int64_t n_rays, double near, double far,
                                    int n_samples, int n_fine, const float* lat1, const float* rand, const float* noise_coarse,
                                    const float* noise_fine, const PhiloxSrc* jitter, const PhiloxSrc* prng_c, const PhiloxSrc* prng_f,
                                    float* rgb_coarse, float* rgb_fine, void* workspace, size_t workspace_bytes, tgtc_stream stream) {
  STYLE_TRAIN_PROLOGUE();
  CHECK_PTR(rays_o, "rays_o"); CHECK_PTR(rays_d, "rays_d"); CHECK_PTR(lat1, "lat1");
  CHECK_PTR(rgb_coarse, "rgb_coarse"); CHECK_PTR(rgb_fine, "rgb_fine");
  float* bias_rays = reinterpret_cast<float*>(base + ws.off_bias);
  float* w_c = reinterpret_cast<float*>(base + ws.off_w_c);
  int rc = launch_style_bias_rays(ctx, lat1, n_rays, bias_rays, st);
  if (rc) return rc;
  float* ts_c = reinterpret_cast<float*>(base + ws.off_ts[0]);
  float* ts_f = reinterpret_cast<float*>(base + ws.off_ts[1]);
  const bool per_ray_ts = rand != nullptr || jitter != nullptr;
  const int64_t ts_c_stride = per_ray_ts ? S : 0;
  rc = launch_sample_uniform(ctx, nullptr, nullptr, per_ray_ts ? n_rays : 1, S, near, far, rand, nullptr, ts_c, st, jitter);
  if (rc) return rc;
  for (int p = 0; p < 2; ++p) {
    const StyleStash stash = style_stash_of(ws, base, p);
    uint8_t* remap = base + ws.off_remap[p];
    float* rs = reinterpret_cast<float*>(base + ws.off_rs[p]);
    MlpIO io;
    io.rays_o = rays_o; io.rays_d = rays_d;
    io.ts = p == 0 ? (ts_c_stride ? ts_c : nullptr) : ts_f;
    io.t_scale = (float)(far - near); io.t_near = (float)near;
    io.n_rays = n_rays; io.S = p == 0 ? S : T;
    io.rgbsigma = rs;
    const double samples = (double)n_rays * io.S;
    cudaEvent_t e1 = nullptr;
    rc = prof_begin(ctx, 0, samples * 2.0 * (593408.0 - 36224.0 - 384.0), st, &e1);
    if (rc) return rc;
    rc = launch_mlp_tc_trunk(ctx, p, io, remap, st);                       // frozen NeRF net: base_remap images + sigma
    if (rc) return rc;
    if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
    rc = prof_begin(ctx, 1, samples * 2.0 * (335360.0 + 614752.0), st, &e1);
    if (rc) return rc;
    rc = launch_style_concat_train(ctx, io, bias_rays, stash, st);
    if (rc) return rc;
    rc = launch_style_wild_train(ctx, io, bias_rays, remap, stash, st);
    if (rc) return rc;
    if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
    if (p == 0) {
      rc = launch_composite(ctx, nullptr, nullptr, rs, ts_c, ts_c_stride, noise_coarse, 0, n_rays, S, rgb_coarse, nullptr, nullptr, w_c, st,
                            prng_c);
      if (rc) return rc;
      rc = launch_sample_fine(ctx, nullptr, nullptr, ts_c, ts_c_stride, w_c, n_rays, S, F, nullptr, ts_f, nullptr, nullptr, st);
      if (rc) return rc;
    } else {
      rc = launch_composite(ctx, nullptr, nullptr, rs, ts_f, T, noise_fine, 0, n_rays, T, rgb_fine, nullptr, nullptr, nullptr, st, prng_f);
      if (rc) return rc;
    }
  }
  // remember what this workspace holds (at most a handful of pending batches: the oldest record is recycled)
  tgtc_ctx::StyleFwdRec rec = {workspace, n_rays, S, F, per_ray_ts ? 1 : 0};
  bool found = false;
  for (auto& r : ctx->style_fwd)
    if (r.ws == workspace) { r = rec; found = true; }
  if (!found) {
    if (ctx->style_fwd.size() >= 16) ctx->style_fwd.erase(ctx->style_fwd.begin());
    ctx->style_fwd.push_back(rec);
  }
  return TGTC_OK;
}

extern "C" int tgtc_style_train_forward(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, int64_t n_rays, double near, double far,
                                        int n_samples, int n_fine, const float* lat1, const float* rand, const float* noise_coarse,
                                        const float* noise_fine, float* rgb_coarse, float* rgb_fine, void* workspace,
                                        size_t workspace_bytes, tgtc_stream stream) {
  return style_train_forward_impl(ctx, rays_o, rays_d, n_rays, near, far, n_samples, n_fine, lat1, rand, noise_coarse, noise_fine, nullptr,
                                  nullptr, nullptr, rgb_coarse, rgb_fine, workspace, workspace_bytes, stream);
}

// the stochastic options drawn inside the kernels (Philox4x32-10, philox.cuh) instead of read from caller tensors
extern "C" int tgtc_style_train_forward_seeded(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, int64_t n_rays, double near,
                                               double far, int n_samples, int n_fine, const float* lat1, unsigned long long seed,
                                               int perturb, double sigma_noise_std, float* rgb_coarse, float* rgb_fine, void* workspace,
                                               size_t workspace_bytes, tgtc_stream stream) {
  const PhiloxSrc jit = {seed, kStreamJitter, 1.0f, 1};
  const PhiloxSrc nc = {seed, kStreamNoiseCoarse, (float)sigma_noise_std, 1};
  const PhiloxSrc nf = {seed, kStreamNoiseFine, (float)sigma_noise_std, 1};
  const bool noise = sigma_noise_std > 0.0;
  return style_train_forward_impl(ctx, rays_o, rays_d, n_rays, near, far, n_samples, n_fine, lat1, nullptr, nullptr, nullptr,
                                  perturb ? &jit : nullptr, noise ? &nc : nullptr, noise ? &nf : nullptr, rgb_coarse, rgb_fine, workspace,
                                  workspace_bytes, stream);
}

static int style_train_backward_impl(tgtc_ctx* ctx, int64_t n_rays, int n_samples, int n_fine, const float* lat1, int has_rand,
                                     const float* noise_coarse, const float* noise_fine, const PhiloxSrc* prng_c, const PhiloxSrc* prng_f,
                                     const float* d_rgb_coarse, const float* d_rgb_fine, float* grads, int accumulate, float* dlat1,
                                     void* workspace, size_t workspace_bytes, tgtc_stream stream) {
  if (ctx != nullptr && n_rays == 0 && !accumulate && grads != nullptr) {
    // an empty shard still owns a gradient buffer that the caller all-reduces: it must read as zero (no d latent rows exist)
    DeviceGuard g0(ctx->device);
    TGTC_CUDA(cudaMemsetAsync(grads, 0, style_flat_floats() * sizeof(float), (cudaStream_t)stream));
  }
  STYLE_TRAIN_PROLOGUE();
  CHECK_PTR(lat1, "lat1"); CHECK_PTR(d_rgb_coarse, "d_rgb_coarse"); CHECK_PTR(d_rgb_fine, "d_rgb_fine"); CHECK_PTR(grads, "grads");
  {
    const tgtc_ctx::StyleFwdRec* rec = nullptr;
    for (const auto& r : ctx->style_fwd)
      if (r.ws == workspace) rec = &r;
    TGTC_REQUIRE(rec != nullptr, TGTC_ERR_STATE, "tgtc_style_train_backward: no tgtc_style_train_forward has filled this workspace");
    TGTC_REQUIRE(rec->n == n_rays && rec->S == S && rec->F == F && rec->has_rand == (has_rand ? 1 : 0), TGTC_ERR_STATE,
                 "tgtc_style_train_backward: the workspace holds the stash of a different batch (n=%lld, %d+%d samples, has_rand=%d)",
                 (long long)rec->n, rec->S, rec->F, rec->has_rand);
  }
  StyleDz dz;
  dz.dz = base + ws.off_dz; dz.dhead = base + ws.off_dhead;
  float* drs = reinterpret_cast<float*>(base + ws.off_drs);
  float* R = reinterpret_cast<float*>(base + ws.off_R);
  float* partial = reinterpret_cast<float*>(base + ws.off_partial);
  for (int p = 0; p < 2; ++p) {
    const StyleStash stash = style_stash_of(ws, base, p);
    const float* rs = reinterpret_cast<const float*>(base + ws.off_rs[p]);
    const float* ts = reinterpret_cast<const float*>(base + ws.off_ts[p]);
    const int Sp = p == 0 ? S : T;
    const int64_t ts_stride = p == 0 ? (has_rand ? S : 0) : T;
    int rc = launch_composite_backward(ctx, rs, ts, ts_stride, p == 0 ? noise_coarse : noise_fine, 0, n_rays, Sp,
                                       p == 0 ? d_rgb_coarse : d_rgb_fine, nullptr, nullptr, drs, st, p == 0 ? prng_c : prng_f);
    if (rc) return rc;
    const double samples = (double)n_rays * Sp;
    cudaEvent_t e1 = nullptr;
    rc = prof_begin(ctx, 2, samples * 2.0 * (11.0 * 65536.0 + 768.0), st, &e1);   // hidden-to-hidden slices + the head
    if (rc) return rc;
    rc = launch_style_dgrad(ctx, rs, drs, stash, dz, n_rays * Sp, st);
    if (rc) return rc;
    if (e1) TGTC_CUDA(cudaEventRecord(e1, st));
    rc = prof_begin(ctx, 3, samples * 2.0 * (335360.0 + 614752.0), st, &e1);
    if (rc) return rc;
    rc = launch_style_wgrad(ctx, stash, base + ws.off_remap[p], dz, lat1, n_rays, Sp, partial, R,
                            reinterpret_cast<float*>(base + ws.off_wlat), grads, (p == 0) ? accumulate : 1, dlat1,
                            p == 0 ? 0 : 1, st, e1);
    if (rc) return rc;
  }
  return TGTC_OK;
}

extern "C" int tgtc_style_train_backward(tgtc_ctx* ctx, int64_t n_rays, int n_samples, int n_fine, const float* lat1, int has_rand,
                                         const float* noise_coarse, const float* noise_fine, const float* d_rgb_coarse,
                                         const float* d_rgb_fine, float* grads, int accumulate, float* dlat1, void* workspace,
                                         size_t workspace_bytes, tgtc_stream stream) {
  return style_train_backward_impl(ctx, n_rays, n_samples, n_fine, lat1, has_rand, noise_coarse, noise_fine, nullptr, nullptr, d_rgb_coarse,
                                   d_rgb_fine, grads, accumulate, dlat1, workspace, workspace_bytes, stream);
}

extern "C" int tgtc_style_train_backward_seeded(tgtc_ctx* ctx, int64_t n_rays, int n_samples, int n_fine, const float* lat1,
                                                unsigned long long seed, int perturb, double sigma_noise_std, const float* d_rgb_coarse,
                                                const float* d_rgb_fine, float* grads, int accumulate, float* dlat1, void* workspace,
                                                size_t workspace_bytes, tgtc_stream stream) {
  const PhiloxSrc nc = {seed, kStreamNoiseCoarse, (float)sigma_noise_std, 1};
  const PhiloxSrc nf = {seed, kStreamNoiseFine, (float)sigma_noise_std, 1};
  const bool noise = sigma_noise_std > 0.0;
  return style_train_backward_impl(ctx, n_rays, n_samples, n_fine, lat1, perturb ? 1 : 0, nullptr, nullptr, noise ? &nc : nullptr,
                                   noise ? &nf : nullptr, d_rgb_coarse, d_rgb_fine, grads, accumulate, dlat1, workspace, workspace_bytes,
                                   stream);
}

// torch.optim.Adam on flat fp32 buffers (the optimizer of train_tgtcs.py:39)
extern "C" int tgtc_adam_step(tgtc_ctx* ctx, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                              double beta1, double beta2, double eps, int64_t step, tgtc_stream stream) {
  CHECK_CTX(ctx);
  TGTC_REQUIRE(n >= 0 && step >= 1, TGTC_ERR_ARG, "bad n=%lld / step=%lld", (long long)n, (long long)step);
  if (n == 0) return TGTC_OK;
  CHECK_PTR(params, "params"); CHECK_PTR(grads, "grads"); CHECK_PTR(exp_avg, "exp_avg"); CHECK_PTR(exp_avg_sq, "exp_avg_sq");
  DeviceGuard g(ctx->device);
  return launch_adam(ctx, params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, (cudaStream_t)stream);
}
