// The per-ray losses of Style_train (train_tgtcs.py:397-404, :425, :449-459, :480-484) on the composited [N,3] maps and their
// gradients, in two small kernels (the host all-reduces two scalars in between when the batch is sharded over ranks):
//   loss_rgb = lambda_rgb * (img2mse(rgb_coarse, gt) + img2mse(rgb_fine, gt))                          utils.py:460
//   loss_coh = L2_norm(cos(c2, x) - cos(org2, x_org)) + L2_norm(cos(f2, y) - cos(org2, org2))         utils.py:459, VGGNet.py:204-210
// with cos(a, b) = sum_k a_k b_k / ((|a| + 1e-8)(|b| + 1e-8)) per row and L2_norm(v) = sqrt(sum v^2 + 1e-8).
// x, y, x_org are the previous coherence batch's maps (constants); the second reference similarity uses THIS batch's originals
// because train_tgtcs.py:403 has already replaced x_origin when :456 runs.
// At the reference's batch sizes (256..1024 rays) the iteration is bound by host launch overhead: these two launches replace
// about forty small torch kernels and an autograd pass.
#include "common.cuh"

namespace {

struct Row3 { float x, y, z; };
__device__ __forceinline__ Row3 ld3(const float* p, int64_t i) { return {p[i * 3 + 0], p[i * 3 + 1], p[i * 3 + 2]}; }
__device__ __forceinline__ float norm3(Row3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ float cos_rows(Row3 a, Row3 b) {
  const float ia = 1.0f / (norm3(a) + 1e-8f), ib = 1.0f / (norm3(b) + 1e-8f);
  return (a.x * b.x + a.y * b.y + a.z * b.z) * ia * ib;
}
// d cos(a, b) / d a  (b constant); torch's norm has a zero subgradient at a = 0
__device__ __forceinline__ Row3 dcos_da(Row3 a, Row3 b) {
  const float na = norm3(a), nb = norm3(b);
  const float ia = 1.0f / (na + 1e-8f), ib = 1.0f / (nb + 1e-8f);
  const Row3 bn = {b.x * ib, b.y * ib, b.z * ib};
  const float dot = a.x * bn.x + a.y * bn.y + a.z * bn.z;
  const float k = na > 0.f ? dot * ia * ia / na : 0.f;
  return {bn.x * ia - k * a.x, bn.y * ia - k * a.y, bn.z * ia - k * a.z};
}

struct LossArgs {
  const float* rgb_c; const float* rgb_f; const float* gt;                 // [N,3]
  const float* c2; const float* f2; const float* x; const float* y;        // [N2,3] or nullptr (no coherence term)
  const float* org2; const float* x_org;
  int64_t n, n2;
};

// sums[0] = sum (rgb_c - gt)^2, sums[1] = sum (rgb_f - gt)^2, sums[2] / sums[3] = sum of squared similarity differences (coarse / fine)
__global__ void __launch_bounds__(1024) style_loss_sums_kernel(const LossArgs A, float* __restrict__ sums) {
  __shared__ float red[4][32];
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t i = threadIdx.x; i < A.n; i += blockDim.x) {
    const Row3 g = ld3(A.gt, i), c = ld3(A.rgb_c, i), f = ld3(A.rgb_f, i);
    s[0] += (c.x - g.x) * (c.x - g.x) + (c.y - g.y) * (c.y - g.y) + (c.z - g.z) * (c.z - g.z);
    s[1] += (f.x - g.x) * (f.x - g.x) + (f.y - g.y) * (f.y - g.y) + (f.z - g.z) * (f.z - g.z);
  }
  if (A.c2 != nullptr) {
    for (int64_t i = threadIdx.x; i < A.n2; i += blockDim.x) {
      const Row3 o = ld3(A.org2, i);
      const float vc = cos_rows(ld3(A.c2, i), ld3(A.x, i)) - cos_rows(o, ld3(A.x_org, i));
      const float vf = cos_rows(ld3(A.f2, i), ld3(A.y, i)) - cos_rows(o, o);
      s[2] += vc * vc;
      s[3] += vf * vf;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], d);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];   // fixed order
    sums[threadIdx.x] = t;
  }
}

// gradients of  scale_rgb * (sums0 + sums1)  +  scale_coh * (sqrt(ss_c + 1e-8) + sqrt(ss_f + 1e-8))  w.r.t. the four maps, with
// ss_c / ss_f = coh_ss[0] / coh_ss[1] (the sums over ALL ranks' rows)
__global__ void style_loss_grads_kernel(const LossArgs A, const float* __restrict__ coh_ss, float scale_rgb, float scale_coh,
                                        float* __restrict__ d_c, float* __restrict__ d_f, float* __restrict__ d_c2,
                                        float* __restrict__ d_f2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < A.n) {
    const Row3 g = ld3(A.gt, i), c = ld3(A.rgb_c, i), f = ld3(A.rgb_f, i);
    const float k = 2.0f * scale_rgb;
    d_c[i * 3 + 0] = k * (c.x - g.x); d_c[i * 3 + 1] = k * (c.y - g.y); d_c[i * 3 + 2] = k * (c.z - g.z);
    d_f[i * 3 + 0] = k * (f.x - g.x); d_f[i * 3 + 1] = k * (f.y - g.y); d_f[i * 3 + 2] = k * (f.z - g.z);
  }
  if (A.c2 != nullptr && d_c2 != nullptr && i < A.n2) {
    const Row3 o = ld3(A.org2, i);
    const Row3 c2 = ld3(A.c2, i), f2 = ld3(A.f2, i), x = ld3(A.x, i), y = ld3(A.y, i);
    const float vc = cos_rows(c2, x) - cos_rows(o, ld3(A.x_org, i));
    const float vf = cos_rows(f2, y) - cos_rows(o, o);
    const float kc = scale_coh * vc / sqrtf(coh_ss[0] + 1e-8f), kf = scale_coh * vf / sqrtf(coh_ss[1] + 1e-8f);
    const Row3 gc = dcos_da(c2, x), gf = dcos_da(f2, y);
    d_c2[i * 3 + 0] = kc * gc.x; d_c2[i * 3 + 1] = kc * gc.y; d_c2[i * 3 + 2] = kc * gc.z;
    d_f2[i * 3 + 0] = kf * gf.x; d_f2[i * 3 + 1] = kf * gf.y; d_f2[i * 3 + 2] = kf * gf.z;
  }
}


// ---------------------------------------------------------------------------
// models.StyleLatents_variational (models.py:475-549) as two kernels: the per-ray latents
//   lat_i = mu[s_i] + sigma_scale * (table[(s_i * frame_num + f_i) mod rows] - mu[s_i])                     models.py:490-506
// (the modulo is the reference's 7x tiling of the LLFF table, models.py:496: table_tiles = 7; 1 for other dataset types), minus_logp (models.py:531-537)
//   logp = mean_i sum_k (lat_ik - mu_k)^2 / (exp(0.5 logvar_k) + 1e-3)
// and the gradient of  (upstream d lat) + logp_scale * logp_sum  w.r.t. the table, reduced row by row in ray order (no atomics).
struct LatArgs {
  const float* table; const float* mu; const float* logvar;   // [rows,32], [style_num,32] x 2
  const int64_t* sid; const int64_t* fid;                      // [n]
  int64_t n, n_logp;                                           // rays; the first n_logp of them enter minus_logp
  int rows, frame_num;
  int64_t limit;                                               // rows * table_tiles: flat ids at or beyond it are out of range
  float sigma_scale;
};
// the reference indexes `latents.reshape(-1, 32).repeat((7, 1))` for dataset_type == 'llff' and the plain table otherwise
// (models.py:495-498): a flat id wraps modulo `rows` inside the tiled range, and is an IndexError beyond it -- here a trap
__device__ __forceinline__ int lat_row(const LatArgs& A, int64_t i) {
  const int64_t flat = A.sid[i] * A.frame_num + A.fid[i];
  if (flat < 0 || flat >= A.limit) {
    printf("tgtc style latents: id %lld of ray %lld outside the table (%lld rows incl. tiling)\n", (long long)flat, (long long)i, (long long)A.limit);
    __trap();
  }
  return (int)(flat % A.rows);
}

__global__ void __launch_bounds__(1024) style_latents_forward_kernel(const LatArgs A, float* __restrict__ lat, float* __restrict__ logp_sum) {
  __shared__ float red[32];
  const int k = threadIdx.x & 31;
  float acc = 0.f;
  for (int64_t i = threadIdx.x >> 5; i < A.n; i += blockDim.x >> 5) {
    const int64_t s = A.sid[i];
    const float m = A.mu[s * 32 + k];
    const float v = m + A.sigma_scale * (A.table[(size_t)lat_row(A, i) * 32 + k] - m);
    lat[i * 32 + k] = v;
    if (i < A.n_logp) acc += (v - m) * (v - m) / (expf(0.5f * A.logvar[s * 32 + k]) + 1e-3f);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (k == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    *logp_sum = t;
  }
}

// one block per table row: grad[row][k] = sigma_scale * sum_{i: row(i) == row} (dlat[i][k] + [i < n_logp] logp_scale * d logp_sum / d lat_ik)
__global__ void __launch_bounds__(256) style_latents_backward_kernel(const LatArgs A, const float* __restrict__ dlat, float logp_scale,
                                                                     float* __restrict__ grad, int accumulate) {
  __shared__ float red[8][32];
  const int row = blockIdx.x;
  const int k = threadIdx.x & 31, g = threadIdx.x >> 5;
  float acc = 0.f;
  for (int64_t i0 = 0; i0 < A.n; i0 += 8) {       // the 8 warps take consecutive rays; each warp's partial is summed in ray order
    const int64_t i = i0 + g;
    if (i < A.n && lat_row(A, i) == row) {
      float v = dlat != nullptr ? dlat[i * 32 + k] : 0.f;
      if (i < A.n_logp) {
        const int64_t s = A.sid[i];
        const float m = A.mu[s * 32 + k];
        const float d = A.sigma_scale * (A.table[(size_t)row * 32 + k] - m);          // lat - mu
        v += logp_scale * 2.0f * d / (expf(0.5f * A.logvar[s * 32 + k]) + 1e-3f);
      }
      acc += v;
    }
  }
  red[g][k] = acc;
  __syncthreads();
  if (g == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][k];
    t *= A.sigma_scale;
    float* dst = grad + (size_t)row * 32 + k;
    *dst = accumulate ? *dst + t : t;
  }
}

}  // namespace

namespace {
struct DevGuard {   // launches go to the context's device whatever the caller's current device is
  int prev = -1, dev;
  explicit DevGuard(int d) : dev(d) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
  ~DevGuard() { if (prev != dev && prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

static LossArgs loss_args(const float* rgb_c, const float* rgb_f, const float* gt, int64_t n, const float* c2, const float* f2, const float* x,
                          const float* y, const float* org2, const float* x_org, int64_t n2) {
  LossArgs A;
  A.rgb_c = rgb_c; A.rgb_f = rgb_f; A.gt = gt; A.n = n;
  A.c2 = c2; A.f2 = f2; A.x = x; A.y = y; A.org2 = org2; A.x_org = x_org; A.n2 = c2 != nullptr ? n2 : 0;
  return A;
}

extern "C" int tgtc_style_loss_sums(tgtc_ctx* ctx, const float* rgb_coarse, const float* rgb_fine, const float* rgb_gt, int64_t n,
                                    const float* coh_coarse, const float* coh_fine, const float* prev_coarse, const float* prev_fine,
                                    const float* rgb_origin, const float* prev_origin, int64_t n_coh, float* sums, tgtc_stream stream) {
  if (ctx == nullptr) { tgtc_set_error("ctx is null"); return TGTC_ERR_ARG; }
  TGTC_REQUIRE(n > 0 && n_coh >= 0 && rgb_coarse && rgb_fine && rgb_gt && sums, TGTC_ERR_ARG, "tgtc_style_loss_sums: null argument or n <= 0");
  TGTC_REQUIRE(coh_coarse == nullptr || (coh_fine && prev_coarse && prev_fine && rgb_origin && prev_origin && n_coh > 0), TGTC_ERR_ARG,
               "tgtc_style_loss_sums: the coherence term needs all six maps");
  DevGuard guard(ctx->device);
  style_loss_sums_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
      loss_args(rgb_coarse, rgb_fine, rgb_gt, n, coh_coarse, coh_fine, prev_coarse, prev_fine, rgb_origin, prev_origin, n_coh), sums);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

extern "C" int tgtc_style_loss_grads(tgtc_ctx* ctx, const float* rgb_coarse, const float* rgb_fine, const float* rgb_gt, int64_t n,
                                     const float* coh_coarse, const float* coh_fine, const float* prev_coarse, const float* prev_fine,
                                     const float* rgb_origin, const float* prev_origin, int64_t n_coh, const float* coh_ss,
                                     double scale_rgb, double scale_coh, float* d_rgb_coarse, float* d_rgb_fine, float* d_coh_coarse,
                                     float* d_coh_fine, tgtc_stream stream) {
  if (ctx == nullptr) { tgtc_set_error("ctx is null"); return TGTC_ERR_ARG; }
  TGTC_REQUIRE(n > 0 && rgb_coarse && rgb_fine && rgb_gt && d_rgb_coarse && d_rgb_fine, TGTC_ERR_ARG,
               "tgtc_style_loss_grads: null argument or n <= 0");
  TGTC_REQUIRE(coh_coarse == nullptr || (coh_fine && prev_coarse && prev_fine && rgb_origin && prev_origin && coh_ss && d_coh_coarse &&
                                         d_coh_fine && n_coh > 0),
               TGTC_ERR_ARG, "tgtc_style_loss_grads: the coherence term needs all six maps, coh_ss and both outputs");
  DevGuard guard(ctx->device);
  const int64_t m = n > n_coh ? n : n_coh;
  style_loss_grads_kernel<<<(unsigned)((m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      loss_args(rgb_coarse, rgb_fine, rgb_gt, n, coh_coarse, coh_fine, prev_coarse, prev_fine, rgb_origin, prev_origin, n_coh), coh_ss,
      (float)scale_rgb, (float)scale_coh, d_rgb_coarse, d_rgb_fine, d_coh_coarse, d_coh_fine);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

static int lat_args(tgtc_ctx* ctx, LatArgs& A, const float* table, const float* mu, const float* logvar, const int64_t* style_id,
                    const int64_t* frame_id, int64_t n, int64_t n_logp, int rows, int frame_num, int table_tiles, double sigma_scale) {
  TGTC_REQUIRE(table && mu && logvar && style_id && frame_id, TGTC_ERR_ARG, "style latents: null argument");
  TGTC_REQUIRE(n > 0 && n_logp >= 0 && n_logp <= n && rows > 0 && frame_num > 0 && table_tiles >= 1, TGTC_ERR_ARG,
               "style latents: bad n=%lld / n_logp=%lld / rows=%d / table_tiles=%d", (long long)n, (long long)n_logp, rows, table_tiles);
  (void)ctx;
  A.table = table; A.mu = mu; A.logvar = logvar; A.sid = style_id; A.fid = frame_id;
  A.n = n; A.n_logp = n_logp; A.rows = rows; A.frame_num = frame_num; A.limit = (int64_t)rows * table_tiles; A.sigma_scale = (float)sigma_scale;
  return TGTC_OK;
}

extern "C" int tgtc_style_latents_forward(tgtc_ctx* ctx, const float* table, const float* mu, const float* logvar, const int64_t* style_id,
                                          const int64_t* frame_id, int64_t n, int64_t n_logp, int rows, int frame_num, int table_tiles, double sigma_scale,
                                          float* lat, float* logp_sum, tgtc_stream stream) {
  if (ctx == nullptr) { tgtc_set_error("ctx is null"); return TGTC_ERR_ARG; }
  LatArgs A;
  int rc = lat_args(ctx, A, table, mu, logvar, style_id, frame_id, n, n_logp, rows, frame_num, table_tiles, sigma_scale);
  if (rc) return rc;
  TGTC_REQUIRE(lat && logp_sum, TGTC_ERR_ARG, "tgtc_style_latents_forward: null output");
  DevGuard guard(ctx->device);
  style_latents_forward_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(A, lat, logp_sum);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

extern "C" int tgtc_style_latents_backward(tgtc_ctx* ctx, const float* table, const float* mu, const float* logvar, const int64_t* style_id,
                                           const int64_t* frame_id, int64_t n, int64_t n_logp, int rows, int frame_num, int table_tiles, double sigma_scale,
                                           const float* dlat, double logp_scale, float* table_grad, int accumulate, tgtc_stream stream) {
  if (ctx == nullptr) { tgtc_set_error("ctx is null"); return TGTC_ERR_ARG; }
  LatArgs A;
  int rc = lat_args(ctx, A, table, mu, logvar, style_id, frame_id, n, n_logp, rows, frame_num, table_tiles, sigma_scale);
  if (rc) return rc;
  TGTC_REQUIRE(table_grad != nullptr, TGTC_ERR_ARG, "tgtc_style_latents_backward: null output");
  DevGuard guard(ctx->device);
  style_latents_backward_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(A, dlat, (float)logp_scale, table_grad, accumulate);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}
