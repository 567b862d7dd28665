// K5 -- alpha compositing.  Replaces utils.alpha_composition (utils.py:354-386):
//   delta_i = t_{i+1}-t_i, delta_last = 1e10
//   alpha   = 1 - exp(-relu(relu(sigma+noise)) * delta)
//   T_i     = prod_{j<i} (1 - alpha_j + 1e-10)         (exclusive cumprod)
//   w       = alpha*T ; rgb = sum w*c ; depth = sum w*t ; acc = sum w
// One warp per ray; the transmittance product is a warp-shuffle scan over
// 32-sample chunks with a running carry.  Early termination: once the carry
// underflows to exactly 0 every later weight is exactly 0, so the remaining
// chunks are not read (bit-identical to not terminating).
// HBM-bound: fine pass reads 16 B/sample + 4 B/sample ts, writes 4 B/sample + 20 B/ray.
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int kWarps = 8;

template <bool PACKED>
__global__ void __launch_bounds__(32 * kWarps) composite_kernel(
    const float* __restrict__ rgb, const float* __restrict__ sigma, const float4* __restrict__ rgbsigma,
    const float* __restrict__ ts, int64_t ts_stride, const float* __restrict__ noise, const PhiloxSrc prng, int white_bkgd, int64_t n,
    int S, float* __restrict__ rgb_out, float* __restrict__ depth_out, float* __restrict__ acc_out,
    float* __restrict__ weights_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t ray = warp; ray < n; ray += nwarps) {
    const float* tsr = ts + ray * ts_stride;
    float carry = 1.0f;  // product of (1-alpha+1e-10) over all previous chunks
    float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
    for (int base = 0; base < S; base += 32) {
      const int i = base + lane;
      const bool valid = i < S;
      if (carry == 0.0f) {  // exact early termination (warp-uniform)
        if (weights_out != nullptr && valid) weights_out[ray * S + i] = 0.0f;
        continue;
      }
      float sg = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, t = 0.f, tn = 0.f;
      if (valid) {
        if (PACKED) {
          const float4 v = rgbsigma[ray * S + i];
          cr = v.x; cg = v.y; cb = v.z; sg = v.w;
        } else {
          sg = sigma[ray * S + i];
          const float* c = rgb + (ray * S + i) * 3;
          cr = c[0]; cg = c[1]; cb = c[2];
        }
        if (noise != nullptr) sg = __fadd_rn(sg, noise[ray * S + i]);
        else if (prng.on) sg = __fadd_rn(sg, __fmul_rn(prng.std, philox_normal(prng.seed, prng.stream, (uint64_t)(ray * S + i))));
        t = tsr[i];
        tn = (i + 1 < S) ? tsr[i + 1] : 0.f;
      }
      const float delta = (i + 1 < S) ? __fsub_rn(tn, t) : 1e10f;
      const float act = fmaxf(sg, 0.0f);
      float alpha = valid ? __fsub_rn(1.0f, expf(-__fmul_rn(act, delta))) : 0.0f;
      const float fac = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
      // inclusive product scan across the warp
      float incl = fac;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl = __fmul_rn(incl, o);
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = __fmul_rn(carry, excl);
      const float w = __fmul_rn(alpha, T);
      carry = __fmul_rn(carry, __shfl_sync(0xffffffffu, incl, 31));
      if (valid) {
        if (weights_out != nullptr) weights_out[ray * S + i] = w;
        // explicit FMAs: the fused compositing in mlp_tc.cu's last epilogue repeats this exact sequence (bit-identical results)
        ar = __fmaf_rn(w, cr, ar); ag = __fmaf_rn(w, cg, ag); ab = __fmaf_rn(w, cb, ab); ad = __fmaf_rn(w, t, ad); aa = __fadd_rn(aa, w);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      ar += __shfl_xor_sync(0xffffffffu, ar, d);
      ag += __shfl_xor_sync(0xffffffffu, ag, d);
      ab += __shfl_xor_sync(0xffffffffu, ab, d);
      ad += __shfl_xor_sync(0xffffffffu, ad, d);
      aa += __shfl_xor_sync(0xffffffffu, aa, d);
    }
    if (lane == 0) {
      if (white_bkgd) { const float bg = __fsub_rn(1.0f, aa); ar = __fadd_rn(ar, bg); ag = __fadd_rn(ag, bg); ab = __fadd_rn(ab, bg); }
      if (rgb_out != nullptr) { rgb_out[ray * 3 + 0] = ar; rgb_out[ray * 3 + 1] = ag; rgb_out[ray * 3 + 2] = ab; }
      if (depth_out != nullptr) depth_out[ray] = ad;
      if (acc_out != nullptr) acc_out[ray] = aa;
    }
  }
}

}  // namespace

int launch_composite(tgtc_ctx* ctx, const float* rgb, const float* sigma, const float* rgbsigma, const float* ts,
                     int64_t ts_stride, const float* noise, int white_bkgd, int64_t n, int S, float* rgb_out,
                     float* depth_out, float* acc_out, float* weights_out, cudaStream_t st, const PhiloxSrc* prng_p) {
  const PhiloxSrc prng = prng_p != nullptr ? *prng_p : PhiloxSrc{0, 0, 0.f, 0};
  const int64_t blocks_needed = (n + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)ctx->num_sms * 8 * 4;  // 8 resident 256-thread CTAs per SM, 4 waves
  const int64_t grid = blocks_needed < cap ? blocks_needed : cap;
  if (rgbsigma != nullptr) {
    composite_kernel<true><<<(unsigned)grid, 32 * kWarps, 0, st>>>(nullptr, nullptr, reinterpret_cast<const float4*>(rgbsigma), ts,
                                                                  ts_stride, noise, prng, white_bkgd, n, S, rgb_out, depth_out,
                                                                  acc_out, weights_out);
  } else {
    composite_kernel<false><<<(unsigned)grid, 32 * kWarps, 0, st>>>(rgb, sigma, nullptr, ts, ts_stride, noise, prng, white_bkgd, n,
                                                                   S, rgb_out, depth_out, acc_out, weights_out);
  }
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

// ---------------------------------------------------------------------------
// K5 backward -- reverse of alpha_composition for the training step (train_tgtcs.py:236-255: the loss reaches the
// MLP outputs through rgb = sum w*c and through sigma -> alpha -> w).  Given dL/d(rgb_map) [n,3] (and optionally
// dL/d(depth), dL/d(acc)), writes dL/d(c_r, c_g, c_b, sigma) per sample as float4.
//   gw_i      = g . c_i + g_depth * t_i + g_acc            (- sum(g) with white background)
//   dL/dc_i   = w_i * g
//   dL/da_i   = gw_i * T_i - (sum_{k>i} gw_k w_k) / (1 - a_i + 1e-10)
//   dL/dsig_i = dL/da_i * delta_i * exp(-relu(sig_i) * delta_i) * 1[sig_i + noise_i > 0]
// One warp per ray: pass 1 recomputes (e, T, w) exactly as the forward kernel and parks them in shared memory,
// pass 2 runs the suffix sum from the far end.  No gradient flows to t (utils.py:576-579).
namespace {

constexpr int kBwdMaxS = 256;

__global__ void __launch_bounds__(32 * kWarps) composite_backward_kernel(
    const float4* __restrict__ rgbsigma, const float* __restrict__ ts, int64_t ts_stride, const float* __restrict__ noise,
    const PhiloxSrc prng, int white_bkgd, int64_t n, int S, const float* __restrict__ g_rgb, const float* __restrict__ g_depth,
    const float* __restrict__ g_acc, float4* __restrict__ d_rgbsigma) {
  __shared__ float s_e[kWarps][kBwdMaxS], s_T[kWarps][kBwdMaxS], s_gw[kWarps][kBwdMaxS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * kWarps + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  const int nchunk = (S + 31) >> 5;
  for (int64_t ray = warp; ray < n; ray += nwarps) {
    const float* tsr = ts + ray * ts_stride;
    const float gr = g_rgb[ray * 3 + 0], gg = g_rgb[ray * 3 + 1], gb = g_rgb[ray * 3 + 2];
    const float gd = g_depth != nullptr ? g_depth[ray] : 0.f;
    const float ga = (g_acc != nullptr ? g_acc[ray] : 0.f) - (white_bkgd ? (gr + gg + gb) : 0.f);
    // ---- pass 1: forward recompute
    float carry = 1.0f;
    for (int c = 0; c < nchunk; ++c) {
      const int i = c * 32 + lane;
      const bool valid = i < S;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      float t = 0.f, tn = 0.f;
      if (valid) {
        v = rgbsigma[ray * S + i];
        if (noise != nullptr) v.w = __fadd_rn(v.w, noise[ray * S + i]);
        else if (prng.on) v.w = __fadd_rn(v.w, __fmul_rn(prng.std, philox_normal(prng.seed, prng.stream, (uint64_t)(ray * S + i))));
        t = tsr[i];
        tn = (i + 1 < S) ? tsr[i + 1] : 0.f;
      }
      const float delta = (i + 1 < S) ? __fsub_rn(tn, t) : 1e10f;
      const float e = valid ? expf(-__fmul_rn(fmaxf(v.w, 0.0f), delta)) : 1.0f;
      const float alpha = __fsub_rn(1.0f, e);
      const float fac = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
      float incl = fac;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl = __fmul_rn(incl, o);
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = carry * excl;
      carry = carry * __shfl_sync(0xffffffffu, incl, 31);
      if (valid) {
        s_e[wib][i] = e;
        s_T[wib][i] = T;
        s_gw[wib][i] = gr * v.x + gg * v.y + gb * v.z + gd * t + ga;
      }
    }
    __syncwarp();
    // ---- pass 2: suffix sums from the far end
    float tail = 0.f;  // sum_{k > current chunk} gw_k w_k
    for (int c = nchunk - 1; c >= 0; --c) {
      const int i = c * 32 + lane;
      const bool valid = i < S;
      const float e = valid ? s_e[wib][i] : 1.0f;
      const float T = valid ? s_T[wib][i] : 0.f;
      const float gw = valid ? s_gw[wib][i] : 0.f;
      const float alpha = __fsub_rn(1.0f, e);
      const float w = alpha * T;
      const float x = gw * w;
      float incl = x;  // inclusive suffix sum inside the chunk
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_down_sync(0xffffffffu, incl, d);
        if (lane + d < 32) incl += o;
      }
      const float after = tail + (incl - x);
      tail += __shfl_sync(0xffffffffu, incl, 0);
      if (valid) {
        const float4 v = rgbsigma[ray * S + i];
        float sg = v.w;
        if (noise != nullptr) sg = __fadd_rn(sg, noise[ray * S + i]);
        else if (prng.on) sg = __fadd_rn(sg, __fmul_rn(prng.std, philox_normal(prng.seed, prng.stream, (uint64_t)(ray * S + i))));
        const float t = tsr[i];
        const float delta = (i + 1 < S) ? __fsub_rn(tsr[i + 1], t) : 1e10f;
        const float fac = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
        const float dalpha = gw * T - after / fac;
        const float dsig = sg > 0.0f ? dalpha * delta * e : 0.0f;
        d_rgbsigma[ray * S + i] = make_float4(w * gr, w * gg, w * gb, dsig);
      }
    }
    __syncwarp();
  }
}

}  // namespace

int launch_composite_backward(tgtc_ctx* ctx, const float* rgbsigma, const float* ts, int64_t ts_stride, const float* noise,
                              int white_bkgd, int64_t n, int S, const float* g_rgb, const float* g_depth, const float* g_acc,
                              float* d_rgbsigma, cudaStream_t st, const PhiloxSrc* prng_p) {
  const PhiloxSrc prng = prng_p != nullptr ? *prng_p : PhiloxSrc{0, 0, 0.f, 0};
  TGTC_REQUIRE(S <= kBwdMaxS, TGTC_ERR_UNSUPPORTED, "composite_backward: S=%d > %d", S, kBwdMaxS);
  const int64_t blocks_needed = (n + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)ctx->num_sms * 8 * 4;
  const int64_t grid = blocks_needed < cap ? blocks_needed : cap;
  composite_backward_kernel<<<(unsigned)grid, 32 * kWarps, 0, st>>>(reinterpret_cast<const float4*>(rgbsigma), ts, ts_stride, noise, prng,
                                                                   white_bkgd, n, S, g_rgb, g_depth, g_acc,
                                                                   reinterpret_cast<float4*>(d_rgbsigma));
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

// ---------------------------------------------------------------------------
// the same streams as stand-alone tensors (tests, and callers that want to look at what a seeded step drew)
namespace {
__global__ void philox_fill_kernel(unsigned long long seed, uint32_t stream, int normal, float std, int64_t n, float* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = normal ? __fmul_rn(std, philox_normal(seed, stream, (uint64_t)e)) : philox_uniform(seed, stream, (uint64_t)e);
}
}  // namespace
int launch_philox_fill(tgtc_ctx* ctx, unsigned long long seed, uint32_t stream, int normal, float std, int64_t n, float* out, cudaStream_t st) {
  if (n == 0) return TGTC_OK;
  const int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  philox_fill_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(seed, stream, normal, std, n, out);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}
