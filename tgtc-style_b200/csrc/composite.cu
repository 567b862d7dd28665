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

namespace {

constexpr int kWarps = 8;

template <bool PACKED>
__global__ void __launch_bounds__(32 * kWarps) composite_kernel(
    const float* __restrict__ rgb, const float* __restrict__ sigma, const float4* __restrict__ rgbsigma,
    const float* __restrict__ ts, int64_t ts_stride, const float* __restrict__ noise, int white_bkgd, int64_t n, int S,
    float* __restrict__ rgb_out, float* __restrict__ depth_out, float* __restrict__ acc_out,
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
        if (lane >= d) incl *= o;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = carry * excl;
      const float w = alpha * T;
      carry = carry * __shfl_sync(0xffffffffu, incl, 31);
      if (valid) {
        if (weights_out != nullptr) weights_out[ray * S + i] = w;
        ar += w * cr; ag += w * cg; ab += w * cb; ad += w * t; aa += w;
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
      if (white_bkgd) { const float bg = 1.0f - aa; ar += bg; ag += bg; ab += bg; }
      if (rgb_out != nullptr) { rgb_out[ray * 3 + 0] = ar; rgb_out[ray * 3 + 1] = ag; rgb_out[ray * 3 + 2] = ab; }
      if (depth_out != nullptr) depth_out[ray] = ad;
      if (acc_out != nullptr) acc_out[ray] = aa;
    }
  }
}

}  // namespace

int launch_composite(tgtc_ctx* ctx, const float* rgb, const float* sigma, const float* rgbsigma, const float* ts,
                     int64_t ts_stride, const float* noise, int white_bkgd, int64_t n, int S, float* rgb_out,
                     float* depth_out, float* acc_out, float* weights_out, cudaStream_t st) {
  const int64_t blocks_needed = (n + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)ctx->num_sms * 8 * 4;  // 8 resident 256-thread CTAs per SM, 4 waves
  const int64_t grid = blocks_needed < cap ? blocks_needed : cap;
  if (rgbsigma != nullptr) {
    composite_kernel<true><<<(unsigned)grid, 32 * kWarps, 0, st>>>(nullptr, nullptr, reinterpret_cast<const float4*>(rgbsigma), ts,
                                                                  ts_stride, noise, white_bkgd, n, S, rgb_out, depth_out,
                                                                  acc_out, weights_out);
  } else {
    composite_kernel<false><<<(unsigned)grid, 32 * kWarps, 0, st>>>(rgb, sigma, nullptr, ts, ts_stride, noise, white_bkgd, n, S,
                                                                   rgb_out, depth_out, acc_out, weights_out);
  }
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}
