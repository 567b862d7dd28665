// K2 -- stratified sample positions (utils.py:509-531)
// K6+K7 -- inverse-CDF importance resampling + sorted union (utils.py:573-609)
//
// Both are HBM-bound when run stand-alone (sample_fine: read w 4S + write
// ts_fine 4(S+F) bytes per ray = 768 B/ray at S=F=64).  One warp per ray.
#include "common.cuh"
#include "philox.cuh"
#include "sample_fine.cuh"

namespace {

// ---------------------------------------------------------------------------
// K2
// utils.py:513-516: ts = linspace*(far-near)+near, or with harmony=True ts = 1/(1/near*(1-ts) + 1/far*ts) (fp32 tensor ops with
// the Python-float scalars 1/near, 1/far rounded to fp32)
struct TRow { float t_scale, t_near, inv_near, inv_far; int harmony; };
__device__ __forceinline__ float row_t(int k, int S, const TRow& r) {
  if (!r.harmony) return coarse_t(k, S, r.t_scale, r.t_near);
  const float u = linspace01(k, S);
  return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(r.inv_near, __fsub_rn(1.0f, u)), __fmul_rn(r.inv_far, u)));
}

__global__ void __launch_bounds__(256) sample_uniform_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                             int64_t n, int S, const TRow row,
                                                             const float* __restrict__ rnd, const PhiloxSrc prng,
                                                             float* __restrict__ pts, float* __restrict__ ts) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * S) return;
  const int64_t ray = idx / S;
  const int k = (int)(idx - ray * S);
  float t = row_t(k, S, row);
  if (rnd != nullptr || prng.on) {
    // utils.py:518-524: mid=(ts[1:]+ts[:-1])/2; upper=[mid, ts[-1]]; lower=[ts[0], mid]; ts=lower+(upper-lower)*rand
    const float tp = row_t(k > 0 ? k - 1 : 0, S, row);
    const float tn = row_t(k < S - 1 ? k + 1 : S - 1, S, row);
    const float lower = (k == 0) ? t : __fdiv_rn(__fadd_rn(t, tp), 2.0f);
    const float upper = (k == S - 1) ? t : __fdiv_rn(__fadd_rn(tn, t), 2.0f);
    const float u = rnd != nullptr ? rnd[idx] : philox_uniform(prng.seed, prng.stream, (uint64_t)idx);
    t = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));
  }
  ts[idx] = t;
  if (pts != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c)  // pts = rays_o + ts*rays_d : separate mul and add
      pts[idx * 3 + c] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(t, rays_d[ray * 3 + c]));
  }
}

// ---------------------------------------------------------------------------
// K6+K7 (the per-ray arithmetic lives in sample_fine.cuh, shared with the fused form inside the coarse MLP kernel)
constexpr int kWarpsPerBlock = 4;

// kS / kF: compile-time sample counts (0 = runtime S_rt / F_rt).  The 64 + 64 instantiation is the reference's configuration
// (configs/fern.txt:16-17): constant trip counts unroll every per-ray loop (the kernel is instruction-bound), same arithmetic.
template <int kS, int kF>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) sample_fine_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ ts_in,
    int64_t ts_stride, const float* __restrict__ weights, int64_t n, int S_rt, int F_rt, int sort_n,
    float* __restrict__ pts_out, float* __restrict__ ts_out, int64_t* __restrict__ inds_out,
    float* __restrict__ samples_out) {
  const int S = kS ? kS : S_rt;
  const int F = kF ? kF : F_rt;
  __shared__ FineSmem smem[kWarpsPerBlock];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  FineSmem& sm = smem[wib];
  const int nw = S - 2;    // pdf entries
  const int nb = S - 1;    // bins (midpoints) == cdf entries
  const int total = S + F;

  for (int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + wib; ray < n; ray += (int64_t)gridDim.x * kWarpsPerBlock) {
    // ---- stage the ray: ts[S], w[j] = weights[j+1] + 1e-5
    for (int i = lane; i < S; i += 32) sm.ts[i] = ts_in[ray * ts_stride + i];
    for (int i = lane; i < nw; i += 32) sm.w[i] = __fadd_rn(weights[ray * S + 1 + i], 1e-5f);
    __syncwarp();

    sample_fine_core<kS, kF>(sm, lane, S, F, sort_n, inds_out != nullptr ? inds_out + ray * F : nullptr,
                             samples_out != nullptr ? samples_out + ray * F : nullptr);

    // ---- write ts_fine (coalesced) and optionally pts = o + d*t
    for (int i = lane; i < total; i += 32) ts_out[ray * total + i] = sm.out[i];
    if (pts_out != nullptr) {
      const float ox = rays_o[ray * 3 + 0], oy = rays_o[ray * 3 + 1], oz = rays_o[ray * 3 + 2];
      const float dx = rays_d[ray * 3 + 0], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
      for (int i = lane; i < total; i += 32) {
        const float t = sm.out[i];
        float* p = pts_out + (ray * total + i) * 3;
        p[0] = __fadd_rn(ox, __fmul_rn(dx, t));
        p[1] = __fadd_rn(oy, __fmul_rn(dy, t));
        p[2] = __fadd_rn(oz, __fmul_rn(dz, t));
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// K6+K7, 64 + 64 samples, ts_fine only (the render / training configuration): the same arithmetic as sample_fine_core -- same
// fp32 summation order, same exact fp64 prefix, same searchsorted(right=True), same interpolation -- with the per-ray instruction
// count cut by (a) keeping w / pdf / cdf in registers (two contiguous elements per lane) and doing the ATen-order sum with
// shuffles, (b) fixed-trip branch-free searches, and (c) no searches at all for the sorted union: a new sample lies between the
// bin midpoints below/above it, so the number of coarse positions <= it is below+1 (+1, +2) -- checked against the neighbours,
// with the full bitonic sort as the fallback for anything unexpected -- and the union is written by a 128-bit occupancy mask:
// slot p holds a new sample if its bit is set, otherwise the (p - popc(bits below p))-th coarse position.
constexpr int kFastWarps = 4;   // one ray per warp, no ray loop and no early exit (index clamped, stores predicated): straight-line code, so
                                // ptxas sees every shuffle / vote / reduction as convergent (inside a grid-stride loop over a thread-derived
                                // ray index each collective cost a WARPSYNC + ENDCOLLECTIVE pair and a duplicated shuffle)
__global__ void __launch_bounds__(32 * kFastWarps, 12) sample_fine64_kernel(const float* __restrict__ ts_in, int64_t ts_stride,
                                                                        const float* __restrict__ weights, int64_t n,
                                                                        float* __restrict__ ts_out) {
  __shared__ __align__(16) float s_ts[kFastWarps][64];
  __shared__ __align__(16) float s_cdf[kFastWarps][64];
  __shared__ __align__(16) float s_out[kFastWarps][128];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* const ts = s_ts[wib];
  float* const cdf = s_cdf[wib];
  float* const out = s_out[wib];
  const int64_t ray_raw = (int64_t)blockIdx.x * kFastWarps + wib;
  const bool live = ray_raw < n;                 // a padding warp of the last CTA recomputes the last ray and stores nothing
  const int64_t ray = live ? ray_raw : n - 1;
  {
    // ---- this lane's two coarse positions and raw weights (elements 2*lane, 2*lane+1)
    const float2 t2 = *reinterpret_cast<const float2*>(ts_in + ray * ts_stride + 2 * lane);
    const float2 x2 = *reinterpret_cast<const float2*>(weights + ray * 64 + 2 * lane);
    *reinterpret_cast<float2*>(ts + 2 * lane) = t2;
    __syncwarp();
    float r[4];
    sample_fine64_lean_core(ts, cdf, out, x2, t2, lane, r);
    if (live) {
#pragma unroll
      for (int w = 0; w < 4; ++w) ts_out[ray * 128 + lane + 32 * w] = r[w];
    }
  }
}

}  // namespace

int launch_sample_uniform(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, int64_t n, int S, double near,
                          double far, const float* rnd, float* pts, float* ts, cudaStream_t st, const PhiloxSrc* prng, int harmony) {
  const int64_t total = n * S;
  const int block = 256;
  const int64_t grid = (total + block - 1) / block;
  TGTC_REQUIRE(grid <= 0x7fffffff, TGTC_ERR_UNSUPPORTED, "sample_uniform: too many samples in one call");
  TRow row = {(float)(far - near), (float)near, 0.f, 0.f, harmony ? 1 : 0};
  if (harmony) {
    TGTC_REQUIRE(near != 0.0 && far != 0.0, TGTC_ERR_ARG, "harmony sampling divides by near and far (utils.py:516): both must be non-zero");
    row.inv_near = (float)(1.0 / near);
    row.inv_far = (float)(1.0 / far);
  }
  sample_uniform_kernel<<<(unsigned)grid, block, 0, st>>>(rays_o, rays_d, n, S, row, rnd,
                                                          prng != nullptr ? *prng : PhiloxSrc{0, 0, 0.f, 0}, pts, ts);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

// tests / A-B: route every call through the general kernel
static bool g_force_general_sample_fine = false;
extern "C" void tgtc_debug_sample_fine_general(int on) { g_force_general_sample_fine = on != 0; }

int launch_sample_fine(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, const float* ts, int64_t ts_stride,
                       const float* weights, int64_t n, int S, int n_fine, float* pts_out, float* ts_out,
                       int64_t* inds_out, float* samples_out, cudaStream_t st) {
  TGTC_REQUIRE(S >= 10 && S <= kFineMaxS, TGTC_ERR_UNSUPPORTED, "sample_fine: S=%d outside [10,%d]", S, kFineMaxS);
  TGTC_REQUIRE(n_fine >= 1 && S + n_fine <= kFineMaxOut, TGTC_ERR_UNSUPPORTED, "sample_fine: S+n_fine=%d > %d", S + n_fine, kFineMaxOut);
  int sort_n = 2;
  while (sort_n < S + n_fine) sort_n <<= 1;
  const int64_t blocks_needed = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int64_t cap = (int64_t)ctx->num_sms * 16;   // 16 resident 128-thread CTAs per SM
  const int64_t grid = blocks_needed < cap ? blocks_needed : cap;
  // the lean 64 + 64 kernel when only ts_fine is asked for and the rows are 8-byte aligned (float2 loads)
  if (S == 64 && n_fine == 64 && pts_out == nullptr && inds_out == nullptr && samples_out == nullptr && (ts_stride % 2) == 0 &&
      (reinterpret_cast<uintptr_t>(ts) & 7) == 0 && (reinterpret_cast<uintptr_t>(weights) & 7) == 0 && !g_force_general_sample_fine) {
    const int64_t fb = (n + kFastWarps - 1) / kFastWarps;   // one ray per warp
    TGTC_REQUIRE(fb <= 0x7fffffff, TGTC_ERR_UNSUPPORTED, "sample_fine: too many rays in one call");
    if (n > 0) sample_fine64_kernel<<<(unsigned)fb, 32 * kFastWarps, 0, st>>>(ts, ts_stride, weights, n, ts_out);
    TGTC_LAUNCH_CHECK(ctx);
    return TGTC_OK;
  }
  if (S == 64 && n_fine == 64)
    sample_fine_kernel<64, 64><<<(unsigned)grid, 32 * kWarpsPerBlock, 0, st>>>(rays_o, rays_d, ts, ts_stride, weights, n, S, n_fine,
                                                                              sort_n, pts_out, ts_out, inds_out, samples_out);
  else
    sample_fine_kernel<0, 0><<<(unsigned)grid, 32 * kWarpsPerBlock, 0, st>>>(rays_o, rays_d, ts, ts_stride, weights, n, S, n_fine,
                                                                            sort_n, pts_out, ts_out, inds_out, samples_out);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}
