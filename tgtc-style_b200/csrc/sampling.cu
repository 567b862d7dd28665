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
  const unsigned full = 0xffffffffu;
  const int64_t ray_raw = (int64_t)blockIdx.x * kFastWarps + wib;
  const bool live = ray_raw < n;                 // a padding warp of the last CTA recomputes the last ray and stores nothing
  const int64_t ray = live ? ray_raw : n - 1;
  {
    // ---- this lane's two coarse positions and raw weights (elements 2*lane, 2*lane+1)
    const float2 t2 = *reinterpret_cast<const float2*>(ts_in + ray * ts_stride + 2 * lane);
    const float2 x2 = *reinterpret_cast<const float2*>(weights + ray * 64 + 2 * lane);
    *reinterpret_cast<float2*>(ts + 2 * lane) = t2;
    // w[m] = weights[m + 1] + 1e-5 for m = 0..61 lives at raw index r = m + 1: lane r >> 1, slot r & 1
    const float x0 = __fadd_rn(x2.x, 1e-5f), x1 = __fadd_rn(x2.y, 1e-5f);
    // ---- torch.sum(-1) in ATen's order (8 lanes, 4-way ILP over the first 4 vectors, then vectors 4..6, then a1 a2 a3):
    // lane k < 8 needs w[k + off] for off = 0, 32, 40, 48, 8, 16, 24, i.e. raw k + 1 + off in lane ((k+1) >> 1) + off/2
    const int src0 = (lane + 1) >> 1;
    const bool odd = (lane + 1) & 1;
    auto w_at = [&](int off) {
      const float a = __shfl_sync(full, x0, src0 + (off >> 1)), b = __shfl_sync(full, x1, src0 + (off >> 1));
      return odd ? b : a;
    };
    float acc = w_at(0);
    acc = __fadd_rn(acc, w_at(32));
    acc = __fadd_rn(acc, w_at(40));
    acc = __fadd_rn(acc, w_at(48));
    acc = __fadd_rn(acc, w_at(8));
    acc = __fadd_rn(acc, w_at(16));
    acc = __fadd_rn(acc, w_at(24));
    // tail w[56..61] = raw 57..62, then the eight lane sums in lane order
    float fin = __shfl_sync(full, x1, 28);
    fin = __fadd_rn(fin, __shfl_sync(full, x0, 29));
    fin = __fadd_rn(fin, __shfl_sync(full, x1, 29));
    fin = __fadd_rn(fin, __shfl_sync(full, x0, 30));
    fin = __fadd_rn(fin, __shfl_sync(full, x1, 30));
    fin = __fadd_rn(fin, __shfl_sync(full, x0, 31));
#pragma unroll
    for (int l = 0; l < 8; ++l) fin = __fadd_rn(fin, __shfl_sync(full, acc, l));
    // ---- pdf, exact fp64 prefix (every partial sum is representable: any order gives torch's bits), cdf[r] for raw r = 0..62
    const float d0 = __fdiv_rn(x0, fin), d1 = __fdiv_rn(x1, fin);
    const float p0 = lane >= 1 ? d0 : 0.f;      // raw 2*lane     (raw 0 is not part of the pdf)
    const float p1 = lane <= 30 ? d1 : 0.f;     // raw 2*lane + 1 (raw 63 is not part of the pdf)
    const double run = (double)p0 + (double)p1;
    double incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double o = fine_shfl_up_f64(incl, d);
      if (lane >= d) incl += o;
    }
    const double c0 = (incl - run) + (double)p0;
    const double c1 = c0 + (double)p1;
    *reinterpret_cast<float2*>(cdf + 2 * lane) = make_float2((float)c0, (float)c1);   // cdf[0] = 0; cdf[63] is never read
    __syncwarp();
    // ---- inverse CDF at u_k, k = 2*lane + e: ind = #{cdf[0..62] <= u} by a fixed-trip branch-free search
    float smp[2];
    int cnt[2];          // #{coarse positions <= sample}
    bool ok = true;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = 2 * lane + e;
      const float u = linspace01(k, 64);
      int pos = 0;   // pos + step <= 63 at every step of 32, 16, ..., 1: no bound check, no branch
#pragma unroll
      for (int step = 32; step > 0; step >>= 1) pos += (cdf[pos + step - 1] <= u) ? step : 0;
      const int ind = pos;                       // >= 1: cdf[0] = 0 <= u
      const int below = max(0, ind - 1);
      const int above = min(62, ind);
      const float cb = cdf[below], ca = cdf[above];
      const float tb0 = ts[below], tb1 = ts[below + 1], ta0 = ts[above], ta1 = ts[above + 1];
      const float bb = __fmul_rn(0.5f, __fadd_rn(tb1, tb0));
      const float ba = __fmul_rn(0.5f, __fadd_rn(ta1, ta0));
      float denom = __fsub_rn(ca, cb);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = __fdiv_rn(__fsub_rn(u, cb), denom);
      const float sv = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
      smp[e] = sv;
      // coarse positions <= sv: everything up to `below` (ts[below] <= its midpoint <= sv), then the next one or two by comparison
      const bool in1 = tb1 <= sv, in2 = in1 && above != below && ta1 <= sv;
      const int c = below + 1 + (in1 ? 1 : 0) + (in2 ? 1 : 0);
      cnt[e] = c;
      // the count must be exact for the mask merge: ts[c-1] <= sv < ts[c]
      const float tlo = ts[c - 1], thi = ts[min(c, 63)];
      ok = ok & (tlo <= sv) & ((c >= 64) | (thi > sv));
    }
    // both lists ascending?  (the new samples are, except in rare fp32 corner cases; the coarse positions by construction)
    {
      const float s_next = __shfl_down_sync(full, smp[0], 1), t_next = __shfl_down_sync(full, t2.x, 1);
      ok = ok & (smp[0] <= smp[1]) & ((lane == 31) | (smp[1] <= s_next));
      ok = ok & (t2.x <= t2.y) & ((lane == 31) | (t2.y <= t_next));
    }
    float r[4];
    if (__all_sync(full, ok)) {
      // ---- sorted union through a 128-bit occupancy mask: sample k goes to slot k + cnt_k
      const int q0 = 2 * lane + cnt[0], q1 = 2 * lane + 1 + cnt[1];
      out[q0] = smp[0];
      out[q1] = smp[1];
      unsigned m[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const unsigned mine = ((q0 >> 5) == w ? 1u << (q0 & 31) : 0u) | ((q1 >> 5) == w ? 1u << (q1 & 31) : 0u);
        m[w] = __reduce_or_sync(full, mine);
      }
      __syncwarp();
      int before = 0;      // new samples in the words below
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int pslot = lane + 32 * w;
        const bool taken = (m[w] >> lane) & 1u;
        const int i = pslot - (before + __popc(m[w] & ((1u << lane) - 1u)));   // coarse index for a free slot
        const float* src = taken ? out + pslot : ts + min(i, 63);     // one load from the selected row
        r[w] = *src;
        before += __popc(m[w]);
      }
    } else {
      // anything unexpected: torch.sort of the concatenation (values only) by the bitonic network, as sample_fine_core does
      out[2 * lane] = t2.x; out[2 * lane + 1] = t2.y;
      out[64 + 2 * lane] = smp[0]; out[64 + 2 * lane + 1] = smp[1];
      __syncwarp();
      fine_bitonic_sort(out, 128, lane);
#pragma unroll
      for (int w = 0; w < 4; ++w) r[w] = out[lane + 32 * w];
    }
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
