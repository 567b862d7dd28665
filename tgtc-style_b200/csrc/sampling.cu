// K2 -- stratified sample positions (utils.py:509-531)
// K6+K7 -- inverse-CDF importance resampling + sorted union (utils.py:573-609)
//
// Both are HBM-bound when run stand-alone (sample_fine: read w 4S + write
// ts_fine 4(S+F) bytes per ray = 768 B/ray at S=F=64).  One warp per ray.
#include "common.cuh"
#include "philox.cuh"

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
// K6+K7
constexpr int kMaxS = 128;       // coarse samples per ray
constexpr int kMaxOut = 256;     // S + n_fine, padded to a power of two for the sort
constexpr int kWarpsPerBlock = 4;

struct FineSmem {
  float ts[kMaxS];
  float w[kMaxS];        // weights[1:-1] + 1e-5, later pdf
  float cdf[kMaxS];      // S-1 entries
  float out[kMaxOut];    // union to sort
  float smp[kMaxOut];    // the new inverse-CDF samples (F entries)
};

__device__ __forceinline__ double shfl_up_f64(double v, int delta) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, delta);
  hi = __shfl_up_sync(0xffffffffu, hi, delta);
  return __hiloint2double(hi, lo);
}
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

    // ---- normaliser: torch.sum(-1) in ATen's order (vectorized_inner_sum, 8 lanes, ILP 4)
    const int nvec = nw >> 3, nilp = nvec >> 2;
    float acc0 = 0.f;
    if (lane < 8) {
      float a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int i = 0; i < nilp; ++i) {
        acc0 = __fadd_rn(acc0, sm.w[8 * (4 * i + 0) + lane]);
        a1 = __fadd_rn(a1, sm.w[8 * (4 * i + 1) + lane]);
        a2 = __fadd_rn(a2, sm.w[8 * (4 * i + 2) + lane]);
        a3 = __fadd_rn(a3, sm.w[8 * (4 * i + 3) + lane]);
      }
      for (int j = 4 * nilp; j < nvec; ++j) acc0 = __fadd_rn(acc0, sm.w[8 * j + lane]);
      acc0 = __fadd_rn(acc0, a1);
      acc0 = __fadd_rn(acc0, a2);
      acc0 = __fadd_rn(acc0, a3);
    }
    float fin = 0.f;
    for (int k = 8 * nvec; k < nw; ++k) fin = __fadd_rn(fin, sm.w[k]);   // every lane, same value
#pragma unroll
    for (int l = 0; l < 8; ++l) fin = __fadd_rn(fin, __shfl_sync(0xffffffffu, acc0, l));
    __syncwarp();

    // ---- pdf = w / sum ; cdf = [0, cumsum(pdf)] with an fp64 accumulator rounded per prefix.
    // Every partial sum of these <=126 non-negative fp32 values in [~1e-7, 1] is exactly
    // representable in fp64 (span < 53 bits), so the warp-parallel scan is bit-identical to
    // torch's sequential fp64 accumulation.
    const int per = (nw + 31) >> 5;            // contiguous elements per lane (<= 4: S <= 128)
    const int j0 = lane * per;
    float pdfv[4];
    double run = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + q;
      pdfv[q] = (q < per && j < nw) ? __fdiv_rn(sm.w[j], fin) : 0.f;
      run += (double)pdfv[q];
    }
    double incl = run;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const double o = shfl_up_f64(incl, dlt);
      if (lane >= dlt) incl += o;
    }
    double pre = incl - run;                   // exclusive prefix of this lane (exact)
    if (lane == 0) sm.cdf[0] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + q;
      if (q < per && j < nw) {
        pre += (double)pdfv[q];
        sm.cdf[j + 1] = (float)pre;
      }
    }
    __syncwarp();

    // ---- inverse CDF for u = linspace(0,1,F)
    for (int k = lane; k < F; k += 32) {
      const float u = linspace01(k, F);
      // searchsorted(cdf, u, right=True) = #{cdf <= u}; cdf is non-decreasing
      int lo = 0, hi = nb;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sm.cdf[mid] <= u) lo = mid + 1; else hi = mid;
      }
      const int ind = lo;
      const int below = max(0, ind - 1);
      const int above = min(nb - 1, ind);
      const float cb = sm.cdf[below], ca = sm.cdf[above];
      const float bb = __fmul_rn(0.5f, __fadd_rn(sm.ts[below + 1], sm.ts[below]));
      const float ba = __fmul_rn(0.5f, __fadd_rn(sm.ts[above + 1], sm.ts[above]));
      float denom = __fsub_rn(ca, cb);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = __fdiv_rn(__fsub_rn(u, cb), denom);
      const float s = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
      sm.smp[k] = s;
      if (inds_out != nullptr) inds_out[ray * F + k] = ind;
      if (samples_out != nullptr) samples_out[ray * F + k] = s;
    }
    __syncwarp();
    // ---- torch.sort(cat(ts, t_samples)) (utils.py:577), values only.  ts is ascending by construction; the new
    // samples are non-decreasing except in rare fp32 corner cases.  When they are (warp vote) the union is a rank merge;
    // otherwise fall back to the full bitonic network (log2^2 n steps).
    bool mono = true;
    for (int k = lane; k + 1 < F; k += 32) mono = mono && (sm.smp[k] <= sm.smp[k + 1]);
    for (int i = lane; i + 1 < S; i += 32) mono = mono && (sm.ts[i] <= sm.ts[i + 1]);
    const bool sorted_inputs = __all_sync(0xffffffffu, mono);
    if (sorted_inputs) {
      // both lists ascending: every element's place in the union is its own index plus the number of elements of the OTHER
      // list that precede it (ties: ts first) -- two binary searches per element instead of a sorting network
      for (int i = lane; i < S; i += 32) {
        const float v = sm.ts[i];
        int lo = 0, hi = F;                      // #{samples < v}
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.smp[mid] < v) lo = mid + 1; else hi = mid; }
        sm.out[i + lo] = v;
      }
      for (int k = lane; k < F; k += 32) {
        const float v = sm.smp[k];
        int lo = 0, hi = S;                      // #{ts <= v}
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.ts[mid] <= v) lo = mid + 1; else hi = mid; }
        sm.out[k + lo] = v;
      }
      __syncwarp();
    } else {
      for (int i = lane; i < S; i += 32) sm.out[i] = sm.ts[i];
      for (int k = lane; k < F; k += 32) sm.out[S + k] = sm.smp[k];
      for (int i = total + lane; i < sort_n; i += 32) sm.out[i] = __int_as_float(0x7f800000);  // +inf padding
      __syncwarp();
      for (int k = 2; k <= sort_n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = lane; i < sort_n; i += 32) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const float a = sm.out[i], b = sm.out[ixj];
              const bool up = ((i & k) == 0);
              if ((a > b) == up) { sm.out[i] = b; sm.out[ixj] = a; }
            }
          }
          __syncwarp();
        }
      }
    }

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

int launch_sample_fine(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, const float* ts, int64_t ts_stride,
                       const float* weights, int64_t n, int S, int n_fine, float* pts_out, float* ts_out,
                       int64_t* inds_out, float* samples_out, cudaStream_t st) {
  TGTC_REQUIRE(S >= 10 && S <= kMaxS, TGTC_ERR_UNSUPPORTED, "sample_fine: S=%d outside [10,%d]", S, kMaxS);
  TGTC_REQUIRE(n_fine >= 1 && S + n_fine <= kMaxOut, TGTC_ERR_UNSUPPORTED, "sample_fine: S+n_fine=%d > %d", S + n_fine, kMaxOut);
  int sort_n = 2;
  while (sort_n < S + n_fine) sort_n <<= 1;
  const int64_t blocks_needed = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int64_t cap = (int64_t)ctx->num_sms * 16;   // 16 resident 128-thread CTAs per SM
  const int64_t grid = blocks_needed < cap ? blocks_needed : cap;
  if (S == 64 && n_fine == 64)
    sample_fine_kernel<64, 64><<<(unsigned)grid, 32 * kWarpsPerBlock, 0, st>>>(rays_o, rays_d, ts, ts_stride, weights, n, S, n_fine,
                                                                              sort_n, pts_out, ts_out, inds_out, samples_out);
  else
    sample_fine_kernel<0, 0><<<(unsigned)grid, 32 * kWarpsPerBlock, 0, st>>>(rays_o, rays_d, ts, ts_stride, weights, n, S, n_fine,
                                                                            sort_n, pts_out, ts_out, inds_out, samples_out);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}
