// K6+K7 core -- utils.sample_pdf (utils.py:583-609, det=True) + the sorted union of utils.sampling_pts_fine_torch (utils.py:577)
// for ONE ray by ONE warp, on a shared-memory scratch the caller has filled with the ray's coarse ts and w = weights[1:-1] + 1e-5.
// Shared by the stand-alone sample_fine_kernel (sampling.cu) and by the input-producer warps of the coarse MLP kernel
// (mlp_tc.cu: fused K6/K7), so both produce the same bits.  Bit-deciding orders: torch.sum(-1) in ATen's
// vectorised order, the cdf as an fp64 prefix rounded per element, searchsorted(right=True).
#pragma once
#include "common.cuh"

constexpr int kFineMaxS = 128;       // coarse samples per ray
constexpr int kFineMaxOut = 256;     // S + n_fine, padded to a power of two for the sort

struct FineSmem {                    // general scratch (runtime S <= 128, S + F <= 256)
  float ts[kFineMaxS];
  float w[kFineMaxS];        // weights[1:-1] + 1e-5, later pdf
  float cdf[kFineMaxS];      // S-1 entries
  float out[kFineMaxOut];    // union to sort
  float smp[kFineMaxOut];    // the new inverse-CDF samples (F entries)
};
struct FineSmem64 {                  // compact scratch of the 64 + 64 configuration (1 536 bytes)
  float ts[64];
  float w[64];
  float cdf[64];
  float out[128];
  float smp[64];
};

__device__ __forceinline__ double fine_shfl_up_f64(double v, int delta) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, delta);
  hi = __shfl_up_sync(0xffffffffu, hi, delta);
  return __hiloint2double(hi, lo);
}

// the rare fallback of the sorted union: out of line and not unrolled, so that it does not sit in the instruction stream of the
// kernels that inline sample_fine_core (the fine MLP kernel's producer warps among them)
static __device__ __noinline__ void fine_bitonic_sort(float* out, int sort_n, int lane) {
#pragma unroll 1
  for (int k = 2; k <= sort_n; k <<= 1) {
#pragma unroll 1
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll 1
      for (int i = lane; i < sort_n; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = out[i], b = out[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { out[i] = b; out[ixj] = a; }
        }
      }
      __syncwarp();
    }
  }
}

// in: sm.ts[0..S), sm.w[0..S-2) filled and visible to the warp.  out: sm.out[0..S+F) = sorted union, sm.smp[0..F) = the new
// samples; inds_row / samples_row (may be nullptr): this ray's rows of the optional outputs.
template <int kS, int kF, typename Smem>
__device__ __forceinline__ void sample_fine_core(Smem& sm, const int lane, const int S_rt, const int F_rt, const int sort_n,
                                                 int64_t* __restrict__ inds_row, float* __restrict__ samples_row) {
  const int S = kS ? kS : S_rt;
  const int F = kF ? kF : F_rt;
  const int nw = S - 2;    // pdf entries
  const int nb = S - 1;    // bins (midpoints) == cdf entries
  const int total = S + F;
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
    const double o = fine_shfl_up_f64(incl, dlt);
    if (lane >= dlt) incl += o;
  }
  double pre = incl - run;                   // exclusive prefix of this lane (exact)
  __syncwarp();                              // every lane holds its pdf values: sm.cdf may alias sm.w (fused form)
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
    if (inds_row != nullptr) inds_row[k] = ind;
    if (samples_row != nullptr) samples_row[k] = s;
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
    fine_bitonic_sort(sm.out, sort_n, lane);
  }

}
