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

// ---------------------------------------------------------------------------
// The lean form of the same per-ray arithmetic for 64 + 64 samples (see sample_fine64_kernel in sampling.cu for the idea): one
// warp, two contiguous elements per lane in registers, fixed-trip searches, the sorted union through a 128-bit occupancy mask.
// ts: the ray's 64 coarse positions in shared memory (visible to the warp); cdf (64 floats) / out (128 floats): scratch;
// x2 / t2: this lane's raw weights and coarse positions 2*lane, 2*lane+1.  r[w] = ts_fine[lane + 32 w].  Bit-identical to
// sample_fine_core<64, 64> (tests/test_gpu_stages.py).
__device__ __forceinline__ void sample_fine64_lean_core(const float* ts, float* cdf, float* out, const float2 x2, const float2 t2,
                                                        const int lane, float (&r)[4]) {
  const unsigned full = 0xffffffffu;
  {
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
  }
}
