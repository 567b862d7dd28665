// K1 -- per-pixel ray generation + NDC warp.
// Replaces dataset.get_rays_np (dataset.py:33-42) and dataset.ndc_rays_np
// (dataset.py:44-61).  The reference does this in NumPy fp64 and the rays are
// then cast to fp32; we do the same fp64 operations in the same order with
// explicit round-to-nearest intrinsics (no FMA contraction), so the fp32 rays
// are bit-identical.  HBM-bound: 24 B/ray written, nothing read.
#include "common.cuh"

namespace {

struct RaygenParams {
  double fx, fy, cx, cy;
  double r[9];     // c2w[:3,:3] row-major
  double t[3];     // c2w[:3,3]
  double near_;    // NDC near plane
  double aw, ah;   // -1/(W/(2f)), -1/(H/(2f))  (python-float scalars of dataset.py:50-55)
  double two_near; // 2*near
  double m2near;   // -2*near
  int W;
  int ndc;
  int pixel_alignment;
  int64_t pix_begin;
  int64_t n;
};

__device__ __forceinline__ void one_ray(const RaygenParams& p, int64_t pix, float* o, float* d) {
  const int64_t row = pix / p.W;
  const int col = (int)(pix - row * p.W);
  double ci = (double)(float)col;  // meshgrid of float32 aranges
  double cj = (double)(float)row;
  if (p.pixel_alignment) { ci = __dadd_rn(ci, 0.5); cj = __dadd_rn(cj, 0.5); }
  // dirs = [(i-cx)/fx, -(j-cy)/fy, -1]
  const double v0 = __ddiv_rn(__dsub_rn(ci, p.cx), p.fx);
  const double v1 = -__ddiv_rn(__dsub_rn(cj, p.cy), p.fy);
  const double v2 = -1.0;
  // rays_d[c] = sum_k dirs[k]*c2w[c,k]   (np.sum over 3 products: ((p0+p1)+p2))
  double dd[3], oo[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double p0 = __dmul_rn(v0, p.r[3 * c + 0]);
    const double p1 = __dmul_rn(v1, p.r[3 * c + 1]);
    const double p2 = __dmul_rn(v2, p.r[3 * c + 2]);
    dd[c] = __dadd_rn(__dadd_rn(p0, p1), p2);
    oo[c] = p.t[c];
  }
  if (p.ndc) {
    // t = -(near + o_z)/d_z ; o = o + t*d
    const double tt = __ddiv_rn(-__dadd_rn(p.near_, oo[2]), dd[2]);
#pragma unroll
    for (int c = 0; c < 3; ++c) oo[c] = __dadd_rn(oo[c], __dmul_rn(tt, dd[c]));
    const double o0 = __ddiv_rn(__dmul_rn(p.aw, oo[0]), oo[2]);
    const double o1 = __ddiv_rn(__dmul_rn(p.ah, oo[1]), oo[2]);
    const double o2 = __dadd_rn(1.0, __ddiv_rn(p.two_near, oo[2]));
    const double d0 = __dmul_rn(p.aw, __dsub_rn(__ddiv_rn(dd[0], dd[2]), __ddiv_rn(oo[0], oo[2])));
    const double d1 = __dmul_rn(p.ah, __dsub_rn(__ddiv_rn(dd[1], dd[2]), __ddiv_rn(oo[1], oo[2])));
    const double d2 = __ddiv_rn(p.m2near, oo[2]);
    oo[0] = o0; oo[1] = o1; oo[2] = o2;
    dd[0] = d0; dd[1] = d1; dd[2] = d2;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) { o[c] = (float)oo[c]; d[c] = (float)dd[c]; }
}

// each thread makes 4 consecutive rays = 12 floats per array = 3 float4 stores
__global__ void __launch_bounds__(256) raygen_kernel(RaygenParams p, float* __restrict__ rays_o, float* __restrict__ rays_d, int vec_ok) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = q * 4;
  if (i0 >= p.n) return;
  float o[12], d[12];
  const int cnt = (int)min((int64_t)4, p.n - i0);
  for (int k = 0; k < cnt; ++k) one_ray(p, p.pix_begin + i0 + k, o + 3 * k, d + 3 * k);
  if (cnt == 4 && vec_ok) {
    float4* po = reinterpret_cast<float4*>(rays_o + 3 * i0);
    float4* pd = reinterpret_cast<float4*>(rays_d + 3 * i0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      po[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
      pd[k] = make_float4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]);
    }
  } else {
    for (int k = 0; k < 3 * cnt; ++k) { rays_o[3 * i0 + k] = o[k]; rays_d[3 * i0 + k] = d[k]; }
  }
}

}  // namespace

int launch_raygen(tgtc_ctx* ctx, int H, int W, const double* K, const double* c2w, int ndc, double ndc_near,
                  int pixel_alignment, int64_t pix_begin, int64_t n, float* rays_o, float* rays_d, cudaStream_t st) {
  RaygenParams p;
  p.fx = K[0]; p.cx = K[2]; p.fy = K[4]; p.cy = K[5];
  for (int c = 0; c < 3; ++c) {
    for (int k = 0; k < 3; ++k) p.r[3 * c + k] = c2w[4 * c + k];
    p.t[c] = c2w[4 * c + 3];
  }
  // dataset.py:50-58 evaluates these scalar sub-expressions in Python floats (fp64);
  // ndc_rays_np is called with focal=K[0][0] (dataset.py:117)
  const double focal = K[0];
  p.near_ = ndc_near;
  p.aw = -1. / (W / (2. * focal));
  p.ah = -1. / (H / (2. * focal));
  p.two_near = 2. * ndc_near;
  p.m2near = -2. * ndc_near;
  p.W = W; p.ndc = ndc; p.pixel_alignment = pixel_alignment;
  p.pix_begin = pix_begin; p.n = n;
  const int64_t quads = (n + 3) / 4;
  const int block = 256;
  const int64_t grid = (quads + block - 1) / block;
  const int vec_ok = aligned16(rays_o) && aligned16(rays_d);
  raygen_kernel<<<(unsigned)grid, block, 0, st>>>(p, rays_o, rays_d, vec_ok);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}
