// K3+K4, fp32 mode -- positional encoding + the 8x256 NeRF MLP on CUDA cores.
// Replaces models.Embedder.forward (models.py:46-60), models.MLP_style.forward
// (models.py:95-117) and models.StyleNerf.forward (models.py:216-223) with fp32
// FFMA arithmetic (the reference-grade path: <=1e-3 end to end, also the
// general path: any S, per-sample view directions, optional feature outputs).
//
// Persistent CTAs, one 64-sample tile at a time.  Activations stay in shared
// memory across all 12 layers (row-major, padded strides); weights stream from
// L2 as [16 x N] fp32 chunks of the pre-transposed image through a 2-stage
// cp.async ring.  256 threads, each owning an 8x8 (or 8x4) register tile.
#include "common.cuh"

namespace {

constexpr int kTile = 64;
constexpr int kThreads = 256;
constexpr int kStrPE = 68;    // 64 + 4 pad (keeps float4 alignment, spreads rows over banks)
constexpr int kStrH = 260;    // 256 + 4
constexpr int kStrD = 36;     // 32 + 4
constexpr int kChunkK = 16;

struct SmemF32 {
  float pe[kTile * kStrPE];
  float ha[kTile * kStrH];
  float hb[kTile * kStrH];
  float dpe[kTile * kStrD];
  float wbuf[2][kChunkK * 256];
  float red[4][kTile][4];
  float p[kTile][3];
  float d[kTile][3];
  float sig[kTile];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// one dense layer: dst[r][0..N) = act(bias + sum_k src(r,k) * Wt[k][n]),  k runs over
// segment A (kA columns of srcA) then segment B (K-kA columns of srcB)
template <int N>
__device__ __forceinline__ void gemm_layer(SmemF32& sm, const float* __restrict__ Wt, int K, const float* srcA, int strA,
                                           int kA, const float* srcB, int strB, float* dst, int strD,
                                           const float* __restrict__ bias, bool relu) {
  constexpr int NC = N / 32;           // columns per thread: 8 (N=256) or 4 (N=128)
  constexpr int F4 = kChunkK * N / 4;  // float4 per chunk
  const int tid = threadIdx.x;
  const int ty = tid >> 5, tx = tid & 31;
  float acc[8][NC];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[i][j] = 0.f;

  const int nchunks = K / kChunkK;
  // prologue: chunk 0
  for (int i = tid; i < F4; i += kThreads) cp_async16(&sm.wbuf[0][i * 4], Wt + (size_t)i * 4);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) {
      const float* g = Wt + (size_t)(c + 1) * kChunkK * N;
      for (int i = tid; i < F4; i += kThreads) cp_async16(&sm.wbuf[buf ^ 1][i * 4], g + (size_t)i * 4);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int k0 = c * kChunkK;
    const float* src;
    int str, kk0;
    if (k0 < kA) { src = srcA; str = strA; kk0 = k0; } else { src = srcB; str = strB; kk0 = k0 - kA; }
    const float* w = sm.wbuf[buf];
#pragma unroll
    for (int k4 = 0; k4 < kChunkK; k4 += 4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(src + (ty * 8 + i) * str + kk0 + k4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float wv[NC];
        const float4 w0 = *reinterpret_cast<const float4*>(w + (k4 + q) * N + tx * 4);
        wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
        if constexpr (NC == 8) {
          const float4 w1 = *reinterpret_cast<const float4*>(w + (k4 + q) * N + 128 + tx * 4);
          wv[NC - 4] = w1.x; wv[NC - 3] = w1.y; wv[NC - 2] = w1.z; wv[NC - 1] = w1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float av = q == 0 ? a[i].x : (q == 1 ? a[i].y : (q == 2 ? a[i].z : a[i].w));
#pragma unroll
          for (int j = 0; j < NC; ++j) acc[i][j] = fmaf(av, wv[j], acc[i][j]);
        }
      }
    }
    __syncthreads();  // everyone done with wbuf[buf] before it is refilled
  }
  // epilogue: bias, activation, store row-major
  float b[NC];
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = bias[tx * 4 + j];
  if constexpr (NC == 8) {
#pragma unroll
    for (int j = 0; j < 4; ++j) b[4 + j] = bias[128 + tx * 4 + j];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      v[j] = acc[i][j] + b[j];
      if (relu) v[j] = fmaxf(v[j], 0.f);
    }
    float* drow = dst + (ty * 8 + i) * strD;
    *reinterpret_cast<float4*>(drow + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
    if constexpr (NC == 8) *reinterpret_cast<float4*>(drow + 128 + tx * 4) = make_float4(v[NC - 4], v[NC - 3], v[NC - 2], v[NC - 1]);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) mlp_fp32_kernel(const float* __restrict__ gemm, const float* __restrict__ smalls,
                                                             MlpIO io, int64_t M) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SmemF32& sm = *reinterpret_cast<SmemF32*>(smem_raw);
  const int tid = threadIdx.x;
  const int64_t ntiles = (M + kTile - 1) / kTile;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t m0 = tile * kTile;
    // ---- per-sample inputs
    if (tid < kTile) {
      const int64_t m = m0 + tid;
      float p[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
      if (m < M) {
        const int64_t ray = m / io.S;
        if (io.rays_o != nullptr) {
          const int k = (int)(m - ray * io.S);
          const float t = io.ts != nullptr ? io.ts[ray * io.S + k] : coarse_t(k, io.S, io.t_scale, io.t_near);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            d[c] = io.rays_d[ray * 3 + c];
            p[c] = __fadd_rn(io.rays_o[ray * 3 + c], __fmul_rn(t, d[c]));
          }
        } else {
          const int64_t dr = io.dirs_per_ray ? ray : m;
#pragma unroll
          for (int c = 0; c < 3; ++c) { p[c] = io.pts[m * 3 + c]; d[c] = io.dirs[dr * 3 + c]; }
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) { sm.p[tid][c] = p[c]; sm.d[tid][c] = d[c]; }
    }
    __syncthreads();
    // ---- positional encoding (models.py:46-60): [x, sin(2^k x), cos(2^k x)]_k, precise sinf/cosf
    {
      const int row = tid & 63, q = tid >> 6;
      float* pe = sm.pe + row * kStrPE;
      float* de = sm.dpe + row * kStrD;
      const float x[3] = {sm.p[row][0], sm.p[row][1], sm.p[row][2]};
      const float v[3] = {sm.d[row][0], sm.d[row][1], sm.d[row][2]};
      if (q == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { pe[c] = x[c]; de[c] = v[c]; }
        pe[63] = 0.f;
#pragma unroll
        for (int c = kDirEmb; c < kDirEmbPad; ++c) de[c] = 0.f;
      }
      for (int f = q; f < 10; f += 4) {
        const float fr = (float)(1 << f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float s, co;
          sincosf(__fmul_rn(x[c], fr), &s, &co);
          pe[3 + 6 * f + c] = s;
          pe[3 + 6 * f + 3 + c] = co;
        }
      }
      {
        const int f = q;  // 4 dir frequencies, one per quarter
        const float fr = (float)(1 << f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float s, co;
          sincosf(__fmul_rn(v[c], fr), &s, &co);
          de[3 + 6 * f + c] = s;
          de[3 + 6 * f + 3 + c] = co;
        }
      }
    }
    __syncthreads();

    const float* bias = smalls + kSmBias;
    // L0: PE -> ha
    gemm_layer<256>(sm, gemm + f32_layer_off(0), 64, sm.pe, kStrPE, 64, nullptr, 0, sm.ha, kStrH, bias + 0 * 256, true);
    gemm_layer<256>(sm, gemm + f32_layer_off(1), 256, sm.ha, kStrH, 256, nullptr, 0, sm.hb, kStrH, bias + 1 * 256, true);
    gemm_layer<256>(sm, gemm + f32_layer_off(2), 256, sm.hb, kStrH, 256, nullptr, 0, sm.ha, kStrH, bias + 2 * 256, true);
    gemm_layer<256>(sm, gemm + f32_layer_off(3), 256, sm.ha, kStrH, 256, nullptr, 0, sm.hb, kStrH, bias + 3 * 256, true);
    gemm_layer<256>(sm, gemm + f32_layer_off(4), 256, sm.hb, kStrH, 256, nullptr, 0, sm.ha, kStrH, bias + 4 * 256, true);
    // L5: cat(PE, h) -> hb   (models.py:100)
    gemm_layer<256>(sm, gemm + f32_layer_off(5), 320, sm.pe, kStrPE, 64, sm.ha, kStrH, sm.hb, kStrH, bias + 5 * 256, true);
    gemm_layer<256>(sm, gemm + f32_layer_off(6), 256, sm.hb, kStrH, 256, nullptr, 0, sm.ha, kStrH, bias + 6 * 256, true);
    gemm_layer<256>(sm, gemm + f32_layer_off(7), 256, sm.ha, kStrH, 256, nullptr, 0, sm.hb, kStrH, bias + 7 * 256, true);
    // sigma head from h7 (hb): no activation (models.py:103)
    {
      const int row = tid & 63, part = tid >> 6;
      const float* h = sm.hb + row * kStrH + part * 64;
      const float* w = smalls + kSmWSigma + part * 64;
      float s = 0.f;
#pragma unroll 4
      for (int k = 0; k < 64; k += 4) {
        const float4 a = *reinterpret_cast<const float4*>(h + k);
        s = fmaf(a.x, w[k], s); s = fmaf(a.y, w[k + 1], s); s = fmaf(a.z, w[k + 2], s); s = fmaf(a.w, w[k + 3], s);
      }
      sm.red[part][row][0] = s;
    }
    __syncthreads();
    if (tid < kTile) sm.sig[tid] = ((sm.red[0][tid][0] + sm.red[1][tid][0]) + (sm.red[2][tid][0] + sm.red[3][tid][0])) + smalls[kSmBSigma];
    // remap: hb -> ha   (models.py:106)
    gemm_layer<256>(sm, gemm + f32_layer_off(8), 256, sm.hb, kStrH, 256, nullptr, 0, sm.ha, kStrH, bias + 8 * 256, true);
    // rgb0: cat(remap, dirPE) -> hb[:, 0:128]   (models.py:108)
    gemm_layer<128>(sm, gemm + f32_layer_off(9), 288, sm.ha, kStrH, 256, sm.dpe, kStrD, sm.hb, kStrH, smalls + kSmBiasRgb0, true);
    // rgb1 + sigmoid (models.py:111)
    {
      const int row = tid & 63, part = tid >> 6;
      const float* h = sm.hb + row * kStrH + part * 32;
      const float* w = smalls + kSmWRgb1 + part * 32;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        const float a = h[k];
        s0 = fmaf(a, w[k], s0); s1 = fmaf(a, w[128 + k], s1); s2 = fmaf(a, w[256 + k], s2);
      }
      sm.red[part][row][0] = s0; sm.red[part][row][1] = s1; sm.red[part][row][2] = s2;
    }
    __syncthreads();
    if (tid < kTile && m0 + tid < M) {
      const int64_t m = m0 + tid;
      float c[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float z = ((sm.red[0][tid][j] + sm.red[1][tid][j]) + (sm.red[2][tid][j] + sm.red[3][tid][j])) + smalls[kSmBRgb1 + j];
        c[j] = 1.0f / (1.0f + expf(-z));
      }
      if (io.rgbsigma != nullptr) {
        reinterpret_cast<float4*>(io.rgbsigma)[m] = make_float4(c[0], c[1], c[2], sm.sig[tid]);
      } else {
        io.rgb[m * 3 + 0] = c[0]; io.rgb[m * 3 + 1] = c[1]; io.rgb[m * 3 + 2] = c[2];
        io.sigma[m] = sm.sig[tid];
      }
    }
    // optional dict entries of the reference (models.py:113-116, :222)
    if (io.base_remap != nullptr) {
      for (int i = tid; i < kTile * 256; i += kThreads) {
        const int r = i >> 8, c = i & 255;
        if (m0 + r < M) io.base_remap[(m0 + r) * 256 + c] = sm.ha[r * kStrH + c];
      }
    }
    if (io.pts_embed != nullptr) {
      for (int i = tid; i < kTile * kPtsEmb; i += kThreads) {
        const int r = i / kPtsEmb, c = i % kPtsEmb;
        if (m0 + r < M) io.pts_embed[(m0 + r) * kPtsEmb + c] = sm.pe[r * kStrPE + c];
      }
    }
    if (io.dirs_embed != nullptr) {
      for (int i = tid; i < kTile * kDirEmb; i += kThreads) {
        const int r = i / kDirEmb, c = i % kDirEmb;
        if (m0 + r < M) io.dirs_embed[(m0 + r) * kDirEmb + c] = sm.dpe[r * kStrD + c];
      }
    }
    __syncthreads();
  }
}

}  // namespace

int launch_mlp_fp32(tgtc_ctx* ctx, int net, const MlpIO& io, cudaStream_t st) {
  const NetImage& im = ctx->net[net];
  const int64_t M = io.n_rays * io.S;
  if (M == 0) return TGTC_OK;
  static bool attr_set[64] = {};
  if (!attr_set[ctx->device & 63]) {
    TGTC_CUDA(cudaFuncSetAttribute(mlp_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemF32)));
    attr_set[ctx->device & 63] = true;
  }
  const int64_t ntiles = (M + kTile - 1) / kTile;
  const int grid = (int)(ntiles < ctx->num_sms ? ntiles : ctx->num_sms);
  mlp_fp32_kernel<<<grid, kThreads, sizeof(SmemF32), st>>>(im.f32_gemm, im.smalls, io, M);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}
