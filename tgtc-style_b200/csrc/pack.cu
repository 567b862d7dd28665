// Weight packer: fp32 nn.Linear tensors ([out,in] row-major, the reference's
// state_dict layout, models.py:75-91) -> the three device images the kernels read.
//   f32_gemm : per GEMM layer, transposed [Kpad][N] fp32        (mlp_fp32.cu)
//   tc_blob  : per GEMM layer, chunks of [N x 64] bf16 in the 128B-swizzled
//              K-major UMMA shared-memory layout, consumption order (mlp_tc.cu)
//   smalls   : biases, sigma head, rgb1 head, dir-PE slice of rgb0 (fp32)
// K padding: PE 63->64 (zero column 63; layer 5 = [PE(64) | hidden(256)]),
// dir-PE 27->32.
#include "common.cuh"

namespace {

struct Params24 { const float* p[TGTC_NUM_PARAMS]; };

// param index of GEMM layer l (0..9): L0..L7 -> 0..7, remap -> 9, rgb0 -> 10
__device__ __forceinline__ int gemm_param(int l) { return l < 8 ? l : l + 1; }

// value of GEMM layer l's weight at (n, k) in the padded-K numbering
__device__ __forceinline__ float gemm_w(const Params24& P, int l, int n, int k, bool with_dir) {
  const float* W = P.p[2 * gemm_param(l)];
  if (l == 0) return k < kPtsEmb ? W[n * kPtsEmb + k] : 0.f;
  if (l == 5) {
    if (k < kPtsEmb) return W[n * 319 + k];
    if (k < kPtsEmbPad) return 0.f;
    return W[n * 319 + kPtsEmb + (k - kPtsEmbPad)];
  }
  if (l == 9) {
    if (k < 256) return W[n * 283 + k];
    if (with_dir && k < 256 + kDirEmb) return W[n * 283 + k];
    return 0.f;
  }
  return W[n * 256 + k];
}

__global__ void pack_f32_kernel(Params24 P, float* __restrict__ out) {
  const size_t total = kF32GemmFloats;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int l = 0;
    size_t off = 0;
    while (l + 1 < kF32NumGemm && idx >= off + (size_t)f32_layer_k(l) * f32_layer_n(l)) {
      off += (size_t)f32_layer_k(l) * f32_layer_n(l);
      ++l;
    }
    const size_t r = idx - off;
    const int N = f32_layer_n(l);
    const int k = (int)(r / N), n = (int)(r % N);
    out[idx] = gemm_w(P, l, n, k, /*with_dir=*/true);
  }
}

template <typename T> __device__ __forceinline__ T to_op(float v);
template <> __device__ __forceinline__ __nv_bfloat16 to_op<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half to_op<__half>(float v) { return __float2half_rn(v); }

template <typename T>
__global__ void pack_tc_kernel(Params24 P, T* __restrict__ out) {
  const size_t total = kTcBlobBytes / 2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int l = 0;
    size_t off = 0;  // in elements
    while (l + 1 < kTcNumGemm && idx >= off + (size_t)tc_layer_k(l) * tc_layer_n(l)) {
      off += (size_t)tc_layer_k(l) * tc_layer_n(l);
      ++l;
    }
    const size_t r = idx - off;
    const int N = tc_layer_n(l);
    const size_t chunk_elems = (size_t)N * kTcChunkK;
    const int chunk = (int)(r / chunk_elems);
    const int e = (int)(r % chunk_elems);       // element offset inside the chunk image
    const int byte = e * 2;
    // invert: byte = (n/8)*1024 + (n%8)*128 + (((kk/8) ^ (n%8)) * 16) + (kk%8)*2
    const int grp = byte >> 10, rem = byte & 1023;
    const int rr = rem >> 7, inrow = rem & 127;
    const int c16p = inrow >> 4, within = (inrow & 15) >> 1;
    const int c16 = c16p ^ rr;
    const int n = grp * 8 + rr;
    const int k = chunk * kTcChunkK + c16 * 8 + within;
    out[idx] = to_op<T>(gemm_w(P, l, n, k, /*with_dir=*/false));
  }
}

// transposed image for the activation-gradient chain (mlp_bwd.cu): GEMM g multiplies dz [128 x K=out features] by
// B[n = input feature][k = output feature] = W[k][n]; chunks of [256 rows x 64 K], 128B-swizzled, in the order
// rgb0 (remap columns; K=128), remap, L7, L6, L5 (hidden columns), L4, L3, L2, L1.
__global__ void pack_tcT_kernel(Params24 P, __nv_bfloat16* __restrict__ out, size_t total) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t e0 = 2 * 256 * 64;                 // elements of GEMM 0 (two chunks)
    int g;
    size_t r;
    if (idx < e0) { g = 0; r = idx; } else { g = 1 + (int)((idx - e0) / (4 * 256 * 64)); r = (idx - e0) % (4 * 256 * 64); }
    const int chunk = (int)(r / (256 * 64));
    const int byte = (int)(r % (256 * 64)) * 2;
    const int grp = byte >> 10, rem = byte & 1023;
    const int rr = rem >> 7, inrow = rem & 127;
    const int c16 = (inrow >> 4) ^ rr, within = (inrow & 15) >> 1;
    const int n = grp * 8 + rr;                     // input feature
    const int k = chunk * 64 + c16 * 8 + within;    // output feature
    float v;
    switch (g) {
      case 0: v = P.p[2 * 10][k * 283 + n]; break;          // rgb0 [128, 283]
      case 1: v = P.p[2 * 9][k * 256 + n]; break;           // remap
      case 2: v = P.p[2 * 7][k * 256 + n]; break;
      case 3: v = P.p[2 * 6][k * 256 + n]; break;
      case 4: v = P.p[2 * 5][k * 319 + 63 + n]; break;      // L5: hidden columns follow the 63 PE columns (models.py:100)
      default: v = P.p[2 * (9 - g)][k * 256 + n]; break;    // g=5..8 -> L4..L1
    }
    out[idx] = __float2bfloat16_rn(v);
  }
}

__global__ void pack_smalls_kernel(Params24 P, float* __restrict__ out) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < kSmallFloats; idx += gridDim.x * blockDim.x) {
    float v;
    if (idx < kSmBiasRgb0) {
      const int j = idx / 256, c = idx % 256;           // L0..L7, remap
      v = P.p[2 * (j < 8 ? j : 9) + 1][c];
    } else if (idx < kSmWSigma) {
      v = P.p[2 * 10 + 1][idx - kSmBiasRgb0];
    } else if (idx < kSmWRgb1) {
      v = P.p[2 * 8][idx - kSmWSigma];
    } else if (idx < kSmBSigma) {
      v = P.p[2 * 11][idx - kSmWRgb1];
    } else if (idx < kSmBRgb1) {
      v = P.p[2 * 8 + 1][0];
    } else if (idx < kSmWDir) {
      v = P.p[2 * 11 + 1][idx - kSmBRgb1];
    } else {
      const int r = idx - kSmWDir;
      const int k = r / 128, n = r % 128;
      v = k < kDirEmb ? P.p[2 * 10][n * 283 + 256 + k] : 0.f;
    }
    out[idx] = v;
  }
}

}  // namespace

int pack_weights(tgtc_ctx* ctx, int net, const float* const* params, cudaStream_t st) {
  NetImage& im = ctx->net[net];
  if (im.f32_gemm == nullptr) {
    TGTC_CUDA(cudaMalloc(&im.f32_gemm, kF32GemmFloats * sizeof(float)));
    TGTC_CUDA(cudaMalloc(&im.smalls, kSmallFloats * sizeof(float)));
    TGTC_CUDA(cudaMalloc(&im.tc_blob, kTcBlobBytes));
    TGTC_CUDA(cudaMalloc(&im.tc_blob_h, kTcBlobBytes));
    TGTC_CUDA(cudaMalloc(&im.tc_blobT, bwd_blobT_bytes()));
  }
  Params24 P;
  for (int i = 0; i < TGTC_NUM_PARAMS; ++i) P.p[i] = params[i];
  pack_f32_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(P, im.f32_gemm);
  TGTC_LAUNCH_CHECK(ctx);
  pack_tc_kernel<__nv_bfloat16><<<ctx->num_sms * 4, 256, 0, st>>>(P, reinterpret_cast<__nv_bfloat16*>(im.tc_blob));
  TGTC_LAUNCH_CHECK(ctx);
  pack_tc_kernel<__half><<<ctx->num_sms * 4, 256, 0, st>>>(P, reinterpret_cast<__half*>(im.tc_blob_h));
  TGTC_LAUNCH_CHECK(ctx);
  pack_tcT_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(P, reinterpret_cast<__nv_bfloat16*>(im.tc_blobT), bwd_blobT_bytes() / 2);
  TGTC_LAUNCH_CHECK(ctx);
  pack_smalls_kernel<<<32, 256, 0, st>>>(P, im.smalls);
  TGTC_LAUNCH_CHECK(ctx);
  im.set = true;
  return TGTC_OK;
}
