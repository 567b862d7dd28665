// Shared declarations of libtgtc_b200 (internal; the public ABI is include/tgtc_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <vector>

#include "../../include/tgtc_b200.h"
#include "philox.cuh"

// ---------------------------------------------------------------------------
// network geometry (models.py:63-117 with D=8, W=256, skips=[4], L_pts=10, L_dir=4)
constexpr int kHidden = 256;
constexpr int kPtsEmb = 63;      // 3 + 3*2*10
constexpr int kPtsEmbPad = 64;
constexpr int kDirEmb = 27;      // 3 + 3*2*4
constexpr int kDirEmbPad = 32;
constexpr int kRgbHidden = 128;
constexpr int kNumLayers = 12;   // base 0..7, sigma, remap, rgb0, rgb1

// ---- fp32 image (mlp_fp32.cu): every GEMM layer transposed to [Kpad][N] fp32
// order: L0 (64x256), L1..L4 (256x256), L5 (320x256: 64 PE rows then 256 hidden),
//        L6, L7, remap (256x256), rgb0 (288x128: 256 remap rows then 32 dir-PE rows)
constexpr int kF32NumGemm = 10;
__host__ __device__ constexpr int f32_layer_k(int l) { return l == 0 ? 64 : (l == 5 ? 320 : (l == 9 ? 288 : 256)); }
__host__ __device__ constexpr int f32_layer_n(int l) { return l == 9 ? 128 : 256; }
__host__ __device__ constexpr size_t f32_layer_off(int l) {
  size_t o = 0;
  for (int i = 0; i < l; ++i) o += (size_t)f32_layer_k(i) * f32_layer_n(i);
  return o;
}
constexpr size_t kF32GemmFloats = f32_layer_off(kF32NumGemm);

// small fp32 vectors shared by both MLP kernels ("smalls"), offsets in floats
constexpr int kSmBias = 0;                       // 9 x 256: bias of L0..L7, remap
constexpr int kSmBiasRgb0 = 9 * 256;             // 128
constexpr int kSmWSigma = kSmBiasRgb0 + 128;     // 256
constexpr int kSmWRgb1 = kSmWSigma + 256;        // 3 x 128
constexpr int kSmBSigma = kSmWRgb1 + 384;        // 1
constexpr int kSmBRgb1 = kSmBSigma + 1;          // 3
constexpr int kSmWDir = kSmBRgb1 + 3 + 0;        // 32 x 128 (dir-PE rows of rgb0, transposed, rows 27..31 zero)
constexpr int kSmallFloats = kSmWDir + 32 * 128;
static_assert(kSmWDir % 4 == 0, "WDir must stay float4 aligned");

// ---- bf16 image (mlp_tc.cu): chunks of [N rows x 64 K] bf16 in the 128B-swizzled
// K-major UMMA shared-memory layout, in consumption order (see mlp_tc.cu); the first
// half of a chunk's bytes holds rows [0, N/2) (CTA 0 of a pair), the second half the rest
constexpr int kTcChunkK = 64;
constexpr int kTcNumGemm = 10;  // L0..L7, remap, rgb0
__host__ __device__ constexpr int tc_layer_k(int l) { return l == 0 ? 64 : (l == 5 ? 320 : 256); }
__host__ __device__ constexpr int tc_layer_n(int l) { return l == 9 ? 128 : 256; }
__host__ __device__ constexpr int tc_layer_chunks(int l) { return tc_layer_k(l) / kTcChunkK; }
__host__ __device__ constexpr size_t tc_layer_off_bytes(int l) {
  size_t o = 0;
  for (int i = 0; i < l; ++i) o += (size_t)tc_layer_k(i) * tc_layer_n(i) * 2;
  return o;
}
constexpr size_t kTcBlobBytes = tc_layer_off_bytes(kTcNumGemm);

struct NetImage {
  float* f32_gemm = nullptr;   // kF32GemmFloats
  float* smalls = nullptr;     // kSmallFloats
  uint8_t* tc_blob = nullptr;  // kTcBlobBytes
  uint8_t* tc_blob_h = nullptr;  // the same image with fp16 elements (TGTC_MLP_F16)
  uint8_t* tc_blobT = nullptr; // transposed weights for the activation-gradient kernel (mlp_bwd.cu), bwd_blobT_bytes()
  bool set = false;
};

// per-ray style head (style_tc.cu): packed images of models.StyleMLP_before_concat / StyleMLP_Wild_multilayers
struct StyleImage {
  uint8_t* blob_c = nullptr;   // module 1: 18 chunks [256 x 64] bf16
  uint8_t* blob_w = nullptr;   // module 2: 34 chunks
  uint8_t* blob_c_h = nullptr; // the same two images with fp16 elements (TGTC_MLP_F16 stylised render)
  uint8_t* blob_w_h = nullptr;
  uint8_t* blob_T = nullptr;   // 45 transposed chunks for the style dgrad (style_bwd.cu)
  float* head_w = nullptr;     // [3][256] output layer of module 2
  float* bias_c = nullptr;     // [5][256] effective biases for the current latents
  float* bias_w = nullptr;     // [7][256]
  float* head_b = nullptr;     // [3]
  float* latents = nullptr;    // [2][32]
  uint8_t* tables = nullptr;   // device scratch for the packing / bias tables
  float* wlat = nullptr;       // per layer: the 32 latent columns [256][32] and the bias [256] (owned copies)
  float* wlatT = nullptr;      // per layer: the latent columns transposed [32][256] and their row sums [256]
  const void* bias_table = nullptr;  // device table the per-call effective-bias kernel walks
  const float* src[26] = {};         // the caller's parameter pointers the device tables were built for
  bool set = false;
};

struct tgtc_ctx {
  int device = 0;
  int num_sms = 0;
  NetImage net[2];
  StyleImage style;
  int64_t launches = 0;
  // staging arena for the *_host entry points
  void* arena = nullptr;
  size_t arena_bytes = 0;
  // optional per-launch timing of the MLP kernel (tgtc_profile_*)
  bool profile = false;
  std::vector<cudaEvent_t> ev_pool;   // pairs: [2i] start, [2i+1] stop
  size_t ev_used = 0;                 // events handed out since the last read
  std::vector<int> ev_kind;           // kind of pair i (TGTC_PROF_*)
  double prof_flops = 0.0;            // forward MLP (kind 0) algorithmic FLOPs since the last read
  double prof_work[4] = {0, 0, 0, 0}; // algorithmic FLOPs per kind
  cudaEvent_t coarse_done = nullptr;  // caller-owned: recorded by tgtc_train_step when the coarse net's gradient half is final
  // stylised render in several passes: odd passes run on this library-owned side stream (forked from / joined to the caller's
  // stream by events), so that one pass's kernel prologues and tails fill under the other's kernels
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
  // Style_train: which workspaces hold a forward stash (tgtc_style_train_backward refuses anything else)
  struct StyleFwdRec { const void* ws; int64_t n; int S, F, has_rand; };
  std::vector<StyleFwdRec> style_fwd;
};

// ---------------------------------------------------------------------------
// error plumbing
void tgtc_set_error(const char* fmt, ...);
int tgtc_cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define TGTC_CUDA(expr)                                                      \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) return tgtc_cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define TGTC_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      tgtc_set_error(__VA_ARGS__);     \
      return (code);                   \
    }                                  \
  } while (0)

#define TGTC_LAUNCH_CHECK(ctx)                                                      \
  do {                                                                              \
    (ctx)->launches++;                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) return tgtc_cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3) == 0; }

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------------------
// internal launchers (one per .cu file)

// pack.cu
int pack_weights(tgtc_ctx* ctx, int net, const float* const* params, cudaStream_t st);

// raygen.cu
int launch_raygen(tgtc_ctx* ctx, int H, int W, const double* K, const double* c2w, int ndc, double ndc_near,
                  int pixel_alignment, int64_t pix_begin, int64_t n, float* rays_o, float* rays_d, cudaStream_t st);

// sampling.cu
int launch_sample_uniform(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, int64_t n, int S, double near,
                          double far, const float* rnd, float* pts, float* ts, cudaStream_t st, const PhiloxSrc* prng = nullptr,
                          int harmony = 0);
int launch_sample_fine(tgtc_ctx* ctx, const float* rays_o, const float* rays_d, const float* ts, int64_t ts_stride,
                       const float* weights, int64_t n, int S, int n_fine, float* pts_out, float* ts_out,
                       int64_t* inds_out, float* samples_out, cudaStream_t st);

// composite.cu
int launch_composite(tgtc_ctx* ctx, const float* rgb, const float* sigma, const float* rgbsigma, const float* ts,
                     int64_t ts_stride, const float* noise, int white_bkgd, int64_t n, int S, float* rgb_out,
                     float* depth_out, float* acc_out, float* weights_out, cudaStream_t st, const PhiloxSrc* prng = nullptr);
int launch_philox_fill(tgtc_ctx* ctx, unsigned long long seed, uint32_t stream, int normal, float std, int64_t n, float* out, cudaStream_t st);

int launch_composite_backward(tgtc_ctx* ctx, const float* rgbsigma, const float* ts, int64_t ts_stride, const float* noise,
                              int white_bkgd, int64_t n, int S, const float* g_rgb, const float* g_depth, const float* g_acc,
                              float* d_rgbsigma, cudaStream_t st, const PhiloxSrc* prng = nullptr);

// how the MLP kernels get their per-sample inputs and where results go
struct MlpIO {
  // explicit-points mode (rays_o == nullptr): pts [M,3]; dirs [n_rays,3] if dirs_per_ray else [M,3]
  const float* pts = nullptr;
  const float* dirs = nullptr;
  int dirs_per_ray = 1;
  // ray mode: rays_o/rays_d [n_rays,3]; ts [n_rays,S] or nullptr -> uniform(t_scale, t_near)
  const float* rays_o = nullptr;
  const float* rays_d = nullptr;
  const float* ts = nullptr;
  float t_scale = 1.f;  // (float)(far-near)
  float t_near = 0.f;   // (float)near
  int64_t n_rays = 0;
  int S = 0;
  // outputs (rgbsigma packed, or rgb+sigma split; optional extras)
  float* rgbsigma = nullptr;
  float* rgb = nullptr;
  float* sigma = nullptr;
  float* base_remap = nullptr;
  float* pts_embed = nullptr;
  float* dirs_embed = nullptr;
  // fused K5 (mlp_tc.cu inference instantiations, ray mode, S in {64,128}: a 128-sample tile is one fine ray / two coarse rays):
  // when comp_rgb != nullptr the last epilogue composites the tile's rays (utils.alpha_composition, utils.py:354-386, with the
  // arithmetic of composite.cu bit for bit) and writes per-RAY outputs; rgbsigma may then be nullptr (nothing per-sample
  // leaves the SM except the optional weights)
  float* comp_rgb = nullptr;      // [n_rays,3]
  float* comp_depth = nullptr;    // [n_rays] or nullptr
  float* comp_acc = nullptr;      // [n_rays] or nullptr
  float* comp_weights = nullptr;  // [n_rays,S] or nullptr
  int comp_white_bkgd = 0;
  // fused K6+K7 (mlp_tc.cu inference instantiation of the FINE pass, 64 coarse + 64 new samples): when fine_weights != nullptr
  // the input-producer warps run sample_pdf + the sorted union (sample_fine.cuh, the stand-alone kernel's arithmetic) for every
  // ray one iteration ahead of its tile, write the row to fine_ts (== ts, [n_rays,128]) and read it back when they form the
  // tile's sample positions: no resampling kernel between the coarse and the fine launch
  const float* fine_weights = nullptr;   // [n_rays,64] coarse weights
  float* fine_ts = nullptr;              // [n_rays,128], the same buffer as ts
  float fine_t_scale = 1.f, fine_t_near = 0.f;   // the coarse ts row: linspace(0,1,64) * scale + near
};

// Activation stash of the training forward (mlp_tc.cu writes, mlp_bwd.cu reads).  Every array is a sequence of
// "tile images": per 128-sample tile and per 64-column block, [128 rows x 128 B] bf16 in the 128-byte-swizzled
// layout of the kernels' shared-memory operand tiles (byte = (row>>3)*1024 + (row&7)*128 + ((chunk ^ (row&7))<<4)),
// so a bulk copy brings a ready tcgen05 operand (K-major as [sample x feature], MN-major as [feature x sample]).
struct TcStash {
  uint8_t* h = nullptr;    // [ntiles][9][4 blocks][16 KB]  post-ReLU outputs of L0..L7 and remap
  uint8_t* f = nullptr;    // [ntiles][2 blocks][16 KB]     post-ReLU output of rgb0
  uint8_t* pe = nullptr;   // [ntiles][16 KB]               positional encoding (column 63 = 0)
  // ReLU masks as bits: [ntiles][10 layers: L0..L7, remap, rgb0][8 column blocks of 32][128 rows] u32; bit (31-c) of a word =
  // sign bit of the fp32 pre-activation of column 32*block+c (1 = the gradient is blocked).  The dgrad kernel reads these
  // 40 KB per tile instead of the 608 KB of activation images (which only the wgrad kernel still needs).
  uint32_t* mask = nullptr;
};
constexpr size_t kStashHBytesPerTile = 9 * 65536, kStashFBytesPerTile = 32768, kStashPeBytesPerTile = 16384,
                 kStashMaskBytesPerTile = 10 * 8 * 128 * 4;

// gradients of the pre-activations, written by the dgrad kernel and read by the wgrad kernel (tile images like TcStash)
struct TcDz {
  uint8_t* dz = nullptr;     // [ntiles][9][64 KB]  dz_0..dz_7, dz_remap
  uint8_t* dzf = nullptr;    // [ntiles][32 KB]     dz of rgb0
  uint8_t* dhead = nullptr;  // [ntiles][16 KB]     columns 0..2 = dz of rgb1, column 3 = d_sigma
};

// Style_train stash (tile images / mask words written by the training forward of the style head, read by style_bwd.cu)
constexpr int kStyleMaskLayers = 12;   // module 2 layers 0..6, then module 1 layers 0..4
struct StyleStash {
  uint8_t* c = nullptr;      // [ntiles][5][64 KB]  module 1 (StyleMLP_before_concat) outputs; image 4 = concat_features
  uint8_t* w = nullptr;      // [ntiles][7][64 KB]  module 2 (StyleMLP_Wild_multilayers) hidden outputs h0..h6
  uint8_t* pe = nullptr;     // [ntiles][16 KB]
  uint32_t* mask = nullptr;  // [ntiles][12][8][128]
};
struct StyleDz {
  uint8_t* dz = nullptr;     // [ntiles][12][64 KB]
  uint8_t* dhead = nullptr;  // [ntiles][16 KB]
};
size_t style_flat_floats();
size_t style_partial_floats();
int launch_style_dgrad(tgtc_ctx* ctx, const float* rgbsigma, const float* d_rgbsigma, const StyleStash& stash, const StyleDz& dz, int64_t M,
                       cudaStream_t st);
int launch_style_wgrad(tgtc_ctx* ctx, const StyleStash& stash, const uint8_t* remap_img, const StyleDz& dz, const float* lat1,
                       int64_t n_rays, int S, float* partial, float* R, float* wlat_part, float* grads, int accumulate, float* dlat,
                       int dlat_accumulate, cudaStream_t st, cudaEvent_t ev_after_kernel = nullptr);
size_t style_wlat_part_floats();
int launch_style_bias_rays(tgtc_ctx* ctx, const float* lat1, int64_t n_rays, float* bias_rays, cudaStream_t st);
int launch_style_concat_train(tgtc_ctx* ctx, const MlpIO& io, const float* bias_rays, const StyleStash& stash, cudaStream_t st);
int launch_style_wild_train(tgtc_ctx* ctx, const MlpIO& io, const float* bias_rays, const uint8_t* remap_img, const StyleStash& stash,
                            cudaStream_t st);

// style_tc.cu
int style_set_weights(tgtc_ctx* ctx, const float* const* params, cudaStream_t st);
int style_set_latents(tgtc_ctx* ctx, const float* latent1, const float* latent2, cudaStream_t st);
int launch_style_concat(tgtc_ctx* ctx, const MlpIO& io, uint8_t* cf_img, cudaStream_t st, bool f16 = false, const float* bias_rays = nullptr);
int launch_style_wild(tgtc_ctx* ctx, const MlpIO& io, const uint8_t* remap_img, const uint8_t* cf_img, cudaStream_t st, bool f16 = false,
                      const float* bias_rays = nullptr);

size_t style_stage_workspace_bytes(int64_t M);
int launch_style_concat_explicit(tgtc_ctx* ctx, const float* x, int64_t M, float* cf_out, uint8_t* ws, bool f16, cudaStream_t st);
int launch_style_wild_explicit(tgtc_ctx* ctx, const float* x, const float* concated, int64_t M, float* rgb_out, uint8_t* ws, bool f16,
                               cudaStream_t st);

// mlp_bwd.cu
size_t bwd_partial_floats();
size_t bwd_blobT_bytes();
size_t bwd_flat_floats();
int launch_mlp_dgrad(tgtc_ctx* ctx, int net, const float* rgbsigma, const float* d_rgbsigma, const TcStash& stash, const TcDz& dz,
                     int64_t M, cudaStream_t st);
int launch_mlp_wgrad(tgtc_ctx* ctx, const TcStash& stash, const TcDz& dz, const float* rays_d, const float* d_rgbsigma, int64_t M, int S,
                     float* partial, float* grads, int accumulate, cudaStream_t st);
int launch_adam(tgtc_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t n, double lr, double b1, double b2, double eps,
                int64_t step, cudaStream_t st);
int launch_mse_grad(tgtc_ctx* ctx, const float* rgb, const float* gt, int64_t n, float scale, float* g, float* sq_sum, cudaStream_t st);

// mlp_fp32.cu / mlp_tc.cu
int launch_mlp_fp32(tgtc_ctx* ctx, int net, const MlpIO& io, cudaStream_t st);
int launch_mlp_tc(tgtc_ctx* ctx, int net, const MlpIO& io, cudaStream_t st, bool f16 = false);
int launch_mlp_tc_train(tgtc_ctx* ctx, int net, const MlpIO& io, const TcStash& stash, cudaStream_t st);
int launch_mlp_tc_trunk(tgtc_ctx* ctx, int net, const MlpIO& io, uint8_t* remap_img, cudaStream_t st, bool f16 = false);
bool mlp_tc_supports(const MlpIO& io);

// ---------------------------------------------------------------------------
// device helpers shared by kernels

// torch.linspace(0,1,steps)[i] in fp32, bit for bit (fma form, symmetric halves)
__device__ __forceinline__ float linspace01(int i, int steps) {
  if (steps == 1) return 0.f;
  const float step = __fdiv_rn(1.0f, (float)(steps - 1));
  return (i < steps / 2) ? __fmaf_rn(step, (float)i, 0.0f) : __fmaf_rn(-step, (float)(steps - 1 - i), 1.0f);
}

// ts = linspace*(far-near)+near exactly as torch evaluates it (mul, then add)
__device__ __forceinline__ float coarse_t(int i, int steps, float t_scale, float t_near) {
  return __fadd_rn(__fmul_rn(linspace01(i, steps), t_scale), t_near);
}
