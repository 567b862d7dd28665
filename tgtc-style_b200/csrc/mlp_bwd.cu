// K9 -- backward of the fused NeRF MLP for the training step (the autograd graph of train_tgtcs.py:236-255
// through models.MLP_style.forward, models.py:95-117), bf16 operands on tcgen05, fp32 accumulation in TMEM.
//
// The training forward (mlp_tc.cu, stash mode) leaves every post-ReLU activation in HBM as "tile images"
// (common.cuh: TcStash).  The backward is three kernels per network pass:
//
//   mlp_dgrad_kernel   activation-gradient chain, same CTA-pair / two-slot / weight-ring structure as the
//                      forward kernel, 9 GEMMs per tile with TRANSPOSED weights:
//                        dz_f = (d_rgb . rgb(1-rgb) . W_rgb1) * 1[f>0]             CUDA cores (input producers)
//                        dz_r = (dz_f . W_rgb0[:, :256])      * 1[r>0]             g0  (K=128)
//                        dz_7 = (dz_r . W_remap + d_sigma w_sigma) * 1[h7>0]       g1
//                        dz_6 = (dz_7 . W_7) * 1[h6>0]  ...  dz_0 = (dz_1 . W_1) * 1[h0>0]   g2..g8
//                      (layer 5 uses only the hidden columns of W_5; the positional encoding and the view
//                      directions are not differentiable inputs).  Every dz tile is written to HBM as a tile image.
//   mlp_wgrad_kernel   dW_l = dz_l^T . x_l summed over all samples: per CTA a 256x256 fp32 accumulator in TMEM, the
//                      tile images of dz_l and x_l used directly as MN-major operands (K = samples), 12 jobs;
//                      bias gradients and the view-direction columns of rgb0 as column sums on CUDA cores.
//   grad_reduce_kernel sums the per-CTA partials into the flat fp32 gradient buffer in nn.Linear layout
//                      (deterministic: no atomics anywhere).
//
// HBM-bound by the stash traffic (about 25 KB per sample for forward + backward); see DESIGN.md.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <math.h>

namespace {
using namespace tcptx;

constexpr int kTileM = 128;

// ===========================================================================================================
// dgrad
// ===========================================================================================================
constexpr int kStages = 3;
constexpr int kStageBytes = 16384;   // this CTA's half (128 rows) of a [256 x 64] transposed-weight chunk
constexpr int kNumThreads = 512;      // w0 weights, w1 MMA/forwarder, w2-5 input producers, w6-13 epilogue, w14-15 tile movers
constexpr int kInWarp0 = 2, kEpiWarp0 = 6, kMoveWarp0 = 14;
constexpr int kNumEpiThreads = 256, kNumInThreads = 128;
constexpr int kNumGemm = 9;

constexpr int kOffAct = 0;                                  // 2 x [4 kblocks][128 x 128 B]
constexpr int kActBytes = 65536;
constexpr int kOffIn = kOffAct + 2 * kActBytes;             // dz_f staging: [2 kblocks][128 x 128 B], shared by both slots
constexpr int kOffW = kOffIn + 32768;                       // 3 x 16384
constexpr int kOffWRgb1 = kOffW + kStages * kStageBytes;    // 3 x 128 fp32
constexpr int kOffWSig = kOffWRgb1 + 384 * 4;               // 256 fp32
constexpr int kOffBars = kOffWSig + 256 * 4;
constexpr int kBarWFull = 0, kBarWEmpty = kStages, kBarInReady = 2 * kStages, kBarInFree = kBarInReady + 1,
              kBarActReady = kBarInFree + 1, kBarAccFull = kBarActReady + 2, kBarSlotFree = kBarAccFull + 2, kBarDzDone = kBarSlotFree + 2, kNumBars = kBarDzDone + 2;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");

struct DgradParams {
  const uint8_t* blobT;       // transposed weights, consumption order (pack.cu: pack_tcT_kernel)
  const float* smalls;
  const float4* rgbsigma;     // [M] forward outputs (r,g,b,sigma)
  const float4* d_rgbsigma;   // [M] dL/d(r,g,b,sigma) from the compositing backward
  const uint32_t* mask;       // forward ReLU mask words (common.cuh: TcStash.mask)
  uint8_t* dz;                // [ntiles][9][64 KB]  dz_0..dz_7, dz_r
  uint8_t* dzf;               // [ntiles][32 KB]
  uint8_t* dhead;             // [ntiles][16 KB]     columns 0..2 = d_rgb * rgb(1-rgb), column 3 = d_sigma
  int64_t M;
  int64_t ntiles;
  long long* dbg_trace;       // timing experiments: [4 roles][4 iters][9 gemms][2 slots][3] clock64 stamps of CTA 0
};
#define BW_TRACE(role, it, g, t, k)                                                                        \
  do {                                                                                                     \
    if constexpr (kDbg)                                                                                    \
    if (P.dbg_trace != nullptr && blockIdx.x == 0 && (it) < 4)                                             \
      P.dbg_trace[(((((role)*4 + (int)(it)) * 9 + (g)) * 2 + (t)) * 3) + (k)] = clock64();                 \
  } while (0)

__device__ __forceinline__ int64_t pair_tile(int64_t it, int t, uint32_t rank) {
  const int64_t quad = (int64_t)(blockIdx.x >> 1) + it * (int64_t)(gridDim.x >> 1);
  return quad * 4 + 2 * (int64_t)rank + t;
}

__host__ __device__ constexpr size_t bwd_layer_off_bytes(int g) { return g == 0 ? 0 : 65536 + (size_t)(g - 1) * 131072; }
__host__ __device__ constexpr int bwd_layer_chunks(int g) { return g == 0 ? 2 : 4; }

// kDbg: the role clock trace of tools/bwd_trace.py -- its own instantiation (the stamps sit in the issue / epilogue loops)
template <bool kDbg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1) mlp_dgrad_kernel(const DgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bars = sbase + kOffBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  const uint32_t rank = cluster_ctarank();

  const int64_t nquads = (P.ntiles + 3) / 4;
  const int64_t ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const int64_t iters = nquads > cid ? (nquads - cid + ncl - 1) / ncl : 0;

  if (threadIdx.x == 0) {
    if ((sbase & 1023u) != 0) { printf("tgtc mlp_dgrad: shared memory base not 1024-aligned\n"); __trap(); }
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(kBarWFull + s), rank == 0 ? 2 : 1); mbar_init(bar(kBarWEmpty + s), 1); }
    mbar_init(bar(kBarInReady), 2 * (kNumInThreads / 32));
    mbar_init(bar(kBarInFree), 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar(kBarActReady + t), 2 * (kNumEpiThreads / 32));
      mbar_init(bar(kBarAccFull + t), 1);
      mbar_init(bar(kBarSlotFree + t), 1);
      mbar_init(bar(kBarDzDone + t), kNumEpiThreads / 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(sbase + kOffTmemPtr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
  }
  {
    float* wr = reinterpret_cast<float*>(smem + kOffWRgb1);
    for (int i = threadIdx.x; i < 384; i += kNumThreads) wr[i] = P.smalls[kSmWRgb1 + i];
    float* ws = reinterpret_cast<float*>(smem + kOffWSig);
    for (int i = threadIdx.x; i < 256; i += kNumThreads) ws[i] = P.smalls[kSmWSigma + i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (*reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr) != 0u) {
    if (threadIdx.x == 0) printf("tgtc mlp_dgrad: unexpected TMEM base\n");
    __trap();
  }
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0) {
    // ===================================================================== transposed-weight producer
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t it = 0; it < iters; ++it) {
      for (int g = 0; g < kNumGemm; ++g) {
        const uint8_t* src = P.blobT + bwd_layer_off_bytes(g) + (size_t)rank * kStageBytes;
        const int nch = bwd_layer_chunks(g);
        for (int t = 0; t < 2; ++t) {
          for (int c = 0; c < nch; ++c) {
            mbar_wait(bar(kBarWEmpty + stage), phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(bar(kBarWFull + stage), kStageBytes);
              bulk_g2s(sbase + kOffW + stage * kStageBytes, src + (size_t)c * 2 * kStageBytes, kStageBytes, bar(kBarWFull + stage));
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ===================================================================== peer: forward "my half landed"
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t leader_wfull = mapa_cluster(bar(kBarWFull), 0);
    for (int64_t it = 0; it < iters; ++it) {
      for (int c = 0; c < 2 * 34; ++c) {
        mbar_wait(bar(kBarWFull + stage), phase);
        if (elect_one()) mbar_arrive_cluster(leader_wfull + 8u * stage);
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== leader: MMA issuer for the pair
    int stage = 0;
    uint32_t phase = 0;
    uint32_t act_par0 = 0, act_par1 = 0, in_par = 0;
    const uint32_t w_lo0 = (((sbase + kOffW) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t in_lo = (((sbase + kOffIn) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t idesc = make_idesc(2 * kTileM, 256);
    auto issue_chunk = [&](uint32_t d_tmem, uint32_t a_lo, uint32_t accumulate) {
      mbar_wait_uniform(bar(kBarWFull + stage), phase);
      tc_fence_after();
      const uint32_t b_lo = w_lo0 + (uint32_t)stage * (kStageBytes >> 4);
      if (elect_one()) {
        umma_bf16_lohi(d_tmem, a_lo, kDescHiSW128, b_lo, kDescHiSW128, idesc, accumulate);
        umma_bf16_lohi(d_tmem, a_lo + 2u, kDescHiSW128, b_lo + 2u, kDescHiSW128, idesc, 1u);
        umma_bf16_lohi(d_tmem, a_lo + 4u, kDescHiSW128, b_lo + 4u, kDescHiSW128, idesc, 1u);
        umma_bf16_lohi(d_tmem, a_lo + 6u, kDescHiSW128, b_lo + 6u, kDescHiSW128, idesc, 1u);
        umma_commit(bar(kBarWEmpty + stage));
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    };
    for (int64_t it = 0; it < iters; ++it) {
      for (int g = 0; g < kNumGemm; ++g) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          uint32_t& act_par = t ? act_par1 : act_par0;
          if (g == 0) {
            mbar_wait_uniform(bar(kBarInReady), in_par); in_par ^= 1;
            if (it > 0) { mbar_wait_uniform(bar(kBarActReady + t), act_par); act_par ^= 1; }
          } else {
            mbar_wait_uniform(bar(kBarActReady + t), act_par); act_par ^= 1;
          }
          tc_fence_after();
          BW_TRACE(0, it, g, t, 0);
          const uint32_t d_tmem = tmem_base + (uint32_t)(256 * t);
          const uint32_t act_lo = (((sbase + kOffAct + t * kActBytes) & 0x3FFFFu) >> 4) | (1u << 16);
          if (g == 0) {
            issue_chunk(d_tmem, in_lo, 0u);
            issue_chunk(d_tmem, in_lo + 1024u, 1u);
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) issue_chunk(d_tmem, act_lo + 1024u * (uint32_t)c, c > 0 ? 1u : 0u);
          }
          if (elect_one()) {
            if (g == 0) umma_commit(bar(kBarInFree));
            umma_commit(bar(kBarAccFull + t));
          }
          __syncwarp();
          BW_TRACE(0, it, g, t, 1);
        }
      }
    }
  } else if (warp < kEpiWarp0) {
    // ===================================================================== input producers: dz_f and the head tile
    const int r = (warp - kInWarp0) * 32 + lane;
    const float* wr = reinterpret_cast<const float*>(smem + kOffWRgb1);
    const uint32_t leader_inready = mapa_cluster(bar(kBarInReady), 0);
    const uint32_t rowoff = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const uint32_t rx = (uint32_t)(r & 7);
    int64_t use = 0;
    for (int64_t it = 0; it < iters; ++it) {
      for (int t = 0; t < 2; ++t, ++use) {
        const int64_t tile = pair_tile(it, t, rank);
        const bool tile_ok = tile < P.ntiles;
        const int64_t m = tile * kTileM + r;
        const bool valid = tile_ok && m < P.M;
        float dzc[3] = {0.f, 0.f, 0.f};
        float dsig = 0.f;
        if (valid) {
          const float4 d = P.d_rgbsigma[m];
          const float4 o = P.rgbsigma[m];
          dzc[0] = d.x * o.x * (1.0f - o.x);   // sigmoid' (models.py:111)
          dzc[1] = d.y * o.y * (1.0f - o.y);
          dzc[2] = d.z * o.z * (1.0f - o.z);
          dsig = d.w;
        }
        // ReLU mask of the rgb0 output (4 words = 128 columns); fetched before waiting for the staging buffer
        uint32_t fm[4] = {~0u, ~0u, ~0u, ~0u};
        if (tile_ok) {
          const uint32_t* ms = P.mask + (((size_t)tile * 10 + 9) * 8) * 128 + r;
#pragma unroll
          for (int w = 0; w < 4; ++w) fm[w] = __ldg(ms + w * 128);
        }
        if (r == 0) BW_TRACE(2, it, 0, t, 0);
        if (use > 0) mbar_wait_relaxed(bar(kBarInFree), (uint32_t)((use - 1) & 1), 64);
        if (r == 0) BW_TRACE(2, it, 0, t, 1);
        uint8_t* gdst = tile_ok ? P.dzf + (size_t)tile * 32768 + rowoff : nullptr;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const int j0 = kb * 64 + ch * 8;
            float df[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              df[e] = fmaf(dzc[2], wr[256 + j0 + e], fmaf(dzc[1], wr[128 + j0 + e], dzc[0] * wr[j0 + e]));
            const uint32_t mb = fm[kb * 2 + (ch >> 2)] << ((ch & 3) * 8);   // bit 31 = first of these 8 columns
            const uint32_t q0 = mask_pack(mb, df[0], df[1]), q1 = mask_pack(mb << 2, df[2], df[3]), q2 = mask_pack(mb << 4, df[4], df[5]),
                           q3 = mask_pack(mb << 6, df[6], df[7]);
            const uint32_t off = (uint32_t)kb * 16384u + (((uint32_t)ch ^ rx) << 4);
            st_shared_v4(sbase + kOffIn + rowoff + off, q0, q1, q2, q3);
            if (gdst != nullptr) st_global_v4(gdst + off, q0, q1, q2, q3);
          }
        }
        if (tile_ok) {
          uint8_t* hd = P.dhead + (size_t)tile * 16384 + rowoff;
          st_global_v4(hd + ((0u ^ rx) << 4), pack_bf16(dzc[0], dzc[1]), pack_bf16(dzc[2], dsig), 0u, 0u);
#pragma unroll
          for (int ch = 1; ch < 8; ++ch) st_global_v4(hd + (((uint32_t)ch ^ rx) << 4), 0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_inready);
        if (r == 0) BW_TRACE(2, it, 0, t, 2);
      }
    }
  } else if (warp >= kMoveWarp0) {
    // ===================================================================== tile movers: one thread per slot
    // When the epilogue has filled act[t] with the masked gradient of a layer, send that tile image to HBM with one bulk store
    // and tell the epilogue warps when the store has finished reading shared memory (the next epilogue overwrites the slot).
    if (lane == 0) {
      const int t = warp - kMoveWarp0;
      uint32_t dz_par = 0;
      mbar_arrive(bar(kBarSlotFree + t));   // the slot starts out free
      for (int64_t it = 0; it < iters; ++it) {
        const int64_t tile = pair_tile(it, t, rank);
        const bool tile_ok = tile < P.ntiles;
        for (int g = 0; g < kNumGemm; ++g) {
          const int ml = 8 - g;
          mbar_wait(bar(kBarDzDone + t), dz_par); dz_par ^= 1;
          if (tile_ok) {
            bulk_s2g(P.dz + ((size_t)tile * 9 + ml) * 65536, sbase + kOffAct + t * kActBytes, 65536u);
            bulk_commit_group();
            bulk_wait_read0();
          }
          mbar_arrive(bar(kBarSlotFree + t));
        }
      }
      bulk_wait_all0();
    }
  } else {
    // ===================================================================== epilogue warps
    const int q = warp & 3;
    const int hc = (warp - kEpiWarp0) >> 2;
    const int row = q * 32 + lane;
    const float* wsig_s = reinterpret_cast<const float*>(smem + kOffWSig) + hc * 128;
    const uint32_t leader_actready = mapa_cluster(bar(kBarActReady), 0);
    const uint32_t rowoff = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + hc * 2 * 16384);
    const uint32_t rx = (uint32_t)(row & 7);
    uint32_t acc_par[2] = {0, 0}, h_par[2] = {0, 0};
    for (int64_t it = 0; it < iters; ++it) {
      for (int g = 0; g < kNumGemm; ++g) {
        const int ml = 8 - g;   // stash / dz index of the layer whose pre-activation gradient this GEMM produces
        for (int t = 0; t < 2; ++t) {
          const int64_t tile = pair_tile(it, t, rank);
          const bool tile_ok = tile < P.ntiles;
          const int64_t m = tile * kTileM + row;
          float dsig = 0.f;
          if (g == 1 && tile_ok && m < P.M) dsig = P.d_rgbsigma[m].w;
          // this thread's four ReLU mask words of the layer (fetched ahead of the accumulator wait)
          uint32_t mw[4] = {~0u, ~0u, ~0u, ~0u};
          if (tile_ok) {
            const uint32_t* ms = P.mask + (((size_t)tile * 10 + ml) * 8 + hc * 4) * 128 + row;
#pragma unroll
            for (int w = 0; w < 4; ++w) mw[w] = __ldg(ms + w * 128);
          }
          mbar_wait(bar(kBarAccFull + t), acc_par[t]); acc_par[t] ^= 1;
          tc_fence_after();
          if (lane == 0 && q == 2 && hc == 0) BW_TRACE(1, it, g, t, 0);
          // AccFull: the MMAs that read act[t] have retired; SlotFree: the slot's tile mover (warp 14/15) has finished reading the
          // previous gradient tile out of act[t].  From here the slot is this epilogue's to overwrite.
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(256 * t) + (uint32_t)(hc * 128);
          const uint32_t sdst = sbase + kOffAct + t * kActBytes + rowoff;
          mbar_wait(bar(kBarSlotFree + t), h_par[t]); h_par[t] ^= 1;
          if (lane == 0 && q == 2 && hc == 0) BW_TRACE(1, it, g, t, 1);
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) {
            const uint32_t kboff = (uint32_t)(blk >> 1) * 16384u;
            uint32_t v[32];
            tmem_ld32(taddr + blk * 32, v);
            tmem_ld_wait_dep(v);
            if (g == 1) {   // sigma head: dh7 += d_sigma * w_sigma (models.py:103; fp32)
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(fmaf(dsig, wsig_s[blk * 32 + c], __uint_as_float(v[c])));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t coff = kboff + ((((uint32_t)(blk & 1) * 4u + (uint32_t)j) ^ rx) << 4);
              const uint32_t mb = mw[blk] << (8 * j);   // bit 31 = column 8j of this block
              const uint32_t q0 = mask_pack(mb, __uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
              const uint32_t q1 = mask_pack(mb << 2, __uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
              const uint32_t q2 = mask_pack(mb << 4, __uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
              const uint32_t q3 = mask_pack(mb << 6, __uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
              st_shared_v4(sdst + coff, q0, q1, q2, q3);   // A operand of the next GEMM (in place) and source of the dz store
            }
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar(kBarDzDone + t));                      // local: the mover may store the dz tile
            mbar_arrive_cluster(leader_actready + 8u * t);         // pair: the leader may issue the next GEMM
          }
          if (lane == 0 && q == 2 && hc == 0) BW_TRACE(1, it, g, t, 2);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
  }
}

// ===========================================================================================================
// wgrad
// ===========================================================================================================
// One CTA per SM, cta_group::1.  12 jobs; per job the CTA walks its tiles in 64-sample stages:
//   A = image blocks of dz (out features, MN-major, K = samples), B = image blocks of the layer input.
constexpr int kWStages = 3;
constexpr int kWStageBytes = 65536;      // A: up to 4 blocks x 8 KB, then B: up to 4 blocks x 8 KB
constexpr int kWThreads = 448;           // w0 producer, w1 MMA, w2-5 drain, w6-13 column sums
constexpr int kNumJobs = 12;
constexpr int kWOffBars = kWStages * kWStageBytes;
constexpr int kWBarFull = 0, kWBarEmpty = kWStages, kWBarAccDone = 2 * kWStages, kWBarAccFree = kWBarAccDone + 1,
              kWNumBars = kWBarAccFree + 1;
constexpr int kWOffTmemPtr = kWOffBars + kWNumBars * 8;
constexpr int kWSmemBytes = kWOffTmemPtr + 16;

enum { SRC_DZ = 0, SRC_DZF = 1, SRC_DHEAD = 2, SRC_H = 3, SRC_F = 4, SRC_PE = 5 };
struct WJob { int a_src, a_idx, a_blocks, b_src, b_idx, b_blocks, colsum; };
__constant__ WJob c_jobs[kNumJobs] = {
    {SRC_DZ, 0, 4, SRC_PE, 0, 1, 1},     // 0  L0        dW[256 x 64]
    {SRC_DZ, 1, 4, SRC_H, 0, 4, 1},      // 1  L1
    {SRC_DZ, 2, 4, SRC_H, 1, 4, 1},      // 2  L2
    {SRC_DZ, 3, 4, SRC_H, 2, 4, 1},      // 3  L3
    {SRC_DZ, 4, 4, SRC_H, 3, 4, 1},      // 4  L4
    {SRC_DZ, 5, 4, SRC_PE, 0, 1, 1},     // 5  L5 (PE columns)
    {SRC_DZ, 5, 4, SRC_H, 4, 4, 0},      // 6  L5 (hidden columns)
    {SRC_DZ, 6, 4, SRC_H, 5, 4, 1},      // 7  L6
    {SRC_DZ, 7, 4, SRC_H, 6, 4, 1},      // 8  L7
    {SRC_DZ, 8, 4, SRC_H, 7, 4, 1},      // 9  remap
    {SRC_DZF, 0, 2, SRC_H, 8, 4, 1},     // 10 rgb0 (remap columns) + view-direction columns from the column sums
    {SRC_DHEAD, 0, 1, SRC_F, 0, 2, 1},   // 11 rgb1 (rows 0..2); column sums: rgb1 bias (0..2), sigma bias (3)
};
// The sigma head's weight gradient  dw_sigma[i] = sum_s d_sigma[s] * h7[s,i]  needs no job of its own: h7 is the B operand of
// job 9, so the column-sum warps take it as a d_sigma-weighted column sum of the staged tile (fp32 d_sigma, no extra HBM read).
// offsets (floats) of each job's [Mo x Ni] partial inside a CTA's partial block; Mo = 256 (4 A blocks) else 128
__host__ __device__ constexpr int job_mo(int j) { return (j <= 9) ? 256 : 128; }
__host__ __device__ constexpr int job_ni(int j) { return (j == 0 || j == 5) ? 64 : (j == 11 ? 128 : 256); }
__host__ __device__ constexpr size_t job_off(int j) {
  size_t o = 0;
  for (int i = 0; i < j; ++i) o += (size_t)job_mo(i) * job_ni(i);
  return o;
}
constexpr size_t kPartMat = job_off(kNumJobs);            // 638 976 floats
constexpr size_t kPartColsum = kPartMat;                  // [12][256]
constexpr size_t kPartDir = kPartColsum + kNumJobs * 256; // [128][32]
constexpr size_t kPartSig = kPartDir + 128 * 32;          // [256] sigma-head weight gradient
constexpr size_t kPartFloats = kPartSig + 256;

struct WgradParams {
  const uint8_t* stash_h;
  const uint8_t* stash_f;
  const uint8_t* stash_pe;
  const uint8_t* dz;
  const uint8_t* dzf;
  const uint8_t* dhead;
  const float* rays_d;     // [n_rays,3]
  const float4* d_rgbsigma; // [M] (d_sigma in .w)
  int64_t M;
  float* partial;          // [gridDim.x][kPartFloats]
  int64_t ntiles;
  int64_t n_rays;
  int S;
};

__device__ __forceinline__ const uint8_t* wsrc_block(const WgradParams& P, int src, int idx, int64_t tile, int blk, int half) {
  const uint8_t* base;
  switch (src) {
    case SRC_DZ: base = P.dz + ((size_t)tile * 9 + idx) * 65536; break;
    case SRC_DZF: base = P.dzf + (size_t)tile * 32768; break;
    case SRC_DHEAD: base = P.dhead + (size_t)tile * 16384; break;
    case SRC_H: base = P.stash_h + ((size_t)tile * 9 + idx) * 65536; break;
    case SRC_F: base = P.stash_f + (size_t)tile * 32768; break;
    default: base = P.stash_pe + (size_t)tile * 16384; break;
  }
  return base + (size_t)blk * 16384 + (size_t)half * 8192;   // rows [64*half, 64*half+64) of a block are contiguous
}


__global__ void __launch_bounds__(kWThreads, 1) mlp_wgrad_kernel(const WgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bars = sbase + kWOffBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  const int64_t n_my = P.ntiles > (int64_t)blockIdx.x ? (P.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t nst = 2 * n_my;   // 64-sample stages per job
  float* part = P.partial + (size_t)blockIdx.x * kPartFloats;

  if (threadIdx.x == 0) {
    if ((sbase & 1023u) != 0) { printf("tgtc mlp_wgrad: shared memory base not 1024-aligned\n"); __trap(); }
    for (int s = 0; s < kWStages; ++s) { mbar_init(bar(kWBarFull + s), 1); mbar_init(bar(kWBarEmpty + s), 1 + 8); }
    mbar_init(bar(kWBarAccDone), 1);
    mbar_init(bar(kWBarAccFree), 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(sbase + kWOffTmemPtr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*reinterpret_cast<volatile uint32_t*>(smem + kWOffTmemPtr) != 0u) {
    if (threadIdx.x == 0) printf("tgtc mlp_wgrad: unexpected TMEM base\n");
    __trap();
  }

  if (warp == 0) {
    // ===================================================================== producer
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < kNumJobs; ++j) {
      const WJob jb = c_jobs[j];
      const uint32_t bytes = (uint32_t)(jb.a_blocks + jb.b_blocks) * 8192u;
      for (int64_t s = 0; s < nst; ++s) {
        const int64_t tile = (int64_t)blockIdx.x + (s >> 1) * gridDim.x;
        const int half = (int)(s & 1);
        mbar_wait(bar(kWBarEmpty + stage), phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar(kWBarFull + stage), bytes);
          const uint32_t dst = sbase + stage * kWStageBytes;
          for (int b = 0; b < jb.a_blocks; ++b) bulk_g2s(dst + b * 8192, wsrc_block(P, jb.a_src, jb.a_idx, tile, b, half), 8192, bar(kWBarFull + stage));
          for (int b = 0; b < jb.b_blocks; ++b)
            bulk_g2s(dst + 32768 + b * 8192, wsrc_block(P, jb.b_src, jb.b_idx, tile, b, half), 8192, bar(kWBarFull + stage));
        }
        __syncwarp();
        if (++stage == kWStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < kNumJobs; ++j) {
      const WJob jb = c_jobs[j];
      const int nm = jb.a_blocks >= 4 ? 2 : 1;                          // 128-row halves of the output
      const uint32_t a_lbo = jb.a_blocks >= 2 ? 8192u : 0u;            // a single 64-feature block is replicated (rows 64..127 unused)
      const uint32_t idesc = make_idesc_mn(128, jb.b_blocks * 64);
      if (j > 0) { mbar_wait_uniform(bar(kWBarAccFree), (uint32_t)((j - 1) & 1)); tc_fence_after(); }
      for (int64_t s = 0; s < nst; ++s) {
        mbar_wait_uniform(bar(kWBarFull + stage), phase);
        tc_fence_after();
        const uint32_t st = sbase + stage * kWStageBytes;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {        // 16 samples per MMA = two 8-row groups = 2 KB
            const uint32_t b_lo = mn_desc_lo(st + 32768 + ks * 2048, 8192u);
            for (int mh = 0; mh < nm; ++mh) {
              const uint32_t a_lo = mn_desc_lo(st + mh * 16384 + ks * 2048, a_lbo);
              umma1_bf16((uint32_t)(256 * mh), a_lo, mn_desc_hi(), b_lo, mn_desc_hi(), idesc, (s > 0 || ks > 0) ? 1u : 0u);
            }
          }
          umma1_commit(bar(kWBarEmpty + stage));
          if (s == nst - 1) umma1_commit(bar(kWBarAccDone));
        }
        __syncwarp();
        if (++stage == kWStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ===================================================================== drain: TMEM -> per-CTA partial
    const int q = warp & 3;
    const int row = q * 32 + lane;
    for (int j = 0; j < kNumJobs; ++j) {
      const WJob jb = c_jobs[j];
      const int nm = jb.a_blocks >= 4 ? 2 : 1;
      const int ni = jb.b_blocks * 64;
      float* dst = part + job_off(j);
      if (nst > 0) {
        mbar_wait(bar(kWBarAccDone), (uint32_t)(j & 1));
        tc_fence_after();
      }
      for (int mh = 0; mh < nm; ++mh) {
        float* drow = dst + (size_t)(mh * 128 + row) * ni;
        for (int cb = 0; cb < ni; cb += 32) {
          uint32_t v[32];
          if (nst > 0) {
            tmem_ld32(((uint32_t)(q * 32) << 16) + (uint32_t)(256 * mh + cb), v);
            tmem_ld_wait_dep(v);
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = 0u;
          }
#pragma unroll
          for (int c = 0; c < 32; c += 4) *reinterpret_cast<uint4*>(drow + cb + c) = make_uint4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kWBarAccFree));
    }
  } else {
    // ===================================================================== column sums of A (bias gradients, dir columns)
    const int c = threadIdx.x - 6 * 32;          // feature column 0..255
    const int cb = c >> 6, cc = c & 63;
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < kNumJobs; ++j) {
      const WJob jb = c_jobs[j];
      const bool mine = jb.colsum && cb < jb.a_blocks;
      float total = 0.f, sig_total = 0.f;
      float dir_acc[27];
      if (j == 10) {
#pragma unroll
        for (int k = 0; k < 27; ++k) dir_acc[k] = 0.f;
      }
      for (int64_t s = 0; s < nst; ++s) {
        float ds_lo = 0.f, ds_hi = 0.f;
        if (j == 9) {
          // each lane fetches two of the stage's 64 d_sigma values (before the wait, so the latency hides behind it)
          const int64_t m0 = ((int64_t)blockIdx.x + (s >> 1) * gridDim.x) * kTileM + (s & 1) * 64;
          if (m0 + lane < P.M) ds_lo = __ldg(&P.d_rgbsigma[m0 + lane].w);
          if (m0 + 32 + lane < P.M) ds_hi = __ldg(&P.d_rgbsigma[m0 + 32 + lane].w);
        }
        mbar_wait(bar(kWBarFull + stage), phase);
        if (mine) {
          const uint8_t* blk = smem + stage * kWStageBytes + cb * 8192;
          float acc = 0.f;
#pragma unroll 8
          for (int r = 0; r < 64; ++r) {
            const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((cc >> 3) ^ (r & 7))) << 4) + (cc & 7) * 2);
            acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(blk + off));
          }
          total += acc;
          if (j == 9) {
            // B operand of this job = h7: d_sigma-weighted column sum = the sigma head's weight gradient (models.py:103)
            const uint8_t* hb = smem + stage * kWStageBytes + 32768 + cb * 8192;
            float sacc = 0.f;   // d_sigma broadcast lane -> warp by shuffle
#pragma unroll
            for (int r = 0; r < 64; ++r) {
              const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((cc >> 3) ^ (r & 7))) << 4) + (cc & 7) * 2);
              const float ds = __shfl_sync(0xffffffffu, r < 32 ? ds_lo : ds_hi, r & 31);
              sacc = fmaf(ds, __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(hb + off)), sacc);
            }
            sig_total += sacc;
          }
          if (j == 10 && c < 128) {
            // all 64 samples of a stage belong to one ray (S is a multiple of 64): view-direction columns of rgb0
            // (models.py:108) get  sum_s dz_f[s,o] * dirPE_k(ray)  =  colsum * dirPE_k
            const int64_t tile = (int64_t)blockIdx.x + (s >> 1) * gridDim.x;
            int64_t ray = (tile * kTileM + (s & 1) * 64) / P.S;
            if (ray >= P.n_rays) ray = P.n_rays - 1;   // padding half of the last tile: its dz rows are zero
            const float v[3] = {P.rays_d[ray * 3 + 0], P.rays_d[ray * 3 + 1], P.rays_d[ray * 3 + 2]};
#pragma unroll
            for (int a = 0; a < 3; ++a) dir_acc[a] = fmaf(acc, v[a], dir_acc[a]);
#pragma unroll
            for (int f = 0; f < 4; ++f) {
              const float fr = (float)(1 << f);
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                float sn, cs;
                fast_sincos(__fmul_rn(v[a], fr), &sn, &cs);
                dir_acc[3 + 6 * f + a] = fmaf(acc, sn, dir_acc[3 + 6 * f + a]);
                dir_acc[3 + 6 * f + 3 + a] = fmaf(acc, cs, dir_acc[3 + 6 * f + 3 + a]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kWBarEmpty + stage));
        if (++stage == kWStages) { stage = 0; phase ^= 1; }
      }
      part[kPartColsum + j * 256 + c] = total;
      if (j == 9) part[kPartSig + c] = sig_total;
      if (j == 10 && c < 128) {
#pragma unroll
        for (int k = 0; k < 27; ++k) part[kPartDir + c * 32 + k] = dir_acc[k];
#pragma unroll
        for (int k = 27; k < 32; ++k) part[kPartDir + c * 32 + k] = 0.f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(0u), "r"(512));
  }
}

// ===========================================================================================================
// reduction of the per-CTA partials into the flat gradient buffer (nn.Linear layout, models.py:75-91)
// ===========================================================================================================
// flat order: for layer in (L0..L7, sigma, remap, rgb0, rgb1): weight [out,in] row-major, then bias [out]
__host__ __device__ constexpr int p_out(int l) { return l < 8 ? 256 : (l == 8 ? 1 : (l == 9 ? 256 : (l == 10 ? 128 : 3))); }
__host__ __device__ constexpr int p_in(int l) { return l == 0 ? 63 : (l == 5 ? 319 : (l == 10 ? 283 : (l == 11 ? 128 : 256))); }
__host__ __device__ constexpr size_t p_off(int l) {
  size_t o = 0;
  for (int i = 0; i < l; ++i) o += (size_t)p_out(i) * p_in(i) + p_out(i);
  return o;
}
constexpr size_t kFlatFloats = p_off(12);
static_assert(kFlatFloats == 595844, "parameter count of one StyleNerf (SURVEY.md 8a)");

// where flat element idx lives inside a CTA's partial block (-1: always zero)
__device__ __forceinline__ long long flat_to_partial(size_t idx) {
  int l = 0;
  while (l + 1 < 12 && idx >= p_off(l + 1)) ++l;
  const size_t r = idx - p_off(l);
  const int no = p_out(l), ni = p_in(l);
  if (r >= (size_t)no * ni) {                      // bias
    const int o = (int)(r - (size_t)no * ni);
    switch (l) {
      case 0: case 1: case 2: case 3: case 4: case 5: return kPartColsum + l * 256 + o;
      case 6: return kPartColsum + 7 * 256 + o;
      case 7: return kPartColsum + 8 * 256 + o;
      case 8: return kPartColsum + 11 * 256 + 3;   // d_sigma column of the head tile
      case 9: return kPartColsum + 9 * 256 + o;
      case 10: return kPartColsum + 10 * 256 + o;
      default: return kPartColsum + 11 * 256 + o;  // rgb1 bias: columns 0..2 of the head tile
    }
  }
  const int o = (int)(r / ni), i = (int)(r % ni);
  switch (l) {
    case 0: return job_off(0) + (size_t)o * 64 + i;
    case 1: case 2: case 3: case 4: return job_off(l) + (size_t)o * 256 + i;
    case 5: return i < 63 ? job_off(5) + (size_t)o * 64 + i : job_off(6) + (size_t)o * 256 + (i - 63);
    case 6: return job_off(7) + (size_t)o * 256 + i;
    case 7: return job_off(8) + (size_t)o * 256 + i;
    case 8: return kPartSig + i;
    case 9: return job_off(9) + (size_t)o * 256 + i;
    case 10: return i < 256 ? job_off(10) + (size_t)o * 256 + i : kPartDir + (size_t)o * 32 + (i - 256);
    default: return job_off(11) + (size_t)o * 128 + i;
  }
}

__global__ void grad_reduce_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ grads, int accumulate) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < kFlatFloats; idx += (size_t)gridDim.x * blockDim.x) {
    const long long src = flat_to_partial(idx);
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * kPartFloats + src];
    grads[idx] = accumulate ? grads[idx] + s : s;
  }
}

// torch.optim.Adam (the reference's optimizer, train_tgtcs.py:39: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad)
// on the flat fp32 parameter buffer of both nets: one pass, 28 B per parameter.  step = 1-based step count.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                            float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;            // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;           // (exp_avg_sq.sqrt() / sqrt(bias_correction2)).add_(eps)
    p[i] = p[i] - (lr / bc1) * (mi / denom);                  // param.addcdiv_(exp_avg, denom, value=-lr / bias_correction1)
  }
}

// g = scale * (rgb - gt); the squared error is added to *sq_sum (the loss value).  One block, fixed summation order: the loss is
// bit-reproducible like everything else of the step (an atomicAdd over blocks made it depend on scheduling).
__global__ void __launch_bounds__(1024) mse_grad_kernel(const float* __restrict__ rgb, const float* __restrict__ gt, int64_t n3, float scale,
                                                        float* __restrict__ g, float* __restrict__ sq_sum) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n3; i += blockDim.x) {
    const float d = rgb[i] - gt[i];
    g[i] = scale * d;
    acc += d * d;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && sq_sum != nullptr) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    *sq_sum += t;
  }
}

}  // namespace

static long long* g_bwd_trace = nullptr;
extern "C" void tgtc_debug_bwd_trace(long long* dev_buf) { g_bwd_trace = dev_buf; }

size_t bwd_partial_floats() { return kPartFloats; }
size_t bwd_blobT_bytes() { return bwd_layer_off_bytes(kNumGemm); }
size_t bwd_flat_floats() { return kFlatFloats; }

int launch_mlp_dgrad(tgtc_ctx* ctx, int net, const float* rgbsigma, const float* d_rgbsigma, const TcStash& stash, const TcDz& dz,
                     int64_t M, cudaStream_t st) {
  const NetImage& im = ctx->net[net];
  if (M == 0) return TGTC_OK;
  DgradParams P;
  P.blobT = im.tc_blobT;
  P.smalls = im.smalls;
  P.rgbsigma = reinterpret_cast<const float4*>(rgbsigma);
  P.d_rgbsigma = reinterpret_cast<const float4*>(d_rgbsigma);
  P.mask = stash.mask;
  P.dz = dz.dz; P.dzf = dz.dzf; P.dhead = dz.dhead;
  P.M = M;
  P.ntiles = (M + kTileM - 1) / kTileM;
  P.dbg_trace = g_bwd_trace;
  static bool attr_set[64] = {};
  if (!attr_set[ctx->device & 63]) {
    TGTC_CUDA(cudaFuncSetAttribute(mlp_dgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    TGTC_CUDA(cudaFuncSetAttribute(mlp_dgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set[ctx->device & 63] = true;
  }
  const int64_t nquads = (P.ntiles + 3) / 4;
  const int64_t max_pairs = ctx->num_sms / 2;
  const int grid = 2 * (int)(nquads < max_pairs ? nquads : max_pairs);
  if (g_bwd_trace != nullptr) mlp_dgrad_kernel<true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else mlp_dgrad_kernel<false><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

int launch_mlp_wgrad(tgtc_ctx* ctx, const TcStash& stash, const TcDz& dz, const float* rays_d, const float* d_rgbsigma, int64_t M, int S,
                     float* partial, float* grads, int accumulate, cudaStream_t st) {
  if (M == 0) return TGTC_OK;
  WgradParams P;
  P.stash_h = stash.h; P.stash_f = stash.f; P.stash_pe = stash.pe;
  P.dz = dz.dz; P.dzf = dz.dzf; P.dhead = dz.dhead;
  P.rays_d = rays_d;
  P.d_rgbsigma = reinterpret_cast<const float4*>(d_rgbsigma);
  P.M = M;
  P.partial = partial;
  P.ntiles = (M + kTileM - 1) / kTileM;
  P.n_rays = M / S;
  P.S = S;
  static bool attr_set[64] = {};
  if (!attr_set[ctx->device & 63]) {
    TGTC_CUDA(cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmemBytes));
    attr_set[ctx->device & 63] = true;
  }
  const int grid = (int)(P.ntiles < ctx->num_sms ? P.ntiles : ctx->num_sms);
  mlp_wgrad_kernel<<<grid, kWThreads, kWSmemBytes, st>>>(P);
  TGTC_LAUNCH_CHECK(ctx);
  grad_reduce_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(partial, grid, grads, accumulate);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

int launch_adam(tgtc_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t n, double lr, double b1, double b2, double eps,
                int64_t step, cudaStream_t st) {
  if (n == 0) return TGTC_OK;
  const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
  adam_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(p, g, m, v, n, (float)lr, (float)b1, (float)b2, (float)eps, (float)bc1, (float)sqrt(bc2));
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

int launch_mse_grad(tgtc_ctx* ctx, const float* rgb, const float* gt, int64_t n, float scale, float* g, float* sq_sum, cudaStream_t st) {
  if (n == 0) return TGTC_OK;
  const int64_t n3 = n * 3;
  mse_grad_kernel<<<1, 1024, 0, st>>>(rgb, gt, n3, scale, g, sq_sum);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}
