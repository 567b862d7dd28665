// Backward of the per-ray style head for Style_train (train_tgtcs.py:311-495): gradients of the stylised rgb w.r.t. the
// parameters of StyleMLP_Wild_multilayers (models.py:149-180) and StyleMLP_before_concat (models.py:120-147) and w.r.t. the
// per-ray latents.  The NeRF nets are frozen in this phase (style_optimizer holds only the two style modules,
// train_tgtcs.py:54), so nothing flows into base_remap / sigma.  bf16 operands on tcgen05, fp32 accumulation in TMEM.
//
// Same three-kernel structure as mlp_bwd.cu, on the stash written by mlp_chain_kernel<true> (style_tc.cu):
//   style_dgrad_kernel   12 GEMMs per tile with transposed weights (CTA pairs, two slots, weight ring):
//        g0   dz_w6 = (dz_head . W_out[:, :256]) * 1[h_w6>0]        K = 64 (the head tile: columns 0..2 = d_rgb rgb(1-rgb))
//        g1-6 dz_w5 .. dz_w0                                         (hidden columns of W6 .. W1)
//        g7   dz_c4 = (dz_w0 . W0[:, 256:512]) * 1[concat_features>0]
//        g8-11 dz_c3 .. dz_c0                                         (hidden columns of C4 .. C1)
//      every dz tile goes to HBM as a tile image [ntiles][12][64 KB].
//   style_wgrad_kernel   17 jobs dW = dz^T . x (MN-major operands straight from the tile images, K = samples), 256x256 fp32
//      accumulators in TMEM; the column-sum warps also leave the per-64-sample column sums of every dz ("R" rows): biases,
//      latent columns and latent gradients are all linear in those.
//   style_reduce_kernel  per-CTA partials -> flat gradient buffer (nn.Linear layout, deterministic)
//   style_latgrad_kernel / style_wlat_partial+reduce kernels   d latent[ray] and the latent columns of every weight from the R rows.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace tcptx;

constexpr int kTileM = 128;

// ===========================================================================================================
// dgrad
// ===========================================================================================================
constexpr int kStages = 3;
constexpr int kStageBytes = 16384;
constexpr int kNumThreads = 512;      // w0 weights, w1 MMA/forwarder, w2-5 head-tile producers, w6-13 epilogue, w14-15 tile movers
constexpr int kInWarp0 = 2, kEpiWarp0 = 6, kMoveWarp0 = 14;
constexpr int kNumEpiThreads = 256, kNumInThreads = 128;
constexpr int kNumGemm = 12;
constexpr int kChunksPerTile = 1 + 11 * 4;

constexpr int kOffAct = 0;                                  // 2 x [4 kblocks][128 x 128 B]
constexpr int kActBytes = 65536;
constexpr int kOffIn = kOffAct + 2 * kActBytes;             // head tile staging: [128 x 128 B], shared by both slots
constexpr int kOffW = kOffIn + 16384;                       // 3 x 16384
constexpr int kOffBars = kOffW + kStages * kStageBytes;
constexpr int kBarWFull = 0, kBarWEmpty = kStages, kBarInReady = 2 * kStages, kBarInFree = kBarInReady + 1,
              kBarActReady = kBarInFree + 1, kBarAccFull = kBarActReady + 2, kBarSlotFree = kBarAccFull + 2, kBarDzDone = kBarSlotFree + 2,
              kNumBars = kBarDzDone + 2;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;

struct SDgradParams {
  const uint8_t* blobT;       // 45 transposed chunks in consumption order (style_tc.cu: style_set_weights)
  const float4* rgbsigma;     // [M] forward outputs (stylised r,g,b; sigma)
  const float4* d_rgbsigma;   // [M] dL/d(r,g,b,sigma) from the compositing backward (sigma ignored: the NeRF nets are frozen)
  const uint32_t* mask;       // [ntiles][12][8][128]
  uint8_t* dz;                // [ntiles][12][64 KB]  image g = gradient at the pre-activation of w6,w5,...,w0,c4,c3,c2,c1,c0
  uint8_t* dhead;             // [ntiles][16 KB]      columns 0..2 = d_rgb * rgb(1-rgb)
  int64_t M;
  int64_t ntiles;
};

__device__ __forceinline__ int64_t pair_tile(int64_t it, int t, uint32_t rank) {
  const int64_t quad = (int64_t)(blockIdx.x >> 1) + it * (int64_t)(gridDim.x >> 1);
  return quad * 4 + 2 * (int64_t)rank + t;
}
__host__ __device__ constexpr size_t s_layer_off_bytes(int g) { return g == 0 ? 0 : 32768 + (size_t)(g - 1) * 131072; }
__host__ __device__ constexpr int s_layer_chunks(int g) { return g == 0 ? 1 : 4; }
__host__ __device__ constexpr int s_mask_id(int g) { return g <= 6 ? 6 - g : 18 - g; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1) style_dgrad_kernel(const SDgradParams P) {
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
    if ((sbase & 1023u) != 0) { printf("tgtc style_dgrad: shared memory base not 1024-aligned\n"); __trap(); }
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
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (*reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr) != 0u) {
    if (threadIdx.x == 0) printf("tgtc style_dgrad: unexpected TMEM base\n");
    __trap();
  }
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0) {
    // ===================================================================== transposed-weight producer
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t it = 0; it < iters; ++it) {
      for (int g = 0; g < kNumGemm; ++g) {
        const uint8_t* src = P.blobT + s_layer_off_bytes(g) + (size_t)rank * kStageBytes;
        const int nch = s_layer_chunks(g);
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
      for (int c = 0; c < 2 * kChunksPerTile; ++c) {
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
          const uint32_t d_tmem = tmem_base + (uint32_t)(256 * t);
          const uint32_t act_lo = (((sbase + kOffAct + t * kActBytes) & 0x3FFFFu) >> 4) | (1u << 16);
          if (g == 0) {
            issue_chunk(d_tmem, in_lo, 0u);
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) issue_chunk(d_tmem, act_lo + 1024u * (uint32_t)c, c > 0 ? 1u : 0u);
          }
          if (elect_one()) {
            if (g == 0) umma_commit(bar(kBarInFree));
            umma_commit(bar(kBarAccFull + t));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < kEpiWarp0) {
    // ===================================================================== head-tile producers (one thread per tile row)
    const int r = (warp - kInWarp0) * 32 + lane;
    const uint32_t leader_inready = mapa_cluster(bar(kBarInReady), 0);
    const uint32_t rowoff = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const uint32_t rx = (uint32_t)(r & 7);
    int64_t use = 0;
    for (int64_t it = 0; it < iters; ++it) {
      for (int t = 0; t < 2; ++t, ++use) {
        const int64_t tile = pair_tile(it, t, rank);
        const bool tile_ok = tile < P.ntiles;
        const int64_t m = tile * kTileM + r;
        float dzc[3] = {0.f, 0.f, 0.f};
        if (tile_ok && m < P.M) {
          const float4 d = P.d_rgbsigma[m];
          const float4 o = P.rgbsigma[m];
          dzc[0] = d.x * o.x * (1.0f - o.x);   // sigmoid' (models.py:179)
          dzc[1] = d.y * o.y * (1.0f - o.y);
          dzc[2] = d.z * o.z * (1.0f - o.z);
        }
        if (use > 0) mbar_wait_relaxed(bar(kBarInFree), (uint32_t)((use - 1) & 1), 64);
        const uint32_t q0 = pack_bf16(dzc[0], dzc[1]), q1 = pack_bf16(dzc[2], 0.f);
        uint8_t* hd = tile_ok ? P.dhead + (size_t)tile * 16384 + rowoff : nullptr;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t off = ((uint32_t)ch ^ rx) << 4;
          st_shared_v4(sbase + kOffIn + rowoff + off, ch == 0 ? q0 : 0u, ch == 0 ? q1 : 0u, 0u, 0u);
          if (hd != nullptr) st_global_v4(hd + off, ch == 0 ? q0 : 0u, ch == 0 ? q1 : 0u, 0u, 0u);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_inready);
      }
    }
  } else if (warp >= kMoveWarp0) {
    // ===================================================================== tile movers: one thread per slot
    if (lane == 0) {
      const int t = warp - kMoveWarp0;
      uint32_t dz_par = 0;
      mbar_arrive(bar(kBarSlotFree + t));   // the slot starts out free
      for (int64_t it = 0; it < iters; ++it) {
        const int64_t tile = pair_tile(it, t, rank);
        const bool tile_ok = tile < P.ntiles;
        for (int g = 0; g < kNumGemm; ++g) {
          mbar_wait(bar(kBarDzDone + t), dz_par); dz_par ^= 1;
          if (tile_ok) {
            bulk_s2g(P.dz + ((size_t)tile * kNumGemm + g) * 65536, sbase + kOffAct + t * kActBytes, 65536u);
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
    const uint32_t leader_actready = mapa_cluster(bar(kBarActReady), 0);
    const uint32_t rowoff = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + hc * 2 * 16384);
    const uint32_t rx = (uint32_t)(row & 7);
    uint32_t acc_par[2] = {0, 0}, sf_par[2] = {0, 0};
    for (int64_t it = 0; it < iters; ++it) {
      for (int g = 0; g < kNumGemm; ++g) {
        const int ml = s_mask_id(g);
        for (int t = 0; t < 2; ++t) {
          const int64_t tile = pair_tile(it, t, rank);
          const bool tile_ok = tile < P.ntiles;
          uint32_t mw[4] = {~0u, ~0u, ~0u, ~0u};
          if (tile_ok) {
            const uint32_t* ms = P.mask + (((size_t)tile * kStyleMaskLayers + ml) * 8 + hc * 4) * 128 + row;
#pragma unroll
            for (int w = 0; w < 4; ++w) mw[w] = __ldg(ms + w * 128);
          }
          mbar_wait(bar(kBarAccFull + t), acc_par[t]); acc_par[t] ^= 1;
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(256 * t) + (uint32_t)(hc * 128);
          const uint32_t sdst = sbase + kOffAct + t * kActBytes + rowoff;
          mbar_wait(bar(kBarSlotFree + t), sf_par[t]); sf_par[t] ^= 1;
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) {
            const uint32_t kboff = (uint32_t)(blk >> 1) * 16384u;
            uint32_t v[32];
            tmem_ld32(taddr + blk * 32, v);
            tmem_ld_wait_dep(v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t coff = kboff + ((((uint32_t)(blk & 1) * 4u + (uint32_t)j) ^ rx) << 4);
              const uint32_t mb = mw[blk] << (8 * j);
              const uint32_t q0 = mask_pack(mb, __uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
              const uint32_t q1 = mask_pack(mb << 2, __uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
              const uint32_t q2 = mask_pack(mb << 4, __uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
              const uint32_t q3 = mask_pack(mb << 6, __uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
              st_shared_v4(sdst + coff, q0, q1, q2, q3);
            }
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar(kBarDzDone + t));
            mbar_arrive_cluster(leader_actready + 8u * t);
          }
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
constexpr int kWStages = 3;
constexpr int kWStageBytes = 65536;      // A: up to 4 blocks x 8 KB, then B: up to 4 blocks x 8 KB
constexpr int kWThreads = 448;           // w0 producer, w1 MMA, w2-5 drain, w6-13 column sums
constexpr int kMaxJobs = 17;
constexpr int kWOffBars = kWStages * kWStageBytes;
constexpr int kWBarFull = 0, kWBarEmpty = kWStages, kWBarAccDone = 2 * kWStages, kWBarAccFree = kWBarAccDone + 1,
              kWNumBars = kWBarAccFree + 1;
constexpr int kWOffTmemPtr = kWOffBars + kWNumBars * 8;
constexpr int kWSmemBytes = kWOffTmemPtr + 16;

struct SJob {
  const uint8_t* a;     // tile images of dz (out features): tile t at a + t * a_stride, blocks of 16 KB
  const uint8_t* b;     // tile images of the layer input
  int64_t a_stride, b_stride;
  int a_blocks, b_blocks;
  int r_slot;           // >= 0: column sums of A -> bias partial and per-stage R rows of slot r_slot; < 0: none
  int out_off;          // floats, inside a CTA's partial block
};
struct SWgradParams {
  SJob job[kMaxJobs];
  int njobs;
  float* partial;       // [gridDim.x][part_floats]
  int64_t part_floats;
  int64_t colsum_off;   // floats: [njobs][256]
  float* R;             // [13][2 * ntiles][256] per-64-sample column sums of dz
  int64_t ntiles;
  // Work split.  job_parallel = 0: every CTA walks all jobs over its tiles (tile = blockIdx.x + k * gridDim.x): best HBM
  // streaming at large batches, but each CTA flushes 17 partial matrices (3.5 MB).  job_parallel = 1 (small batches, the
  // reference's batch_size_style is 256..1024 rays): CTAs first[j] .. first[j+1]-1 share job j, a CTA's partial is one matrix.
  int job_parallel;
  int first[kMaxJobs + 1];
};
constexpr int64_t kCtaFloatsParallel = 256 * 256 + 256;   // job-parallel partial of one CTA: matrix, then column sums

__global__ void __launch_bounds__(kWThreads, 1) style_wgrad_kernel(const __grid_constant__ SWgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bars = sbase + kWOffBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  int j0 = 0, j1 = P.njobs;
  int64_t slice = blockIdx.x, nsl = gridDim.x;
  if (P.job_parallel) {
    while (j0 + 1 < P.njobs && (int)blockIdx.x >= P.first[j0 + 1]) ++j0;
    j1 = j0 + 1;
    slice = (int)blockIdx.x - P.first[j0];
    nsl = P.first[j0 + 1] - P.first[j0];
  }
  const int64_t n_my = P.ntiles > slice ? (P.ntiles - slice + nsl - 1) / nsl : 0;   // tiles slice, slice + nsl, ...
  const int64_t nst = 2 * n_my;   // 64-sample stages per job
  float* part = P.partial + (size_t)blockIdx.x * (P.job_parallel ? kCtaFloatsParallel : P.part_floats);

  if (threadIdx.x == 0) {
    if ((sbase & 1023u) != 0) { printf("tgtc style_wgrad: shared memory base not 1024-aligned\n"); __trap(); }
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
    if (threadIdx.x == 0) printf("tgtc style_wgrad: unexpected TMEM base\n");
    __trap();
  }

  if (warp == 0) {
    // ===================================================================== producer
    int stage = 0;
    uint32_t phase = 0;
    for (int j = j0; j < j1; ++j) {
      const SJob& jb = P.job[j];
      const uint32_t bytes = (uint32_t)(jb.a_blocks + jb.b_blocks) * 8192u;
      for (int64_t s = 0; s < nst; ++s) {
        const int64_t tile = slice + (s >> 1) * nsl;
        const size_t half = (size_t)(s & 1) * 8192;   // rows [64*half, 64*half+64) of a block are contiguous
        mbar_wait(bar(kWBarEmpty + stage), phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar(kWBarFull + stage), bytes);
          const uint32_t dst = sbase + stage * kWStageBytes;
          const uint8_t* pa = jb.a + (size_t)tile * jb.a_stride + half;
          const uint8_t* pb = jb.b + (size_t)tile * jb.b_stride + half;
          for (int b = 0; b < jb.a_blocks; ++b) bulk_g2s(dst + b * 8192, pa + (size_t)b * 16384, 8192, bar(kWBarFull + stage));
          for (int b = 0; b < jb.b_blocks; ++b) bulk_g2s(dst + 32768 + b * 8192, pb + (size_t)b * 16384, 8192, bar(kWBarFull + stage));
        }
        __syncwarp();
        if (++stage == kWStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    for (int j = j0; j < j1; ++j) {
      const SJob& jb = P.job[j];
      const int nm = jb.a_blocks >= 4 ? 2 : 1;                          // 128-row halves of the output
      const uint32_t a_lbo = jb.a_blocks >= 2 ? 8192u : 0u;            // a single 64-feature block is replicated (rows 64..127 unused)
      const uint32_t idesc = make_idesc_mn(128, jb.b_blocks * 64);
      if (j > j0) { mbar_wait_uniform(bar(kWBarAccFree), (uint32_t)((j - j0 - 1) & 1)); tc_fence_after(); }
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
    for (int j = j0; j < j1; ++j) {
      const SJob& jb = P.job[j];
      const int nm = jb.a_blocks >= 4 ? 2 : 1;
      const int ni = jb.b_blocks * 64;
      float* dst = part + (P.job_parallel ? 0 : jb.out_off);
      if (nst > 0) {
        mbar_wait(bar(kWBarAccDone), (uint32_t)((j - j0) & 1));
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
    // ===================================================================== column sums of A (biases; R rows for the latents)
    const int c = threadIdx.x - 6 * 32;          // feature column 0..255
    const int cb = c >> 6, cc = c & 63;
    int stage = 0;
    uint32_t phase = 0;
    const int64_t nstages_total = 2 * P.ntiles;
    for (int j = j0; j < j1; ++j) {
      const SJob& jb = P.job[j];
      const bool mine = jb.r_slot >= 0 && cb < jb.a_blocks;
      float total = 0.f;
      float* Rj = jb.r_slot >= 0 ? P.R + (size_t)jb.r_slot * nstages_total * 256 : nullptr;
      for (int64_t s = 0; s < nst; ++s) {
        mbar_wait(bar(kWBarFull + stage), phase);
        if (jb.r_slot >= 0) {
          float acc = 0.f;
          if (mine) {
            const uint8_t* blk = smem + stage * kWStageBytes + cb * 8192;
#pragma unroll 8
            for (int r = 0; r < 64; ++r) {
              const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((cc >> 3) ^ (r & 7))) << 4) + (cc & 7) * 2);
              acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(blk + off));
            }
            total += acc;
          }
          const int64_t sg = 2 * (slice + (s >> 1) * nsl) + (s & 1);
          Rj[sg * 256 + c] = acc;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kWBarEmpty + stage));
        if (++stage == kWStages) { stage = 0; phase ^= 1; }
      }
      part[P.job_parallel ? 256 * 256 + c : P.colsum_off + j * 256 + c] = total;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(0u), "r"(512));
  }
}

// ---------------------------------------------------------------------------
// partial -> flat gradients.  One region = a [rows x cols] sub-matrix of a parameter.
struct SRegion { int64_t part_off; int job; int colsum; int ni; int rows; int cols; int64_t dst_off; int dst_ld; int dst_col0; };
constexpr int kMaxRegions = 40;
struct SReduceParams {
  SRegion reg[kMaxRegions];
  int nreg;
  const float* partial;
  int64_t part_floats;
  int nparts;
  float* grads;
  int accumulate;
  int job_parallel;
  int first[kMaxJobs + 1];
};
__global__ void style_reduce_kernel(const __grid_constant__ SReduceParams P) {
  const SRegion& rg = P.reg[blockIdx.y];
  const int64_t total = (int64_t)rg.rows * rg.cols;
  // job-serial: every CTA holds a partial of this region at part_off; job-parallel: only the CTAs of the region's job do
  const int p0 = P.job_parallel ? P.first[rg.job] : 0, p1 = P.job_parallel ? P.first[rg.job + 1] : P.nparts;
  const int64_t stride = P.job_parallel ? kCtaFloatsParallel : P.part_floats;
  const int64_t base = P.job_parallel ? (rg.colsum ? 256 * 256 : 0) : rg.part_off;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / rg.cols), col = (int)(i % rg.cols);
    const float* src = P.partial + base + (int64_t)row * rg.ni + col;
    float acc = 0.f;
    for (int p = p0; p < p1; ++p) acc += src[(int64_t)p * stride];
    float* dst = P.grads + rg.dst_off + (int64_t)row * rg.dst_ld + rg.dst_col0 + col;
    *dst = P.accumulate ? *dst + acc : acc;
  }
}

// ---------------------------------------------------------------------------
// latents.  R [13][nstages][256]: slot l = column sums of dz of layer l over each 64-sample stage; a ray owns S/64 consecutive
// stages.  Slots 0..4 = module 1 (latent = lat1[ray]), 5..11 = module 2 layers 0..6, 12 = head (3 rows) (latent = mean(lat1[ray])).
//   d lat1[ray][k] = sum_{l<5} sum_j Rray_l[j] Wlat_l[j][k]  +  (1/32) sum_{l>=5} sum_j Rray_l[j] rowsum(Wlat_l[j][:])
constexpr int kLatRays = 8;   // rays per block
__global__ void __launch_bounds__(256) style_latgrad_kernel(const float* __restrict__ R, int64_t nstages, int spr, const float* __restrict__ wlat,
                                                            int64_t n_rays, float* __restrict__ dlat, int accumulate) {
  __shared__ float rr[kLatRays][256];
  __shared__ float red[kLatRays][8][33];
  const int64_t ray0 = (int64_t)blockIdx.x * kLatRays;
  const int k = threadIdx.x & 31, jg = threadIdx.x >> 5;
  float a1[kLatRays], a2[kLatRays];
#pragma unroll
  for (int q = 0; q < kLatRays; ++q) a1[q] = a2[q] = 0.f;
  for (int l = 0; l < 13; ++l) {
    // all of the layer's R rows of this block's rays in flight at once (spr <= 2: S is 64 or 128)
    float v[kLatRays][2];
#pragma unroll
    for (int q = 0; q < kLatRays; ++q) {
      const int64_t ray = ray0 + q;
      const bool ok = ray < n_rays && !(l == 12 && threadIdx.x >= 3);
#pragma unroll
      for (int s = 0; s < 2; ++s) v[q][s] = (ok && s < spr) ? __ldg(R + ((size_t)l * nstages + ray * spr + s) * 256 + threadIdx.x) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kLatRays; ++q) rr[q][threadIdx.x] = v[q][0] + v[q][1];
    __syncthreads();
    const float* w = wlat + (size_t)l * 256 * 33;
    float acc[kLatRays];
#pragma unroll
    for (int q = 0; q < kLatRays; ++q) acc[q] = 0.f;
    for (int jj = jg; jj < 256; jj += 8) {
      const float wv = w[jj * 32 + k];
#pragma unroll
      for (int q = 0; q < kLatRays; ++q) acc[q] = fmaf(rr[q][jj], wv, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < kLatRays; ++q) {
      if (l < 5) a1[q] += acc[q]; else a2[q] += acc[q];
    }
  }
#pragma unroll
  for (int q = 0; q < kLatRays; ++q) {
    float m = a2[q];   // summed over k: the gradient of the scalar mean
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    red[q][jg][k] = a1[q];
    if (k == 0) red[q][jg][32] = m;
  }
  __syncthreads();
  {
    const int q = threadIdx.x >> 5;   // 8 warps = 8 rays
    const int64_t ray = ray0 + q;
    if (ray < n_rays) {
      float g = 0.f, gm = 0.f;
      for (int i = 0; i < 8; ++i) { g += red[q][i][k]; gm += red[q][i][32]; }
      g += gm * (1.0f / 32.0f);
      float* d = dlat + ray * 32 + k;
      *d = accumulate ? *d + g : g;
    }
  }
}

// latent columns of every weight:  dW_l[j][lat0 + k] = sum_ray Rray_l[j] * lat_l(ray)[k]  -- a [256 x stages] x [stages x 32] product
// per layer, split over stage chunks (partials), then summed in chunk order (deterministic).
constexpr int kWlatChunks = 64;
__global__ void __launch_bounds__(256) style_wlat_partial_kernel(const float* __restrict__ R, int64_t nstages, int spr, const float* __restrict__ lat1,
                                                                 int64_t n_rays, float* __restrict__ part) {
  __shared__ float lat_s[8][32];
  const int l = blockIdx.y;
  const int j = threadIdx.x;
  const int64_t nvalid = n_rays * spr;
  const int64_t per = (nvalid + gridDim.x - 1) / gridDim.x;
  const int64_t s0 = (int64_t)blockIdx.x * per;
  const int64_t s1 = s0 + per < nvalid ? s0 + per : nvalid;
  float acc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) acc[k] = 0.f;
  const float* Rl = R + (size_t)l * nstages * 256 + j;
  for (int64_t s = s0; s < s1; s += 8) {
    __syncthreads();
    {
      const int q = threadIdx.x >> 5, k = threadIdx.x & 31;
      float v = 0.f;
      if (s + q < s1) {
        v = lat1[((s + q) / spr) * 32 + k];
        if (l >= 5) {   // module 2: every latent input is mean(lat1[ray])
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          v *= (1.0f / 32.0f);
        }
      }
      lat_s[q][k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (s + q < s1) {
        const float r = Rl[(s + q) * 256];
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = fmaf(r, lat_s[q][k], acc[k]);
      }
    }
  }
  float* dst = part + (((size_t)blockIdx.x * 13 + l) * 256 + j) * 32;
#pragma unroll
  for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(acc[k], acc[k + 1], acc[k + 2], acc[k + 3]);
}
struct SLatDst { int64_t dst_off; int ld; int lat0; int nout; };
struct SWlatParams { SLatDst d[13]; const float* part; int nchunks; float* grads; int accumulate; };
__global__ void style_wlat_reduce_kernel(const __grid_constant__ SWlatParams P) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over 13 * 256 * 32
  if (idx >= 13 * 256 * 32) return;
  const int l = idx / (256 * 32), j = (idx / 32) % 256, k = idx % 32;
  const SLatDst& d = P.d[l];
  if (j >= d.nout) return;
  float acc = 0.f;
  for (int c = 0; c < P.nchunks; ++c) acc += P.part[(size_t)c * 13 * 256 * 32 + idx];
  float* dst = P.grads + d.dst_off + (int64_t)j * d.ld + d.lat0 + k;
  *dst = P.accumulate ? *dst + acc : acc;
}

}  // namespace

// ---------------------------------------------------------------------------
// host side

// flat gradient layout = the 26 tensors of tgtc_set_style_weights in order: module 1 (W,b) x 5, module 2 (W,b) x 8
static const int kCIn[5] = {95, 288, 288, 288, 351};
static const int kWIn[8] = {607, 288, 288, 288, 351, 288, 288, 288};
struct StyleFlat { int64_t w[13], b[13]; int64_t total; };
static StyleFlat style_flat() {
  StyleFlat f;
  int64_t o = 0;
  for (int l = 0; l < 5; ++l) { f.w[l] = o; o += 256 * (int64_t)kCIn[l]; f.b[l] = o; o += 256; }
  for (int l = 0; l < 8; ++l) {
    const int nout = l < 7 ? 256 : 3;
    f.w[5 + l] = o; o += (int64_t)nout * kWIn[l]; f.b[5 + l] = o; o += nout;
  }
  f.total = o;
  return f;
}
size_t style_flat_floats() { return (size_t)style_flat().total; }

struct StyleJobPlan {
  int njobs;
  int out_off[kMaxJobs], mo[kMaxJobs], ni[kMaxJobs];
  int64_t colsum_off, part_floats;
};
// job order (see launch_style_wgrad): the plan only depends on the block counts
static const int kJobABlocks[kMaxJobs] = {1, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4};
static const int kJobBBlocks[kMaxJobs] = {4, 4, 4, 4, 1, 4, 4, 4, 4, 4, 1, 4, 1, 4, 4, 4, 1};
static StyleJobPlan style_plan() {
  StyleJobPlan p;
  p.njobs = kMaxJobs;
  int64_t o = 0;
  for (int j = 0; j < kMaxJobs; ++j) {
    p.mo[j] = kJobABlocks[j] >= 4 ? 256 : 128;
    p.ni[j] = kJobBBlocks[j] * 64;
    p.out_off[j] = (int)o;
    o += (int64_t)p.mo[j] * p.ni[j];
  }
  p.colsum_off = o;
  p.part_floats = o + (int64_t)kMaxJobs * 256;
  return p;
}
size_t style_partial_floats() { return (size_t)style_plan().part_floats; }
size_t style_wlat_part_floats() { return (size_t)kWlatChunks * 13 * 256 * 32; }

int launch_style_dgrad(tgtc_ctx* ctx, const float* rgbsigma, const float* d_rgbsigma, const StyleStash& stash, const StyleDz& dz, int64_t M,
                       cudaStream_t st) {
  if (M == 0) return TGTC_OK;
  SDgradParams P;
  P.blobT = ctx->style.blob_T;
  P.rgbsigma = reinterpret_cast<const float4*>(rgbsigma);
  P.d_rgbsigma = reinterpret_cast<const float4*>(d_rgbsigma);
  P.mask = stash.mask;
  P.dz = dz.dz; P.dhead = dz.dhead;
  P.M = M;
  P.ntiles = (M + kTileM - 1) / kTileM;
  static bool attr_set[64] = {};
  if (!attr_set[ctx->device & 63]) {
    TGTC_CUDA(cudaFuncSetAttribute(style_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    TGTC_CUDA(cudaFuncSetAttribute(style_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmemBytes));
    attr_set[ctx->device & 63] = true;
  }
  const int64_t nquads = (P.ntiles + 3) / 4;
  const int64_t max_pairs = ctx->num_sms / 2;
  const int grid = 2 * (int)(nquads < max_pairs ? nquads : max_pairs);
  style_dgrad_kernel<<<grid, kNumThreads, kSmemBytes, st>>>(P);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

static int g_style_wgrad_mode = -1;   // test hook: -1 automatic, 0 force job-serial, 1 force job-parallel
extern "C" void tgtc_debug_style_wgrad_mode(int m) { g_style_wgrad_mode = m; }

// dW of both style modules into grads (flat, style_flat layout) and d lat1 [n_rays][32]
int launch_style_wgrad(tgtc_ctx* ctx, const StyleStash& stash, const uint8_t* remap_img, const StyleDz& dz, const float* lat1,
                       int64_t n_rays, int S, float* partial, float* R, float* wlat_part, float* grads, int accumulate, float* dlat,
                       int dlat_accumulate, cudaStream_t st, cudaEvent_t ev_after_kernel) {
  const int64_t M = n_rays * S;
  if (M == 0) return TGTC_OK;
  TGTC_REQUIRE(S == 64 || S == 128, TGTC_ERR_UNSUPPORTED, "style wgrad needs 64 or 128 samples per ray");
  const int64_t ntiles = (M + kTileM - 1) / kTileM;
  const StyleJobPlan plan = style_plan();
  const StyleFlat flat = style_flat();
  SWgradParams P = {};
  P.njobs = plan.njobs;
  const int64_t DZ = (int64_t)kNumGemm * 65536, SW = 7 * 65536, SC = 5 * 65536;
  auto dzimg = [&](int g) { return dz.dz + (size_t)g * 65536; };
  auto job = [&](int j, const uint8_t* a, int64_t as, const uint8_t* b, int64_t bs, int r_slot) {
    P.job[j].a = a; P.job[j].a_stride = as; P.job[j].a_blocks = kJobABlocks[j];
    P.job[j].b = b; P.job[j].b_stride = bs; P.job[j].b_blocks = kJobBBlocks[j];
    P.job[j].r_slot = r_slot; P.job[j].out_off = plan.out_off[j];
  };
  // module 2 (dz image g: w6=0 .. w0=6), R slots 5..12
  job(0, dz.dhead, 16384, stash.w + 6 * 65536, SW, 12);          // head       x h_w6
  job(1, dzimg(0), DZ, stash.w + 5 * 65536, SW, 11);             // W6         x h_w5
  job(2, dzimg(1), DZ, stash.w + 4 * 65536, SW, 10);             // W5         x h_w4
  job(3, dzimg(2), DZ, stash.w + 3 * 65536, SW, 9);              // W4 hidden  x h_w3
  job(4, dzimg(2), DZ, stash.pe, 16384, -1);                     // W4 x-columns
  job(5, dzimg(3), DZ, stash.w + 2 * 65536, SW, 8);              // W3
  job(6, dzimg(4), DZ, stash.w + 1 * 65536, SW, 7);              // W2
  job(7, dzimg(5), DZ, stash.w + 0 * 65536, SW, 6);              // W1
  job(8, dzimg(6), DZ, remap_img, 65536, 5);                     // W0 base_remap columns
  job(9, dzimg(6), DZ, stash.c + 4 * 65536, SC, -1);             // W0 concat_features columns
  job(10, dzimg(6), DZ, stash.pe, 16384, -1);                    // W0 x-columns
  // module 1 (dz image g: c4=7 .. c0=11), R slots 0..4
  job(11, dzimg(7), DZ, stash.c + 3 * 65536, SC, 4);             // C4 hidden
  job(12, dzimg(7), DZ, stash.pe, 16384, -1);                    // C4 x-columns
  job(13, dzimg(8), DZ, stash.c + 2 * 65536, SC, 3);             // C3
  job(14, dzimg(9), DZ, stash.c + 1 * 65536, SC, 2);             // C2
  job(15, dzimg(10), DZ, stash.c + 0 * 65536, SC, 1);            // C1
  job(16, dzimg(11), DZ, stash.pe, 16384, 0);                    // C0 x-columns
  P.partial = partial;
  P.part_floats = plan.part_floats;
  P.colsum_off = plan.colsum_off;
  P.R = R;
  P.ntiles = ntiles;
  // small batches: one job per CTA (CTAs per job proportional to the job's bytes per stage)
  P.job_parallel = (ntiles <= 1536 && ctx->num_sms >= kMaxJobs) ? 1 : 0;
  if (g_style_wgrad_mode >= 0 && ctx->num_sms >= kMaxJobs) P.job_parallel = g_style_wgrad_mode;
  int grid = (int)(ntiles < ctx->num_sms ? ntiles : ctx->num_sms);
  if (P.job_parallel) {
    grid = ctx->num_sms;
    int k[kMaxJobs], total = 0, csum = 0;
    for (int j = 0; j < kMaxJobs; ++j) csum += kJobABlocks[j] + kJobBBlocks[j];
    for (int j = 0; j < kMaxJobs; ++j) { k[j] = grid * (kJobABlocks[j] + kJobBBlocks[j]) / csum; if (k[j] < 1) k[j] = 1; total += k[j]; }
    while (total < grid) {   // the remaining CTAs go to the jobs with the most bytes per CTA
      int best = 0;
      for (int j = 1; j < kMaxJobs; ++j)
        if ((kJobABlocks[j] + kJobBBlocks[j]) * k[best] > (kJobABlocks[best] + kJobBBlocks[best]) * k[j]) best = j;
      ++k[best]; ++total;
    }
    while (total > grid) {
      int best = -1;
      for (int j = 0; j < kMaxJobs; ++j)
        if (k[j] > 1 && (best < 0 || (kJobABlocks[j] + kJobBBlocks[j]) * k[best] < (kJobABlocks[best] + kJobBBlocks[best]) * k[j])) best = j;
      --k[best]; --total;
    }
    P.first[0] = 0;
    for (int j = 0; j < kMaxJobs; ++j) P.first[j + 1] = P.first[j] + k[j];
  }
  style_wgrad_kernel<<<grid, kWThreads, kWSmemBytes, st>>>(P);
  TGTC_LAUNCH_CHECK(ctx);
  if (ev_after_kernel != nullptr) TGTC_CUDA(cudaEventRecord(ev_after_kernel, st));   // profiling: the main kernel alone

  // regions: (job, parameter, destination column offset, valid columns, rows)
  SReduceParams Q = {};
  int nr = 0;
  auto region = [&](int j, int64_t dst_off, int ld, int col0, int cols, int rows) {
    Q.reg[nr++] = {(int64_t)plan.out_off[j], j, 0, plan.ni[j], rows, cols, dst_off, ld, col0};
  };
  auto bias = [&](int j, int64_t dst_off, int rows) { Q.reg[nr++] = {plan.colsum_off + (int64_t)j * 256, j, 1, 1, rows, 1, dst_off, 1, 0}; };
  const int W0 = 5;   // index of module 2's first layer in flat.w / flat.b
  region(0, flat.w[W0 + 7], 288, 0, 256, 3);      bias(0, flat.b[W0 + 7], 3);
  region(1, flat.w[W0 + 6], 288, 0, 256, 256);    bias(1, flat.b[W0 + 6], 256);
  region(2, flat.w[W0 + 5], 288, 0, 256, 256);    bias(2, flat.b[W0 + 5], 256);
  region(3, flat.w[W0 + 4], 351, 0, 256, 256);    bias(3, flat.b[W0 + 4], 256);
  region(4, flat.w[W0 + 4], 351, 288, 63, 256);
  region(5, flat.w[W0 + 3], 288, 0, 256, 256);    bias(5, flat.b[W0 + 3], 256);
  region(6, flat.w[W0 + 2], 288, 0, 256, 256);    bias(6, flat.b[W0 + 2], 256);
  region(7, flat.w[W0 + 1], 288, 0, 256, 256);    bias(7, flat.b[W0 + 1], 256);
  region(8, flat.w[W0 + 0], 607, 0, 256, 256);    bias(8, flat.b[W0 + 0], 256);
  region(9, flat.w[W0 + 0], 607, 256, 256, 256);
  region(10, flat.w[W0 + 0], 607, 512, 63, 256);
  region(11, flat.w[4], 351, 0, 256, 256);        bias(11, flat.b[4], 256);
  region(12, flat.w[4], 351, 288, 63, 256);
  region(13, flat.w[3], 288, 0, 256, 256);        bias(13, flat.b[3], 256);
  region(14, flat.w[2], 288, 0, 256, 256);        bias(14, flat.b[2], 256);
  region(15, flat.w[1], 288, 0, 256, 256);        bias(15, flat.b[1], 256);
  region(16, flat.w[0], 95, 0, 63, 256);          bias(16, flat.b[0], 256);
  Q.nreg = nr;
  Q.partial = partial;
  Q.part_floats = plan.part_floats;
  Q.nparts = grid;
  Q.grads = grads;
  Q.accumulate = accumulate;
  Q.job_parallel = P.job_parallel;
  for (int j = 0; j <= kMaxJobs; ++j) Q.first[j] = P.first[j];
  style_reduce_kernel<<<dim3(64, nr), 256, 0, st>>>(Q);
  TGTC_LAUNCH_CHECK(ctx);

  // latent columns and latent gradients from the R rows
  static const int clat[5] = {63, 256, 256, 256, 256};
  static const int wlat0[8] = {575, 256, 256, 256, 256, 256, 256, 256};
  style_wlat_partial_kernel<<<dim3(kWlatChunks, 13), 256, 0, st>>>(R, 2 * ntiles, S / 64, lat1, n_rays, wlat_part);
  TGTC_LAUNCH_CHECK(ctx);
  SWlatParams L = {};
  for (int l = 0; l < 5; ++l) L.d[l] = {flat.w[l], kCIn[l], clat[l], 256};
  for (int l = 0; l < 8; ++l) L.d[5 + l] = {flat.w[5 + l], kWIn[l], wlat0[l], l < 7 ? 256 : 3};
  L.part = wlat_part; L.nchunks = kWlatChunks; L.grads = grads; L.accumulate = accumulate;
  style_wlat_reduce_kernel<<<(13 * 256 * 32 + 255) / 256, 256, 0, st>>>(L);
  TGTC_LAUNCH_CHECK(ctx);
  if (dlat != nullptr) {
    style_latgrad_kernel<<<(unsigned)((n_rays + kLatRays - 1) / kLatRays), 256, 0, st>>>(R, 2 * ntiles, S / 64, ctx->style.wlat, n_rays, dlat,
                                                                                        dlat_accumulate);
    TGTC_LAUNCH_CHECK(ctx);
  }
  return TGTC_OK;
}
