// Per-ray style head -- models.StyleMLP_before_concat (models.py:120-147) and
// models.StyleMLP_Wild_multilayers (models.py:149-180) as called by render_style (rendering.py:118-178),
// bf16 operands on tcgen05 (CTA pairs), fp32 accumulation in TMEM.
//
// One table-driven chain kernel runs either module on 128-sample tiles (two slots per CTA, weight ring, epilogue warps:
// the structure of mlp_tc.cu).  A layer is a sequence of K segments:
//     PE    one 64-column chunk: the positional-encoding tile (recomputed from the rays, never stored)
//     ACT   four chunks: the activation tile of the previous layer (in place, shared memory)
//     IMGk  four chunks: a tile image from HBM (base_remap written by the NeRF trunk kernel, concat_features written by the
//           first module) staged into the activation buffer by the slot's tile-mover thread with one 64 KB bulk copy
// and an output kind: ACT (bias + ReLU -> next A operand), ACT+IMG (also one 64 KB bulk store of the tile image), HEAD
// (bias + ReLU, then the 3-row fp32 output layer + sigmoid on CUDA cores -> rgb).
//     module 1: [PE] [ACT] [ACT] [ACT] [PE,ACT]->concat_features image
//     module 2: [PE,IMG0=base_remap,IMG1=concat_features] [ACT] [ACT] [ACT] [PE,ACT] [ACT] [ACT]->HEAD
// The 32-d latent inputs of every layer are constant per (style, frame): their products with the latent columns of the
// weights are folded into per-call effective biases by style_bias_kernel (fp32).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace tcptx;

constexpr int kTileM = 128;
constexpr int kStages = 3;
constexpr int kStageBytes = 16384;
constexpr int kNumThreads = 512;   // w0 weights, w1 MMA / forwarder, w2-5 PE producers, w6-13 epilogue, w14-15 tile movers
constexpr int kPeWarp0 = 2, kEpiWarp0 = 6, kMoveWarp0 = 14;
constexpr int kNumEpiThreads = 256, kNumPeThreads = 128;
constexpr int kMaxLayers = 8;

enum { SEG_PE = 1, SEG_ACT = 2, SEG_IMG0 = 3, SEG_IMG1 = 4 };
enum { OUT_ACT = 0, OUT_ACT_IMG = 1, OUT_HEAD = 2 };

struct ChainLayer {
  uint8_t seg[3];
  uint8_t nseg;
  uint8_t out;
  uint8_t pe_last;   // this layer's PE segment is the tile's last use of the PE buffer
};
struct ChainParams {
  ChainLayer layer[kMaxLayers];
  int nlayers;
  const uint8_t* blob;       // chunks [256 x 64] bf16 SW128 in consumption order (first half = rows 0..127)
  const float* bias;         // [nlayers][256] effective biases (bias + latent columns . latent)
  const float* head_w;       // [3][256] output layer (HEAD), fp32
  const float* head_b;       // [3] effective
  const float* rays_o;
  const float* rays_d;
  const float* ts;           // [n_rays,S] or nullptr -> uniform coarse positions
  float t_scale, t_near;
  int S;
  int64_t M;
  int64_t ntiles;
  const uint8_t* pe_img;     // explicit-input mode (stage shims): [ntiles][16 KB] PE tile images instead of rays (rays_o == nullptr)
  const uint8_t* img0;       // [ntiles][64 KB]
  const uint8_t* img1;
  uint8_t* img_out;          // [ntiles][64 KB]
  float* rgbsigma;           // [M][4]: HEAD writes .xyz
  int64_t img0_stride, img1_stride;   // bytes between consecutive tiles of img0 / img1
  // training forward (kTrain): per-ray latents and the activation stash the style backward reads
  const float* bias_rays;    // [n_rays][bias_ray_stride] effective biases of this module's layers (then the head's 3), per ray
  int64_t bias_ray_stride;   // floats
  uint8_t* stash;            // [ntiles][stash_layers][64 KB]: the post-ReLU output image of every layer
  int stash_layers;
  uint32_t* mask;            // [ntiles][mask_layers][8][128] ReLU mask words (common.cuh: TcStash.mask), this module at mask_slot0
  int mask_layers, mask_slot0;
  uint8_t* stash_pe;         // [ntiles][16 KB] positional-encoding tile image, or nullptr
  int dbg_flags;             // timing experiments (results garbage): 1 no mask words, 4 no image store, 8 no drain wait
};

constexpr int kOffAct = 0;
constexpr int kActBytes = 65536;
constexpr int kOffPe = kOffAct + 2 * kActBytes;
constexpr int kPeBytes = 16384;
constexpr int kOffW = kOffPe + 2 * kPeBytes;
constexpr int kOffBias = kOffW + kStages * kStageBytes;     // 8 x 256 fp32
constexpr int kOffHeadW = kOffBias + kMaxLayers * 256 * 4;  // 3 x 256 fp32
constexpr int kOffHeadPart = kOffHeadW + 3 * 256 * 4;       // [128][4] fp32
constexpr int kOffBars = kOffHeadPart + 128 * 4 * 4;
constexpr int kBarWFull = 0, kBarWEmpty = kStages, kBarPeReady = 2 * kStages /*leader, count 8*/, kBarPeFree = kBarPeReady + 2,
              kBarActReady = kBarPeFree + 2, kBarAccFull = kBarActReady + 2, kBarAFull = kBarAccFull + 2 /*leader: both CTAs staged*/,
              kBarAFree = kBarAFull + 2, kBarOutDone = kBarAFree + 2, kBarOutFree = kBarOutDone + 2 /*leader*/,
              kBarLoad = kBarOutFree + 2 /*mover's own bulk loads*/, kBarSlotFree = kBarLoad + 2 /*training: store drained*/,
              kNumBars = kBarSlotFree + 2;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");

__device__ __forceinline__ int64_t pair_tile(int64_t it, int t, uint32_t rank) {
  const int64_t quad = (int64_t)(blockIdx.x >> 1) + it * (int64_t)(gridDim.x >> 1);
  return quad * 4 + 2 * (int64_t)rank + t;
}
__device__ __forceinline__ int seg_chunks(int s) { return s == SEG_PE ? 1 : 4; }


// kTrain: the training forward of Style_train (train_tgtcs.py:311-483) -- per-ray latents (effective biases per ray from
// global memory), every layer's output image and ReLU mask words stashed for the backward, the PE tile stashed once.
// kF16 (inference only): fp16 instead of bf16 operands -- weights image, PE tile, activations and the feature tile images that
// travel between the trunk / module 1 / module 2 launches (TGTC_MLP_F16).
// kRayBias (inference): per-ray latents inside one call -- the per-ray effective-bias path of the training instantiation
// (bias_rays) without its stash, so a batch may mix (style, frame) latents (rendering.py:125-127).
template <bool kTrain, bool kF16 = false, bool kRayBias = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1) mlp_chain_kernel(const __grid_constant__ ChainParams P) {
  static_assert(!(kTrain && kF16), "the Style_train stash / backward kernels are bf16");
  static_assert(!(kTrain && kRayBias), "the training instantiation already takes per-ray biases");
  constexpr bool kPerRay = kTrain || kRayBias;
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
  const int nl = P.nlayers;
  int chunks_per_tile = 0, imgs_per_tile = 0;
  for (int l = 0; l < nl; ++l)
    for (int s = 0; s < P.layer[l].nseg; ++s) {
      chunks_per_tile += seg_chunks(P.layer[l].seg[s]);
      imgs_per_tile += P.layer[l].seg[s] >= SEG_IMG0 ? 1 : 0;
    }
  const bool has_out_img = !kTrain && P.layer[nl - 1].out == OUT_ACT_IMG;

  if (threadIdx.x == 0) {
    if ((sbase & 1023u) != 0) { printf("tgtc mlp_chain: shared memory base not 1024-aligned\n"); __trap(); }
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(kBarWFull + s), rank == 0 ? 2 : 1); mbar_init(bar(kBarWEmpty + s), 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar(kBarPeReady + t), 2 * (kNumPeThreads / 32));
      mbar_init(bar(kBarPeFree + t), 1);
      mbar_init(bar(kBarActReady + t), 2 * (kNumEpiThreads / 32));
      mbar_init(bar(kBarAccFull + t), 1);
      mbar_init(bar(kBarAFull + t), 2);                         // one remote arrive per CTA's mover after its copy landed
      mbar_init(bar(kBarAFree + t), 1);
      mbar_init(bar(kBarOutDone + t), kNumEpiThreads / 32);
      mbar_init(bar(kBarOutFree + t), 2);
      mbar_init(bar(kBarLoad + t), 1);
      mbar_init(bar(kBarSlotFree + t), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(sbase + kOffTmemPtr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
  }
  {
    float* dst = reinterpret_cast<float*>(smem + kOffBias);
    if (!kPerRay)
      for (int i = threadIdx.x; i < nl * 256; i += kNumThreads) dst[i] = P.bias[i];
    if (P.head_w != nullptr) {
      float* hw = reinterpret_cast<float*>(smem + kOffHeadW);
      for (int i = threadIdx.x; i < 3 * 256; i += kNumThreads) hw[i] = P.head_w[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (*reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr) != 0u) {
    if (threadIdx.x == 0) printf("tgtc mlp_chain: unexpected TMEM base\n");
    __trap();
  }
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0) {
    // ===================================================================== weight producer
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t it = 0; it < iters; ++it) {
      int c0 = 0;
      for (int l = 0; l < nl; ++l) {
        int nch = 0;
        for (int s = 0; s < P.layer[l].nseg; ++s) nch += seg_chunks(P.layer[l].seg[s]);
        for (int t = 0; t < 2; ++t) {
          for (int c = 0; c < nch; ++c) {
            mbar_wait(bar(kBarWEmpty + stage), phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(bar(kBarWFull + stage), kStageBytes);
              bulk_g2s(sbase + kOffW + stage * kStageBytes, P.blob + (size_t)(c0 + c) * 2 * kStageBytes + (size_t)rank * kStageBytes,
                       kStageBytes, bar(kBarWFull + stage));
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
        c0 += nch;
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ===================================================================== peer: forward "my half of the stage landed"
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t leader_wfull = mapa_cluster(bar(kBarWFull), 0);
    for (int64_t it = 0; it < iters; ++it) {
      for (int c = 0; c < 2 * chunks_per_tile; ++c) {
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
    uint32_t act_par[2] = {0, 0}, pe_par[2] = {0, 0}, af_par[2] = {0, 0}, of_par[2] = {0, 0};
    const uint32_t w_lo0 = (((sbase + kOffW) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t idesc = make_idesc_op<kF16>(2 * kTileM, 256);
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
      for (int l = 0; l < nl; ++l) {
        const ChainLayer L = P.layer[l];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(256 * t);
          const uint32_t pe_lo = (((sbase + kOffPe + t * kPeBytes) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t act_lo = (((sbase + kOffAct + t * kActBytes) & 0x3FFFFu) >> 4) | (1u << 16);
          // the accumulator of this slot is drained / the activation tile is written (epilogue of the previous layer)
          if (l > 0 || it > 0) { mbar_wait_uniform(bar(kBarActReady + t), act_par[t]); act_par[t] ^= 1; }
          // first layer of a tile: the previous tile's image store out of act[t] must have finished reading it before this
          // tile's first epilogue overwrites it
          if (l == 0 && it > 0 && has_out_img) { mbar_wait_uniform(bar(kBarOutFree + t), of_par[t]); of_par[t] ^= 1; }
          tc_fence_after();
          uint32_t acc = 0u;
          for (int s = 0; s < L.nseg; ++s) {
            const int sg = L.seg[s];
            if (sg == SEG_PE) {
              mbar_wait_uniform(bar(kBarPeReady + t), pe_par[t]);   // completes once per tile; the parity advances at its last use
              tc_fence_after();
              issue_chunk(d_tmem, pe_lo, acc);
              acc = 1u;
              if (L.pe_last) {
                pe_par[t] ^= 1;
                if (elect_one()) umma_commit(bar(kBarPeFree + t));
                __syncwarp();
              }
            } else {
              if (sg >= SEG_IMG0) { mbar_wait_uniform(bar(kBarAFull + t), af_par[t]); af_par[t] ^= 1; tc_fence_after(); }
#pragma unroll
              for (int c = 0; c < 4; ++c) { issue_chunk(d_tmem, act_lo + 1024u * (uint32_t)c, acc); acc = 1u; }
              // the image just consumed may be replaced by the next staged image (or, at the end of the tile, by the next tile's)
              const bool next_is_img = (s + 1 < L.nseg && L.seg[s + 1] >= SEG_IMG0);
              if (next_is_img || (imgs_per_tile > 0 && l == nl - 1 && s == L.nseg - 1)) {
                if (elect_one()) umma_commit(bar(kBarAFree + t));
                __syncwarp();
              }
            }
          }
          if (elect_one()) umma_commit(bar(kBarAccFull + t));
          __syncwarp();
        }
      }
    }
  } else if (warp < kEpiWarp0) {
    // ===================================================================== PE producers (one thread per tile row)
    const int r = (warp - kPeWarp0) * 32 + lane;
    const int S = P.S;
    const uint32_t leader_peready = mapa_cluster(bar(kBarPeReady), 0);
    for (int64_t it = 0; it < iters; ++it) {
      for (int t = 0; t < 2; ++t) {
        const int64_t tile = pair_tile(it, t, rank);
        if (it > 0) mbar_wait_relaxed(bar(kBarPeFree + t), (uint32_t)((it - 1) & 1), 128);
        int64_t m = tile * kTileM + r;
        if (m >= P.M) m = P.M - 1;
        const uint32_t prow = sbase + kOffPe + t * kPeBytes + (r >> 3) * 1024 + (r & 7) * 128;
        if (P.pe_img != nullptr) {
          // explicit-input mode: the caller's embedded points, already converted to tile images -- copy this row's 128 bytes
          const uint4* src = reinterpret_cast<const uint4*>(P.pe_img + (size_t)(tile < P.ntiles ? tile : P.ntiles - 1) * 16384 + (r >> 3) * 1024 + (r & 7) * 128);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint4 q = __ldg(src + ch);
            st_shared_v4(prow + (ch << 4), q.x, q.y, q.z, q.w);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader_peready + 8u * t);
          continue;
        }
        const int64_t ray = m / S;
        const int k = (int)(m - ray * S);
        const float tt = P.ts != nullptr ? P.ts[m] : coarse_t(k, S, P.t_scale, P.t_near);
        float x[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(P.rays_o[ray * 3 + c], __fmul_rn(tt, P.rays_d[ray * 3 + c]));
        float e[64];
        e[0] = x[0]; e[1] = x[1]; e[2] = x[2];
#pragma unroll
        for (int f = 0; f < 10; ++f) {
          const float fr = (float)(1 << f);
#pragma unroll
          for (int a = 0; a < 3; ++a) fast_sincos(__fmul_rn(x[a], fr), &e[3 + 6 * f + a], &e[3 + 6 * f + 3 + a]);
        }
        e[63] = 0.f;
        uint8_t* gpe = nullptr;
        if constexpr (kTrain) {
          if (P.stash_pe != nullptr && tile < P.ntiles) gpe = P.stash_pe + (size_t)tile * 16384 + (r >> 3) * 1024 + (r & 7) * 128;
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t q0 = pack_op<kF16>(e[8 * ch + 0], e[8 * ch + 1]), q1 = pack_op<kF16>(e[8 * ch + 2], e[8 * ch + 3]),
                         q2 = pack_op<kF16>(e[8 * ch + 4], e[8 * ch + 5]), q3 = pack_op<kF16>(e[8 * ch + 6], e[8 * ch + 7]);
          st_shared_v4(prow + ((ch ^ (r & 7)) << 4), q0, q1, q2, q3);
          if (gpe != nullptr) *reinterpret_cast<uint4*>(gpe + ((ch ^ (r & 7)) << 4)) = make_uint4(q0, q1, q2, q3);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_peready + 8u * t);
      }
    }
  } else if (warp >= kMoveWarp0) {
    // ===================================================================== tile movers: one thread per slot
    if (lane == 0) {
      const int t = warp - kMoveWarp0;
      const uint32_t leader_afull = mapa_cluster(bar(kBarAFull), 0) + 8u * t;
      const uint32_t leader_outfree = mapa_cluster(bar(kBarOutFree), 0) + 8u * t;
      uint32_t afree_par = 0, od_par = 0, ld_par = 0;
      bool first_stage = true;
      if (kTrain) mbar_arrive(bar(kBarSlotFree + t));   // the slot starts out free
      for (int64_t it = 0; it < iters; ++it) {
        const int64_t tile = pair_tile(it, t, rank);
        const bool tile_ok = tile < P.ntiles;
        for (int l = 0; l < nl; ++l) {
          const ChainLayer L = P.layer[l];
          for (int s = 0; s < L.nseg; ++s) {
            if (L.seg[s] < SEG_IMG0) continue;
            if (!first_stage) { mbar_wait(bar(kBarAFree + t), afree_par); afree_par ^= 1; }
            first_stage = false;
            const uint8_t* src = L.seg[s] == SEG_IMG0 ? P.img0 + (size_t)(tile_ok ? tile : 0) * P.img0_stride
                                                      : P.img1 + (size_t)(tile_ok ? tile : 0) * P.img1_stride;
            // one 64 KB bulk copy into act[t]; when it has landed here, tell the leader's MMA warp (it needs both CTAs' tiles)
            mbar_arrive_expect_tx(bar(kBarLoad + t), 65536u);
            bulk_g2s(sbase + kOffAct + t * kActBytes, src, 65536u, bar(kBarLoad + t));
            mbar_wait(bar(kBarLoad + t), ld_par); ld_par ^= 1;
            mbar_arrive_cluster(leader_afull);
          }
          if constexpr (kTrain) {
            // every layer's output image goes to the stash; the epilogue of the next layer waits for the store to have drained
            mbar_wait(bar(kBarOutDone + t), od_par); od_par ^= 1;
            if (tile_ok && !(P.dbg_flags & 4)) {
              bulk_s2g(P.stash + ((size_t)tile * P.stash_layers + l) * 65536, sbase + kOffAct + t * kActBytes, 65536u);
              bulk_commit_group();
            }
            bulk_wait_read0();
            mbar_arrive(bar(kBarSlotFree + t));
          } else if (L.out == OUT_ACT_IMG) {
            mbar_wait(bar(kBarOutDone + t), od_par); od_par ^= 1;
            if (tile_ok) {
              bulk_s2g(P.img_out + (size_t)tile * 65536, sbase + kOffAct + t * kActBytes, 65536u);
              bulk_commit_group();
            }
            bulk_wait_read0();
            mbar_arrive_cluster(leader_outfree);
          }
        }
      }
      bulk_wait_all0();
    }
  } else {
    // ===================================================================== epilogue warps
    const int q = warp & 3;
    const int hc = (warp - kEpiWarp0) >> 2;
    const int row = q * 32 + lane;
    const float* bias_s = reinterpret_cast<const float*>(smem + kOffBias);
    const float* headw_s = reinterpret_cast<const float*>(smem + kOffHeadW);
    float* part_s = reinterpret_cast<float*>(smem + kOffHeadPart);
    const uint32_t leader_actready = mapa_cluster(bar(kBarActReady), 0);
    const uint32_t rx = (uint32_t)(row & 7) << 4;
    uint32_t acc_par[2] = {0, 0}, sf_par[2] = {0, 0};
    float hb[3] = {0.f, 0.f, 0.f};
    if (P.head_b != nullptr) { hb[0] = P.head_b[0]; hb[1] = P.head_b[1]; hb[2] = P.head_b[2]; }
    // kTrain: the per-ray effective biases of a step (one layer of one slot: the two half-tiles' rays x 256 outputs) are staged
    // in shared memory one step ahead -- [slot][parity][half][256] fp32 in the (otherwise unused) bias area; fetching them
    // from global memory inside the epilogue cost 35 % of the kernel
    float* bias_w = reinterpret_cast<float*>(smem + kOffBias);
    const int e256 = threadIdx.x - kEpiWarp0 * 32;
    uint32_t use[2] = {0, 0};
    auto bias_rows = [&](int64_t it_, int l_, int t_, float& b0, float& b1) {
      const int64_t tile_ = pair_tile(it_, t_, rank);
      int64_t m0 = tile_ * kTileM, m1 = tile_ * kTileM + 64;
      if (m0 >= P.M) m0 = P.M - 1;
      if (m1 >= P.M) m1 = P.M - 1;
      b0 = __ldg(P.bias_rays + (m0 / P.S) * P.bias_ray_stride + l_ * 256 + e256);
      b1 = __ldg(P.bias_rays + (m1 / P.S) * P.bias_ray_stride + l_ * 256 + e256);
    };
    if constexpr (kPerRay) {
      if (iters > 0) {
        for (int t = 0; t < 2; ++t) {
          float b0, b1;
          bias_rows(0, 0, t, b0, b1);
          bias_w[((t * 2 + 0) * 2 + 0) * 256 + e256] = b0;
          bias_w[((t * 2 + 0) * 2 + 1) * 256 + e256] = b1;
        }
      }
    }
    for (int64_t it = 0; it < iters; ++it) {
      for (int l = 0; l < nl; ++l) {
        const int okind = P.layer[l].out;
        for (int t = 0; t < 2; ++t) {
          const int64_t tile = pair_tile(it, t, rank);
          const int64_t m = tile * kTileM + row;
          float nb0 = 0.f, nb1 = 0.f;
          const bool has_next = kPerRay && !(it == iters - 1 && l == nl - 1);
          if constexpr (kPerRay) {
            if (has_next) bias_rows(l + 1 < nl ? it : it + 1, l + 1 < nl ? l + 1 : 0, t, nb0, nb1);
            named_bar_sync(3, kNumEpiThreads);   // this step's rows (written at the end of the slot's previous step) are visible
          }
          mbar_wait(bar(kBarAccFull + t), acc_par[t]); acc_par[t] ^= 1;
          tc_fence_after();
          const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(256 * t) + (uint32_t)(hc * 128);
          const uint32_t arow = sbase + kOffAct + t * kActBytes + (row >> 3) * 1024 + (row & 7) * 128 + hc * 2 * 16384;
          float p0 = 0.f, p1 = 0.f, p2 = 0.f;
          if constexpr (kPerRay) {
            // per-ray effective biases (the latent columns folded per ray); all 32 rows of a warp belong to one ray
            const float* bl = bias_w + ((t * 2 + (int)(use[t] & 1)) * 2 + (row >> 6)) * 256 + hc * 128;
            uint32_t* mrow = nullptr;
            if constexpr (kTrain) {
              mrow = (tile < P.ntiles && !(P.dbg_flags & 1)) ? P.mask + (((size_t)tile * P.mask_layers + P.mask_slot0 + l) * 8 + hc * 4) * 128 + row : nullptr;
              if (!(P.dbg_flags & 8)) { mbar_wait(bar(kBarSlotFree + t), sf_par[t]); sf_par[t] ^= 1; }   // the previous image store has drained act[t]
            }
#pragma unroll 1
            for (int blk = 0; blk < 4; ++blk) {
              uint32_t v[32];
              tmem_ld32(tcol + blk * 32, v);
              tmem_ld_wait_dep(v);
              const uint32_t kb = arow + (uint32_t)(blk >> 1) * 16384u;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(bl + blk * 32 + 4 * j);
                add2(v[4 * j + 0], v[4 * j + 1], b.x, b.y);
                add2(v[4 * j + 2], v[4 * j + 3], b.z, b.w);
              }
              if (mrow != nullptr) mrow[blk * 128] = sign_mask32(v);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (okind == OUT_HEAD) {
                  const float* w = headw_s + hc * 128 + blk * 32 + 8 * j;
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const float h = fmaxf(__uint_as_float(v[8 * j + e]), 0.f);
                    p0 = fmaf(h, w[e], p0);
                    p1 = fmaf(h, w[256 + e], p1);
                    p2 = fmaf(h, w[512 + e], p2);
                  }
                }
                const uint32_t dst = kb + ((uint32_t)((((blk & 1) * 4) + j) << 4) ^ rx);
                st_shared_v4(dst, pack_op_relu<kF16>(v[8 * j + 0], v[8 * j + 1]), pack_op_relu<kF16>(v[8 * j + 2], v[8 * j + 3]),
                             pack_op_relu<kF16>(v[8 * j + 4], v[8 * j + 5]), pack_op_relu<kF16>(v[8 * j + 6], v[8 * j + 7]));
              }
            }
            if (has_next) {
              bias_w[((t * 2 + (int)((use[t] + 1) & 1)) * 2 + 0) * 256 + e256] = nb0;
              bias_w[((t * 2 + (int)((use[t] + 1) & 1)) * 2 + 1) * 256 + e256] = nb1;
            }
            ++use[t];
          } else if (okind != OUT_HEAD) {
            // hidden layer: packed bias add + ReLU + bf16 pack, the TMEM load of block b+1 in flight while block b is processed
            const float* bl = bias_s + l * 256 + hc * 128;
            // every group's bias loads are issued BEFORE the previous group's activation store (ptxas cannot tell the bias table
            // from the activation buffer: a load written after the store waits behind it and its latency is exposed -- mlp_tc.cu)
            float4 pb0 = *reinterpret_cast<const float4*>(bl), pb1 = *reinterpret_cast<const float4*>(bl + 4);
            auto blk_store = [&](uint32_t (&v)[32], int blk) {
              const uint32_t kb = arow + (uint32_t)(blk >> 1) * 16384u;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int c = blk * 32 + 8 * j;
                const float4 b0 = pb0, b1 = pb1;
                if (c + 8 < 128) {
                  pb0 = *reinterpret_cast<const float4*>(bl + c + 8);
                  pb1 = *reinterpret_cast<const float4*>(bl + c + 12);
                }
                add2(v[8 * j + 0], v[8 * j + 1], b0.x, b0.y);
                add2(v[8 * j + 2], v[8 * j + 3], b0.z, b0.w);
                add2(v[8 * j + 4], v[8 * j + 5], b1.x, b1.y);
                add2(v[8 * j + 6], v[8 * j + 7], b1.z, b1.w);
                const uint32_t dst = kb + ((uint32_t)((((blk & 1) * 4) + j) << 4) ^ rx);
                st_shared_v4(dst, pack_op_relu<kF16>(v[8 * j + 0], v[8 * j + 1]), pack_op_relu<kF16>(v[8 * j + 2], v[8 * j + 3]),
                             pack_op_relu<kF16>(v[8 * j + 4], v[8 * j + 5]), pack_op_relu<kF16>(v[8 * j + 6], v[8 * j + 7]));
              }
            };
            uint32_t va[32], vb[32];
            tmem_ld32(tcol, va);
            tmem_ld_wait_dep(va);
            tmem_ld32(tcol + 32, vb);
            blk_store(va, 0);
            tmem_ld_wait_dep(vb);
            tmem_ld32(tcol + 64, va);
            blk_store(vb, 1);
            tmem_ld_wait_dep(va);
            tmem_ld32(tcol + 96, vb);
            blk_store(va, 2);
            tmem_ld_wait_dep(vb);
            blk_store(vb, 3);
          } else {
          const float* bl = bias_s + l * 256 + hc * 128;
#pragma unroll 1
          for (int blk = 0; blk < 4; ++blk) {
            uint32_t v[32];
            tmem_ld32(tcol + blk * 32, v);
            tmem_ld_wait_dep(v);
            const uint32_t kb = arow + (uint32_t)(blk >> 1) * 16384u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = blk * 32 + 8 * j;
              const float4 b0 = *reinterpret_cast<const float4*>(bl + c);
              const float4 b1 = *reinterpret_cast<const float4*>(bl + c + 4);
              float h[8];
              h[0] = fmaxf(__uint_as_float(v[8 * j + 0]) + b0.x, 0.f); h[1] = fmaxf(__uint_as_float(v[8 * j + 1]) + b0.y, 0.f);
              h[2] = fmaxf(__uint_as_float(v[8 * j + 2]) + b0.z, 0.f); h[3] = fmaxf(__uint_as_float(v[8 * j + 3]) + b0.w, 0.f);
              h[4] = fmaxf(__uint_as_float(v[8 * j + 4]) + b1.x, 0.f); h[5] = fmaxf(__uint_as_float(v[8 * j + 5]) + b1.y, 0.f);
              h[6] = fmaxf(__uint_as_float(v[8 * j + 6]) + b1.z, 0.f); h[7] = fmaxf(__uint_as_float(v[8 * j + 7]) + b1.w, 0.f);
              if (okind == OUT_HEAD) {
                const float* w = headw_s + hc * 128 + c;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  p0 = fmaf(h[e], w[e], p0);
                  p1 = fmaf(h[e], w[256 + e], p1);
                  p2 = fmaf(h[e], w[512 + e], p2);
                }
              } else {
                const uint32_t dst = kb + ((uint32_t)((((blk & 1) * 4) + j) << 4) ^ rx);
                st_shared_v4(dst, pack_op<kF16>(h[0], h[1]), pack_op<kF16>(h[2], h[3]), pack_op<kF16>(h[4], h[5]), pack_op<kF16>(h[6], h[7]));
              }
            }
          }
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kTrain || okind == OUT_ACT_IMG) mbar_arrive(bar(kBarOutDone + t));   // local: the mover may store the tile image
            mbar_arrive_cluster(leader_actready + 8u * t);
          }
          if (okind == OUT_HEAD) {
            // models.py:177-179: sigmoid(W_out . [h, latent] + b); the latent columns live in the effective bias
            if (hc == 1) *reinterpret_cast<float4*>(part_s + row * 4) = make_float4(p0, p1, p2, 0.f);
            named_bar_sync(1, kNumEpiThreads);
            if (hc == 0 && m < P.M) {
              const float4 o = *reinterpret_cast<const float4*>(part_s + row * 4);
              if constexpr (kPerRay) {
                const float* hbr = P.bias_rays + (m / P.S) * P.bias_ray_stride + nl * 256;
                hb[0] = hbr[0]; hb[1] = hbr[1]; hb[2] = hbr[2];
              }
              const float z0 = p0 + o.x + hb[0], z1 = p1 + o.y + hb[1], z2 = p2 + o.z + hb[2];
              float* dst = P.rgbsigma + m * 4;
              dst[0] = 1.0f / (1.0f + expf(-z0));
              dst[1] = 1.0f / (1.0f + expf(-z1));
              dst[2] = 1.0f / (1.0f + expf(-z2));
            }
            named_bar_sync(2, kNumEpiThreads);
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

// ---------------------------------------------------------------------------
// packing: one chunk = [256 rows x 64 K] bf16 SW128; value = W[n][col0 + k] for k < ncols else 0
// transposed chunks (the dgrad of style_bwd.cu): row n = input feature col0 + n, value = W[k0 + k][col0 + n] for k < ncols else 0
struct ChunkSrc { const float* W; int ld; int col0; int ncols; int k0; int trans; };

template <typename T> __device__ __forceinline__ T to_op(float v);
template <> __device__ __forceinline__ __nv_bfloat16 to_op<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half to_op<__half>(float v) { return __float2half_rn(v); }

template <typename T>
__global__ void pack_chunks_kernel(const ChunkSrc* __restrict__ table, int nchunks, T* __restrict__ out) {
  const size_t total = (size_t)nchunks * 256 * 64;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int chunk = (int)(idx / (256 * 64));
    const int byte = (int)(idx % (256 * 64)) * 2;
    const int grp = byte >> 10, rem = byte & 1023;
    const int rr = rem >> 7, inrow = rem & 127;
    const int c16 = (inrow >> 4) ^ rr, within = (inrow & 15) >> 1;
    const int n = grp * 8 + rr;
    const int k = c16 * 8 + within;
    const ChunkSrc s = table[chunk];
    float v = 0.f;
    if (k < s.ncols) v = s.trans ? s.W[(size_t)(s.k0 + k) * s.ld + s.col0 + n] : s.W[(size_t)n * s.ld + s.col0 + k];
    out[idx] = to_op<T>(v);
  }
}

// effective biases: out[j] = b[j] + sum_k Wlat[j][k] * latent[k]   (the 32 latent columns of a layer, copied at set time)
struct BiasSrc { const float* wlat; const float* b; int nout; const float* latent; float* out; };
__global__ void style_bias_kernel(const BiasSrc* __restrict__ table, int n) {
  const BiasSrc s = table[blockIdx.x];
  for (int j = threadIdx.x; j < s.nout; j += blockDim.x) {
    float acc = s.b[j];
    for (int k = 0; k < 32; ++k) acc = fmaf(s.wlat[j * 32 + k], s.latent[k], acc);
    s.out[j] = acc;
  }
}
// own copies of what the per-call bias kernel needs, so nothing of the caller's tensors is referenced after set_weights
// wlatT: per layer the same latent columns transposed [32][256] followed by their row sums [256] (the per-ray bias kernel of the
// training forward reads them coalesced over the output feature)
struct LatSrc { const float* W; const float* b; int ld; int lat0; int nout; float* wlat; float* bout; };
__global__ void style_latcopy_kernel(const LatSrc* __restrict__ table, float* __restrict__ wlatT) {
  const LatSrc s = table[blockIdx.x];
  float* T = wlatT + (size_t)blockIdx.x * 256 * 33;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
    const int j = i / 32, k = i % 32;
    const float v = j < s.nout ? s.W[(size_t)j * s.ld + s.lat0 + k] : 0.f;
    s.wlat[i] = v;
    T[k * 256 + j] = v;
  }
  for (int j = threadIdx.x; j < 256; j += blockDim.x) {
    s.bout[j] = j < s.nout ? s.b[j] : 0.f;
    float rs = 0.f;
    if (j < s.nout)
      for (int k = 0; k < 32; ++k) rs += s.W[(size_t)j * s.ld + s.lat0 + k];
    T[32 * 256 + j] = rs;
  }
}

// training: effective biases per RAY.  lat1 [n_rays][32] = the module-1 latent of each ray (latents_model_1(style_id, frame_id));
// module 2 sees mean(lat1) broadcast to 32 dims (train_tgtcs.py:376, :410).  out [n_rays][13][256]: module 1 layers 0..4,
// module 2 layers 0..6, head (3 used).  wlat: 13 slots of [256][32] latent columns followed by [256] biases.
__global__ void style_bias_rays_kernel(const float* __restrict__ wlat, const float* __restrict__ wlatT, const float* __restrict__ lat1,
                                       int64_t n_rays, float* __restrict__ out) {
  __shared__ float lat[33];
  const int64_t ray = blockIdx.x;
  if (threadIdx.x < 32) lat[threadIdx.x] = lat1[ray * 32 + threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < 32; ++k) s += lat[k];
    lat[32] = s * (1.0f / 32.0f);
  }
  __syncthreads();
  const int j = threadIdx.x;
  for (int l = 0; l < 13; ++l) {
    const float* T = wlatT + (size_t)l * 256 * 33;
    float acc = wlat[(size_t)l * 256 * 33 + 256 * 32 + j];   // bias (0 beyond the layer's outputs)
    if (l < 5) {
#pragma unroll 8
      for (int k = 0; k < 32; ++k) acc = fmaf(T[k * 256 + j], lat[k], acc);
    } else {
      acc = fmaf(T[32 * 256 + j], lat[32], acc);
    }
    out[(ray * 13 + l) * 256 + j] = acc;
  }
}

__global__ void head_copy_kernel(const float* __restrict__ W, int ld, float* __restrict__ out) {   // [3][256] <- W[3][ld][:256]
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) out[i] = W[(size_t)(i / 256) * ld + (i % 256)];
}

}  // namespace

// ---------------------------------------------------------------------------
// host side

static const int kCIn[5] = {95, 288, 288, 288, 351};
static const int kWIn[8] = {607, 288, 288, 288, 351, 288, 288, 288};
constexpr int kCChunks = 1 + 4 + 4 + 4 + 5;            // 18
constexpr int kWChunks = 9 + 4 + 4 + 4 + 5 + 4 + 4;    // 34
constexpr int kTChunks = 1 + 11 * 4;                    // 45: transposed chunks of the 12 dgrad GEMMs (style_bwd.cu)

int style_set_weights(tgtc_ctx* ctx, const float* const* params, cudaStream_t st) {
  StyleImage& im = ctx->style;
  if (im.blob_c == nullptr) {
    TGTC_CUDA(cudaMalloc(&im.blob_c, (size_t)kCChunks * 32768));
    TGTC_CUDA(cudaMalloc(&im.blob_w, (size_t)kWChunks * 32768));
    TGTC_CUDA(cudaMalloc(&im.blob_c_h, (size_t)kCChunks * 32768));
    TGTC_CUDA(cudaMalloc(&im.blob_w_h, (size_t)kWChunks * 32768));
    TGTC_CUDA(cudaMalloc(&im.head_w, 3 * 256 * sizeof(float)));
    TGTC_CUDA(cudaMalloc(&im.bias_c, 5 * 256 * sizeof(float)));
    TGTC_CUDA(cudaMalloc(&im.bias_w, 7 * 256 * sizeof(float)));
    TGTC_CUDA(cudaMalloc(&im.head_b, 4 * sizeof(float)));
    TGTC_CUDA(cudaMalloc(&im.latents, 64 * sizeof(float)));
    TGTC_CUDA(cudaMalloc(&im.blob_T, (size_t)kTChunks * 32768));
    TGTC_CUDA(cudaMalloc(&im.tables, 16384));
    TGTC_CUDA(cudaMalloc(&im.wlat, 13 * 256 * 33 * sizeof(float)));   // per layer: [256][32] latent columns, then [256] bias
    TGTC_CUDA(cudaMalloc(&im.wlatT, 13 * 256 * 33 * sizeof(float)));  // per layer: the same transposed [32][256], then row sums [256]
  }
  // The device tables hold the caller's parameter pointers; they are rebuilt only when those change (a trainer re-packs the
  // same 26 tensors after every optimizer step: then this call is five kernel launches, fully stream-ordered).
  bool same = im.set;
  for (int i = 0; i < 26 && same; ++i) same = (im.src[i] == params[i]);
  ChunkSrc* dtab = reinterpret_cast<ChunkSrc*>(im.tables);
  LatSrc* dlat = reinterpret_cast<LatSrc*>(im.tables + 192 * sizeof(ChunkSrc));
  BiasSrc* dbias = reinterpret_cast<BiasSrc*>(im.tables + 192 * sizeof(ChunkSrc) + 13 * sizeof(LatSrc));
  static_assert(sizeof(ChunkSrc) == 32 && 192 * sizeof(ChunkSrc) + 13 * (sizeof(BiasSrc) + sizeof(LatSrc)) <= 16384, "style tables do not fit");
  const float* const* C = params;        // concat module: (W,b) x 5
  const float* const* Wp = params + 10;  // wild module: (W,b) x 8
  if (!same) {
    // chunk tables (consumption order, see the header comment)
    std::vector<ChunkSrc> tc, tw;
    auto act4 = [](std::vector<ChunkSrc>& v, const float* W, int ld, int col0) { for (int k = 0; k < 4; ++k) v.push_back({W, ld, col0 + 64 * k, 64, 0, 0}); };
    auto actT = [](std::vector<ChunkSrc>& v, const float* W, int ld, int col0) { for (int k = 0; k < 4; ++k) v.push_back({W, ld, col0, 64, 64 * k, 1}); };
    tc.push_back({C[0], 95, 0, 63, 0, 0});
    for (int l = 1; l <= 3; ++l) act4(tc, C[2 * l], 288, 0);
    tc.push_back({C[8], 351, 288, 63, 0, 0});                 // skip layer: [h(256), latent(32), x(63)] (models.py:141-144)
    act4(tc, C[8], 351, 0);
    tw.push_back({Wp[0], 607, 512, 63, 0, 0});                // layer 0: [base_remap(256), concat_features(256), x(63), latent(32)]
    act4(tw, Wp[0], 607, 0);
    act4(tw, Wp[0], 607, 256);
    for (int l = 1; l <= 3; ++l) act4(tw, Wp[2 * l], 288, 0);
    tw.push_back({Wp[8], 351, 288, 63, 0, 0});
    act4(tw, Wp[8], 351, 0);
    act4(tw, Wp[10], 288, 0);
    act4(tw, Wp[12], 288, 0);
    // transposed chunks, in the order the style dgrad consumes them: head, W6..W1, W0 (concat_features columns), C4..C1
    std::vector<ChunkSrc> tt;
    tt.push_back({Wp[14], 288, 0, 3, 0, 1});
    for (int l = 6; l >= 1; --l) actT(tt, Wp[2 * l], kWIn[l], 0);
    actT(tt, Wp[0], 607, 256);
    for (int l = 4; l >= 1; --l) actT(tt, C[2 * l], kCIn[l], 0);
    TGTC_REQUIRE((int)tc.size() == kCChunks && (int)tw.size() == kWChunks && (int)tt.size() == kTChunks, TGTC_ERR_STATE,
                 "style chunk tables inconsistent");
    // latent columns + biases -> owned buffers; the per-call bias table (device)
    static const int clat[5] = {63, 256, 256, 256, 256};
    static const int wlat0[8] = {575, 256, 256, 256, 256, 256, 256, 256};
    std::vector<LatSrc> tl;
    std::vector<BiasSrc> tb;
    auto slot = [&](int i) { return im.wlat + (size_t)i * 256 * 33; };
    for (int l = 0; l < 5; ++l) {
      tl.push_back({C[2 * l], C[2 * l + 1], kCIn[l], clat[l], 256, slot(l), slot(l) + 256 * 32});
      tb.push_back({slot(l), slot(l) + 256 * 32, 256, im.latents, im.bias_c + l * 256});
    }
    for (int l = 0; l < 8; ++l) {
      const int nout = l < 7 ? 256 : 3;
      tl.push_back({Wp[2 * l], Wp[2 * l + 1], kWIn[l], wlat0[l], nout, slot(5 + l), slot(5 + l) + 256 * 32});
      tb.push_back({slot(5 + l), slot(5 + l) + 256 * 32, nout, im.latents + 32, l < 7 ? im.bias_w + l * 256 : im.head_b});
    }
    TGTC_CUDA(cudaMemcpyAsync(dtab, tc.data(), tc.size() * sizeof(ChunkSrc), cudaMemcpyHostToDevice, st));
    TGTC_CUDA(cudaMemcpyAsync(dtab + 64, tw.data(), tw.size() * sizeof(ChunkSrc), cudaMemcpyHostToDevice, st));
    TGTC_CUDA(cudaMemcpyAsync(dtab + 128, tt.data(), tt.size() * sizeof(ChunkSrc), cudaMemcpyHostToDevice, st));
    TGTC_CUDA(cudaMemcpyAsync(dlat, tl.data(), tl.size() * sizeof(LatSrc), cudaMemcpyHostToDevice, st));
    TGTC_CUDA(cudaMemcpyAsync(dbias, tb.data(), tb.size() * sizeof(BiasSrc), cudaMemcpyHostToDevice, st));
    TGTC_CUDA(cudaStreamSynchronize(st));   // the host vectors go out of scope
    for (int i = 0; i < 26; ++i) im.src[i] = params[i];
  }
  pack_chunks_kernel<__half><<<ctx->num_sms * 2, 256, 0, st>>>(dtab, kCChunks, reinterpret_cast<__half*>(im.blob_c_h));
  TGTC_LAUNCH_CHECK(ctx);
  pack_chunks_kernel<__half><<<ctx->num_sms * 2, 256, 0, st>>>(dtab + 64, kWChunks, reinterpret_cast<__half*>(im.blob_w_h));
  TGTC_LAUNCH_CHECK(ctx);
  pack_chunks_kernel<__nv_bfloat16><<<ctx->num_sms * 2, 256, 0, st>>>(dtab, kCChunks, reinterpret_cast<__nv_bfloat16*>(im.blob_c));
  TGTC_LAUNCH_CHECK(ctx);
  pack_chunks_kernel<__nv_bfloat16><<<ctx->num_sms * 2, 256, 0, st>>>(dtab + 64, kWChunks, reinterpret_cast<__nv_bfloat16*>(im.blob_w));
  TGTC_LAUNCH_CHECK(ctx);
  pack_chunks_kernel<__nv_bfloat16><<<ctx->num_sms * 2, 256, 0, st>>>(dtab + 128, kTChunks, reinterpret_cast<__nv_bfloat16*>(im.blob_T));
  TGTC_LAUNCH_CHECK(ctx);
  head_copy_kernel<<<1, 256, 0, st>>>(Wp[14], 288, im.head_w);
  TGTC_LAUNCH_CHECK(ctx);
  style_latcopy_kernel<<<13, 256, 0, st>>>(dlat, im.wlatT);
  TGTC_LAUNCH_CHECK(ctx);
  im.bias_table = dbias;
  im.set = true;
  return TGTC_OK;
}

// latent1 / latent2: device pointers to 32 floats (module 1 / module 2 latents of this call).  Fully stream-ordered.
int style_set_latents(tgtc_ctx* ctx, const float* latent1, const float* latent2, cudaStream_t st) {
  StyleImage& im = ctx->style;
  TGTC_CUDA(cudaMemcpyAsync(im.latents, latent1, 32 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  TGTC_CUDA(cudaMemcpyAsync(im.latents + 32, latent2, 32 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  style_bias_kernel<<<13, 256, 0, st>>>(reinterpret_cast<const BiasSrc*>(im.bias_table), 13);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

static void fill_common(ChainParams& P, const MlpIO& io) {
  P.rays_o = io.rays_o; P.rays_d = io.rays_d; P.ts = io.ts;
  P.t_scale = io.t_scale; P.t_near = io.t_near; P.S = io.S;
  P.M = io.n_rays * io.S;
  P.ntiles = (P.M + kTileM - 1) / kTileM;
  P.img0_stride = P.img1_stride = 65536;
}

static int g_chain_dbg = 0;
extern "C" void tgtc_debug_chain_flags(int f) { g_chain_dbg = f; }

static int launch_chain(tgtc_ctx* ctx, const ChainParams& P_in, cudaStream_t st, bool train = false, bool f16 = false, bool ray_bias = false) {
  ChainParams P = P_in;
  P.dbg_flags = g_chain_dbg;
  static bool attr_set[64] = {};
  if (!attr_set[ctx->device & 63]) {
    TGTC_CUDA(cudaFuncSetAttribute(mlp_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    TGTC_CUDA(cudaFuncSetAttribute(mlp_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    TGTC_CUDA((cudaFuncSetAttribute(mlp_chain_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
    TGTC_CUDA((cudaFuncSetAttribute(mlp_chain_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
    TGTC_CUDA((cudaFuncSetAttribute(mlp_chain_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
    attr_set[ctx->device & 63] = true;
  }
  const int64_t nquads = (P.ntiles + 3) / 4;
  const int64_t max_pairs = ctx->num_sms / 2;
  const int grid = 2 * (int)(nquads < max_pairs ? nquads : max_pairs);
  if (train) mlp_chain_kernel<true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else if (ray_bias && f16) mlp_chain_kernel<false, true, true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else if (ray_bias) mlp_chain_kernel<false, false, true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else if (f16) mlp_chain_kernel<false, true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else mlp_chain_kernel<false><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

// module 1: concat_features tile images for the samples of io -> cf_img [ntiles][64 KB]
int launch_style_concat(tgtc_ctx* ctx, const MlpIO& io, uint8_t* cf_img, cudaStream_t st, bool f16, const float* bias_rays) {
  if (io.n_rays == 0) return TGTC_OK;
  ChainParams P = {};
  fill_common(P, io);
  P.nlayers = 5;
  P.layer[0] = {{SEG_PE, 0, 0}, 1, OUT_ACT, 0};
  for (int l = 1; l <= 3; ++l) P.layer[l] = {{SEG_ACT, 0, 0}, 1, OUT_ACT, 0};
  P.layer[4] = {{SEG_PE, SEG_ACT, 0}, 2, OUT_ACT_IMG, 1};
  P.blob = f16 ? ctx->style.blob_c_h : ctx->style.blob_c;
  P.bias = ctx->style.bias_c;
  P.img_out = cf_img;
  if (bias_rays != nullptr) { P.bias_rays = bias_rays; P.bias_ray_stride = 13 * 256; }
  return launch_chain(ctx, P, st, false, f16, bias_rays != nullptr);
}

// module 2: stylised rgb -> rgbsigma[.].xyz from base_remap images (NeRF trunk) and concat_features images (module 1)
int launch_style_wild(tgtc_ctx* ctx, const MlpIO& io, const uint8_t* remap_img, const uint8_t* cf_img, cudaStream_t st, bool f16,
                      const float* bias_rays) {
  if (io.n_rays == 0) return TGTC_OK;
  ChainParams P = {};
  fill_common(P, io);
  P.nlayers = 7;
  P.layer[0] = {{SEG_PE, SEG_IMG0, SEG_IMG1}, 3, OUT_ACT, 0};
  for (int l = 1; l <= 3; ++l) P.layer[l] = {{SEG_ACT, 0, 0}, 1, OUT_ACT, 0};
  P.layer[4] = {{SEG_PE, SEG_ACT, 0}, 2, OUT_ACT, 1};
  P.layer[5] = {{SEG_ACT, 0, 0}, 1, OUT_ACT, 0};
  P.layer[6] = {{SEG_ACT, 0, 0}, 1, OUT_HEAD, 0};
  P.blob = f16 ? ctx->style.blob_w_h : ctx->style.blob_w;
  P.bias = ctx->style.bias_w;
  P.head_w = ctx->style.head_w;
  P.head_b = ctx->style.head_b;
  P.img0 = remap_img;
  P.img1 = cf_img;
  P.rgbsigma = io.rgbsigma;
  if (bias_rays != nullptr) { P.bias_rays = bias_rays + 5 * 256; P.bias_ray_stride = 13 * 256; }
  return launch_chain(ctx, P, st, false, f16, bias_rays != nullptr);
}

// ---------------------------------------------------------------------------
// explicit-input stage entries (the reference's concat_style_forward / style_forward callables, rendering.py:129-140): features
// arrive as fp32 rows; two small kernels convert rows <-> the 128B-swizzled tile images the chain kernel stages
template <typename T>
__global__ void rows_to_images_kernel(const float* __restrict__ x, int64_t M, int C, int ld, int col0, int nblk, T* __restrict__ img) {
  // img: [ntiles][nblk][128 rows x 64] elements; element (row, k) of block b at byte (row>>3)*1024 + (row&7)*128 + (((k>>3)^(row&7))<<4) + (k&7)*2
  const int64_t ntiles = (M + 127) / 128;
  const int64_t total = ntiles * nblk * 128 * 8;          // one thread per 16-byte chunk
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(idx & 7);
    const int row = (int)((idx >> 3) & 127);
    const int b = (int)((idx >> 10) % nblk);
    const int64_t tile = idx / ((int64_t)nblk * 1024);
    const int64_t m = tile * 128 + row;
    T v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = b * 64 + ch * 8 + e;
      v[e] = to_op<T>((m < M && c < C) ? x[m * ld + col0 + c] : 0.f);
    }
    uint8_t* dst = reinterpret_cast<uint8_t*>(img) + ((size_t)tile * nblk + b) * 16384 + (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
  }
}
template <typename T> __device__ __forceinline__ float from_op(T v);
template <> __device__ __forceinline__ float from_op<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float from_op<__half>(__half v) { return __half2float(v); }
template <typename T>
__global__ void images_to_rows_kernel(const T* __restrict__ img, int64_t M, int nblk, float* __restrict__ out) {   // out [M][nblk*64]
  const int64_t ntiles = (M + 127) / 128;
  const int64_t total = ntiles * nblk * 128 * 8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(idx & 7);
    const int row = (int)((idx >> 3) & 127);
    const int b = (int)((idx >> 10) % nblk);
    const int64_t tile = idx / ((int64_t)nblk * 1024);
    const int64_t m = tile * 128 + row;
    if (m >= M) continue;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(img) + ((size_t)tile * nblk + b) * 16384 + (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4);
    T v[8];
    *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(src);
    float* o = out + m * (int64_t)(nblk * 64) + b * 64 + ch * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = from_op<T>(v[e]);
  }
}

static int rows_to_images(tgtc_ctx* ctx, const float* x, int64_t M, int C, int ld, int col0, int nblk, uint8_t* img, bool f16, cudaStream_t st) {
  const int64_t total = ((M + 127) / 128) * nblk * 1024;
  const int64_t grid = (total + 255) / 256;
  const unsigned g = (unsigned)(grid < (int64_t)ctx->num_sms * 16 ? grid : (int64_t)ctx->num_sms * 16);
  if (f16) rows_to_images_kernel<__half><<<g, 256, 0, st>>>(x, M, C, ld, col0, nblk, reinterpret_cast<__half*>(img));
  else rows_to_images_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(x, M, C, ld, col0, nblk, reinterpret_cast<__nv_bfloat16*>(img));
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

// workspace: pe images [ntiles][16 KB] | img0 [ntiles][64 KB] | img1 [ntiles][64 KB] | rgbsigma [M][4] fp32
size_t style_stage_workspace_bytes(int64_t M) {
  const size_t tiles = (size_t)((M + 127) / 128);
  return tiles * (16384 + 2 * 65536) + (size_t)M * 16 + 1024;
}

// StyleMLP_before_concat.forward(x = embedded pts [M,63], latent) -> concat_features [M,256]   (models.py:137-147)
int launch_style_concat_explicit(tgtc_ctx* ctx, const float* x, int64_t M, float* cf_out, uint8_t* ws, bool f16, cudaStream_t st) {
  if (M == 0) return TGTC_OK;
  const size_t tiles = (size_t)((M + 127) / 128);
  uint8_t* pe = ws;
  uint8_t* cf = ws + tiles * 16384;
  int rc = rows_to_images(ctx, x, M, 63, 63, 0, 1, pe, f16, st);
  if (rc) return rc;
  ChainParams P = {};
  P.S = 128; P.M = M; P.ntiles = (int64_t)tiles; P.img0_stride = P.img1_stride = 65536;
  P.pe_img = pe;
  P.nlayers = 5;
  P.layer[0] = {{SEG_PE, 0, 0}, 1, OUT_ACT, 0};
  for (int l = 1; l <= 3; ++l) P.layer[l] = {{SEG_ACT, 0, 0}, 1, OUT_ACT, 0};
  P.layer[4] = {{SEG_PE, SEG_ACT, 0}, 2, OUT_ACT_IMG, 1};
  P.blob = f16 ? ctx->style.blob_c_h : ctx->style.blob_c;
  P.bias = ctx->style.bias_c;
  P.img_out = cf;
  rc = launch_chain(ctx, P, st, false, f16);
  if (rc) return rc;
  const int64_t total = (int64_t)tiles * 4 * 1024;
  const unsigned g = (unsigned)((total + 255) / 256 < (int64_t)ctx->num_sms * 16 ? (total + 255) / 256 : (int64_t)ctx->num_sms * 16);
  if (f16) images_to_rows_kernel<__half><<<g, 256, 0, st>>>(reinterpret_cast<const __half*>(cf), M, 4, cf_out);
  else images_to_rows_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(cf), M, 4, cf_out);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

// StyleMLP_Wild_multilayers.forward(x = embedded pts [M,63], concated = [base_remap | concat_features] [M,512], latent) -> rgb   (models.py:165-180)
int launch_style_wild_explicit(tgtc_ctx* ctx, const float* x, const float* concated, int64_t M, float* rgb_out, uint8_t* ws, bool f16,
                               cudaStream_t st) {
  if (M == 0) return TGTC_OK;
  const size_t tiles = (size_t)((M + 127) / 128);
  uint8_t* pe = ws;
  uint8_t* i0 = ws + tiles * 16384;
  uint8_t* i1 = i0 + tiles * 65536;
  float* rs = reinterpret_cast<float*>(i1 + tiles * 65536);
  int rc = rows_to_images(ctx, x, M, 63, 63, 0, 1, pe, f16, st);
  if (rc) return rc;
  rc = rows_to_images(ctx, concated, M, 256, 512, 0, 4, i0, f16, st);
  if (rc) return rc;
  rc = rows_to_images(ctx, concated, M, 256, 512, 256, 4, i1, f16, st);
  if (rc) return rc;
  ChainParams P = {};
  P.S = 128; P.M = M; P.ntiles = (int64_t)tiles; P.img0_stride = P.img1_stride = 65536;
  P.pe_img = pe;
  P.nlayers = 7;
  P.layer[0] = {{SEG_PE, SEG_IMG0, SEG_IMG1}, 3, OUT_ACT, 0};
  for (int l = 1; l <= 3; ++l) P.layer[l] = {{SEG_ACT, 0, 0}, 1, OUT_ACT, 0};
  P.layer[4] = {{SEG_PE, SEG_ACT, 0}, 2, OUT_ACT, 1};
  P.layer[5] = {{SEG_ACT, 0, 0}, 1, OUT_ACT, 0};
  P.layer[6] = {{SEG_ACT, 0, 0}, 1, OUT_HEAD, 0};
  P.blob = f16 ? ctx->style.blob_w_h : ctx->style.blob_w;
  P.bias = ctx->style.bias_w;
  P.head_w = ctx->style.head_w;
  P.head_b = ctx->style.head_b;
  P.img0 = i0;
  P.img1 = i1;
  P.rgbsigma = rs;
  rc = launch_chain(ctx, P, st, false, f16);
  if (rc) return rc;
  TGTC_CUDA(cudaMemcpy2DAsync(rgb_out, 12, rs, 16, 12, (size_t)M, cudaMemcpyDeviceToDevice, st));
  return TGTC_OK;
}

// ---------------------------------------------------------------------------
// training forward (Style_train): per-ray latents, stash for the backward
int launch_style_bias_rays(tgtc_ctx* ctx, const float* lat1, int64_t n_rays, float* bias_rays, cudaStream_t st) {
  if (n_rays == 0) return TGTC_OK;
  style_bias_rays_kernel<<<(unsigned)n_rays, 256, 0, st>>>(ctx->style.wlat, ctx->style.wlatT, lat1, n_rays, bias_rays);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

int launch_style_concat_train(tgtc_ctx* ctx, const MlpIO& io, const float* bias_rays, const StyleStash& stash, cudaStream_t st) {
  if (io.n_rays == 0) return TGTC_OK;
  ChainParams P = {};
  fill_common(P, io);
  P.nlayers = 5;
  P.layer[0] = {{SEG_PE, 0, 0}, 1, OUT_ACT, 0};
  for (int l = 1; l <= 3; ++l) P.layer[l] = {{SEG_ACT, 0, 0}, 1, OUT_ACT, 0};
  P.layer[4] = {{SEG_PE, SEG_ACT, 0}, 2, OUT_ACT_IMG, 1};
  P.blob = ctx->style.blob_c;
  P.bias_rays = bias_rays; P.bias_ray_stride = 13 * 256;
  P.stash = stash.c; P.stash_layers = 5;
  P.mask = stash.mask; P.mask_layers = kStyleMaskLayers; P.mask_slot0 = 7;
  P.stash_pe = stash.pe;
  return launch_chain(ctx, P, st, true);
}

int launch_style_wild_train(tgtc_ctx* ctx, const MlpIO& io, const float* bias_rays, const uint8_t* remap_img, const StyleStash& stash,
                            cudaStream_t st) {
  if (io.n_rays == 0) return TGTC_OK;
  ChainParams P = {};
  fill_common(P, io);
  P.nlayers = 7;
  P.layer[0] = {{SEG_PE, SEG_IMG0, SEG_IMG1}, 3, OUT_ACT, 0};
  for (int l = 1; l <= 3; ++l) P.layer[l] = {{SEG_ACT, 0, 0}, 1, OUT_ACT, 0};
  P.layer[4] = {{SEG_PE, SEG_ACT, 0}, 2, OUT_ACT, 1};
  P.layer[5] = {{SEG_ACT, 0, 0}, 1, OUT_ACT, 0};
  P.layer[6] = {{SEG_ACT, 0, 0}, 1, OUT_HEAD, 0};
  P.blob = ctx->style.blob_w;
  P.head_w = ctx->style.head_w;
  P.bias_rays = bias_rays + 5 * 256; P.bias_ray_stride = 13 * 256;
  P.img0 = remap_img;
  P.img1 = stash.c + 4 * 65536;            // concat_features = the last image of module 1's stash
  P.img1_stride = 5 * 65536;
  P.stash = stash.w; P.stash_layers = 7;
  P.mask = stash.mask; P.mask_layers = kStyleMaskLayers; P.mask_slot0 = 0;
  P.rgbsigma = io.rgbsigma;
  return launch_chain(ctx, P, st, true);
}
