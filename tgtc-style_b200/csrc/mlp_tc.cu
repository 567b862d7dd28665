// K3+K4, bf16 mode -- positional encoding + the 8x256 NeRF MLP as ONE persistent,
// warp-specialised tcgen05 kernel running on CTA PAIRS (cta_group::2).  Replaces
// models.Embedder.forward (models.py:46-60), models.MLP_style.forward
// (models.py:95-117) and models.StyleNerf.forward (models.py:216-223);
// utils.batchify (utils.py:435-456) disappears (the CTAs stream 128-sample tiles).
//
// A cluster of two CTAs (the two SMs of a TPC) works on four 128-sample tiles at a
// time: each CTA owns two tiles ("slots").  One tcgen05.mma.cta_group::2
// (M=256, N=256|128, K=16, bf16 x bf16 -> fp32) multiplies slot t of BOTH CTAs
// by the same weight slice: every CTA stages only HALF of the weight rows
// (N/2) in its shared memory, which halves the L2->SMEM weight stream and the
// operand-read pressure per SM -- the limiter of the single-CTA version.
//
// Per CTA (1 per SM, 448 threads):
//   warp 0      weight producer: cp.async.bulk (UBLKCP) of this CTA's half of each
//               pre-swizzled [N x 64] bf16 chunk (L2-resident blob) into a 3-stage ring
//   warp 1      leader CTA: MMA issuer -- the warp runs converged, one elected lane
//               issues tcgen05.mma / tcgen05.commit (multicast to both CTAs' barriers);
//               peer CTA: forwards "my half of the stage has landed" to the leader
//   warps 2-5   input producers: pts = o + t*d, positional encoding written as the
//               bf16 A operand of layer 0 / the skip of layer 5 (128B-swizzled K-major),
//               and the per-ray view-direction term of rgb0 (fp32, CUDA cores)
//   warps 6-13  epilogue: tcgen05.ld the accumulator, +bias, ReLU, bf16, store as
//               the next layer's A operand (in place); fp32 sigma head at layer 7,
//               fp32 rgb1 head + sigmoid at rgb0; float4 (r,g,b,sigma) to HBM
// While the tensor cores run layer l of slot 0, the epilogue warps turn slot 1's
// accumulator (TMEM columns [256,512)) into its next A operand and vice versa, so
// activations never leave the SM between layers.
//
// Tensor-roofline kernel: 1 186 816 algorithmic FLOP per sample; HBM traffic is
// 24 B/ray in + 16 B/sample out.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "sample_fine.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kStages = 3;
constexpr int kStageBytes = 128 * kTcChunkK * 2;  // 16384: this CTA's half (N/2 rows) of a [256 x 64] chunk
constexpr int kNumThreads = 448;
constexpr int kPeWarp0 = 2, kEpiWarp0 = 6;
constexpr int kNumEpiThreads = 256;
constexpr int kNumPeThreads = 128;
constexpr int kMaxRaysPerTile = 2;  // S >= 64

// ---- shared memory map (bytes from a 1024-aligned base)
constexpr int kOffAct = 0;                                   // 2 x [4 kblocks][128 rows x 128 B]  SW128
constexpr int kActBytes = kTileM * 256 * 2;                  // 65536
constexpr int kOffPe = kOffAct + 2 * kActBytes;              // 2 x [128 rows x 128 B]             SW128
constexpr int kPeBytes = kTileM * 64 * 2;                    // 16384
constexpr int kOffW = kOffPe + 2 * kPeBytes;                 // 3 x 16384                          SW128
constexpr int kOffBias = kOffW + kStages * kStageBytes;      // 9 x 256 fp32
constexpr int kOffWSig = kOffBias + 9 * 256 * 4;             // 256 fp32
constexpr int kOffWRgb1 = kOffWSig + 256 * 4;                // 3 x 128 fp32
constexpr int kOffDirBias = kOffWRgb1 + 384 * 4;             // [2 slots][2 bufs][2 rays][128] fp32
constexpr int kOffSigPart = kOffDirBias + 2 * 2 * kMaxRaysPerTile * 128 * 4;  // [2 slots][128] fp32
constexpr int kOffRgbPart = kOffSigPart + 2 * 128 * 4;       // [128][4] fp32
constexpr int kCompStageOff = kActBytes - kTileM * 32;       // fused K5 staging rows: the last 4 KB of an activation buffer (K block 3, rows 96..127)
constexpr int kOffCompTot = kOffRgbPart + 128 * 4 * 4;       // [4] fp32: per-warp transmittance products of the fused compositing
constexpr int kOffTsRow = kOffCompTot + 4 * 4 + 16;          // [64] fp32: the coarse ts row of the fused resampling (K6+K7)
constexpr int kOffBars = kOffTsRow + 64 * 4;                 // mbarriers
constexpr int kNumBars = 2 * kStages + 14;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");

// barrier indices
//   WFull[s]    this CTA's half of ring stage s landed (leader: + the peer's half, forwarded)      leader count 2, peer 1
//   WEmpty[s]   the MMAs that read stage s retired (commit, multicast to both CTAs)                  count 1
//   PeReady[t]  slot t's PE tile + dir term written by this CTA's producers (local: epilogue waits)  count 128
//   PePair[t]   leader only: both CTAs' PE tiles of slot t written                                   count 8 (warps)
//   PeFree[t]   layer 5 of slot t retired: PE tile reusable (commit, multicast)                      count 1
//   ActReady[t] leader only: both CTAs' epilogues stored slot t's next A operand / drained TMEM      count 16 (warps)
//   AccFull[t]  slot t's accumulator complete (commit, multicast)                                    count 1
//   CompReady[t] fused K5: the tile's (r,g,b,sigma) rows are staged in act[t] (local: epilogue -> producers) count 4 (warps)
//   CompDone[t]  fused K5: the producers have composited the tile, act[t] may be overwritten (local)       count 4 (warps)
constexpr int kBarWFull = 0, kBarWEmpty = kStages, kBarPeReady = 2 * kStages, kBarPeFree = kBarPeReady + 2,
              kBarActReady = kBarPeFree + 2, kBarAccFull = kBarActReady + 2, kBarPePair = kBarAccFull + 2,
              kBarCompReady = kBarPePair + 2, kBarCompDone = kBarCompReady + 2;

using namespace tcptx;

// one 32-column block of a hidden-layer epilogue: +bias, ReLU, bf16, 4 x 16-byte swizzled stores.
// blk = 32-column block index inside this thread's 128 columns; kb = row base of the 64-column K block.
// mrow: when training, the block's ReLU mask word (bit 31-c = sign of pre-activation c, i.e. 1 = gradient blocked) goes to the
// mask stash in HBM that mlp_dgrad_kernel reads instead of the 16x larger activation image; nullptr otherwise.
// nb0 / nb1 (inference form): the bias values of this block's FIRST 8-column group, loaded by the caller; on return they hold the
// next block's first group.  Every group's bias (and sigma-head weight) loads are issued BEFORE the previous group's activation
// store: ptxas cannot tell the bias table from the activation buffer, so a load written after the store is issued after it and
// the packed add that needs it waits out the full shared-memory latency -- 16 exposed round trips per layer and slot, 40 % of
// the epilogue warps' time in the round-2 ncu source view (stall_short_sb on the FADD2s).
template <bool kSigma, bool kMask, bool kF16>
__device__ __forceinline__ void epi_block(uint32_t (&v)[32], const float* bl, const float* wsig, uint32_t kb, uint32_t rx, int blk,
                                          float& sig, uint32_t* mrow, float4& nb0, float4& nb1) {
  if constexpr (kMask) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = blk * 32 + 8 * j;  // column inside the 128-column half
      const float4 b0 = *reinterpret_cast<const float4*>(bl + c);
      const float4 b1 = *reinterpret_cast<const float4*>(bl + c + 4);
      add2(v[8 * j + 0], v[8 * j + 1], b0.x, b0.y);
      add2(v[8 * j + 2], v[8 * j + 3], b0.z, b0.w);
      add2(v[8 * j + 4], v[8 * j + 5], b1.x, b1.y);
      add2(v[8 * j + 6], v[8 * j + 7], b1.z, b1.w);
    }
    if (mrow != nullptr) mrow[blk * 128] = sign_mask32(v);
  }
  float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
  if constexpr (kSigma) {
    w0 = *reinterpret_cast<const float4*>(wsig + blk * 32);
    w1 = *reinterpret_cast<const float4*>(wsig + blk * 32 + 4);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = blk * 32 + 8 * j;
    const uint32_t coff = (uint32_t)((((blk & 1) * 4) + j) << 4) ^ rx;
    const uint32_t dst = kb + coff;
    if constexpr (!kMask) {
      const float4 b0 = nb0, b1 = nb1;
      if (c + 8 < 128) {   // the next group's bias: in flight while this group is converted and stored
        nb0 = *reinterpret_cast<const float4*>(bl + c + 8);
        nb1 = *reinterpret_cast<const float4*>(bl + c + 12);
      }
      if constexpr (kSigma) {
        const float ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        if (j < 3) {
          w0 = *reinterpret_cast<const float4*>(wsig + c + 8);
          w1 = *reinterpret_cast<const float4*>(wsig + c + 12);
        }
        float h[8];
        h[0] = fmaxf(__uint_as_float(v[8 * j + 0]) + b0.x, 0.f);
        h[1] = fmaxf(__uint_as_float(v[8 * j + 1]) + b0.y, 0.f);
        h[2] = fmaxf(__uint_as_float(v[8 * j + 2]) + b0.z, 0.f);
        h[3] = fmaxf(__uint_as_float(v[8 * j + 3]) + b0.w, 0.f);
        h[4] = fmaxf(__uint_as_float(v[8 * j + 4]) + b1.x, 0.f);
        h[5] = fmaxf(__uint_as_float(v[8 * j + 5]) + b1.y, 0.f);
        h[6] = fmaxf(__uint_as_float(v[8 * j + 6]) + b1.z, 0.f);
        h[7] = fmaxf(__uint_as_float(v[8 * j + 7]) + b1.w, 0.f);
        // fp32 sigma head on the un-rounded activations (models.py:103)
#pragma unroll
        for (int e = 0; e < 8; ++e) sig = fmaf(h[e], ws[e], sig);
        const uint32_t q0 = pack_op<kF16>(h[0], h[1]), q1 = pack_op<kF16>(h[2], h[3]), q2 = pack_op<kF16>(h[4], h[5]), q3 = pack_op<kF16>(h[6], h[7]);
        st_shared_v4(dst, q0, q1, q2, q3);
      } else {
        add2(v[8 * j + 0], v[8 * j + 1], b0.x, b0.y);
        add2(v[8 * j + 2], v[8 * j + 3], b0.z, b0.w);
        add2(v[8 * j + 4], v[8 * j + 5], b1.x, b1.y);
        add2(v[8 * j + 6], v[8 * j + 7], b1.z, b1.w);
        const uint32_t q0 = pack_op_relu<kF16>(v[8 * j + 0], v[8 * j + 1]), q1 = pack_op_relu<kF16>(v[8 * j + 2], v[8 * j + 3]),
                       q2 = pack_op_relu<kF16>(v[8 * j + 4], v[8 * j + 5]), q3 = pack_op_relu<kF16>(v[8 * j + 6], v[8 * j + 7]);
        st_shared_v4(dst, q0, q1, q2, q3);
      }
    } else {
      // bias already added above (the mask wants the pre-activations)
      if constexpr (kSigma) {
        // fp32 sigma head on the un-rounded activations (models.py:103)
        const float4 s0 = *reinterpret_cast<const float4*>(wsig + c);
        const float4 s1 = *reinterpret_cast<const float4*>(wsig + c + 4);
        const float ws[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) sig = fmaf(fmaxf(__uint_as_float(v[8 * j + e]), 0.f), ws[e], sig);
      }
      const uint32_t q0 = pack_op_relu<kF16>(v[8 * j + 0], v[8 * j + 1]), q1 = pack_op_relu<kF16>(v[8 * j + 2], v[8 * j + 3]),
                     q2 = pack_op_relu<kF16>(v[8 * j + 4], v[8 * j + 5]), q3 = pack_op_relu<kF16>(v[8 * j + 6], v[8 * j + 7]);
      st_shared_v4(dst, q0, q1, q2, q3);
    }
  }
}

// this thread's 128 columns of one hidden layer: TMEM loads software-pipelined against the math
// (the load of block b+1 is in flight while block b is converted and stored)
// wait_bar != 0: an mbarrier (address, parity) to pass before this thread's LAST two blocks are stored -- the fused-compositing
// staging rows live in the tail of the activation buffer (K block 3, rows 96..127), which only those stores overwrite.
template <bool kSigma, bool kMask, bool kF16>
__device__ __forceinline__ float hidden_epilogue(uint32_t tcol, const float* bl, const float* wsig, uint32_t arow, uint32_t rx,
                                                 uint32_t* mrow, uint32_t wait_bar = 0u, uint32_t wait_par = 0u) {
  float sig = 0.f;
  uint32_t va[32], vb[32];
  float4 nb0 = make_float4(0.f, 0.f, 0.f, 0.f), nb1 = nb0;
  if constexpr (!kMask) {
    nb0 = *reinterpret_cast<const float4*>(bl);
    nb1 = *reinterpret_cast<const float4*>(bl + 4);
  }
  tmem_ld32(tcol, va);
  tmem_ld_wait_dep(va);
  // (four unrolled blocks: a two-pass loop over block pairs halves the code but measured 2.3 % slower, A/B on one box)
  tmem_ld32(tcol + 32, vb);
  epi_block<kSigma, kMask, kF16>(va, bl, wsig, arow, rx, 0, sig, mrow, nb0, nb1);
  tmem_ld_wait_dep(vb);
  tmem_ld32(tcol + 64, va);
  epi_block<kSigma, kMask, kF16>(vb, bl, wsig, arow, rx, 1, sig, mrow, nb0, nb1);
  tmem_ld_wait_dep(va);
  tmem_ld32(tcol + 96, vb);
  if (wait_bar != 0u) mbar_wait(wait_bar, wait_par);
  epi_block<kSigma, kMask, kF16>(va, bl, wsig, arow + 16384, rx, 2, sig, mrow, nb0, nb1);
  tmem_ld_wait_dep(vb);
  epi_block<kSigma, kMask, kF16>(vb, bl, wsig, arow + 16384, rx, 3, sig, mrow, nb0, nb1);
  return sig;
}

struct TcParams {
  const uint8_t* blob;
  const float* smalls;
  MlpIO io;
  int64_t M;        // samples
  int64_t ntiles;
  int rays_per_tile;  // 128/S when S<128 else 1
  // training: activation stash (tile images, see mlp_bwd.cu); all nullptr for inference
  uint8_t* stash_h;   // [ntiles][9][128 x 256 bf16]  post-ReLU outputs of L0..L7 and remap
  uint8_t* stash_f;   // [ntiles][128 x 128 bf16]      post-ReLU output of rgb0
  uint8_t* stash_pe;  // [ntiles][128 x 64 bf16]       positional encoding tile
  uint32_t* stash_mask;  // [ntiles][10][8][128] ReLU mask words (sign bits of the pre-activations), see common.cuh: TcStash
  int trunk;          // style path: run L0..L7 + sigma + remap only; stash ONLY the remap tile ([ntiles][64 KB]) and write sigma
  int dbg_flags;      // timing experiments (results garbage): 2 = skip the hidden-layer epilogue work, 16 = no weight ring at all,
                      // training: 32 = no activation-image store, 64 = no store-drain wait/barrier, 128 = no mask words
  int dbg_layers;     // >0: stop after this many GEMM layers and dump the fp32 accumulator (tests)
  float* dbg_out;     // [ntiles*128, 256]
  long long* dbg_trace;  // timing experiments: clock64 stamps of CTA 0's roles, [4 roles][4 iters][10 layers][2 slots][2]
};
#define TC_TRACE(role, it, l, t, k)                                                                              \
  do {                                                                                                           \
    if constexpr (kDbg)                                                                                          \
    if (P.dbg_trace != nullptr && blockIdx.x == 0 && (it) < 4)                                                   \
      P.dbg_trace[(((((role)*4 + (int)(it)) * 10 + (l)) * 2 + (t)) * 2) + (k)] = clock64();                      \
  } while (0)

// tile of (cluster iteration it, slot t) for this CTA: a cluster works on 4 consecutive tiles per iteration
// (CTA rank r owns tiles 4q+2r and 4q+2r+1); tiles >= ntiles are padding (computed on clamped samples, never stored)
__device__ __forceinline__ int64_t my_tile(int64_t it, int t, uint32_t rank) {
  const int64_t quad = (int64_t)(blockIdx.x >> 1) + it * (int64_t)(gridDim.x >> 1);
  return quad * 4 + 2 * (int64_t)rank + t;
}

// ---------------------------------------------------------------------------
// kTrain: also write the activation stash (training forward); the inference instantiation carries none of that code.
// kTrunk (style path; implies P.trunk): the inference epilogue on L0..L7 + remap, and only the remap tile leaves as an image --
// two barriers per tile instead of the training instantiation's one per layer.
// kF16: fp16 instead of bf16 operands (weights image, PE tile, activations); inference only.
// kDbg: the test / timing hooks (layer dump, role clock trace, work-skipping flags) -- a separate instantiation, so that the
// production kernels carry none of that code (their instruction footprint is performance-relevant).
template <bool kTrain, bool kTrunk = false, bool kF16 = false, bool kDbg = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1) mlp_tc_kernel(const TcParams P) {
  static_assert(!(kTrain && kF16), "the training stash / backward kernels are bf16");
  const int dbg_flags = kDbg ? P.dbg_flags : 0;
  const int dbg_layers = kDbg ? P.dbg_layers : 0;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bars = sbase + kOffBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  const uint32_t rank = cluster_ctarank();   // 0 = leader (issues the MMAs of the pair)

  const int64_t nquads = (P.ntiles + 3) / 4;
  const int64_t ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const int64_t iters = nquads > cid ? (nquads - cid + ncl - 1) / ncl : 0;   // identical in both CTAs of the pair
  const int nlayers = dbg_layers > 0 ? dbg_layers : (P.trunk ? 9 : kTcNumGemm);

  // ---- one-time setup
  if (threadIdx.x == 0) {
    if ((sbase & 1023u) != 0) { printf("tgtc mlp_tc: shared memory base not 1024-aligned\n"); __trap(); }
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(kBarWFull + s), rank == 0 ? 2 : 1); mbar_init(bar(kBarWEmpty + s), 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar(kBarPeReady + t), kNumPeThreads);
      mbar_init(bar(kBarPePair + t), 2 * (kNumPeThreads / 32));
      mbar_init(bar(kBarPeFree + t), 1);
      mbar_init(bar(kBarActReady + t), 2 * (kNumEpiThreads / 32));
      mbar_init(bar(kBarAccFull + t), 1);
      // fused K5: one arrival per epilogue warp that stages partial sums (tail_off: all eight; otherwise the four hc == 0 warps)
      mbar_init(bar(kBarCompReady + t), (!kTrain && !kTrunk && P.io.comp_rgb != nullptr && P.io.rgbsigma == nullptr && P.io.rgb == nullptr)
                                            ? kNumEpiThreads / 32 : kNumEpiThreads / 64);
      mbar_init(bar(kBarCompDone + t), kNumPeThreads / 32);     // one arrival per producer warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {  // TMEM: all 512 columns in both CTAs of the pair (two 128x256 fp32 accumulators per CTA)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(sbase + kOffTmemPtr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
  }
  // biases / head weights -> shared memory (fp32), once per CTA
  {
    float* dst = reinterpret_cast<float*>(smem + kOffBias);
    for (int i = threadIdx.x; i < 9 * 256; i += kNumThreads) dst[i] = P.smalls[kSmBias + i];
    float* ws = reinterpret_cast<float*>(smem + kOffWSig);
    for (int i = threadIdx.x; i < 256; i += kNumThreads) ws[i] = P.smalls[kSmWSigma + i];
    float* wr = reinterpret_cast<float*>(smem + kOffWRgb1);
    for (int i = threadIdx.x; i < 384; i += kNumThreads) wr[i] = P.smalls[kSmWRgb1 + i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  // the pair owns all 512 columns of both SMs' tensor memory, so the allocation starts at lane 0 / column 0; using the
  // literal keeps every TMEM address a compile-time/uniform value in the issue loops
  if (*reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr) != 0u) {
    if (threadIdx.x == 0) printf("tgtc mlp_tc: unexpected TMEM base\n");
    __trap();
  }
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0) {
    // =====================================================================
    // weight producer (whole warp converged, one elected lane issues the bulk copies): this CTA's half of the rows
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t it = 0; it < ((dbg_flags & 16) ? 0 : iters); ++it) {
      for (int l = 0; l < nlayers; ++l) {
        const uint32_t hbytes = (uint32_t)tc_layer_n(l) * kTcChunkK;   // half of a [N x 64] bf16 chunk
        const uint8_t* src = P.blob + tc_layer_off_bytes(l) + (size_t)rank * hbytes;
        const int nch = tc_layer_chunks(l);
        for (int t = 0; t < 2; ++t) {
          for (int c = 0; c < nch; ++c) {
            mbar_wait(bar(kBarWEmpty + stage), phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(bar(kBarWFull + stage), hbytes);
              bulk_g2s(sbase + kOffW + stage * kStageBytes, src + (size_t)c * 2 * hbytes, hbytes, bar(kBarWFull + stage));
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && rank != 0) {
    // =====================================================================
    // peer CTA: tell the leader's MMA warp when this CTA's half of each ring stage has landed
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t leader_wfull = mapa_cluster(bar(kBarWFull), 0);
    for (int64_t it = 0; it < ((dbg_flags & 16) ? 0 : iters); ++it) {
      for (int l = 0; l < nlayers; ++l) {
        const int nch = 2 * tc_layer_chunks(l);
        for (int c = 0; c < nch; ++c) {
          mbar_wait(bar(kBarWFull + stage), phase);
          if (elect_one()) mbar_arrive_cluster(leader_wfull + 8u * stage);
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =====================================================================
    // leader CTA: MMA issuer for the pair.  The warp runs converged with warp-uniform control flow (vote-based waits), so
    // descriptors and barrier addresses live in uniform registers; one elected lane issues tcgen05.mma / commit.
    // Descriptor low words advance by (bytes >> 4); everything per K step is an integer add.
    int stage = 0;
    uint32_t phase = 0;
    uint32_t act_par0 = 0, act_par1 = 0, pe_par0 = 0, pe_par1 = 0;
    const uint32_t w_lo0 = (((sbase + kOffW) & 0x3FFFFu) >> 4) | (1u << 16);
    const bool ring = !(dbg_flags & 16);
    // one K=64 chunk: wait for the ring stage, four MMAs (K=16 each: +32 B inside the 128 B swizzled rows), release the stage
    auto issue_chunk = [&](uint32_t d_tmem, uint32_t a_lo, uint32_t idesc, uint32_t accumulate) {
      if (ring) { mbar_wait_uniform(bar(kBarWFull + stage), phase); tc_fence_after(); }
      const uint32_t b_lo = w_lo0 + (uint32_t)stage * (kStageBytes >> 4);
      if (elect_one()) {
        umma_bf16_lohi(d_tmem, a_lo, kDescHiSW128, b_lo, kDescHiSW128, idesc, accumulate);
        umma_bf16_lohi(d_tmem, a_lo + 2u, kDescHiSW128, b_lo + 2u, kDescHiSW128, idesc, 1u);
        umma_bf16_lohi(d_tmem, a_lo + 4u, kDescHiSW128, b_lo + 4u, kDescHiSW128, idesc, 1u);
        umma_bf16_lohi(d_tmem, a_lo + 6u, kDescHiSW128, b_lo + 6u, kDescHiSW128, idesc, 1u);
        if (ring) umma_commit(bar(kBarWEmpty + stage));  // frees the ring stage in both CTAs when these MMAs retire
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    };
    for (int64_t it = 0; it < iters; ++it) {
      for (int l = 0; l < nlayers; ++l) {
        const uint32_t idesc = make_idesc_op<kF16>(2 * kTileM, tc_layer_n(l));
        const bool has_pe = (l == 0 || l == 5);
        const bool has_act = (l != 0);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          // A operands of both CTAs ready?  (also: this slot's accumulators drained by both epilogues)
          uint32_t& act_par = t ? act_par1 : act_par0;
          uint32_t& pe_par = t ? pe_par1 : pe_par0;
          if (l == 0) {
            mbar_wait_uniform(bar(kBarPePair + t), pe_par); pe_par ^= 1;
            if (it > 0) { mbar_wait_uniform(bar(kBarActReady + t), act_par); act_par ^= 1; }
          } else {
            mbar_wait_uniform(bar(kBarActReady + t), act_par); act_par ^= 1;
          }
          tc_fence_after();
          TC_TRACE(0, it, l, t, 0);
          const uint32_t d_tmem = tmem_base + (uint32_t)(256 * t);
          const uint32_t pe_lo = (((sbase + kOffPe + t * kPeBytes) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t act_lo = (((sbase + kOffAct + t * kActBytes) & 0x3FFFFu) >> 4) | (1u << 16);
          // chunk = one 64-column K block: the PE tile (layers 0 and 5) or K block c of the activations (16 KB apart)
          if (has_pe) issue_chunk(d_tmem, pe_lo, idesc, 0u);
          if (has_act) {
#pragma unroll
            for (int c = 0; c < 4; ++c) issue_chunk(d_tmem, act_lo + 1024u * (uint32_t)c, idesc, (has_pe || c > 0) ? 1u : 0u);
          }
          if (elect_one()) {
            if (l == 5 || (l == nlayers - 1 && nlayers <= 5)) umma_commit(bar(kBarPeFree + t));
            umma_commit(bar(kBarAccFull + t));
          }
          __syncwarp();
          TC_TRACE(0, it, l, t, 1);
        }
      }
    }
  } else if (warp < kEpiWarp0) {
    // =====================================================================
    // input producers: one thread per tile row
    const int r = (warp - kPeWarp0) * 32 + lane;
    const MlpIO& io = P.io;
    const int S = io.S;
    const uint32_t leader_pepair = mapa_cluster(bar(kBarPePair), 0);
    const bool comp = !kTrain && !kTrunk && io.comp_rgb != nullptr;
    // tail_off: nothing per-sample is asked for, so the epilogue warps hand the rgb1 head's partial sums over and these warps finish
    // the sample (sum, biases, sigmoid) before compositing it
    const bool tail_off = comp && io.rgbsigma == nullptr && io.rgb == nullptr;
    const float comp_b_sigma = P.smalls[kSmBSigma];
    const float comp_b_rgb[3] = {P.smalls[kSmBRgb1 + 0], P.smalls[kSmBRgb1 + 1], P.smalls[kSmBRgb1 + 2]};
    // ---- fused K5: utils.alpha_composition (utils.py:354-386) for the rays of tile (it_, t_) (S in {64,128}: rows [0,S) are one
    // ray, a warp owns 32 consecutive samples), run by these producer warps while they would otherwise wait for the tile's PE
    // buffer to be released.  Every operation and its order is composite.cu's, so the results are bit-identical to the
    // stand-alone kernel: per-warp inclusive product scan of (1-alpha+1e-10), sequential carry over the ray's 32-sample
    // chunks, per-lane FMA sums over the chunks, xor-butterfly.
    auto composite_tile = [&](int64_t it_, int t_) {
      const int64_t m = my_tile(it_, t_, rank) * kTileM + r;
      const bool valid = m < P.M;
      const int k = r & (S - 1);
      float tcur = 0.f, tnext = 0.f;              // sample positions: fetched before the wait (L2 latency off the hand-over)
      if (io.ts != nullptr) {
        if (valid) { tcur = __ldcg(io.ts + m); if (k + 1 < S) tnext = __ldcg(io.ts + m + 1); }
      } else {
        tcur = coarse_t(k, S, io.t_scale, io.t_near);
        if (k + 1 < S) tnext = coarse_t(k + 1, S, io.t_scale, io.t_near);
      }
      mbar_wait_relaxed(bar(kBarCompReady + t_), (uint32_t)(it_ & 1), 100);   // may be microseconds away: do not steal issue slots
      uint8_t* const stg = smem + kOffAct + t_ * kActBytes + kCompStageOff;      // [128 rows][32 B]: (r,g,b,sigma) from the epilogue
      float4 v = *reinterpret_cast<const float4*>(stg + r * 32);
      if (tail_off) {
        // the epilogue warps staged the two column halves' partial sums of the rgb1 head (and the first half's share of sigma)
        // instead of the finished sample: the sum, biases and sigmoid (models.py:103, :111) run here, off the epilogue's
        // critical loop -- the same expressions in the same order as the epilogue's own tail, so the results are bit-identical
        const float4 o = *reinterpret_cast<const float4*>(stg + r * 32 + 16);
        const float z0 = v.x + o.x + comp_b_rgb[0], z1 = v.y + o.y + comp_b_rgb[1], z2 = v.z + o.z + comp_b_rgb[2];
        const float sg = v.w + reinterpret_cast<const float*>(smem + kOffSigPart)[t_ * 128 + r] + comp_b_sigma;
        v = make_float4(1.0f / (1.0f + expf(-z0)), 1.0f / (1.0f + expf(-z1)), 1.0f / (1.0f + expf(-z2)), sg);
      }
      const int pw = r >> 5;                      // producer warp = 32-row quarter of the tile
      const int c = k >> 5;                       // 32-sample chunk of this warp inside its ray
      const float delta = (k + 1 < S) ? __fsub_rn(tnext, tcur) : 1e10f;
      const float act = fmaxf(v.w, 0.0f);
      const float alpha = valid ? __fsub_rn(1.0f, expf(-__fmul_rn(act, delta))) : 0.0f;
      const float fac = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
      float incl = fac;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl = __fmul_rn(incl, up);
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      float* tot_s = reinterpret_cast<float*>(smem + kOffCompTot);
      if (lane == 31) tot_s[pw] = incl;
      named_bar_sync(4, kNumPeThreads);
      float carry = 1.0f;
      for (int j = 0; j < c; ++j) carry = __fmul_rn(carry, tot_s[pw - c + j]);
      const float T = __fmul_rn(carry, excl);
      const float w = __fmul_rn(alpha, T);
      if (valid && io.comp_weights != nullptr) io.comp_weights[m] = w;
      *reinterpret_cast<float4*>(stg + r * 32) = make_float4(v.x, v.y, v.z, w);
      *reinterpret_cast<float*>(stg + r * 32 + 16) = tcur;
      named_bar_sync(4, kNumPeThreads);
      if (c == 0) {
        float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
        for (int j = 0; j < (S >> 5); ++j) {
          const float4 u = *reinterpret_cast<const float4*>(stg + (r + 32 * j) * 32);
          const float tt = *reinterpret_cast<const float*>(stg + (r + 32 * j) * 32 + 16);
          ar = __fmaf_rn(u.w, u.x, ar); ag = __fmaf_rn(u.w, u.y, ag); ab = __fmaf_rn(u.w, u.z, ab);
          ad = __fmaf_rn(u.w, tt, ad); aa = __fadd_rn(aa, u.w);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          ar += __shfl_xor_sync(0xffffffffu, ar, d);
          ag += __shfl_xor_sync(0xffffffffu, ag, d);
          ab += __shfl_xor_sync(0xffffffffu, ab, d);
          ad += __shfl_xor_sync(0xffffffffu, ad, d);
          aa += __shfl_xor_sync(0xffffffffu, aa, d);
        }
        const int64_t ray = m / S;                 // lane 0: the ray's first sample
        if (lane == 0 && ray < io.n_rays) {
          if (io.comp_white_bkgd) { const float bg = __fsub_rn(1.0f, aa); ar = __fadd_rn(ar, bg); ag = __fadd_rn(ag, bg); ab = __fadd_rn(ab, bg); }
          io.comp_rgb[ray * 3 + 0] = ar; io.comp_rgb[ray * 3 + 1] = ag; io.comp_rgb[ray * 3 + 2] = ab;
          if (io.comp_depth != nullptr) io.comp_depth[ray] = ad;
          if (io.comp_acc != nullptr) io.comp_acc[ray] = aa;
        }
      }
      __syncwarp();                               // every lane's reads of the staging rows are done
      if (lane == 0) mbar_arrive(bar(kBarCompDone + t_));
    };
    // ---- fused K6+K7 (fine pass, S == 128: one ray per tile): producer warp t runs utils.sample_pdf + the sorted union
    // (sample_fine.cuh: the stand-alone kernel's code, bit-identical) for the ray of tile (it_, t) and writes its ts_fine row; the
    // rows of iteration it+1 are computed at the end of iteration it, while these warps would otherwise wait for the next PE
    // release.  Scratch: the second-ray halves of the view-direction buffers (unused when a tile holds one ray) + a shared ts row.
    const bool fuse_sf = !kTrain && !kTrunk && io.fine_weights != nullptr;
    struct FinePtrs { float* ts; float* w; float* cdf; float* out; float* smp; };
    auto sample_tile = [&](int64_t it_, int t_) {
      const int64_t tile = my_tile(it_, t_, rank);
      if (tile >= P.ntiles) return;
      float* const spare = reinterpret_cast<float*>(smem + kOffDirBias) + (t_ * 2 * kMaxRaysPerTile + 1) * 128;   // (slot t_, buffer 0, ray 1)
      FinePtrs fs;
      fs.ts = reinterpret_cast<float*>(smem + kOffTsRow);
      fs.w = spare;                                  // 62 values; the cdf (63) takes its place once the pdf is in registers
      fs.cdf = spare;
      fs.smp = spare + 64;
      fs.out = spare + kMaxRaysPerTile * 128;        // (slot t_, buffer 1, ray 1): 128 floats
      const float* wsrc = io.fine_weights + tile * 64;   // one ray per tile
      // the lean 64 + 64 form of sample_fine_core (sample_fine.cuh; bit-identical, a third of the instructions): this lane's two raw
      // weights / coarse positions in registers, the shared coarse row in fs.ts, cdf and union scratch in the spare rows
      const float2 x2 = __ldcg(reinterpret_cast<const float2*>(wsrc) + lane);
      const float2 t2 = *reinterpret_cast<const float2*>(fs.ts + 2 * lane);
      float rr[4];
      sample_fine64_lean_core(fs.ts, fs.cdf, fs.out, x2, t2, lane, rr);
      float* dst = io.fine_ts + tile * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[lane + 32 * i] = rr[i];
      __syncwarp();
    };
    if (fuse_sf) {
      if (r < 64) reinterpret_cast<float*>(smem + kOffTsRow)[r] = coarse_t(r, 64, io.fine_t_scale, io.fine_t_near);
      named_bar_sync(4, kNumPeThreads);
      if ((r >> 5) < 2 && iters > 0) sample_tile(0, r >> 5);
      named_bar_sync(4, kNumPeThreads);          // the rows are visible to every producer thread (CTA-scope ordering)
    }
    for (int64_t it = 0; it <= iters; ++it) {
#pragma unroll 1
      for (int t = 0; t < 2 && it < iters; ++t) {
        const int64_t tile = my_tile(it, t, rank);
        if (it > 0) mbar_wait_relaxed(bar(kBarPeFree + t), (uint32_t)((it - 1) & 1), 128);
        if (r == 0) TC_TRACE(3, it, 0, t, 0);
        // ---- sample position of this row
        int64_t m = tile * kTileM + r;
        if (m >= P.M) m = P.M - 1;  // padding rows of the last tile: recompute a valid sample, never stored
        const int64_t ray = m / S;
        float x[3];
        if (io.rays_o != nullptr) {
          const int k = (int)(m - ray * S);
          const float tt = io.ts != nullptr ? __ldcg(io.ts + m) : coarse_t(k, S, io.t_scale, io.t_near);   // (.cg: fused K6+K7 rows come from another warp)
#pragma unroll
          for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(io.rays_o[ray * 3 + c], __fmul_rn(tt, io.rays_d[ray * 3 + c]));
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) x[c] = io.pts[m * 3 + c];
        }
        // ---- positional encoding row (63 values + zero pad) -> bf16, SW128 K-major
        const uint32_t prow = sbase + kOffPe + t * kPeBytes + (r >> 3) * 1024 + (r & 7) * 128;
        float e[64];
        e[0] = x[0]; e[1] = x[1]; e[2] = x[2];
#pragma unroll
        for (int f = 0; f < 10; ++f) {
          const float fr = (float)(1 << f);
#pragma unroll
          for (int a = 0; a < 3; ++a) fast_sincos(__fmul_rn(x[a], fr), &e[3 + 6 * f + a], &e[3 + 6 * f + 3 + a]);
        }
        e[63] = 0.f;
        uint8_t* const gpe = (kTrain && tile < P.ntiles) ? P.stash_pe + (size_t)tile * 16384 + (r >> 3) * 1024 + (r & 7) * 128 : nullptr;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t q0 = pack_op<kF16>(e[8 * ch + 0], e[8 * ch + 1]), q1 = pack_op<kF16>(e[8 * ch + 2], e[8 * ch + 3]),
                         q2 = pack_op<kF16>(e[8 * ch + 4], e[8 * ch + 5]), q3 = pack_op<kF16>(e[8 * ch + 6], e[8 * ch + 7]);
          st_shared_v4(prow + ((ch ^ (r & 7)) << 4), q0, q1, q2, q3);
          if (gpe != nullptr) st_global_v4(gpe + ((ch ^ (r & 7)) << 4), q0, q1, q2, q3);
        }
        // ---- per-ray view-direction term of rgb0: b_rgb0[j] + sum_k W_dir[k][j] * dirPE[k]   (j = r)
        {
          float* db = reinterpret_cast<float*>(smem + kOffDirBias) + ((t * 2 + (int)(it & 1)) * kMaxRaysPerTile) * 128;
          const float* wd = P.smalls + kSmWDir;
          const int64_t ray0 = (tile * kTileM) / S;
          for (int rs = 0; rs < P.rays_per_tile; ++rs) {
            int64_t rr = ray0 + rs;
            if (rr >= io.n_rays) rr = io.n_rays - 1;
            const float* dsrc = io.rays_o != nullptr ? io.rays_d : io.dirs;
            const float v[3] = {dsrc[rr * 3 + 0], dsrc[rr * 3 + 1], dsrc[rr * 3 + 2]};
            float acc = P.smalls[kSmBiasRgb0 + r];
#pragma unroll
            for (int a = 0; a < 3; ++a) acc = fmaf(wd[a * 128 + r], v[a], acc);
#pragma unroll 1
            for (int f = 0; f < 4; ++f) {
              const float fr = (float)(1 << f);
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                float sn, cs;
                fast_sincos(__fmul_rn(v[a], fr), &sn, &cs);
                acc = fmaf(wd[(3 + 6 * f + a) * 128 + r], sn, acc);
                acc = fmaf(wd[(3 + 6 * f + 3 + a) * 128 + r], cs, acc);
              }
            }
            db[rs * 128 + r] = acc;
          }
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(bar(kBarPeReady + t));          // local: this CTA's epilogue acquires the dir term
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_pepair + 8u * t);   // pair: the leader's MMA warp
        if (r == 0) TC_TRACE(3, it, 0, t, 1);
      }
      // fused K5: this iteration's PE tiles are out (they were released at the previous tiles' layer 5); the previous tiles' last
      // epilogues run about now -- composite them while waiting for the next PE release (the pass after the last iteration
      // composites the last tiles).  One call site: the kernel's instruction footprint is performance-relevant.
      if (comp && it > 0) {
#pragma unroll 1
        for (int t = 0; t < 2; ++t) composite_tile(it - 1, t);
      }
      // fused K6+K7: the next iteration's ts_fine rows
      if (fuse_sf && it + 1 < iters) {
        if ((r >> 5) < 2) sample_tile(it + 1, r >> 5);
        named_bar_sync(4, kNumPeThreads);
      }
    }
  } else {
    // =====================================================================
    // epilogue warps
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int hc = (warp - kEpiWarp0) >> 2;   // which half of the columns
    const int row = q * 32 + lane;
    const float* bias_s = reinterpret_cast<const float*>(smem + kOffBias);
    const float* wsig_s = reinterpret_cast<const float*>(smem + kOffWSig);
    const float* wrgb_s = reinterpret_cast<const float*>(smem + kOffWRgb1);
    float* sigpart_s = reinterpret_cast<float*>(smem + kOffSigPart);
    float* rgbpart_s = reinterpret_cast<float*>(smem + kOffRgbPart);
    const float b_sigma = P.smalls[kSmBSigma];
    const float b_rgb[3] = {P.smalls[kSmBRgb1 + 0], P.smalls[kSmBRgb1 + 1], P.smalls[kSmBRgb1 + 2]};
    uint32_t acc_par[2] = {0, 0};
    float sig_keep[2] = {0.f, 0.f};
    const int S = P.io.S;

    const uint32_t leader_actready = mapa_cluster(bar(kBarActReady), 0);
    // per-warp arrival on the leader's barrier: every lane has fenced its own writes / TMEM loads before the warp sync
    auto act_arrive = [&](int t) {
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_actready + 8u * t);
    };
    for (int64_t it = 0; it < iters; ++it) {
      for (int l = 0; l < nlayers; ++l) {
        for (int t = 0; t < 2; ++t) {
          if (l == 0) mbar_wait(bar(kBarPeReady + t), (uint32_t)(it & 1));  // acquire the producers' dir-bias writes
          mbar_wait(bar(kBarAccFull + t), acc_par[t]); acc_par[t] ^= 1;
          tc_fence_after();
          if (lane == 0 && q == 2) TC_TRACE(1 + hc, it, l, t, 0);
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(256 * t);
          const int64_t tile = my_tile(it, t, rank);
          const int64_t m = tile * kTileM + row;
          if constexpr (kTrunk) {
            // the previous tile's remap image store out of act[t] must have drained before this tile's first epilogue overwrites it:
            // slot 0 may leave the store of slot 1 (issued one epilogue ago) in flight, slot 1 waits for everything
            if (l == 0 && it > 0) {
              if (warp == kEpiWarp0 && lane == 0) { if (t == 0) bulk_wait_read1(); else bulk_wait_read0(); }
              named_bar_sync(3, kNumEpiThreads);
            }
          }

          if (dbg_layers > 0 && l == nlayers - 1) {
            // test hook: dump the raw fp32 accumulator of the last executed layer
            const int ncols = tc_layer_n(l) / 2;
            for (int b = 0; b < ncols / 32; ++b) {
              uint32_t v[32];
              const int c0 = hc * ncols + b * 32;
              tmem_ld32(taddr + c0, v);
              tmem_ld_wait();
              if (m < P.M)
                for (int j = 0; j < 32; ++j) P.dbg_out[m * 256 + c0 + j] = __uint_as_float(v[j]);
            }
            tc_fence_before();
            act_arrive(t);
            continue;
          }

          if (l < 9) {
            // hidden layer: bias + ReLU -> bf16 A operand of the next layer (in place)
            const float* bl = bias_s + l * 256 + hc * 128;
            const uint32_t arow = sbase + kOffAct + t * kActBytes + (row >> 3) * 1024 + (row & 7) * 128 + hc * 2 * 16384;
            const uint32_t tcol = taddr + hc * 128;
            const uint32_t rx = (uint32_t)(row & 7) << 4;
            uint32_t* mrow = nullptr;   // training: this thread's first mask word of the layer (the image itself leaves by bulk store)
            if constexpr (kTrain) {
              if (P.stash_mask != nullptr && tile < P.ntiles && !(dbg_flags & 128)) mrow = P.stash_mask + (((size_t)tile * 10 + l) * 8 + hc * 4) * 128 + row;
            }
            if (dbg_flags & 2) {
            } else if (l == 7) {
              const float sig = hidden_epilogue<true, kTrain, kF16>(tcol, bl, wsig_s + hc * 128, arow, rx, mrow);
              if (hc == 1) sigpart_s[t * 128 + row] = sig; else sig_keep[t] = sig;
            } else {
              // fused K5: the previous tile's staged (r,g,b,sigma) rows occupy the tail of act[t] (K block 3, rows 96..127), which
              // only the last two blocks of the (hc 1, quarter 3) warp overwrite at layer 0: the producers must have composited
              // them by then (long done).  One call site: a second inlined copy of the epilogue costs more than the branch.
              uint32_t wbar = 0u, wpar = 0u;
              if constexpr (!kTrain && !kTrunk) {
                if (l == 0 && it > 0 && hc == 1 && q == 3 && P.io.comp_rgb != nullptr) { wbar = bar(kBarCompDone + t); wpar = (uint32_t)((it - 1) & 1); }
              }
              hidden_epilogue<false, kTrain, kF16>(tcol, bl, nullptr, arow, rx, mrow, wbar, wpar);
            }
            fence_proxy_async();
            if constexpr (kTrain) {
              // activation stash: the finished tile image leaves as ONE 64 KB bulk store.  The issuer first waits until its
              // previous store (the other slot, one epilogue ago) has finished reading shared memory -- that buffer is the one
              // the NEXT epilogue overwrites, and every thread passes the named barrier after this wait.
              const bool issuer = (warp == kEpiWarp0 && lane == 0);
              if (!(dbg_flags & 64)) {
                if (issuer) bulk_wait_read0();
                named_bar_sync(3, kNumEpiThreads);
              }
              if (issuer && tile < P.ntiles && !(dbg_flags & 32)) {
                bulk_s2g(P.stash_h + ((size_t)tile * 9 + l) * 65536, sbase + kOffAct + t * kActBytes, 65536u);
                bulk_commit_group();
              }
            }
            if constexpr (kTrunk) {
              if (l == 8) {
                named_bar_sync(3, kNumEpiThreads);   // every thread's tile writes are fenced: the image may leave
                if (warp == kEpiWarp0 && lane == 0) {
                  if (tile < P.ntiles) bulk_s2g(P.stash_h + (size_t)tile * 65536, sbase + kOffAct + t * kActBytes, 65536u);
                  bulk_commit_group();   // one group per slot and tile (empty for padding tiles): the drain waits count groups
                }
                // trunk mode ends here: sigma (fp32 head, models.py:103) goes to the .w lane of the per-sample float4; the style head
                // kernels fill in (r,g,b).  sigpart_s was written by the hc==1 threads one layer ago (ordered through the
                // ActReady -> MMA -> AccFull chain).
                if (hc == 0 && m < P.M)
                  reinterpret_cast<float*>(P.io.rgbsigma)[m * 4 + 3] = sig_keep[t] + sigpart_s[t * 128 + row] + b_sigma;
              }
            }
            tc_fence_before();
            act_arrive(t);
            if (lane == 0 && q == 2) TC_TRACE(1 + hc, it, l, t, 1);
          } else {
            // rgb0 (N=128): + per-ray dir term, ReLU, then the 3x128 rgb1 head + sigmoid in fp32 (models.py:108-111)
            const float* db = reinterpret_cast<const float*>(smem + kOffDirBias) + ((t * 2 + (int)(it & 1)) * kMaxRaysPerTile) * 128;
            const int rs = (S < kTileM) ? (row / S) : 0;
            const float* dbr = db + rs * 128;
            float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
              uint32_t v[32];
              const int c0 = hc * 64 + b * 32;
              tmem_ld32(taddr + c0, v);
              tmem_ld_wait();
              uint32_t fq[16];  // bf16 pairs of this block's 32 outputs (training stash)
              uint32_t fmask = 0u;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int c = c0 + j;
                const float4 bb = *reinterpret_cast<const float4*>(dbr + c);
                const float4 w0 = *reinterpret_cast<const float4*>(wrgb_s + c);
                const float4 w1 = *reinterpret_cast<const float4*>(wrgb_s + 128 + c);
                const float4 w2 = *reinterpret_cast<const float4*>(wrgb_s + 256 + c);
                const float z0 = __uint_as_float(v[j + 0]) + bb.x, z1 = __uint_as_float(v[j + 1]) + bb.y;
                const float z2 = __uint_as_float(v[j + 2]) + bb.z, z3 = __uint_as_float(v[j + 3]) + bb.w;
                if constexpr (kTrain) {
                  fmask = __funnelshift_l(__float_as_uint(z0), fmask, 1);
                  fmask = __funnelshift_l(__float_as_uint(z1), fmask, 1);
                  fmask = __funnelshift_l(__float_as_uint(z2), fmask, 1);
                  fmask = __funnelshift_l(__float_as_uint(z3), fmask, 1);
                }
                const float f0 = fmaxf(z0, 0.f), f1 = fmaxf(z1, 0.f), f2 = fmaxf(z2, 0.f), f3 = fmaxf(z3, 0.f);
                p0 = fmaf(f0, w0.x, p0); p0 = fmaf(f1, w0.y, p0); p0 = fmaf(f2, w0.z, p0); p0 = fmaf(f3, w0.w, p0);
                p1 = fmaf(f0, w1.x, p1); p1 = fmaf(f1, w1.y, p1); p1 = fmaf(f2, w1.z, p1); p1 = fmaf(f3, w1.w, p1);
                p2 = fmaf(f0, w2.x, p2); p2 = fmaf(f1, w2.y, p2); p2 = fmaf(f2, w2.z, p2); p2 = fmaf(f3, w2.w, p2);
                fq[j / 2] = pack_bf16(f0, f1);
                fq[j / 2 + 1] = pack_bf16(f2, f3);
              }
              if (kTrain && tile < P.ntiles) {
                if (P.stash_mask != nullptr) P.stash_mask[(((size_t)tile * 10 + 9) * 8 + hc * 2 + b) * 128 + row] = fmask;
                // rgb0 output tile image: K block hc (64 columns), 16-byte chunks b*4 .. b*4+3 of this row
                uint8_t* gf = P.stash_f + (size_t)tile * 32768 + hc * 16384 + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                  st_global_v4(gf + (((b * 4 + ch) ^ (row & 7)) << 4), fq[4 * ch], fq[4 * ch + 1], fq[4 * ch + 2], fq[4 * ch + 3]);
              }
            }
            tc_fence_before();
            act_arrive(t);  // accumulator drained: the next tile's layer 0 may start
            if (lane == 0 && q == 2) TC_TRACE(1 + hc, it, l, t, 1);
            if (!kTrain && !kTrunk && P.io.comp_rgb != nullptr && P.io.rgbsigma == nullptr && P.io.rgb == nullptr) {
              // fused K5, nothing per-sample requested (tail_off): stage this column half's partial sums of the rgb1 head (hc 0: + its
              // share of sigma) in the row's 32 staging bytes and go on -- the producer warps sum the halves, add the biases and
              // take the sigmoid before they composite.  No barrier among the epilogue warps, no expf on this critical loop.
              *reinterpret_cast<float4*>(smem + kOffAct + t * kActBytes + kCompStageOff + row * 32 + hc * 16) =
                  make_float4(p0, p1, p2, hc == 0 ? sig_keep[t] : 0.f);
              __syncwarp();
              if (lane == 0) mbar_arrive(bar(kBarCompReady + t));
              continue;
            }
            if (hc == 1) *reinterpret_cast<float4*>(rgbpart_s + row * 4) = make_float4(p0, p1, p2, 0.f);
            named_bar_sync(1, kNumEpiThreads);
            if (hc == 0) {
              const float4 o = *reinterpret_cast<const float4*>(rgbpart_s + row * 4);
              const float z0 = p0 + o.x + b_rgb[0], z1 = p1 + o.y + b_rgb[1], z2 = p2 + o.z + b_rgb[2];
              const float sg = sig_keep[t] + sigpart_s[t * 128 + row] + b_sigma;
              const float r0 = 1.0f / (1.0f + expf(-z0)), r1 = 1.0f / (1.0f + expf(-z1)), r2 = 1.0f / (1.0f + expf(-z2));
              if (m < P.M) {
                if (P.io.rgbsigma != nullptr) {
                  reinterpret_cast<float4*>(P.io.rgbsigma)[m] = make_float4(r0, r1, r2, sg);
                } else if (P.io.rgb != nullptr) {
                  P.io.rgb[m * 3 + 0] = r0; P.io.rgb[m * 3 + 1] = r1; P.io.rgb[m * 3 + 2] = r2;
                  P.io.sigma[m] = sg;
                }
              }
              if (!kTrain && P.io.comp_rgb != nullptr) {
                // fused K5: the row's (r,g,b,sigma) is staged in the slot's activation buffer (free between the rgb0 MMA and the
                // next tile's first epilogue) for the input-producer warps, which composite the tile off this critical path
                *reinterpret_cast<float4*>(smem + kOffAct + t * kActBytes + kCompStageOff + row * 32) = make_float4(r0, r1, r2, sg);
                __syncwarp();                              // one arrival per warp (128 arrivals on one word would serialise)
                if (lane == 0) mbar_arrive(bar(kBarCompReady + t));
              }
            }
            named_bar_sync(2, kNumEpiThreads);  // partial buffers free for the next tile
          }
        }
      }
    }
  }

  // ---- teardown
  if constexpr (kTrain || kTrunk) {
    if (warp == kEpiWarp0 && lane == 0) bulk_wait_all0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA may exit (or free TMEM) while the other can still signal it / read its shared memory
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

bool mlp_tc_supports(const MlpIO& io) {
  if (io.base_remap || io.pts_embed || io.dirs_embed) return false;
  if (io.rays_o == nullptr && !io.dirs_per_ray) return false;
  const int S = io.S;
  if (!(S == 64 || S == 128 || (S > 128 && S % 128 == 0))) return false;
  if (io.rgbsigma != nullptr && !aligned16(io.rgbsigma)) return false;
  if (io.comp_rgb != nullptr && !(io.rays_o != nullptr && (S == 64 || S == 128))) return false;   // fused compositing: whole rays per tile
  if (io.fine_weights != nullptr && !(io.rays_o != nullptr && S == 128 && io.fine_ts != nullptr && io.fine_ts == io.ts)) return false;   // fused resampling
  return true;
}

static int g_dbg_flags = 0;
static long long* g_dbg_trace = nullptr;
extern "C" void tgtc_debug_tc_flags(int f) { g_dbg_flags = f; }
extern "C" void tgtc_debug_tc_trace(long long* dev_buf) { g_dbg_trace = dev_buf; }

static int launch_tc_common(tgtc_ctx* ctx, int net, const MlpIO& io, int dbg_layers, float* dbg_out, const TcStash* stash,
                            cudaStream_t st, int trunk = 0, bool f16 = false) {
  const NetImage& im = ctx->net[net];
  TcParams P;
  P.blob = f16 ? im.tc_blob_h : im.tc_blob;
  P.smalls = im.smalls;
  P.io = io;
  P.M = io.n_rays * io.S;
  if (P.M == 0) return TGTC_OK;
  P.ntiles = (P.M + kTileM - 1) / kTileM;
  P.rays_per_tile = io.S < kTileM ? kTileM / io.S : 1;
  P.dbg_layers = dbg_layers;
  P.trunk = trunk;
  P.dbg_flags = g_dbg_flags;
  P.stash_h = stash != nullptr ? stash->h : nullptr;
  P.stash_f = stash != nullptr ? stash->f : nullptr;
  P.stash_pe = stash != nullptr ? stash->pe : nullptr;
  P.stash_mask = stash != nullptr ? stash->mask : nullptr;
  P.dbg_trace = g_dbg_trace;
  P.dbg_out = dbg_out;
  static bool attr_set[64] = {};
  if (!attr_set[ctx->device & 63]) {
    TGTC_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    TGTC_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    TGTC_CUDA((cudaFuncSetAttribute(mlp_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
    TGTC_CUDA((cudaFuncSetAttribute(mlp_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
    TGTC_CUDA((cudaFuncSetAttribute(mlp_tc_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
    TGTC_CUDA((cudaFuncSetAttribute(mlp_tc_kernel<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
    TGTC_CUDA((cudaFuncSetAttribute(mlp_tc_kernel<false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
    TGTC_CUDA((cudaFuncSetAttribute(mlp_tc_kernel<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
    attr_set[ctx->device & 63] = true;
  }
  const int64_t nquads = (P.ntiles + 3) / 4;
  const int64_t max_pairs = ctx->num_sms / 2;
  const int grid = 2 * (int)(nquads < max_pairs ? nquads : max_pairs);   // CTA pairs (clusters of 2)
  const bool dbg = dbg_layers > 0 || g_dbg_flags != 0 || g_dbg_trace != nullptr;   // test / timing hooks: their own instantiations
  if (trunk && f16) mlp_tc_kernel<false, true, true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else if (trunk) mlp_tc_kernel<false, true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else if (stash != nullptr && dbg) mlp_tc_kernel<true, false, false, true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else if (stash != nullptr) mlp_tc_kernel<true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else if (f16 && dbg) mlp_tc_kernel<false, false, true, true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else if (f16) mlp_tc_kernel<false, false, true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else if (dbg) mlp_tc_kernel<false, false, false, true><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  else mlp_tc_kernel<false><<<grid, kNumThreads, kSmemBytes, st>>>(P);
  TGTC_LAUNCH_CHECK(ctx);
  return TGTC_OK;
}

// style path: NeRF trunk only (L0..L7, sigma head, remap); remap tile images -> remap_img [ntiles][64 KB], sigma -> rgbsigma[.].w
int launch_mlp_tc_trunk(tgtc_ctx* ctx, int net, const MlpIO& io, uint8_t* remap_img, cudaStream_t st, bool f16) {
  TcStash stash;
  stash.h = remap_img;
  return launch_tc_common(ctx, net, io, 0, nullptr, &stash, st, 1, f16);
}

// training forward: same kernel, additionally writing the activation stash the backward kernels read
int launch_mlp_tc_train(tgtc_ctx* ctx, int net, const MlpIO& io, const TcStash& stash, cudaStream_t st) {
  return launch_tc_common(ctx, net, io, 0, nullptr, &stash, st);
}

int launch_mlp_tc(tgtc_ctx* ctx, int net, const MlpIO& io, cudaStream_t st, bool f16) {
  return launch_tc_common(ctx, net, io, 0, nullptr, nullptr, st, 0, f16);
}

// test hook (not part of the public header): run the first `layers` GEMM layers of the bf16 kernel and dump
// the raw fp32 accumulator of the last one ([n_rays*S, 256], columns >= N untouched)
static int g_dbg_f16 = 0;
extern "C" void tgtc_debug_tc_f16(int on) { g_dbg_f16 = on; }   // the hook below runs the fp16-operand instantiation
extern "C" int tgtc_debug_tc_layers(tgtc_ctx* ctx, int net, const float* rays_o, const float* rays_d, const float* ts,
                                    int64_t n_rays, int S, double near, double far, int layers, float* acc_out, void* stream) {
  if (ctx == nullptr || !ctx->net[net].set || layers < 1 || layers > kTcNumGemm) { tgtc_set_error("bad debug args"); return TGTC_ERR_ARG; }
  DeviceGuard g(ctx->device);
  MlpIO io;
  io.rays_o = rays_o; io.rays_d = rays_d; io.ts = ts;
  io.t_scale = (float)(far - near); io.t_near = (float)near;
  io.n_rays = n_rays; io.S = S;
  if (!mlp_tc_supports(io)) { tgtc_set_error("unsupported S"); return TGTC_ERR_UNSUPPORTED; }
  return launch_tc_common(ctx, net, io, layers, acc_out, nullptr, (cudaStream_t)stream, 0, g_dbg_f16 != 0);
}
