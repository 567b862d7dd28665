// PTX wrappers shared by the tcgen05 kernels (mlp_tc.cu, mlp_bwd.cu): mbarriers, bulk copies, cluster addressing,
// tcgen05 MMA / commit / TMEM loads, UMMA shared-memory and instruction descriptors.
#pragma once
#include "common.cuh"

namespace tcptx {

// ---------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// the slow path of every bounded wait, out of line: a printf call inlined at ~20 wait sites per kernel costs instruction-cache
// footprint that the hot loops pay for (removing the test hooks' code from the production kernels alone was worth 4 %)
static __device__ __noinline__ void mbar_timeout_trap(uint32_t bar, uint32_t parity) {
  printf("tgtc tcgen05 kernel: mbarrier timeout (block %d thread %d bar@%u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, bar, parity);
  __trap();
}
// bounded wait: a protocol bug traps (visible as a launch failure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout_trap(bar, parity);  // ~2 s
  }
}
// warp-uniform wait for converged single-issuer warps: the loop condition is a vote, so control flow stays uniform and
// ptxas can keep descriptors / barrier addresses in uniform registers across the wait
__device__ __forceinline__ void mbar_wait_uniform(uint32_t bar, uint32_t parity) {
  if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return;
  const long long t0 = clock64();
  while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout_trap(bar, parity);
  }
}
// same, for roles that are far off the critical path: sleep between probes so the spin does not steal issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned ns) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if (clock64() - t0 > 4000000000LL) mbar_timeout_trap(bar, parity);
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// shared -> global bulk store (TMA engine), tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// all of this thread's earlier bulk stores have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// address of a shared-memory object of CTA `rank` of this cluster, in the shared::cluster window
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release at CTA scope), as cutlass::arch::ClusterBarrier::arrive(cta_id): a cluster-scope release would
  // make ptxas emit an L1 invalidate + membar on every arrive (measured: ~1 us per arrive)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// tcgen05.mma with the descriptors given as (lo, hi) halves: the hi halves are compile-time constants and the lo
// halves advance by (byte offset >> 4), so one K step costs one integer add per operand
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (the same lane every time for a full mask)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xFFFFFFFF;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred));
  return pred != 0;
}
// sin/cos for the bf16 operand path: two-constant Cody-Waite reduction to [-pi, pi] (exact to ~2e-7 for |a| < 4096),
// then the SFU approximations (abs error 2^-21.4 on that interval) -- three orders below a bf16 ulp
__device__ __forceinline__ void fast_sincos(float a, float* s, float* c) {
  const float n = rintf(a * 0.15915494309189535f);
  float r = fmaf(n, -6.2831854820251465f, a);
  r = fmaf(n, 1.7484555e-7f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}
// arrive (once the MMAs issued so far by this thread retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// wait::ld that also "defines" the 32 destination registers, so no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                 "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                 "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
// packed fp32x2 add (sm_100 FADD2): {x0,x1} += {b0,b1}
__device__ __forceinline__ void add2(uint32_t& x0, uint32_t& x1, float b0, float b1) {
  asm("{\n.reg .b64 a, b, d;\nmov.b64 a, {%0, %1};\nmov.b64 b, {%2, %3};\nadd.rn.f32x2 d, a, b;\nmov.b64 {%0, %1}, d;\n}\n"
      : "+r"(x0), "+r"(x1)
      : "f"(b0), "f"(b1));
}
// relu + round-to-nearest bf16 + pack in one instruction (F2FP.RELU)
__device__ __forceinline__ uint32_t pack_bf16_relu(uint32_t lo, uint32_t hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): K-major, swizzled; built as (lo, hi) words by the kernels:
//   lo = (addr & 0x3FFFF) >> 4 | LBO field, hi = kDescHiSW128
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major; 1 as CUTLASS) | [32,46) SBO>>4 |
//   [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
constexpr uint32_t kDescHiSW128 = (1024u >> 4) | (1u << 14) | (2u << 29);

// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// ---- operand format of the inference kernels: bf16 (the training format) or fp16 (TGTC_MLP_F16: 11-bit significand, 8x smaller
// operand rounding error at the same tensor-core rate; fp32 values beyond +-65504 saturate instead of becoming inf)
template <bool kF16>
__device__ __forceinline__ uint32_t pack_op(float lo, float hi) {
  uint32_t r;
  if constexpr (kF16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <bool kF16>
__device__ __forceinline__ uint32_t pack_op_relu(uint32_t lo, uint32_t hi) {
  uint32_t r;
  if constexpr (kF16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}
// instruction descriptor with the A/B format field: 1 = bf16, 0 = f16 (cute::UMMA::F16F32Format)
template <bool kF16>
__host__ __device__ constexpr uint32_t make_idesc_op(int M, int N) {
  return (1u << 4) | (kF16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }

__device__ __forceinline__ void st_global_v4(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}
// ReLU mask word of 32 fp32 pre-activations: bit (31-c) = sign bit of v[c] (1 = the gradient is blocked)
__device__ __forceinline__ uint32_t sign_mask32(const uint32_t (&v)[32]) {
  uint32_t m = 0u;
#pragma unroll
  for (int c = 0; c < 32; ++c) m = __funnelshift_l(v[c], m, 1);
  return m;
}
// two fp32 gradients -> bf16 pair, zeroed where the ReLU blocked them: bit 31 of mb belongs to lo, bit 30 to hi (1 = blocked)
__device__ __forceinline__ uint32_t mask_pack(uint32_t mb, float lo, float hi) {
  const float l = (int32_t)mb < 0 ? 0.f : lo;
  const float h = (int32_t)(mb << 1) < 0 ? 0.f : hi;
  return pack_bf16(l, h);
}

// ---- cta_group::1 MMAs with MN-major operands (weight-gradient kernels: K = samples)
// MN-major, 128B-swizzled operand descriptor: 64-feature atoms 8 KB apart (LBO), 8-sample row groups 1 KB apart (SBO)
__device__ __forceinline__ uint32_t mn_desc_hi() { return (1024u >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint32_t mn_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) {   // both operands MN-major (bits 15, 16)
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma1_bf16(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma1_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

}  // namespace tcptx
