// Counter-based RNG for the reference's stochastic training options: stratified jitter
// (torch.nn.init.uniform_ in utils.py:519-520) and the sigma noise (torch.randn * std in utils.py:372-374), generated inside
// the kernels that consume them.  Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11):
// the value of element e of stream s under seed k is a pure function of (k, s, e), so the compositing backward regenerates
// exactly the noise the forward saw, results do not depend on the launch geometry, and a CPU restatement
// (oracle/philox_oracle.py) reproduces every uniform bit for bit.  The reference draws from torch's global generator; its
// streams are not reproduced (they depend on torch's launch geometry) -- only the distributions are.
#pragma once
#include <stdint.h>

struct PhiloxSrc {
  unsigned long long seed;   // key
  uint32_t stream;           // which tensor of the step (TGTC_STREAM_*)
  float std;                 // scale of the normal draws
  int on;
};
constexpr uint32_t kStreamJitter = 0, kStreamNoiseCoarse = 1, kStreamNoiseFine = 2;

#ifdef __CUDACC__
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ uint32_t philox_word(const uint4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
// element e of (seed, stream): counter = (e/4 lo, e/4 hi, stream, 0), word e%4
__device__ __forceinline__ float philox_uniform(unsigned long long seed, uint32_t stream, uint64_t e) {
  const uint64_t blk = e >> 2;
  const uint4 v = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), stream, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return (float)(philox_word(v, (int)(e & 3)) >> 8) * 5.9604644775390625e-8f;   // 24 bits -> [0,1)
}
// element e: counter = (e/2 lo, e/2 hi, stream, 1), words 2*(e%2), 2*(e%2)+1 -> one Box-Muller cosine branch
__device__ __forceinline__ float philox_normal(unsigned long long seed, uint32_t stream, uint64_t e) {
  const uint64_t blk = e >> 1;
  const uint4 v = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), stream, 1u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const int w = (int)(e & 1) * 2;
  const float u1 = (float)((philox_word(v, w) >> 8) + 1u) * 5.9604644775390625e-8f;     // (0,1]
  const float u2 = (float)(philox_word(v, w + 1) >> 8) * 5.9604644775390625e-8f;        // [0,1)
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
#endif
