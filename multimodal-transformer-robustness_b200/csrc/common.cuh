// common.cuh -- shared device helpers for libmultb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <utility>
#include "../../include/multb200.h"

#ifndef __CUDA_ARCH__
#define MTB_HOST_ONLY 1
#endif

namespace mtb {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
#define MTB_CHECK(cond, ...)                  \
  do {                                        \
    if (!(cond)) {                            \
      mtb::set_error(__VA_ARGS__);            \
      return -1;                              \
    }                                         \
  } while (0)
#define MTB_CUDA(expr)                                                         \
  do {                                                                         \
    cudaError_t e__ = (expr);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      mtb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),  \
                     __FILE__, __LINE__);                                      \
      return -2;                                                               \
    }                                                                          \
  } while (0)

template <typename D>
struct Group {          // descriptors travel by value in the kernel parameter block
  D d[MTB_MAX_GROUP];
  int start[MTB_MAX_GROUP + 1];   // prefix sum of work items (CTAs) per problem
  int n;
};

// blockIdx -> (problem, local block) by linear scan (n <= 24)
template <typename D>
__device__ __forceinline__ int find_problem(const Group<D>& g, int blk, int& local) {
  int p = 0;
#pragma unroll 1
  while (p + 1 < g.n && blk >= g.start[p + 1]) ++p;
  local = blk - g.start[p];
  return p;
}

// ---------------------------------------------------------------- Philox4x32-7
struct Philox {
  uint32_t k0, k1;
  __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    // exactly one IMAD.WIDE.U32 (the C++ form below makes the compiler append a 64-bit "+ 0" per product)
    asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#else
    const uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
  }
  __host__ __device__ inline uint4 operator()(uint64_t ctr) const {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x5bd1e995u, c3 = 0u;
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 7; ++r) {          // Philox4x32-7 (the round count curand also offers; passes BigCrush)
      uint32_t h0, l0, h1, l1;
      mulhilo(0xD2511F53u, c0, h0, l0);
      mulhilo(0xCD9E8D57u, c2, h1, l1);
      uint32_t n0 = h1 ^ c1 ^ a, n1 = l1, n2 = h0 ^ c3 ^ b, n3 = l0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

struct DropCtx {        // resolved per kernel: effective seed/offset, integer threshold, 1/(1-p)
  uint64_t seed, offset;
  uint32_t thr;
  float inv_keep;
  bool on;
};

__device__ __forceinline__ DropCtx make_drop(const mtb_rng& r, float p) {
  DropCtx c;
  c.on = p > 0.f;
  c.seed = r.seed; c.offset = r.offset;
  if (c.on && r.dev != nullptr) { c.seed += r.dev[0]; c.offset += r.dev[1]; }
  double t = (double)p * 4294967296.0;
  c.thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  c.inv_keep = c.on ? 1.f / (1.f - p) : 1.f;
  return c;
}

// keep-bits for the 4 elements of group `grp` (elements 4*grp .. 4*grp+3)
__device__ __forceinline__ uint4 drop_rand4(const DropCtx& c, uint64_t grp) {
  return Philox(c.seed)(c.offset + grp);
}
__device__ __forceinline__ bool drop_keep1(const DropCtx& c, uint64_t idx) {
  uint4 r = drop_rand4(c, idx >> 2);
  uint32_t v = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return v >= c.thr;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- bf16 <-> fp32 (bf16 data path: activations between kernels)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {      // round-to-nearest-even, lo in bits [0,16)
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float bf16_round(float x) { return bf16_lo(pack_bf16x2(x, 0.f)); }
// 4 consecutive elements at ELEMENT offset `off` of a tensor that is fp32 or bf16 (8- / 16-byte aligned accesses)
__device__ __forceinline__ float4 ld4_any(const void* base, int64_t off, bool bf16) {
  if (bf16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + off);
    return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
}
__device__ __forceinline__ void st4_any(void* base, int64_t off, float4 v, bool bf16) {
  if (bf16) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + off) = u;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = v;
  }
}
__device__ __forceinline__ float ld1_any(const void* base, int64_t off, bool bf16) {
  return bf16 ? __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(base)[off] << 16) : reinterpret_cast<const float*>(base)[off];
}
__device__ __forceinline__ void st1_any(void* base, int64_t off, float v, bool bf16) {
  if (bf16) reinterpret_cast<uint16_t*>(base)[off] = (uint16_t)(pack_bf16x2(v, 0.f) & 0xffffu);
  else reinterpret_cast<float*>(base)[off] = v;
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every hot kernel waits (griddepcontrol.wait) until the preceding kernel of the stream has completed and flushed its
// writes, and lets the NEXT kernel's CTAs be scheduled (griddepcontrol.launch_dependents) while this one is still
// running -- they park in their own wait.  Launched through launch_k() with the programmatic-stream-serialization
// attribute this removes the ~2-3 us launch bubble between the small dependent kernels of a step; without the
// attribute (or after a kernel that never triggers) both instructions are no-ops / the edge is a normal full
// serialisation.
//   pdl_sync()                : wait, then trigger -- the first statement of the HBM-bound kernels
//   pdl_begin() / pdl_ready() : the tcgen05 kernels trigger first, run their prologue (mbarrier init, TMEM allocation,
//                               tensor-map prefetch: nothing that touches global memory) under the tail of the
//                               preceding kernel, and wait right before their first global access.  EVERY thread of
//                               every CTA executes pdl_ready() before it exits, so "kernel N complete" still implies
//                               "kernel N-1 complete" for the kernel after it.
//   -DMTB_PDL_EARLY_WAIT      : variant build with the wait back at the top (A/B measurement).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }
#ifdef MTB_PDL_EARLY_WAIT
__device__ __forceinline__ void pdl_begin() { pdl_sync(); }
__device__ __forceinline__ void pdl_ready() {}
#else
__device__ __forceinline__ void pdl_begin() { pdl_trigger(); }
__device__ __forceinline__ void pdl_ready() { pdl_wait(); }
#endif
extern int g_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int sm_count();
void note_launch();
extern int g_gemm_mode;

}  // namespace mtb
