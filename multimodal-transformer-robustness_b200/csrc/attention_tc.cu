// attention_tc.cu -- fused flash-style attention on tcgen05 / TMEM (tensor-core mode).
// modules/dynamic_multihead_attention.py:91-116, modules/transformer.py:145-157.
//
// One CTA (128 threads) per (batch, head, 128-row query tile); thread t owns query row t, which
// is TMEM lane t.  Per 64-row key tile:
//   S  = Q K^T    tcgen05.mma.kind::tf32, A = Q tile, B = K tile (both K-major, head_dim zero-padded
//                 25 -> 32 IN SHARED MEMORY ONLY), fp32 accumulator in TMEM columns [0, 64)
//   softmax       tcgen05.ld of the row, causal-with-offset predicate, online max / sum in fp32,
//                 Philox attention dropout (same element indexing as the fp32 engine)
//   P~ -> smem    written as the K-major A operand of the second MMA (128-byte swizzle done by hand)
//   PV = P~ V     tcgen05.mma, B = V tile MN-major (32-bit MN-major operands need the
//                 SWIZZLE_128B_BASE32B layout), accumulator in TMEM columns [64, 96)
//   O  = O * corr + PV   in registers (32 fp32 per thread)
// Scores and probabilities never touch HBM.  q / k / v rows are 25 contiguous floats at a
// 100-byte aligned offset, which TMA cannot address, so tiles are staged with plain coalesced
// loads and written to shared memory in the swizzled operand layouts.
#include "common.cuh"
#include <math_constants.h>

namespace mtb {

constexpr int TQ = 128;            // query rows per CTA (= UMMA M = TMEM lanes)
constexpr int TK = 64;             // key rows per inner tile
constexpr int HP = 32;             // padded head dim (one 128-byte swizzle row of fp32)
constexpr int ATC_THREADS = 128;
constexpr int ATC_TMEM_COLS = 128;

__device__ __forceinline__ uint32_t a_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void a_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool a_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void a_mbar_wait(uint32_t bar, uint32_t parity) {
  for (int it = 0; it < (1 << 22); ++it)
    if (a_mbar_try(bar, parity)) return;
  printf("mtb attn_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void a_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void a_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void a_tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// round-to-nearest TF32 (the tensor core would otherwise truncate the low 13 mantissa bits)
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// byte offset of (row r, byte column cb < 128) inside a tile of 128-byte rows
__device__ __forceinline__ uint32_t swz128(int r, int cb) {        // SWIZZLE_128B (16-byte atoms, 8-row period)
  return (uint32_t)(r * 128 + ((((cb >> 4) ^ (r & 7)) << 4) | (cb & 15)));
}
__device__ __forceinline__ uint32_t swz128_32(int r, int cb) {     // SWIZZLE_128B_BASE32B (32-byte atoms, 4-row period)
  return (uint32_t)(r * 128 + ((((cb >> 5) ^ (r & 3)) << 5) | (cb & 31)));
}
// smem matrix descriptors (see linear_tc.cu)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(16u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(4096u >> 4) << 16) | ((uint64_t)(512u >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)1 << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn ? 1 : 0) << 15) | ((uint32_t)(b_mn ? 1 : 0) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// stage `rows` token rows (token l0 + r, batch b, head h) of a token-major matrix into a tile of
// 128-byte rows; MNSW selects the 32-byte-atom swizzle.  Zero padding for d >= hd and l >= L.
// Rows are 25 floats at 100-byte offsets (TMA cannot address them), so each element travels as a
// 4-byte cp.async (LDGSTS) with zero-fill straight into its swizzled slot: no registers, and ALL
// copies of a tile are in flight at once -- one L2 latency per tile instead of one per element
// (the first version's load -> convert -> store loop spent 70 % of its stall samples here).
template <bool MNSW>
__device__ __forceinline__ void stage_tile(uint8_t* tile, const float* base, int64_t ld, int B, int b, int h, int hd, int l0, int L,
                                           int rows) {
  const uint32_t tbase = a_smem_u32(tile);
  for (int e = threadIdx.x; e < rows * HP; e += ATC_THREADS) {
    const int r = e >> 5, c = e & 31;
    const int l = l0 + r;
    const bool ok = (l < L) && (c < hd);
    const float* src = ok ? base + ((int64_t)l * B + b) * ld + h * hd + c : base;
    const uint32_t dst = tbase + (MNSW ? swz128_32(r, c * 4) : swz128(r, c * 4));
    const int nbytes = ok ? 4 : 0;                           // src-size 0 -> the 4 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
  }
}
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ int a_round4(int x) { return (x + 3) & ~3; }

__global__ void __launch_bounds__(ATC_THREADS) attn_fwd_tc_kernel(const __grid_constant__ Group<mtb_attn_desc> g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_attn_desc& d = g.d[pi];
  const int qtiles = (d.Lq + TQ - 1) / TQ;
  const int bh = local / qtiles, qt = local - bh * qtiles;
  const int b = bh / d.H, h = bh - b * d.H;
  const int i0 = qt * TQ;
  const int off = abs(d.Lk - d.Lq);
  const int Lk4 = a_round4(d.Lk);
  const DropCtx dc = make_drop(d.rng, d.p);
  const int tid = threadIdx.x, warp = tid >> 5;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Qs = smem;                       // [128][128 B]  K-major
  uint8_t* Ks = Qs + TQ * 128;              // [ 64][128 B]  K-major
  uint8_t* Vs = Ks + TK * 128;              // [ 64][128 B]  MN-major (32 B-atom swizzle)
  uint8_t* Ps = Vs + TK * 128;              // 2 regions of [128][128 B], K-major, 32 key columns each

  if (tid == 0) {
    a_mbar_init(a_smem_u32(&bar_s), 1);
    a_mbar_init(a_smem_u32(&bar_o), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(&tmem_slot)), "r"(ATC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  stage_tile<false>(Qs, d.q, d.ldq, d.B, b, h, d.hd, i0, d.Lq, TQ);
  stage_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_s = tmem, t_o = tmem + 64;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;

  const int i = i0 + tid;                    // my query row
  float o[HP];
#pragma unroll
  for (int c = 0; c < HP; ++c) o[c] = 0.f;
  float m_run = -CUDART_INF_F, l_run = 0.f;

  const int i_last = min(d.Lq, i0 + TQ) - 1;
  const int j_end = min(d.Lk, i_last + off + 1);
  const uint32_t id_s = idesc_tf32(TQ, TK, false, false);
  const uint32_t id_o = idesc_tf32(TQ, HP, false, true);
  uint32_t phase = 0;
  for (int j0 = 0; j0 < j_end; j0 += TK, phase ^= 1u) {
    stage_tile<false>(Ks, d.k, d.ldk, d.B, b, h, d.hd, j0, d.Lk, TK);
    stage_tile<true>(Vs, d.v, d.ldv, d.B, b, h, d.hd, j0, d.Lk, TK);
    stage_wait();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HP / 8; ++k)
        a_mma_tf32(t_s, desc_kmajor(a_smem_u32(Qs) + k * 32), desc_kmajor(a_smem_u32(Ks) + k * 32), id_s, k != 0 ? 1u : 0u);
      a_commit(a_smem_u32(&bar_s));
    }
    a_mbar_wait(a_smem_u32(&bar_s), phase);
    tc_fence_after();
    // ---- softmax on my row -----------------------------------------------------------------
    float s[TK];
    {
      float t0[32], t1[32];
      a_tmem_ld32(t_s + lane_addr, t0);
      a_tmem_ld32(t_s + lane_addr + 32, t1);
#pragma unroll
      for (int c = 0; c < 32; ++c) { s[c] = t0[c]; s[32 + c] = t1[c]; }
    }
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int c = 0; c < TK; ++c) {
      const int j = j0 + c;
      const bool open = (j < d.Lk) && (j - i < 1 + off);
      s[c] = open ? s[c] * d.scale : -CUDART_INF_F;
      mx = fmaxf(mx, s[c]);
    }
    const float m_new = fmaxf(m_run, mx);
    const float corr = (m_new == -CUDART_INF_F) ? 1.f : __expf(m_run - m_new);
    float rs = 0.f;
    const int irow = min(i, d.Lq - 1);
#pragma unroll
    for (int c4 = 0; c4 < TK; c4 += 4) {
      float keep[4] = {dc.inv_keep, dc.inv_keep, dc.inv_keep, dc.inv_keep};
      if (dc.on) {
        const uint64_t idx = ((uint64_t)((int64_t)bh * d.Lq + irow)) * (uint64_t)Lk4 + (uint64_t)(j0 + c4);
        const uint4 r = drop_rand4(dc, idx >> 2);
        keep[0] = r.x >= dc.thr ? dc.inv_keep : 0.f; keep[1] = r.y >= dc.thr ? dc.inv_keep : 0.f;
        keep[2] = r.z >= dc.thr ? dc.inv_keep : 0.f; keep[3] = r.w >= dc.thr ? dc.inv_keep : 0.f;
      }
      float pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = (s[c4 + e] == -CUDART_INF_F) ? 0.f : __expf(s[c4 + e] - m_new);
        rs += p;
        pk[e] = to_tf32(p * keep[e]);
      }
      uint8_t* region = Ps + (c4 >> 5) * (TQ * 128);
      *reinterpret_cast<float4*>(region + swz128(tid, (c4 & 31) * 4)) = make_float4(pk[0], pk[1], pk[2], pk[3]);
    }
    l_run = l_run * corr + rs;
    m_run = m_new;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < TK / 8; ++kk)
        a_mma_tf32(t_o, desc_kmajor(a_smem_u32(Ps) + (kk >> 2) * (TQ * 128) + (kk & 3) * 32),
                   desc_mnmajor(a_smem_u32(Vs) + kk * 1024), id_o, kk != 0 ? 1u : 0u);
      a_commit(a_smem_u32(&bar_o));
    }
    a_mbar_wait(a_smem_u32(&bar_o), phase);
    tc_fence_after();
    float pv[32];
    a_tmem_ld32(t_o + lane_addr, pv);
#pragma unroll
    for (int c = 0; c < HP; ++c) o[c] = o[c] * corr + pv[c];
    tc_fence_before();          // my TMEM reads are done before the next tile's MMAs overwrite S / PV
  }
  if (i < d.Lq) {
    const float inv = 1.f / l_run;
    float* op = d.o + ((int64_t)i * d.B + b) * d.ldo + h * d.hd;
#pragma unroll
    for (int c = 0; c < HP; ++c)
      if (c < d.hd) op[c] = o[c] * inv;
    if (d.lse) d.lse[(int64_t)bh * d.Lq + i] = m_run + logf(l_run);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ATC_TMEM_COLS) : "memory");
  }
}


// ============================================================================ backward: dQ (+ delta)
// Per 64-row key tile: S = Q K^T and dP = dO V^T on the tensor core, dS = P * (dP * keep - delta) * scale
// on CUDA cores, dQ += dS K on the tensor core (accumulating in TMEM across key tiles).
constexpr int ATC_DQ_TMEM = 256;
constexpr int ATC_DQ_SMEM = 2 * TQ * 128 + 3 * TK * 128 + (TK / 32) * TQ * 128 + 1024;

__global__ void __launch_bounds__(ATC_THREADS) attn_bwd_dq_tc_kernel(const __grid_constant__ Group<mtb_attn_bwd_desc> g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a, bar_b;
  __shared__ uint32_t tmem_slot;
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_attn_bwd_desc& d = g.d[pi];
  const int qtiles = (d.Lq + TQ - 1) / TQ;
  const int bh = local / qtiles, qt = local - bh * qtiles;
  const int b = bh / d.H, h = bh - b * d.H;
  const int i0 = qt * TQ;
  const int off = abs(d.Lk - d.Lq);
  const int Lk4 = a_round4(d.Lk);
  const DropCtx dc = make_drop(d.rng, d.p);
  const int tid = threadIdx.x, warp = tid >> 5;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Qs = smem;                        // [128][128 B] K-major   (A of S)
  uint8_t* dOs = Qs + TQ * 128;              // [128][128 B] K-major   (A of dP)
  uint8_t* Ks = dOs + TQ * 128;              // [ 64][128 B] K-major   (B of S)
  uint8_t* Vs = Ks + TK * 128;               // [ 64][128 B] K-major   (B of dP)
  uint8_t* Kmn = Vs + TK * 128;              // [ 64][128 B] MN-major  (B of dQ)
  uint8_t* dSs = Kmn + TK * 128;             // 2 regions [128][128 B] K-major (A of dQ)

  if (tid == 0) {
    a_mbar_init(a_smem_u32(&bar_a), 1);
    a_mbar_init(a_smem_u32(&bar_b), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(&tmem_slot)), "r"(ATC_DQ_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  stage_tile<false>(Qs, d.q, d.ldq, d.B, b, h, d.hd, i0, d.Lq, TQ);
  stage_tile<false>(dOs, d.d_o, d.lddo, d.B, b, h, d.hd, i0, d.Lq, TQ);
  const int i = i0 + tid;
  float delta = 0.f, lse = 0.f;
  if (i < d.Lq) {
    const float* op = d.o + ((int64_t)i * d.B + b) * d.ldo + h * d.hd;
    const float* gp = d.d_o + ((int64_t)i * d.B + b) * d.lddo + h * d.hd;
    for (int c = 0; c < d.hd; ++c) delta = fmaf(op[c], gp[c], delta);
    lse = d.lse[(int64_t)bh * d.Lq + i];
    d.delta[(int64_t)bh * d.Lq + i] = delta;
  }
  stage_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_s = tmem, t_dp = tmem + 64, t_dq = tmem + 128;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  const int i_last = min(d.Lq, i0 + TQ) - 1;
  const int j_end = min(d.Lk, i_last + off + 1);
  const uint32_t id_s = idesc_tf32(TQ, TK, false, false);
  const uint32_t id_q = idesc_tf32(TQ, HP, false, true);
  const int irow = min(i, d.Lq - 1);
  uint32_t phase = 0;
  int tile = 0;
  for (int j0 = 0; j0 < j_end; j0 += TK, phase ^= 1u, ++tile) {
    stage_tile<false>(Ks, d.k, d.ldk, d.B, b, h, d.hd, j0, d.Lk, TK);
    stage_tile<false>(Vs, d.v, d.ldv, d.B, b, h, d.hd, j0, d.Lk, TK);
    stage_tile<true>(Kmn, d.k, d.ldk, d.B, b, h, d.hd, j0, d.Lk, TK);
    stage_wait();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HP / 8; ++k)
        a_mma_tf32(t_s, desc_kmajor(a_smem_u32(Qs) + k * 32), desc_kmajor(a_smem_u32(Ks) + k * 32), id_s, k != 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < HP / 8; ++k)
        a_mma_tf32(t_dp, desc_kmajor(a_smem_u32(dOs) + k * 32), desc_kmajor(a_smem_u32(Vs) + k * 32), id_s, k != 0 ? 1u : 0u);
      a_commit(a_smem_u32(&bar_a));
    }
    a_mbar_wait(a_smem_u32(&bar_a), phase);
    tc_fence_after();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float s[32], dp[32];
      a_tmem_ld32(t_s + lane_addr + half * 32, s);
      a_tmem_ld32(t_dp + lane_addr + half * 32, dp);
#pragma unroll
      for (int c4 = 0; c4 < 32; c4 += 4) {
        const int jb = j0 + half * 32 + c4;
        float keep[4] = {dc.inv_keep, dc.inv_keep, dc.inv_keep, dc.inv_keep};
        if (dc.on) {
          const uint64_t idx = ((uint64_t)((int64_t)bh * d.Lq + irow)) * (uint64_t)Lk4 + (uint64_t)jb;
          const uint4 r = drop_rand4(dc, idx >> 2);
          keep[0] = r.x >= dc.thr ? dc.inv_keep : 0.f; keep[1] = r.y >= dc.thr ? dc.inv_keep : 0.f;
          keep[2] = r.z >= dc.thr ? dc.inv_keep : 0.f; keep[3] = r.w >= dc.thr ? dc.inv_keep : 0.f;
        }
        float ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = jb + e;
          const bool open = (i < d.Lq) && (j < d.Lk) && (j - i < 1 + off);
          const float p = open ? __expf(s[c4 + e] * d.scale - lse) : 0.f;
          ds[e] = to_tf32(p * (dp[c4 + e] * keep[e] - delta) * d.scale);
        }
        *reinterpret_cast<float4*>(dSs + half * (TQ * 128) + swz128(tid, c4 * 4)) = make_float4(ds[0], ds[1], ds[2], ds[3]);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < TK / 8; ++kk)
        a_mma_tf32(t_dq, desc_kmajor(a_smem_u32(dSs) + (kk >> 2) * (TQ * 128) + (kk & 3) * 32),
                   desc_mnmajor(a_smem_u32(Kmn) + kk * 1024), id_q, (tile | kk) != 0 ? 1u : 0u);
      a_commit(a_smem_u32(&bar_b));
    }
    a_mbar_wait(a_smem_u32(&bar_b), phase);       // operands free again; accumulator keeps growing in TMEM
    tc_fence_after();
    tc_fence_before();
  }
  {
    float dq[32];
    a_tmem_ld32(t_dq + lane_addr, dq);
    if (i < d.Lq) {
      float* qp = d.dq + ((int64_t)i * d.B + b) * d.lddq + h * d.hd;
#pragma unroll
      for (int c = 0; c < HP; ++c)
        if (c < d.hd) qp[c] = dq[c];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ATC_DQ_TMEM) : "memory");
  }
}

// ============================================================================ backward: dK, dV
// One CTA per (batch, head, 128-row KEY tile); thread t owns key row t.  Per 32-row query tile:
// S^T = K Q^T and dP^T = V dO^T on the tensor core, P~^T and dS^T on CUDA cores, then
// dV += P~^T dO and dK += dS^T Q on the tensor core (accumulators stay in TMEM).
constexpr int TI = 32;
constexpr int ATC_DKV_TMEM = 128;
constexpr int ATC_DKV_SMEM = 2 * TQ * 128 + 4 * TI * 128 + 2 * TQ * 128 + 2 * TI * 4 + 1024;

__global__ void __launch_bounds__(ATC_THREADS) attn_bwd_dkv_tc_kernel(const __grid_constant__ Group<mtb_attn_bwd_desc> g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a, bar_b;
  __shared__ uint32_t tmem_slot;
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_attn_bwd_desc& d = g.d[pi];
  const int ktiles = (d.Lk + TQ - 1) / TQ;
  const int bh = local / ktiles, kt = local - bh * ktiles;
  const int b = bh / d.H, h = bh - b * d.H;
  const int j0 = kt * TQ;
  const int off = abs(d.Lk - d.Lq);
  const int Lk4 = a_round4(d.Lk);
  const DropCtx dc = make_drop(d.rng, d.p);
  const int tid = threadIdx.x, warp = tid >> 5;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Ks = smem;                        // [128][128 B] K-major (A of S^T)
  uint8_t* Vs = Ks + TQ * 128;               // [128][128 B] K-major (A of dP^T)
  uint8_t* Qs = Vs + TQ * 128;               // [ 32][128 B] K-major (B of S^T)
  uint8_t* dOs = Qs + TI * 128;              // [ 32][128 B] K-major (B of dP^T)
  uint8_t* Qmn = dOs + TI * 128;             // [ 32][128 B] MN-major (B of dK)
  uint8_t* dOmn = Qmn + TI * 128;            // [ 32][128 B] MN-major (B of dV)
  uint8_t* PTs = dOmn + TI * 128;            // [128][128 B] K-major (A of dV), 32 query columns
  uint8_t* dSTs = PTs + TQ * 128;            // [128][128 B] K-major (A of dK)
  float* col_lse = reinterpret_cast<float*>(dSTs + TQ * 128);
  float* col_delta = col_lse + TI;

  if (tid == 0) {
    a_mbar_init(a_smem_u32(&bar_a), 1);
    a_mbar_init(a_smem_u32(&bar_b), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(&tmem_slot)), "r"(ATC_DKV_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  stage_tile<false>(Ks, d.k, d.ldk, d.B, b, h, d.hd, j0, d.Lk, TQ);
  stage_tile<false>(Vs, d.v, d.ldv, d.B, b, h, d.hd, j0, d.Lk, TQ);
  stage_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_st = tmem, t_dpt = tmem + 32, t_dv = tmem + 64, t_dk = tmem + 96;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  const int j = j0 + tid;                     // my key row
  const uint32_t id_s = idesc_tf32(TQ, TI, false, false);
  const uint32_t id_o = idesc_tf32(TQ, HP, false, true);
  const int i_first = max(0, j0 - off);
  uint32_t phase = 0;
  int tile = 0;
  for (int i0 = (i_first / TI) * TI; i0 < d.Lq; i0 += TI, phase ^= 1u, ++tile) {
    stage_tile<false>(Qs, d.q, d.ldq, d.B, b, h, d.hd, i0, d.Lq, TI);
    stage_tile<false>(dOs, d.d_o, d.lddo, d.B, b, h, d.hd, i0, d.Lq, TI);
    stage_tile<true>(Qmn, d.q, d.ldq, d.B, b, h, d.hd, i0, d.Lq, TI);
    stage_tile<true>(dOmn, d.d_o, d.lddo, d.B, b, h, d.hd, i0, d.Lq, TI);
    if (tid < TI) {
      const int ii = i0 + tid;
      col_lse[tid] = ii < d.Lq ? d.lse[(int64_t)bh * d.Lq + ii] : 0.f;
      col_delta[tid] = ii < d.Lq ? d.delta[(int64_t)bh * d.Lq + ii] : 0.f;
    }
    stage_wait();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HP / 8; ++k)
        a_mma_tf32(t_st, desc_kmajor(a_smem_u32(Ks) + k * 32), desc_kmajor(a_smem_u32(Qs) + k * 32), id_s, k != 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < HP / 8; ++k)
        a_mma_tf32(t_dpt, desc_kmajor(a_smem_u32(Vs) + k * 32), desc_kmajor(a_smem_u32(dOs) + k * 32), id_s, k != 0 ? 1u : 0u);
      a_commit(a_smem_u32(&bar_a));
    }
    a_mbar_wait(a_smem_u32(&bar_a), phase);
    tc_fence_after();
    {
      float st[32], dpt[32];
      a_tmem_ld32(t_st + lane_addr, st);
      a_tmem_ld32(t_dpt + lane_addr, dpt);
#pragma unroll
      for (int c4 = 0; c4 < TI; c4 += 4) {
        // Dropout keeps for 4 query columns x my key row.  The 4 lanes of a quad own key rows j..j+3 =
        // one Philox group per query row, so lane k of the quad draws the group of query row c4+k and
        // the quad exchanges components by shuffle: 1/4 Philox call + 4 shuffles per element group
        // instead of one call per element.
        float keep[4] = {dc.inv_keep, dc.inv_keep, dc.inv_keep, dc.inv_keep};
        if (dc.on) {
          const int lane = tid & 31;
          const int iq = min(i0 + c4 + (lane & 3), d.Lq - 1);
          const uint64_t idx = ((uint64_t)((int64_t)bh * d.Lq + iq)) * (uint64_t)Lk4 + (uint64_t)((j0 + tid) & ~3);
          const uint4 r = drop_rand4(dc, idx >> 2);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int srcl = (lane & ~3) + m;
            const uint32_t rx = __shfl_sync(0xffffffffu, r.x, srcl), ry = __shfl_sync(0xffffffffu, r.y, srcl);
            const uint32_t rz = __shfl_sync(0xffffffffu, r.z, srcl), rw = __shfl_sync(0xffffffffu, r.w, srcl);
            const int k = lane & 3;
            const uint32_t mine = k == 0 ? rx : k == 1 ? ry : k == 2 ? rz : rw;
            keep[m] = mine >= dc.thr ? dc.inv_keep : 0.f;
          }
        }
        float pt[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ii = i0 + c4 + e;
          const bool open = (ii < d.Lq) && (j < d.Lk) && (j - ii < 1 + off);
          const float p = open ? __expf(st[c4 + e] * d.scale - col_lse[c4 + e]) : 0.f;
          pt[e] = p * keep[e];
          ds[e] = p * (dpt[c4 + e] * keep[e] - col_delta[c4 + e]) * d.scale;
        }
        *reinterpret_cast<float4*>(PTs + swz128(tid, c4 * 4)) = make_float4(pt[0], pt[1], pt[2], pt[3]);
        *reinterpret_cast<float4*>(dSTs + swz128(tid, c4 * 4)) = make_float4(ds[0], ds[1], ds[2], ds[3]);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < TI / 8; ++kk)
        a_mma_tf32(t_dv, desc_kmajor(a_smem_u32(PTs) + kk * 32), desc_mnmajor(a_smem_u32(dOmn) + kk * 1024), id_o,
                   (tile | kk) != 0 ? 1u : 0u);
#pragma unroll
      for (int kk = 0; kk < TI / 8; ++kk)
        a_mma_tf32(t_dk, desc_kmajor(a_smem_u32(dSTs) + kk * 32), desc_mnmajor(a_smem_u32(Qmn) + kk * 1024), id_o,
                   (tile | kk) != 0 ? 1u : 0u);
      a_commit(a_smem_u32(&bar_b));
    }
    a_mbar_wait(a_smem_u32(&bar_b), phase);
    tc_fence_after();
    tc_fence_before();
  }
  {
    float dv[32], dk[32];
    if (tile > 0) {
      a_tmem_ld32(t_dv + lane_addr, dv);
      a_tmem_ld32(t_dk + lane_addr, dk);
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) { dv[c] = 0.f; dk[c] = 0.f; }
    }
    if (j < d.Lk) {
      float* kp = d.dk + ((int64_t)j * d.B + b) * d.lddk + h * d.hd;
      float* vp = d.dv + ((int64_t)j * d.B + b) * d.lddv + h * d.hd;
#pragma unroll
      for (int c = 0; c < HP; ++c)
        if (c < d.hd) { kp[c] = dk[c]; vp[c] = dv[c]; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ATC_DKV_TMEM) : "memory");
  }
}

constexpr int ATC_FWD_SMEM = TQ * 128 + 2 * TK * 128 + (TK / 32) * TQ * 128 + 1024;

int attn_fwd_simt(const mtb_attn_desc* d, int n, cudaStream_t st);

int attn_fwd_tc(const mtb_attn_desc* d, int n, cudaStream_t st) {
  mtb_attn_desc rest[MTB_MAX_GROUP];
  Group<mtb_attn_desc> g;
  int ntc = 0, nrest = 0, tot = 0;
  for (int i = 0; i < n; ++i) {
    if (d[i].hd > HP) { rest[nrest++] = d[i]; continue; }
    g.d[ntc] = d[i];
    g.start[ntc] = tot;
    tot += d[i].B * d[i].H * ((d[i].Lq + TQ - 1) / TQ);
    ++ntc;
  }
  g.n = ntc;
  g.start[ntc] = tot;
  if (tot > 0) {
    static bool attr = false;
    if (!attr) {
      MTB_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_FWD_SMEM));
      attr = true;
    }
    attn_fwd_tc_kernel<<<tot, ATC_THREADS, ATC_FWD_SMEM, st>>>(g);
    mtb::note_launch();
    MTB_CUDA(cudaGetLastError());
  }
  if (nrest) return attn_fwd_simt(rest, nrest, st);
  return 0;
}

int attn_bwd_simt(const mtb_attn_bwd_desc* d, int n, cudaStream_t st);

int attn_bwd_tc(const mtb_attn_bwd_desc* d, int n, cudaStream_t st) {
  mtb_attn_bwd_desc rest[MTB_MAX_GROUP];
  Group<mtb_attn_bwd_desc> gq, gk;
  int ntc = 0, nrest = 0, totq = 0, totk = 0;
  for (int i = 0; i < n; ++i) {
    if (d[i].hd > HP) { rest[nrest++] = d[i]; continue; }
    gq.d[ntc] = d[i]; gk.d[ntc] = d[i];
    gq.start[ntc] = totq; gk.start[ntc] = totk;
    totq += d[i].B * d[i].H * ((d[i].Lq + TQ - 1) / TQ);
    totk += d[i].B * d[i].H * ((d[i].Lk + TQ - 1) / TQ);
    ++ntc;
  }
  gq.n = gk.n = ntc;
  gq.start[ntc] = totq; gk.start[ntc] = totk;
  if (totq > 0 && totk > 0) {
    static bool attr = false;
    if (!attr) {
      MTB_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_DQ_SMEM));
      MTB_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_DKV_SMEM));
      attr = true;
    }
    attn_bwd_dq_tc_kernel<<<totq, ATC_THREADS, ATC_DQ_SMEM, st>>>(gq);
    mtb::note_launch();
    MTB_CUDA(cudaGetLastError());
    attn_bwd_dkv_tc_kernel<<<totk, ATC_THREADS, ATC_DKV_SMEM, st>>>(gk);
    mtb::note_launch();
    MTB_CUDA(cudaGetLastError());
  }
  if (nrest) return attn_bwd_simt(rest, nrest, st);
  return 0;
}

}  // namespace mtb

namespace mtb {
int preload_attention_tc() {
  int bad = 0;
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_fwd_tc_kernel) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_bwd_dq_tc_kernel) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_bwd_dkv_tc_kernel) != cudaSuccess) ++bad; }
  return bad;
}
}  // namespace mtb
