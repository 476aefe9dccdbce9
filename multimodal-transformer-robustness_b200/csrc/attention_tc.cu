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
#include <stdlib.h>

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"     // %3: suspend-time hint -- the warp sleeps in
      "selp.u32 %0, 1, 0, p;\n\t}"                                        // hardware instead of spinning on issue slots
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
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
  // round-to-nearest (ties away) at the 10-bit TF32 mantissa: add half an ulp (bit 12) to the magnitude; the tensor core
  // then truncates bits [0, 13).  Same result as cvt.rna.tf32.f32 for every finite input (all values rounded here are
  // finite: probabilities and their products); one integer add instead of a compare + predicated add.
  return __uint_as_float(__float_as_uint(x) + 0x1000u);
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
template <bool MNSW, int NT = ATC_THREADS>
__device__ __forceinline__ void stage_tile(uint8_t* tile, const float* base, int64_t ld, int B, int b, int h, int hd, int l0, int L,
                                           int rows) {
  const uint32_t tbase = a_smem_u32(tile);
  for (int e = threadIdx.x; e < rows * HP; e += NT) {
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

// 256-thread staging with incremental addressing: thread t always moves head-dim column t & 31 of rows
// (t >> 5) + 8 k, so the column test, the swizzle term (row & 7 / row & 3 are invariant under + 8) and the
// source offset are computed once; each copy is then a compare, a select and a cp.async.
template <bool MNSW>
__device__ __forceinline__ void stage_rows256(uint8_t* tile, const float* base, int64_t ld, int B, int b, int h, int hd, int l0, int L,
                                              int rows, bool bf = false) {
  const int c = threadIdx.x & 31, r0 = threadIdx.x >> 5;
  const bool ok_c = c < hd;
  uint32_t dst = a_smem_u32(tile) + (MNSW ? swz128_32(r0, c * 4) : swz128(r0, c * 4));
  if (!bf) {
    const float* src = base + ((int64_t)(l0 + r0) * B + b) * ld + h * hd + c;
    const int64_t sstep = (int64_t)8 * B * ld;
#pragma unroll 4
    for (int r = r0; r < rows; r += 8) {
      const bool ok = ok_c && (l0 + r < L);
      const float* sp = ok ? src : base;
      const int nbytes = ok ? 4 : 0;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(sp), "r"(nbytes) : "memory");
      src += sstep;
      dst += 8 * 128;
    }
  } else {
    // bf16 activations (bf16 data path): rows are 2 * hd bytes at 2-byte aligned offsets, below cp.async's 4-byte
    // granularity.  Each element's slot receives the 4-byte-aligned WORD that contains the element (asynchronously, like
    // the fp32 path); fix_rows256() -- run by the same thread once its copies have landed -- keeps the right half and
    // widens it to fp32 in place (exact: a 16-bit shift; a bf16 value is also a valid tf32 operand).  The neighbouring
    // element read along is always inside the tensor: leading dimensions and element counts are even.
    const uint16_t* src = reinterpret_cast<const uint16_t*>(base) + ((int64_t)(l0 + r0) * B + b) * ld + h * hd + c;
    const int64_t sstep = (int64_t)8 * B * ld;
#pragma unroll 4
    for (int r = r0; r < rows; r += 8) {
      const bool ok = ok_c && (l0 + r < L);
      const void* sp = ok ? reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3) : reinterpret_cast<const void*>(base);
      const int nbytes = ok ? 4 : 0;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(sp), "r"(nbytes) : "memory");
      src += sstep;
      dst += 8 * 128;
    }
  }
}
// second half of the bf16 staging: every thread converts the slots IT copied (cp.async.wait_group makes a thread's own
// copies visible to it), so no barrier is needed between the wait and this pass.  Zero-filled slots stay zero.
template <bool MNSW>
__device__ __forceinline__ void fix_rows256(uint8_t* tile, const float* base, int64_t ld, int B, int b, int h, int hd, int l0, int rows) {
  const int c = threadIdx.x & 31, r0 = threadIdx.x >> 5;
  // parity of the element's 2-byte index in memory: rows advance by 8 * B * ld elements (even), so it is per-thread constant
  const uint64_t e0 = (uint64_t)(reinterpret_cast<uintptr_t>(base) >> 1) + (uint64_t)(((int64_t)(l0 + r0) * B + b) * ld + h * hd + c);
  const bool hi = (e0 & 1) != 0;
  uint32_t dst = a_smem_u32(tile) + (MNSW ? swz128_32(r0, c * 4) : swz128(r0, c * 4));
#pragma unroll 4
  for (int r = r0; r < rows; r += 8) {
    uint32_t w;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(dst) : "memory");
    w = hi ? (w & 0xffff0000u) : (w << 16);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(w) : "memory");
    dst += 8 * 128;
  }
}
__device__ __forceinline__ void a_tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Store vals[0 .. n) (n <= N) as one head's row at ELEMENT offset `off` of a tensor that is fp32 or bf16.  bf16 rows start at
// 2-byte aligned offsets: elements are paired into 4-byte words (one cvt.rn.bf16x2 + one store per pair) around an optional
// leading / trailing single element; the parity is uniform across a CTA (leading dimensions are even, so it is h * hd & 1).
template <int N>
__device__ __forceinline__ void store_row(void* base, int64_t off, const float (&vals)[N], int n, bool bf16) {
  if (!bf16) {
    float* p = reinterpret_cast<float*>(base) + off;
#pragma unroll
    for (int c = 0; c < N; ++c)
      if (c < n) p[c] = vals[c];
    return;
  }
  uint16_t* p = reinterpret_cast<uint16_t*>(base) + off;
  if ((reinterpret_cast<uintptr_t>(p) & 2) == 0) {
#pragma unroll
    for (int c = 0; c + 1 < N; c += 2) {
      if (c + 1 < n) *reinterpret_cast<uint32_t*>(p + c) = pack_bf16x2(vals[c], vals[c + 1]);
      else if (c < n) p[c] = (uint16_t)(pack_bf16x2(vals[c], 0.f) & 0xffffu);
    }
    if ((N & 1) && N - 1 < n) p[N - 1] = (uint16_t)(pack_bf16x2(vals[N - 1], 0.f) & 0xffffu);
  } else {
    if (0 < n) p[0] = (uint16_t)(pack_bf16x2(vals[0], 0.f) & 0xffffu);
#pragma unroll
    for (int c = 1; c + 1 < N; c += 2) {
      if (c + 1 < n) *reinterpret_cast<uint32_t*>(p + c) = pack_bf16x2(vals[c], vals[c + 1]);
      else if (c < n) p[c] = (uint16_t)(pack_bf16x2(vals[c], 0.f) & 0xffffu);
    }
    if (!(N & 1) && N - 1 < n) p[N - 1] = (uint16_t)(pack_bf16x2(vals[N - 1], 0.f) & 0xffffu);
  }
}

#ifdef MTB_TC_TRACE
__device__ unsigned long long g_attn_trace[128];
__device__ __forceinline__ unsigned long long a_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t) :: "memory"); return t; }
#define ATRACE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (i) < 128) g_attn_trace[(i)] = a_gtime(); } while (0)
extern "C" int mtb_debug_attn_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(unsigned long long) * 128);
}
#define QTRACE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (i) < 128) g_attn_trace[(i)] = a_gtime(); } while (0)
#else
#define ATRACE(i) do { } while (0)
#define QTRACE(i) do { } while (0)
#endif

// ---------------------------------------------------------------------------- forward
// 256 threads: TWO threads per query row.  Warp w serves TMEM lane quadrant w & 3 (rows) and key-column half
// w >> 2 of every 64-key tile; the two half-row threads run INDEPENDENT online softmaxes (own running max /
// sum, own PV accumulator fed by its own K = 32 MMA group) and are merged once after the last tile, so no
// per-tile exchange is needed.  The loop is software pipelined: K/V of tile t+1 are fetched (cp.async) and
// S_{t+1} = Q K_{t+1}^T is issued while tile t is in its softmax / PV phase; one __syncthreads per tile.
constexpr int AF_THREADS = 256;
constexpr int ATC_FWD_SMEM = TQ * 128 + 4 * TK * 128 + (TK / 32) * TQ * 128 + 1024;

__device__ __forceinline__ float fast_exp2(float x) {          // one MUFU.EX2; exp2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t row_keep_bits32(const DropCtx& dc, uint64_t idx0) {   // keep bits of 32 consecutive columns
  uint32_t bits = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 r = drop_rand4(dc, (idx0 >> 2) + (uint64_t)q);
    bits |= ((r.x >= dc.thr ? 1u : 0u) | (r.y >= dc.thr ? 2u : 0u) | (r.z >= dc.thr ? 4u : 0u) | (r.w >= dc.thr ? 8u : 0u)) << (4 * q);
  }
  return bits;
}

template <bool MASKED>
__device__ __forceinline__ void fwd_softmax_half(float (&s)[32], const float c2, const int i, const int jb, const int Lk, const int off,
                                                 float& m_run, float& l_run, float& corr, const bool drop_on, const float inv_keep,
                                                 const uint32_t kbits, uint8_t* prow_region, const int row) {
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    if (MASKED) {
      const int j = jb + c;
      const bool open = (j < Lk) && (j - i < 1 + off);
      s[c] = open ? s[c] * c2 : -CUDART_INF_F;
    } else {
      s[c] *= c2;
    }
    mx = fmaxf(mx, s[c]);
  }
  const float m_new = fmaxf(m_run, mx);
  const bool dead = (m_new == -CUDART_INF_F);                 // no open column seen so far in my half
  corr = dead ? 1.f : fast_exp2(m_run - m_new);
  const float m_use = dead ? 0.f : m_new;
  float rs = 0.f;
#pragma unroll
  for (int c4 = 0; c4 < 32; c4 += 4) {
    float pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      pk[e] = fast_exp2(s[c4 + e] - m_use);                      // exp2(-inf) = 0 for masked entries
      rs += pk[e];
    }
    if (drop_on) {
#pragma unroll
      for (int e = 0; e < 4; ++e) pk[e] = ((kbits >> (c4 + e)) & 1u) ? pk[e] * inv_keep : 0.f;
    }
    *reinterpret_cast<float4*>(prow_region + swz128(row, c4 * 4)) =
        make_float4(to_tf32(pk[0]), to_tf32(pk[1]), to_tf32(pk[2]), to_tf32(pk[3]));
  }
  l_run = l_run * corr + rs;
  m_run = m_new;
}

// Output rows leave through shared memory: a thread owns one accumulator ROW (TMEM lane), so direct stores are hd scalar
// stores per thread that each touch 32 scattered sectors (~25 x 32 L1 wavefronts per warp).  Instead every thread parks its
// values in an (idle) [128][32] fp32 tile -- element (r, c) at r * 32 + (c ^ (r & 31)): conflict-free for thread-per-row
// writes and for warp-per-row reads -- and after a barrier each warp writes whole rows, one lane per column (one or two
// sectors per row).  Values and rounding are those of store_row().
template <int N>
__device__ __forceinline__ void park_row(float* tile, int r, int c0, const float (&vals)[N]) {
#pragma unroll
  for (int c = 0; c < N; ++c) tile[r * 32 + ((c0 + c) ^ (r & 31))] = vals[c];
}
template <int NT>
__device__ __forceinline__ void rows_out(const float* tile, void* base, int64_t ld, int B, int b, int h, int hd, int l0, int L, int rows,
                                         bool bf16) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane >= hd) return;
  const int n = min(rows, L - l0);
  int64_t off = ((int64_t)(l0 + w) * B + b) * ld + h * hd + lane;
  const int64_t step = (int64_t)(NT / 32) * B * ld;
  if (bf16) {
    uint16_t* p = reinterpret_cast<uint16_t*>(base);
#pragma unroll 4
    for (int r = w; r < n; r += NT / 32, off += step) p[off] = (uint16_t)(pack_bf16x2(tile[r * 32 + (lane ^ (r & 31))], 0.f) & 0xffffu);
  } else {
    float* p = reinterpret_cast<float*>(base);
#pragma unroll 4
    for (int r = w; r < n; r += NT / 32, off += step) p[off] = tile[r * 32 + (lane ^ (r & 31))];
  }
}

__global__ void __launch_bounds__(AF_THREADS, 2) attn_fwd_tc_kernel(const __grid_constant__ Group<mtb_attn_desc> g) {
  pdl_begin();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_attn_desc& d = g.d[pi];
  const int qtiles = (d.Lq + TQ - 1) / TQ;
  const int BH = d.B * d.H;
  // heaviest query tiles (most key tiles under the causal-offset mask) get the lowest block indices
  const int bh = local % BH, qt = qtiles - 1 - local / BH;
  const int b = bh / d.H, h = bh - b * d.H;
  const int i0 = qt * TQ;
  const int Lq = d.Lq, Lk = d.Lk, hd = d.hd;
  const int off = abs(Lk - Lq);
  const int Lk4 = a_round4(Lk);
  DropCtx dc;                                       // filled after pdl_ready(): the device counter lives in global memory
  const bool bf = (d.bf16 & 1) != 0;                // q / k / v are bfloat16 in HBM
  const bool bf_o = (d.bf16 & 2) != 0;              // o is bfloat16
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;                 // query row inside the tile = TMEM lane

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Qs = smem;                       // [128][128 B]  K-major
  uint8_t* KV = Qs + TQ * 128;              // 2 buffers x { K [64][128 B] K-major, V [64][128 B] MN-major (32 B-atom swizzle) }
  uint8_t* Ps = KV + 4 * TK * 128;          // 2 regions of [128][128 B], K-major, 32 key columns each

  if (tid == 0) {
    a_mbar_init(a_smem_u32(&bar_s), 1);
    a_mbar_init(a_smem_u32(&bar_o), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(&tmem_slot)), "r"(ATC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_ready();                 // barrier init and TMEM allocation ran under the tail of the preceding kernel
  dc = make_drop(d.rng, d.p);
  const int i_last = min(Lq, i0 + TQ) - 1;
  const int j_end = min(Lk, i_last + off + 1);
  const int T = (j_end + TK - 1) / TK;
  stage_rows256<false>(Qs, d.q, d.ldq, d.B, b, h, hd, i0, Lq, TQ, bf);
  stage_rows256<false>(KV, d.k, d.ldk, d.B, b, h, hd, 0, Lk, TK, bf);
  stage_rows256<true>(KV + TK * 128, d.v, d.ldv, d.B, b, h, hd, 0, Lk, TK, bf);
  stage_wait();
  if (bf) {
    fix_rows256<false>(Qs, d.q, d.ldq, d.B, b, h, hd, i0, TQ);
    fix_rows256<false>(KV, d.k, d.ldk, d.B, b, h, hd, 0, TK);
    fix_rows256<true>(KV + TK * 128, d.v, d.ldv, d.B, b, h, hd, 0, TK);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_s = tmem, t_o = tmem + 64;        // S: 64 columns; PV accumulators of the two halves: 32 + 32
  const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
  const uint32_t id_s = idesc_tf32(TQ, TK, false, false);
  const uint32_t id_o = idesc_tf32(TQ, HP, false, true);
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < HP / 8; ++k)
      a_mma_tf32(t_s, desc_kmajor(a_smem_u32(Qs) + k * 32), desc_kmajor(a_smem_u32(KV) + k * 32), id_s, k != 0 ? 1u : 0u);
    a_commit(a_smem_u32(&bar_s));
  }

  const int i = i0 + row;                    // my query row
  const int irow = min(i, Lq - 1);
  const float c2 = d.scale * 1.4426950408889634f;   // softmax in the exp2 domain
  float o[HP];
#pragma unroll
  for (int c = 0; c < HP; ++c) o[c] = 0.f;
  float m_run = -CUDART_INF_F, l_run = 0.f;
  uint8_t* my_p = Ps + half * (TQ * 128);
  const uint64_t idx_row = ((uint64_t)((int64_t)bh * Lq + irow)) * (uint64_t)Lk4 + (uint64_t)(half * 32);
  // dropout keep bits of my 32 columns of the current tile: drawn one tile ahead, while the MMAs run; optionally stored
  // for the backward kernels (one word per thread and tile)
  uint32_t kbits = dc.on ? row_keep_bits32(dc, idx_row) : 0u;
  const int KW = (Lk + 31) >> 5;
  uint32_t* wrow = (d.keep_bits != nullptr && dc.on && i < Lq) ? d.keep_bits + ((int64_t)bh * Lq + i) * KW : nullptr;

  ATRACE(0);
  for (int t = 0; t < T; ++t) {
    const int j0 = t * TK;
    const uint32_t ph = (uint32_t)t & 1u;
    uint8_t* cur = KV + (t & 1) * (2 * TK * 128);
    uint8_t* nxt = KV + ((t + 1) & 1) * (2 * TK * 128);
    ATRACE(8 + t * 8 + 0);
    if (t + 1 < T) {                         // prefetch the next key / value tile behind this tile's softmax
      stage_rows256<false>(nxt, d.k, d.ldk, d.B, b, h, hd, j0 + TK, Lk, TK, bf);
      stage_rows256<true>(nxt + TK * 128, d.v, d.ldv, d.B, b, h, hd, j0 + TK, Lk, TK, bf);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    ATRACE(8 + t * 8 + 1);
    a_mbar_wait(a_smem_u32(&bar_s), ph);
    tc_fence_after();
    ATRACE(8 + t * 8 + 2);
    float s[32];
    a_tmem_ld32(t_s + lane_addr + half * 32, s);
    tc_fence_before();
    float corr;
    const bool tile_open = (j0 + TK <= Lk) && (j0 + TK - 1 - i0 < 1 + off);     // CTA-uniform: no entry of this tile is masked
    if (tile_open)
      fwd_softmax_half<false>(s, c2, i, j0 + half * 32, Lk, off, m_run, l_run, corr, dc.on, dc.inv_keep, kbits, my_p, row);
    else
      fwd_softmax_half<true>(s, c2, i, j0 + half * 32, Lk, off, m_run, l_run, corr, dc.on, dc.inv_keep, kbits, my_p, row);
    if (wrow != nullptr && 2 * t + half < KW) wrow[2 * t + half] = kbits;
    ATRACE(8 + t * 8 + 3);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (bf && t + 1 < T) {
      fix_rows256<false>(nxt, d.k, d.ldk, d.B, b, h, hd, j0 + TK, TK);
      fix_rows256<true>(nxt + TK * 128, d.v, d.ldv, d.B, b, h, hd, j0 + TK, TK);
    }
    fence_async_smem();
    __syncthreads();                          // P written, S read by everyone, next K / V landed
    ATRACE(8 + t * 8 + 4);
    // Issuing one of these small tcgen05.mma costs ~0.1 us of a single thread: the two independent MMA groups of a tile
    // are issued by two threads of different warps in parallel (each commits to its own barrier).
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < TK / 8; ++kk)     // half hf = kk >> 2 accumulates P[:, 32 hf : 32 hf + 32] V[32 hf : 32 hf + 32, :] on its own
        a_mma_tf32(t_o + (kk >> 2) * 32, desc_kmajor(a_smem_u32(Ps) + (kk >> 2) * (TQ * 128) + (kk & 3) * 32),
                   desc_mnmajor(a_smem_u32(cur + TK * 128) + kk * 1024), id_o, (kk & 3) != 0 ? 1u : 0u);
      a_commit(a_smem_u32(&bar_o));
    } else if (tid == 128 && t + 1 < T) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HP / 8; ++k)
        a_mma_tf32(t_s, desc_kmajor(a_smem_u32(Qs) + k * 32), desc_kmajor(a_smem_u32(nxt) + k * 32), id_s, k != 0 ? 1u : 0u);
      a_commit(a_smem_u32(&bar_s));
    }
    if (dc.on && t + 1 < T) kbits = row_keep_bits32(dc, idx_row + (uint64_t)(j0 + TK));     // next tile's bits, behind the PV MMAs
    a_mbar_wait(a_smem_u32(&bar_o), ph);
    tc_fence_after();
    ATRACE(8 + t * 8 + 5);
    float pv[32];
    a_tmem_ld32(t_o + lane_addr + half * 32, pv);
#pragma unroll
    for (int c = 0; c < HP; ++c) o[c] = o[c] * corr + pv[c];
    tc_fence_before();          // my TMEM reads are done before the next tile's MMAs overwrite S / PV
    ATRACE(8 + t * 8 + 6);
  }
  ATRACE(1);
  // ---- merge the two half-row states: half 1 hands (m, l, o) to half 0 through the (idle) P region ----------
  float* ex = reinterpret_cast<float*>(Ps);            // [128][36]
  if (half == 1) {
    float* e = ex + row * 36;
    e[0] = m_run; e[1] = l_run;
#pragma unroll
    for (int c = 0; c < HP; c += 4) *reinterpret_cast<float4*>(e + 4 + c) = make_float4(o[c], o[c + 1], o[c + 2], o[c + 3]);
  }
  __syncthreads();
  if (half == 0 && i < Lq) {
    const float* e = ex + row * 36;
    const float m_b = e[0], l_b = e[1];
    const float m = fmaxf(m_run, m_b);                 // finite: key 0 is open for every row and belongs to half 0
    const float wa = fast_exp2(m_run - m), wb = fast_exp2(m_b - m);
    const float l = l_run * wa + l_b * wb;
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < HP; ++c) o[c] = (o[c] * wa + e[4 + c] * wb) * inv;
    park_row<HP>(reinterpret_cast<float*>(Qs), row, 0, o);          // every MMA that read Q has retired
    if (d.lse) d.lse[(int64_t)bh * Lq + i] = m * 0.6931471805599453f + logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ATC_TMEM_COLS) : "memory");
  }
  rows_out<AF_THREADS>(reinterpret_cast<const float*>(Qs), d.o, d.ldo, d.B, b, h, hd, i0, Lq, TQ, bf_o);
}


// ---------------------------------------------------------------------------- forward, short sequences (Lq, Lk <= 64)
// At L <= 64 a (batch, head) problem fills at most half of the 128 TMEM lanes and a single key tile, and a launch is
// thousands of one-tile CTAs whose cost is their fixed part (TMEM allocation, barrier setup, staging latency, drain):
// EA fitness at 2048 samples x 8 heads and L = 50 spent 36 % of its kernel time there.  This variant packs TWO
// (batch, head) problems into one CTA: rows 0-63 are the queries of problem A, rows 64-127 those of problem B; the
// key / value tiles hold A's keys in rows 0-63 and B's in rows 64-127.  S_A = Q K_A^T and S_B = Q K_B^T are both formed
// for all 128 rows (the cross blocks are never read), every thread reads the block of its own problem, and
// PV runs once per (problem, 32-key half) into four accumulators of which a thread reads its own.  Twice the MMAs of
// the general kernel per useful score -- the tensor pipe idles at this size anyway -- for half the CTAs.
constexpr int SQ = 64;
constexpr int ATC_SHORT_TMEM = 256;          // S_A, S_B: 2 x 64 columns; PV: 4 x 32 columns
constexpr int ATC_SHORT_SMEM = 3 * TQ * 128 + (TK / 32) * TQ * 128 + 1024;

__global__ void __launch_bounds__(AF_THREADS, 2) attn_fwd_tc_short_kernel(const __grid_constant__ Group<mtb_attn_desc> g) {
  pdl_begin();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_attn_desc& d = g.d[pi];
  const int BH = d.B * d.H;
  const int Lq = d.Lq, Lk = d.Lk, hd = d.hd;
  const int off = abs(Lk - Lq);
  const int Lk4 = a_round4(Lk);
  DropCtx dc;                                       // filled after pdl_ready(): the device counter lives in global memory
  const bool bf = (d.bf16 & 1) != 0, bf_o = (d.bf16 & 2) != 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;                  // TMEM lane
  const int sub = row >> 6;                          // which of the two packed problems this row belongs to
  const int i = row & (SQ - 1);                      // query index inside that problem
  const int bh_a = 2 * local, bh_b = 2 * local + 1;  // bh_b may be one past the end (odd B * H): staged as zeros
  const int bh = sub ? bh_b : bh_a;
  const bool pvalid = bh < BH;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Qs = smem;                       // [128][128 B] K-major: rows 0-63 problem A, 64-127 problem B
  uint8_t* Ks = Qs + TQ * 128;              // [128][128 B] K-major: keys of A, keys of B
  uint8_t* Vs = Ks + TQ * 128;              // [128][128 B] MN-major (32 B-atom swizzle): values of A, values of B
  uint8_t* Ps = Vs + TQ * 128;              // 2 regions of [128][128 B], K-major, 32 key columns each

  if (tid == 0) {
    a_mbar_init(a_smem_u32(&bar_s), 1);
    a_mbar_init(a_smem_u32(&bar_o), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(&tmem_slot)), "r"(ATC_SHORT_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_ready();                 // barrier init and TMEM allocation ran under the tail of the preceding kernel
  dc = make_drop(d.rng, d.p);
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    const int bh2 = 2 * local + s2;
    const bool ok2 = bh2 < BH;
    const int b2 = ok2 ? bh2 / d.H : 0, h2 = ok2 ? bh2 - b2 * d.H : 0;
    const int lq2 = ok2 ? Lq : 0, lk2 = ok2 ? Lk : 0;        // a missing second problem is staged as zeros
    stage_rows256<false>(Qs + s2 * SQ * 128, d.q, d.ldq, d.B, b2, h2, hd, 0, lq2, SQ, bf);
    stage_rows256<false>(Ks + s2 * SQ * 128, d.k, d.ldk, d.B, b2, h2, hd, 0, lk2, SQ, bf);
    stage_rows256<true>(Vs + s2 * SQ * 128, d.v, d.ldv, d.B, b2, h2, hd, 0, lk2, SQ, bf);
  }
  stage_wait();
  if (bf) {
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      const int bh2 = 2 * local + s2;
      const bool ok2 = bh2 < BH;
      const int b2 = ok2 ? bh2 / d.H : 0, h2 = ok2 ? bh2 - b2 * d.H : 0;
      fix_rows256<false>(Qs + s2 * SQ * 128, d.q, d.ldq, d.B, b2, h2, hd, 0, SQ);
      fix_rows256<false>(Ks + s2 * SQ * 128, d.k, d.ldk, d.B, b2, h2, hd, 0, SQ);
      fix_rows256<true>(Vs + s2 * SQ * 128, d.v, d.ldv, d.B, b2, h2, hd, 0, SQ);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_s = tmem, t_o = tmem + 128;
  const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
  const uint32_t id_s = idesc_tf32(TQ, SQ, false, false);
  const uint32_t id_o = idesc_tf32(TQ, HP, false, true);
  if (tid == 0) {
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int k = 0; k < HP / 8; ++k)
        a_mma_tf32(t_s + s2 * SQ, desc_kmajor(a_smem_u32(Qs) + k * 32), desc_kmajor(a_smem_u32(Ks) + s2 * SQ * 128 + k * 32), id_s, k != 0 ? 1u : 0u);
    a_commit(a_smem_u32(&bar_s));
  }
  const int irow = min(i, Lq - 1);
  const float c2 = d.scale * 1.4426950408889634f;
  float m_run = -CUDART_INF_F, l_run = 0.f;
  uint8_t* my_p = Ps + half * (TQ * 128);
  const int bhc = pvalid ? bh : 0;
  const uint64_t idx_row = ((uint64_t)((int64_t)bhc * Lq + irow)) * (uint64_t)Lk4 + (uint64_t)(half * 32);
  const uint32_t kbits = dc.on ? row_keep_bits32(dc, idx_row) : 0u;        // drawn while the S MMAs run
  const int KW = (Lk + 31) >> 5;
  a_mbar_wait(a_smem_u32(&bar_s), 0);
  tc_fence_after();
  float s[32];
  a_tmem_ld32(t_s + lane_addr + sub * SQ + half * 32, s);
  tc_fence_before();
  float corr;
  fwd_softmax_half<true>(s, c2, i, half * 32, Lk, off, m_run, l_run, corr, dc.on, dc.inv_keep, kbits, my_p, row);
  if (d.keep_bits != nullptr && dc.on && pvalid && i < Lq && half < KW) d.keep_bits[((int64_t)bh * Lq + i) * KW + half] = kbits;
  fence_async_smem();
  __syncthreads();                            // P written, S read by everyone
  if (tid == 0) {
    tc_fence_after();
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)        // accumulator (s2, hf) = P[:, 32 hf : 32 hf + 32] . V_{s2}[32 hf : 32 hf + 32, :]
          a_mma_tf32(t_o + (s2 * 2 + hf) * 32, desc_kmajor(a_smem_u32(Ps) + hf * (TQ * 128) + k4 * 32),
                     desc_mnmajor(a_smem_u32(Vs) + s2 * SQ * 128 + (hf * 4 + k4) * 1024), id_o, k4 != 0 ? 1u : 0u);
    a_commit(a_smem_u32(&bar_o));
  }
  a_mbar_wait(a_smem_u32(&bar_o), 0);
  tc_fence_after();
  float o[HP];
  a_tmem_ld32(t_o + lane_addr + (sub * 2 + half) * 32, o);
  tc_fence_before();
  // ---- merge the two half-row states (see the general kernel) ----------------------------------------------------------
  __syncthreads();                            // the PV MMAs have read P: its region is free for the exchange
  float* ex = reinterpret_cast<float*>(Ps);   // [128][36]
  if (half == 1) {
    float* e = ex + row * 36;
    e[0] = m_run; e[1] = l_run;
#pragma unroll
    for (int c = 0; c < HP; c += 4) *reinterpret_cast<float4*>(e + 4 + c) = make_float4(o[c], o[c + 1], o[c + 2], o[c + 3]);
  }
  __syncthreads();
  if (half == 0 && pvalid && i < Lq) {
    const float* e = ex + row * 36;
    const float m_b = e[0], l_b = e[1];
    const float m = fmaxf(m_run, m_b);        // finite: key 0 is open for every row and belongs to half 0
    const float wa = fast_exp2(m_run - m), wb = fast_exp2(m_b - m);
    const float l = l_run * wa + l_b * wb;
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < HP; ++c) o[c] = (o[c] * wa + e[4 + c] * wb) * inv;
    park_row<HP>(reinterpret_cast<float*>(Qs), row, 0, o);
    if (d.lse) d.lse[(int64_t)bh * Lq + i] = m * 0.6931471805599453f + logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ATC_SHORT_TMEM) : "memory");
  }
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {                  // rows 0-63: problem A, rows 64-127: problem B
    const int bh2 = 2 * local + s2;
    if (bh2 < BH)
      rows_out<AF_THREADS>(reinterpret_cast<const float*>(Qs) + s2 * SQ * 32, d.o, d.ldo, d.B, bh2 / d.H, bh2 % d.H, hd, 0, Lq, SQ, bf_o);
  }
}


// ============================================================================ backward: dQ (+ delta)
// One CTA (256 threads) per (batch, head, 128-row query tile); two threads per query row (key-column halves of
// every 64-key tile).  Per key tile: S = Q K^T and dP = dO V^T on the tensor core,
// dS = P * (dP * keep - delta) * scale on CUDA cores, dQ += dS K on the tensor core (accumulating in TMEM).
// Software pipelined like the forward kernel: K / V of tile t+1 are fetched and S_{t+1} / dP_{t+1} issued while
// tile t is processed, the Philox keep bits of tile t+1 are drawn while the MMAs run; one __syncthreads per tile.
constexpr int ATC_DQ_TMEM = 256;
constexpr int ATC_DQ_SMEM = 2 * TQ * 128 + 4 * TK * 128 + (TK / 32) * TQ * 128 + 1024;
constexpr int AQ_THREADS = 256;

template <bool MASKED>
__device__ __forceinline__ void dq_math(const float (&s)[32], const float (&dp)[32], float lse2, float delta, uint32_t kbits, bool drop_on,
                                        float inv_keep, float c2, float scale, int i, int jb, int Lk, int off, uint8_t* ds_region, int row) {
#pragma unroll
  for (int c4 = 0; c4 < 32; c4 += 4) {
    float ds[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c4 + e;
      float p = fast_exp2(fmaf(s[c], c2, -lse2));
      if (MASKED) {
        const int j = jb + c;
        const bool open = (j < Lk) && (j - i < 1 + off);
        p = open ? p : 0.f;
      }
      const float keep = (!drop_on || ((kbits >> c) & 1u)) ? inv_keep : 0.f;
      ds[e] = p * (dp[c] * keep - delta) * scale;
    }
    *reinterpret_cast<float4*>(ds_region + swz128(row, c4 * 4)) = make_float4(to_tf32(ds[0]), to_tf32(ds[1]), to_tf32(ds[2]), to_tf32(ds[3]));
  }
}

__global__ void __launch_bounds__(AQ_THREADS, 2) attn_bwd_dq_tc_kernel(const __grid_constant__ Group<mtb_attn_bwd_desc> g) {
  pdl_begin();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a, bar_b;
  __shared__ uint32_t tmem_slot;
  __shared__ float delta_s[TQ];
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_attn_bwd_desc& d = g.d[pi];
  const int qtiles = (d.Lq + TQ - 1) / TQ;
  const int BH = d.B * d.H;
  const int bh = local % BH, qt = qtiles - 1 - local / BH;      // heaviest query tiles first
  const int b = bh / d.H, h = bh - b * d.H;
  const int i0 = qt * TQ;
  const int Lq = d.Lq, Lk = d.Lk, hd = d.hd;
  const int off = abs(Lk - Lq);
  const int Lk4 = a_round4(Lk);
  DropCtx dc;                                       // filled after pdl_ready(): the device counter lives in global memory
  const bool bf = (d.bf16 & 1) != 0;                // q / k / v are bfloat16 in HBM
  const bool bf_o = (d.bf16 & 2) != 0, bf_do = (d.bf16 & 4) != 0, bf_dx = (d.bf16 & 8) != 0;   // o, d_o, dq likewise
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Qs = smem;                        // [128][128 B] K-major   (A of S)
  uint8_t* dOs = Qs + TQ * 128;              // [128][128 B] K-major   (A of dP)
  uint8_t* Ks = dOs + TQ * 128;              // [ 64][128 B] K-major   (B of S)
  uint8_t* Vs = Ks + TK * 128;               // [ 64][128 B] K-major   (B of dP)
  uint8_t* Kmn = Vs + TK * 128;              // 2 x [ 64][128 B] MN-major  (B of dQ), double buffered
  uint8_t* dSs = Kmn + 2 * TK * 128;         // 2 regions [128][128 B] K-major (A of dQ)

  if (tid == 0) {
    a_mbar_init(a_smem_u32(&bar_a), 1);
    a_mbar_init(a_smem_u32(&bar_b), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(&tmem_slot)), "r"(ATC_DQ_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_ready();                 // barrier init and TMEM allocation ran under the tail of the preceding kernel
  dc = make_drop(d.rng, d.p);
  stage_rows256<false>(Qs, d.q, d.ldq, d.B, b, h, hd, i0, Lq, TQ, bf);
  stage_rows256<false>(dOs, d.d_o, d.lddo, d.B, b, h, hd, i0, Lq, TQ, bf_do);
  stage_rows256<false>(Ks, d.k, d.ldk, d.B, b, h, hd, 0, Lk, TK, bf);
  stage_rows256<false>(Vs, d.v, d.ldv, d.B, b, h, hd, 0, Lk, TK, bf);
  stage_rows256<true>(Kmn, d.k, d.ldk, d.B, b, h, hd, 0, Lk, TK, bf);
  const int i = i0 + row;
  const int irow = min(i, Lq - 1);
  // delta[r] = sum_c o[r, c] * d_o[r, c]: one warp per row, one lane per column (coalesced 100-byte reads and a shuffle
  // reduction) -- a thread-per-row loop costs 2 * hd scalar loads of 32 scattered sectors each, ~7 us per CTA of L1
  // wavefronts.  Four rows in flight per warp; the rows' owners pick their value up from shared memory after the barrier.
  {
    constexpr int NW = AQ_THREADS / 32;
#pragma unroll 1
    for (int r0 = warp; r0 < TQ; r0 += 4 * NW) {
      float pr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ii = i0 + r0 + u * NW;
        pr[u] = 0.f;
        if (r0 + u * NW < TQ && ii < Lq && lane < hd) {
          const int64_t oo = ((int64_t)ii * d.B + b) * d.ldo + h * hd + lane, og = ((int64_t)ii * d.B + b) * d.lddo + h * hd + lane;
          pr[u] = ld1_any(d.o, oo, bf_o) * ld1_any(d.d_o, og, bf_do);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float sum = warp_sum(pr[u]);
        const int r = r0 + u * NW;
        if (lane == 0 && r < TQ) {
          delta_s[r] = sum;
          if (i0 + r < Lq) d.delta[(int64_t)bh * Lq + i0 + r] = sum;
        }
      }
    }
  }
  float lse2 = 0.f;
  if (i < Lq) lse2 = d.lse[(int64_t)bh * Lq + i] * 1.4426950408889634f;
  stage_wait();
  if (bf_do) fix_rows256<false>(dOs, d.d_o, d.lddo, d.B, b, h, hd, i0, TQ);
  if (bf) {
    fix_rows256<false>(Qs, d.q, d.ldq, d.B, b, h, hd, i0, TQ);
    fix_rows256<false>(Ks, d.k, d.ldk, d.B, b, h, hd, 0, TK);
    fix_rows256<false>(Vs, d.v, d.ldv, d.B, b, h, hd, 0, TK);
    fix_rows256<true>(Kmn, d.k, d.ldk, d.B, b, h, hd, 0, TK);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const float delta = delta_s[row];
  const uint32_t tmem = tmem_slot;
  const uint32_t t_s = tmem, t_dp = tmem + 64, t_dq = tmem + 128;
  const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
  const int i_last = min(Lq, i0 + TQ) - 1;
  const int j_end = min(Lk, i_last + off + 1);
  const int T = (j_end + TK - 1) / TK;
  const uint32_t id_s = idesc_tf32(TQ, TK, false, false);
  const uint32_t id_q = idesc_tf32(TQ, HP, false, true);
  const float c2 = d.scale * 1.4426950408889634f;
  auto issue_s = [&]() {                      // S and dP of the tile currently in Ks / Vs
#pragma unroll
    for (int k = 0; k < HP / 8; ++k)
      a_mma_tf32(t_s, desc_kmajor(a_smem_u32(Qs) + k * 32), desc_kmajor(a_smem_u32(Ks) + k * 32), id_s, k != 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < HP / 8; ++k)
      a_mma_tf32(t_dp, desc_kmajor(a_smem_u32(dOs) + k * 32), desc_kmajor(a_smem_u32(Vs) + k * 32), id_s, k != 0 ? 1u : 0u);
    a_commit(a_smem_u32(&bar_a));
  };
  if (tid == 0) issue_s();
  const uint64_t idx_row = ((uint64_t)((int64_t)bh * Lq + irow)) * (uint64_t)Lk4 + (uint64_t)(half * 32);
  const int KW = (Lk + 31) >> 5;
  const uint32_t* wrow = (d.keep_bits != nullptr && dc.on) ? d.keep_bits + ((int64_t)bh * Lq + irow) * KW : nullptr;
  uint32_t kbits = !dc.on ? 0u : wrow != nullptr ? (half < KW ? wrow[half] : 0u) : row_keep_bits32(dc, idx_row);
  uint8_t* my_ds = dSs + half * (TQ * 128);

  QTRACE(0);
  for (int t = 0; t < T; ++t) {
    const int j0 = t * TK;
    const uint32_t ph = (uint32_t)t & 1u;
    QTRACE(8 + t * 8 + 0);
    a_mbar_wait(a_smem_u32(&bar_a), ph);       // S_t, dP_t ready; Ks / Vs are free
    if (t > 0) a_mbar_wait(a_smem_u32(&bar_b), (uint32_t)(t - 1) & 1u);   // dQ_{t-1} retired: dS and the other Kmn buffer are free
    QTRACE(8 + t * 8 + 1);
    tc_fence_after();
    if (t + 1 < T) {
      stage_rows256<false>(Ks, d.k, d.ldk, d.B, b, h, hd, j0 + TK, Lk, TK, bf);
      stage_rows256<false>(Vs, d.v, d.ldv, d.B, b, h, hd, j0 + TK, Lk, TK, bf);
      stage_rows256<true>(Kmn + ((t + 1) & 1) * (TK * 128), d.k, d.ldk, d.B, b, h, hd, j0 + TK, Lk, TK, bf);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    uint32_t kbits_next = 0u;                  // stored keep bits of the next tile: loaded now, used one iteration later
    if (wrow != nullptr && t + 1 < T && 2 * (t + 1) + half < KW) kbits_next = wrow[2 * (t + 1) + half];
    QTRACE(8 + t * 8 + 2);
    float s[32], dp[32];
    a_tmem_ld32(t_s + lane_addr + half * 32, s);
    a_tmem_ld32(t_dp + lane_addr + half * 32, dp);
    tc_fence_before();
    QTRACE(8 + t * 8 + 3);
    const bool tile_open = (j0 + TK <= Lk) && (j0 + TK - 1 - i0 < 1 + off);      // CTA-uniform
    if (tile_open)
      dq_math<false>(s, dp, lse2, delta, kbits, dc.on, dc.inv_keep, c2, d.scale, i, j0 + half * 32, Lk, off, my_ds, row);
    else
      dq_math<true>(s, dp, lse2, delta, kbits, dc.on, dc.inv_keep, c2, d.scale, i, j0 + half * 32, Lk, off, my_ds, row);
    QTRACE(8 + t * 8 + 4);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (bf && t + 1 < T) {
      fix_rows256<false>(Ks, d.k, d.ldk, d.B, b, h, hd, j0 + TK, TK);
      fix_rows256<false>(Vs, d.v, d.ldv, d.B, b, h, hd, j0 + TK, TK);
      fix_rows256<true>(Kmn + ((t + 1) & 1) * (TK * 128), d.k, d.ldk, d.B, b, h, hd, j0 + TK, TK);
    }
    QTRACE(8 + t * 8 + 5);
    fence_async_smem();
    __syncthreads();                           // dS written, S / dP read by everyone, next K / V / Kmn landed
    QTRACE(8 + t * 8 + 6);
    if (tid == 0) {                            // critical path first: the next tile's S / dP
      tc_fence_after();
      if (t + 1 < T) issue_s();
    } else if (tid == 128) {                   // dQ_t on a second issuing thread (own barrier)
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < TK / 8; ++kk)
        a_mma_tf32(t_dq, desc_kmajor(a_smem_u32(dSs) + (kk >> 2) * (TQ * 128) + (kk & 3) * 32),
                   desc_mnmajor(a_smem_u32(Kmn + (t & 1) * (TK * 128)) + kk * 1024), id_q, (t | kk) != 0 ? 1u : 0u);
      a_commit(a_smem_u32(&bar_b));
    }
    if (dc.on && t + 1 < T)                    // forward kernel's bits if it stored them, else re-drawn while the MMAs run
      kbits = wrow != nullptr ? kbits_next : row_keep_bits32(dc, idx_row + (uint64_t)(j0 + TK));
    QTRACE(8 + t * 8 + 7);
  }
  QTRACE(1);
  {
    a_mbar_wait(a_smem_u32(&bar_b), (uint32_t)(T - 1) & 1u);
    tc_fence_after();
    float dq[16];
    a_tmem_ld16(t_dq + lane_addr + half * 16, dq);
    if (i < Lq) park_row<16>(reinterpret_cast<float*>(Qs), row, half * 16, dq);      // all MMAs have retired: Q's tile is idle
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ATC_DQ_TMEM) : "memory");
  }
  rows_out<AQ_THREADS>(reinterpret_cast<const float*>(Qs), d.dq, d.lddq, d.B, b, h, hd, i0, Lq, TQ, bf_dx);
}

// ============================================================================ backward: dK, dV
// One CTA (256 threads) per (batch, head, 128-row KEY tile).  Warp w serves TMEM lane quadrant w & 3 (key rows)
// and query-column half w >> 2 of every 32-row query tile.  Per query tile: S^T = K Q^T and dP^T = V dO^T on
// the tensor core, P~^T and dS^T on CUDA cores, then dV += P~^T dO and dK += dS^T Q on the tensor core
// (accumulators stay in TMEM).  Software pipelined: the four operand tiles of query tile t+1 are fetched
// (cp.async, double buffered) and its S^T / dP^T MMAs are issued while tile t is processed; the Philox keep
// bits of tile t+1 are drawn while the MMAs run; one __syncthreads per tile.
constexpr int TI = 32;
constexpr int ATC_DKV_TMEM = 128;
constexpr int ATC_DKV_SMEM = 2 * TQ * 128 + 2 * 4 * TI * 128 + 2 * TQ * 128 + 1024;
constexpr int AB_THREADS = 256;

// keep bits of the 16 query columns (half * 16 ..) x my key row for the query tile starting at i0.  The 4 lanes
// of a quad own key rows j..j+3 = one Philox group per query row: lane k draws the groups of query rows
// 4 q + k (q = 0..3), packs the 16 compare results, and the quad exchanges the packed words (4 shuffles).
__device__ __forceinline__ void dkv_keep_bits(const DropCtx& dc, int bh, int Lq, int Lk4, int i0h, int jrow, uint32_t (&w)[4]) {
  const int lane = threadIdx.x & 31, k = lane & 3;
  uint32_t bits = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int iq = min(i0h + 4 * q + k, Lq - 1);
    const uint64_t idx = ((uint64_t)((int64_t)bh * Lq + iq)) * (uint64_t)Lk4 + (uint64_t)(jrow & ~3);
    const uint4 r = drop_rand4(dc, idx >> 2);
    bits |= ((r.x >= dc.thr ? 1u : 0u) | (r.y >= dc.thr ? 2u : 0u) | (r.z >= dc.thr ? 4u : 0u) | (r.w >= dc.thr ? 8u : 0u)) << (4 * q);
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) w[m] = __shfl_sync(0xffffffffu, bits, (lane & ~3) + m) >> k;   // bit 4 q of w[m]: query 4 q + m
}

// same result from the words the forward kernel stored: query c of my 16 columns, bit (key j) of word (j / 32)
__device__ __forceinline__ void dkv_keep_bits_stored(const uint32_t* bits, int KW, int bh, int Lq, int i0h, int jrow, uint32_t (&w)[4]) {
  const int jw = jrow >> 5, jb = jrow & 31;
  uint32_t wb[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int ii = i0h + c;
    wb[c] = (ii < Lq && jw < KW) ? __ldg(bits + ((int64_t)bh * Lq + ii) * KW + jw) : 0u;
  }
  w[0] = w[1] = w[2] = w[3] = 0u;
#pragma unroll
  for (int c = 0; c < 16; ++c) w[c & 3] |= ((wb[c] >> jb) & 1u) << (c & ~3);
}

template <bool MASKED>
__device__ __forceinline__ void dkv_math(const float (&st)[16], const float (&dpt)[16], const float* lse2, const float* dlt,
                                         const uint32_t (&kw)[4], bool drop_on, float inv_keep, float c2, float scale, int i0h, int j,
                                         int Lq, int Lk, int off, uint8_t* pt_row, uint8_t* ds_row, int row, int cb0) {
#pragma unroll
  for (int c4 = 0; c4 < 16; c4 += 4) {
    float pt[4], ds[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c4 + e;
      float p = fast_exp2(fmaf(st[c], c2, -lse2[c]));
      if (MASKED) {
        const int ii = i0h + c;
        const bool open = (ii < Lq) && (j < Lk) && (j - ii < 1 + off);
        p = open ? p : 0.f;
      }
      const float keep = (!drop_on || ((kw[e] >> c4) & 1u)) ? inv_keep : 0.f;      // query c = 4 (c4 / 4) + e -> word e, bit c4
      pt[e] = p * keep;
      ds[e] = p * (dpt[c] * keep - dlt[c]) * scale;
    }
    *reinterpret_cast<float4*>(pt_row + swz128(row, (cb0 + c4) * 4)) = make_float4(to_tf32(pt[0]), to_tf32(pt[1]), to_tf32(pt[2]), to_tf32(pt[3]));
    *reinterpret_cast<float4*>(ds_row + swz128(row, (cb0 + c4) * 4)) = make_float4(to_tf32(ds[0]), to_tf32(ds[1]), to_tf32(ds[2]), to_tf32(ds[3]));
  }
}

__global__ void __launch_bounds__(AB_THREADS, 2) attn_bwd_dkv_tc_kernel(const __grid_constant__ Group<mtb_attn_bwd_desc> g) {
  pdl_begin();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a, bar_b;
  __shared__ uint32_t tmem_slot;
  __shared__ float col_lse2[2][TI], col_delta[2][TI];
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_attn_bwd_desc& d = g.d[pi];
  const int BH = d.B * d.H;
  const int bh = local % BH, kt = local / BH;       // key tile 0 sees every query tile: heaviest first
  const int b = bh / d.H, h = bh - b * d.H;
  const int j0 = kt * TQ;
  const int Lq = d.Lq, Lk = d.Lk, hd = d.hd;
  const int off = abs(Lk - Lq);
  const int Lk4 = a_round4(Lk);
  DropCtx dc;                                       // filled after pdl_ready(): the device counter lives in global memory
  const bool bf = (d.bf16 & 1) != 0;                // q / k / v are bfloat16 in HBM
  const bool bf_do = (d.bf16 & 4) != 0, bf_dx = (d.bf16 & 8) != 0;   // d_o, dk / dv likewise
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;                 // key row inside the tile = TMEM lane
  const int j = j0 + row;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Ks = smem;                        // [128][128 B] K-major (A of S^T)
  uint8_t* Vs = Ks + TQ * 128;               // [128][128 B] K-major (A of dP^T)
  uint8_t* QB = Vs + TQ * 128;               // 2 buffers x { Q K-major (B of S^T), dO K-major (B of dP^T), Q MN-major (B of dK), dO MN-major (B of dV) }, [32][128 B] each
  uint8_t* PTs = QB + 2 * 4 * TI * 128;      // [128][128 B] K-major (A of dV), 32 query columns
  uint8_t* dSTs = PTs + TQ * 128;            // [128][128 B] K-major (A of dK)

  if (tid == 0) {
    a_mbar_init(a_smem_u32(&bar_a), 1);
    a_mbar_init(a_smem_u32(&bar_b), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(&tmem_slot)), "r"(ATC_DKV_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_ready();                 // barrier init and TMEM allocation ran under the tail of the preceding kernel
  dc = make_drop(d.rng, d.p);
  const int i_begin = (max(0, j0 - off) / TI) * TI;
  const int T = (Lq - i_begin + TI - 1) / TI;
  const float c2 = d.scale * 1.4426950408889634f;
  auto fix_q = [&](int t) {                   // bf16 staging, second half (see fix_rows256)
    uint8_t* qb = QB + (t & 1) * (4 * TI * 128);
    const int i0 = i_begin + t * TI;
    if (bf) {
      fix_rows256<false>(qb, d.q, d.ldq, d.B, b, h, hd, i0, TI);
      fix_rows256<true>(qb + 2 * TI * 128, d.q, d.ldq, d.B, b, h, hd, i0, TI);
    }
    if (bf_do) {
      fix_rows256<false>(qb + TI * 128, d.d_o, d.lddo, d.B, b, h, hd, i0, TI);
      fix_rows256<true>(qb + 3 * TI * 128, d.d_o, d.lddo, d.B, b, h, hd, i0, TI);
    }
  };
  auto stage_q = [&](int t) {                 // operand tiles + per-query-row lse / delta of query tile t
    uint8_t* qb = QB + (t & 1) * (4 * TI * 128);
    const int i0 = i_begin + t * TI;
    stage_rows256<false>(qb, d.q, d.ldq, d.B, b, h, hd, i0, Lq, TI, bf);
    stage_rows256<false>(qb + TI * 128, d.d_o, d.lddo, d.B, b, h, hd, i0, Lq, TI, bf_do);
    stage_rows256<true>(qb + 2 * TI * 128, d.q, d.ldq, d.B, b, h, hd, i0, Lq, TI, bf);
    stage_rows256<true>(qb + 3 * TI * 128, d.d_o, d.lddo, d.B, b, h, hd, i0, Lq, TI, bf_do);
    if (tid < TI) {
      const int ii = i0 + tid;
      col_lse2[t & 1][tid] = ii < Lq ? d.lse[(int64_t)bh * Lq + ii] * 1.4426950408889634f : 0.f;
      col_delta[t & 1][tid] = ii < Lq ? d.delta[(int64_t)bh * Lq + ii] : 0.f;
    }
  };
  stage_rows256<false>(Ks, d.k, d.ldk, d.B, b, h, hd, j0, Lk, TQ, bf);
  stage_rows256<false>(Vs, d.v, d.ldv, d.B, b, h, hd, j0, Lk, TQ, bf);
  if (T > 0) stage_q(0);
  stage_wait();
  if (bf) {
    fix_rows256<false>(Ks, d.k, d.ldk, d.B, b, h, hd, j0, TQ);
    fix_rows256<false>(Vs, d.v, d.ldv, d.B, b, h, hd, j0, TQ);
  }
  if ((bf || bf_do) && T > 0) fix_q(0);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_st = tmem, t_dpt = tmem + 32, t_dv = tmem + 64, t_dk = tmem + 96;
  const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
  const uint32_t id_s = idesc_tf32(TQ, TI, false, false);
  const uint32_t id_o = idesc_tf32(TQ, HP, false, true);
  auto issue_s = [&](int t) {                 // S^T_t and dP^T_t
    uint8_t* qb = QB + (t & 1) * (4 * TI * 128);
#pragma unroll
    for (int k = 0; k < HP / 8; ++k)
      a_mma_tf32(t_st, desc_kmajor(a_smem_u32(Ks) + k * 32), desc_kmajor(a_smem_u32(qb) + k * 32), id_s, k != 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < HP / 8; ++k)
      a_mma_tf32(t_dpt, desc_kmajor(a_smem_u32(Vs) + k * 32), desc_kmajor(a_smem_u32(qb + TI * 128) + k * 32), id_s, k != 0 ? 1u : 0u);
    a_commit(a_smem_u32(&bar_a));
  };
  if (tid == 0 && T > 0) issue_s(0);
  uint32_t kw[4] = {0u, 0u, 0u, 0u};
  const int KW = (Lk + 31) >> 5;
  const uint32_t* kbits_g = dc.on ? d.keep_bits : nullptr;
  if (dc.on && T > 0) {
    if (kbits_g != nullptr) dkv_keep_bits_stored(kbits_g, KW, bh, Lq, i_begin + half * 16, j, kw);
    else dkv_keep_bits(dc, bh, Lq, Lk4, i_begin + half * 16, j, kw);
  }

  for (int t = 0; t < T; ++t) {
    const int i0 = i_begin + t * TI;
    const uint32_t ph = (uint32_t)t & 1u;
    uint8_t* qb = QB + (t & 1) * (4 * TI * 128);
    a_mbar_wait(a_smem_u32(&bar_a), ph);       // S^T_t, dP^T_t ready
    if (t > 0) a_mbar_wait(a_smem_u32(&bar_b), (uint32_t)(t - 1) & 1u);   // dV / dK of tile t-1 retired: P~^T / dS^T and the other operand buffer are free
    tc_fence_after();
    if (t + 1 < T) stage_q(t + 1);             // the other buffer was last read by tile t-1's MMAs
    asm volatile("cp.async.commit_group;" ::: "memory");
    float st[16], dpt[16];
    a_tmem_ld16(t_st + lane_addr + half * 16, st);
    a_tmem_ld16(t_dpt + lane_addr + half * 16, dpt);
    tc_fence_before();
    const bool tile_open = (i0 + TI <= Lq) && (j0 + TQ <= Lk) && (j0 + TQ - 1 - i0 < 1 + off);   // CTA-uniform
    const float* l2 = &col_lse2[t & 1][half * 16];
    const float* dl = &col_delta[t & 1][half * 16];
    if (tile_open)
      dkv_math<false>(st, dpt, l2, dl, kw, dc.on, dc.inv_keep, c2, d.scale, i0 + half * 16, j, Lq, Lk, off, PTs, dSTs, row, half * 16);
    else
      dkv_math<true>(st, dpt, l2, dl, kw, dc.on, dc.inv_keep, c2, d.scale, i0 + half * 16, j, Lq, Lk, off, PTs, dSTs, row, half * 16);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if ((bf || bf_do) && t + 1 < T) fix_q(t + 1);
    fence_async_smem();
    __syncthreads();                           // P~^T / dS^T written, S^T / dP^T read by everyone, next operand tiles landed
    if (tid == 0) {                            // critical path first: the next tile's S^T / dP^T
      tc_fence_after();
      if (t + 1 < T) issue_s(t + 1);
    } else if (tid == 128) {                   // dV_t / dK_t on a second issuing thread (own barrier)
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < TI / 8; ++kk)
        a_mma_tf32(t_dv, desc_kmajor(a_smem_u32(PTs) + kk * 32), desc_mnmajor(a_smem_u32(qb + 3 * TI * 128) + kk * 1024), id_o,
                   (t | kk) != 0 ? 1u : 0u);
#pragma unroll
      for (int kk = 0; kk < TI / 8; ++kk)
        a_mma_tf32(t_dk, desc_kmajor(a_smem_u32(dSTs) + kk * 32), desc_mnmajor(a_smem_u32(qb + 2 * TI * 128) + kk * 1024), id_o,
                   (t | kk) != 0 ? 1u : 0u);
      a_commit(a_smem_u32(&bar_b));
    }
    if (dc.on && t + 1 < T) {                  // next tile's keep bits while the MMAs run: stored words, else re-drawn
      if (kbits_g != nullptr) dkv_keep_bits_stored(kbits_g, KW, bh, Lq, i0 + TI + half * 16, j, kw);
      else dkv_keep_bits(dc, bh, Lq, Lk4, i0 + TI + half * 16, j, kw);
    }
  }
  {
    float dv[16], dk[16];
    if (T > 0) {
      a_mbar_wait(a_smem_u32(&bar_b), (uint32_t)(T - 1) & 1u);
      tc_fence_after();
      a_tmem_ld16(t_dv + lane_addr + half * 16, dv);
      a_tmem_ld16(t_dk + lane_addr + half * 16, dk);
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) { dv[c] = 0.f; dk[c] = 0.f; }
    }
    if (j < Lk) {                                     // all MMAs have retired: the K and V tiles are idle
      park_row<16>(reinterpret_cast<float*>(Ks), row, half * 16, dk);
      park_row<16>(reinterpret_cast<float*>(Vs), row, half * 16, dv);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ATC_DKV_TMEM) : "memory");
  }
  rows_out<AB_THREADS>(reinterpret_cast<const float*>(Ks), d.dk, d.lddk, d.B, b, h, hd, j0, Lk, TQ, bf_dx);
  rows_out<AB_THREADS>(reinterpret_cast<const float*>(Vs), d.dv, d.lddv, d.B, b, h, hd, j0, Lk, TQ, bf_dx);
}

int attn_fwd_simt(const mtb_attn_desc* d, int n, cudaStream_t st);

int attn_fwd_tc(const mtb_attn_desc* d, int n, cudaStream_t st) {
  mtb_attn_desc rest[MTB_MAX_GROUP];
  Group<mtb_attn_desc> g, gs;
  int ntc = 0, nshort = 0, nrest = 0, tot = 0, tots = 0;
  static const bool no_short = getenv("MTB_ATTN_NO_SHORT") != nullptr;
  for (int i = 0; i < n; ++i) {
    MTB_CHECK(d[i].hd <= HP || !d[i].bf16, "attn_fwd: bf16 operands need head_dim <= %d (problem %d)", HP, i);
    if (d[i].hd > HP) { rest[nrest++] = d[i]; continue; }
    if (!no_short && d[i].Lq <= SQ && d[i].Lk <= SQ && d[i].B * d[i].H >= 2) {   // two (batch, head) problems per CTA
      gs.d[nshort] = d[i];
      gs.start[nshort] = tots;
      tots += (d[i].B * d[i].H + 1) / 2;
      ++nshort;
      continue;
    }
    g.d[ntc] = d[i];
    g.start[ntc] = tot;
    tot += d[i].B * d[i].H * ((d[i].Lq + TQ - 1) / TQ);
    ++ntc;
  }
  g.n = ntc;
  g.start[ntc] = tot;
  gs.n = nshort;
  gs.start[nshort] = tots;
  static bool attr = false;
  if (!attr) {
    MTB_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_FWD_SMEM));
    MTB_CUDA(cudaFuncSetAttribute(attn_fwd_tc_short_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SHORT_SMEM));
    attr = true;
  }
  if (tot > 0) {
    MTB_CUDA(launch_k(attn_fwd_tc_kernel, dim3(tot), dim3(AF_THREADS), ATC_FWD_SMEM, st, g));
    mtb::note_launch();
    MTB_CUDA(cudaGetLastError());
  }
  if (tots > 0) {
    MTB_CUDA(launch_k(attn_fwd_tc_short_kernel, dim3(tots), dim3(AF_THREADS), ATC_SHORT_SMEM, st, gs));
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
    MTB_CHECK(d[i].hd <= HP || !d[i].bf16, "attn_bwd: bf16 operands need head_dim <= %d (problem %d)", HP, i);
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
    MTB_CUDA(launch_k(attn_bwd_dq_tc_kernel, dim3(totq), dim3(AQ_THREADS), ATC_DQ_SMEM, st, gq));
    mtb::note_launch();
    MTB_CUDA(cudaGetLastError());
    MTB_CUDA(launch_k(attn_bwd_dkv_tc_kernel, dim3(totk), dim3(AB_THREADS), ATC_DKV_SMEM, st, gk));
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
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_fwd_tc_short_kernel) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_bwd_dq_tc_kernel) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_bwd_dkv_tc_kernel) != cudaSuccess) ++bad; }
  return bad;
}
}  // namespace mtb
