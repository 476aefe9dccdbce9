// linear_tc.cu -- tcgen05 / TMEM / TMA grouped GEMM engine behind mtb_linear_fwd / mtb_linear_bwd
// (gemm mode 1).  modules/dynamic_multihead_attention.py:259-282, modules/dynamic_layers.py:15-25.
//
// One warp-specialised kernel (192 threads, 1 CTA / SM, one 128 x BJ output tile per CTA):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor loads of fp32 operand tiles into a 4-stage
//              128B-swizzled shared-memory ring, completion on mbarriers
//   warp 1   : TMEM allocator + MMA issuer -- one elected lane issues tcgen05.mma.kind::tf32
//              (A and B straight from shared memory, fp32 accumulator in tensor memory) and
//              tcgen05.commit to release ring slots / publish the accumulator
//   warps 2-5: epilogue -- tcgen05.ld of the accumulator, bias / ReLU / Philox dropout /
//              accumulate / atomic split-K, 128-bit stores
// The same kernel serves forward (both operands K-major), dgrad (B MN-major) and wgrad (both
// MN-major, split over the token axis) by switching the shared-memory matrix descriptors; fp32
// data is consumed directly as TF32 (no conversion pass, no bf16 copies of weights in HBM).
// Partial tiles (M % 128, N % BJ, K % 32 such as K = 200) rely on TMA zero fill.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <time.h>

namespace mtb {

int linear_fwd_simt(const mtb_linear_desc* d, int n, cudaStream_t st);
int linear_bwd_simt(const mtb_linear_bwd_desc* d, int n, cudaStream_t st);

constexpr int TC_BI = 128;                 // UMMA M
constexpr int TC_BR = 32;                  // reduction elements per stage (32 fp32 = one 128 B swizzle row)
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_A_BYTES = TC_BI * 128;    // 16 KB
constexpr int TC_SMEM_BUDGET = 100 * 1024;  // <= half an SM's shared memory: two CTAs per SM overlap
                                           // one tile's epilogue with the other's main loop
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 256;

constexpr int TC_MAXSEG = 16;

// Every logical axis (I rows of D, J cols of D, R reduction) may be SEGMENTED: the compact
// axis is n_seg blocks of seg_len elements and block s lives at physical block index seg[s] of
// the (larger) parameter tensor.  This is how the reference's active_mask gathers
// (src/dynamic_models2.py:243-251: unions of d-wide column blocks) reach TMA: operands are 4-D
// tensor maps {col_in_block, col_block, row_in_block, row_block}, tiles never straddle a block
// and TMA zero-fills the tail of a partially covered block.  Unsegmented axes use one block.
struct alignas(64) TcProblem {
  CUtensorMap mapA, mapB, mapC;   // operands (loads) and the output (store / reduce-add)
  float* C; int64_t ldc;
  const float* bias;
  int I, J, R;         // compact sizes
  int i_len, i_nseg, j_len, j_nseg, r_len, r_nseg;
  uint8_t a_iseg[TC_MAXSEG], a_rseg[TC_MAXSEG], b_jseg[TC_MAXSEG], b_rseg[TC_MAXSEG];
  uint8_t c_iseg[TC_MAXSEG], c_jseg[TC_MAXSEG], bias_seg[TC_MAXSEG];
  int BJ;              // tile width along J (multiple of 16, <= 256)
  int stages;          // shared-memory ring depth (2..4), stage = 16 KB (A) + BJ * 128 B (B)
  int a_mn, b_mn;      // operand is MN-major in shared memory
  int splits;          // split of the reduction range (epi == 2)
  int epi;             // 0 store, 1 +=, 2 atomicAdd
  int act; float p; mtb_rng rng;
  int ab16;            // operands are bf16: 64 reduction elements per 128-byte row, UMMA_K = 16, tcgen05.mma.kind::f16
  int c16;             // output is bf16 (fp32 accumulator converted in the epilogue, 64-column TMA boxes; BJ % 64 == 0)
};

// The problem list travels in the kernel parameter block.  Parameter blocks above 4 KB take a slower launch path
// (measured: ~6.5 us of GPU idle before every GEMM kernel with the 24-problem, 15 KB block vs ~2 us for the other
// kernels), so the kernel is instantiated for several capacities and a launch uses the smallest that fits.
template <int CAP>
struct TcGroupT {
  TcProblem d[CAP];
  int start[CAP + 1];
  int n;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"     // %3: suspend-time hint -- the warp sleeps in
      "selp.u32 %0, 1, 0, p;\n\t}"                                        // hardware instead of spinning on issue slots
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must trap, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (int it = 0; it < (1 << 22); ++it)
    if (mbar_try_wait(bar, parity)) return;
  printf("mtb gemm_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), 128-byte swizzle:
//   [0,14) start >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) layout = 2
//   K-major : SWIZZLE_128B (2); rows 128 B apart, 8-row groups SBO = 1024 B apart (LBO unused)
//   MN-major: 32-bit operands only exist in the SWIZZLE_128B_BASE32B (1) layout (cutlass
//             sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only available smem
//             layout"): atom = 32 MN elements (128 B) x 4 reduction rows; MN atoms LBO = 4096 B apart
//             (one 32 x 32 TMA box each), 4-row reduction groups SBO = 512 B apart
//   MN-major, 16-bit operands: plain SWIZZLE_128B (2): atom = 64 MN elements (128 B) x 8 reduction rows; MN atoms LBO =
//             8192 B apart (one 64-wide x 64-row TMA box each), 8-row reduction groups SBO = 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, bool mn_major, bool ab16 = false) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((mn_major ? (ab16 ? 8192u : 4096u) : 16u) >> 4) << 16;
  d |= (uint64_t)((mn_major ? (ab16 ? 1024u : 512u) : 1024u) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)((mn_major && !ab16) ? 1 : 2) << 61;
  return d;
}


#ifdef MTB_TC_TRACE
__device__ unsigned long long g_tc_trace[64];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TRACE(i) do { if (blockIdx.x == 0) { asm volatile("" ::: "memory"); g_tc_trace[(i)] = gtime(); asm volatile("" ::: "memory"); } } while (0)
#else
#define TRACE(i) do { } while (0)
#endif
#ifdef MTB_TC_TRACE_FINE       // per-slab / per-chunk stamps: every stamp costs ~0.3 us, so the coarse build is the one to read totals from
#define TRACE_FINE(i) TRACE(i)
#else
#define TRACE_FINE(i) do { } while (0)
#endif

template <int CAP>
__global__ void __launch_bounds__(TC_THREADS, 2) gemm_tc_kernel(const __grid_constant__ TcGroupT<CAP> g) {
  pdl_begin();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[256];   // bias of this tile's columns, staged while the main loop runs

  if (threadIdx.x == 0) TRACE(0);
  int p = 0;
  while (p + 1 < g.n && (int)blockIdx.x >= g.start[p + 1]) ++p;
  const int local = blockIdx.x - g.start[p];
  const TcProblem& P = g.d[p];
  const int tpi = (P.i_len + TC_BI - 1) / TC_BI, tpj = (P.j_len + P.BJ - 1) / P.BJ;   // tiles per block
  const int ti_n = tpi * P.i_nseg, tj_n = tpj * P.j_nseg;
  const int split = local / (ti_n * tj_n);
  const int tile = local - split * (ti_n * tj_n);
  const int ti = tile / tj_n, tj = tile - ti * tj_n;
  const int is = ti / tpi, il0 = (ti - is * tpi) * TC_BI;        // block / offset inside block along I
  const int js = tj / tpj, jl0 = (tj - js * tpj) * P.BJ;
  const bool ab16 = P.ab16 != 0;
  const int BR = ab16 ? 64 : TC_BR;                              // reduction elements per stage (one 128-byte row)
  const int MNW = ab16 ? 64 : 32;                                // MN-major operands: elements per 128-byte box row
  const uint32_t MNBOX = ab16 ? 8192u : 4096u;                   // bytes of one MN-major box (MNW wide x BR rows)
  const int kbps = (P.r_len + BR - 1) / BR;                      // reduction slabs per block
  const int nkb_total = kbps * P.r_nseg;
  const int per = (nkb_total + P.splits - 1) / P.splits;
  const int kb0 = split * per;
  const int kb1 = min(nkb_total, kb0 + per);
  const int nkb = kb1 - kb0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_boxes = P.b_mn ? (P.BJ + MNW - 1) / MNW : 1;
  const uint32_t b_bytes = P.b_mn ? (uint32_t)b_boxes * MNBOX : (uint32_t)P.BJ * 128u;
  const int TC_STAGES = P.stages;
  const uint32_t TC_STAGE_BYTES = (uint32_t)TC_A_BYTES + (((uint32_t)P.BJ * 128u + 1023u) & ~1023u);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.mapB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.mapC) : "memory");
    for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;
  if (threadIdx.x == 0) TRACE(1);
  pdl_ready();                 // everything above ran under the tail of the preceding kernel; global memory from here on

  if (nkb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        // ===== TMA producer =====
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = kb % TC_STAGES;
          const uint32_t ph = (uint32_t)(kb / TC_STAGES) & 1u;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          if (kb < 16) TRACE_FINE(8 + kb);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, (uint32_t)TC_A_BYTES + b_bytes);
          const uint32_t sa = smem_base + (uint32_t)s * TC_STAGE_BYTES;
          const uint32_t sb = sa + TC_A_BYTES;
          const int rs = (kb0 + kb) / kbps;
          const int rl0 = ((kb0 + kb) - rs * kbps) * BR;
          // 4-D coordinates {col_in_block, col_block, row_in_block, row_block}; K-major operands have the
          // reduction on the (contiguous) column axis, MN-major operands on the row axis
          if (P.a_mn) {
            for (int b = 0; b < TC_BI / MNW; ++b)
              tma_load_4d(sa + b * MNBOX, &P.mapA, fb, il0 + b * MNW, P.a_iseg[is], rl0, P.a_rseg[rs]);
          } else {
            tma_load_4d(sa, &P.mapA, fb, rl0, P.a_rseg[rs], il0, P.a_iseg[is]);
          }
          if (P.b_mn) {
            for (int b = 0; b < b_boxes; ++b)
              tma_load_4d(sb + b * MNBOX, &P.mapB, fb, jl0 + b * MNW, P.b_jseg[js], rl0, P.b_rseg[rs]);
          } else {
            tma_load_4d(sb, &P.mapB, fb, rl0, P.b_rseg[rs], jl0, P.b_jseg[js]);
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // ===== MMA issuer =====
        // instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, majors, N >> 3, M >> 4
        //   operand format (bits [7,10) and [10,13)): 1 = bf16, 2 = tf32
        const uint32_t fmt = ab16 ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(P.a_mn ? 1 : 0) << 15) |
                               ((uint32_t)(P.b_mn ? 1 : 0) << 16) | ((uint32_t)(P.BJ >> 3) << 17) |
                               ((uint32_t)(TC_BI >> 4) << 24);
        // bytes per UMMA_K step (8 tf32 / 16 bf16 = 32 B of a K-major row; 8 / 16 reduction rows of 128 B when MN-major)
        const uint32_t mn_step = ab16 ? 2048u : 1024u;
        const uint32_t a_step = P.a_mn ? mn_step : 32u;
        const uint32_t b_step = P.b_mn ? mn_step : 32u;
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = kb % TC_STAGES;
          const uint32_t ph = (uint32_t)(kb / TC_STAGES) & 1u;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          if (kb < 16) TRACE_FINE(24 + kb);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + (uint32_t)s * TC_STAGE_BYTES;
          const uint32_t sb = sa + TC_A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BR / 8; ++k) {
            const uint64_t ad = make_desc(sa + k * a_step, P.a_mn, ab16);
            const uint64_t bd = make_desc(sb + k * b_step, P.b_mn, ab16);
            if (ab16) tcgen05_mma_f16(tmem_d, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
            else tcgen05_mma_tf32(tmem_d, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tcgen05_commit(smem_u32(&empty_bar[s]));        // frees the ring slot when these MMAs retire
        }
        tcgen05_commit(smem_u32(&tmem_full_bar));         // accumulator complete
      }
    } else {
      // ===== epilogue warps: TMEM lane quadrant = warp_id % 4 =====
      // tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns per chunk).  Bias / ReLU /
      // dropout are applied in registers, the 32 x 32 chunk is written to a 128B-swizzled staging tile in the
      // (now idle) operand ring with conflict-free 128-bit stores, and one lane hands the tile to the TMA
      // unit: a tensor store for plain outputs, a tensor reduce-add (performed in L2) for `+=` and split-K
      // outputs.  TMA clips rows >= i_len and columns >= j_len, so partial tiles need no predicates.  Two
      // staging tiles per warp keep one store in flight while the next chunk is prepared.
      const int q = warp & 3;
      const int il_w = il0 + q * 32;                      // first row of this warp inside the block
      const int i_len = P.i_len, j_len = P.j_len, BJ = P.BJ, Jtot = P.J, act = P.act, epi = P.epi;
      const int64_t row_c = (int64_t)is * i_len + il_w + lane;   // compact row (dropout index)
      const int cj = P.c_jseg[js], ci = P.c_iseg[is];
      const float* brow = (P.bias && split == 0) ? P.bias + (int64_t)P.bias_seg[js] * j_len : nullptr;   // split-K: bias once
      const DropCtx dc = make_drop(P.rng, P.p);
      const bool rows_any = il_w < i_len;                 // warp-uniform
      // staging tiles per warp: a quarter of the operand ring, 4 KB each (2..8)
      int nstage = (int)(((uint32_t)TC_STAGES * TC_STAGE_BYTES) >> 14);
      nstage = nstage > 8 ? 8 : nstage;
      const uint32_t stage0 = smem_base + (uint32_t)q * ((uint32_t)nstage << 12);
      const uint32_t my_row = (uint32_t)lane * 128u;
      const uint32_t sw = (uint32_t)(lane & 7);
      if (brow) {
        for (int c = threadIdx.x - 64; c < BJ; c += 128) bias_s[c] = jl0 + c < j_len ? __ldg(brow + jl0 + c) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");    // the four epilogue warps only
      }
      mbar_wait(smem_u32(&tmem_full_bar), 0);
      if (threadIdx.x == 64) TRACE(2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      int nbuf = 0;
      // one 32-column group of my accumulator row: TMEM -> registers, + bias, ReLU + dropout
      auto load32 = [&](int c0, uint32_t (&v)[32]) {
        tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        const int jb = jl0 + c0;
        if (brow) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {               // warp-wide broadcast reads
            const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[c0 + c]);
            v[c] = __float_as_uint(__uint_as_float(v[c]) + b4.x); v[c + 1] = __float_as_uint(__uint_as_float(v[c + 1]) + b4.y);
            v[c + 2] = __float_as_uint(__uint_as_float(v[c + 2]) + b4.z); v[c + 3] = __float_as_uint(__uint_as_float(v[c + 3]) + b4.w);
          }
        }
        if (act == 1) {                                   // ReLU + dropout (Philox groups of 4 consecutive columns)
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const int j = jb + c;
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = fmaxf(__uint_as_float(v[c + e]), 0.f);
            if (dc.on && j < j_len) {
              const uint64_t idx = (uint64_t)row_c * (uint64_t)Jtot + (uint64_t)((int64_t)js * j_len + j);
              if ((idx & 3) == 0) {
                const uint4 r = drop_rand4(dc, idx >> 2);
                o[0] = r.x >= dc.thr ? o[0] * dc.inv_keep : 0.f; o[1] = r.y >= dc.thr ? o[1] * dc.inv_keep : 0.f;
                o[2] = r.z >= dc.thr ? o[2] * dc.inv_keep : 0.f; o[3] = r.w >= dc.thr ? o[3] * dc.inv_keep : 0.f;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = drop_keep1(dc, idx + e) ? o[e] * dc.inv_keep : 0.f;
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) v[c + e] = __float_as_uint(o[e]);
          }
        }
      };
      // the store issued `nstage` chunks ago must have finished reading its staging tile before it is overwritten
      auto wait_stage = [&]() {
        if (nbuf >= nstage) {
          if (lane == 0) {
            switch (nstage) {                             // wait_group takes an immediate: allow nstage - 1 stores in flight
              case 2: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
              case 3: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
              case 4: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
              case 5: asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); break;
              case 6: asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory"); break;
              case 7: asm volatile("cp.async.bulk.wait_group.read 6;" ::: "memory"); break;
              default: asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); break;
            }
          }
          __syncwarp();
        }
      };
      auto tma_out = [&](uint32_t st, int jb) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (epi == 0)
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(&P.mapC), "r"(st), "r"(jb), "r"(cj), "r"(il_w), "r"(ci) : "memory");
          else
            asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(&P.mapC), "r"(st), "r"(jb), "r"(cj), "r"(il_w), "r"(ci) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      };
      if (P.c16) {
        // bf16 output: two 32-column groups are converted (round-to-nearest-even) and packed into ONE 128-byte staging
        // row = a 64-column x 32-row TMA box of the bf16 output map
        for (int c0 = 0; c0 < BJ; c0 += 64) {
          if (jl0 + c0 >= j_len || !rows_any) break;      // warp-uniform
          const uint32_t st = stage0 + (uint32_t)(nbuf % nstage) * 4096u;
          wait_stage();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            load32(c0 + hh * 32, v);
#pragma unroll
            for (int c = 0; c < 4; ++c) {                 // 16-byte chunk (4 hh + c) of row r lives at chunk ((4 hh + c) ^ (r & 7))
              const uint32_t w0 = pack_bf16x2(__uint_as_float(v[8 * c]), __uint_as_float(v[8 * c + 1]));
              const uint32_t w1 = pack_bf16x2(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3]));
              const uint32_t w2 = pack_bf16x2(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5]));
              const uint32_t w3 = pack_bf16x2(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7]));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + my_row + ((((uint32_t)(4 * hh + c)) ^ sw) << 4)),
                           "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
            }
          }
          tma_out(st, jl0 + c0);
          ++nbuf;
        }
      } else {
        for (int c0 = 0; c0 < BJ; c0 += 32) {
          if (jl0 + c0 >= j_len || !rows_any) break;      // warp-uniform
          uint32_t v[32];
          if (threadIdx.x == 64 && nbuf < 3) TRACE_FINE(40 + nbuf * 6);
          load32(c0, v);
          if (threadIdx.x == 64 && nbuf < 3) TRACE_FINE(42 + nbuf * 6);
          const uint32_t st = stage0 + (uint32_t)(nbuf % nstage) * 4096u;
          wait_stage();
#pragma unroll
          for (int c = 0; c < 8; ++c)                     // 16-byte chunk c of row r lives at chunk (c ^ (r & 7)): SWIZZLE_128B
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + my_row + (((uint32_t)c ^ sw) << 4)),
                         "r"(v[4 * c]), "r"(v[4 * c + 1]), "r"(v[4 * c + 2]), "r"(v[4 * c + 3]) : "memory");
          if (threadIdx.x == 64 && nbuf < 3) TRACE_FINE(43 + nbuf * 6);
          tma_out(st, jl0 + c0);
          if (threadIdx.x == 64 && nbuf < 3) TRACE_FINE(45 + nbuf * 6);
          ++nbuf;
        }
      }
      if (threadIdx.x == 64) TRACE(5);
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
  }
  if (threadIdx.x == 64) TRACE(3);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) TRACE(4);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TC_TMEM_COLS) : "memory");
  }
}

#ifdef MTB_TC_TRACE
extern "C" int mtb_debug_tc_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(unsigned long long) * 64);
}
#endif

// ------------------------------------------------------------------ small helpers (backward)
// scratch = dY * [Y > 0] * inv_keep  (ReLU + dropout backward applied once, feeds dgrad and wgrad) and, fused,
// db[phys(n)] += sum_m scratch[m, n]: one thread per column, 64 rows per block, coalesced along n.
struct ActgradArgs {
  const void* dY; int64_t ldy; const void* Y; int64_t ldyy; void* out; float* db;
  int M, N, seg_len; float inv_keep; uint8_t seg[TC_MAXSEG];
};
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
constexpr int ACT_ROWS = 32;           // rows per block: 4 row-lanes x 8 rows, all 8 loads of a lane in flight
template <typename T>
__global__ void __launch_bounds__(512) actgrad_kernel(const ActgradArgs a) {
  pdl_sync();
  const T* dYp = reinterpret_cast<const T*>(a.dY);
  const T* Yp = reinterpret_cast<const T*>(a.Y);
  T* outp = reinterpret_cast<T*>(a.out);
  __shared__ float part[4][128];
  const int n = blockIdx.x * 128 + threadIdx.x;
  const int m0 = blockIdx.y * ACT_ROWS + threadIdx.y * 8;
  float s = 0.f;
  if (n < a.N) {
    float dy[8], y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int m = m0 + u;
      const bool ok = m < a.M;
      dy[u] = ok ? ldf(dYp + (int64_t)m * a.ldy + n) : 0.f;
      y[u] = ok ? ldf(Yp + (int64_t)m * a.ldyy + n) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int m = m0 + u;
      float v = y[u] > 0.f ? dy[u] * a.inv_keep : 0.f;
      if (sizeof(T) == 2) v = __bfloat162float(__float2bfloat16_rn(v));      // the bias gradient sums what the GEMMs will read
      if (m < a.M) stf(outp + (int64_t)m * a.N + n, v);
      s += v;
    }
  }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < a.N && a.db) {
    const float t = (part[0][threadIdx.x] + part[1][threadIdx.x]) + (part[2][threadIdx.x] + part[3][threadIdx.x]);
    const int sgi = n / a.seg_len;
    atomicAdd(a.db + (int64_t)a.seg[sgi] * a.seg_len + (n - sgi * a.seg_len), t);
  }
}
// db[phys(n)] += sum_m dY[m, n]; one thread per column, 64 rows per block, 8 independent loads in flight
constexpr int COLSUM_ROWS = 64;
struct ColsumArgs {
  const void* dY; int64_t ldy; float* db; int M, N, seg_len; int bf16; uint8_t seg[TC_MAXSEG];
};
struct ColsumGroup {
  ColsumArgs a[MTB_MAX_GROUP];
  int n;
};
// one launch for every bias gradient of a backward call: blockIdx.z selects the problem, the grid covers the
// largest one (surplus blocks exit at once)
__global__ void __launch_bounds__(128) colsum_kernel(const __grid_constant__ ColsumGroup g) {
  pdl_sync();
  const ColsumArgs& a = g.a[blockIdx.z];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m0 = blockIdx.y * COLSUM_ROWS, m1 = min(a.M, m0 + COLSUM_ROWS);
  if (n >= a.N || m0 >= a.M) return;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int m = m0;
  if (a.bf16) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(a.dY) + n;
    for (; m + 8 <= m1; m += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] += __bfloat162float(p[(int64_t)(m + u) * a.ldy]);
    }
    for (; m < m1; ++m) s[0] += __bfloat162float(p[(int64_t)m * a.ldy]);
  } else {
    const float* p = reinterpret_cast<const float*>(a.dY) + n;
    for (; m + 8 <= m1; m += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] += p[(int64_t)(m + u) * a.ldy];
    }
    for (; m < m1; ++m) s[0] += p[(int64_t)m * a.ldy];
  }
  const float t = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  const int sgi = n / a.seg_len;
  atomicAdd(a.db + (int64_t)a.seg[sgi] * a.seg_len + (n - sgi * a.seg_len), t);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Row-major fp32 matrix with leading dimension ld seen as 4-D {col_in_block, col_block, row_in_block,
// row_block}: column blocks of clen (ncb of them), row blocks of rlen (nrb of them); box = bx cols x by rows.
// esz = bytes per element: 4 (fp32, consumed as tf32) or 2 (bf16).  32-bit MN-major operand tiles need the 32-byte-atom
// variant of the 128-byte swizzle, everything else the plain one.
static bool encode_map(CUtensorMap* m, const void* ptr, int64_t ld, int64_t clen, int64_t ncb, int64_t rlen, int64_t nrb,
                       int bx, int by, bool mn_major, int esz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  if (ncb > 1 && ((clen * esz) % 16) != 0) return false;
  cuuint64_t dims[4] = {(cuuint64_t)clen, (cuuint64_t)ncb, (cuuint64_t)rlen, (cuuint64_t)nrb};
  cuuint64_t strides[3] = {(cuuint64_t)(ncb > 1 ? clen * esz : ld * esz), (cuuint64_t)ld * esz, (cuuint64_t)rlen * ld * esz};
  cuuint32_t box[4] = {(cuuint32_t)bx, 1, (cuuint32_t)by, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, (mn_major && esz == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// A tensor map is a pure function of (address, shape, strides, box, swizzle).  The plan executor launches the same
// operands step after step (persistent arenas), and cuTensorMapEncodeTiled costs ~0.4 us x 3 maps x every problem of
// every GEMM launch -- on the critical path, because the GPU finishes these small kernels faster than the host issues
// them.  Direct-mapped cache of encoded maps (single host thread per device context, like the rest of the library).
struct MapKey {
  const void* ptr; int64_t ld, clen, ncb, rlen, nrb; int32_t bx, by, mn, esz;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && ld == o.ld && clen == o.clen && ncb == o.ncb && rlen == o.rlen && nrb == o.nrb && bx == o.bx && by == o.by &&
           mn == o.mn && esz == o.esz;
  }
};
struct MapEntry { MapKey key; CUtensorMap map; bool valid; };
constexpr int MAP_CACHE = 1 << 14;
static MapEntry* g_map_cache = nullptr;

static bool make_map(CUtensorMap* m, const void* ptr, int64_t ld, int64_t clen, int64_t ncb, int64_t rlen, int64_t nrb,
                     int bx, int by, bool mn_major, int esz = 4) {
  if (!g_map_cache) g_map_cache = (MapEntry*)calloc(MAP_CACHE, sizeof(MapEntry));
  MapKey k{ptr, ld, clen, ncb, rlen, nrb, bx, by, mn_major ? 1 : 0, esz};
  uint64_t h = (uint64_t)(uintptr_t)ptr * 0x9E3779B97F4A7C15ull;
  h ^= ((uint64_t)ld * 0xC2B2AE3D27D4EB4Full) ^ ((uint64_t)clen << 17) ^ ((uint64_t)rlen << 29) ^ ((uint64_t)ncb << 7) ^ ((uint64_t)nrb << 11) ^
       ((uint64_t)bx << 41) ^ ((uint64_t)by << 47) ^ ((uint64_t)k.mn << 53) ^ ((uint64_t)esz << 57);
  h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  MapEntry* e = g_map_cache ? &g_map_cache[h & (MAP_CACHE - 1)] : nullptr;
  if (e && e->valid && e->key == k) { *m = e->map; return true; }
  if (!encode_map(m, ptr, ld, clen, ncb, rlen, nrb, bx, by, mn_major, esz)) return false;
  if (e) { e->key = k; e->map = *m; e->valid = true; }
  return true;
}

static bool tma_ok(const void* p, int64_t ld, int esz = 4) { return p != nullptr && ((((uintptr_t)p) & 15) == 0) && ((ld * esz) % 16 == 0) && ld > 0; }

// Tile width along J: a multiple of 32 (the epilogue moves 32-column TMA boxes), chosen to minimise
// tiles x (width + 128) -- the "+ 128" is the A tile every extra column tile has to load again.
static int pick_bj(int J, int gran = 32) {
  int best = 64, best_cost = 1 << 30;
  for (int c = 256; c >= gran; c -= gran) {
    const int tiles = (J + c - 1) / c;
    const int cost = tiles * (c + 128);
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}
// Launches that cannot fill the GPU with full-width tiles (<= one CTA per SM) are pure latency: halve the tile
// width so twice as many CTAs each run a shorter main loop (3-stage ring at <= 128 columns) and epilogue.
static int pick_bj_narrow(int J, int gran = 32) {
  const int bj = pick_bj(J, gran);
  if (J < 128) return bj;
  const int tiles = (J + bj - 1) / bj;
  int nb = ((J + 2 * tiles - 1) / (2 * tiles) + gran - 1) / gran * gran;
  return nb < 64 ? 64 : nb;
}
static int tiles_of(int I, int i_n, int J, int j_n, int bj) { return ((I + TC_BI - 1) / TC_BI) * i_n * ((J + bj - 1) / bj) * j_n; }
// split of a long reduction across CTAs when a problem has too few tiles to matter (reduce-add epilogue)
static int pick_splitk(int tiles, int nkb) {
  if (tiles > 16 || nkb < 24) return 1;
  int s = nkb / 6;
  const int cap = sm_count() / tiles;
  if (s > cap) s = cap;
  return s < 1 ? 1 : s;
}
static int zero_matrix(float* p, int64_t ld, int rows, int cols, cudaStream_t st) {
  if (ld == cols) { MTB_CUDA(cudaMemsetAsync(p, 0, (size_t)rows * cols * sizeof(float), st)); }
  else { MTB_CUDA(cudaMemset2DAsync(p, (size_t)ld * sizeof(float), 0, (size_t)cols * sizeof(float), (size_t)rows, st)); }
  return 0;
}
static int round_bj_mn(int bj, int mnw = 32) { return ((bj + mnw - 1) / mnw) * mnw; }   // MN-major B tiles are loaded in 32- (tf32) / 64- (bf16) wide boxes

// axis description derived from the public descriptor
struct Axis {
  int len, n;                       // block length, number of compact blocks
  uint8_t phys[TC_MAXSEG];          // physical block index of compact block s
  int nphys;                        // physical blocks addressable (max + 1)
};
static bool make_axis(Axis& a, int total, const int32_t* idx, const mtb_segs& sg, int esz = 4) {
  if (idx == nullptr) {
    a.len = total; a.n = 1; a.phys[0] = 0; a.nphys = 1;
    return true;
  }
  if (sg.n <= 0 || sg.n > TC_MAXSEG || sg.len <= 0 || sg.len * sg.n != total || ((sg.len * esz) % 16) != 0) return false;
  a.len = sg.len; a.n = sg.n; a.nphys = 0;
  for (int s = 0; s < sg.n; ++s) {
    if (sg.seg[s] < 0 || sg.seg[s] > 255) return false;
    a.phys[s] = (uint8_t)sg.seg[s];
    if (sg.seg[s] + 1 > a.nphys) a.nphys = sg.seg[s] + 1;
  }
  return true;
}
static void ident(uint8_t* d, int n) { for (int s = 0; s < TC_MAXSEG; ++s) d[s] = (uint8_t)(s < n ? s : 0); }
static void copy_phys(uint8_t* d, const Axis& a) { for (int s = 0; s < TC_MAXSEG; ++s) d[s] = s < a.n ? a.phys[s] : 0; }

template <int CAP>
static int launch_tc_cap(const TcProblem* probs, int n, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    MTB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET + 2048));
    attr_done = true;
  }
  TcGroupT<CAP> g;
  g.n = n;
  int tot = 0;
  size_t smem = 0;
  for (int i = 0; i < n; ++i) {
    TcProblem& q = g.d[i];
    q = probs[i];
    const int mnw = q.ab16 ? 64 : 32;
    const size_t stage = (size_t)TC_A_BYTES + (((size_t)(q.b_mn ? ((q.BJ + mnw - 1) / mnw) * mnw : q.BJ) * 128 + 1023) & ~(size_t)1023);
    int stages = (int)(TC_SMEM_BUDGET / stage);
    stages = stages > TC_MAX_STAGES ? TC_MAX_STAGES : (stages < 2 ? 2 : stages);
    q.stages = stages;
    if (stage * stages + 1024 > smem) smem = stage * stages + 1024;
    g.start[i] = tot;
    tot += ((q.i_len + TC_BI - 1) / TC_BI) * q.i_nseg * ((q.j_len + q.BJ - 1) / q.BJ) * q.j_nseg * q.splits;
  }
  g.start[n] = tot;
  if (tot == 0) return 0;
  static const bool dbg = getenv("MTB_TC_DEBUG") != nullptr;
  if (dbg) {
    fprintf(stderr, "[tc] grid %d:", tot);
    for (int i = 0; i < n; ++i)
      fprintf(stderr, " {I %d J %d R %d BJ %d st %d sp %d mn %d%d epi %d ab16 %d c16 %d}", g.d[i].I, g.d[i].J, g.d[i].R, g.d[i].BJ, g.d[i].stages,
              g.d[i].splits, g.d[i].a_mn, g.d[i].b_mn, g.d[i].epi, g.d[i].ab16, g.d[i].c16);
    fprintf(stderr, "\n");
  }
  MTB_CUDA(launch_k(gemm_tc_kernel<CAP>, dim3(tot), dim3(TC_THREADS), smem, st, g));
  mtb::note_launch();
  MTB_CUDA(cudaGetLastError());
  return 0;
}
static int launch_tc(const TcProblem* probs, int n, cudaStream_t st) {
  if (n == 0) return 0;
  if (n <= 2) return launch_tc_cap<2>(probs, n, st);          // 1.3 KB of parameters
  if (n <= 6) return launch_tc_cap<6>(probs, n, st);          // 3.9 KB: still on the fast launch path
  if (n <= 12) return launch_tc_cap<12>(probs, n, st);
  return launch_tc_cap<MTB_MAX_GROUP>(probs, n, st);
}

int linear_fwd_tc(const mtb_linear_desc* d, int n, cudaStream_t st) {
  TcProblem tc[MTB_MAX_GROUP];
  mtb_linear_desc rest[MTB_MAX_GROUP];
  Axis ans[MTB_MAX_GROUP], aks[MTB_MAX_GROUP];
  bool oks[MTB_MAX_GROUP];
  int ntc = 0, nrest = 0, full_tiles = 0;
  for (int i = 0; i < n; ++i) {
    const mtb_linear_desc& x = d[i];
    const int ei = x.in_bf16 ? 2 : 4, eo = x.out_bf16 ? 2 : 4;
    oks[i] = x.N >= 16 && x.K >= 8 && x.M >= 1 && tma_ok(x.X, x.ldx, ei) && tma_ok(x.W, x.ldw, ei) && tma_ok(x.Y, x.ldy, eo) &&
             make_axis(ans[i], x.N, x.row_idx, x.row_segs, eo < ei ? eo : ei) && make_axis(aks[i], x.K, x.col_idx, x.col_segs, ei);
    MTB_CHECK(oks[i] || !(x.in_bf16 || x.out_bf16),
              "linear_fwd: problem %d has bf16 operands the TMA engine cannot address (M %d N %d K %d ldx %lld ldw %lld ldy %lld); there is no "
              "bf16 fallback", i, x.M, x.N, x.K, (long long)x.ldx, (long long)x.ldw, (long long)x.ldy);
    if (oks[i]) full_tiles += tiles_of(x.M, 1, ans[i].len, ans[i].n, pick_bj(ans[i].len, x.out_bf16 ? 64 : 32));
  }
  const bool narrow = full_tiles <= sm_count();
  for (int i = 0; i < n; ++i) {
    const mtb_linear_desc& x = d[i];
    const Axis& an = ans[i];
    const Axis& ak = aks[i];
    const int ei = x.in_bf16 ? 2 : 4, eo = x.out_bf16 ? 2 : 4;
    const int br = x.in_bf16 ? 64 : TC_BR;
    const int gran = x.out_bf16 ? 64 : 32;
    bool ok = oks[i];
    TcProblem& q = tc[ntc];
    if (ok) {
      q = TcProblem{};
      q.BJ = narrow ? pick_bj_narrow(an.len, gran) : pick_bj(an.len, gran);
      ok = make_map(&q.mapA, x.X, x.ldx, ak.len, ak.n, x.M, 1, br, TC_BI, false, ei) &&
           make_map(&q.mapB, x.W, x.ldw, ak.len, ak.nphys, an.len, an.nphys, br, q.BJ, false, ei) &&
           make_map(&q.mapC, x.Y, x.ldy, an.len, an.n, x.M, 1, x.out_bf16 ? 64 : 32, 32, false, eo);
      MTB_CHECK(ok || !(x.in_bf16 || x.out_bf16), "linear_fwd: tensor-map encoding failed for bf16 problem %d", i);
    }
    if (!ok) { rest[nrest++] = x; continue; }
    q.C = (float*)x.Y; q.ldc = x.ldy; q.bias = x.bias;
    q.I = x.M; q.J = x.N; q.R = x.K;
    q.i_len = x.M; q.i_nseg = 1; q.j_len = an.len; q.j_nseg = an.n; q.r_len = ak.len; q.r_nseg = ak.n;
    ident(q.a_iseg, 1); ident(q.a_rseg, ak.n); copy_phys(q.b_jseg, an); copy_phys(q.b_rseg, ak);
    ident(q.c_iseg, 1); ident(q.c_jseg, an.n); copy_phys(q.bias_seg, an);
    q.a_mn = 0; q.b_mn = 0; q.splits = 1; q.epi = 0;
    q.act = x.act; q.p = x.p; q.rng = x.rng;
    q.ab16 = x.in_bf16 ? 1 : 0; q.c16 = x.out_bf16 ? 1 : 0;
    if (x.act == 0 && !x.out_bf16) {     // few tiles, long reduction (the head's [B, 3000] inputs): split K, partial sums meet in L2 (fp32 outputs only)
      const int nkb = ((ak.len + br - 1) / br) * ak.n;
      const int sp = pick_splitk(tiles_of(x.M, 1, an.len, an.n, q.BJ), nkb);
      if (sp > 1) {
        q.splits = sp; q.epi = 2;
        if (zero_matrix((float*)x.Y, x.ldy, x.M, x.N, st)) return -2;
      }
    } else if (x.out_bf16 && tiles_of(x.M, 1, an.len, an.n, q.BJ) * 8 <= sm_count()) {
      q.BJ = 64;                         // bf16 outputs cannot be split over K: spread a long skinny problem over more column tiles instead
      ok = make_map(&q.mapB, x.W, x.ldw, ak.len, ak.nphys, an.len, an.nphys, br, q.BJ, false, ei);
      MTB_CHECK(ok, "linear_fwd: tensor-map encoding failed for bf16 problem %d", i);
    }
    ++ntc;
  }
  int rc = launch_tc(tc, ntc, st);
  if (rc) return rc;
  if (nrest) return linear_fwd_simt(rest, nrest, st);
  return 0;
}

int linear_bwd_tc(const mtb_linear_bwd_desc* d, int n, cudaStream_t st) {
  TcProblem dg[MTB_MAX_GROUP], wg[MTB_MAX_GROUP];
  mtb_linear_bwd_desc rest[MTB_MAX_GROUP];
  ColsumGroup cs;
  cs.n = 0;
  int cs_gx = 0, cs_gy = 0;
  int ndg = 0, nwg = 0, nrest = 0;
  for (int i = 0; i < n; ++i) {
    const mtb_linear_bwd_desc& x = d[i];
    Axis an, ak;
    const int ei = x.in_bf16 ? 2 : 4, ex = x.dx_bf16 ? 2 : 4;
    const int br = x.in_bf16 ? 64 : TC_BR, mnw = x.in_bf16 ? 64 : 32;
    bool ok = x.N >= 16 && x.K >= 16 && x.M >= 1 && tma_ok(x.dY, x.ldy, ei) && tma_ok(x.W, x.ldw, ei) &&
              (!x.dX || tma_ok(x.dX, x.lddx, ex)) && (!x.dW || tma_ok(x.X, x.ldx, ei)) && (x.act == 0 || x.scratch != nullptr) &&
              make_axis(an, x.N, x.row_idx, x.row_segs, ei) && make_axis(ak, x.K, x.col_idx, x.col_segs, ei < ex ? ei : ex);
    MTB_CHECK(ok || !(x.in_bf16 || x.dx_bf16), "linear_bwd: problem %d has bf16 operands the TMA engine cannot address (M %d N %d K %d); there is "
              "no bf16 fallback", i, x.M, x.N, x.K);
    if (!ok) {
      // fp32-engine fallback (operands TMA cannot address, e.g. a gather without block structure).  Callers that split
      // dgrad and wgrad into separate calls rely on `scratch` holding dY' = dY * [Y > 0] / (1 - p) afterwards (the
      // deferred weight-gradient call reads it as its dY): materialise it here exactly like the tensor-core path does,
      // and hand the fp32 engine a plain (act = 0) problem over it.
      mtb_linear_bwd_desc y = x;
      if (x.act == 1 && x.scratch != nullptr) {
        Axis arow;
        MTB_CHECK(make_axis(arow, x.N, x.row_idx, x.row_segs, 4), "linear_bwd: problem %d needs ReLU backward over a row gather without "
                  "block structure, which neither engine supports", i);
        ActgradArgs aa{};
        aa.dY = x.dY; aa.ldy = x.ldy; aa.Y = x.Yact; aa.ldyy = x.ldyact; aa.out = x.scratch; aa.db = x.db;
        aa.M = x.M; aa.N = x.N; aa.seg_len = arow.len; aa.inv_keep = x.p > 0.f ? 1.f / (1.f - x.p) : 1.f;
        for (int s2 = 0; s2 < TC_MAXSEG; ++s2) aa.seg[s2] = s2 < arow.n ? arow.phys[s2] : 0;
        dim3 grid((x.N + 127) / 128, (x.M + ACT_ROWS - 1) / ACT_ROWS);
        MTB_CUDA(launch_k(actgrad_kernel<float>, grid, dim3(128, 4), 0, st, aa));
        mtb::note_launch();
        MTB_CUDA(cudaGetLastError());
        y.dY = x.scratch; y.ldy = x.N; y.act = 0; y.Yact = nullptr; y.ldyact = 0; y.db = nullptr; y.p = 0.f;
      }
      if (y.db != nullptr && y.dW == nullptr) {
        // the fp32 engine only forms a bias gradient inside its weight-gradient GEMM: column-sum it here instead
        Axis arow;
        MTB_CHECK(make_axis(arow, x.N, x.row_idx, x.row_segs, 4), "linear_bwd: problem %d needs a bias gradient over a row gather without "
                  "block structure and no weight gradient, which neither engine supports", i);
        ColsumArgs& ca = cs.a[cs.n++];
        ca = ColsumArgs{};
        ca.dY = y.dY; ca.ldy = y.ldy; ca.db = y.db; ca.M = x.M; ca.N = x.N; ca.seg_len = arow.len; ca.bf16 = 0;
        for (int s = 0; s < TC_MAXSEG; ++s) ca.seg[s] = s < arow.n ? arow.phys[s] : 0;
        cs_gx = cs_gx > (x.N + 127) / 128 ? cs_gx : (x.N + 127) / 128;
        cs_gy = cs_gy > (x.M + COLSUM_ROWS - 1) / COLSUM_ROWS ? cs_gy : (x.M + COLSUM_ROWS - 1) / COLSUM_ROWS;
        y.db = nullptr;
      }
      rest[nrest++] = y;
      continue;
    }
    bool built = true;
    TcProblem qd{}, qw{};
    const void* dYp = x.act == 1 ? (const void*)x.scratch : (const void*)x.dY;
    const int64_t ldyp = x.act == 1 ? x.N : x.ldy;
    if (x.dX) {           // dX[M,K] = dY'[M,N] . W'[N,K] : A K-major (reduction n contiguous), B MN-major (k_out contiguous)
      qd.BJ = round_bj_mn(pick_bj(ak.len, x.dx_bf16 ? 64 : 32), mnw);
      if (qd.BJ > 256) qd.BJ = 256;
      built = built && make_map(&qd.mapA, dYp, ldyp, an.len, an.n, x.M, 1, br, TC_BI, false, ei) &&
              make_map(&qd.mapB, x.W, x.ldw, ak.len, ak.nphys, an.len, an.nphys, mnw, br, true, ei) &&
              make_map(&qd.mapC, x.dX, x.lddx, ak.len, ak.n, x.M, 1, x.dx_bf16 ? 64 : 32, 32, false, ex);
      qd.C = (float*)x.dX; qd.ldc = x.lddx; qd.bias = nullptr;
      qd.I = x.M; qd.J = x.K; qd.R = x.N;
      qd.i_len = x.M; qd.i_nseg = 1; qd.j_len = ak.len; qd.j_nseg = ak.n; qd.r_len = an.len; qd.r_nseg = an.n;
      ident(qd.a_iseg, 1); ident(qd.a_rseg, an.n); copy_phys(qd.b_jseg, ak); copy_phys(qd.b_rseg, an);
      ident(qd.c_iseg, 1); ident(qd.c_jseg, ak.n); ident(qd.bias_seg, 1);
      qd.a_mn = 0; qd.b_mn = 1; qd.splits = 1; qd.epi = x.accumulate_dx ? 1 : 0;
      qd.ab16 = x.in_bf16 ? 1 : 0; qd.c16 = x.dx_bf16 ? 1 : 0;
    }
    if (x.dW) {           // dW'[N,K] += dY'^T[N,M] . X[M,K] : both MN-major, reduction over tokens split across CTAs; fp32 output
      qw.BJ = round_bj_mn(pick_bj(ak.len), mnw);
      if (qw.BJ > 256) qw.BJ = 256;
      built = built && make_map(&qw.mapA, dYp, ldyp, an.len, an.n, x.M, 1, mnw, br, true, ei) &&
              make_map(&qw.mapB, x.X, x.ldx, ak.len, ak.n, x.M, 1, mnw, br, true, ei) &&
              make_map(&qw.mapC, x.dW, x.ldw, ak.len, ak.nphys, an.len, an.nphys, 32, 32, false, 4);
      qw.C = x.dW; qw.ldc = x.ldw; qw.bias = nullptr;
      qw.I = x.N; qw.J = x.K; qw.R = x.M;
      qw.i_len = an.len; qw.i_nseg = an.n; qw.j_len = ak.len; qw.j_nseg = ak.n; qw.r_len = x.M; qw.r_nseg = 1;
      ident(qw.a_iseg, an.n); ident(qw.a_rseg, 1); ident(qw.b_jseg, ak.n); ident(qw.b_rseg, 1);
      copy_phys(qw.c_iseg, an); copy_phys(qw.c_jseg, ak); ident(qw.bias_seg, 1);
      qw.a_mn = 1; qw.b_mn = 1; qw.epi = 2;
      qw.splits = 1;        // decided for the whole launch below
      qw.ab16 = x.in_bf16 ? 1 : 0; qw.c16 = 0;
    }
    MTB_CHECK(built || !(x.in_bf16 || x.dx_bf16), "linear_bwd: tensor-map encoding failed for bf16 problem %d", i);
    if (!built) { rest[nrest++] = x; continue; }
    if (x.act == 1) {       // dY' = dY * [Y > 0] / (1 - p), materialised once for dgrad + wgrad; bias grad fused
      ActgradArgs aa{};
      aa.dY = x.dY; aa.ldy = x.ldy; aa.Y = x.Yact; aa.ldyy = x.ldyact; aa.out = x.scratch; aa.db = x.db;
      aa.M = x.M; aa.N = x.N; aa.seg_len = an.len; aa.inv_keep = x.p > 0.f ? 1.f / (1.f - x.p) : 1.f;
      for (int s2 = 0; s2 < TC_MAXSEG; ++s2) aa.seg[s2] = s2 < an.n ? an.phys[s2] : 0;
      dim3 grid((x.N + 127) / 128, (x.M + ACT_ROWS - 1) / ACT_ROWS);
      if (x.in_bf16) { MTB_CUDA(launch_k(actgrad_kernel<__nv_bfloat16>, grid, dim3(128, 4), 0, st, aa)); }
      else { MTB_CUDA(launch_k(actgrad_kernel<float>, grid, dim3(128, 4), 0, st, aa)); }
      mtb::note_launch();
      MTB_CUDA(cudaGetLastError());
    }
    if (x.dX) dg[ndg++] = qd;
    if (x.dW) wg[nwg++] = qw;
    if (x.db && x.act != 1) {
      ColsumArgs& ca = cs.a[cs.n++];
      ca = ColsumArgs{};
      ca.dY = dYp; ca.ldy = ldyp; ca.db = x.db; ca.M = x.M; ca.N = x.N; ca.seg_len = an.len; ca.bf16 = x.in_bf16 ? 1 : 0;
      for (int s = 0; s < TC_MAXSEG; ++s) ca.seg[s] = s < an.n ? an.phys[s] : 0;
      cs_gx = cs_gx > (x.N + 127) / 128 ? cs_gx : (x.N + 127) / 128;
      cs_gy = cs_gy > (x.M + COLSUM_ROWS - 1) / COLSUM_ROWS ? cs_gy : (x.M + COLSUM_ROWS - 1) / COLSUM_ROWS;
    }
  }
  if (cs.n > 0) {
    MTB_CUDA(launch_k(colsum_kernel, dim3(cs_gx, cs_gy, cs.n), dim3(128), 0, st, cs));
    mtb::note_launch();
    MTB_CUDA(cudaGetLastError());
  }
  // ---- launch-wide tiling policy (B operands are MN-major here: 32- / 64-wide TMA boxes, so BJ is free to change) ----
  auto gran_of = [](const TcProblem& q) { return (q.ab16 || q.c16) ? 64 : 32; };
  auto br_of = [](const TcProblem& q) { return q.ab16 ? 64 : TC_BR; };
  int dg_ctas = 0;
  for (int i = 0; i < ndg; ++i) dg_ctas += tiles_of(dg[i].i_len, dg[i].i_nseg, dg[i].j_len, dg[i].j_nseg, dg[i].BJ);
  if (dg_ctas <= sm_count()) {
    dg_ctas = 0;
    for (int i = 0; i < ndg; ++i) {
      dg[i].BJ = pick_bj_narrow(dg[i].j_len, gran_of(dg[i]));
      dg_ctas += tiles_of(dg[i].i_len, dg[i].i_nseg, dg[i].j_len, dg[i].j_nseg, dg[i].BJ);
    }
  }
  for (int i = 0; i < ndg; ++i) {        // few tiles, long reduction (head): split it; partial sums meet in L2 (fp32 outputs only)
    TcProblem& q = dg[i];
    if (q.c16) {
      if (tiles_of(q.i_len, q.i_nseg, q.j_len, q.j_nseg, q.BJ) * 8 <= sm_count()) q.BJ = 64;
      continue;
    }
    const int nkb = ((q.r_len + br_of(q) - 1) / br_of(q)) * q.r_nseg;
    const int sp = pick_splitk(tiles_of(q.i_len, q.i_nseg, q.j_len, q.j_nseg, q.BJ), nkb);
    if (sp > 1) {
      q.splits = sp;
      if (q.epi == 0) {
        q.epi = 2;
        if (zero_matrix(q.C, q.ldc, q.I, q.J, st)) return -2;
      }
      dg_ctas += (sp - 1) * tiles_of(q.i_len, q.i_nseg, q.j_len, q.j_nseg, q.BJ);
    }
  }
  {
    // weight gradients reduce over the token axis: give every CTA >= 8 reduction slabs (>= 4 of the twice as deep bf16
    // slabs) and aim the whole launch at about one wave (2 CTAs / SM) -- more splits only multiply the reduce-add traffic
    long long work = 0;
    for (int i = 0; i < nwg; ++i)
      work += (long long)tiles_of(wg[i].i_len, wg[i].i_nseg, wg[i].j_len, wg[i].j_nseg, wg[i].BJ) * ((wg[i].r_len + br_of(wg[i]) - 1) / br_of(wg[i]));
    int budget = 2 * sm_count() - (dg_ctas < sm_count() ? dg_ctas : sm_count());
    int kpc = (int)((work + budget - 1) / budget);
    for (int i = 0; i < nwg; ++i) {
      const int kmin = wg[i].ab16 ? 4 : 8;
      const int k = kpc < kmin ? kmin : kpc;
      const int nkb = (wg[i].r_len + br_of(wg[i]) - 1) / br_of(wg[i]);
      int sp = (nkb + k - 1) / k;
      wg[i].splits = sp < 1 ? 1 : sp;
    }
  }
  TcProblem all[2 * MTB_MAX_GROUP];
  int nall = 0;
  for (int i = 0; i < ndg; ++i) all[nall++] = dg[i];
  for (int i = 0; i < nwg; ++i) all[nall++] = wg[i];
  for (int off = 0; off < nall; off += MTB_MAX_GROUP) {
    const int m = nall - off < MTB_MAX_GROUP ? nall - off : MTB_MAX_GROUP;
    int rc = launch_tc(all + off, m, st);
    if (rc) return rc;
  }
  if (nrest) return linear_bwd_simt(rest, nrest, st);
  return 0;
}

}  // namespace mtb

namespace mtb {
int preload_linear_tc() {
  int bad = 0;
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, gemm_tc_kernel<2>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, gemm_tc_kernel<6>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, gemm_tc_kernel<12>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, gemm_tc_kernel<MTB_MAX_GROUP>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, actgrad_kernel<float>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, actgrad_kernel<__nv_bfloat16>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, colsum_kernel) != cudaSuccess) ++bad; }
  return bad;
}
}  // namespace mtb
