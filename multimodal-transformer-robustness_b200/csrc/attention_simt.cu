// attention_simt.cu -- fused flash-style attention core in fp32 on CUDA cores (parity mode):
// QK^T, causal-with-offset predicate, online fp32 softmax, Philox attention dropout, PV,
// forward and backward, ragged (Lq, Lk) per problem, head_dim padded to 32/64 in SHARED
// MEMORY only.  modules/dynamic_multihead_attention.py:91-116, modules/transformer.py:145-157.
// Scores/probabilities never touch HBM; algorithmic HBM bytes = q + k + v + o (+ lse).
#include "common.cuh"
#include <math_constants.h>

namespace mtb {

extern int g_gemm_mode;
// exp in the softmax: accurate expf in fp32 parity mode, ex2.approx-based __expf when the run is in
// reduced-precision (tensor-core) mode anyway
template <bool FAST> __device__ __forceinline__ float attn_exp(float x) { return FAST ? __expf(x) : expf(x); }

constexpr int AT = 64;          // q-tile and k-tile edge
constexpr int AP = AT + 4;      // padded smem row (floats), keeps float4 alignment
constexpr int A_THREADS = 256;

// C[4][4] += sum_d A[d][ty*4+u] * B[d][tx*4+v]   (A, B stored [HDP][AP])
template <int HDP>
__device__ __forceinline__ void mini_gemm(const float* __restrict__ A, const float* __restrict__ B, int hd,
                                          int ty, int tx, float (&c)[4][4]) {
#pragma unroll 5
  for (int d = 0; d < hd; ++d) {
    const float4 a = *reinterpret_cast<const float4*>(A + d * AP + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(B + d * AP + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) c[u][v] = fmaf(av[u], bv[v], c[u][v]);
  }
}

// O[4][HDP/16] += sum_j P[ty*4+u][j] * V[j][tx*(HDP/16)+w]   (P stored [AT][AP], V stored [AT][HDP+2])
template <int HDP>
__device__ __forceinline__ void mini_pv(const float* __restrict__ P, const float* __restrict__ V, int jn,
                                        int ty, int tx, float (&o)[4][HDP / 16]) {
  constexpr int W = HDP / 16;
  constexpr int VP = HDP + 2;
  for (int j = 0; j < jn; j += 4) {
    float pv[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 t = *reinterpret_cast<const float4*>(P + (ty * 4 + u) * AP + j);
      pv[u][0] = t.x; pv[u][1] = t.y; pv[u][2] = t.z; pv[u][3] = t.w;
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float vv[W];
#pragma unroll
      for (int w = 0; w < W; ++w) vv[w] = V[(j + jj) * VP + tx * W + w];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int w = 0; w < W; ++w) o[u][w] = fmaf(pv[u][jj], vv[w], o[u][w]);
    }
  }
}

// rows of a token-major matrix for fixed (b, h): element (l, d) at base[(l*B + b)*ld + h*hd + d]
// -> transposed tile T[d][l - l0] (zero padded)
template <int HDP>
__device__ __forceinline__ void load_tile_T(float* T, const float* base, int64_t ld, int B, int b, int h, int hd,
                                            int l0, int L) {
  // 4-byte cp.async with zero-fill: every copy of the tile is in flight at once (one L2 latency per
  // tile) instead of a dependent load -> store chain per element; completed by tile_wait()
  const uint32_t tb = (uint32_t)__cvta_generic_to_shared(T);
  for (int e = threadIdx.x; e < AT * HDP; e += A_THREADS) {
    const int r = e / HDP, dd = e - r * HDP;
    const int l = l0 + r;
    const bool ok = (l < L) && (dd < hd);
    const float* src = ok ? base + ((int64_t)l * B + b) * ld + h * hd + dd : base;
    const int nbytes = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(tb + (uint32_t)(dd * AP + r) * 4u), "l"(src), "r"(nbytes) : "memory");
  }
}
__device__ __forceinline__ void tile_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
// -> natural tile N[l - l0][d] (zero padded), row stride HDP+2
template <int HDP>
__device__ __forceinline__ void load_tile_N(float* N, const float* base, int64_t ld, int B, int b, int h, int hd,
                                            int l0, int L) {
  const uint32_t tb = (uint32_t)__cvta_generic_to_shared(N);
  for (int e = threadIdx.x; e < AT * HDP; e += A_THREADS) {
    const int r = e / HDP, dd = e - r * HDP;
    const int l = l0 + r;
    const bool ok = (l < L) && (dd < hd);
    const float* src = ok ? base + ((int64_t)l * B + b) * ld + h * hd + dd : base;
    const int nbytes = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(tb + (uint32_t)(r * (HDP + 2) + dd) * 4u), "l"(src), "r"(nbytes) : "memory");
  }
}

__device__ __forceinline__ int round4(int x) { return (x + 3) & ~3; }

// keep-factors (inv_keep or 0) for elements (row, j..j+3) of head `bh`
__device__ __forceinline__ void attn_keep4(const DropCtx& dc, int64_t bh, int Lq, int Lk4, int i, int j, float (&k)[4]) {
  k[0] = k[1] = k[2] = k[3] = dc.inv_keep;
  if (dc.on) {
    const uint64_t idx = ((uint64_t)(bh * Lq + i)) * (uint64_t)Lk4 + (uint64_t)j;
    const uint4 r = drop_rand4(dc, idx >> 2);
    k[0] = r.x >= dc.thr ? dc.inv_keep : 0.f; k[1] = r.y >= dc.thr ? dc.inv_keep : 0.f;
    k[2] = r.z >= dc.thr ? dc.inv_keep : 0.f; k[3] = r.w >= dc.thr ? dc.inv_keep : 0.f;
  }
}

// ------------------------------------------------------------------------------ forward
template <int HDP, bool FAST>
__global__ void __launch_bounds__(A_THREADS) attn_fwd_kernel(const __grid_constant__ Group<mtb_attn_desc> g) {
  extern __shared__ __align__(16) float smem[];
  float* Qt = smem;                    // [HDP][AP]
  float* Kt = Qt + HDP * AP;           // [HDP][AP]
  float* Vn = Kt + HDP * AP;           // [AT][HDP+2]
  float* Ps = Vn + AT * (HDP + 2);     // [AT][AP]
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_attn_desc& d = g.d[pi];
  const int qtiles = (d.Lq + AT - 1) / AT;
  const int bh = local / qtiles, qt = local - bh * qtiles;
  const int b = bh / d.H, h = bh - b * d.H;
  const int i0 = qt * AT;
  const int off = abs(d.Lk - d.Lq);
  const int Lk4 = round4(d.Lk);
  const DropCtx dc = make_drop(d.rng, d.p);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  constexpr int W = HDP / 16;

  load_tile_T<HDP>(Qt, d.q, d.ldq, d.B, b, h, d.hd, i0, d.Lq);
  float m[4], l[4], o[4][W];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    m[u] = -CUDART_INF_F; l[u] = 0.f;
#pragma unroll
    for (int w = 0; w < W; ++w) o[u][w] = 0.f;
  }
  const int i_last = min(d.Lq, i0 + AT) - 1;
  const int j_end = min(d.Lk, i_last + off + 1);      // first fully masked column for this q tile
  for (int j0 = 0; j0 < j_end; j0 += AT) {
    __syncthreads();                                   // previous tile's Ps / Vn / Kt consumed
    load_tile_T<HDP>(Kt, d.k, d.ldk, d.B, b, h, d.hd, j0, d.Lk);
    load_tile_N<HDP>(Vn, d.v, d.ldv, d.B, b, h, d.hd, j0, d.Lk);
    tile_wait();
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) s[u][v] = 0.f;
    mini_gemm<HDP>(Qt, Kt, d.hd, ty, tx, s);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + ty * 4 + u;
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int j = j0 + tx * 4 + v;
        const bool open = (j < d.Lk) && (j - i < 1 + off);
        s[u][v] = open ? s[u][v] * d.scale : -CUDART_INF_F;
        mx = fmaxf(mx, s[u][v]);
      }
#pragma unroll
      for (int sh = 8; sh > 0; sh >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
      const float m_new = fmaxf(m[u], mx);
      const float corr = (m_new == -CUDART_INF_F) ? 1.f : attn_exp<FAST>(m[u] - m_new);
      float keep[4];
      attn_keep4(dc, bh, d.Lq, Lk4, min(i, d.Lq - 1), j0 + tx * 4, keep);
      float rs = 0.f, pk[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float p = (s[u][v] == -CUDART_INF_F) ? 0.f : attn_exp<FAST>(s[u][v] - m_new);
        rs += p;
        pk[v] = p * keep[v];
      }
#pragma unroll
      for (int sh = 8; sh > 0; sh >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, sh);
      l[u] = l[u] * corr + rs;
      m[u] = m_new;
#pragma unroll
      for (int w = 0; w < W; ++w) o[u][w] *= corr;
      *reinterpret_cast<float4*>(Ps + (ty * 4 + u) * AP + tx * 4) = make_float4(pk[0], pk[1], pk[2], pk[3]);
    }
    __syncthreads();
    mini_pv<HDP>(Ps, Vn, min(AT, round4(j_end - j0)), ty, tx, o);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ty * 4 + u;
    if (i >= d.Lq) continue;
    const float inv = 1.f / l[u];
    float* op = d.o + ((int64_t)i * d.B + b) * d.ldo + h * d.hd;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int dd = tx * W + w;
      if (dd < d.hd) op[dd] = o[u][w] * inv;
    }
    if (tx == 0 && d.lse) d.lse[(int64_t)bh * d.Lq + i] = m[u] + logf(l[u]);
  }
}

// ------------------------------------------------------------------------------ backward: dQ (+ delta)
template <int HDP, bool FAST>
__global__ void __launch_bounds__(A_THREADS) attn_bwd_dq_kernel(const __grid_constant__ Group<mtb_attn_bwd_desc> g) {
  extern __shared__ __align__(16) float smem[];
  float* Qt = smem;                      // [HDP][AP]
  float* dOt = Qt + HDP * AP;            // [HDP][AP]
  float* Kt = dOt + HDP * AP;            // [HDP][AP]
  float* Vt = Kt + HDP * AP;             // [HDP][AP]
  float* Kn = Vt + HDP * AP;             // [AT][HDP+2]
  float* dSs = Kn + AT * (HDP + 2);      // [AT][AP]
  float* row_lse = dSs + AT * AP;        // [AT]
  float* row_delta = row_lse + AT;       // [AT]
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_attn_bwd_desc& d = g.d[pi];
  const int qtiles = (d.Lq + AT - 1) / AT;
  const int bh = local / qtiles, qt = local - bh * qtiles;
  const int b = bh / d.H, h = bh - b * d.H;
  const int i0 = qt * AT;
  const int off = abs(d.Lk - d.Lq);
  const int Lk4 = round4(d.Lk);
  const DropCtx dc = make_drop(d.rng, d.p);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  constexpr int W = HDP / 16;

  load_tile_T<HDP>(Qt, d.q, d.ldq, d.B, b, h, d.hd, i0, d.Lq);
  load_tile_T<HDP>(dOt, d.d_o, d.lddo, d.B, b, h, d.hd, i0, d.Lq);
  if (tid < AT) {
    const int i = i0 + tid;
    float dl = 0.f, ls = 0.f;
    if (i < d.Lq) {
      const float* op = d.o + ((int64_t)i * d.B + b) * d.ldo + h * d.hd;
      const float* gp = d.d_o + ((int64_t)i * d.B + b) * d.lddo + h * d.hd;
      for (int dd = 0; dd < d.hd; ++dd) dl = fmaf(op[dd], gp[dd], dl);
      ls = d.lse[(int64_t)bh * d.Lq + i];
      d.delta[(int64_t)bh * d.Lq + i] = dl;
    }
    row_lse[tid] = ls; row_delta[tid] = dl;
  }
  float dq[4][W];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int w = 0; w < W; ++w) dq[u][w] = 0.f;
  const int i_last = min(d.Lq, i0 + AT) - 1;
  const int j_end = min(d.Lk, i_last + off + 1);
  for (int j0 = 0; j0 < j_end; j0 += AT) {
    __syncthreads();
    load_tile_T<HDP>(Kt, d.k, d.ldk, d.B, b, h, d.hd, j0, d.Lk);
    load_tile_T<HDP>(Vt, d.v, d.ldv, d.B, b, h, d.hd, j0, d.Lk);
    load_tile_N<HDP>(Kn, d.k, d.ldk, d.B, b, h, d.hd, j0, d.Lk);
    tile_wait();
    __syncthreads();
    float s[4][4], dp[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) { s[u][v] = 0.f; dp[u][v] = 0.f; }
    mini_gemm<HDP>(Qt, Kt, d.hd, ty, tx, s);
    mini_gemm<HDP>(dOt, Vt, d.hd, ty, tx, dp);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = ty * 4 + u, i = i0 + r;
      float keep[4];
      attn_keep4(dc, bh, d.Lq, Lk4, min(i, d.Lq - 1), j0 + tx * 4, keep);
      float ds[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int j = j0 + tx * 4 + v;
        const bool open = (i < d.Lq) && (j < d.Lk) && (j - i < 1 + off);
        const float p = open ? attn_exp<FAST>(s[u][v] * d.scale - row_lse[r]) : 0.f;
        ds[v] = p * (dp[u][v] * keep[v] - row_delta[r]) * d.scale;
      }
      *reinterpret_cast<float4*>(dSs + r * AP + tx * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    }
    __syncthreads();
    mini_pv<HDP>(dSs, Kn, min(AT, round4(j_end - j0)), ty, tx, dq);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ty * 4 + u;
    if (i >= d.Lq) continue;
    float* qp = d.dq + ((int64_t)i * d.B + b) * d.lddq + h * d.hd;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int dd = tx * W + w;
      if (dd < d.hd) qp[dd] = dq[u][w];
    }
  }
}

// ------------------------------------------------------------------------------ backward: dK, dV
// One CTA per (b, h, k-tile); loops over the q tiles that can see this k tile.  Works in the
// transposed [j][i] thread layout so P~^T and dS^T land in shared memory without conflicts.
template <int HDP, bool FAST>
__global__ void __launch_bounds__(A_THREADS) attn_bwd_dkv_kernel(const __grid_constant__ Group<mtb_attn_bwd_desc> g) {
  extern __shared__ __align__(16) float smem[];
  float* Kt = smem;                      // [HDP][AP]
  float* Vt = Kt + HDP * AP;
  float* Qt = Vt + HDP * AP;
  float* dOt = Qt + HDP * AP;
  float* Qn = dOt + HDP * AP;            // [AT][HDP+2]
  float* dOn = Qn + AT * (HDP + 2);      // [AT][HDP+2]
  float* PT = dOn + AT * (HDP + 2);      // [AT(j)][AP(i)]
  float* dST = PT + AT * AP;             // [AT(j)][AP(i)]
  float* col_lse = dST + AT * AP;        // [AT]
  float* col_delta = col_lse + AT;       // [AT]
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_attn_bwd_desc& d = g.d[pi];
  const int ktiles = (d.Lk + AT - 1) / AT;
  const int bh = local / ktiles, kt = local - bh * ktiles;
  const int b = bh / d.H, h = bh - b * d.H;
  const int j0 = kt * AT;
  const int off = abs(d.Lk - d.Lq);
  const int Lk4 = round4(d.Lk);
  const DropCtx dc = make_drop(d.rng, d.p);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  constexpr int W = HDP / 16;

  load_tile_T<HDP>(Kt, d.k, d.ldk, d.B, b, h, d.hd, j0, d.Lk);
  load_tile_T<HDP>(Vt, d.v, d.ldv, d.B, b, h, d.hd, j0, d.Lk);
  float dk[4][W], dv[4][W];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int w = 0; w < W; ++w) { dk[u][w] = 0.f; dv[u][w] = 0.f; }
  // rows i that see column j: j - i < 1 + off  <=>  i > j - 1 - off ; first q tile containing such a row
  const int i_first = max(0, j0 - off);
  for (int i0 = (i_first / AT) * AT; i0 < d.Lq; i0 += AT) {
    __syncthreads();
    load_tile_T<HDP>(Qt, d.q, d.ldq, d.B, b, h, d.hd, i0, d.Lq);
    load_tile_T<HDP>(dOt, d.d_o, d.lddo, d.B, b, h, d.hd, i0, d.Lq);
    load_tile_N<HDP>(Qn, d.q, d.ldq, d.B, b, h, d.hd, i0, d.Lq);
    load_tile_N<HDP>(dOn, d.d_o, d.lddo, d.B, b, h, d.hd, i0, d.Lq);
    if (tid < AT) {
      const int i = i0 + tid;
      col_lse[tid] = i < d.Lq ? d.lse[(int64_t)bh * d.Lq + i] : 0.f;
      col_delta[tid] = i < d.Lq ? d.delta[(int64_t)bh * d.Lq + i] : 0.f;
    }
    tile_wait();
    __syncthreads();
    float st[4][4], dpt[4][4];     // [j = ty*4+u][i = tx*4+v]
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) { st[u][v] = 0.f; dpt[u][v] = 0.f; }
    mini_gemm<HDP>(Kt, Qt, d.hd, ty, tx, st);
    mini_gemm<HDP>(Vt, dOt, d.hd, ty, tx, dpt);
    // keep factors of my 4 x 4 block: the 4 key rows j = j0 + ty*4 .. +3 of one query column share a
    // Philox group (dropout index = (bh*Lq + i) * round4(Lk) + j), so one draw serves 4 elements
    float keepm[4][4];          // [v = query column][u = key row]
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      keepm[v][0] = keepm[v][1] = keepm[v][2] = keepm[v][3] = dc.inv_keep;
      const int i = i0 + tx * 4 + v;
      if (dc.on && i < d.Lq) {
        const uint64_t idx = ((uint64_t)((int64_t)bh * d.Lq + i)) * (uint64_t)Lk4 + (uint64_t)(j0 + ty * 4);
        const uint4 r = drop_rand4(dc, idx >> 2);
        keepm[v][0] = r.x >= dc.thr ? dc.inv_keep : 0.f; keepm[v][1] = r.y >= dc.thr ? dc.inv_keep : 0.f;
        keepm[v][2] = r.z >= dc.thr ? dc.inv_keep : 0.f; keepm[v][3] = r.w >= dc.thr ? dc.inv_keep : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int jr = ty * 4 + u, j = j0 + jr;
      float pt[4], ds[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int ir = tx * 4 + v, i = i0 + ir;
        const bool open = (i < d.Lq) && (j < d.Lk) && (j - i < 1 + off);
        const float keep = keepm[v][u];
        const float p = open ? attn_exp<FAST>(st[u][v] * d.scale - col_lse[ir]) : 0.f;
        pt[v] = p * keep;
        ds[v] = p * (dpt[u][v] * keep - col_delta[ir]) * d.scale;
      }
      *reinterpret_cast<float4*>(PT + jr * AP + tx * 4) = make_float4(pt[0], pt[1], pt[2], pt[3]);
      *reinterpret_cast<float4*>(dST + jr * AP + tx * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    }
    __syncthreads();
    const int in = min(AT, round4(d.Lq - i0));
    mini_pv<HDP>(PT, dOn, in, ty, tx, dv);
    mini_pv<HDP>(dST, Qn, in, ty, tx, dk);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = j0 + ty * 4 + u;
    if (j >= d.Lk) continue;
    float* kp = d.dk + ((int64_t)j * d.B + b) * d.lddk + h * d.hd;
    float* vp = d.dv + ((int64_t)j * d.B + b) * d.lddv + h * d.hd;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int dd = tx * W + w;
      if (dd < d.hd) { kp[dd] = dk[u][w]; vp[dd] = dv[u][w]; }
    }
  }
}

template <int HDP> static size_t fwd_smem() { return sizeof(float) * (2 * HDP * AP + AT * (HDP + 2) + AT * AP); }
template <int HDP> static size_t dq_smem() { return sizeof(float) * (4 * HDP * AP + AT * (HDP + 2) + AT * AP + 2 * AT); }
template <int HDP> static size_t dkv_smem() { return sizeof(float) * (4 * HDP * AP + 2 * AT * (HDP + 2) + 2 * AT * AP + 2 * AT); }

template <int HDP, bool FAST>
static int launch_fwd(const Group<mtb_attn_desc>& g, int tot, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    MTB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HDP, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<HDP>()));
    attr = true;
  }
  attn_fwd_kernel<HDP, FAST><<<tot, A_THREADS, fwd_smem<HDP>(), st>>>(g);
  mtb::note_launch();
  MTB_CUDA(cudaGetLastError());
  return 0;
}
template <int HDP, bool FAST>
static int launch_bwd(const Group<mtb_attn_bwd_desc>& gq, int totq, const Group<mtb_attn_bwd_desc>& gk, int totk,
                      cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    MTB_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<HDP, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dq_smem<HDP>()));
    MTB_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<HDP, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dkv_smem<HDP>()));
    attr = true;
  }
  attn_bwd_dq_kernel<HDP, FAST><<<totq, A_THREADS, dq_smem<HDP>(), st>>>(gq);
  mtb::note_launch();
  MTB_CUDA(cudaGetLastError());
  attn_bwd_dkv_kernel<HDP, FAST><<<totk, A_THREADS, dkv_smem<HDP>(), st>>>(gk);
  mtb::note_launch();
  MTB_CUDA(cudaGetLastError());
  return 0;
}

int attn_fwd_simt(const mtb_attn_desc* d, int n, cudaStream_t st) {
  Group<mtb_attn_desc> g;
  g.n = n;
  int tot = 0, maxhd = 0;
  for (int i = 0; i < n; ++i) {
    g.d[i] = d[i];
    g.start[i] = tot;
    tot += d[i].B * d[i].H * ((d[i].Lq + AT - 1) / AT);
    maxhd = d[i].hd > maxhd ? d[i].hd : maxhd;
  }
  g.start[n] = tot;
  if (tot == 0) return 0;
  MTB_CHECK(maxhd <= 64, "attention: head_dim %d > 64 not supported", maxhd);
  if (g_gemm_mode >= 1) return maxhd <= 32 ? launch_fwd<32, true>(g, tot, st) : launch_fwd<64, true>(g, tot, st);
  return maxhd <= 32 ? launch_fwd<32, false>(g, tot, st) : launch_fwd<64, false>(g, tot, st);
}

int attn_bwd_simt(const mtb_attn_bwd_desc* d, int n, cudaStream_t st) {
  Group<mtb_attn_bwd_desc> gq, gk;
  gq.n = gk.n = n;
  int totq = 0, totk = 0, maxhd = 0;
  for (int i = 0; i < n; ++i) {
    gq.d[i] = d[i]; gk.d[i] = d[i];
    gq.start[i] = totq; gk.start[i] = totk;
    totq += d[i].B * d[i].H * ((d[i].Lq + AT - 1) / AT);
    totk += d[i].B * d[i].H * ((d[i].Lk + AT - 1) / AT);
    maxhd = d[i].hd > maxhd ? d[i].hd : maxhd;
  }
  gq.start[n] = totq; gk.start[n] = totk;
  if (totq == 0 || totk == 0) return 0;
  MTB_CHECK(maxhd <= 64, "attention: head_dim %d > 64 not supported", maxhd);
  if (g_gemm_mode >= 1) return maxhd <= 32 ? launch_bwd<32, true>(gq, totq, gk, totk, st) : launch_bwd<64, true>(gq, totq, gk, totk, st);
  return maxhd <= 32 ? launch_bwd<32, false>(gq, totq, gk, totk, st) : launch_bwd<64, false>(gq, totq, gk, totk, st);
}

}  // namespace mtb

namespace mtb {
int preload_attention_simt() {
  int bad = 0;
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_fwd_kernel<32, false>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_fwd_kernel<32, true>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_fwd_kernel<64, false>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_bwd_dq_kernel<32, true>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_bwd_dq_kernel<32, false>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_bwd_dkv_kernel<32, true>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, attn_bwd_dkv_kernel<32, false>) != cudaSuccess) ++bad; }
  return bad;
}
}  // namespace mtb
