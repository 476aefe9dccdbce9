// linear_simt.cu -- fp32 CUDA-core grouped GEMM engine (parity mode, <= 1e-5 vs the fp32
// reference) behind mtb_linear_fwd / mtb_linear_bwd.
// modules/dynamic_multihead_attention.py:259-282, modules/dynamic_layers.py:15-25.
// One generic kernel computes  C[i,j] (op)= sum_r A(i,r) * B(j,r)  with strided / index-mapped
// operand access, so forward, dgrad and wgrad (and the head/dim slicing + active_mask gathers
// of the reference) are the same code with different strides:
//   fwd   : i=m j=n r=k   A=X          B=W[row(n), col(k)]          C=Y (+bias, ReLU, dropout)
//   dgrad : i=m j=k r=n   A=dY'        B=W[row(n), col(k)] (j<->r)  C=dX
//   wgrad : i=n j=k r=m   A=dY'^T      B=X^T                        C=dW[row(n), col(k)] (atomic, split-R)
#include "common.cuh"
#include <stdlib.h>

namespace mtb {

constexpr int GB = 64;         // tile edge (I and J)
constexpr int GK = 16;         // reduction slab
constexpr int GPAD = 4;
constexpr int G_THREADS = 256;

struct GemmP {
  const float* A; int64_t sAi, sAr;
  const float* A2; int64_t sA2i, sA2r; float a_scale;      // A *= (A2 > 0) * a_scale when A2 != null
  const float* Bm; int64_t sBj, sBr; const int32_t* mapBj; const int32_t* mapBr;
  float* C; int64_t sCi, sCj; const int32_t* mapCi; const int32_t* mapCj;
  const float* bias; const int32_t* mapBias;
  float* rowsum; const int32_t* mapRowsum;                  // rowsum[map(i)] += sum_r A(i,r)   (bias grad)
  int I, J, R;
  int epi;            // 0: C = acc, 1: C += acc, 2: atomicAdd(C, acc)
  int act;            // 1: ReLU + dropout on the stored value (forward only)
  int splits;         // split-R factor (epi == 2 only)
  float p; mtb_rng rng;
};

__global__ void __launch_bounds__(G_THREADS) gemm_simt_kernel(const __grid_constant__ Group<GemmP> g) {
  __shared__ __align__(16) float As[GK][GB + GPAD];
  __shared__ __align__(16) float Bs[GK][GB + GPAD];
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const GemmP& P = g.d[pi];
  const int tj_n = (P.J + GB - 1) / GB, ti_n = (P.I + GB - 1) / GB;
  const int split = local / (ti_n * tj_n);
  const int tile = local - split * (ti_n * tj_n);
  const int ti = tile / tj_n, tj = tile - ti * tj_n;
  const int i0 = ti * GB, j0 = tj * GB;
  // reduction range of this split, in whole slabs
  const int slabs = (P.R + GK - 1) / GK;
  const int per = (slabs + P.splits - 1) / P.splits;
  const int r_begin = split * per * GK;
  const int r_end = min(P.R, (split + 1) * per * GK);

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
  float rsum = 0.f;
  const bool a_r_contig = (P.sAr == 1);
  const bool b_r_contig = (P.sBr == 1) && (P.mapBr == nullptr);
  const bool want_rowsum = (P.rowsum != nullptr) && (tj == 0);

  for (int r0 = r_begin; r0 < r_end; r0 += GK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = tid + u * G_THREADS;
      int i, r;
      if (a_r_contig) { r = e & (GK - 1); i = e >> 4; } else { i = e & (GB - 1); r = e >> 6; }
      float v = 0.f;
      const int gi = i0 + i, gr = r0 + r;
      if (gi < P.I && gr < r_end) {
        v = P.A[(int64_t)gi * P.sAi + (int64_t)gr * P.sAr];
        if (P.A2) v = (P.A2[(int64_t)gi * P.sA2i + (int64_t)gr * P.sA2r] > 0.f) ? v * P.a_scale : 0.f;
      }
      As[r][i] = v;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = tid + u * G_THREADS;
      int j, r;
      if (b_r_contig) { r = e & (GK - 1); j = e >> 4; } else { j = e & (GB - 1); r = e >> 6; }
      float v = 0.f;
      const int gj = j0 + j, gr = r0 + r;
      if (gj < P.J && gr < r_end) {
        const int64_t pj = P.mapBj ? P.mapBj[gj] : gj;
        const int64_t pr = P.mapBr ? P.mapBr[gr] : gr;
        v = P.Bm[pj * P.sBj + pr * P.sBr];
      }
      Bs[r][j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < GK; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
    }
    if (want_rowsum && tid < GB) {
#pragma unroll
      for (int r = 0; r < GK; ++r) rsum += As[r][tid];
    }
    __syncthreads();
  }

  if (want_rowsum && tid < GB && i0 + tid < P.I) {
    const int gi = i0 + tid;
    atomicAdd(&P.rowsum[P.mapRowsum ? P.mapRowsum[gi] : gi], rsum);
  }
  const DropCtx dc = make_drop(P.rng, P.p);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int gi = i0 + ty * 4 + u;
    if (gi >= P.I) continue;
    const int64_t ci = (P.mapCi ? (int64_t)P.mapCi[gi] : (int64_t)gi) * P.sCi;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int gj = j0 + tx * 4 + v;
      if (gj >= P.J) continue;
      float val = acc[u][v];
      if (P.bias && split == 0) val += P.bias[P.mapBias ? P.mapBias[gj] : gj];     // split-R: the bias joins once
      if (P.act == 1) {
        val = fmaxf(val, 0.f);
        if (dc.on) val = drop_keep1(dc, (uint64_t)gi * P.J + gj) ? val * dc.inv_keep : 0.f;
      }
      float* cp = P.C + ci + (P.mapCj ? (int64_t)P.mapCj[gj] : (int64_t)gj) * P.sCj;
      if (P.epi == 0) *cp = val;
      else if (P.epi == 1) *cp += val;
      else atomicAdd(cp, val);
    }
  }
}

int launch_gemm_simt(const GemmP* p, int n, cudaStream_t st) {
  Group<GemmP> g;
  int off = 0;
  while (off < n) {
    const int m = (n - off) < MTB_MAX_GROUP ? (n - off) : MTB_MAX_GROUP;
    int tot = 0;
    g.n = m;
    for (int i = 0; i < m; ++i) {
      g.d[i] = p[off + i];
      g.start[i] = tot;
      if (p[off + i].I > 0 && p[off + i].J > 0)
        tot += ((p[off + i].I + GB - 1) / GB) * ((p[off + i].J + GB - 1) / GB) * p[off + i].splits;
    }
    g.start[m] = tot;
    static const bool dbg = getenv("MTB_TC_DEBUG") != nullptr;
    if (dbg) {
      fprintf(stderr, "[simt] grid %d:", tot);
      for (int i = 0; i < m; ++i) fprintf(stderr, " {I %d J %d R %d sp %d epi %d act %d}", g.d[i].I, g.d[i].J, g.d[i].R, g.d[i].splits, g.d[i].epi, g.d[i].act);
      fprintf(stderr, "\n");
    }
    if (tot > 0) {
      gemm_simt_kernel<<<tot, G_THREADS, 0, st>>>(g);
      mtb::note_launch();
      MTB_CUDA(cudaGetLastError());
    }
    off += m;
  }
  return 0;
}

// ------------------------------------------------------------------ public descs -> GemmP
int linear_fwd_simt(const mtb_linear_desc* d, int n, cudaStream_t st) {
  GemmP p[MTB_MAX_GROUP];
  for (int i = 0; i < n; ++i) {
    const mtb_linear_desc& x = d[i];
    GemmP& q = p[i];
    q = GemmP{};
    q.A = x.X; q.sAi = x.ldx; q.sAr = 1;
    q.Bm = x.W; q.sBj = x.ldw; q.sBr = 1; q.mapBj = x.row_idx; q.mapBr = x.col_idx;
    q.C = x.Y; q.sCi = x.ldy; q.sCj = 1;
    q.bias = x.bias; q.mapBias = x.row_idx;
    q.I = x.M; q.J = x.N; q.R = x.K;
    q.epi = 0; q.act = x.act; q.splits = 1; q.p = x.p; q.rng = x.rng;
    // A single tile with a long, index-gathered reduction (the head's [B, C] x [C] -> [B, 1] output layer) is one
    // CTA walking the slabs serially (~2.5 us each): split the reduction, partial sums meet through atomics.
    const int tiles = ((x.M + GB - 1) / GB) * ((x.N + GB - 1) / GB);
    const int slabs = (x.K + GK - 1) / GK;
    if (x.act == 0 && tiles <= 4 && slabs >= 8 && x.ldy == x.N) {
      int sp = slabs / 2;
      if (sp > 48) sp = 48;
      q.splits = sp; q.epi = 2;
      MTB_CUDA(cudaMemsetAsync(x.Y, 0, (size_t)x.M * x.N * sizeof(float), st));
    }
  }
  return launch_gemm_simt(p, n, st);
}

static int pick_splits(int I, int J, int R) {
  const int tiles = ((I + GB - 1) / GB) * ((J + GB - 1) / GB);
  const int slabs = (R + GK - 1) / GK;
  int s = (2 * sm_count() + tiles - 1) / tiles;
  if (s > slabs / 4) s = slabs / 4;    // at least 4 slabs (64 reduction steps) per split
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return s;
}

int linear_bwd_simt(const mtb_linear_bwd_desc* d, int n, cudaStream_t st) {
  GemmP pd[MTB_MAX_GROUP], pw[MTB_MAX_GROUP];
  int nd = 0, nw = 0;
  for (int i = 0; i < n; ++i) {
    const mtb_linear_bwd_desc& x = d[i];
    const float a_scale = (x.act == 1 && x.p > 0.f) ? 1.f / (1.f - x.p) : 1.f;
    if (x.dX) {
      GemmP& q = pd[nd++];
      q = GemmP{};
      q.A = x.dY; q.sAi = x.ldy; q.sAr = 1;
      if (x.act == 1) { q.A2 = x.Yact; q.sA2i = x.ldyact; q.sA2r = 1; q.a_scale = a_scale; }
      q.Bm = x.W; q.sBj = 1; q.mapBj = x.col_idx; q.sBr = x.ldw; q.mapBr = x.row_idx;
      q.C = x.dX; q.sCi = x.lddx; q.sCj = 1;
      q.I = x.M; q.J = x.K; q.R = x.N;
      q.epi = x.accumulate_dx ? 1 : 0; q.splits = 1;
    }
    if (x.dW) {
      GemmP& q = pw[nw++];
      q = GemmP{};
      q.A = x.dY; q.sAi = 1; q.sAr = x.ldy;
      if (x.act == 1) { q.A2 = x.Yact; q.sA2i = 1; q.sA2r = x.ldyact; q.a_scale = a_scale; }
      q.Bm = x.X; q.sBj = 1; q.sBr = x.ldx;
      q.C = x.dW; q.sCi = x.ldw; q.mapCi = x.row_idx; q.sCj = 1; q.mapCj = x.col_idx;
      q.rowsum = x.db; q.mapRowsum = x.row_idx;
      q.I = x.N; q.J = x.K; q.R = x.M;
      q.epi = 2; q.splits = pick_splits(q.I, q.J, q.R);
    }
  }
  if (nd) { int rc = launch_gemm_simt(pd, nd, st); if (rc) return rc; }
  if (nw) { int rc = launch_gemm_simt(pw, nw, st); if (rc) return rc; }
  return 0;
}

}  // namespace mtb

namespace mtb {
int preload_linear_simt() {
  int bad = 0;
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, gemm_simt_kernel) != cudaSuccess) ++bad; }
  return bad;
}
}  // namespace mtb
