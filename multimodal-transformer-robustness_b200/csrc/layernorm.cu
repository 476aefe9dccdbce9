// layernorm.cu -- fused dropout + residual add + LayerNorm (forward and backward).
// modules/dynamic_transformer.py:163,169-170,173-178,185-187,87; modules/dynamic_layers.py:61-67.
// One warp per token row, the row lives in registers (128-bit loads/stores, warp-shuffle
// reductions, two-pass mean/variance).  HBM-bound: forward 4*4 B per element
// (read branch, read residual, write new residual, write normed) + 8 B per row of stats.
#include "common.cuh"
#include <stdlib.h>

namespace mtb {

constexpr int LN_WARPS = 8;
constexpr int LN_THREADS = LN_WARPS * 32;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 gather4(const float* w, const int32_t* idx, int c) {
  if (idx == nullptr) return ld4(w + c);
  return make_float4(w[idx[c]], w[idx[c + 1]], w[idx[c + 2]], w[idx[c + 3]]);
}
__device__ __forceinline__ float4 keep4(const DropCtx& dc, uint64_t grp) {
  float4 k = make_float4(dc.inv_keep, dc.inv_keep, dc.inv_keep, dc.inv_keep);
  if (dc.on) {
    uint4 r = drop_rand4(dc, grp);
    k.x = r.x >= dc.thr ? dc.inv_keep : 0.f; k.y = r.y >= dc.thr ? dc.inv_keep : 0.f;
    k.z = r.z >= dc.thr ? dc.inv_keep : 0.f; k.w = r.w >= dc.thr ? dc.inv_keep : 0.f;
  }
  return k;
}

// ------------------------------------------------------------------ forward (vector path)
template <int MAXV>
__global__ void __launch_bounds__(LN_THREADS) resln_fwd_kernel(const __grid_constant__ Group<mtb_resln_desc> g) {
  pdl_sync();
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_resln_desc& d = g.d[pi];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = local * LN_WARPS + warp;
  if (t >= d.T) return;
  const int E = d.E, nv = E >> 2;
  const DropCtx dc = make_drop(d.rng, d.p);
  float4 x[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + 32 * i;
    x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v < nv) {
      const int c = v << 2;
      float4 r = d.res ? ld4(d.res + (int64_t)t * d.ld_res + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (d.a) {
        const float4 a = ld4_any(d.a, (int64_t)t * d.ld_a + c, d.a_bf16 != 0);
        const float4 k = keep4(dc, ((uint64_t)t * E + c) >> 2);
        r.x += a.x * k.x; r.y += a.y * k.y; r.z += a.z * k.z; r.w += a.w * k.w;
        if (d.x_new) st4(d.x_new + (int64_t)t * d.ld_x + c, r);
      }
      x[i] = r;
      sum += (r.x + r.y) + (r.z + r.w);
    }
  }
  if (d.gamma == nullptr) return;
  const float mean = warp_sum(sum) / (float)E;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (lane + 32 * i < nv) {
      const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, e = x[i].w - mean;
      sq += (a * a + b * b) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)E + d.eps);
  if (lane == 0 && d.mean) { d.mean[t] = mean; d.rstd[t] = rstd; }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      const int c = v << 2;
      const float4 gm = gather4(d.gamma, d.idx, c), bt = gather4(d.beta, d.idx, c);
      float4 y;
      y.x = (x[i].x - mean) * rstd * gm.x + bt.x; y.y = (x[i].y - mean) * rstd * gm.y + bt.y;
      y.z = (x[i].z - mean) * rstd * gm.z + bt.z; y.w = (x[i].w - mean) * rstd * gm.w + bt.w;
      st4_any(d.y, (int64_t)t * d.ld_y + c, y, d.y_bf16 != 0);
    }
  }
}

// ------------------------------------------------------------------ forward (generic width)
__global__ void __launch_bounds__(LN_THREADS) resln_fwd_generic(const __grid_constant__ Group<mtb_resln_desc> g) {
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_resln_desc& d = g.d[pi];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = local * LN_WARPS + warp;
  if (t >= d.T) return;
  const int E = d.E;
  const DropCtx dc = make_drop(d.rng, d.p);
  auto value = [&](int c) -> float {
    float r = d.res ? d.res[(int64_t)t * d.ld_res + c] : 0.f;
    if (d.a) {
      const float k = dc.on ? (drop_keep1(dc, (uint64_t)t * E + c) ? dc.inv_keep : 0.f) : 1.f;
      r += d.a[(int64_t)t * d.ld_a + c] * k;
    }
    return r;
  };
  float sum = 0.f;
  for (int c = lane; c < E; c += 32) {
    const float r = value(c);
    if (d.a && d.x_new) d.x_new[(int64_t)t * d.ld_x + c] = r;
    sum += r;
  }
  if (d.gamma == nullptr) return;
  const float mean = warp_sum(sum) / (float)E;
  float sq = 0.f;
  for (int c = lane; c < E; c += 32) { const float r = value(c) - mean; sq += r * r; }
  const float rstd = rsqrtf(warp_sum(sq) / (float)E + d.eps);
  if (lane == 0 && d.mean) { d.mean[t] = mean; d.rstd[t] = rstd; }
  for (int c = lane; c < E; c += 32) {
    const int ic = d.idx ? d.idx[c] : c;
    d.y[(int64_t)t * d.ld_y + c] = (value(c) - mean) * rstd * d.gamma[ic] + d.beta[ic];
  }
}

// ------------------------------------------------------------------ backward (vector path)
// Each CTA owns `rows_per_cta` consecutive rows; per-lane partial dgamma/dbeta live in
// registers, are combined through shared memory once per CTA and leave as one atomicAdd
// per feature per CTA.  gamma is read once per CTA, not per row (the gathered read is two dependent loads).
// R = rows a warp keeps in flight: measured on the bench step R = 1 3.04 ms, R = 2 3.05-3.07 ms (122 registers),
// R = 4 3.21 ms (177 registers: one CTA per SM, a second wave) -- only R = 1 is instantiated.
template <int MAXV, bool AFFINE_GRAD, int R>
__global__ void __launch_bounds__(LN_THREADS) resln_bwd_kernel(const __grid_constant__ Group<mtb_resln_bwd_desc> g,
                                                               int rows_per_cta) {
  pdl_sync();
  extern __shared__ float sred[];   // [3][E] when AFFINE_GRAD: dgamma, dbeta, dbias partials
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_resln_bwd_desc& d = g.d[pi];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int E = d.E, nv = E >> 2;
  const DropCtx dc = make_drop(d.rng, d.p);
  const bool has_ln = d.dy != nullptr;
  const bool agrad = AFFINE_GRAD && d.dgamma != nullptr && has_ln;
  const bool bgrad = AFFINE_GRAD && d.dbias != nullptr && d.d_a != nullptr;
  float4 dg[AFFINE_GRAD ? MAXV : 1], db[AFFINE_GRAD ? MAXV : 1], dba[AFFINE_GRAD ? MAXV : 1];
  if (AFFINE_GRAD) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) { dg[i] = make_float4(0, 0, 0, 0); db[i] = make_float4(0, 0, 0, 0); dba[i] = make_float4(0, 0, 0, 0); }
    for (int c = threadIdx.x; c < 3 * E; c += LN_THREADS) sred[c] = 0.f;
    __syncthreads();
  }
  float4 gm[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + 32 * i;
    gm[i] = (v < nv && has_ln) ? gather4(d.gamma, d.idx, v << 2) : make_float4(0, 0, 0, 0);
  }
  const int row0 = local * rows_per_cta;
  const int row1 = min(d.T, row0 + rows_per_cta);
  for (int tb = row0 + warp; tb < row1; tb += LN_WARPS * R) {
    float4 wdy[R][MAXV], xh[R][MAXV], ex[R > 1 ? R : 1][R > 1 ? MAXV : 1];   // R == 1 (wide rows): no register room to prefetch ex
    float mean[R], rstd[R], c1[R], c2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = tb + r * LN_WARPS;
      mean[r] = 0.f; rstd[r] = 0.f;
      if (t < row1 && has_ln) { mean[r] = d.mean[t]; rstd[r] = d.rstd[t]; }
    }
    // all loads of the R rows first (dy and x_new land in wdy / xh, the pass-through gradient in ex)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = tb + r * LN_WARPS;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int v = lane + 32 * i, c = v << 2;
        const bool ok = v < nv && t < row1;
        wdy[r][i] = make_float4(0, 0, 0, 0); xh[r][i] = make_float4(0, 0, 0, 0);
        if (ok && has_ln) {
          wdy[r][i] = ld4_any(d.dy, (int64_t)t * d.ld_dy + c, d.dy_bf16 != 0);
          xh[r][i] = ld4(d.x_new + (int64_t)t * d.ld_x + c);
        }
        if constexpr (R > 1) ex[r][i] = (ok && d.d_xnew) ? ld4(d.d_xnew + (int64_t)t * d.ld_dx + c) : make_float4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s1 = 0.f, s2 = 0.f;
      if (has_ln) {
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
          const int v = lane + 32 * i;
          if (v < nv && tb + r * LN_WARPS < row1) {
            const float4 dy = wdy[r][i], x = xh[r][i];
            xh[r][i] = make_float4((x.x - mean[r]) * rstd[r], (x.y - mean[r]) * rstd[r], (x.z - mean[r]) * rstd[r], (x.w - mean[r]) * rstd[r]);
            wdy[r][i] = make_float4(dy.x * gm[i].x, dy.y * gm[i].y, dy.z * gm[i].z, dy.w * gm[i].w);
            s1 += (wdy[r][i].x * xh[r][i].x + wdy[r][i].y * xh[r][i].y) + (wdy[r][i].z * xh[r][i].z + wdy[r][i].w * xh[r][i].w);
            s2 += (wdy[r][i].x + wdy[r][i].y) + (wdy[r][i].z + wdy[r][i].w);
            if (AFFINE_GRAD && agrad) {
              dg[i].x += dy.x * xh[r][i].x; dg[i].y += dy.y * xh[r][i].y; dg[i].z += dy.z * xh[r][i].z; dg[i].w += dy.w * xh[r][i].w;
              db[i].x += dy.x; db[i].y += dy.y; db[i].z += dy.z; db[i].w += dy.w;
            }
          }
        }
      }
      c1[r] = s1; c2[r] = s2;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) { c1[r] = warp_sum(c1[r]) / (float)E; c2[r] = warp_sum(c2[r]) / (float)E; }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = tb + r * LN_WARPS;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int v = lane + 32 * i;
        if (v < nv && t < row1) {
          const int c = v << 2;
          float4 gx;
          gx.x = rstd[r] * (wdy[r][i].x - c2[r] - xh[r][i].x * c1[r]); gx.y = rstd[r] * (wdy[r][i].y - c2[r] - xh[r][i].y * c1[r]);
          gx.z = rstd[r] * (wdy[r][i].z - c2[r] - xh[r][i].z * c1[r]); gx.w = rstd[r] * (wdy[r][i].w - c2[r] - xh[r][i].w * c1[r]);
          if (d.d_xnew) {
            float4 e;
            if constexpr (R > 1) e = ex[r][i]; else e = ld4(d.d_xnew + (int64_t)t * d.ld_dx + c);
            gx.x += e.x; gx.y += e.y; gx.z += e.z; gx.w += e.w;
          }
          if (d.d_res) st4(d.d_res + (int64_t)t * d.ld_dres + c, gx);
          if (d.d_a) {
            const float4 k = keep4(dc, ((uint64_t)t * E + c) >> 2);
            float4 da = make_float4(gx.x * k.x, gx.y * k.y, gx.z * k.z, gx.w * k.w);
            if (d.da_bf16) {     // the bias gradient sums exactly what the GEMMs will read
              da = make_float4(bf16_round(da.x), bf16_round(da.y), bf16_round(da.z), bf16_round(da.w));
            }
            st4_any(d.d_a, (int64_t)t * d.ld_da + c, da, d.da_bf16 != 0);
            if (AFFINE_GRAD && bgrad) { dba[i].x += da.x; dba[i].y += da.y; dba[i].z += da.z; dba[i].w += da.w; }
          }
        }
      }
    }
  }
  if (AFFINE_GRAD) {
    // combine the 8 warps' register partials in shared memory, one warp at a time (plain adds between
    // barriers: shared-memory float atomics are CAS loops and serialise badly under 8-way contention)
#pragma unroll 1
    for (int w = 0; w < LN_WARPS; ++w) {
      if (warp == w) {
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
          const int v = lane + 32 * i;
          if (v < nv) {
            const int c = v << 2;
            if (agrad) {
              sred[c] += dg[i].x; sred[c + 1] += dg[i].y; sred[c + 2] += dg[i].z; sred[c + 3] += dg[i].w;
              sred[E + c] += db[i].x; sred[E + c + 1] += db[i].y; sred[E + c + 2] += db[i].z; sred[E + c + 3] += db[i].w;
            }
            if (bgrad) {
              sred[2 * E + c] += dba[i].x; sred[2 * E + c + 1] += dba[i].y;
              sred[2 * E + c + 2] += dba[i].z; sred[2 * E + c + 3] += dba[i].w;
            }
          }
        }
      }
      __syncthreads();
    }
    if (agrad || bgrad) {
      for (int c = threadIdx.x; c < E; c += LN_THREADS) {
        const int ic = d.idx ? d.idx[c] : c;
        if (agrad) {
          atomicAdd(&d.dgamma[ic], sred[c]);
          if (d.dbeta) atomicAdd(&d.dbeta[ic], sred[E + c]);
        }
        if (bgrad) atomicAdd(&d.dbias[ic], sred[2 * E + c]);
      }
    }
  }
}

// ------------------------------------------------------------------ backward (generic width)
__global__ void __launch_bounds__(LN_THREADS) resln_bwd_generic(const __grid_constant__ Group<mtb_resln_bwd_desc> g) {
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_resln_bwd_desc& d = g.d[pi];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = local * LN_WARPS + warp;
  if (t >= d.T) return;
  const int E = d.E;
  const DropCtx dc = make_drop(d.rng, d.p);
  const bool has_ln = d.dy != nullptr;
  float mean = 0.f, rstd = 0.f, s1 = 0.f, s2 = 0.f;
  if (has_ln) {
    mean = d.mean[t]; rstd = d.rstd[t];
    for (int c = lane; c < E; c += 32) {
      const int ic = d.idx ? d.idx[c] : c;
      const float dy = d.dy[(int64_t)t * d.ld_dy + c];
      const float xh = (d.x_new[(int64_t)t * d.ld_x + c] - mean) * rstd;
      const float w = dy * d.gamma[ic];
      s1 += w * xh; s2 += w;
      if (d.dgamma) { atomicAdd(&d.dgamma[ic], dy * xh); if (d.dbeta) atomicAdd(&d.dbeta[ic], dy); }
    }
  }
  const float c1 = warp_sum(s1) / (float)E, c2 = warp_sum(s2) / (float)E;
  for (int c = lane; c < E; c += 32) {
    float gx = 0.f;
    if (has_ln) {
      const int ic = d.idx ? d.idx[c] : c;
      const float xh = (d.x_new[(int64_t)t * d.ld_x + c] - mean) * rstd;
      gx = rstd * (d.dy[(int64_t)t * d.ld_dy + c] * d.gamma[ic] - c2 - xh * c1);
    }
    if (d.d_xnew) gx += d.d_xnew[(int64_t)t * d.ld_dx + c];
    if (d.d_res) d.d_res[(int64_t)t * d.ld_dres + c] = gx;
    if (d.d_a) {
      const float k = dc.on ? (drop_keep1(dc, (uint64_t)t * E + c) ? dc.inv_keep : 0.f) : 1.f;
      d.d_a[(int64_t)t * d.ld_da + c] = gx * k;
      if (d.dbias) atomicAdd(&d.dbias[d.idx ? d.idx[c] : c], gx * k);
    }
  }
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace mtb

namespace mtb {
int preload_layernorm() {
  int bad = 0;
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, resln_fwd_kernel<2>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, resln_fwd_kernel<8>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, resln_fwd_generic) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, resln_bwd_kernel<2, true, 1>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, resln_bwd_kernel<2, false, 1>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, resln_bwd_kernel<8, true, 1>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, resln_bwd_kernel<8, false, 1>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, resln_bwd_generic) != cudaSuccess) ++bad; }
  return bad;
}
}  // namespace mtb

extern "C" {

int mtb_resln_fwd(const mtb_resln_desc* d, int n, void* stream) {
  using namespace mtb;
  MTB_CHECK(n >= 1 && n <= MTB_MAX_GROUP, "resln_fwd: group size %d out of range", n);
  Group<mtb_resln_desc> g;
  g.n = n;
  int tot = 0, maxE = 0;
  bool vec = true;
  for (int i = 0; i < n; ++i) {
    const mtb_resln_desc& x = d[i];
    MTB_CHECK(x.res || x.a, "resln_fwd: problem %d has neither residual nor branch input", i);
    MTB_CHECK(!x.gamma || x.y, "resln_fwd: problem %d has gamma but no output", i);
    g.d[i] = x;
    g.start[i] = tot;
    tot += (x.T + LN_WARPS - 1) / LN_WARPS;
    maxE = x.E > maxE ? x.E : maxE;
    const bool v1 = (x.E % 4 == 0) && (!x.res || (aligned16(x.res) && x.ld_res % 4 == 0)) &&
          (!x.a || (aligned16(x.a) && x.ld_a % 4 == 0)) && (!x.x_new || (aligned16(x.x_new) && x.ld_x % 4 == 0)) &&
          (!x.y || (aligned16(x.y) && x.ld_y % 4 == 0)) && (!x.gamma || x.idx || (aligned16(x.gamma) && aligned16(x.beta)));
    MTB_CHECK(v1 || !(x.a_bf16 || x.y_bf16), "resln_fwd: problem %d has bf16 operands but is not 4-element aligned", i);
    MTB_CHECK(x.E <= 1024 || !(x.a_bf16 || x.y_bf16), "resln_fwd: bf16 operands need E <= 1024 (problem %d)", i);
    vec = vec && v1;
  }
  g.start[n] = tot;
  if (tot == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec && maxE <= 256) MTB_CUDA(launch_k(resln_fwd_kernel<2>, dim3(tot), dim3(LN_THREADS), 0, st, g));
  else if (vec && maxE <= 1024) MTB_CUDA(launch_k(resln_fwd_kernel<8>, dim3(tot), dim3(LN_THREADS), 0, st, g));
  else resln_fwd_generic<<<tot, LN_THREADS, 0, st>>>(g);
  mtb::note_launch();
  MTB_CUDA(cudaGetLastError());
  return 0;
}

int mtb_resln_bwd(const mtb_resln_bwd_desc* d, int n, void* stream) {
  using namespace mtb;
  MTB_CHECK(n >= 1 && n <= MTB_MAX_GROUP, "resln_bwd: group size %d out of range", n);
  Group<mtb_resln_bwd_desc> g;
  g.n = n;
  int maxE = 0, maxT = 0;
  bool vec = true, affine = false;
  for (int i = 0; i < n; ++i) {
    const mtb_resln_bwd_desc& x = d[i];
    MTB_CHECK(x.dy || x.d_xnew, "resln_bwd: problem %d has no incoming gradient", i);
    maxE = x.E > maxE ? x.E : maxE;
    maxT = x.T > maxT ? x.T : maxT;
    affine = affine || (x.dgamma != nullptr) || (x.dbias != nullptr);
    const bool v1 = (x.E % 4 == 0) && (!x.dy || (aligned16(x.dy) && x.ld_dy % 4 == 0 && aligned16(x.x_new) && x.ld_x % 4 == 0)) &&
          (!x.d_xnew || (aligned16(x.d_xnew) && x.ld_dx % 4 == 0)) && (!x.d_res || (aligned16(x.d_res) && x.ld_dres % 4 == 0)) &&
          (!x.d_a || (aligned16(x.d_a) && x.ld_da % 4 == 0)) && (!x.dy || x.idx || aligned16(x.gamma));
    MTB_CHECK((v1 && x.E <= 1024) || !(x.dy_bf16 || x.da_bf16), "resln_bwd: problem %d has bf16 operands but is not vectorisable", i);
    vec = vec && v1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int tot = 0;
  if (vec && maxE <= 1024) {
    // ~2 CTAs per SM when column sums (d-gamma / d-beta / bias grads) are accumulated -- fewer, longer CTAs
    // mean fewer contended global atomics per feature; ~4 waves otherwise
    static const int mult_env = getenv("MTB_LN_BWD_CTAS") ? atoi(getenv("MTB_LN_BWD_CTAS")) : 0;
    const int target = mult_env > 0 ? sm_count() * mult_env : (affine ? sm_count() * 2 : sm_count() * 4);
    int rows = (maxT + target - 1) / target;
    rows = ((rows + LN_WARPS - 1) / LN_WARPS) * LN_WARPS;
    if (rows < LN_WARPS) rows = LN_WARPS;
    for (int i = 0; i < n; ++i) { g.d[i] = d[i]; g.start[i] = tot; tot += (d[i].T + rows - 1) / rows; }
    g.start[n] = tot;
    if (tot == 0) return 0;
    const size_t smem = affine ? 3 * (size_t)maxE * sizeof(float) : 0;
    if (maxE <= 256) {
      if (affine) MTB_CUDA(launch_k(resln_bwd_kernel<2, true, 1>, dim3(tot), dim3(LN_THREADS), smem, st, g, rows));
      else MTB_CUDA(launch_k(resln_bwd_kernel<2, false, 1>, dim3(tot), dim3(LN_THREADS), 0, st, g, rows));
    } else {
      if (affine) MTB_CUDA(launch_k(resln_bwd_kernel<8, true, 1>, dim3(tot), dim3(LN_THREADS), smem, st, g, rows));
      else MTB_CUDA(launch_k(resln_bwd_kernel<8, false, 1>, dim3(tot), dim3(LN_THREADS), 0, st, g, rows));
    }
  } else {
    for (int i = 0; i < n; ++i) { g.d[i] = d[i]; g.start[i] = tot; tot += (d[i].T + LN_WARPS - 1) / LN_WARPS; }
    g.start[n] = tot;
    if (tot == 0) return 0;
    resln_bwd_generic<<<tot, LN_THREADS, 0, st>>>(g);
  }
  mtb::note_launch();
  MTB_CUDA(cudaGetLastError());
  return 0;
}
}
