// elementwise.cu -- fused embed-scale + sinusoidal position embedding + dropout
// (modules/dynamic_transformer.py:64-68,72-78; modules/position_embedding.py:8-27,45-83),
// its backward, and the Philox mask materialiser used by the parity tests.
// HBM-bound: 2 * 4 B per element (read x, write y); 128-bit accesses on the contiguous side.
#include "common.cuh"

namespace mtb {

constexpr int EW_THREADS = 256;
constexpr int EW_VEC_PER_THREAD = 4;   // float4s per thread -> 16 elements, 4 independent loads in flight

__device__ __forceinline__ float pe_freq(int c_half, float neg_step) {
  return expf(__fmul_rn((float)c_half, neg_step));
}

template <bool BWD>
__global__ void __launch_bounds__(EW_THREADS) embed_kernel(const __grid_constant__ Group<mtb_embed_desc> g) {
  pdl_sync();
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_embed_desc& d = g.d[pi];
  const int E = d.E, B = d.B;
  const int64_t total = (int64_t)d.L * B * E;
  const DropCtx dc = make_drop(d.rng, d.p);
  const float neg_step = -(float)(9.210340371976184 / (double)(E / 2 - 1));  // -ln(1e4)/(E/2-1)
  if ((E & 3) == 0) {
    const int64_t nvec = total >> 2;
    const int64_t base = (int64_t)local * (EW_THREADS * EW_VEC_PER_THREAD) + threadIdx.x;
#pragma unroll
    for (int u = 0; u < EW_VEC_PER_THREAD; ++u) {
      const int64_t v = base + (int64_t)u * EW_THREADS;
      if (v >= nvec) break;
      const int64_t e0 = v << 2;
      const int64_t t = e0 / E;
      const int c = (int)(e0 - t * E);
      const int l = (int)(t / B), b = (int)(t - (int64_t)l * B);
      float4 keep = make_float4(dc.inv_keep, dc.inv_keep, dc.inv_keep, dc.inv_keep);
      if (dc.on) {
        uint4 r = drop_rand4(dc, (uint64_t)v);
        keep.x = r.x >= dc.thr ? dc.inv_keep : 0.f;
        keep.y = r.y >= dc.thr ? dc.inv_keep : 0.f;
        keep.z = r.z >= dc.thr ? dc.inv_keep : 0.f;
        keep.w = r.w >= dc.thr ? dc.inv_keep : 0.f;
      }
      if (BWD) {
        const float4 dy = *reinterpret_cast<const float4*>(d.x + e0);   // d.x = dy (contiguous)
        float4 o;
        o.x = dy.x * keep.x * d.scale; o.y = dy.y * keep.y * d.scale;
        o.z = dy.z * keep.z * d.scale; o.w = dy.w * keep.w * d.scale;
        *reinterpret_cast<float4*>(d.y + e0) = o;
      } else {
        const float* xp = d.x + (int64_t)l * d.sl + (int64_t)b * d.sb;
        float4 xv;
        if (d.se == 1 && ((((uintptr_t)(xp + c)) & 15) == 0)) {
          xv = *reinterpret_cast<const float4*>(xp + c);
        } else {
          xv.x = xp[(int64_t)c * d.se]; xv.y = xp[(int64_t)(c + 1) * d.se];
          xv.z = xp[(int64_t)(c + 2) * d.se]; xv.w = xp[(int64_t)(c + 3) * d.se];
        }
        const float f0 = xp[0];                       // padding test on feature 0
        float4 pe = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f0 != 0.f) {
          const float pos = (float)(l + 1);
          float s0, c0, s1, c1;
          sincosf(__fmul_rn(pos, pe_freq(c >> 1, neg_step)), &s0, &c0);
          sincosf(__fmul_rn(pos, pe_freq((c >> 1) + 1, neg_step)), &s1, &c1);
          pe = make_float4(s0, c0, s1, c1);
        }
        float4 o;
        o.x = (__fmul_rn(d.scale, xv.x) + pe.x) * keep.x;
        o.y = (__fmul_rn(d.scale, xv.y) + pe.y) * keep.y;
        o.z = (__fmul_rn(d.scale, xv.z) + pe.z) * keep.z;
        o.w = (__fmul_rn(d.scale, xv.w) + pe.w) * keep.w;
        *reinterpret_cast<float4*>(d.y + e0) = o;
      }
    }
  } else {   // generic width (reference toy configs): scalar path
    const int64_t base = (int64_t)local * (EW_THREADS * EW_VEC_PER_THREAD * 4);
    for (int64_t e = base + threadIdx.x; e < base + EW_THREADS * EW_VEC_PER_THREAD * 4 && e < total; e += EW_THREADS) {
      const int64_t t = e / E;
      const int c = (int)(e - t * E);
      const int l = (int)(t / B), b = (int)(t - (int64_t)l * B);
      const float keep = dc.on ? (drop_keep1(dc, (uint64_t)e) ? dc.inv_keep : 0.f) : 1.f;
      if (BWD) {
        d.y[e] = d.x[e] * keep * d.scale;
      } else {
        const float* xp = d.x + (int64_t)l * d.sl + (int64_t)b * d.sb;
        float pe = 0.f;
        if (xp[0] != 0.f) {
          const float a = __fmul_rn((float)(l + 1), pe_freq(c >> 1, neg_step));
          pe = (c & 1) ? cosf(a) : sinf(a);
        }
        d.y[e] = (__fmul_rn(d.scale, xp[(int64_t)c * d.se]) + pe) * keep;
      }
    }
  }
}

template <bool BWD>
static int launch_embed(const mtb_embed_desc* d, int n, cudaStream_t st) {
  MTB_CHECK(n >= 1 && n <= MTB_MAX_GROUP, "embed: group size %d out of range", n);
  Group<mtb_embed_desc> g;
  g.n = n;
  int tot = 0;
  for (int i = 0; i < n; ++i) {
    MTB_CHECK(d[i].E >= 4 || BWD, "embed: E=%d too small for the sinusoid (E/2-1 == 0)", d[i].E);
    g.d[i] = d[i];
    g.start[i] = tot;
    const int64_t total = (int64_t)d[i].L * d[i].B * d[i].E;
    const int64_t per = (int64_t)EW_THREADS * EW_VEC_PER_THREAD * 4;
    tot += (int)((total + per - 1) / per);
  }
  g.start[n] = tot;
  if (tot == 0) return 0;
  MTB_CUDA(launch_k(embed_kernel<BWD>, dim3(tot), dim3(EW_THREADS), 0, st, g));
  mtb::note_launch();
  MTB_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(EW_THREADS) addn_kernel(const __grid_constant__ Group<mtb_addn_desc> g) {
  pdl_sync();
  int local;
  const int pi = find_problem(g, blockIdx.x, local);
  const mtb_addn_desc& d = g.d[pi];
  const int E = d.E;
  const int64_t total = (int64_t)d.T * E;
  const bool vec = (E & 3) == 0 && (d.ld_dst & 3) == 0 && ((((uintptr_t)d.dst) & 15) == 0) &&
                   (d.ld_src[0] & 3) == 0 && ((((uintptr_t)d.src[0]) & 15) == 0) &&
                   (d.n_src < 2 || ((d.ld_src[1] & 3) == 0 && ((((uintptr_t)d.src[1]) & 15) == 0))) &&
                   (d.n_src < 3 || ((d.ld_src[2] & 3) == 0 && ((((uintptr_t)d.src[2]) & 15) == 0)));
  if (vec) {
    const int64_t nvec = total >> 2;
    const int ev = E >> 2;
    for (int u = 0; u < EW_VEC_PER_THREAD; ++u) {
      const int64_t v = (int64_t)local * (EW_THREADS * EW_VEC_PER_THREAD) + threadIdx.x + (int64_t)u * EW_THREADS;
      if (v >= nvec) break;
      const int64_t t = v / ev;
      const int c = (int)(v - t * ev) << 2;
      float4 acc = d.accumulate ? ld4_any(d.dst, t * d.ld_dst + c, d.dst_bf16 != 0) : make_float4(0, 0, 0, 0);
      for (int i = 0; i < d.n_src; ++i) {
        const float4 a = ld4_any(d.src[i], t * d.ld_src[i] + c, d.src_bf16[i] != 0);
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      }
      st4_any(d.dst, t * d.ld_dst + c, acc, d.dst_bf16 != 0);
    }
  } else {
    const int64_t base = (int64_t)local * (EW_THREADS * EW_VEC_PER_THREAD * 4);
    for (int64_t e = base + threadIdx.x; e < base + EW_THREADS * EW_VEC_PER_THREAD * 4 && e < total; e += EW_THREADS) {
      const int64_t t = e / E;
      const int c = (int)(e - t * E);
      float acc = d.accumulate ? ld1_any(d.dst, t * d.ld_dst + c, d.dst_bf16 != 0) : 0.f;
      for (int i = 0; i < d.n_src; ++i) acc += ld1_any(d.src[i], t * d.ld_src[i] + c, d.src_bf16[i] != 0);
      st1_any(d.dst, t * d.ld_dst + c, acc, d.dst_bf16 != 0);
    }
  }
}

__global__ void mask_kernel(mtb_rng rng, float p, int64_t n, uint8_t* keep) {
  const DropCtx dc = make_drop(rng, p);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = dc.on ? (drop_keep1(dc, (uint64_t)i) ? 1 : 0) : 1;
}

__global__ void rng_advance_kernel(uint64_t* st, uint64_t delta) { st[1] += delta; }

}  // namespace mtb

namespace mtb {
int preload_elementwise() {
  int bad = 0;
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, embed_kernel<false>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, embed_kernel<true>) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, addn_kernel) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, mask_kernel) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, rng_advance_kernel) != cudaSuccess) ++bad; }
  return bad;
}
}  // namespace mtb

extern "C" {
int mtb_embed_fwd(const mtb_embed_desc* d, int n, void* stream) {
  return mtb::launch_embed<false>(d, n, (cudaStream_t)stream);
}
int mtb_embed_bwd(const mtb_embed_desc* d, int n, void* stream) {
  return mtb::launch_embed<true>(d, n, (cudaStream_t)stream);
}
int mtb_addn(const mtb_addn_desc* d, int n, void* stream) {
  using namespace mtb;
  MTB_CHECK(n >= 1 && n <= MTB_MAX_GROUP, "addn: group size %d out of range", n);
  Group<mtb_addn_desc> g;
  g.n = n;
  int tot = 0;
  for (int i = 0; i < n; ++i) {
    MTB_CHECK(d[i].n_src >= 1 && d[i].n_src <= 3 && d[i].dst && d[i].src[0], "addn: bad problem %d", i);
    g.d[i] = d[i];
    g.start[i] = tot;
    const int64_t total = (int64_t)d[i].T * d[i].E;
    const int64_t per = (int64_t)EW_THREADS * EW_VEC_PER_THREAD * 4;
    tot += (int)((total + per - 1) / per);
  }
  g.start[n] = tot;
  if (tot == 0) return 0;
  MTB_CUDA(launch_k(addn_kernel, dim3(tot), dim3(EW_THREADS), 0, (cudaStream_t)stream, g));
  mtb::note_launch();
  MTB_CUDA(cudaGetLastError());
  return 0;
}
int mtb_dropout_mask(mtb_rng rng, float p, int64_t n, uint8_t* keep, void* stream) {
  if (n <= 0) return 0;
  mtb::mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rng, p, n, keep);
  mtb::note_launch();
  MTB_CUDA(cudaGetLastError());
  return 0;
}
int mtb_rng_advance(uint64_t* rng_dev, uint64_t delta, void* stream) {
  mtb::rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(rng_dev, delta);
  mtb::note_launch();
  MTB_CUDA(cudaGetLastError());
  return 0;
}
}
