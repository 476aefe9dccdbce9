// optim.cu -- fused gradient-norm clip + Adam over the flat gradient arena (mtb_adam_step).
//
// Reference: src/train.py:181-182 (`clip_grad_norm_(model.parameters(), clip)` then `optimizer.step()`,
// optimizer = torch.optim.Adam(lr) from src/train.py:51).  torch's optimiser only visits parameters whose
// .grad is not None (sub-networks that did not run have no gradient), keeps ONE step counter
// per parameter for the bias corrections, and runs ~10 multi-tensor passes over p / g / m / v.  Here the
// whole update is three launches over static chunk tables:
//   adam_sumsq_kernel   : per-chunk sum of g^2 over the ACTIVE parameters (inactive chunks write 0)
//   adam_finalize_kernel: deterministic tree sum of the partials -> total norm, clip coefficient; bumps the
//                         step counter of every active parameter
//   adam_update_kernel  : g *= coef (written back: p.grad holds the clipped gradient like the reference),
//                         m, v, p updated with torch's formulas; 128-bit loads/stores, 32 B/element traffic
// HBM-bound: algorithmic bytes = 4 B (norm pass) + 32 B (update pass) per active element.
#include "common.cuh"

namespace mtb {

constexpr int AD_THREADS = 256;

__global__ void __launch_bounds__(AD_THREADS) adam_sumsq_kernel(const mtb_adam_desc d) {
  pdl_sync();
  const int c = blockIdx.x;
  const int pid = d.chunk_pid[c];
  __shared__ float red[AD_THREADS / 32];
  float s = 0.f;
  if (d.active[pid]) {
    const float* g = d.grad + d.chunk_off[c];
    const int n = d.chunk_n[c];
    const int n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int i = threadIdx.x; i < n4; i += AD_THREADS) {
      const float4 v = g4[i];
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += AD_THREADS) s += g[i] * g[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < AD_THREADS / 32; ++w) t += red[w];
    d.partial[c] = t;
  }
}

__global__ void __launch_bounds__(1024) adam_finalize_kernel(const mtb_adam_desc d) {
  pdl_sync();
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < d.n_chunks; i += 1024) s += (double)d.partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += red[w];
    const float norm = (float)sqrt(t);
    float coef = 1.f;
    if (d.max_norm > 0.f) {                       // torch: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
      coef = d.max_norm / (norm + 1e-6f);
      coef = coef > 1.f ? 1.f : coef;
    }
    d.scalars[0] = norm;
    d.scalars[1] = coef;
  }
  for (int i = threadIdx.x; i < d.n_params; i += 1024)
    if (d.active[i]) d.steps[i] += 1;
}

__global__ void __launch_bounds__(AD_THREADS) adam_update_kernel(const mtb_adam_desc d) {
  pdl_sync();
  const int c = blockIdx.x;
  const int pid = d.chunk_pid[c];
  if (!d.active[pid]) return;
  const float coef = d.scalars[1];
  const int step = d.steps[pid];
  // torch/optim/adam.py (_multi_tensor_adam / _single_tensor_adam, capturable=False): python-double scalars
  const double bc1 = 1.0 - pow((double)d.beta1, (double)step);
  const double bc2 = 1.0 - pow((double)d.beta2, (double)step);
  const float step_size = (float)((double)d.lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float b1 = d.beta1, b2 = d.beta2, eps = d.eps, wd = d.weight_decay;
  const float omb1 = 1.f - b1, omb2 = 1.f - b2;
  const int64_t off = d.chunk_off[c];
  const int n = d.chunk_n[c];
  float* __restrict__ p = d.chunk_param[c];
  float* __restrict__ g = d.grad + off;
  float* __restrict__ m = d.exp_avg + off;
  float* __restrict__ v = d.exp_avg_sq + off;
  uint16_t* __restrict__ sh = d.shadow ? d.shadow + off : nullptr;      // bf16 shadow of the parameter (bf16 data path)
  auto upd = [&](float& pp, float& gg, float& mm, float& vv) {
    gg *= coef;
    float ge = gg;
    if (wd != 0.f) ge = fmaf(wd, pp, ge);        // grad = grad.add(param, alpha=weight_decay)
    mm = fmaf(ge - mm, omb1, mm);                // exp_avg.lerp_(grad, 1 - beta1)
    vv = fmaf(omb2 * ge, ge, vv * b2);           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= step_size * (mm / denom);              // param.addcdiv_(exp_avg, denom, value=-step_size)
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) & 15) == 0);   // arena slots are 256 B aligned; parameters usually are
  const int n4 = vec ? (n >> 2) : 0;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int i = threadIdx.x; i < n4; i += AD_THREADS) {
    float4 P = p4[i], G = g4[i], M = m4[i], V = v4[i];
    upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
    p4[i] = P; g4[i] = G; m4[i] = M; v4[i] = V;
    if (sh) { uint2 u; u.x = pack_bf16x2(P.x, P.y); u.y = pack_bf16x2(P.z, P.w); *reinterpret_cast<uint2*>(sh + 4 * i) = u; }
  }
  for (int i = (n4 << 2) + threadIdx.x; i < n; i += AD_THREADS) {
    float P = p[i], G = g[i], M = m[i], V = v[i];
    upd(P, G, M, V);
    p[i] = P; g[i] = G; m[i] = M; v[i] = V;
    if (sh) sh[i] = (uint16_t)(pack_bf16x2(P, 0.f) & 0xffffu);
  }
}

int adam_step(const mtb_adam_desc* d, cudaStream_t st) {
  MTB_CHECK(d->n_chunks >= 1 && d->n_params >= 1, "adam_step: empty tables");
  MTB_CHECK(d->chunk_param && d->chunk_off && d->chunk_n && d->chunk_pid && d->active && d->steps && d->grad &&
            d->exp_avg && d->exp_avg_sq && d->partial && d->scalars, "adam_step: null table or arena pointer");
  MTB_CUDA(launch_k(adam_sumsq_kernel, dim3(d->n_chunks), dim3(AD_THREADS), 0, st, *d));
  note_launch();
  MTB_CUDA(launch_k(adam_finalize_kernel, dim3(1), dim3(1024), 0, st, *d));
  note_launch();
  MTB_CUDA(launch_k(adam_update_kernel, dim3(d->n_chunks), dim3(AD_THREADS), 0, st, *d));
  note_launch();
  MTB_CUDA(cudaGetLastError());
  return 0;
}

int preload_optim() {
  int bad = 0;
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, adam_sumsq_kernel) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, adam_finalize_kernel) != cudaSuccess) ++bad; }
  { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, adam_update_kernel) != cudaSuccess) ++bad; }
  return bad;
}

}  // namespace mtb
