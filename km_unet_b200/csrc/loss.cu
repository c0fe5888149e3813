// loss.cu -- HybridLoss of the training scripts (train_shanghai.py:298-326, train_LAPS.py:347-375) as four streaming kernels
// around the two banded GEMMs of the SSIM filter.
//
//   loss = alpha (0.55 MSE + 0.45 mean((p - t)^2 exp(2 t))) + (1 - alpha) (1 - SSIM(p_n, t_n)),
//   p_n = (p - min p) / (max p - min p + 1e-8), t_n likewise (extrema detached).
// The torch formulation issues ~100 elementwise / reduction kernels over 42 MB tensors (2.1 ms per step at (32,20,128,128)); here:
//   stats   one pass over p, t: sum (p-t)^2, sum (p-t)^2 e^{2t}, min / max of p and t  (per-CTA partials, fixed-order finish)
//   stack   p, t -> the five maps the SSIM filter needs (p_n, t_n, p_n^2, t_n^2, p_n t_n) in one pass
//   [caller: the separable 11-tap Gaussian "valid" filter = two batched GEMMs with constant banded matrices, as before]
//   ssim    the five filtered maps -> sum of the SSIM map AND its five partial-derivative maps (analytic), loss value
//   [caller: the transposed filter of the derivative maps = two batched GEMMs]
//   bwd     dp = g [ alpha 2 (p - t) (0.55 + 0.45 e^{2t}) / N + (d0 + 2 p_n d2 + t_n d4) / (max p - min p + 1e-8) ]
// The SSIM term itself is torchmetrics' (not in the reference tree): "parity unpinned", see km_unet_b200/loss.py.
#include <cfloat>

#include "common.cuh"

namespace kmu {
namespace loss {

constexpr int NPART = 592;   // CTAs of the reductions (4 per SM)

__device__ __forceinline__ float block_sum256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) a += red[i];
  return a;
}
__device__ __forceinline__ float block_min256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float a = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) a = fminf(a, red[i]);
  return a;
}

// part[cta][6] = S1, S2, min p, max p, min t, max t over the CTA's grid-stride share
__global__ void __launch_bounds__(256) stats_kernel(const float4* __restrict__ p, const float4* __restrict__ t, long long n4,
                                                    float* __restrict__ part) {
  __shared__ float red[8];
  float s1 = 0.f, s2 = 0.f, pmin = FLT_MAX, pmax = -FLT_MAX, tmin = FLT_MAX, tmax = -FLT_MAX;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 a = __ldg(p + i), b = __ldg(t + i);
    const float pa[4] = {a.x, a.y, a.z, a.w}, tb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d = pa[e] - tb[e], d2 = d * d;
      s1 += d2;
      s2 = fmaf(d2, __expf(2.f * tb[e]), s2);
      pmin = fminf(pmin, pa[e]); pmax = fmaxf(pmax, pa[e]);
      tmin = fminf(tmin, tb[e]); tmax = fmaxf(tmax, tb[e]);
    }
  }
  s1 = block_sum256(s1, red);
  s2 = block_sum256(s2, red);
  pmin = block_min256(pmin, red);
  pmax = -block_min256(-pmax, red);
  tmin = block_min256(tmin, red);
  tmax = -block_min256(-tmax, red);
  if (threadIdx.x == 0) {
    float* o = part + (size_t)blockIdx.x * 6;
    o[0] = s1; o[1] = s2; o[2] = pmin; o[3] = pmax; o[4] = tmin; o[5] = tmax;
  }
}

// scal[0..7] = S1, S2, min p, 1 / (max p - min p + eps), min t, 1 / (max t - min t + eps), (unused), (unused)
__global__ void __launch_bounds__(256) stats_fin_kernel(const float* __restrict__ part, int nparts, float* __restrict__ scal) {
  __shared__ float red[8];
  float s1 = 0.f, s2 = 0.f, pmin = FLT_MAX, pmax = -FLT_MAX, tmin = FLT_MAX, tmax = -FLT_MAX;
  for (int k = threadIdx.x; k < nparts; k += 256) {
    const float* o = part + (size_t)k * 6;
    s1 += o[0]; s2 += o[1];
    pmin = fminf(pmin, o[2]); pmax = fmaxf(pmax, o[3]);
    tmin = fminf(tmin, o[4]); tmax = fmaxf(tmax, o[5]);
  }
  s1 = block_sum256(s1, red);
  s2 = block_sum256(s2, red);
  pmin = block_min256(pmin, red);
  pmax = -block_min256(-pmax, red);
  tmin = block_min256(tmin, red);
  tmax = -block_min256(-tmax, red);
  if (threadIdx.x == 0) {
    scal[0] = s1; scal[1] = s2;
    scal[2] = pmin; scal[3] = 1.f / (pmax - pmin + 1e-8f);
    scal[4] = tmin; scal[5] = 1.f / (tmax - tmin + 1e-8f);
  }
}

// stack[k][i], k = 0..4: p_n, t_n, p_n^2, t_n^2, p_n t_n
__global__ void __launch_bounds__(256) stack_kernel(const float4* __restrict__ p, const float4* __restrict__ t,
                                                    const float* __restrict__ scal, float4* __restrict__ stack, long long n4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float pmin = __ldg(scal + 2), ip = __ldg(scal + 3), tmin = __ldg(scal + 4), it = __ldg(scal + 5);
  const float4 a = __ldg(p + i), b = __ldg(t + i);
  const float4 pn = make_float4((a.x - pmin) * ip, (a.y - pmin) * ip, (a.z - pmin) * ip, (a.w - pmin) * ip);
  const float4 tn = make_float4((b.x - tmin) * it, (b.y - tmin) * it, (b.z - tmin) * it, (b.w - tmin) * it);
  stack[i] = pn;
  stack[n4 + i] = tn;
  stack[2 * n4 + i] = make_float4(pn.x * pn.x, pn.y * pn.y, pn.z * pn.z, pn.w * pn.w);
  stack[3 * n4 + i] = make_float4(tn.x * tn.x, tn.y * tn.y, tn.z * tn.z, tn.w * tn.w);
  stack[4 * n4 + i] = make_float4(pn.x * tn.x, pn.y * tn.y, pn.z * tn.z, pn.w * tn.w);
}

// f[k][i] = the five filtered maps (mu_p, mu_t, E pp, E tt, E pt).  Writes gm[0..2][i] = -(1 - alpha) / n * dm/d{mu_p, E pp, E pt}
// (the target carries no gradient) and per-CTA sums of m.
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ f, float* __restrict__ gm, float* __restrict__ part,
                                                   long long n, float c1, float c2, float gscale) {
  __shared__ float red[8];
  float sm = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float mp = __ldg(f + i), mt = __ldg(f + n + i), pp = __ldg(f + 2 * n + i), tt = __ldg(f + 3 * n + i), pt = __ldg(f + 4 * n + i);
    const float sp = pp - mp * mp, st = tt - mt * mt, spt = pt - mp * mt;
    const float a1 = 2.f * mp * mt + c1, a2 = 2.f * spt + c2, b1 = mp * mp + mt * mt + c1, b2 = sp + st + c2;
    const float inv = 1.f / (b1 * b2);
    const float m = a1 * a2 * inv;
    sm += m;
    const float dpp = -m / b2;                    // = dtt
    const float dpt = 2.f * a1 * inv;
    const float k = 2.f * (a2 - a1) * inv;        // d/dmu of the numerator terms
    const float dmp = mt * k - m * 2.f * mp * (1.f / b1 - 1.f / b2);
    gm[i] = gscale * dmp;            // only the maps that carry gradient to the prediction: d/dmu_p, d/dE[pp], d/dE[pt]
    gm[n + i] = gscale * dpp;
    gm[2 * n + i] = gscale * dpt;
  }
  sm = block_sum256(sm, red);
  if (threadIdx.x == 0) part[blockIdx.x] = sm;
}

// loss = alpha (0.55 S1 + 0.45 S2) / N + (1 - alpha) (1 - sum m / n)
__global__ void __launch_bounds__(256) loss_fin_kernel(const float* __restrict__ part, int nparts, const float* __restrict__ scal,
                                                       float* __restrict__ out, float alpha, float inv_N, float inv_n) {
  __shared__ float red[8];
  float s = 0.f;
  for (int k = threadIdx.x; k < nparts; k += 256) s += part[k];
  s = block_sum256(s, red);
  if (threadIdx.x == 0) out[0] = alpha * (0.55f * scal[0] + 0.45f * scal[1]) * inv_N + (1.f - alpha) * (1.f - s * inv_n);
}

// dp = g [ alpha 2 (p - t) (0.55 + 0.45 e^{2t}) / N + ip (d0 + 2 p_n d2 + t_n d4) ];  ds = the back-filtered derivative maps (3, N): d/dp_n, d/dp_n^2, d/d(p_n t_n)
__global__ void __launch_bounds__(256) bwd_kernel(const float4* __restrict__ p, const float4* __restrict__ t,
                                                  const float* __restrict__ scal, const float4* __restrict__ ds,
                                                  const float* __restrict__ g, float4* __restrict__ dp, long long n4, float alpha,
                                                  float inv_N) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float pmin = __ldg(scal + 2), ip = __ldg(scal + 3), tmin = __ldg(scal + 4), it = __ldg(scal + 5), gv = __ldg(g);
  const float4 a = __ldg(p + i), b = __ldg(t + i), d0 = __ldg(ds + i), d2 = __ldg(ds + n4 + i), d4 = __ldg(ds + 2 * n4 + i);
  const float pa[4] = {a.x, a.y, a.z, a.w}, tb[4] = {b.x, b.y, b.z, b.w};
  const float e0[4] = {d0.x, d0.y, d0.z, d0.w}, e2[4] = {d2.x, d2.y, d2.z, d2.w}, e4[4] = {d4.x, d4.y, d4.z, d4.w};
  float r[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float pn = (pa[e] - pmin) * ip, tn = (tb[e] - tmin) * it;
    const float mse = alpha * 2.f * (pa[e] - tb[e]) * (0.55f + 0.45f * __expf(2.f * tb[e])) * inv_N;
    r[e] = gv * (mse + ip * (e0[e] + 2.f * pn * e2[e] + tn * e4[e]));
  }
  dp[i] = make_float4(r[0], r[1], r[2], r[3]);
}

}  // namespace loss
}  // namespace kmu

using namespace kmu;
using namespace kmu::loss;

extern "C" {

size_t kmu_hybridloss_workspace_bytes(void) { return align_up((size_t)NPART * 6 * 4, 256); }

int kmu_hybridloss_stats(const float* pred, const float* target, int64_t n, float* scal, void* workspace, size_t workspace_bytes,
                         kmu_stream stream) {
  KMU_REQUIRE(pred && target && scal && n > 0 && (n & 3) == 0, KMU_ERR_BAD_ARG, "hybridloss_stats: bad argument (n must be a multiple of 4)");
  KMU_REQUIRE(workspace && workspace_bytes >= kmu_hybridloss_workspace_bytes(), KMU_ERR_WORKSPACE, "hybridloss_stats: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n4 = n / 4;
  const int ctas = (int)(cdiv(n4, 256) < NPART ? cdiv(n4, 256) : NPART);
  stats_kernel<<<ctas, 256, 0, st>>>((const float4*)pred, (const float4*)target, n4, (float*)workspace);
  KMU_LAUNCH_CHECK("hybridloss_stats");
  stats_fin_kernel<<<1, 256, 0, st>>>((const float*)workspace, ctas, scal);
  KMU_LAUNCH_CHECK("hybridloss_stats_fin");
  return KMU_OK;
}

int kmu_hybridloss_stack(const float* pred, const float* target, const float* scal, float* stack5, int64_t n, kmu_stream stream) {
  KMU_REQUIRE(pred && target && scal && stack5 && n > 0 && (n & 3) == 0, KMU_ERR_BAD_ARG, "hybridloss_stack: bad argument");
  const long long n4 = n / 4;
  stack_kernel<<<(unsigned)cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)pred, (const float4*)target, scal, (float4*)stack5, n4);
  KMU_LAUNCH_CHECK("hybridloss_stack");
  return KMU_OK;
}

int kmu_hybridloss_ssim(const float* filtered5, const float* scal, float* gm3, float* loss_out, int64_t n_valid, int64_t n_full, float alpha,
                        float c1, float c2, void* workspace, size_t workspace_bytes, kmu_stream stream) {
  KMU_REQUIRE(filtered5 && scal && gm3 && loss_out && n_valid > 0 && n_full > 0, KMU_ERR_BAD_ARG, "hybridloss_ssim: bad argument");
  KMU_REQUIRE(workspace && workspace_bytes >= kmu_hybridloss_workspace_bytes(), KMU_ERR_WORKSPACE, "hybridloss_ssim: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int ctas = (int)(cdiv(n_valid, 256) < NPART ? cdiv(n_valid, 256) : NPART);
  ssim_kernel<<<ctas, 256, 0, st>>>(filtered5, gm3, (float*)workspace, n_valid, c1, c2, -(1.f - alpha) / (float)n_valid);
  KMU_LAUNCH_CHECK("hybridloss_ssim");
  loss_fin_kernel<<<1, 256, 0, st>>>((const float*)workspace, ctas, scal, loss_out, alpha, 1.f / (float)n_full, 1.f / (float)n_valid);
  KMU_LAUNCH_CHECK("hybridloss_fin");
  return KMU_OK;
}

int kmu_hybridloss_bwd(const float* pred, const float* target, const float* scal, const float* dstack3, const float* grad_out, float* dpred,
                       int64_t n, float alpha, kmu_stream stream) {
  KMU_REQUIRE(pred && target && scal && dstack3 && grad_out && dpred && n > 0 && (n & 3) == 0, KMU_ERR_BAD_ARG, "hybridloss_bwd: bad argument");
  const long long n4 = n / 4;
  bwd_kernel<<<(unsigned)cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)pred, (const float4*)target, scal, (const float4*)dstack3,
                                                                      grad_out, (float4*)dpred, n4, alpha, 1.f / (float)n);
  KMU_LAUNCH_CHECK("hybridloss_bwd");
  return KMU_OK;
}

}  // extern "C"
