// glue.cu -- fused elementwise / normalisation kernels of the callers that sit between the hot operators
// (SURVEY section 8f rank 2): TripleNorm and the DirectionAttention gate of KM_UNetV3_SH.py.
//
// TripleNorm (KM_UNetV3_SH.py:266-284): (GroupNorm(1,C)(x^T)^T + GroupNorm(1,C)(x) + LayerNorm_C(x)) / 3.  GroupNorm(1)
// statistics run over (C,H,W) of a sample and do not see the H/W transposition, so both GroupNorms share them; the
// LayerNorm runs over the C values of one pixel.  PyTorch executes this as two GroupNorms, a permute + contiguous copy +
// LayerNorm + permute back, two adds and a divide (about 10 kernels forward, twice that backward, the permuted LayerNorm
// alone 385 us at (32,16,128,128)).  Here: one statistics pass + one apply pass per direction, thread = pixel with the C
// channel values in registers (coalesced NCHW plane reads), all reductions in fixed order.
//
// DirectionAttention gate (:259-261): attn = sigmoid(q k) v on the three channel thirds of the qkv tensor -- one
// float4 streaming kernel per direction instead of 3 forward / 7 backward elementwise kernels.
#include "common.cuh"

namespace kmu {
namespace glue {

__device__ __forceinline__ float block_sum256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) a += red[i];
  return a;  // valid in every thread
}

// ================================================================================================ TripleNorm
constexpr int TN_SPLIT = 16;   // CTAs per sample in the statistics kernels

// per-sample partial sums: part[(b*TN_SPLIT + s)*2 + {0,1}] = sum x, sum x^2 over the CTA's slice of the (C*HW) sample
__global__ void __launch_bounds__(256) tn_stats_kernel(const float* __restrict__ x, float* __restrict__ part, long long n) {
  __shared__ float red[8];
  const int b = blockIdx.y, s = blockIdx.x;
  const long long len = ((n + TN_SPLIT - 1) / TN_SPLIT + 3) / 4 * 4;
  const long long v0 = s * len;
  long long v1 = v0 + len;
  if (v1 > n) v1 = n;
  const float* xb = x + (size_t)b * n;
  float a = 0.f, q = 0.f;
  if ((n & 3) == 0) {
    for (long long v = v0 + 4 * threadIdx.x; v < v1; v += 1024) {
      const float4 t = *reinterpret_cast<const float4*>(xb + v);
      a += (t.x + t.y) + (t.z + t.w);
      q = fmaf(t.x, t.x, fmaf(t.y, t.y, fmaf(t.z, t.z, fmaf(t.w, t.w, q))));
    }
  } else {
    for (long long v = v0 + threadIdx.x; v < v1; v += 256) {
      const float t = xb[v];
      a += t;
      q = fmaf(t, t, q);
    }
  }
  a = block_sum256(a, red);
  q = block_sum256(q, red);
  if (threadIdx.x == 0) {
    part[((size_t)b * TN_SPLIT + s) * 2] = a;
    part[((size_t)b * TN_SPLIT + s) * 2 + 1] = q;
  }
}

__device__ __forceinline__ float2 tn_sample_stat(const float* __restrict__ part, int b, long long n, float eps) {
  double s = 0.0, q = 0.0;
#pragma unroll
  for (int i = 0; i < TN_SPLIT; ++i) {
    s += (double)part[((size_t)b * TN_SPLIT + i) * 2];
    q += (double)part[((size_t)b * TN_SPLIT + i) * 2 + 1];
  }
  const double mean = s / (double)n;
  double var = q / (double)n - mean * mean;
  if (var < 0.0) var = 0.0;
  return make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}

// forward apply.  grid (ceil(HW/256), B), thread = pixel.  gstat[b] = (mean, rstd) of the sample (saved for backward)
template <int C>
__global__ void __launch_bounds__(256) tn_apply_kernel(const float* __restrict__ x, const float* __restrict__ part,
                                                       const float* __restrict__ gh, const float* __restrict__ bh,
                                                       const float* __restrict__ gw, const float* __restrict__ bw,
                                                       const float* __restrict__ gc, const float* __restrict__ bc,
                                                       float* __restrict__ y, float2* __restrict__ gstat, int HW, float eps_g,
                                                       float eps_l) {
  __shared__ float p_s[3][C];   // Gamma = gh+gw, Beta = bh+bw+bc, gc
  __shared__ float2 st_s;
  const int b = blockIdx.y;
  if (threadIdx.x < C) {
    p_s[0][threadIdx.x] = gh[threadIdx.x] + gw[threadIdx.x];
    p_s[1][threadIdx.x] = bh[threadIdx.x] + bw[threadIdx.x] + bc[threadIdx.x];
    p_s[2][threadIdx.x] = gc[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    st_s = tn_sample_stat(part, b, (long long)C * HW, eps_g);
    if (blockIdx.x == 0) gstat[b] = st_s;
  }
  __syncthreads();
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const float2 st = st_s;
  const float* xp = x + (size_t)b * C * HW + p;
  float v[C];
  float m = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    v[c] = __ldg(xp + (size_t)c * HW);
    m += v[c];
  }
  m *= (1.0f / C);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float dlt = v[c] - m;
    var = fmaf(dlt, dlt, var);
  }
  const float rp = rsqrtf(var * (1.0f / C) + eps_l);
  float* yp = y + (size_t)b * C * HW + p;
  const float third = 1.0f / 3.0f;
#pragma unroll
  for (int c = 0; c < C; ++c)
    yp[(size_t)c * HW] = (p_s[0][c] * (v[c] - st.x) * st.y + p_s[2][c] * (v[c] - m) * rp + p_s[1][c]) * third;
}

// backward statistics.  grid (TN_SPLIT, B).  Per CTA: sample sums (sum u, sum u xg) with u = g Gamma_c, and per-channel sums
// (sum g xg, sum g, sum g xl), g = dy / 3.  part2[(b*TN_SPLIT+s)*(2+3C)]: [su, sux, gxg[C], g[C], gxl[C]]
template <int C>
__global__ void __launch_bounds__(256) tn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float2* __restrict__ gstat, const float* __restrict__ gh,
                                                           const float* __restrict__ gw, float* __restrict__ part2, int HW,
                                                           float eps_l) {
  __shared__ float red[8];
  __shared__ float G_s[C];
  __shared__ float acc_s[8][3 * C];
  const int b = blockIdx.y, s = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x < C) G_s[threadIdx.x] = gh[threadIdx.x] + gw[threadIdx.x];
  __syncthreads();
  const float2 st = gstat[b];
  const int len = (HW + TN_SPLIT - 1) / TN_SPLIT;
  const int p0 = s * len;
  const int p1 = p0 + len < HW ? p0 + len : HW;
  const float* xb = x + (size_t)b * C * HW;
  const float* gb = dy + (size_t)b * C * HW;
  float su = 0.f, sux = 0.f;
  // channels in chunks of 16 so the per-channel accumulators stay in registers; the pixel's LayerNorm statistics are
  // recomputed per chunk from the (L1/L2-resident) channel column
  for (int c0 = 0; c0 < C; c0 += 16) {
    float a0[16], a1[16], a2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a0[k] = a1[k] = a2[k] = 0.f;
    for (int p = p0 + threadIdx.x; p < p1; p += 256) {
      float m = 0.f, q = 0.f;
#pragma unroll 8
      for (int c = 0; c < C; ++c) {
        const float t = __ldg(xb + (size_t)c * HW + p);
        m += t;
        q = fmaf(t, t, q);
      }
      m *= (1.0f / C);
      const float rp = rsqrtf(fmaxf(q * (1.0f / C) - m * m, 0.f) + eps_l);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int c = c0 + k;
        const float t = __ldg(xb + (size_t)c * HW + p);
        const float g = __ldg(gb + (size_t)c * HW + p) * (1.0f / 3.0f);
        const float xg = (t - st.x) * st.y;
        a0[k] = fmaf(g, xg, a0[k]);
        a1[k] += g;
        a2[k] = fmaf(g, (t - m) * rp, a2[k]);
        const float u = g * G_s[c];
        su += u;
        sux = fmaf(u, xg, sux);
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float r0 = warp_sum(a0[k]), r1 = warp_sum(a1[k]), r2 = warp_sum(a2[k]);
      if (lane == 0) {
        acc_s[wid][c0 + k] = r0;
        acc_s[wid][C + c0 + k] = r1;
        acc_s[wid][2 * C + c0 + k] = r2;
      }
    }
  }
  su = block_sum256(su, red);
  sux = block_sum256(sux, red);
  float* out = part2 + ((size_t)b * TN_SPLIT + s) * (2 + 3 * C);
  if (threadIdx.x == 0) { out[0] = su; out[1] = sux; }
  for (int i = threadIdx.x; i < 3 * C; i += 256) {
    float a = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) a += acc_s[wv][i];
    out[2 + i] = a;
  }
}

// backward apply: dx = r_b (u - mean u - xg mean(u xg)) + r_p (w - mean_c w - xl mean_c(w xl)), u = g Gamma, w = g gc
template <int C>
__global__ void __launch_bounds__(256) tn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float2* __restrict__ gstat, const float* __restrict__ part2,
                                                           const float* __restrict__ gh, const float* __restrict__ gw,
                                                           const float* __restrict__ gc, float* __restrict__ dx, int HW, float eps_l) {
  __shared__ float p_s[2][C];
  __shared__ float mu_s[2];
  const int b = blockIdx.y;
  if (threadIdx.x < C) {
    p_s[0][threadIdx.x] = gh[threadIdx.x] + gw[threadIdx.x];
    p_s[1][threadIdx.x] = gc[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    double su = 0.0, sux = 0.0;
#pragma unroll
    for (int i = 0; i < TN_SPLIT; ++i) {
      su += (double)part2[((size_t)b * TN_SPLIT + i) * (2 + 3 * C)];
      sux += (double)part2[((size_t)b * TN_SPLIT + i) * (2 + 3 * C) + 1];
    }
    const double n = (double)C * HW;
    mu_s[0] = (float)(su / n);
    mu_s[1] = (float)(sux / n);
  }
  __syncthreads();
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const float2 st = gstat[b];
  const float mu = mu_s[0], mux = mu_s[1];
  const float* xp = x + (size_t)b * C * HW + p;
  const float* gp = dy + (size_t)b * C * HW + p;
  float v[C], g[C];
  float m = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    v[c] = __ldg(xp + (size_t)c * HW);
    g[c] = __ldg(gp + (size_t)c * HW) * (1.0f / 3.0f);
    m += v[c];
  }
  m *= (1.0f / C);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float dlt = v[c] - m;
    var = fmaf(dlt, dlt, var);
  }
  const float rp = rsqrtf(var * (1.0f / C) + eps_l);
  float mw = 0.f, mwx = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float w = g[c] * p_s[1][c];
    mw += w;
    mwx = fmaf(w, (v[c] - m) * rp, mwx);
  }
  mw *= (1.0f / C);
  mwx *= (1.0f / C);
  float* dxp = dx + (size_t)b * C * HW + p;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float xg = (v[c] - st.x) * st.y, xl = (v[c] - m) * rp;
    const float u = g[c] * p_s[0][c], w = g[c] * p_s[1][c];
    dxp[(size_t)c * HW] = st.y * (u - mu - xg * mux) + rp * (w - mw - xl * mwx);
  }
}

// parameter gradients: d gh = d gw = sum g xg ; d bh = d bw = d bc = sum g ; d gc = sum g xl.  thread = (kind, channel)
__global__ void __launch_bounds__(256) tn_param_reduce_kernel(const float* __restrict__ part2, int nparts, int C, float* __restrict__ d_gh,
                                                              float* __restrict__ d_bh, float* __restrict__ d_gw, float* __restrict__ d_bw,
                                                              float* __restrict__ d_gc, float* __restrict__ d_bc) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= 3 * C) return;
  double s = 0.0;
  for (int k = 0; k < nparts; ++k) s += (double)part2[(size_t)k * (2 + 3 * C) + 2 + i];
  const int kind = i / C, c = i - kind * C;
  const float v = (float)s;
  if (kind == 0) { d_gh[c] = v; d_gw[c] = v; }
  else if (kind == 1) { d_bh[c] = v; d_bw[c] = v; d_bc[c] = v; }
  else d_gc[c] = v;
}

// ================================================================================================ qkv gate
// attn[b,c,p] = sigmoid(q k) v with q,k,v = qkv[b, c | C+c | 2C+c, p].  One thread per float4 of the output.
template <bool VEC>
__global__ void __launch_bounds__(256) gate_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, int C, int HW,
                                                       long long total) {
  long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * (VEC ? 4 : 1);
  if (i >= total) return;
  const long long plane = (long long)C * HW;
  const long long b = i / plane, r = i - b * plane;
  const float* q = qkv + b * 3 * plane + r;
  if (VEC) {
    const float4 a = *reinterpret_cast<const float4*>(q), k = *reinterpret_cast<const float4*>(q + plane),
                 v = *reinterpret_cast<const float4*>(q + 2 * plane);
    *reinterpret_cast<float4*>(out + i) = make_float4(sigmoidf_(a.x * k.x) * v.x, sigmoidf_(a.y * k.y) * v.y,
                                                     sigmoidf_(a.z * k.z) * v.z, sigmoidf_(a.w * k.w) * v.w);
  } else {
    out[i] = sigmoidf_(q[0] * q[plane]) * q[2 * plane];
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) gate_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                       float* __restrict__ dqkv, int C, int HW, long long total) {
  long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * (VEC ? 4 : 1);
  if (i >= total) return;
  const long long plane = (long long)C * HW;
  const long long b = i / plane, r = i - b * plane;
  const float* q = qkv + b * 3 * plane + r;
  float* dq = dqkv + b * 3 * plane + r;
  auto one = [](float qv, float kv, float vv, float g, float& o_q, float& o_k, float& o_v) {
    const float s = sigmoidf_(qv * kv);
    const float t = g * vv * s * (1.f - s);
    o_q = t * kv;
    o_k = t * qv;
    o_v = g * s;
  };
  if (VEC) {
    const float4 a = *reinterpret_cast<const float4*>(q), k = *reinterpret_cast<const float4*>(q + plane),
                 v = *reinterpret_cast<const float4*>(q + 2 * plane), g = *reinterpret_cast<const float4*>(dout + i);
    float4 oq, ok, ov;
    one(a.x, k.x, v.x, g.x, oq.x, ok.x, ov.x);
    one(a.y, k.y, v.y, g.y, oq.y, ok.y, ov.y);
    one(a.z, k.z, v.z, g.z, oq.z, ok.z, ov.z);
    one(a.w, k.w, v.w, g.w, oq.w, ok.w, ov.w);
    *reinterpret_cast<float4*>(dq) = oq;
    *reinterpret_cast<float4*>(dq + plane) = ok;
    *reinterpret_cast<float4*>(dq + 2 * plane) = ov;
  } else {
    one(q[0], q[plane], q[2 * plane], dout[i], dq[0], dq[plane], dq[2 * plane]);
  }
}

static int tn_check(const kmu_triplenorm_desc* d, const char* who) {
  KMU_REQUIRE(d != nullptr, KMU_ERR_BAD_ARG, "%s: null descriptor", who);
  KMU_REQUIRE(d->B > 0 && d->C > 0 && d->HW > 0, KMU_ERR_BAD_ARG, "%s: non-positive shape", who);
  KMU_REQUIRE(d->C == 16 || d->C == 32 || d->C == 64, KMU_ERR_UNSUPPORTED, "%s: C=%d not in {16,32,64}", who, d->C);
  KMU_REQUIRE(d->B <= 65535, KMU_ERR_UNSUPPORTED, "%s: B=%d > 65535", who, d->B);
  return KMU_OK;
}

}  // namespace glue
}  // namespace kmu

using namespace kmu;
using namespace kmu::glue;

extern "C" {

size_t kmu_triplenorm_workspace_bytes(const kmu_triplenorm_desc* d) {
  if (tn_check(d, "triplenorm_workspace_bytes") != KMU_OK) return 0;
  return align_up((size_t)d->B * TN_SPLIT * (2 + 3 * (size_t)d->C) * 4, 256);
}

int kmu_triplenorm_fwd(const kmu_triplenorm_fwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "triplenorm_fwd: null args");
  int rc = tn_check(&a->d, "triplenorm_fwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(a->x && a->gh && a->bh && a->gw && a->bw && a->gc && a->bc && a->y && a->gstat, KMU_ERR_BAD_ARG, "triplenorm_fwd: null tensor");
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= kmu_triplenorm_workspace_bytes(&a->d), KMU_ERR_WORKSPACE, "triplenorm_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = a->d.B, C = a->d.C, HW = a->d.HW;
  float* part = (float*)a->workspace;
  tn_stats_kernel<<<dim3(TN_SPLIT, B), 256, 0, st>>>(a->x, part, (long long)C * HW);
  KMU_LAUNCH_CHECK("tn_stats");
  dim3 grid(cdiv(HW, 256), B);
#define KMU_TN_APPLY(CC) \
  tn_apply_kernel<CC><<<grid, 256, 0, st>>>(a->x, part, a->gh, a->bh, a->gw, a->bw, a->gc, a->bc, a->y, (float2*)a->gstat, HW, a->d.eps_gn, a->d.eps_ln)
  if (C == 16) KMU_TN_APPLY(16);
  else if (C == 32) KMU_TN_APPLY(32);
  else KMU_TN_APPLY(64);
#undef KMU_TN_APPLY
  KMU_LAUNCH_CHECK("tn_apply");
  return KMU_OK;
}

int kmu_triplenorm_bwd(const kmu_triplenorm_bwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "triplenorm_bwd: null args");
  int rc = tn_check(&a->d, "triplenorm_bwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(a->x && a->dy && a->gstat && a->gh && a->gw && a->gc && a->dx && a->d_gh && a->d_bh && a->d_gw && a->d_bw && a->d_gc && a->d_bc,
              KMU_ERR_BAD_ARG, "triplenorm_bwd: null tensor");
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= kmu_triplenorm_workspace_bytes(&a->d), KMU_ERR_WORKSPACE, "triplenorm_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = a->d.B, C = a->d.C, HW = a->d.HW;
  float* part2 = (float*)a->workspace;
  const float2* gstat = (const float2*)a->gstat;
  dim3 sgrid(TN_SPLIT, B), agrid(cdiv(HW, 256), B);
#define KMU_TN_BWD(CC)                                                                                                    \
  do {                                                                                                                    \
    tn_bwd_stats_kernel<CC><<<sgrid, 256, 0, st>>>(a->x, a->dy, gstat, a->gh, a->gw, part2, HW, a->d.eps_ln);             \
    KMU_LAUNCH_CHECK("tn_bwd_stats");                                                                                     \
    tn_bwd_apply_kernel<CC><<<agrid, 256, 0, st>>>(a->x, a->dy, gstat, part2, a->gh, a->gw, a->gc, a->dx, HW, a->d.eps_ln); \
    KMU_LAUNCH_CHECK("tn_bwd_apply");                                                                                     \
  } while (0)
  if (C == 16) KMU_TN_BWD(16);
  else if (C == 32) KMU_TN_BWD(32);
  else KMU_TN_BWD(64);
#undef KMU_TN_BWD
  tn_param_reduce_kernel<<<cdiv(3 * C, 256), 256, 0, st>>>(part2, B * TN_SPLIT, C, a->d_gh, a->d_bh, a->d_gw, a->d_bw, a->d_gc, a->d_bc);
  KMU_LAUNCH_CHECK("tn_param_reduce");
  return KMU_OK;
}

int kmu_qkv_gate_fwd(const float* qkv, float* out, int32_t B, int32_t C, int32_t HW, kmu_stream stream) {
  KMU_REQUIRE(qkv && out && B > 0 && C > 0 && HW > 0, KMU_ERR_BAD_ARG, "qkv_gate_fwd: bad argument");
  const long long total = (long long)B * C * HW;
  if ((HW & 3) == 0) gate_fwd_kernel<true><<<cdiv(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(qkv, out, C, HW, total);
  else gate_fwd_kernel<false><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(qkv, out, C, HW, total);
  KMU_LAUNCH_CHECK("qkv_gate_fwd");
  return KMU_OK;
}

int kmu_qkv_gate_bwd(const float* qkv, const float* dout, float* dqkv, int32_t B, int32_t C, int32_t HW, kmu_stream stream) {
  KMU_REQUIRE(qkv && dout && dqkv && B > 0 && C > 0 && HW > 0, KMU_ERR_BAD_ARG, "qkv_gate_bwd: bad argument");
  const long long total = (long long)B * C * HW;
  if ((HW & 3) == 0) gate_bwd_kernel<true><<<cdiv(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(qkv, dout, dqkv, C, HW, total);
  else gate_bwd_kernel<false><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(qkv, dout, dqkv, C, HW, total);
  KMU_LAUNCH_CHECK("qkv_gate_bwd");
  return KMU_OK;
}

}  // extern "C"
