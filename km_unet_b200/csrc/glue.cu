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

// ================================================================================================ wavelet pooling (IWP)
// WPL/iwp.py:116-132 with the Haar taps of :50-52 and the matrices of :58-103: for every 2x2 block [[a,b],[c,d]]
//   LL = (a+b+c+d)/2, LH = (a-b+c-d)/2, HL = (a+b-c-d)/2, HH = (a-b-c+d)/2, the LAST high-pass row and column are zero,
//   the Softmax2d over a one-channel map is identically 1, so out = fusion_conv([LL ; mean_c(LH,HL,HH)]).
// The reference rebuilds its numpy filter matrices and uploads them on every call (a host sync inside the model) and runs
// 6 matmuls + cat + conv + softmax + mean; here: one kernel forward, one backward (+ the weight-gradient reduction).
// thread = output pixel; LL of all C channels in registers; weights [C+1][C] transposed in shared memory.
template <int C>
__global__ void __launch_bounds__(128) iwp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wf,
                                                      const float* __restrict__ bias, float* __restrict__ out, int B, int H, int W) {
  __shared__ __align__(16) float w_s[(C + 1) * C];   // w_s[i][o] = wf[o][i]
  for (int i = threadIdx.x; i < (C + 1) * C; i += 128) {
    const int ii = i / C, o = i - ii * C;
    w_s[i] = wf[o * (C + 1) + ii];
  }
  __syncthreads();
  const int Ho = H >> 1, Wo = W >> 1;
  const long long idx = (long long)blockIdx.x * 128 + threadIdx.x;
  if (idx >= (long long)B * Ho * Wo) return;
  const int j = (int)(idx % Wo), i = (int)((idx / Wo) % Ho), b = (int)(idx / ((long long)Wo * Ho));
  const float mlh = j < Wo - 1 ? 1.f : 0.f, mhl = i < Ho - 1 ? 1.f : 0.f, mhh = mlh * mhl;
  const float* xp = x + ((size_t)b * C * H + 2 * i) * W + 2 * j;
  float acc[C];
#pragma unroll
  for (int o = 0; o < C; ++o) acc[o] = bias[o];
  float high = 0.f;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float2 r0 = *reinterpret_cast<const float2*>(xp + (size_t)c * H * W);
    const float2 r1 = *reinterpret_cast<const float2*>(xp + (size_t)c * H * W + W);
    const float ll = 0.5f * ((r0.x + r0.y) + (r1.x + r1.y));
    high += 0.5f * (mlh * ((r0.x - r0.y) + (r1.x - r1.y)) + mhl * ((r0.x + r0.y) - (r1.x + r1.y)) + mhh * ((r0.x - r0.y) - (r1.x - r1.y)));
    const float4* w4 = reinterpret_cast<const float4*>(w_s + c * C);
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
      const float4 w = w4[q];
      acc[4 * q + 0] = fmaf(w.x, ll, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(w.y, ll, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(w.z, ll, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(w.w, ll, acc[4 * q + 3]);
    }
  }
  high *= 1.0f / (3 * C);
  float* op = out + ((size_t)b * C * Ho + i) * Wo + j;
#pragma unroll
  for (int o = 0; o < C; ++o) op[(size_t)o * Ho * Wo] = fmaf(w_s[C * C + o], high, acc[o]);
}

// backward: dx and per-CTA partials of d wf [C][C+1] | d bias [C].  CTA = 64 output pixels, 128 threads = 64 pixels x 2 halves
// of the channels for the dx part; the outer products run over the staged tiles.
template <int C>
__global__ void __launch_bounds__(128) iwp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ wf,
                                                      const float* __restrict__ dout, float* __restrict__ dx,
                                                      float* __restrict__ partial, int B, int H, int W) {
  constexpr int PADP = 65;
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;                      // [C][C+1] natural
  float* g_s = w_s + C * (C + 1);         // [C][PADP]      dout tile
  float* in_s = g_s + C * PADP;           // [C+1][PADP]    LL tile | mean-high
  float* dav_s = in_s + (C + 1) * PADP;   // [64]           d(mean-high) per pixel
  const int tid = threadIdx.x, px = tid & 63, half = tid >> 6;
  for (int i = tid; i < C * (C + 1); i += 128) w_s[i] = wf[i];
  const int Ho = H >> 1, Wo = W >> 1;
  const long long idx = (long long)blockIdx.x * 64 + px;
  const bool valid = idx < (long long)B * Ho * Wo;
  const long long id2 = valid ? idx : 0;
  const int j = (int)(id2 % Wo), i = (int)((id2 / Wo) % Ho), b = (int)(id2 / ((long long)Wo * Ho));
  const float mlh = j < Wo - 1 ? 1.f : 0.f, mhl = i < Ho - 1 ? 1.f : 0.f, mhh = mlh * mhl;
  const float* xp = x + ((size_t)b * C * H + 2 * i) * W + 2 * j;
  const float* gp = dout + ((size_t)b * C * Ho + i) * Wo + j;
  // stage dout and LL (this half's channels); the high-pass mean needs all channels: both halves add their part
  float high = 0.f;
#pragma unroll 4
  for (int k = 0; k < C / 2; ++k) {
    const int c = half * (C / 2) + k;
    float ll = 0.f, g = 0.f;
    if (valid) {
      const float2 r0 = *reinterpret_cast<const float2*>(xp + (size_t)c * H * W);
      const float2 r1 = *reinterpret_cast<const float2*>(xp + (size_t)c * H * W + W);
      ll = 0.5f * ((r0.x + r0.y) + (r1.x + r1.y));
      high += 0.5f * (mlh * ((r0.x - r0.y) + (r1.x - r1.y)) + mhl * ((r0.x + r0.y) - (r1.x + r1.y)) + mhh * ((r0.x - r0.y) - (r1.x - r1.y)));
      g = __ldg(gp + (size_t)c * Ho * Wo);
    }
    g_s[c * PADP + px] = g;
    in_s[c * PADP + px] = ll;
  }
  if (half == 1) dav_s[px] = high;
  __syncthreads();
  if (half == 0) in_s[C * PADP + px] = (high + dav_s[px]) * (1.0f / (3 * C));
  __syncthreads();
  // d(mean-high) = sum_o wf[o][C] dout[o]
  if (half == 0) {
    float dv = 0.f;
#pragma unroll 8
    for (int o = 0; o < C; ++o) dv = fmaf(w_s[o * (C + 1) + C], g_s[o * PADP + px], dv);
    dav_s[px] = dv * (1.0f / (6 * C));
  }
  __syncthreads();
  if (valid) {
    const float dav = dav_s[px];
    const float ka = dav * (mlh + mhl + mhh), kb = dav * (-mlh + mhl - mhh), kc = dav * (mlh - mhl - mhh), kd = dav * (-mlh - mhl + mhh);
    float* dxp = dx + ((size_t)b * C * H + 2 * i) * W + 2 * j;
#pragma unroll 2
    for (int k = 0; k < C / 2; ++k) {
      const int c = half * (C / 2) + k;
      float dll = 0.f;
#pragma unroll 8
      for (int o = 0; o < C; ++o) dll = fmaf(w_s[o * (C + 1) + c], g_s[o * PADP + px], dll);
      dll *= 0.5f;
      *reinterpret_cast<float2*>(dxp + (size_t)c * H * W) = make_float2(dll + ka, dll + kb);
      *reinterpret_cast<float2*>(dxp + (size_t)c * H * W + W) = make_float2(dll + kc, dll + kd);
    }
  }
  // weight-gradient partials
  float* pb = partial + (size_t)blockIdx.x * (C * (C + 1) + C);
  for (int e = tid; e < C * (C + 1); e += 128) {
    const int o = e / (C + 1), ii = e - o * (C + 1);
    const float* ar = g_s + o * PADP;
    const float* br = in_s + ii * PADP;
    float a = 0.f;
#pragma unroll 8
    for (int p = 0; p < 64; ++p) a = fmaf(ar[p], br[p], a);
    pb[e] = a;
  }
  for (int o = tid; o < C; o += 128) {
    float a = 0.f;
    for (int p = 0; p < 64; ++p) a += g_s[o * PADP + p];
    pb[C * (C + 1) + o] = a;
  }
}

__global__ void __launch_bounds__(256) iwp_wreduce_kernel(const float* __restrict__ partial, int nparts, int n_w, int n_b,
                                                          float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float red[8][33];
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + o;
  float s = 0.f;
  if (idx < n_w + n_b)
    for (int k = sl; k < nparts; k += 8) s += partial[(size_t)k * (n_w + n_b) + idx];
  red[sl][o] = s;
  __syncthreads();
  if (sl == 0 && idx < n_w + n_b) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][o];
    if (idx < n_w) dw[idx] = t;
    else db[idx - n_w] = t;
  }
}

static int iwp_check(int B, int C, int H, int W, const char* who) {
  KMU_REQUIRE(B > 0 && H > 0 && W > 0, KMU_ERR_BAD_ARG, "%s: non-positive shape", who);
  KMU_REQUIRE(C == 16 || C == 32 || C == 64, KMU_ERR_UNSUPPORTED, "%s: C=%d not in {16,32,64}", who, C);
  KMU_REQUIRE((H & 1) == 0 && (W & 1) == 0, KMU_ERR_UNSUPPORTED, "%s: %dx%d map (even sizes only)", who, H, W);
  return KMU_OK;
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

namespace kmu {
namespace glue {

// ---------------------------------------------------------------------------------------------- combine3
// EnhancedViMBlock.forward (KM_UNetV3_SH.py:349-368): x + DropPath(g0 f0 + g1 f1 + g2 f2) with per-sample gate weights g (softmax
// of the fusion gate) and the per-sample DropPath factor folded into coef[b][i]:   out = x + sum_i coef[b][i] f_i.
// One pass over five tensors instead of three broadcast multiplies, a mask multiply and three adds.
__global__ void __launch_bounds__(256) combine3_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ f0,
                                                           const float4* __restrict__ f1, const float4* __restrict__ f2,
                                                           const float* __restrict__ coef, float4* __restrict__ out, long long n4_per_b,
                                                           long long total4) {
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total4) return;
  const int b = (int)(i / n4_per_b);
  const float c0 = __ldg(coef + b * 3), c1 = __ldg(coef + b * 3 + 1), c2 = __ldg(coef + b * 3 + 2);
  const float4 xv = __ldg(x + i), a = __ldg(f0 + i), bb = __ldg(f1 + i), c = __ldg(f2 + i);
  float4 o;
  o.x = xv.x + c0 * a.x + c1 * bb.x + c2 * c.x;
  o.y = xv.y + c0 * a.y + c1 * bb.y + c2 * c.y;
  o.z = xv.z + c0 * a.z + c1 * bb.z + c2 * c.z;
  o.w = xv.w + c0 * a.w + c1 * bb.w + c2 * c.w;
  out[i] = o;
}

// df_i = coef[b][i] dy ; dcoef partials: part[b][chunk][i] = sum over the chunk of dy . f_i.   grid (chunks, B), 256 threads.
__global__ void __launch_bounds__(256) combine3_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ f0,
                                                           const float4* __restrict__ f1, const float4* __restrict__ f2,
                                                           const float* __restrict__ coef, float4* __restrict__ df0,
                                                           float4* __restrict__ df1, float4* __restrict__ df2, float* __restrict__ part,
                                                           long long n4_per_b, int per_cta) {
  __shared__ float red[8];
  const int b = blockIdx.y;
  const float c0 = __ldg(coef + b * 3), c1 = __ldg(coef + b * 3 + 1), c2 = __ldg(coef + b * 3 + 2);
  const long long base = (long long)b * n4_per_b;
  const long long lo = (long long)blockIdx.x * per_cta;
  long long hi = lo + per_cta;
  if (hi > n4_per_b) hi = n4_per_b;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (long long j = lo + threadIdx.x; j < hi; j += 256) {
    const long long i = base + j;
    const float4 g = __ldg(dy + i), a = __ldg(f0 + i), bb = __ldg(f1 + i), c = __ldg(f2 + i);
    s0 += g.x * a.x + g.y * a.y + g.z * a.z + g.w * a.w;
    s1 += g.x * bb.x + g.y * bb.y + g.z * bb.z + g.w * bb.w;
    s2 += g.x * c.x + g.y * c.y + g.z * c.z + g.w * c.w;
    df0[i] = make_float4(c0 * g.x, c0 * g.y, c0 * g.z, c0 * g.w);
    df1[i] = make_float4(c1 * g.x, c1 * g.y, c1 * g.z, c1 * g.w);
    df2[i] = make_float4(c2 * g.x, c2 * g.y, c2 * g.z, c2 * g.w);
  }
  s0 = block_sum256(s0, red);
  s1 = block_sum256(s1, red);
  s2 = block_sum256(s2, red);
  if (threadIdx.x == 0) {
    float* p = part + ((size_t)b * gridDim.x + blockIdx.x) * 3;
    p[0] = s0; p[1] = s1; p[2] = s2;
  }
}

__global__ void combine3_reduce_kernel(const float* __restrict__ part, int chunks, float* __restrict__ dcoef, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (b, i)
  if (idx >= n) return;
  const int b = idx / 3, i = idx - b * 3;
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += part[((size_t)b * chunks + k) * 3 + i];
  dcoef[idx] = s;
}

// ---------------------------------------------------------------------------------------------- bilinear resize (align_corners)
// F.interpolate(x, size, mode='bilinear', align_corners=True) of the skip connections (KM_UNetV3_SH.py:493-512).  ATen's forward
// kernel parallelises over the output pixels of ONE plane and loops over batch x channels inside (4 CTAs, 370 us for a 2 MB result);
// here a thread is one output element.  src = dst * (in - 1) / (out - 1), the +1 neighbour clamped at the border, as ATen computes it.
__global__ void __launch_bounds__(256) resize_bilinear_ac_kernel(const float* __restrict__ x, float* __restrict__ out, long long planes,
                                                                 int H, int W, int OH, int OW, float sh, float sw) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = planes * OH * OW;
  if (idx >= total) return;
  const int ow = (int)(idx % OW);
  const long long t = idx / OW;
  const int oh = (int)(t % OH);
  const long long plane = t / OH;
  const float fy = sh * oh, fx = sw * ow;
  const int h1 = (int)fy, w1 = (int)fx;
  const int hp = h1 < H - 1 ? 1 : 0, wp = w1 < W - 1 ? 1 : 0;
  const float lh1 = fy - h1, lh0 = 1.f - lh1, lw1 = fx - w1, lw0 = 1.f - lw1;
  const float* p = x + plane * H * W + (size_t)h1 * W + w1;
  out[idx] = lh0 * (lw0 * __ldg(p) + lw1 * __ldg(p + wp)) + lh1 * (lw0 * __ldg(p + hp * W) + lw1 * __ldg(p + hp * W + wp));
}

// ---------------------------------------------------------------------------------------------- GroupNorm forward
// nn.GroupNorm of StableHybridKANConv.pre_norm (4 groups), MultiScaleFusion (1 group) and the output norm (KM_UNetV3_SH.py:72-94,
// 292,455).  ATen's statistics kernel runs ONE CTA per (sample, group) -- 32 CTAs streaming 42 MB at the output norm; here every
// (sample, group) row is split over GN_SPLIT CTAs, a tiny kernel finishes mean / rstd in double, the apply kernel is float4.
// mean / rstd are returned so that the caller can hand them to the library's native_group_norm_backward.
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, float* __restrict__ part, long long row_len, int split) {
  __shared__ float red[8];
  const long long row = blockIdx.y;
  const long long per = ((row_len / 4 + split - 1) / split) * 4;      // slice length, multiple of 4
  const long long lo = (long long)blockIdx.x * per;
  long long hi = lo + per;
  if (hi > row_len) hi = row_len;
  const float* p = x + row * row_len;
  float s = 0.f, q = 0.f;
  for (long long i = lo + 4 * threadIdx.x; i < hi; i += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(p + i);
    s += (v.x + v.y) + (v.z + v.w);
    q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  s = block_sum256(s, red);
  q = block_sum256(q, red);
  if (threadIdx.x == 0) {
    part[(row * split + blockIdx.x) * 2] = s;
    part[(row * split + blockIdx.x) * 2 + 1] = q;
  }
}

__global__ void gn_fin_kernel(const float* __restrict__ part, float* __restrict__ mean, float* __restrict__ rstd, int rows, int split,
                              double inv_n, float eps) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < split; ++k) {
    s += (double)part[((size_t)row * split + k) * 2];
    q += (double)part[((size_t)row * split + k) * 2 + 1];
  }
  const double m = s * inv_n;
  double var = q * inv_n - m * m;
  if (var < 0.0) var = 0.0;
  mean[row] = (float)m;
  rstd[row] = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const float4* __restrict__ x, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float4* __restrict__ y, int C, int G, long long hw4,
                                                       long long total4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total4) return;
  const long long plane = i / hw4;                 // b * C + c
  const int c = (int)(plane % C);
  const long long row = (plane / C) * G + c / (C / G);
  const float m = __ldg(mean + row), r = __ldg(rstd + row);
  const float a = r * (gamma ? __ldg(gamma + c) : 1.f), b = (beta ? __ldg(beta + c) : 0.f) - m * a;
  const float4 v = __ldg(x + i);
  y[i] = make_float4(fmaf(v.x, a, b), fmaf(v.y, a, b), fmaf(v.z, a, b), fmaf(v.w, a, b));
}

static int gn_split(int rows, long long row_len) {
  int split = (148 * 4 + rows - 1) / rows;
  const long long maxs = (row_len + 4095) / 4096;
  if (split > maxs) split = (int)maxs;
  return split < 1 ? 1 : split;
}

// ---------------------------------------------------------------------------------------------- layer-scale mix (lerp)
// EfficientViMBlock.forward (vim_block_init/efficient_vim_init.py:89-90): x <- (1 - sigmoid(alpha_c)) x + sigmoid(alpha_c) mixer(x).
// torch.lerp with a broadcast tensor weight costs one forward and ~6 backward kernels (two products, the weight gradient, its
// reduction to (1,C,1,1), the sigmoid backward); here one float4 pass per direction plus a fixed-order reduction of d(alpha).
__global__ void __launch_bounds__(256) lerpmix_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ m,
                                                          const float* __restrict__ alpha, float4* __restrict__ y, int C, long long hw4,
                                                          long long total4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total4) return;
  const int c = (int)((i / hw4) % C);
  const float s = 1.f / (1.f + __expf(-__ldg(alpha + c)));
  const float4 a = __ldg(x + i), b = __ldg(m + i);
  y[i] = make_float4(a.x + s * (b.x - a.x), a.y + s * (b.y - a.y), a.z + s * (b.z - a.z), a.w + s * (b.w - a.w));
}

// grid (chunks, B * C): dx = (1 - s) dy, dm = s dy, part[plane][chunk] = sum dy (m - x)
__global__ void __launch_bounds__(256) lerpmix_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ m,
                                                          const float4* __restrict__ dy, const float* __restrict__ alpha,
                                                          float4* __restrict__ dx, float4* __restrict__ dm, float* __restrict__ part,
                                                          int C, long long hw4, int per_cta) {
  __shared__ float red[8];
  const long long plane = blockIdx.y;
  const int c = (int)(plane % C);
  const float s = 1.f / (1.f + __expf(-__ldg(alpha + c))), r = 1.f - s;
  const long long lo = (long long)blockIdx.x * per_cta;
  long long hi = lo + per_cta;
  if (hi > hw4) hi = hw4;
  float acc = 0.f;
  for (long long j = lo + threadIdx.x; j < hi; j += 256) {
    const long long i = plane * hw4 + j;
    const float4 a = __ldg(x + i), b = __ldg(m + i), g = __ldg(dy + i);
    acc += g.x * (b.x - a.x) + g.y * (b.y - a.y) + g.z * (b.z - a.z) + g.w * (b.w - a.w);
    dx[i] = make_float4(r * g.x, r * g.y, r * g.z, r * g.w);
    dm[i] = make_float4(s * g.x, s * g.y, s * g.z, s * g.w);
  }
  acc = block_sum256(acc, red);
  if (threadIdx.x == 0) part[plane * gridDim.x + blockIdx.x] = acc;
}

// dalpha[c] = s (1 - s) sum over b and chunks; one warp per channel, fixed order
__global__ void __launch_bounds__(128) lerpmix_reduce_kernel(const float* __restrict__ part, const float* __restrict__ alpha, int B, int C,
                                                             int chunks, float* __restrict__ dalpha) {
  const int c = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  float acc = 0.f;
  const int n = B * chunks;
  for (int i = lane; i < n; i += 32) {
    const int b = i / chunks, k = i - b * chunks;
    acc += part[((size_t)b * C + c) * chunks + k];
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const float s = 1.f / (1.f + __expf(-alpha[c]));
    dalpha[c] = s * (1.f - s) * acc;
  }
}

static int lerpmix_chunks(int planes, long long hw4) {
  int chunks = (148 * 8 + planes - 1) / planes;
  const long long maxc = (hw4 + 255) / 256;
  if (chunks > maxc) chunks = (int)maxc;
  return chunks < 1 ? 1 : chunks;
}

static int combine3_chunks(int B, long long n4_per_b) {
  int chunks = (148 * 8 + B - 1) / B;
  const long long maxc = (n4_per_b + 1023) / 1024;
  if (chunks > maxc) chunks = (int)maxc;
  return chunks < 1 ? 1 : chunks;
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

size_t kmu_iwp_bwd_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W) {
  if (iwp_check(B, C, H, W, "iwp_bwd_workspace_bytes") != KMU_OK) return 0;
  const long long npix = (long long)B * (H / 2) * (W / 2);
  return align_up((size_t)cdiv(npix, 64) * ((size_t)C * (C + 1) + C) * 4, 256);
}

int kmu_iwp_fwd(const float* x, const float* wf, const float* bias, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                kmu_stream stream) {
  int rc = iwp_check(B, C, H, W, "iwp_fwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(x && wf && bias && out, KMU_ERR_BAD_ARG, "iwp_fwd: null tensor");
  const long long npix = (long long)B * (H / 2) * (W / 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 16) iwp_fwd_kernel<16><<<cdiv(npix, 128), 128, 0, st>>>(x, wf, bias, out, B, H, W);
  else if (C == 32) iwp_fwd_kernel<32><<<cdiv(npix, 128), 128, 0, st>>>(x, wf, bias, out, B, H, W);
  else iwp_fwd_kernel<64><<<cdiv(npix, 128), 128, 0, st>>>(x, wf, bias, out, B, H, W);
  KMU_LAUNCH_CHECK("iwp_fwd");
  return KMU_OK;
}

int kmu_iwp_bwd(const float* x, const float* wf, const float* dout, float* dx, float* d_wf, float* d_bias, int32_t B, int32_t C,
                int32_t H, int32_t W, void* workspace, size_t workspace_bytes, kmu_stream stream) {
  int rc = iwp_check(B, C, H, W, "iwp_bwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(x && wf && dout && dx && d_wf && d_bias, KMU_ERR_BAD_ARG, "iwp_bwd: null tensor");
  KMU_REQUIRE(workspace && workspace_bytes >= kmu_iwp_bwd_workspace_bytes(B, C, H, W), KMU_ERR_WORKSPACE, "iwp_bwd: workspace too small");
  const long long npix = (long long)B * (H / 2) * (W / 2);
  const int nblk = cdiv(npix, 64);
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
#define KMU_IWP_BWD(CC)                                                                                          \
  do {                                                                                                           \
    size_t smem = ((size_t)CC * (CC + 1) + (size_t)CC * 65 + (size_t)(CC + 1) * 65 + 64) * 4;                    \
    if (smem > 48 * 1024) cudaFuncSetAttribute(iwp_bwd_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    iwp_bwd_kernel<CC><<<nblk, 128, smem, st>>>(x, wf, dout, dx, partial, B, H, W);                              \
  } while (0)
  if (C == 16) KMU_IWP_BWD(16);
  else if (C == 32) KMU_IWP_BWD(32);
  else KMU_IWP_BWD(64);
#undef KMU_IWP_BWD
  KMU_LAUNCH_CHECK("iwp_bwd");
  const int n_w = C * (C + 1), n_b = C;
  iwp_wreduce_kernel<<<cdiv(n_w + n_b, 32), 256, 0, st>>>(partial, nblk, n_w, n_b, d_wf, d_bias);
  KMU_LAUNCH_CHECK("iwp_wreduce");
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

size_t kmu_combine3_bwd_workspace_bytes(int32_t B, int64_t n_per_b) {
  if (B <= 0 || n_per_b <= 0 || (n_per_b & 3)) return 0;
  return align_up((size_t)B * kmu::glue::combine3_chunks(B, n_per_b / 4) * 3 * 4, 256);
}

int kmu_combine3_fwd(const float* x, const float* f0, const float* f1, const float* f2, const float* coef, float* out, int32_t B,
                     int64_t n_per_b, kmu_stream stream) {
  KMU_REQUIRE(x && f0 && f1 && f2 && coef && out && B > 0 && n_per_b > 0, KMU_ERR_BAD_ARG, "combine3_fwd: bad argument");
  KMU_REQUIRE((n_per_b & 3) == 0, KMU_ERR_UNSUPPORTED, "combine3_fwd: elements per sample (%lld) must be a multiple of 4", (long long)n_per_b);
  const long long n4 = n_per_b / 4, total4 = n4 * B;
  kmu::glue::combine3_fwd_kernel<<<(unsigned)cdiv(total4, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)x, (const float4*)f0, (const float4*)f1, (const float4*)f2, coef, (float4*)out, n4, total4);
  KMU_LAUNCH_CHECK("combine3_fwd");
  return KMU_OK;
}

int kmu_combine3_bwd(const float* dy, const float* f0, const float* f1, const float* f2, const float* coef, float* df0, float* df1,
                     float* df2, float* dcoef, int32_t B, int64_t n_per_b, void* workspace, size_t workspace_bytes, kmu_stream stream) {
  KMU_REQUIRE(dy && f0 && f1 && f2 && coef && df0 && df1 && df2 && dcoef && B > 0 && n_per_b > 0, KMU_ERR_BAD_ARG, "combine3_bwd: bad argument");
  KMU_REQUIRE((n_per_b & 3) == 0 && B <= 65535, KMU_ERR_UNSUPPORTED, "combine3_bwd: unsupported shape");
  KMU_REQUIRE(workspace && workspace_bytes >= kmu_combine3_bwd_workspace_bytes(B, n_per_b), KMU_ERR_WORKSPACE, "combine3_bwd: workspace too small");
  const long long n4 = n_per_b / 4;
  const int chunks = kmu::glue::combine3_chunks(B, n4);
  const int per_cta = (int)cdiv(n4, chunks);
  cudaStream_t st = (cudaStream_t)stream;
  kmu::glue::combine3_bwd_kernel<<<dim3(chunks, B), 256, 0, st>>>((const float4*)dy, (const float4*)f0, (const float4*)f1, (const float4*)f2,
                                                                 coef, (float4*)df0, (float4*)df1, (float4*)df2, (float*)workspace, n4, per_cta);
  KMU_LAUNCH_CHECK("combine3_bwd");
  kmu::glue::combine3_reduce_kernel<<<cdiv(B * 3, 128), 128, 0, st>>>((const float*)workspace, chunks, dcoef, B * 3);
  KMU_LAUNCH_CHECK("combine3_reduce");
  return KMU_OK;
}

int kmu_resize_bilinear_ac_fwd(const float* x, float* out, int64_t planes, int32_t H, int32_t W, int32_t OH, int32_t OW,
                               kmu_stream stream) {
  KMU_REQUIRE(x && out && planes > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, KMU_ERR_BAD_ARG, "resize_bilinear_ac_fwd: bad argument");
  const float sh = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f, sw = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  const long long total = (long long)planes * OH * OW;
  kmu::glue::resize_bilinear_ac_kernel<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, out, planes, H, W, OH, OW, sh, sw);
  KMU_LAUNCH_CHECK("resize_bilinear_ac_fwd");
  return KMU_OK;
}

size_t kmu_groupnorm_fwd_workspace_bytes(int32_t B, int32_t C, int64_t HW, int32_t G) {
  if (B <= 0 || C <= 0 || HW <= 0 || G <= 0 || C % G) return 0;
  return align_up((size_t)B * G * kmu::glue::gn_split(B * G, (long long)(C / G) * HW) * 2 * 4, 256);
}

int kmu_groupnorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int32_t B, int32_t C,
                      int64_t HW, int32_t G, float eps, void* workspace, size_t workspace_bytes, kmu_stream stream) {
  KMU_REQUIRE(x && y && mean && rstd && B > 0 && C > 0 && HW > 0 && G > 0 && C % G == 0, KMU_ERR_BAD_ARG, "groupnorm_fwd: bad argument");
  KMU_REQUIRE((HW & 3) == 0 && (long long)B * G <= 65535, KMU_ERR_UNSUPPORTED, "groupnorm_fwd: HW must be a multiple of 4, B*G <= 65535");
  KMU_REQUIRE(workspace && workspace_bytes >= kmu_groupnorm_fwd_workspace_bytes(B, C, HW, G), KMU_ERR_WORKSPACE, "groupnorm_fwd: workspace too small");
  const int rows = B * G;
  const long long row_len = (long long)(C / G) * HW;
  const int split = kmu::glue::gn_split(rows, row_len);
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)workspace;
  kmu::glue::gn_stats_kernel<<<dim3(split, rows), 256, 0, st>>>(x, part, row_len, split);
  KMU_LAUNCH_CHECK("gn_stats");
  kmu::glue::gn_fin_kernel<<<cdiv(rows, 128), 128, 0, st>>>(part, mean, rstd, rows, split, 1.0 / (double)row_len, eps);
  KMU_LAUNCH_CHECK("gn_fin");
  const long long total4 = (long long)B * C * HW / 4;
  kmu::glue::gn_apply_kernel<<<(unsigned)cdiv(total4, 256), 256, 0, st>>>((const float4*)x, mean, rstd, gamma, beta, (float4*)y, C, G, HW / 4, total4);
  KMU_LAUNCH_CHECK("gn_apply");
  return KMU_OK;
}

size_t kmu_lerpmix_bwd_workspace_bytes(int32_t B, int32_t C, int64_t HW) {
  if (B <= 0 || C <= 0 || HW <= 0 || (HW & 3)) return 0;
  return align_up((size_t)B * C * kmu::glue::lerpmix_chunks(B * C, HW / 4) * 4, 256);
}

int kmu_lerpmix_fwd(const float* x, const float* m, const float* alpha, float* y, int32_t B, int32_t C, int64_t HW, kmu_stream stream) {
  KMU_REQUIRE(x && m && alpha && y && B > 0 && C > 0 && HW > 0 && (HW & 3) == 0, KMU_ERR_BAD_ARG, "lerpmix_fwd: bad argument (HW multiple of 4)");
  const long long hw4 = HW / 4, total4 = hw4 * B * C;
  kmu::glue::lerpmix_fwd_kernel<<<(unsigned)cdiv(total4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)m, alpha,
                                                                                           (float4*)y, C, hw4, total4);
  KMU_LAUNCH_CHECK("lerpmix_fwd");
  return KMU_OK;
}

int kmu_lerpmix_bwd(const float* x, const float* m, const float* dy, const float* alpha, float* dx, float* dm, float* dalpha, int32_t B,
                    int32_t C, int64_t HW, void* workspace, size_t workspace_bytes, kmu_stream stream) {
  KMU_REQUIRE(x && m && dy && alpha && dx && dm && dalpha && B > 0 && C > 0 && HW > 0 && (HW & 3) == 0, KMU_ERR_BAD_ARG, "lerpmix_bwd: bad argument");
  KMU_REQUIRE((long long)B * C <= 65535, KMU_ERR_UNSUPPORTED, "lerpmix_bwd: B*C > 65535");
  KMU_REQUIRE(workspace && workspace_bytes >= kmu_lerpmix_bwd_workspace_bytes(B, C, HW), KMU_ERR_WORKSPACE, "lerpmix_bwd: workspace too small");
  const long long hw4 = HW / 4;
  const int chunks = kmu::glue::lerpmix_chunks(B * C, hw4);
  const int per_cta = (int)cdiv(hw4, chunks);
  cudaStream_t st = (cudaStream_t)stream;
  kmu::glue::lerpmix_bwd_kernel<<<dim3(chunks, B * C), 256, 0, st>>>((const float4*)x, (const float4*)m, (const float4*)dy, alpha, (float4*)dx,
                                                                   (float4*)dm, (float*)workspace, C, hw4, per_cta);
  KMU_LAUNCH_CHECK("lerpmix_bwd");
  kmu::glue::lerpmix_reduce_kernel<<<cdiv(C, 4), 128, 0, st>>>((const float*)workspace, alpha, B, C, chunks, dalpha);
  KMU_LAUNCH_CHECK("lerpmix_reduce");
  return KMU_OK;
}

}  // extern "C"
