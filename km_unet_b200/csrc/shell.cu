// shell.cu -- the EfficientViMBlock shell around HSM-SSD: depthwise 3x3 convolution and train-mode BatchNorm2d with its
// fused epilogues (ReLU, sigmoid layer-scale mix), forward and backward.
//
// Replaces vim_block_init/efficient_vim_init.py:82-96 + vim_utils_init.py:83-89,128-130: every
//     x <- (1 - sigmoid(alpha)) x + sigmoid(alpha) BN(conv(x))        (dwconv1, dwconv2, ffn.fc2)
//     h <- ReLU(BN(conv1x1(x)))                                       (ffn.fc1)
// is one kmu_bnmix call on the convolution's output, and the depthwise 3x3 itself is kmu_dwconv3x3 (also used by
// DirectionAttention.conv, KM_UNetV3_SH.py:223,262).  The reference issues cuDNN's per-channel BatchNorm kernels here
// (one CTA per channel: C = 16..64 CTAs on 148 SMs) and PyTorch's generic depthwise weight-gradient kernel; both are
// HBM-bound streaming ops, so the kernels below split every channel over many CTAs, move float4s and keep all
// reductions in a fixed order (per-CTA partials -> double-precision finalize; no atomics).
#include "common.cuh"

namespace kmu {
namespace shell {

// ================================================================================================ BatchNorm + epilogue
struct BnDims {
  int B, C, HW;
  int nsplit;      // CTAs per channel in the reduction kernels
  long long per;   // B*HW elements per channel
  long long len;   // elements per CTA (multiple of 4 when HW % 4 == 0)
};

// KMU_BN_ORDER=0 restores the plain traversal (channel-major statistics, ascending applies) for A/B timing
static int bn_order() {
  static const int v = [] {
    const char* e = getenv("KMU_BN_ORDER");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return v;
}

static BnDims bn_dims(const kmu_bnmix_desc& s) {
  BnDims d;
  d.B = s.B; d.C = s.C; d.HW = s.HW;
  d.per = (long long)s.B * s.HW;
  int want = cdiv(148 * 8, s.C);
  long long maxsplit = d.per / 2048;
  if (maxsplit < 1) maxsplit = 1;
  d.nsplit = (int)(want < maxsplit ? want : maxsplit);
  if (d.nsplit < 1) d.nsplit = 1;
  long long len = (d.per + d.nsplit - 1) / d.nsplit;
  d.len = (len + 3) / 4 * 4;
  d.nsplit = (int)((d.per + d.len - 1) / d.len);
  return d;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float a = 0.f;
  if (threadIdx.x < 32) {
    a = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    a = warp_sum(a);
  }
  return a;  // valid in warp 0
}

// element v of channel c -> offset in the (B,C,HW) tensor
__device__ __forceinline__ size_t chan_off(long long v, int c, int C, int HW) {
  long long b = v / HW;
  return (size_t)((b * C + c) * (long long)HW + (v - b * HW));
}

// ---- forward statistics: part[(c*nsplit + s)*2 + {0,1}] = sum, sum of squares of (x - k), k = the channel's first element.
//      The shift keeps var = E[(x-k)^2] - E[x-k]^2 free of the catastrophic cancellation of E[x^2] - mean^2 when |mean| >> std
//      (k is a sample of the channel, so |E[x-k]| is a few std at most); nn.BatchNorm's Welford has no such problem either.
//      grid (nsplit, C), 256 threads
//      Traversal order (`order` != 0, default): the passes over one tensor are laid out so that each starts where the previous one
//      ended and finds the tail of that pass in L2 (the FFN's hidden tensors are 134 MB against 126 MB of L2: re-reading them from the
//      start evicts every line just before it is needed).  The producing convolution writes batch-major ascending; the forward statistics
//      walk the splits DESCENDING with the split index slowest (all channels of the last samples first); the forward apply walks planes
//      ascending; the backward statistics ascending with the split index slowest; the backward apply descending.
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, float* __restrict__ part, BnDims d, int order) {
  const int c = order ? blockIdx.x : blockIdx.y, s = order ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.x;
  const long long v0 = (long long)s * d.len;
  long long v1 = v0 + d.len;
  if (v1 > d.per) v1 = d.per;
  const float k = __ldg(x + (size_t)c * d.HW);
  float a = 0.f, q = 0.f;
  if ((d.HW & 3) == 0) {
    // (sample b, offset in the plane) advance incrementally: the 64-bit division of chan_off per 16-byte load made these
    // kernels instruction-bound
    long long v = v0 + 4 * threadIdx.x;
    int b = (int)(v / d.HW), off = (int)(v - (long long)b * d.HW);
    for (; v < v1; v += 1024, off += 1024) {
      while (off >= d.HW) { off -= d.HW; ++b; }
      float4 t = *reinterpret_cast<const float4*>(x + ((size_t)b * d.C + c) * d.HW + off);
      t.x -= k; t.y -= k; t.z -= k; t.w -= k;
      a += (t.x + t.y) + (t.z + t.w);
      q = fmaf(t.x, t.x, fmaf(t.y, t.y, fmaf(t.z, t.z, fmaf(t.w, t.w, q))));
    }
  } else {
    for (long long v = v0 + threadIdx.x; v < v1; v += 256) {
      float t = x[chan_off(v, c, d.C, d.HW)] - k;
      a += t;
      q = fmaf(t, t, q);
    }
  }
  // one barrier for both sums: warp shuffles, then the eight warp partials in warp order
  __shared__ float red2[8][2];
  a = warp_sum(a);
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) { red2[threadIdx.x >> 5][0] = a; red2[threadIdx.x >> 5][1] = q; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += red2[w8][threadIdx.x];
    part[((size_t)c * d.nsplit + s) * 2 + threadIdx.x] = t;
  }
}

// ---- finalize: stat[c] = (mean, rstd); running statistics as torch.nn.BatchNorm2d (momentum, unbiased variance).  grid C, 32 thr
__global__ void __launch_bounds__(32) bn_fin_kernel(const float* __restrict__ x, const float* __restrict__ part, float2* __restrict__ stat,
                                                    float* __restrict__ rmean, float* __restrict__ rvar, BnDims d, int training,
                                                    float momentum, float eps) {
  const int c = blockIdx.x;
  double mean, var;
  if (training) {
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < d.nsplit; i += 32) {
      s += (double)part[((size_t)c * d.nsplit + i) * 2];
      q += (double)part[((size_t)c * d.nsplit + i) * 2 + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const double ms = s / (double)d.per;      // mean of the shifted values
    var = q / (double)d.per - ms * ms;
    if (var < 0.0) var = 0.0;
    mean = ms + (double)x[(size_t)c * d.HW];
  } else {
    mean = (double)rmean[c];
    var = (double)rvar[c];
  }
  if (threadIdx.x == 0) {
    stat[c] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    if (training && rmean && rvar) {
      double unb = d.per > 1 ? var * (double)d.per / (double)(d.per - 1) : var;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
    }
  }
}

// ---- apply: y = BN(x) [-> ReLU] [-> (1-a) res + a y].  One thread per float4 (or scalar) of the tensor.
template <bool VEC>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float2* __restrict__ stat,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ res, const float* __restrict__ alpha,
                                                       float* __restrict__ y, int C, int HW, long long total, int relu) {
  long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * (VEC ? 4 : 1);
  if (i >= total) return;
  const int c = (int)((i / HW) % C);
  const float2 st = stat[c];
  const float sc = gamma[c] * st.y, sh = beta[c] - st.x * sc;
  float a = 1.f;
  if (alpha) a = 1.f / (1.f + expf(-alpha[c]));
  if (VEC) {
    float4 t = *reinterpret_cast<const float4*>(x + i);
    t.x = fmaf(t.x, sc, sh); t.y = fmaf(t.y, sc, sh); t.z = fmaf(t.z, sc, sh); t.w = fmaf(t.w, sc, sh);
    if (relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
    if (res) {
      const float4 r = *reinterpret_cast<const float4*>(res + i);
      const float b = 1.f - a;
      t.x = fmaf(a, t.x, b * r.x); t.y = fmaf(a, t.y, b * r.y); t.z = fmaf(a, t.z, b * r.z); t.w = fmaf(a, t.w, b * r.w);
    }
    *reinterpret_cast<float4*>(y + i) = t;
  } else {
    float t = fmaf(x[i], sc, sh);
    if (relu) t = fmaxf(t, 0.f);
    if (res) t = fmaf(a, t, (1.f - a) * res[i]);
    y[i] = t;
  }
}

// Same, HW % 4 == 0: grid (chunks of 4096 elements, B * C planes) -- the channel is a property of the CTA, no per-thread divisions.
// The statistics finalize is folded in: warp 0 of EVERY CTA reduces its channel's partials in double (same order everywhere ->
// identical values) instead of a separate one-warp-per-channel kernel between the two passes; the CTA of (sample 0, chunk 0) writes
// stat[c] for the backward and updates the running statistics.
__global__ void __launch_bounds__(256) bn_apply_plane_kernel(const float* __restrict__ x, const float* __restrict__ part,
                                                             float2* __restrict__ stat, float* __restrict__ rmean,
                                                             float* __restrict__ rvar, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ res,
                                                             const float* __restrict__ alpha, float* __restrict__ y, BnDims d, int training,
                                                             float momentum, float eps, int relu) {
  __shared__ float2 st_s;
  const int plane = blockIdx.y, c = plane % d.C, HW = d.HW;
  if (threadIdx.x < 32) {
    double mean, var;
    if (training) {
      double s = 0.0, q = 0.0;
      for (int i = threadIdx.x; i < d.nsplit; i += 32) {
        s += (double)part[((size_t)c * d.nsplit + i) * 2];
        q += (double)part[((size_t)c * d.nsplit + i) * 2 + 1];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      const double ms = s / (double)d.per;
      var = q / (double)d.per - ms * ms;
      if (var < 0.0) var = 0.0;
      mean = ms + (double)x[(size_t)c * HW];
    } else {
      mean = (double)rmean[c];
      var = (double)rvar[c];
    }
    if (threadIdx.x == 0) {
      st_s = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
      if (plane == c && blockIdx.x == 0) {
        stat[c] = st_s;
        if (training && rmean && rvar) {
          const double unb = d.per > 1 ? var * (double)d.per / (double)(d.per - 1) : var;
          rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
          rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
        }
      }
    }
  }
  __syncthreads();
  const float2 st = st_s;
  const float sc = gamma[c] * st.y, sh = beta[c] - st.x * sc;
  const float a = alpha ? 1.f / (1.f + expf(-alpha[c])) : 1.f, b = 1.f - a;
  const size_t base = (size_t)plane * HW;
  const int n4 = HW >> 2;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = blockIdx.x * 1024 + k * 256 + threadIdx.x;
    if (i < n4) {
      float4 t = *reinterpret_cast<const float4*>(x + base + 4 * (size_t)i);
      t.x = fmaf(t.x, sc, sh); t.y = fmaf(t.y, sc, sh); t.z = fmaf(t.z, sc, sh); t.w = fmaf(t.w, sc, sh);
      if (relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
      if (res) {
        const float4 r = *reinterpret_cast<const float4*>(res + base + 4 * (size_t)i);
        t.x = fmaf(a, t.x, b * r.x); t.y = fmaf(a, t.y, b * r.y); t.z = fmaf(a, t.z, b * r.z); t.w = fmaf(a, t.w, b * r.w);
      }
      *reinterpret_cast<float4*>(y + base + 4 * (size_t)i) = t;
    }
  }
}

// ---- backward statistics: part[(c*nsplit+s)*3 + {0,1,2}] = sum g, sum g xhat, sum dy (bn_out - res)
//      g = a dy masked by the ReLU.  grid (nsplit, C), 256 threads
__global__ void __launch_bounds__(256) bn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float2* __restrict__ stat, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ res,
                                                           float* __restrict__ part, BnDims d, int relu, int order) {
  const int c = order ? blockIdx.x : blockIdx.y, s = order ? blockIdx.y : blockIdx.x;
  const long long v0 = (long long)s * d.len;
  long long v1 = v0 + d.len;
  if (v1 > d.per) v1 = d.per;
  const float2 st = stat[c];
  const float sc = gamma[c] * st.y, sh = beta[c] - st.x * sc;
  float sg = 0.f, sgx = 0.f, sa = 0.f;
  auto one = [&](float xv, float g, float rv) {
    float yb = fmaf(xv, sc, sh);
    if (relu) yb = fmaxf(yb, 0.f);
    sa = fmaf(g, yb - rv, sa);          // d alpha sees the unmasked upstream gradient
    if (relu && yb <= 0.f) g = 0.f;
    sg += g;
    sgx = fmaf(g, (xv - st.x) * st.y, sgx);
  };
  if ((d.HW & 3) == 0) {
    long long v = v0 + 4 * threadIdx.x;
    int b = (int)(v / d.HW), o = (int)(v - (long long)b * d.HW);
    for (; v < v1; v += 1024, o += 1024) {
      while (o >= d.HW) { o -= d.HW; ++b; }
      const size_t off = ((size_t)b * d.C + c) * d.HW + o;
      float4 xv = *reinterpret_cast<const float4*>(x + off);
      float4 g = *reinterpret_cast<const float4*>(dy + off);
      float4 r = res ? *reinterpret_cast<const float4*>(res + off) : make_float4(0.f, 0.f, 0.f, 0.f);
      one(xv.x, g.x, r.x); one(xv.y, g.y, r.y); one(xv.z, g.z, r.z); one(xv.w, g.w, r.w);
    }
  } else {
    for (long long v = v0 + threadIdx.x; v < v1; v += 256) {
      size_t off = chan_off(v, c, d.C, d.HW);
      one(x[off], dy[off], res ? res[off] : 0.f);
    }
  }
  __shared__ float red3[8][3];
  sg = warp_sum(sg);
  sgx = warp_sum(sgx);
  sa = warp_sum(sa);
  if ((threadIdx.x & 31) == 0) { red3[threadIdx.x >> 5][0] = sg; red3[threadIdx.x >> 5][1] = sgx; red3[threadIdx.x >> 5][2] = sa; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += red3[w8][threadIdx.x];
    part[((size_t)c * d.nsplit + s) * 3 + threadIdx.x] = t;
  }
}

// ---- backward finalize: dgamma, dbeta, dalpha and bstat[c] = (a mean g, a mean g xhat) (zeros in eval mode).  grid C, 32 threads
//      note: the sums above are of the UNSCALED upstream (dy masked); the layer-scale factor a is applied here.
__global__ void __launch_bounds__(32) bn_bwd_fin_kernel(const float* __restrict__ part, float2* __restrict__ bstat,
                                                        const float* __restrict__ alpha, float* __restrict__ dgamma,
                                                        float* __restrict__ dbeta, float* __restrict__ dalpha, BnDims d,
                                                        int training) {
  const int c = blockIdx.x;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < d.nsplit; i += 32) {
    const float* p = part + ((size_t)c * d.nsplit + i) * 3;
    s0 += (double)p[0];
    s1 += (double)p[1];
    s2 += (double)p[2];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (threadIdx.x == 0) {
    double a = 1.0;
    if (alpha) a = 1.0 / (1.0 + exp(-(double)alpha[c]));
    dbeta[c] = (float)(a * s0);
    dgamma[c] = (float)(a * s1);
    if (dalpha) dalpha[c] = (float)(s2 * a * (1.0 - a));
    bstat[c] = training ? make_float2((float)(a * s0 / (double)d.per), (float)(a * s1 / (double)d.per)) : make_float2(0.f, 0.f);
  }
}

// ---- backward apply: dx = gamma rstd (a g - mean - xhat mean') ; dres = (1 - a) dy
template <bool VEC>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float2* __restrict__ stat, const float2* __restrict__ bstat,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ alpha, float* __restrict__ dx,
                                                           float* __restrict__ dres, int C, int HW, long long total, int relu) {
  long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * (VEC ? 4 : 1);
  if (i >= total) return;
  const int c = (int)((i / HW) % C);
  const float2 st = stat[c], bs = bstat[c];
  const float sc = gamma[c] * st.y, sh = beta[c] - st.x * sc;
  float a = 1.f;
  if (alpha) a = 1.f / (1.f + expf(-alpha[c]));
  auto one = [&](float xv, float g) {
    if (relu && fmaf(xv, sc, sh) <= 0.f) g = 0.f;
    return sc * (a * g - bs.x - (xv - st.x) * st.y * bs.y);
  };
  if (VEC) {
    const float4 xv = *reinterpret_cast<const float4*>(x + i);
    const float4 g = *reinterpret_cast<const float4*>(dy + i);
    *reinterpret_cast<float4*>(dx + i) = make_float4(one(xv.x, g.x), one(xv.y, g.y), one(xv.z, g.z), one(xv.w, g.w));
    if (dres) {
      const float b = 1.f - a;
      *reinterpret_cast<float4*>(dres + i) = make_float4(b * g.x, b * g.y, b * g.z, b * g.w);
    }
  } else {
    const float g = dy[i];
    dx[i] = one(x[i], g);
    if (dres) dres[i] = (1.f - a) * g;
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_plane_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 const float2* __restrict__ stat, const float* __restrict__ part,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ alpha, float* __restrict__ dx,
                                                                 float* __restrict__ dres, float* __restrict__ dgamma,
                                                                 float* __restrict__ dbeta, float* __restrict__ dalpha, BnDims d, int training,
                                                                 int relu, int order) {
  // the backward finalize folded in (see bn_apply_plane_kernel): every CTA reduces its channel's three partial sums in double
  __shared__ float2 bs_s;
  const int plane = order ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y, c = plane % d.C, HW = d.HW;
  const int chunk = order ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const float a = alpha ? 1.f / (1.f + expf(-alpha[c])) : 1.f, b = 1.f - a;
  if (threadIdx.x < 32) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < d.nsplit; i += 32) {
      const float* p = part + ((size_t)c * d.nsplit + i) * 3;
      s0 += (double)p[0];
      s1 += (double)p[1];
      s2 += (double)p[2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (threadIdx.x == 0) {
      const double ad = alpha ? 1.0 / (1.0 + exp(-(double)alpha[c])) : 1.0;
      bs_s = training ? make_float2((float)(ad * s0 / (double)d.per), (float)(ad * s1 / (double)d.per)) : make_float2(0.f, 0.f);
      if (plane == c && chunk == 0) {
        dbeta[c] = (float)(ad * s0);
        dgamma[c] = (float)(ad * s1);
        if (dalpha) dalpha[c] = (float)(s2 * ad * (1.0 - ad));
      }
    }
  }
  __syncthreads();
  const float2 st = stat[c], bs = bs_s;
  const float sc = gamma[c] * st.y, sh = beta[c] - st.x * sc;
  auto one = [&](float xv, float g) {
    if (relu && fmaf(xv, sc, sh) <= 0.f) g = 0.f;
    return sc * (a * g - bs.x - (xv - st.x) * st.y * bs.y);
  };
  const size_t base = (size_t)plane * HW;
  const int n4 = HW >> 2;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = chunk * 1024 + k * 256 + threadIdx.x;
    if (i < n4) {
      const float4 xv = *reinterpret_cast<const float4*>(x + base + 4 * (size_t)i);
      const float4 g = *reinterpret_cast<const float4*>(dy + base + 4 * (size_t)i);
      *reinterpret_cast<float4*>(dx + base + 4 * (size_t)i) = make_float4(one(xv.x, g.x), one(xv.y, g.y), one(xv.z, g.z), one(xv.w, g.w));
      if (dres) *reinterpret_cast<float4*>(dres + base + 4 * (size_t)i) = make_float4(b * g.x, b * g.y, b * g.z, b * g.w);
    }
  }
}

static int bn_check(const kmu_bnmix_desc* d, const char* who) {
  KMU_REQUIRE(d != nullptr, KMU_ERR_BAD_ARG, "%s: null descriptor", who);
  KMU_REQUIRE(d->B > 0 && d->C > 0 && d->HW > 0, KMU_ERR_BAD_ARG, "%s: non-positive shape", who);
  KMU_REQUIRE(d->C <= 65535, KMU_ERR_UNSUPPORTED, "%s: C=%d > 65535", who, d->C);
  return KMU_OK;
}

// ================================================================================================ depthwise 3x3
struct DwDims {
  int B, C, H, W;
};

// y[b,c,h,w] = bias[c] + sum_{ky,kx} w[c][ky*3+kx] x[b,c,h+ky-1,w+kx-1] (zero padding).  FLIP: correlate with the flipped
// kernel (= the input gradient of the same convolution).  Thread = 4 consecutive output columns of one row.
// scale (optional, one factor per (b, c) plane) multiplies the result: y = scale * (conv + bias) forward, dx = scale * conv^T(dy).
// add (optional, may alias y) is added to the result: the residual branch's gradient joins dx here instead of in a separate kernel.
template <bool FLIP>
__global__ void __launch_bounds__(256) dw3x3_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                    const float* __restrict__ bias, const float* __restrict__ scale,
                                                    const float* add, float* y, DwDims d) {
  const int wq = (d.W + 3) >> 2;
  long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  long long total = (long long)d.B * d.C * d.H * wq;
  if (idx >= total) return;
  const int q = (int)(idx % wq);
  long long t = idx / wq;
  const int h = (int)(t % d.H);
  const long long plane = t / d.H;
  const int c = (int)(plane % d.C);
  const int w0 = q * 4;
  float k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + (FLIP ? 8 - i : i));
  const float* xp = x + (size_t)plane * d.H * d.W;
  const float bv = (bias && !FLIP) ? __ldg(bias + c) : 0.f;
  float acc[4] = {bv, bv, bv, bv};
  const bool vec = (d.W & 3) == 0;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int hh = h + ky - 1;
    if (hh < 0 || hh >= d.H) continue;
    const float* row = xp + (size_t)hh * d.W;
    float v[6];
    v[0] = w0 > 0 ? __ldg(row + w0 - 1) : 0.f;
    if (vec) {
      const float4 m = *reinterpret_cast<const float4*>(row + w0);
      v[1] = m.x; v[2] = m.y; v[3] = m.z; v[4] = m.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[1 + e] = (w0 + e < d.W) ? __ldg(row + w0 + e) : 0.f;
    }
    v[5] = (w0 + 4 < d.W) ? __ldg(row + w0 + 4) : 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      acc[e] = fmaf(k[ky * 3], v[e], fmaf(k[ky * 3 + 1], v[e + 1], fmaf(k[ky * 3 + 2], v[e + 2], acc[e])));
  }
  if (scale) {
    const float sc = __ldg(scale + plane);
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] *= sc;
  }
  float* yp = y + (size_t)plane * d.H * d.W + (size_t)h * d.W + w0;
  if (add) {
    const float* ap = add + (size_t)plane * d.H * d.W + (size_t)h * d.W + w0;
    if (vec) {
      const float4 m = *reinterpret_cast<const float4*>(ap);
      acc[0] += m.x; acc[1] += m.y; acc[2] += m.z; acc[3] += m.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (w0 + e < d.W) acc[e] += ap[e];
    }
  }
  if (vec) {
    *reinterpret_cast<float4*>(yp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (w0 + e < d.W) yp[e] = acc[e];
  }
}

// Row-sliding forms (W % 4 == 0).  CTA = (row slice, plane); thread = (column quad, sub-slice of the rows): it walks DOWN its rows with a
// three-row register window, so every new output row costs one row of loads (a 128-bit vector + the two neighbouring scalars)
// instead of three, and there is no index arithmetic beyond a row pointer increment.
struct Row6 { float v[6]; };
__device__ __forceinline__ Row6 load_row6(const float* __restrict__ plane, int h, int H, int W, int w0) {
  Row6 r;
  if (h < 0 || h >= H) {
#pragma unroll
    for (int e = 0; e < 6; ++e) r.v[e] = 0.f;
    return r;
  }
  const float* row = plane + (size_t)h * W + w0;
  const float4 m = *reinterpret_cast<const float4*>(row);
  r.v[0] = w0 > 0 ? __ldg(row - 1) : 0.f;
  r.v[1] = m.x; r.v[2] = m.y; r.v[3] = m.z; r.v[4] = m.w;
  r.v[5] = w0 + 4 < W ? __ldg(row + 4) : 0.f;
  return r;
}

__global__ void __launch_bounds__(256) dw3x3_rows_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ scale,
                                                             float* __restrict__ y, DwDims d, int rows_per_cta) {
  const int plane = blockIdx.y, c = plane % d.C;
  const int wq = d.W >> 2, nsub = 256 / wq;                 // wq divides 256 (host check)
  const int q = threadIdx.x % wq, sub = threadIdx.x / wq;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, d.H);
  const int per = (r1 - r0 + nsub - 1) / nsub;
  const int h0 = r0 + sub * per, h1 = min(h0 + per, r1);
  if (h0 >= h1) return;
  float k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
  const float bv = bias ? __ldg(bias + c) : 0.f;
  const float sc = scale ? __ldg(scale + plane) : 1.f;
  const float* xp = x + (size_t)plane * d.H * d.W;
  float* yp = y + (size_t)plane * d.H * d.W;
  const int w0 = q * 4;
  Row6 a = load_row6(xp, h0 - 1, d.H, d.W, w0), b = load_row6(xp, h0, d.H, d.W, w0);
  for (int h = h0; h < h1; ++h) {
    const Row6 cc = load_row6(xp, h + 1, d.H, d.W, w0);
    float acc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[e] = bv;
      acc[e] = fmaf(k[0], a.v[e], fmaf(k[1], a.v[e + 1], fmaf(k[2], a.v[e + 2], acc[e])));
      acc[e] = fmaf(k[3], b.v[e], fmaf(k[4], b.v[e + 1], fmaf(k[5], b.v[e + 2], acc[e])));
      acc[e] = fmaf(k[6], cc.v[e], fmaf(k[7], cc.v[e + 1], fmaf(k[8], cc.v[e + 2], acc[e])));
    }
    *reinterpret_cast<float4*>(yp + (size_t)h * d.W + w0) = make_float4(acc[0] * sc, acc[1] * sc, acc[2] * sc, acc[3] * sc);
    a = b;
    b = cc;
  }
}

// dx AND the weight / bias gradient partials with the same row walk: windows of x and dy (three rows each).
__global__ void __launch_bounds__(256) dw3x3_rows_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                             const float* __restrict__ w, const float* __restrict__ scale,
                                                             const float* add, float* dx, float* __restrict__ part, DwDims d,
                                                             int rows_per_cta) {
  __shared__ float red[8][10];
  const int plane = blockIdx.y, c = plane % d.C;
  const int wq = d.W >> 2, nsub = 256 / wq;
  const int q = threadIdx.x % wq, sub = threadIdx.x / wq;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, d.H);
  const int per = (r1 - r0 + nsub - 1) / nsub;
  const int h0 = r0 + sub * per, h1 = min(h0 + per, r1);
  float k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + 8 - i);        // flipped taps: dx = correlate(dy, flip(w))
  const float sc = scale ? __ldg(scale + plane) : 1.f;
  const float* xp = x + (size_t)plane * d.H * d.W;
  const float* gp = dy + (size_t)plane * d.H * d.W;
  const int w0 = q * 4;
  float a[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) a[i] = 0.f;
  if (h0 < h1) {
    Row6 xa = load_row6(xp, h0 - 1, d.H, d.W, w0), xb = load_row6(xp, h0, d.H, d.W, w0);
    Row6 ga = load_row6(gp, h0 - 1, d.H, d.W, w0), gb = load_row6(gp, h0, d.H, d.W, w0);
    for (int h = h0; h < h1; ++h) {
      const Row6 xc = load_row6(xp, h + 1, d.H, d.W, w0), gc = load_row6(gp, h + 1, d.H, d.W, w0);
      // weight gradient: dy of THIS row (gb, centre 4) against the 3x3 neighbourhood of x
      a[9] += (gb.v[1] + gb.v[2]) + (gb.v[3] + gb.v[4]);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        a[kx] += gb.v[1] * xa.v[kx] + gb.v[2] * xa.v[kx + 1] + gb.v[3] * xa.v[kx + 2] + gb.v[4] * xa.v[kx + 3];
        a[3 + kx] += gb.v[1] * xb.v[kx] + gb.v[2] * xb.v[kx + 1] + gb.v[3] * xb.v[kx + 2] + gb.v[4] * xb.v[kx + 3];
        a[6 + kx] += gb.v[1] * xc.v[kx] + gb.v[2] * xc.v[kx + 1] + gb.v[3] * xc.v[kx + 2] + gb.v[4] * xc.v[kx + 3];
      }
      if (dx) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[e] = fmaf(k[0], ga.v[e], fmaf(k[1], ga.v[e + 1], fmaf(k[2], ga.v[e + 2], acc[e])));
          acc[e] = fmaf(k[3], gb.v[e], fmaf(k[4], gb.v[e + 1], fmaf(k[5], gb.v[e + 2], acc[e])));
          acc[e] = fmaf(k[6], gc.v[e], fmaf(k[7], gc.v[e + 1], fmaf(k[8], gc.v[e + 2], acc[e])));
        }
        const size_t off = (size_t)plane * d.H * d.W + (size_t)h * d.W + w0;
        float4 o = make_float4(acc[0] * sc, acc[1] * sc, acc[2] * sc, acc[3] * sc);
        if (add) {
          const float4 m = *reinterpret_cast<const float4*>(add + off);
          o.x += m.x; o.y += m.y; o.z += m.z; o.w += m.w;
        }
        *reinterpret_cast<float4*>(dx + off) = o;
      }
      xa = xb; xb = xc;
      ga = gb; gb = gc;
    }
  }
  // one barrier for all ten sums: warp shuffles, then the eight warp partials are added in warp order
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const float s = warp_sum(a[i]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    float s = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8][threadIdx.x];
    part[((size_t)plane * gridDim.x + blockIdx.x) * 10 + threadIdx.x] = s;
  }
}
static bool dw_rows_ok(const DwDims& d) { return (d.W & 3) == 0 && d.W >= 16 && d.W <= 1024 && 256 % (d.W >> 2) == 0; }

// weight / bias gradient partials: part[(plane*RS + rs)*10 + t] = sum over the CTA's rows of dy(p) x(p + tap t) (t < 9), sum dy (t = 9)
// grid (RS, B*C), 256 threads; thread = 4 consecutive columns of a row, rows strided over the CTA's row slice
__global__ void __launch_bounds__(256) dw3x3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ part, DwDims d, int rows_per_cta) {
  __shared__ float red[8];
  const long long plane = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta;
  int r1 = r0 + rows_per_cta;
  if (r1 > d.H) r1 = d.H;
  const int wq = (d.W + 3) >> 2;
  const float* xp = x + (size_t)plane * d.H * d.W;
  const float* gp = dy + (size_t)plane * d.H * d.W;
  float a[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) a[i] = 0.f;
  const int items = (r1 - r0) * wq;
  for (int it = threadIdx.x; it < items; it += 256) {
    const int h = r0 + it / wq, w0 = (it % wq) * 4;
    float g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) g[e] = (w0 + e < d.W) ? __ldg(gp + (size_t)h * d.W + w0 + e) : 0.f;
    a[9] += (g[0] + g[1]) + (g[2] + g[3]);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      if (hh < 0 || hh >= d.H) continue;
      const float* row = xp + (size_t)hh * d.W;
      float v[6];
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        const int ww = w0 - 1 + e;
        v[e] = (ww >= 0 && ww < d.W) ? __ldg(row + ww) : 0.f;
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
        a[ky * 3 + kx] += g[0] * v[kx] + g[1] * v[kx + 1] + g[2] * v[kx + 2] + g[3] * v[kx + 3];
    }
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    float s = block_sum(a[i], red);
    if (threadIdx.x == 0) part[((size_t)plane * gridDim.x + blockIdx.x) * 10 + i] = s;
  }
}

// dx AND the weight / bias gradient partials in one pass (W % 4 == 0): thread = 4 consecutive columns of a row, rows strided over the
// CTA's row slice.  x and dy are each read once (three rows x (one 128-bit load + two scalars) per item) instead of dy twice and x
// once by the dx kernel + the weight-gradient kernel.  dx = scale * conv^T(dy) [+ add]; partials as dw3x3_wgrad_kernel writes them.
__global__ void __launch_bounds__(256) dw3x3_bwd_fused_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              const float* __restrict__ w, const float* __restrict__ scale,
                                                              const float* add, float* dx, float* __restrict__ part, DwDims d,
                                                              int rows_per_cta) {
  __shared__ float red[8];
  const long long plane = blockIdx.y;
  const int c = (int)(plane % d.C);
  const int r0 = blockIdx.x * rows_per_cta;
  int r1 = r0 + rows_per_cta;
  if (r1 > d.H) r1 = d.H;
  const int wq = d.W >> 2;
  const float* xp = x + (size_t)plane * d.H * d.W;
  const float* gp = dy + (size_t)plane * d.H * d.W;
  float k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + 8 - i);        // flipped taps: dx = correlate(dy, flip(w))
  const float sc = scale ? __ldg(scale + plane) : 1.f;
  float a[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) a[i] = 0.f;
  const int items = (r1 - r0) * wq;
  for (int it = threadIdx.x; it < items; it += 256) {
    const int h = r0 + it / wq, w0 = (it % wq) * 4;
    float g[3][6], v[3][6];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      const bool ok = hh >= 0 && hh < d.H;
      const float* grow = gp + (size_t)hh * d.W + w0;
      const float* xrow = xp + (size_t)hh * d.W + w0;
      const float4 gm = ok ? *reinterpret_cast<const float4*>(grow) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 xm = ok ? *reinterpret_cast<const float4*>(xrow) : make_float4(0.f, 0.f, 0.f, 0.f);
      g[ky][0] = (ok && w0 > 0) ? __ldg(grow - 1) : 0.f;
      v[ky][0] = (ok && w0 > 0) ? __ldg(xrow - 1) : 0.f;
      g[ky][1] = gm.x; g[ky][2] = gm.y; g[ky][3] = gm.z; g[ky][4] = gm.w;
      v[ky][1] = xm.x; v[ky][2] = xm.y; v[ky][3] = xm.z; v[ky][4] = xm.w;
      g[ky][5] = (ok && w0 + 4 < d.W) ? __ldg(grow + 4) : 0.f;
      v[ky][5] = (ok && w0 + 4 < d.W) ? __ldg(xrow + 4) : 0.f;
    }
    // weight gradient: dy of THIS row against the 3x3 neighbourhood of x
    a[9] += (g[1][1] + g[1][2]) + (g[1][3] + g[1][4]);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
        a[ky * 3 + kx] += g[1][1] * v[ky][kx] + g[1][2] * v[ky][kx + 1] + g[1][3] * v[ky][kx + 2] + g[1][4] * v[ky][kx + 3];
    if (dx) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          acc[e] = fmaf(k[ky * 3], g[ky][e], fmaf(k[ky * 3 + 1], g[ky][e + 1], fmaf(k[ky * 3 + 2], g[ky][e + 2], acc[e])));
      const size_t off = (size_t)plane * d.H * d.W + (size_t)h * d.W + w0;
      float4 o = make_float4(acc[0] * sc, acc[1] * sc, acc[2] * sc, acc[3] * sc);
      if (add) {
        const float4 m = *reinterpret_cast<const float4*>(add + off);
        o.x += m.x; o.y += m.y; o.z += m.z; o.w += m.w;
      }
      *reinterpret_cast<float4*>(dx + off) = o;
    }
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    float s = block_sum(a[i], red);
    if (threadIdx.x == 0) part[((size_t)plane * gridDim.x + blockIdx.x) * 10 + i] = s;
  }
}

// dw[c][t] = sum over b and row slices; warp = (c, t), lanes stride over the B * RS partials, fixed-order tree (deterministic).
// With a per-plane scale (y = scale * (conv + bias)) every plane's partial is weighted by its scale.
__global__ void __launch_bounds__(128) dw3x3_wreduce_kernel(const float* __restrict__ part, const float* __restrict__ scale, int B, int C,
                                                            int RS, float* __restrict__ dw, float* __restrict__ dbias) {
  const int idx = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (idx >= C * 10) return;
  const int c = idx / 10, t = idx - c * 10;
  double s = 0.0;
  const int n = B * RS;
  for (int i = lane; i < n; i += 32) {
    const int b = i / RS, r = i - b * RS;
    const float sc = scale ? __ldg(scale + (size_t)b * C + c) : 1.f;
    s += (double)(sc * part[(((size_t)b * C + c) * RS + r) * 10 + t]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (t < 9) dw[c * 9 + t] = (float)s;
    else if (dbias) dbias[c] = (float)s;
  }
}

// dscale[b, c] = sum_p dy * (conv + bias) = sum_t w[c][t] part[b, c][t] + bias[c] part[b, c][9]: free from the same partials
__global__ void __launch_bounds__(256) dw3x3_dscale_kernel(const float* __restrict__ part, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int planes, int C, int RS,
                                                           float* __restrict__ dscale) {
  const int plane = blockIdx.x * 256 + threadIdx.x;
  if (plane >= planes) return;
  const int c = plane % C;
  float a[10];
#pragma unroll
  for (int t = 0; t < 10; ++t) a[t] = 0.f;
  for (int r = 0; r < RS; ++r)
#pragma unroll
    for (int t = 0; t < 10; ++t) a[t] += part[((size_t)plane * RS + r) * 10 + t];
  float s = bias ? __ldg(bias + c) * a[9] : 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) s = fmaf(__ldg(w + c * 9 + t), a[t], s);
  dscale[plane] = s;
}

static int dw_check(const kmu_dwconv3x3_desc* d, const char* who) {
  KMU_REQUIRE(d != nullptr, KMU_ERR_BAD_ARG, "%s: null descriptor", who);
  KMU_REQUIRE(d->B > 0 && d->C > 0 && d->H > 0 && d->W > 0, KMU_ERR_BAD_ARG, "%s: non-positive shape", who);
  KMU_REQUIRE((long long)d->B * d->C <= 65535, KMU_ERR_UNSUPPORTED, "%s: B*C=%lld > 65535", who, (long long)d->B * d->C);
  return KMU_OK;
}
static int dw_row_slices(const kmu_dwconv3x3_desc& s) {
  // enough CTAs to fill the machine: B*C planes x RS row slices ~ 148*8, at least 8 rows per slice
  int rs = cdiv(148 * 8, s.B * s.C);
  int maxrs = s.H / 8 > 0 ? s.H / 8 : 1;
  if (rs > maxrs) rs = maxrs;
  return rs < 1 ? 1 : rs;
}

}  // namespace shell
}  // namespace kmu

using namespace kmu;
using namespace kmu::shell;

extern "C" {

size_t kmu_bnmix_workspace_bytes(const kmu_bnmix_desc* dd) {
  if (bn_check(dd, "bnmix_workspace_bytes") != KMU_OK) return 0;
  BnDims d = bn_dims(*dd);
  return align_up((size_t)d.C * d.nsplit * 3 * 4, 256) + align_up((size_t)d.C * 8, 256);
}

int kmu_bnmix_fwd(const kmu_bnmix_fwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "bnmix_fwd: null args");
  int rc = bn_check(&a->d, "bnmix_fwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(a->x && a->weight && a->bias && a->y && a->stat, KMU_ERR_BAD_ARG, "bnmix_fwd: null tensor");
  KMU_REQUIRE(a->d.training || (a->running_mean && a->running_var), KMU_ERR_BAD_ARG, "bnmix_fwd: eval mode needs running statistics");
  KMU_REQUIRE(!a->d.mix || (a->res && a->alpha), KMU_ERR_BAD_ARG, "bnmix_fwd: mix needs res and alpha");
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= kmu_bnmix_workspace_bytes(&a->d), KMU_ERR_WORKSPACE, "bnmix_fwd: workspace too small");
  BnDims d = bn_dims(a->d);
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)a->workspace;
  float2* stat = (float2*)a->stat;
  if (a->d.training) {
    const int order = bn_order();
    bn_stats_kernel<<<order ? dim3(d.C, d.nsplit) : dim3(d.nsplit, d.C), 256, 0, st>>>(a->x, part, d, order);
    KMU_LAUNCH_CHECK("bn_stats");
  }
  const long long total = (long long)d.B * d.C * d.HW;
  const float* res = a->d.mix ? a->res : nullptr;
  const float* alpha = a->d.mix ? a->alpha : nullptr;
  const bool plane_path = (d.HW & 3) == 0 && (long long)d.B * d.C <= 65535;
  if (!plane_path) {
    bn_fin_kernel<<<d.C, 32, 0, st>>>(a->x, part, stat, a->running_mean, a->running_var, d, a->d.training, a->d.momentum, a->d.eps);
    KMU_LAUNCH_CHECK("bn_fin");
  }
  if (plane_path)
    bn_apply_plane_kernel<<<dim3(cdiv(d.HW, 4096), d.B * d.C), 256, 0, st>>>(a->x, part, stat, a->running_mean, a->running_var, a->weight,
                                                                             a->bias, res, alpha, a->y, d, a->d.training, a->d.momentum,
                                                                             a->d.eps, a->d.relu);
  else if ((d.HW & 3) == 0)
    bn_apply_kernel<true><<<cdiv(total / 4, 256), 256, 0, st>>>(a->x, stat, a->weight, a->bias, res, alpha, a->y, d.C, d.HW, total, a->d.relu);
  else
    bn_apply_kernel<false><<<cdiv(total, 256), 256, 0, st>>>(a->x, stat, a->weight, a->bias, res, alpha, a->y, d.C, d.HW, total, a->d.relu);
  KMU_LAUNCH_CHECK("bn_apply");
  return KMU_OK;
}

int kmu_bnmix_bwd(const kmu_bnmix_bwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "bnmix_bwd: null args");
  int rc = bn_check(&a->d, "bnmix_bwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(a->x && a->dy && a->weight && a->bias && a->stat && a->dx && a->d_weight && a->d_bias, KMU_ERR_BAD_ARG,
              "bnmix_bwd: null tensor");
  KMU_REQUIRE(!a->d.mix || (a->res && a->alpha && a->d_res && a->d_alpha), KMU_ERR_BAD_ARG, "bnmix_bwd: mix needs res, alpha, d_res, d_alpha");
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= kmu_bnmix_workspace_bytes(&a->d), KMU_ERR_WORKSPACE, "bnmix_bwd: workspace too small");
  BnDims d = bn_dims(a->d);
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)a->workspace;
  float2* bstat = (float2*)((char*)a->workspace + align_up((size_t)d.C * d.nsplit * 3 * 4, 256));
  const float2* stat = (const float2*)a->stat;
  const float* res = a->d.mix ? a->res : nullptr;
  const float* alpha = a->d.mix ? a->alpha : nullptr;
  const int order = bn_order();
  bn_bwd_stats_kernel<<<order ? dim3(d.C, d.nsplit) : dim3(d.nsplit, d.C), 256, 0, st>>>(a->x, a->dy, stat, a->weight, a->bias, res, part, d,
                                                                                          a->d.relu, order);
  KMU_LAUNCH_CHECK("bn_bwd_stats");
  const long long total = (long long)d.B * d.C * d.HW;
  float* dres = a->d.mix ? a->d_res : nullptr;
  const bool plane_path = (d.HW & 3) == 0 && (long long)d.B * d.C <= 65535;
  if (!plane_path) {
    bn_bwd_fin_kernel<<<d.C, 32, 0, st>>>(part, bstat, alpha, a->d_weight, a->d_bias, a->d.mix ? a->d_alpha : nullptr, d, a->d.training);
    KMU_LAUNCH_CHECK("bn_bwd_fin");
  }
  if (plane_path)
    bn_bwd_apply_plane_kernel<<<dim3(cdiv(d.HW, 4096), d.B * d.C), 256, 0, st>>>(a->x, a->dy, stat, part, a->weight, a->bias, alpha, a->dx,
                                                                                 dres, a->d_weight, a->d_bias,
                                                                                 a->d.mix ? a->d_alpha : nullptr, d, a->d.training, a->d.relu,
                                                                                 order);
  else if ((d.HW & 3) == 0)
    bn_bwd_apply_kernel<true><<<cdiv(total / 4, 256), 256, 0, st>>>(a->x, a->dy, stat, bstat, a->weight, a->bias, alpha, a->dx, dres, d.C,
                                                                    d.HW, total, a->d.relu);
  else
    bn_bwd_apply_kernel<false><<<cdiv(total, 256), 256, 0, st>>>(a->x, a->dy, stat, bstat, a->weight, a->bias, alpha, a->dx, dres, d.C,
                                                                 d.HW, total, a->d.relu);
  KMU_LAUNCH_CHECK("bn_bwd_apply");
  return KMU_OK;
}

size_t kmu_dwconv3x3_bwd_workspace_bytes(const kmu_dwconv3x3_desc* dd) {
  if (dw_check(dd, "dwconv3x3_bwd_workspace_bytes") != KMU_OK) return 0;
  return align_up((size_t)dd->B * dd->C * dw_row_slices(*dd) * 10 * 4, 256);
}

static int dw_fwd(const kmu_dwconv3x3_desc* dd, const float* x, const float* w, const float* bias, const float* scale, float* y,
                  kmu_stream stream) {
  int rc = dw_check(dd, "dwconv3x3_fwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(x && w && y, KMU_ERR_BAD_ARG, "dwconv3x3_fwd: null tensor");
  DwDims d{dd->B, dd->C, dd->H, dd->W};
  if (dw_rows_ok(d) && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0) {
    const int rs = dw_row_slices(*dd);
    const int rows = cdiv(d.H, rs);
    dw3x3_rows_fwd_kernel<<<dim3(cdiv(d.H, rows), d.B * d.C), 256, 0, (cudaStream_t)stream>>>(x, w, bias, scale, y, d, rows);
    KMU_LAUNCH_CHECK("dw3x3_fwd");
    return KMU_OK;
  }
  long long total = (long long)d.B * d.C * d.H * ((d.W + 3) / 4);
  dw3x3_kernel<false><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, w, bias, scale, nullptr, y, d);
  KMU_LAUNCH_CHECK("dw3x3_fwd");
  return KMU_OK;
}

static int dw_bwd(const kmu_dwconv3x3_desc* dd, const float* x, const float* dy, const float* w, const float* bias, const float* scale,
                  const float* dx_add, float* dx, float* dw, float* dbias, float* dscale, void* workspace, size_t workspace_bytes,
                  kmu_stream stream) {
  int rc = dw_check(dd, "dwconv3x3_bwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(x && dy && w, KMU_ERR_BAD_ARG, "dwconv3x3_bwd: null tensor");
  DwDims d{dd->B, dd->C, dd->H, dd->W};
  cudaStream_t st = (cudaStream_t)stream;
  const bool fused = (dw || dscale) && (d.W & 3) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 &&
                     (!dx || ((uintptr_t)dx & 15) == 0) && (!dx_add || ((uintptr_t)dx_add & 15) == 0);
  if (dx && !fused) {
    long long total = (long long)d.B * d.C * d.H * ((d.W + 3) / 4);
    dw3x3_kernel<true><<<cdiv(total, 256), 256, 0, st>>>(dy, w, nullptr, scale, dx_add, dx, d);
    KMU_LAUNCH_CHECK("dw3x3_bwd_dx");
  }
  if (dw || dscale) {
    KMU_REQUIRE(workspace && workspace_bytes >= kmu_dwconv3x3_bwd_workspace_bytes(dd), KMU_ERR_WORKSPACE, "dwconv3x3_bwd: workspace too small");
    const int rs = dw_row_slices(*dd);
    const int rows = cdiv(d.H, rs);
    const int rs2 = cdiv(d.H, rows);
    if (fused && dw_rows_ok(d))
      dw3x3_rows_bwd_kernel<<<dim3(rs2, d.B * d.C), 256, 0, st>>>(x, dy, w, scale, dx_add, dx, (float*)workspace, d, rows);
    else if (fused)
      dw3x3_bwd_fused_kernel<<<dim3(rs2, d.B * d.C), 256, 0, st>>>(x, dy, w, scale, dx_add, dx, (float*)workspace, d, rows);
    else
      dw3x3_wgrad_kernel<<<dim3(rs2, d.B * d.C), 256, 0, st>>>(x, dy, (float*)workspace, d, rows);
    KMU_LAUNCH_CHECK("dw3x3_wgrad");
    if (dw) {
      dw3x3_wreduce_kernel<<<cdiv(d.C * 10, 4), 128, 0, st>>>((const float*)workspace, scale, d.B, d.C, rs2, dw, dbias);
      KMU_LAUNCH_CHECK("dw3x3_wreduce");
    }
    if (dscale) {
      dw3x3_dscale_kernel<<<cdiv(d.B * d.C, 256), 256, 0, st>>>((const float*)workspace, w, bias, d.B * d.C, d.C, rs2, dscale);
      KMU_LAUNCH_CHECK("dw3x3_dscale");
    }
  }
  return KMU_OK;
}

int kmu_dwconv3x3_fwd(const kmu_dwconv3x3_desc* dd, const float* x, const float* w, const float* bias, float* y, kmu_stream stream) {
  return dw_fwd(dd, x, w, bias, nullptr, y, stream);
}

int kmu_dwconv3x3_bwd(const kmu_dwconv3x3_desc* dd, const float* x, const float* dy, const float* w, float* dx, float* dw, float* dbias,
                      void* workspace, size_t workspace_bytes, kmu_stream stream) {
  return dw_bwd(dd, x, dy, w, nullptr, nullptr, nullptr, dx, dw, dbias, nullptr, workspace, workspace_bytes, stream);
}

int kmu_dwconv3x3_bwd_add(const kmu_dwconv3x3_desc* dd, const float* x, const float* dy, const float* w, const float* dx_add, float* dx,
                          float* dw, float* dbias, void* workspace, size_t workspace_bytes, kmu_stream stream) {
  KMU_REQUIRE(dx_add != nullptr && dx != nullptr, KMU_ERR_BAD_ARG, "dwconv3x3_bwd_add: null dx_add / dx");
  return dw_bwd(dd, x, dy, w, nullptr, nullptr, dx_add, dx, dw, dbias, nullptr, workspace, workspace_bytes, stream);
}

int kmu_dwconv3x3_scaled_fwd(const kmu_dwconv3x3_desc* dd, const float* x, const float* w, const float* bias, const float* scale,
                             float* y, kmu_stream stream) {
  KMU_REQUIRE(scale != nullptr, KMU_ERR_BAD_ARG, "dwconv3x3_scaled_fwd: null scale");
  return dw_fwd(dd, x, w, bias, scale, y, stream);
}

int kmu_dwconv3x3_scaled_bwd(const kmu_dwconv3x3_desc* dd, const float* x, const float* dy, const float* w, const float* bias,
                             const float* scale, float* dx, float* dw, float* dbias, float* dscale, void* workspace,
                             size_t workspace_bytes, kmu_stream stream) {
  KMU_REQUIRE(scale != nullptr, KMU_ERR_BAD_ARG, "dwconv3x3_scaled_bwd: null scale");
  return dw_bwd(dd, x, dy, w, bias, scale, nullptr, dx, dw, dbias, dscale, workspace, workspace_bytes, stream);
}

}  // extern "C"
