// pwconv.cu -- pointwise (1x1) convolution on NCHW tensors, forward / input gradient / weight+bias gradient.
//
// The EfficientViMBlock FFN (vim_utils_init.py:122-130: 1x1 C->4C, 1x1 4C->C), DirectionAttention.qkv (KM_UNetV3_SH.py:221),
// the EnhancedViMBlock FFN (:118-122), DirectionViM.proj in 'channel' mode (:178) and StableHybridKANConv.residual (:59) are
// all y[b,o,p] = bias[o] + sum_c W[o,c] x[b,c,p] with 16..256 channels on 1k..16k pixels: a few hundred FLOP per byte moved
// at most, i.e. HBM-bound streaming.  cuDNN runs them as NHWC implicit GEMMs bracketed by two layout-transposing kernels
// per call; here the NCHW planes are read as they lie (pixel-contiguous, coalesced), the weights sit in shared memory
// and every thread keeps a 2-pixel x 32-output register tile.  The weight gradient is a pixel reduction: persistent
// CTAs keep their Cout x Cin partial in registers across all their pixel tiles and a second kernel folds the partials in
// a fixed order (deterministic, no atomics).  fp32 FMA throughout: matches the reference's fp32 modules to rounding.
#include "common.cuh"

namespace kmu {
namespace pw {

constexpr int OTMAX = 32; // outputs per thread pass (16 when the layer has <= 16 outputs)
constexpr int NTH = 128;  // threads per CTA (forward / dgrad); each thread owns 2 pixels

// y[b, j0+j, p] = bias[j0+j] + sum_i Wt[i][j] in[b, i, p]; Wt[i][j] = W[j*NI + i] (forward) or W[i*NJ_total + j] (dgrad: in = dy)
// grid (ceil(HW / (2*NTH)), ceil(NJ/OT), B)
template <int OT>
__global__ void __launch_bounds__(NTH) pw_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                 const float* __restrict__ bias, float* __restrict__ out, int NI, int NJ, int HW,
                                                 int dgrad) {
  extern __shared__ __align__(16) float w_s[];  // [NI][OT]
  const int j0 = blockIdx.y * OT;
  for (int i = threadIdx.x; i < NI * OT; i += NTH) {
    int ii = i / OT, j = i - ii * OT;
    float v = 0.f;
    if (j0 + j < NJ) v = dgrad ? w[(size_t)ii * NJ + j0 + j] : w[(size_t)(j0 + j) * NI + ii];
    w_s[i] = v;
  }
  __syncthreads();
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * (2 * NTH) + threadIdx.x, p1 = p0 + NTH;
  const bool ok0 = p0 < HW, ok1 = p1 < HW;
  const float* ib = in + (size_t)b * NI * HW;
  float a0[OT], a1[OT];
#pragma unroll
  for (int j = 0; j < OT; ++j) {
    float bv = (bias && j0 + j < NJ) ? __ldg(bias + j0 + j) : 0.f;
    a0[j] = bv;
    a1[j] = bv;
  }
#pragma unroll 4
  for (int i = 0; i < NI; ++i) {
    const float x0 = ok0 ? __ldg(ib + (size_t)i * HW + p0) : 0.f;
    const float x1 = ok1 ? __ldg(ib + (size_t)i * HW + p1) : 0.f;
    const float4* w4 = reinterpret_cast<const float4*>(w_s + i * OT);
#pragma unroll
    for (int q = 0; q < OT / 4; ++q) {
      const float4 ww = w4[q];
      a0[4 * q + 0] = fmaf(ww.x, x0, a0[4 * q + 0]); a1[4 * q + 0] = fmaf(ww.x, x1, a1[4 * q + 0]);
      a0[4 * q + 1] = fmaf(ww.y, x0, a0[4 * q + 1]); a1[4 * q + 1] = fmaf(ww.y, x1, a1[4 * q + 1]);
      a0[4 * q + 2] = fmaf(ww.z, x0, a0[4 * q + 2]); a1[4 * q + 2] = fmaf(ww.z, x1, a1[4 * q + 2]);
      a0[4 * q + 3] = fmaf(ww.w, x0, a0[4 * q + 3]); a1[4 * q + 3] = fmaf(ww.w, x1, a1[4 * q + 3]);
    }
  }
  float* ob = out + (size_t)b * NJ * HW;
#pragma unroll
  for (int j = 0; j < OT; ++j) {
    if (j0 + j < NJ) {
      if (ok0) ob[(size_t)(j0 + j) * HW + p0] = a0[j];
      if (ok1) ob[(size_t)(j0 + j) * HW + p1] = a1[j];
    }
  }
}

// ---- weight / bias gradient.  dW[o][c] = sum_{b,p} dy[b,o,p] x[b,c,p], db[o] = sum dy.
// Thread tile = NO outputs x 4 inputs (NO >= 4), kept in registers across all the CTA's pixel tiles of TP pixels.  The
// (Cin/4) x (Cout/NO) thread grid may be smaller than the CTA: the CTA then holds PG = 256/T copies of it, copy g taking the
// pixels p = g (mod PG) of every tile and writing its own partial.  partial[cta*PG + g][Cout*Cin + Cout].
constexpr int TP = 64;
template <int NO>
__global__ void __launch_bounds__(256) pw_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                       float* __restrict__ partial, int Cin, int Cout, int HW, int B, int ntiles,
                                                       int TOn, int PG) {
  extern __shared__ __align__(16) float smem[];
  const int OC = TOn * NO;                    // outputs covered by the thread grid (>= Cout, multiple of 4)
  const int XP = Cin + 4, YP = OC + 4;        // row pitches ([pixel][channel], padded, 16-byte aligned)
  float* x_s = smem;            // [TP][XP]
  float* y_s = x_s + TP * XP;   // [TP][YP]
  const int TC = Cin >> 2, T = TC * TOn, tid = threadIdx.x;
  const int pg = tid / T, r = tid - pg * T;
  const int tc = r % TC, to = r / TC;
  const int o0 = to * NO;
  const bool active = pg < PG;
  float acc[NO][4], bacc[NO];
#pragma unroll
  for (int k = 0; k < NO; ++k) {
    bacc[k] = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[k][e] = 0.f;
  }
  const int tiles_per_img = (HW + TP - 1) / TP;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * TP;
    __syncthreads();
    // transposing stage-in: a warp covers 8 consecutive pixels x 4 channels, so its 32 stores ([pixel][channel] rows whose
    // pitch is an odd number of 16-byte groups) hit 32 different banks while every global access is a full 32-byte sector
    for (int i = tid; i < Cin * TP; i += 256) {
      const int rest = i >> 5;
      const int p = (rest % (TP / 8)) * 8 + (i & 7), c = (rest / (TP / 8)) * 4 + ((i >> 3) & 3);
      x_s[p * XP + c] = (p0 + p < HW) ? __ldg(x + ((size_t)b * Cin + c) * HW + p0 + p) : 0.f;
    }
    for (int i = tid; i < OC * TP; i += 256) {
      const int rest = i >> 5;
      const int p = (rest % (TP / 8)) * 8 + (i & 7), o = (rest / (TP / 8)) * 4 + ((i >> 3) & 3);
      y_s[p * YP + o] = (o < Cout && p0 + p < HW) ? __ldg(dy + ((size_t)b * Cout + o) * HW + p0 + p) : 0.f;
    }
    __syncthreads();
    if (active) {
      for (int p = pg; p < TP; p += PG) {
        const float4 xv = *reinterpret_cast<const float4*>(x_s + p * XP + 4 * tc);
        float g[NO];
#pragma unroll
        for (int q = 0; q < NO / 4; ++q) {
          const float4 t = *reinterpret_cast<const float4*>(y_s + p * YP + o0 + 4 * q);
          g[4 * q] = t.x; g[4 * q + 1] = t.y; g[4 * q + 2] = t.z; g[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int k = 0; k < NO; ++k) {
          acc[k][0] = fmaf(g[k], xv.x, acc[k][0]);
          acc[k][1] = fmaf(g[k], xv.y, acc[k][1]);
          acc[k][2] = fmaf(g[k], xv.z, acc[k][2]);
          acc[k][3] = fmaf(g[k], xv.w, acc[k][3]);
          bacc[k] += g[k];
        }
      }
    }
  }
  if (active) {
    float* pb = partial + ((size_t)blockIdx.x * PG + pg) * ((size_t)Cout * Cin + Cout);
#pragma unroll
    for (int k = 0; k < NO; ++k) {
      const int o = o0 + k;
      if (o < Cout) {
#pragma unroll
        for (int e = 0; e < 4; ++e) pb[(size_t)o * Cin + 4 * tc + e] = acc[k][e];
        if (tc == 0) pb[(size_t)Cout * Cin + o] = bacc[k];
      }
    }
  }
}

// fixed-order sum of the partials: CTA = 32 outputs x 8 interleaved part slices
__global__ void __launch_bounds__(256) pw_wreduce_kernel(const float* __restrict__ partial, int nparts, int n_w, int n_b,
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
    else if (db) db[idx - n_w] = t;
  }
}

static int check(const kmu_pwconv_desc* d, const char* who) {
  KMU_REQUIRE(d != nullptr, KMU_ERR_BAD_ARG, "%s: null descriptor", who);
  KMU_REQUIRE(d->B > 0 && d->Cin > 0 && d->Cout > 0 && d->HW > 0, KMU_ERR_BAD_ARG, "%s: non-positive shape", who);
  KMU_REQUIRE(d->B <= 65535, KMU_ERR_UNSUPPORTED, "%s: B=%d > 65535", who, d->B);
  KMU_REQUIRE((size_t)(d->Cin > d->Cout ? d->Cin : d->Cout) * OTMAX * 4 <= 200 * 1024, KMU_ERR_UNSUPPORTED, "%s: too many channels", who);
  return KMU_OK;
}
static void launch_pw(const float* in, const float* w, const float* bias, float* out, int NI, int NJ, int HW, int B, int dgrad,
                      cudaStream_t st) {
  if (NJ <= 16) {
    size_t smem = (size_t)NI * 16 * 4;
    if (smem > 48 * 1024) cudaFuncSetAttribute(pw_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pw_kernel<16><<<dim3(cdiv(HW, 2 * NTH), 1, B), NTH, smem, st>>>(in, w, bias, out, NI, NJ, HW, dgrad);
  } else {
    size_t smem = (size_t)NI * 32 * 4;
    if (smem > 48 * 1024) cudaFuncSetAttribute(pw_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pw_kernel<32><<<dim3(cdiv(HW, 2 * NTH), cdiv(NJ, 32), B), NTH, smem, st>>>(in, w, bias, out, NI, NJ, HW, dgrad);
  }
}

struct WgradPlan {
  int NO, TOn, PG, ctas;
};
static bool wgrad_ok(const kmu_pwconv_desc& d) {
  const int c = d.Cin;
  const bool pow2 = c >= 16 && c <= 1024 && (c & (c - 1)) == 0;
  return pow2 && (long long)d.Cout * c <= 16 * 1024;
}
static WgradPlan wgrad_plan(const kmu_pwconv_desc& d) {
  WgradPlan p;
  p.NO = 4;
  while ((d.Cin / 4) * cdiv(d.Cout, p.NO) > 256) p.NO <<= 1;   // <= 16 by wgrad_ok
  p.TOn = cdiv(d.Cout, p.NO);
  p.PG = 256 / ((d.Cin / 4) * p.TOn);
  if (p.PG > 8) p.PG = 8;
  long long tiles = (long long)d.B * cdiv(d.HW, TP);
  p.ctas = (int)(tiles < 296 ? tiles : 296);
  return p;
}

template <int NO>
static void launch_wgrad(const kmu_pwconv_desc& d, const WgradPlan& pl, const float* x, const float* dy, float* partial, cudaStream_t st) {
  const int XP = d.Cin + 4, YP = pl.TOn * NO + 4;
  size_t smem = (size_t)TP * (XP + YP) * 4;
  if (smem > 48 * 1024) cudaFuncSetAttribute(pw_wgrad_kernel<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int ntiles = d.B * cdiv(d.HW, TP);
  pw_wgrad_kernel<NO><<<pl.ctas, 256, smem, st>>>(x, dy, partial, d.Cin, d.Cout, d.HW, d.B, ntiles, pl.TOn, pl.PG);
}

}  // namespace pw
}  // namespace kmu

using namespace kmu;
using namespace kmu::pw;

extern "C" {

int kmu_pwconv_wgrad_supported(const kmu_pwconv_desc* d) { return (d && wgrad_ok(*d)) ? 1 : 0; }

size_t kmu_pwconv_bwd_workspace_bytes(const kmu_pwconv_desc* d) {
  if (check(d, "pwconv_bwd_workspace_bytes") != KMU_OK) return 0;
  if (!wgrad_ok(*d)) return 256;
  WgradPlan pl = wgrad_plan(*d);
  return align_up((size_t)pl.ctas * pl.PG * ((size_t)d->Cout * d->Cin + d->Cout) * 4, 256);
}

int kmu_pwconv_fwd(const kmu_pwconv_desc* d, const float* x, const float* w, const float* bias, float* y, kmu_stream stream) {
  int rc = check(d, "pwconv_fwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(x && w && y, KMU_ERR_BAD_ARG, "pwconv_fwd: null tensor");
  launch_pw(x, w, bias, y, d->Cin, d->Cout, d->HW, d->B, 0, (cudaStream_t)stream);
  KMU_LAUNCH_CHECK("pw_fwd");
  return KMU_OK;
}

int kmu_pwconv_bwd(const kmu_pwconv_desc* d, const float* x, const float* dy, const float* w, float* dx, float* dw, float* dbias,
                   void* workspace, size_t workspace_bytes, kmu_stream stream) {
  int rc = check(d, "pwconv_bwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(dy && w, KMU_ERR_BAD_ARG, "pwconv_bwd: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) {
    launch_pw(dy, w, nullptr, dx, d->Cout, d->Cin, d->HW, d->B, 1, st);
    KMU_LAUNCH_CHECK("pw_dgrad");
  }
  if (dw) {
    KMU_REQUIRE(x != nullptr, KMU_ERR_BAD_ARG, "pwconv_bwd: weight gradient needs x");
    KMU_REQUIRE(wgrad_ok(*d), KMU_ERR_UNSUPPORTED, "pwconv_bwd: weight gradient needs Cin a power of two in [16,1024] and Cin*Cout <= 16384 "
                "(got %d -> %d)", d->Cin, d->Cout);
    KMU_REQUIRE(workspace && workspace_bytes >= kmu_pwconv_bwd_workspace_bytes(d), KMU_ERR_WORKSPACE, "pwconv_bwd: workspace too small");
    float* partial = (float*)workspace;
    const WgradPlan pl = wgrad_plan(*d);
    switch (pl.NO) {
      case 4: launch_wgrad<4>(*d, pl, x, dy, partial, st); break;
      case 8: launch_wgrad<8>(*d, pl, x, dy, partial, st); break;
      default: launch_wgrad<16>(*d, pl, x, dy, partial, st); break;
    }
    KMU_LAUNCH_CHECK("pw_wgrad");
    const int n_w = d->Cout * d->Cin, n_b = d->Cout;
    pw_wreduce_kernel<<<cdiv(n_w + n_b, 32), 256, 0, st>>>(partial, pl.ctas * pl.PG, n_w, n_b, dw, dbias);
    KMU_LAUNCH_CHECK("pw_wreduce");
  }
  return KMU_OK;
}

}  // extern "C"
