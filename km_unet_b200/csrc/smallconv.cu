// smallconv.cu -- dense "same" convolutions with at most 9 taps (1x3, 3x1, 3x3; stride 1, zero padding k/2) on NCHW tensors,
// forward / input gradient / weight+bias gradient.
//
// The callers of the hot path use them everywhere the tensors are still wide but the channel counts are tiny:
// DirectionViM.proj (KM_UNetV3_SH.py:172-176: 3x1 / 1x3, C -> C), the decoder's 3x3 convs (:427,437,439) and
// MultiScaleFusion's 3x3 (:292,299).  With 16..64 channels these are streaming problems (tens of FLOP per byte); cuDNN runs them
// as NHWC implicit GEMMs with a layout-transposing kernel on each side.  Same scheme as pwconv.cu: pixel-contiguous NCHW
// planes are read as they lie, a tap is just a shifted, bounds-checked read of the same plane ("virtual input channel"
// (c, tap)), weights sit in shared memory, every thread keeps a 2-pixel x OT-output register tile; the weight gradient is one
// persistent pixel reduction per tap with deterministic partials.
#include "common.cuh"

namespace kmu {
namespace sc {

constexpr int NTH = 128;
constexpr int MAXT = 9;

struct Taps {
  int T;
  int dy[MAXT], dx[MAXT];
};

static Taps make_taps(int kh, int kw, bool negate) {
  Taps t;
  t.T = kh * kw;
  for (int i = 0; i < kh; ++i)
    for (int j = 0; j < kw; ++j) {
      int k = i * kw + j;
      t.dy[k] = (i - kh / 2) * (negate ? -1 : 1);
      t.dx[k] = (j - kw / 2) * (negate ? -1 : 1);
    }
  for (int k = t.T; k < MAXT; ++k) t.dy[k] = t.dx[k] = 0;
  return t;
}

// out[b, j0+j, p] = bias + sum_{c,t} Wt[(c,t)][j] in[b, c, p + delta_t]
//   forward: Wt[(c,t)][j] = w[((j0+j)*NC + c)*T + t]                 (in = x, NC = Cin, NJ = Cout)
//   dgrad:   Wt[(o,t)][j] = w[((o)*NJ + (j0+j))*T + t], delta negated  (in = dy, NC = Cout, NJ = Cin)
template <int OT>
__global__ void __launch_bounds__(NTH) sc_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                 const float* __restrict__ bias, float* __restrict__ out, int NC, int NJ, int H, int W,
                                                 Taps taps, int dgrad) {
  extern __shared__ __align__(16) float w_s[];  // [NC*T][OT]
  const int T = taps.T, j0 = blockIdx.y * OT, HW = H * W;
  for (int i = threadIdx.x; i < NC * T * OT; i += NTH) {
    const int vi = i / OT, j = i - vi * OT;
    const int c = vi / T, t = vi - c * T;
    float v = 0.f;
    if (j0 + j < NJ) v = dgrad ? w[((size_t)c * NJ + j0 + j) * T + t] : w[((size_t)(j0 + j) * NC + c) * T + t];
    w_s[i] = v;
  }
  __syncthreads();
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * (2 * NTH) + threadIdx.x, p1 = p0 + NTH;
  const int h0 = p0 / W, c0 = p0 - h0 * W, h1 = p1 / W, c1 = p1 - h1 * W;
  int off[MAXT];
  unsigned ok0 = 0, ok1 = 0;
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    off[t] = taps.dy[t] * W + taps.dx[t];
    if (t < T) {
      if (p0 < HW && (unsigned)(h0 + taps.dy[t]) < (unsigned)H && (unsigned)(c0 + taps.dx[t]) < (unsigned)W) ok0 |= 1u << t;
      if (p1 < HW && (unsigned)(h1 + taps.dy[t]) < (unsigned)H && (unsigned)(c1 + taps.dx[t]) < (unsigned)W) ok1 |= 1u << t;
    }
  }
  const float* ib = in + (size_t)b * NC * HW;
  float a0[OT], a1[OT];
#pragma unroll
  for (int j = 0; j < OT; ++j) {
    const float bv = (bias && j0 + j < NJ) ? __ldg(bias + j0 + j) : 0.f;
    a0[j] = bv;
    a1[j] = bv;
  }
  for (int c = 0; c < NC; ++c) {
    const float* ic = ib + (size_t)c * HW;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      if (t < T) {
        const float x0 = ((ok0 >> t) & 1u) ? __ldg(ic + p0 + off[t]) : 0.f;
        const float x1 = ((ok1 >> t) & 1u) ? __ldg(ic + p1 + off[t]) : 0.f;
        const float4* w4 = reinterpret_cast<const float4*>(w_s + (c * T + t) * OT);
#pragma unroll
        for (int q = 0; q < OT / 4; ++q) {
          const float4 ww = w4[q];
          a0[4 * q + 0] = fmaf(ww.x, x0, a0[4 * q + 0]); a1[4 * q + 0] = fmaf(ww.x, x1, a1[4 * q + 0]);
          a0[4 * q + 1] = fmaf(ww.y, x0, a0[4 * q + 1]); a1[4 * q + 1] = fmaf(ww.y, x1, a1[4 * q + 1]);
          a0[4 * q + 2] = fmaf(ww.z, x0, a0[4 * q + 2]); a1[4 * q + 2] = fmaf(ww.z, x1, a1[4 * q + 2]);
          a0[4 * q + 3] = fmaf(ww.w, x0, a0[4 * q + 3]); a1[4 * q + 3] = fmaf(ww.w, x1, a1[4 * q + 3]);
        }
      }
    }
  }
  float* ob = out + (size_t)b * NJ * HW;
#pragma unroll
  for (int j = 0; j < OT; ++j) {
    if (j0 + j < NJ) {
      if (p0 < HW) ob[(size_t)(j0 + j) * HW + p0] = a0[j];
      if (p1 < HW) ob[(size_t)(j0 + j) * HW + p1] = a1[j];
    }
  }
}

// ---- weight / bias gradient of ONE tap: dW[o][c][t] = sum_{b,p} dy[b,o,p] x[b,c,p+delta_t]; db[o] = sum dy (tap 0 only).
// Same thread layout as pwconv.cu's pw_wgrad_kernel (NO outputs x 4 inputs per thread, PG pixel groups per CTA).
constexpr int TP = 64;
template <int NO>
__global__ void __launch_bounds__(256) sc_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                       float* __restrict__ partial, int Cin, int Cout, int H, int W, int ntiles,
                                                       int TOn, int PG, int sdy, int sdx) {
  extern __shared__ __align__(16) float smem[];
  const int HW = H * W;
  const int OC = TOn * NO;
  const int XP = Cin + 4, YP = OC + 4;
  float* x_s = smem;            // [TP][XP]
  float* y_s = x_s + TP * XP;   // [TP][YP]
  const int TC = Cin >> 2, Tn = TC * TOn, tid = threadIdx.x;
  const int pg = tid / Tn, r = tid - pg * Tn;
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
  const int shift = sdy * W + sdx;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * TP;
    __syncthreads();
    for (int i = tid; i < Cin * TP; i += 256) {
      const int rest = i >> 5;
      const int p = (rest % (TP / 8)) * 8 + (i & 7), c = (rest / (TP / 8)) * 4 + ((i >> 3) & 3);
      const int pp = p0 + p, hh = pp / W, ww = pp - hh * W;
      const bool ok = pp < HW && (unsigned)(hh + sdy) < (unsigned)H && (unsigned)(ww + sdx) < (unsigned)W;
      x_s[p * XP + c] = ok ? __ldg(x + ((size_t)b * Cin + c) * HW + pp + shift) : 0.f;
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

// fixed-order sum of one tap's partials into dw[(o*Cin + c)*T + t] (and db from tap 0)
__global__ void __launch_bounds__(256) sc_wreduce_kernel(const float* __restrict__ partial, int nparts, int n_w, int n_b, int T, int t,
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
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) a += red[q][o];
    if (idx < n_w) dw[(size_t)idx * T + t] = a;
    else if (db && t == 0) db[idx - n_w] = a;
  }
}

static int check(const kmu_smallconv_desc* d, const char* who) {
  KMU_REQUIRE(d != nullptr, KMU_ERR_BAD_ARG, "%s: null descriptor", who);
  KMU_REQUIRE(d->B > 0 && d->Cin > 0 && d->Cout > 0 && d->H > 0 && d->W > 0, KMU_ERR_BAD_ARG, "%s: non-positive shape", who);
  KMU_REQUIRE(d->kh >= 1 && d->kw >= 1 && (d->kh & 1) && (d->kw & 1) && d->kh * d->kw <= MAXT, KMU_ERR_UNSUPPORTED,
              "%s: kernel %dx%d (odd sizes with at most 9 taps)", who, d->kh, d->kw);
  KMU_REQUIRE(d->B <= 65535, KMU_ERR_UNSUPPORTED, "%s: B=%d > 65535", who, d->B);
  const size_t ch = (size_t)(d->Cin > d->Cout ? d->Cin : d->Cout);
  KMU_REQUIRE(ch * d->kh * d->kw * 32 * 4 <= 200 * 1024, KMU_ERR_UNSUPPORTED, "%s: %zu channels x %d taps exceed shared memory", who, ch,
              d->kh * d->kw);
  return KMU_OK;
}

static void launch(const float* in, const float* w, const float* bias, float* out, int NC, int NJ, int H, int W, int B, const Taps& taps,
                   int dgrad, cudaStream_t st) {
  const int HW = H * W;
  if (NJ <= 16) {
    size_t smem = (size_t)NC * taps.T * 16 * 4;
    if (smem > 48 * 1024) cudaFuncSetAttribute(sc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sc_kernel<16><<<dim3(cdiv(HW, 2 * NTH), 1, B), NTH, smem, st>>>(in, w, bias, out, NC, NJ, H, W, taps, dgrad);
  } else {
    size_t smem = (size_t)NC * taps.T * 32 * 4;
    if (smem > 48 * 1024) cudaFuncSetAttribute(sc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sc_kernel<32><<<dim3(cdiv(HW, 2 * NTH), cdiv(NJ, 32), B), NTH, smem, st>>>(in, w, bias, out, NC, NJ, H, W, taps, dgrad);
  }
}

struct WgradPlan {
  int NO, TOn, PG, ctas;
};
static bool wgrad_ok(const kmu_smallconv_desc& d) {
  const int c = d.Cin;
  const bool pow2 = c >= 16 && c <= 1024 && (c & (c - 1)) == 0;
  return pow2 && (long long)d.Cout * c <= 16 * 1024;
}
static WgradPlan wgrad_plan(const kmu_smallconv_desc& d) {
  WgradPlan p;
  p.NO = 4;
  while ((d.Cin / 4) * cdiv(d.Cout, p.NO) > 256) p.NO <<= 1;
  p.TOn = cdiv(d.Cout, p.NO);
  p.PG = 256 / ((d.Cin / 4) * p.TOn);
  if (p.PG > 8) p.PG = 8;
  long long tiles = (long long)d.B * cdiv(d.H * d.W, TP);
  p.ctas = (int)(tiles < 296 ? tiles : 296);
  return p;
}

template <int NO>
static void launch_wgrad(const kmu_smallconv_desc& d, const WgradPlan& pl, const float* x, const float* dy, float* partial, int sdy, int sdx,
                         cudaStream_t st) {
  const int XP = d.Cin + 4, YP = pl.TOn * NO + 4;
  size_t smem = (size_t)TP * (XP + YP) * 4;
  if (smem > 48 * 1024) cudaFuncSetAttribute(sc_wgrad_kernel<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int ntiles = d.B * cdiv(d.H * d.W, TP);
  sc_wgrad_kernel<NO><<<pl.ctas, 256, smem, st>>>(x, dy, partial, d.Cin, d.Cout, d.H, d.W, ntiles, pl.TOn, pl.PG, sdy, sdx);
}

}  // namespace sc
}  // namespace kmu

using namespace kmu;
using namespace kmu::sc;

extern "C" {

int kmu_smallconv_supported(const kmu_smallconv_desc* d) {
  if (!d || d->kh < 1 || d->kw < 1 || !(d->kh & 1) || !(d->kw & 1) || d->kh * d->kw > MAXT) return 0;
  const size_t ch = (size_t)(d->Cin > d->Cout ? d->Cin : d->Cout);
  if (ch * d->kh * d->kw * 32 * 4 > 200 * 1024) return 0;
  return wgrad_ok(*d) ? 1 : 0;
}

size_t kmu_smallconv_bwd_workspace_bytes(const kmu_smallconv_desc* d) {
  if (check(d, "smallconv_bwd_workspace_bytes") != KMU_OK) return 0;
  if (!wgrad_ok(*d)) return 256;
  WgradPlan pl = wgrad_plan(*d);
  return align_up((size_t)pl.ctas * pl.PG * ((size_t)d->Cout * d->Cin + d->Cout) * 4, 256);
}

int kmu_smallconv_fwd(const kmu_smallconv_desc* d, const float* x, const float* w, const float* bias, float* y, kmu_stream stream) {
  int rc = check(d, "smallconv_fwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(x && w && y, KMU_ERR_BAD_ARG, "smallconv_fwd: null tensor");
  launch(x, w, bias, y, d->Cin, d->Cout, d->H, d->W, d->B, make_taps(d->kh, d->kw, false), 0, (cudaStream_t)stream);
  KMU_LAUNCH_CHECK("sc_fwd");
  return KMU_OK;
}

int kmu_smallconv_bwd(const kmu_smallconv_desc* d, const float* x, const float* dy, const float* w, float* dx, float* dw, float* dbias,
                      void* workspace, size_t workspace_bytes, kmu_stream stream) {
  int rc = check(d, "smallconv_bwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(dy && w, KMU_ERR_BAD_ARG, "smallconv_bwd: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) {
    launch(dy, w, nullptr, dx, d->Cout, d->Cin, d->H, d->W, d->B, make_taps(d->kh, d->kw, true), 1, st);
    KMU_LAUNCH_CHECK("sc_dgrad");
  }
  if (dw) {
    KMU_REQUIRE(x != nullptr, KMU_ERR_BAD_ARG, "smallconv_bwd: weight gradient needs x");
    KMU_REQUIRE(wgrad_ok(*d), KMU_ERR_UNSUPPORTED, "smallconv_bwd: weight gradient needs Cin a power of two in [16,1024] and "
                "Cin*Cout <= 16384 (got %d -> %d)", d->Cin, d->Cout);
    KMU_REQUIRE(workspace && workspace_bytes >= kmu_smallconv_bwd_workspace_bytes(d), KMU_ERR_WORKSPACE, "smallconv_bwd: workspace too small");
    float* partial = (float*)workspace;
    const WgradPlan pl = wgrad_plan(*d);
    const Taps taps = make_taps(d->kh, d->kw, false);
    const int n_w = d->Cout * d->Cin, n_b = d->Cout;
    for (int t = 0; t < taps.T; ++t) {
      switch (pl.NO) {
        case 4: launch_wgrad<4>(*d, pl, x, dy, partial, taps.dy[t], taps.dx[t], st); break;
        case 8: launch_wgrad<8>(*d, pl, x, dy, partial, taps.dy[t], taps.dx[t], st); break;
        default: launch_wgrad<16>(*d, pl, x, dy, partial, taps.dy[t], taps.dx[t], st); break;
      }
      KMU_LAUNCH_CHECK("sc_wgrad");
      sc_wreduce_kernel<<<cdiv(n_w + n_b, 32), 256, 0, st>>>(partial, pl.ctas * pl.PG, n_w, n_b, taps.T, t, dw, dbias);
      KMU_LAUNCH_CHECK("sc_wreduce");
    }
  }
  return KMU_OK;
}

}  // extern "C"
