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
constexpr int TP = 64;      // pixels per tile of the weight-gradient kernels

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

// ---- 1x3 / 3x1 specialisation (the DirectionViM projections): thread = 4 consecutive pixels of a row x 16 outputs.
//      The centre tap is one 128-bit load per input channel; a horizontal conv builds its two shifted vectors from that load plus
//      one scalar on each side, a vertical conv takes two more 128-bit loads.  64 accumulators per thread, 192 FMA per channel for
//      3 (vertical) / 3 (horizontal: 1 vector + 2 scalars) loads and 12 broadcast LDS.128 of weights.
//      grid (ceil(HW / 512), ceil(NJ / 16), B), 128 threads.  Needs W % 4 == 0.
__global__ void __launch_bounds__(NTH) sc3_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                  float* __restrict__ out, int NC, int NJ, int H, int W, int vertical, int dgrad) {
  extern __shared__ __align__(16) float w_s[];  // [NC * 3][16]
  const int j0 = blockIdx.y * 16, HW = H * W;
  for (int i = threadIdx.x; i < NC * 3 * 16; i += NTH) {
    const int vi = i >> 4, j = i & 15;
    const int c = vi / 3, t = vi - c * 3;
    float v = 0.f;
    // dgrad runs the same correlation on dy with the taps mirrored: tap index 2 - t
    if (j0 + j < NJ) v = dgrad ? w[((size_t)c * NJ + j0 + j) * 3 + (2 - t)] : w[((size_t)(j0 + j) * NC + c) * 3 + t];
    w_s[i] = v;
  }
  __syncthreads();
  const int b = blockIdx.z;
  const int p = (blockIdx.x * NTH + threadIdx.x) * 4;
  if (p >= HW) return;
  const int h = p / W, col = p - h * W;
  const bool lo_ok = vertical ? h > 0 : col > 0;            // tap 0 (offset -1) in range for the first pixel
  const bool hi_ok = vertical ? h < H - 1 : col + 4 < W;    // tap 2 (offset +1) in range for the last pixel
  const float* ib = in + (size_t)b * NC * HW + p;
  float acc[16][4];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float bv = (bias && j0 + j < NJ) ? __ldg(bias + j0 + j) : 0.f;
    acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = bv;
  }
  for (int c = 0; c < NC; ++c) {
    const float* ic = ib + (size_t)c * HW;
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(ic));
    float4 v0, v2;
    if (vertical) {
      v0 = lo_ok ? __ldg(reinterpret_cast<const float4*>(ic - W)) : make_float4(0.f, 0.f, 0.f, 0.f);
      v2 = hi_ok ? __ldg(reinterpret_cast<const float4*>(ic + W)) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const float l = lo_ok ? __ldg(ic - 1) : 0.f, r = hi_ok ? __ldg(ic + 4) : 0.f;
      v0 = make_float4(l, v1.x, v1.y, v1.z);
      v2 = make_float4(v1.y, v1.z, v1.w, r);
    }
    const float4* w4 = reinterpret_cast<const float4*>(w_s + c * 48);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const float4 xv = t == 0 ? v0 : (t == 1 ? v1 : v2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 ww = w4[t * 4 + q];
        const float wj[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[4 * q + e][0] = fmaf(wj[e], xv.x, acc[4 * q + e][0]);
          acc[4 * q + e][1] = fmaf(wj[e], xv.y, acc[4 * q + e][1]);
          acc[4 * q + e][2] = fmaf(wj[e], xv.z, acc[4 * q + e][2]);
          acc[4 * q + e][3] = fmaf(wj[e], xv.w, acc[4 * q + e][3]);
        }
      }
    }
  }
  float* ob = out + (size_t)b * NJ * HW + p;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j0 + j < NJ) *reinterpret_cast<float4*>(ob + (size_t)(j0 + j) * HW) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
}

// ---- weight / bias gradient of ALL THREE taps of a 1x3 / 3x1 kernel in one pass over x and dy (the per-tap version below streams
//      both tensors once per tap).  Same thread layout as sc_wgrad_kernel with NO = 4: thread = 4 outputs x 4 inputs x 3 taps.
__global__ void __launch_bounds__(256) sc3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        float* __restrict__ partial, int Cin, int Cout, int H, int W, int ntiles, int TOn,
                                                        int PG, int vertical) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NO = 4;
  const int HW = H * W;
  const int OC = TOn * NO;
  const int XP = Cin + 4, YP = OC + 4;
  float* x_s = smem;                // [3][TP][XP]
  float* y_s = x_s + 3 * TP * XP;   // [TP][YP]
  const int TC = Cin >> 2, Tn = TC * TOn, tid = threadIdx.x;
  const int pg = tid / Tn, r = tid - pg * Tn;
  const int tc = r % TC, to = r / TC;
  const int o0 = to * NO;
  const bool active = pg < PG;
  float acc[3][NO][4], bacc[NO];
#pragma unroll
  for (int k = 0; k < NO; ++k) {
    bacc[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[t][k][e] = 0.f;
  }
  const int tiles_per_img = (HW + TP - 1) / TP;
  const int sdy = vertical ? 1 : 0, sdx = vertical ? 0 : 1;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * TP;
    __syncthreads();
    for (int i = tid; i < 3 * Cin * TP; i += 256) {
      const int t = i / (Cin * TP), ii = i - t * Cin * TP;
      const int rest = ii >> 5;
      const int p = (rest % (TP / 8)) * 8 + (ii & 7), c = (rest / (TP / 8)) * 4 + ((ii >> 3) & 3);
      const int pp = p0 + p, hh = pp / W, ww = pp - hh * W;
      const int dyv = (t - 1) * sdy, dxv = (t - 1) * sdx;
      const bool ok = pp < HW && (unsigned)(hh + dyv) < (unsigned)H && (unsigned)(ww + dxv) < (unsigned)W;
      x_s[(t * TP + p) * XP + c] = ok ? __ldg(x + ((size_t)b * Cin + c) * HW + pp + dyv * W + dxv) : 0.f;
    }
    for (int i = tid; i < OC * TP; i += 256) {
      const int rest = i >> 5;
      const int p = (rest % (TP / 8)) * 8 + (i & 7), o = (rest / (TP / 8)) * 4 + ((i >> 3) & 3);
      y_s[p * YP + o] = (o < Cout && p0 + p < HW) ? __ldg(dy + ((size_t)b * Cout + o) * HW + p0 + p) : 0.f;
    }
    __syncthreads();
    if (active) {
      for (int p = pg; p < TP; p += PG) {
        const float4 g4 = *reinterpret_cast<const float4*>(y_s + p * YP + o0);
        const float g[NO] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const float4 xv = *reinterpret_cast<const float4*>(x_s + (t * TP + p) * XP + 4 * tc);
#pragma unroll
          for (int k = 0; k < NO; ++k) {
            acc[t][k][0] = fmaf(g[k], xv.x, acc[t][k][0]);
            acc[t][k][1] = fmaf(g[k], xv.y, acc[t][k][1]);
            acc[t][k][2] = fmaf(g[k], xv.z, acc[t][k][2]);
            acc[t][k][3] = fmaf(g[k], xv.w, acc[t][k][3]);
          }
        }
#pragma unroll
        for (int k = 0; k < NO; ++k) bacc[k] += g[k];
      }
    }
  }
  if (active) {      // partial layout per (CTA, pixel group): dW as [o][c][t] (the weight tensor's own layout) | db
    float* pb = partial + ((size_t)blockIdx.x * PG + pg) * ((size_t)Cout * Cin * 3 + Cout);
#pragma unroll
    for (int k = 0; k < NO; ++k) {
      const int o = o0 + k;
      if (o < Cout) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int t = 0; t < 3; ++t) pb[((size_t)o * Cin + 4 * tc + e) * 3 + t] = acc[t][k][e];
        if (tc == 0) pb[(size_t)Cout * Cin * 3 + o] = bacc[k];
      }
    }
  }
}

// ---- weight / bias gradient of ONE tap: dW[o][c][t] = sum_{b,p} dy[b,o,p] x[b,c,p+delta_t]; db[o] = sum dy (tap 0 only).
// Same thread layout as pwconv.cu's pw_wgrad_kernel (NO outputs x 4 inputs per thread, PG pixel groups per CTA).
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

// ---- 3-tap weight gradient, warp-per-strip form.  A warp stages a strip of 128 consecutive pixels of one image row -- 16 input
//      channels (for a vertical kernel: the rows above and below as well) and 16 output-gradient channels -- in its own slice of
//      shared memory with coalesced 128-bit loads, then lane = (output channel o = lane & 15, half of the strip): per pixel quad one
//      LDS.128 of dy[o] and, per input channel, one broadcast LDS.128 of x (+ two scalars for the horizontal taps) feed 48 FMA x 4
//      pixels into 48 accumulators (o fixed, 16 c x 3 taps) that live for the warp's whole pixel range.  Every byte of x and dy is
//      read from HBM once; the arithmetic (0.8 GFLOP at (32,16,16,128,128)) is no issue.
//      grid (slices, (Cin/16) * (Cout/16)), 128 threads; partial[(slice * 4 + warp)][o][c][t] (+ bias), reduced in fixed order.
// strip width: 128 pixels for a horizontal kernel (one staged x row), 64 for a vertical one (three rows): ~17 KB of shared memory per
// warp either way, three 4-warp CTAs per SM
static inline int strip_width(int W, int vertical) { const int s = vertical ? 64 : 128; return W < s ? W : s; }
static inline size_t strip_floats(int sw, int vertical) { return (size_t)(vertical ? 3 : 1) * 16 * (sw + 8) + (size_t)16 * (sw + 4); }
__global__ void __launch_bounds__(128) sc3_wgrad_strip_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              float* __restrict__ partial, int Cin, int Cout, int H, int W, int B,
                                                              int vertical) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sw = W < (vertical ? 64 : 128) ? W : (vertical ? 64 : 128);      // strip width
  const int rows = vertical ? 3 : 1, xst = sw + 8, yst = sw + 4;
  float* xs = smem + (size_t)warp * ((size_t)rows * 16 * xst + (size_t)16 * yst);
  float* ys = xs + (size_t)rows * 16 * xst;
  const int cblocks = Cin >> 4;
  const int ob = blockIdx.y / cblocks, cb = blockIdx.y - ob * cblocks;
  const int c0 = cb * 16, o0 = ob * 16;
  const int HW = H * W;
  const int strips_per_row = W / sw;             // W % sw == 0 (host check)
  const long long nstrips = (long long)B * H * strips_per_row;
  const int o = lane & 15, half = lane >> 4;
  float acc[16][3], bacc = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c][0] = acc[c][1] = acc[c][2] = 0.f;
  for (long long st = (long long)blockIdx.x * 4 + warp; st < nstrips; st += (long long)gridDim.x * 4) {
    const int b = (int)(st / ((long long)H * strips_per_row));
    const int rem = (int)(st - (long long)b * H * strips_per_row);
    const int h = rem / strips_per_row, w0 = (rem - h * strips_per_row) * sw;
    __syncwarp();
    // stage: x rows (h - 1, h, h + 1 for a vertical kernel; h only otherwise) and dy row h, 16 channels each, 128-bit loads
    const int q4 = sw >> 2;
    for (int i = lane; i < rows * 16 * q4; i += 32) {
      const int q = i % q4, rc = i / q4, c = rc & 15, r = rc >> 4;
      const int hh = vertical ? h + r - 1 : h;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hh >= 0 && hh < H) v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)b * Cin + c0 + c) * HW + (size_t)hh * W + w0) + q);
      *reinterpret_cast<float4*>(xs + (r * 16 + c) * xst + 4 + 4 * q) = v;
    }
    if (!vertical && lane < 16) {               // the two halo pixels of a horizontal kernel
      const float* row = x + ((size_t)b * Cin + c0 + lane) * HW + (size_t)h * W;
      xs[lane * xst + 3] = w0 > 0 ? __ldg(row + w0 - 1) : 0.f;
      xs[lane * xst + 4 + sw] = w0 + sw < W ? __ldg(row + w0 + sw) : 0.f;
    }
    for (int i = lane; i < 16 * q4; i += 32) {
      const int q = i % q4, oo = i / q4;
      *reinterpret_cast<float4*>(ys + oo * yst + 4 * q) =
          __ldg(reinterpret_cast<const float4*>(dy + ((size_t)b * Cout + o0 + oo) * HW + (size_t)h * W + w0) + q);
    }
    __syncwarp();
    const int qh = q4 >> 1;                      // quads per half (sw % 8 == 0)
    for (int q = half * qh; q < (half + 1) * qh; ++q) {
      const float4 g = *reinterpret_cast<const float4*>(ys + o * yst + 4 * q);
      bacc += (g.x + g.y) + (g.z + g.w);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float4 v0, v1, v2;
        if (vertical) {
          v0 = *reinterpret_cast<const float4*>(xs + (0 * 16 + c) * xst + 4 + 4 * q);
          v1 = *reinterpret_cast<const float4*>(xs + (1 * 16 + c) * xst + 4 + 4 * q);
          v2 = *reinterpret_cast<const float4*>(xs + (2 * 16 + c) * xst + 4 + 4 * q);
        } else {
          const float* xr = xs + c * xst + 4 + 4 * q;
          v1 = *reinterpret_cast<const float4*>(xr);
          v0 = make_float4(xr[-1], v1.x, v1.y, v1.z);
          v2 = make_float4(v1.y, v1.z, v1.w, xr[4]);
        }
        acc[c][0] += g.x * v0.x + g.y * v0.y + g.z * v0.z + g.w * v0.w;
        acc[c][1] += g.x * v1.x + g.y * v1.y + g.z * v1.z + g.w * v1.w;
        acc[c][2] += g.x * v2.x + g.y * v2.y + g.z * v2.z + g.w * v2.w;
      }
    }
  }
  // the two halves of the strip, then out: partial[(slice, warp)][o][c][t] | bias
  float* pb = partial + ((size_t)blockIdx.x * 4 + warp) * ((size_t)Cout * Cin * 3 + Cout);
#pragma unroll
  for (int c = 0; c < 16; ++c)
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const float v = acc[c][t] + __shfl_xor_sync(0xffffffffu, acc[c][t], 16);
      if (half == 0) pb[((size_t)(o0 + o) * Cin + c0 + c) * 3 + t] = v;
    }
  const float bv = bacc + __shfl_xor_sync(0xffffffffu, bacc, 16);
  if (half == 0 && cb == 0) pb[(size_t)Cout * Cin * 3 + o0 + o] = bv;
}
static bool three_tap_strip(const kmu_smallconv_desc& d) {
  const int sw = strip_width(d.W, d.kh == 3);
  return d.kh * d.kw == 3 && d.Cin % 16 == 0 && d.Cout % 16 == 0 && sw % 8 == 0 && d.W % sw == 0;
}
static int strip_slices(const kmu_smallconv_desc& d) {
  const int blocks = (d.Cin / 16) * (d.Cout / 16);
  int s = (148 * 3 + blocks - 1) / blocks;
  const long long strips = (long long)d.B * d.H * (d.W / strip_width(d.W, d.kh == 3));
  if ((long long)s * 4 > strips) s = (int)((strips + 3) / 4);
  return s < 1 ? 1 : s;
}

// 1x3 / 3x1 fast path: one forward / dgrad kernel with 128-bit traffic, one weight-gradient pass for all three taps
static bool three_tap(const kmu_smallconv_desc& d) { return d.kh * d.kw == 3 && d.W % 4 == 0 && d.Cin % 4 == 0 && d.Cout % 4 == 0; }
static bool three_tap_wgrad(const kmu_smallconv_desc& d) {
  return three_tap(d) && wgrad_ok(d) && (d.Cin / 4) * cdiv(d.Cout, 4) <= 256 && (size_t)TP * (3 * (d.Cin + 4) + d.Cout + 8) * 4 <= 200 * 1024;
}
static void launch3(const float* in, const float* w, const float* bias, float* out, int NC, int NJ, int H, int W, int B, int vertical,
                    int dgrad, cudaStream_t st) {
  const size_t smem = (size_t)NC * 3 * 16 * 4;
  if (smem > 48 * 1024) cudaFuncSetAttribute(sc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sc3_kernel<<<dim3(cdiv(H * W, 4 * NTH), cdiv(NJ, 16), B), NTH, smem, st>>>(in, w, bias, out, NC, NJ, H, W, vertical, dgrad);
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
  size_t parts = (size_t)pl.ctas * pl.PG;
  if (three_tap_strip(*d) && (size_t)strip_slices(*d) * 4 > parts) parts = (size_t)strip_slices(*d) * 4;
  return align_up(parts * ((size_t)d->Cout * d->Cin * 3 + d->Cout) * 4, 256);
}

int kmu_smallconv_fwd(const kmu_smallconv_desc* d, const float* x, const float* w, const float* bias, float* y, kmu_stream stream) {
  int rc = check(d, "smallconv_fwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(x && w && y, KMU_ERR_BAD_ARG, "smallconv_fwd: null tensor");
  if (three_tap(*d) && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0) {
    launch3(x, w, bias, y, d->Cin, d->Cout, d->H, d->W, d->B, d->kh == 3, 0, (cudaStream_t)stream);
    KMU_LAUNCH_CHECK("sc3_fwd");
    return KMU_OK;
  }
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
    if (three_tap(*d) && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0)
      launch3(dy, w, nullptr, dx, d->Cout, d->Cin, d->H, d->W, d->B, d->kh == 3, 1, st);
    else
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
    if (three_tap_strip(*d) && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0) {
      const int slices = strip_slices(*d);
      const size_t smem = (size_t)4 * strip_floats(strip_width(d->W, d->kh == 3), d->kh == 3) * 4;
      cudaFuncSetAttribute(sc3_wgrad_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      sc3_wgrad_strip_kernel<<<dim3(slices, (d->Cin / 16) * (d->Cout / 16)), 128, smem, st>>>(x, dy, partial, d->Cin, d->Cout, d->H, d->W,
                                                                                            d->B, d->kh == 3);
      KMU_LAUNCH_CHECK("sc3_wgrad_strip");
      sc_wreduce_kernel<<<cdiv(3 * n_w + n_b, 32), 256, 0, st>>>(partial, slices * 4, 3 * n_w, n_b, 1, 0, dw, dbias);
      KMU_LAUNCH_CHECK("sc_wreduce");
      return KMU_OK;
    }
    if (three_tap_wgrad(*d)) {
      WgradPlan p3 = pl;
      p3.NO = 4;
      p3.TOn = cdiv(d->Cout, 4);
      p3.PG = 256 / ((d->Cin / 4) * p3.TOn);
      if (p3.PG > 8) p3.PG = 8;
      if (p3.PG > pl.PG) p3.PG = pl.PG;                         // the workspace was sized for pl.PG pixel groups
      const size_t smem = (size_t)TP * (3 * (d->Cin + 4) + p3.TOn * 4 + 4) * 4;
      if (smem > 48 * 1024) cudaFuncSetAttribute(sc3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      const int ntiles = d->B * cdiv(d->H * d->W, TP);
      sc3_wgrad_kernel<<<pl.ctas, 256, smem, st>>>(x, dy, partial, d->Cin, d->Cout, d->H, d->W, ntiles, p3.TOn, p3.PG, d->kh == 3);
      KMU_LAUNCH_CHECK("sc3_wgrad");
      // the partial already has the weight tensor's [o][c][t] layout: reduce it as ONE "tap" of 3 * n_w values
      sc_wreduce_kernel<<<cdiv(3 * n_w + n_b, 32), 256, 0, st>>>(partial, pl.ctas * p3.PG, 3 * n_w, n_b, 1, 0, dw, dbias);
      KMU_LAUNCH_CHECK("sc_wreduce");
      return KMU_OK;
    }
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
