// deform.cu -- DAGEM's deformable 3x3 convolution (DAGEM_md.py:46,98-101: offset = offset_conv(x); DeformConv2d(C, C, 3, padding=1)
// from torchvision 0.14, one offset group, stride 1, dilation 1), forward and backward.
//
//   col[b, c, t, p] = bilinear(x[b, c], y(p) - 1 + ki + off[b, 2t, p], x(p) - 1 + kj + off[b, 2t + 1, p])     t = ki * 3 + kj
//   out[b, o, p]    = bias[o] + sum_{c, t} W[o, c, t] col[b, c, t, p]
//   bilinear: zero when the point lies outside (-1, H) x (-1, W); corners outside the image contribute zero.
//
// Why it exists: torchvision's CUDA op launches its im2col / col2im kernels on the legacy default stream, not on the current one.
// Under CUDA-graph capture those launches are simply not captured, and the replayed training step multiplies stale columns
// (measured: 99 % error of the deformable branch in every replay; tools/graph_bisect.py).  It also scatters dX with fp32 atomics.
// Here every kernel runs on the caller's stream and every reduction has a fixed order (bit-reproducible):
//   forward   CTA = (16 output pixels, b): columns built in shared memory, then a 64 x 576 x 16 product on CUDA cores.
//   backward  dcol = W^T dout per 16-pixel tile (+ the offset gradient from the same tile), written once as [b][t][p][c];
//             dX: thread = (channel, slice of the (p, t) sample list) owns a private shared-memory image, slices summed in order;
//             dW: CTA = (b, 64 columns) recomputes its columns, accumulates 64 x 64 in registers, partials over b reduced in order.
// The tensor is tiny (B x 64 x 16 x 16 at the bridge): these kernels are launch / latency bound, not bandwidth bound.
#include "common.cuh"

namespace kmu {
namespace deform {

constexpr int PT = 16;      // output pixels per CTA (forward, dcol)
constexpr int KC = 32;      // weight columns staged per step

struct Geo {
  int i00;            // index of the top-left corner (clamped into the image)
  float w00, w01, w10, w11;   // corner weights, already zero for invalid corners / outside points
  float gy0, gy1, gx0, gx1;   // d/dpy = gy0 * (v10 - v00) + gy1 * (v11 - v01) ; d/dpx = gx0 * (v01 - v00) + gx1 * (v11 - v10)   (validity folded in below)
  int o01, o10, o11;  // offsets of the other corners from i00 (0 when the corner is invalid: its weight is zero)
  float m00, m01, m10, m11;   // 1 for valid corners (for the offset gradient)
};

__device__ __forceinline__ Geo geometry(const float* __restrict__ off_b, int t, int p, int H, int W) {
  const int HW = H * W;
  const int oy = p / W, ox = p - oy * W;
  const int ki = t / 3, kj = t - ki * 3;
  const float py = (float)(oy - 1 + ki) + off_b[(size_t)(2 * t) * HW + p];
  const float px = (float)(ox - 1 + kj) + off_b[(size_t)(2 * t + 1) * HW + p];
  Geo g;
  const bool inside = py > -1.f && py < (float)H && px > -1.f && px < (float)W;
  const float fy = floorf(py), fx = floorf(px);
  const float ly = py - fy, lx = px - fx;
  const int y0 = (int)fy, x0 = (int)fx;
  const bool y0ok = inside && y0 >= 0 && y0 <= H - 1, y1ok = inside && y0 + 1 >= 0 && y0 + 1 <= H - 1;
  const bool x0ok = inside && x0 >= 0 && x0 <= W - 1, x1ok = inside && x0 + 1 >= 0 && x0 + 1 <= W - 1;
  g.m00 = (y0ok && x0ok) ? 1.f : 0.f;
  g.m01 = (y0ok && x1ok) ? 1.f : 0.f;
  g.m10 = (y1ok && x0ok) ? 1.f : 0.f;
  g.m11 = (y1ok && x1ok) ? 1.f : 0.f;
  g.w00 = g.m00 * (1.f - ly) * (1.f - lx);
  g.w01 = g.m01 * (1.f - ly) * lx;
  g.w10 = g.m10 * ly * (1.f - lx);
  g.w11 = g.m11 * ly * lx;
  g.gy0 = 1.f - lx; g.gy1 = lx; g.gx0 = 1.f - ly; g.gx1 = ly;
  const int yc = min(max(y0, 0), H - 1), xc = min(max(x0, 0), W - 1);
  g.i00 = (y0ok && x0ok) ? y0 * W + x0 : yc * W + xc;
  // the other corners relative to a base that is always in range: use absolute indices folded into offsets from i00
  const int i01 = (y0ok && x1ok) ? y0 * W + x0 + 1 : g.i00;
  const int i10 = (y1ok && x0ok) ? (y0 + 1) * W + x0 : g.i00;
  const int i11 = (y1ok && x1ok) ? (y0 + 1) * W + x0 + 1 : g.i00;
  g.o01 = i01 - g.i00; g.o10 = i10 - g.i00; g.o11 = i11 - g.i00;
  return g;
}

// ---------------------------------------------------------------------------------------------------------------- forward
// grid (ceil(HW / PT), B), 256 threads.  smem: col[C*9][PT] + wt[KC][Cout+1] + geo[9*PT]
__global__ void __launch_bounds__(256) deform_fwd_kernel(const float* __restrict__ x, const float* __restrict__ off,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ out, int C, int H, int W, int Cout) {
  extern __shared__ __align__(16) float sm[];
  const int HW = H * W, K = C * 9;
  float* col = sm;                         // [K][PT]
  float* wt = col + (size_t)K * PT;        // [KC][Cout + 1]
  Geo* geo = reinterpret_cast<Geo*>(wt + (size_t)KC * (Cout + 1));
  const int b = blockIdx.y, p0 = blockIdx.x * PT, tid = threadIdx.x;
  const float* off_b = off + (size_t)b * 18 * HW;
  const float* xb = x + (size_t)b * C * HW;
  for (int i = tid; i < 9 * PT; i += 256) {
    const int t = i / PT, pl = i - t * PT;
    const int p = min(p0 + pl, HW - 1);
    geo[i] = geometry(off_b, t, p, H, W);
  }
  __syncthreads();
  for (int i = tid; i < K * PT; i += 256) {
    const int k = i / PT, pl = i - k * PT;
    const int c = k / 9, t = k - c * 9;
    const Geo g = geo[t * PT + pl];
    const float* xc = xb + (size_t)c * HW + g.i00;
    col[i] = g.w00 * __ldg(xc) + g.w01 * __ldg(xc + g.o01) + g.w10 * __ldg(xc + g.o10) + g.w11 * __ldg(xc + g.o11);
  }
  // out[o][4 pixels] per thread: o = tid % 64 (+64 ...), pixel quad = tid / 64
  const int pq = (tid >> 6) * 4;
  const int NO = (Cout + 63) / 64;         // output channels per thread (Cout <= 256)
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  for (int k0 = 0; k0 < K; k0 += KC) {
    __syncthreads();
    for (int i = tid; i < KC * Cout; i += 256) {      // W[o][k0 + kk], coalesced along kk
      const int o = i / KC, kk = i - o * KC;
      wt[kk * (Cout + 1) + o] = (k0 + kk < K) ? __ldg(w + (size_t)o * K + k0 + kk) : 0.f;
    }
    __syncthreads();
    const int kend = min(KC, K - k0);
    for (int kk = 0; kk < kend; ++kk) {
      const float4 cv = *reinterpret_cast<const float4*>(col + (size_t)(k0 + kk) * PT + pq);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < NO) {
          const int o = (tid & 63) + 64 * j;
          const float wv = o < Cout ? wt[kk * (Cout + 1) + o] : 0.f;
          acc[j][0] = fmaf(wv, cv.x, acc[j][0]);
          acc[j][1] = fmaf(wv, cv.y, acc[j][1]);
          acc[j][2] = fmaf(wv, cv.z, acc[j][2]);
          acc[j][3] = fmaf(wv, cv.w, acc[j][3]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int o = (tid & 63) + 64 * j;
    if (j < NO && o < Cout) {
      const float bv = bias ? bias[o] : 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int p = p0 + pq + e;
        if (p < HW) out[((size_t)b * Cout + o) * HW + p] = acc[j][e] + bv;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- backward
// dcol[b][t][p][c] = sum_o W[o][c*9+t] dout[b][o][p]; doff[b][2t / 2t+1][p] = sum_c dcol * d(sample)/d(py / px).
// grid (ceil(HW / PT), B), 256 threads.  smem: dout_s[Cout][PT] + dcol_s[9][PT][C + 1] + geo[9 * PT]
__global__ void __launch_bounds__(256) deform_dcol_kernel(const float* __restrict__ x, const float* __restrict__ off,
                                                          const float* __restrict__ w, const float* __restrict__ dout,
                                                          float* __restrict__ dcol, float* __restrict__ doff, int C, int H, int W,
                                                          int Cout) {
  extern __shared__ __align__(16) float sm[];
  const int HW = H * W, K = C * 9;
  float* dout_s = sm;                                   // [Cout][PT]
  float* dcol_s = dout_s + (size_t)Cout * PT;           // [9][PT][C + 1]
  Geo* geo = reinterpret_cast<Geo*>(dcol_s + (size_t)9 * PT * (C + 1));
  const int b = blockIdx.y, p0 = blockIdx.x * PT, tid = threadIdx.x;
  const float* off_b = off + (size_t)b * 18 * HW;
  for (int i = tid; i < Cout * PT; i += 256) {
    const int o = i / PT, pl = i - o * PT;
    dout_s[i] = (p0 + pl < HW) ? __ldg(dout + ((size_t)b * Cout + o) * HW + p0 + pl) : 0.f;
  }
  for (int i = tid; i < 9 * PT; i += 256) {
    const int t = i / PT, pl = i - t * PT;
    geo[i] = geometry(off_b, t, min(p0 + pl, HW - 1), H, W);
  }
  __syncthreads();
  for (int kp = tid; kp < K; kp += 256) {                // kp = t * C + c: consecutive threads -> consecutive channels
    const int t = kp / C, c = kp - t * C;
    const float* wk = w + (size_t)c * 9 + t;
    float acc[PT];
#pragma unroll
    for (int e = 0; e < PT; ++e) acc[e] = 0.f;
    for (int o = 0; o < Cout; ++o) {
      const float wv = __ldg(wk + (size_t)o * K);
      const float4* d4 = reinterpret_cast<const float4*>(dout_s + (size_t)o * PT);
#pragma unroll
      for (int q = 0; q < PT / 4; ++q) {
        const float4 d = d4[q];
        acc[4 * q + 0] = fmaf(wv, d.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(wv, d.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(wv, d.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(wv, d.w, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int e = 0; e < PT; ++e) {
      dcol_s[((size_t)t * PT + e) * (C + 1) + c] = acc[e];
      if (p0 + e < HW) dcol[(((size_t)b * 9 + t) * HW + p0 + e) * C + c] = acc[e];
    }
  }
  __syncthreads();
  // offset gradient: thread = (t, pixel), fixed-order sum over channels
  if (tid < 9 * PT) {
    const int t = tid / PT, pl = tid - t * PT;
    const int p = p0 + pl;
    if (p < HW) {
      const Geo g = geo[tid];
      const float* xb = x + (size_t)b * C * HW + g.i00;
      const float* dc = dcol_s + ((size_t)t * PT + pl) * (C + 1);
      float sy = 0.f, sx = 0.f;
      for (int c = 0; c < C; ++c) {
        const float* xc = xb + (size_t)c * HW;
        const float v00 = g.m00 * __ldg(xc), v01 = g.m01 * __ldg(xc + g.o01), v10 = g.m10 * __ldg(xc + g.o10),
                    v11 = g.m11 * __ldg(xc + g.o11);
        const float d = dc[c];
        sy = fmaf(d, g.gy0 * (v10 - v00) + g.gy1 * (v11 - v01), sy);
        sx = fmaf(d, g.gx0 * (v01 - v00) + g.gx1 * (v11 - v10), sx);
      }
      doff[((size_t)b * 18 + 2 * t) * HW + p] = sy;
      doff[((size_t)b * 18 + 2 * t + 1) * HW + p] = sx;
    }
  }
}

// dX: grid (C / CPG, B), CPG * S threads: thread (cl, s) owns img[s][cl][HW] in shared memory and walks samples s, s + S, ...
// of the (p, t) list; the S images of a channel are then summed in slice order.
template <int S>
__global__ void deform_dx_kernel(const float* __restrict__ off, const float* __restrict__ dcol, float* __restrict__ dx, int C, int H,
                                 int W, int CPG) {
  extern __shared__ __align__(16) float sm[];
  const int HW = H * W;
  const int HWP = HW | 1;      // odd image pitch: the CPG threads of a slice hit the same pixel of different channels -> different banks
  const int b = blockIdx.y, c0 = blockIdx.x * CPG;
  const int nthr = CPG * S;
  const int cl = threadIdx.x % CPG, s = threadIdx.x / CPG;
  for (int i = threadIdx.x; i < S * CPG * HWP; i += nthr) sm[i] = 0.f;
  __syncthreads();
  float* img = sm + ((size_t)s * CPG + cl) * HWP;
  const float* off_b = off + (size_t)b * 18 * HW;
  const float* dc_b = dcol + (size_t)b * 9 * HW * C + c0 + cl;
  if (c0 + cl < C) {
    for (int i = s; i < 9 * HW; i += S) {
      const int p = i / 9, t = i - p * 9;
      const Geo g = geometry(off_b, t, p, H, W);
      const float d = __ldg(dc_b + ((size_t)t * HW + p) * C);
      float* q = img + g.i00;
      q[0] += d * g.w00;           // invalid corners alias i00 with weight zero
      q[g.o01] += d * g.w01;
      q[g.o10] += d * g.w10;
      q[g.o11] += d * g.w11;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CPG * HW; i += nthr) {
    const int c = i / HW, p = i - c * HW;
    if (c0 + c < C) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < S; ++k) a += sm[((size_t)k * CPG + c) * HWP + p];
      dx[((size_t)b * C + c0 + c) * HW + p] = a;
    }
  }
}

// dW partial: grid (K / 64, B), 256 threads: CTA recomputes col rows [k0, k0 + 64) of image b tile by tile and accumulates
// part[b][o][k] (64 x 64 per CTA; thread = 4 o x 4 k).  Cout <= 64 per pass (looped for larger Cout).
constexpr int WP = 32;   // pixels per tile of the weight-gradient pass
__global__ void __launch_bounds__(256) deform_dw_kernel(const float* __restrict__ x, const float* __restrict__ off,
                                                        const float* __restrict__ dout, float* __restrict__ part, int C, int H, int W,
                                                        int Cout) {
  __shared__ __align__(16) float col_s[64][WP + 4];
  __shared__ __align__(16) float do_s[64][WP + 4];
  const int HW = H * W, K = C * 9;
  const int b = blockIdx.y, k0 = blockIdx.x * 64, tid = threadIdx.x;
  const float* off_b = off + (size_t)b * 18 * HW;
  const float* xb = x + (size_t)b * C * HW;
  const int to = (tid >> 4) * 4, tk = (tid & 15) * 4;
  for (int o0 = 0; o0 < Cout; o0 += 64) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int p0 = 0; p0 < HW; p0 += WP) {
      __syncthreads();
      for (int i = tid; i < 64 * WP; i += 256) {
        const int r = i / WP, pl = i - r * WP;
        const int p = p0 + pl, k = k0 + r, o = o0 + r;
        float cv = 0.f;
        if (p < HW && k < K) {
          const int c = k / 9, t = k - c * 9;
          const Geo g = geometry(off_b, t, p, H, W);
          const float* xc = xb + (size_t)c * HW + g.i00;
          cv = g.w00 * __ldg(xc) + g.w01 * __ldg(xc + g.o01) + g.w10 * __ldg(xc + g.o10) + g.w11 * __ldg(xc + g.o11);
        }
        col_s[r][pl] = cv;
        do_s[r][pl] = (p < HW && o < Cout) ? __ldg(dout + ((size_t)b * Cout + o) * HW + p) : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int pl = 0; pl < WP; ++pl) {
        float dv[4], cv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { dv[i] = do_s[to + i][pl]; cv[i] = col_s[tk + i][pl]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dv[i], cv[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int o = o0 + to + i, k = k0 + tk + j;
        if (o < Cout && k < K) part[((size_t)b * Cout + o) * K + k] = acc[i][j];
      }
  }
}

// dW[o][k] = sum_b part[b][o][k] (fixed order); dbias[o] = sum_{b,p} dout[b][o][p] (one warp per o, fixed order).
__global__ void __launch_bounds__(256) deform_dw_reduce_kernel(const float* __restrict__ part, const float* __restrict__ dout,
                                                               float* __restrict__ dw, float* __restrict__ dbias, int B, int Cout, int K,
                                                               int HW) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < Cout * K) {
    float a = 0.f;
    for (int b = 0; b < B; ++b) a += part[(size_t)b * Cout * K + i];
    dw[i] = a;
  }
  if (dbias && blockIdx.x < (unsigned)((Cout + 7) / 8)) {
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o < Cout) {
      float a = 0.f;
      for (int b = 0; b < B; ++b) {
        const float* d = dout + ((size_t)b * Cout + o) * HW;
        float s = 0.f;
        for (int p = lane; p < HW; p += 32) s += d[p];
        a += warp_sum(s);
      }
      if (lane == 0) dbias[o] = a;
    }
  }
}

constexpr int DX_S = 8;
static int dx_cpg(int HW) {
  int cpg = 4096 / HW;
  if (cpg < 1) cpg = 1;
  if (cpg > 16) cpg = 16;
  return cpg;
}

static int check(const kmu_deform_desc* d, const char* who) {
  KMU_REQUIRE(d != nullptr, KMU_ERR_BAD_ARG, "%s: null descriptor", who);
  KMU_REQUIRE(d->B > 0 && d->C > 0 && d->H > 0 && d->W > 0 && d->Cout > 0, KMU_ERR_BAD_ARG, "%s: non-positive shape", who);
  KMU_REQUIRE(d->B <= 65535, KMU_ERR_UNSUPPORTED, "%s: B=%d > 65535", who, d->B);
  KMU_REQUIRE(d->C <= 256 && d->Cout <= 256, KMU_ERR_UNSUPPORTED, "%s: C=%d / Cout=%d > 256", who, d->C, d->Cout);
  KMU_REQUIRE((long long)d->H * d->W <= 4096, KMU_ERR_UNSUPPORTED,
              "%s: H*W=%d > 4096 (the deterministic dX pass keeps whole channel images in shared memory)", who, d->H * d->W);
  return KMU_OK;
}

}  // namespace deform
}  // namespace kmu

using namespace kmu;
using namespace kmu::deform;

extern "C" {

size_t kmu_deformconv3x3_bwd_workspace_bytes(const kmu_deform_desc* d) {
  if (check(d, "deformconv3x3_bwd_workspace_bytes") != KMU_OK) return 0;
  const size_t HW = (size_t)d->H * d->W, K = (size_t)d->C * 9;
  return align_up((size_t)d->B * 9 * HW * d->C * 4, 256) + align_up((size_t)d->B * d->Cout * K * 4, 256);
}

int kmu_deformconv3x3_fwd(const kmu_deform_desc* d, const float* x, const float* offset, const float* weight, const float* bias,
                          float* out, kmu_stream stream) {
  int rc = check(d, "deformconv3x3_fwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(x && offset && weight && out, KMU_ERR_BAD_ARG, "deformconv3x3_fwd: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = d->H * d->W, K = d->C * 9;
  const size_t smem = ((size_t)K * PT + (size_t)KC * (d->Cout + 1)) * 4 + (size_t)9 * PT * sizeof(Geo);
  KMU_REQUIRE(smem <= 200 * 1024, KMU_ERR_UNSUPPORTED, "deformconv3x3_fwd: C=%d needs %zu B of shared memory", d->C, smem);
  if (smem > 48 * 1024) cudaFuncSetAttribute(deform_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  deform_fwd_kernel<<<dim3(cdiv(HW, PT), d->B), 256, smem, st>>>(x, offset, weight, bias, out, d->C, d->H, d->W, d->Cout);
  KMU_LAUNCH_CHECK("deform_fwd");
  return KMU_OK;
}

int kmu_deformconv3x3_bwd(const kmu_deform_desc* d, const float* x, const float* offset, const float* weight, const float* dout,
                          float* dx, float* doffset, float* dweight, float* dbias, void* workspace, size_t workspace_bytes,
                          kmu_stream stream) {
  int rc = check(d, "deformconv3x3_bwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(x && offset && weight && dout && dx && doffset && dweight, KMU_ERR_BAD_ARG, "deformconv3x3_bwd: null tensor");
  KMU_REQUIRE(workspace && workspace_bytes >= kmu_deformconv3x3_bwd_workspace_bytes(d), KMU_ERR_WORKSPACE,
              "deformconv3x3_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = d->H * d->W, K = d->C * 9;
  float* dcol = (float*)workspace;
  float* part = (float*)((char*)workspace + align_up((size_t)d->B * 9 * HW * d->C * 4, 256));
  {
    const size_t smem = ((size_t)d->Cout * PT + (size_t)9 * PT * (d->C + 1)) * 4 + (size_t)9 * PT * sizeof(Geo);
    KMU_REQUIRE(smem <= 200 * 1024, KMU_ERR_UNSUPPORTED, "deformconv3x3_bwd: C=%d needs %zu B of shared memory", d->C, smem);
    if (smem > 48 * 1024) cudaFuncSetAttribute(deform_dcol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    deform_dcol_kernel<<<dim3(cdiv(HW, PT), d->B), 256, smem, st>>>(x, offset, weight, dout, dcol, doffset, d->C, d->H, d->W, d->Cout);
    KMU_LAUNCH_CHECK("deform_dcol");
  }
  {
    const int cpg = dx_cpg(HW);
    const size_t smem = (size_t)DX_S * cpg * (HW | 1) * 4;
    if (smem > 48 * 1024) cudaFuncSetAttribute(deform_dx_kernel<DX_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    deform_dx_kernel<DX_S><<<dim3(cdiv(d->C, cpg), d->B), cpg * DX_S, smem, st>>>(offset, dcol, dx, d->C, d->H, d->W, cpg);
    KMU_LAUNCH_CHECK("deform_dx");
  }
  deform_dw_kernel<<<dim3(cdiv(K, 64), d->B), 256, 0, st>>>(x, offset, dout, part, d->C, d->H, d->W, d->Cout);
  KMU_LAUNCH_CHECK("deform_dw");
  const int nred = cdiv(d->Cout * K, 256);
  deform_dw_reduce_kernel<<<nred > cdiv(d->Cout, 8) ? nred : cdiv(d->Cout, 8), 256, 0, st>>>(part, dout, dweight, dbias, d->B, d->Cout, K,
                                                                                              HW);
  KMU_LAUNCH_CHECK("deform_dw_reduce");
  return KMU_OK;
}

}  // extern "C"
