// dysample.cu -- DySample ('lp' style) offset generation + bilinear point-sampling upsample, forward and backward.
//
// Replaces DySample_md.py:49-68: the reference's conv1x1 -> *0.25 + init_pos -> meshgrid/normalise/pixel_shuffle/permute
// -> F.grid_sample(bilinear, border, align_corners=False) chain collapses to
//     out[b, g*Cg+c, s*h+i, s*w+j] = bilinear(x[b, g*Cg+c]; row = clamp(h + off_y, 0, H-1), col = clamp(w + off_x, 0, W-1))
// with off_x = offset[b, (g*s+i)*s+j, h, w], off_y = offset[b, G*s*s + (g*s+i)*s+j, h, w]   (SURVEY appendix A.3).
// HBM-bound gather: reads x once, writes s*s times as much; stores are 128-bit, loads of the four taps of
// neighbouring output pixels fall in the same cache lines.
#include <cstdlib>

#include "common.cuh"

namespace kmu {
namespace dys {

struct Dims {
  int B, C, H, W, s, G, Cg, NOFF, OH, OW;
};

static Dims make_dims(const kmu_dysample_desc& d) {
  Dims r;
  r.B = d.B; r.C = d.C; r.H = d.H; r.W = d.W; r.s = d.scale; r.G = d.groups;
  r.Cg = d.C / d.groups;
  r.NOFF = 2 * d.groups * d.scale * d.scale;
  r.OH = d.H * d.scale; r.OW = d.W * d.scale;
  return r;
}

// ---------------------------------------------------------------------------------------------- offset conv (fwd)
// thread = input pixel; 32 offset channels per pass (blockIdx.y); weights transposed in smem as [c][32].
__global__ void __launch_bounds__(128) dys_offset_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ init_pos,
                                                             float* __restrict__ offset, Dims d) {
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;  // [C][32]
  const int j0 = blockIdx.y * 32;
  for (int i = threadIdx.x; i < d.C * 32; i += 128) {
    int c = i >> 5, j = i & 31;
    w_s[i] = (j0 + j < d.NOFF) ? w[(size_t)(j0 + j) * d.C + c] : 0.f;
  }
  __syncthreads();
  const long long HW = (long long)d.H * d.W;
  long long p = (long long)blockIdx.x * 128 + threadIdx.x;
  if (p >= d.B * HW) return;
  int b = (int)(p / HW);
  long long r = p - b * HW;
  const float* xp = x + (size_t)b * d.C * HW + r;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  for (int c = 0; c < d.C; ++c) {
    float xv = __ldg(xp + (size_t)c * HW);
    const float4* wr = reinterpret_cast<const float4*>(w_s + c * 32);
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      float4 ww = wr[j4];
      acc[j4 * 4 + 0] = fmaf(xv, ww.x, acc[j4 * 4 + 0]);
      acc[j4 * 4 + 1] = fmaf(xv, ww.y, acc[j4 * 4 + 1]);
      acc[j4 * 4 + 2] = fmaf(xv, ww.z, acc[j4 * 4 + 2]);
      acc[j4 * 4 + 3] = fmaf(xv, ww.w, acc[j4 * 4 + 3]);
    }
  }
  float* op = offset + (size_t)b * d.NOFF * HW + r;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j0 + j < d.NOFF) op[(size_t)(j0 + j) * HW] = (acc[j] + bias[j0 + j]) * 0.25f + init_pos[j0 + j];
}

// ---------------------------------------------------------------------------------------------- sampling helpers
struct Coord {
  int x0, y0;
  float fx, fy;
  bool x1ok, y1ok;   // the +1 neighbour lies inside the image
  bool gx, gy;       // coordinate was not clipped -> offset gradient flows (PyTorch clip_coordinates_set_grad)
};

__device__ __forceinline__ Coord make_coord(float rx, float ry, int W, int H) {
  Coord c;
  c.gx = rx > 0.f && rx < (float)(W - 1);
  c.gy = ry > 0.f && ry < (float)(H - 1);
  float sx = fminf(fmaxf(rx, 0.f), (float)(W - 1));
  float sy = fminf(fmaxf(ry, 0.f), (float)(H - 1));
  float fx0 = floorf(sx), fy0 = floorf(sy);
  c.x0 = (int)fx0;
  c.y0 = (int)fy0;
  c.fx = sx - fx0;
  c.fy = sy - fy0;
  c.x1ok = c.x0 + 1 < W;
  c.y1ok = c.y0 + 1 < H;
  return c;
}

// ---------------------------------------------------------------------------------------------- sample (fwd)
// thread = VEC consecutive output columns of one (b, offset group, output row): the sampling coordinates depend on the
// group only, so they are computed ONCE and reused by the group's Cg channels (the first version recomputed them per
// channel and was instruction-bound at 80 % issue utilisation); per channel: 4*VEC tap loads, VEC outputs, one 128-bit store.
template <int VEC, int CPT>
__global__ void __launch_bounds__(256, 2) dys_sample_fwd_kernel(const float* __restrict__ x, const float* __restrict__ offset,
                                                                float* __restrict__ out, Dims d) {
  // CPT channels of the group per thread (CPT = 0: all of them): fewer channels per thread = more threads in flight for the
  // latency-bound tap loads, at the price of recomputing the coordinates Cg / CPT times
  const int cpt = CPT > 0 ? CPT : d.Cg;
  const int cq_n = d.Cg / cpt;
  const int owv = d.OW / VEC;
  long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  long long total = (long long)d.B * d.G * cq_n * d.OH * owv;
  if (idx >= total) return;
  int ow0 = (int)(idx % owv) * VEC;
  long long t = idx / owv;
  int oh = (int)(t % d.OH);
  t /= d.OH;
  int cq = (int)(t % cq_n);
  t /= cq_n;
  int g = (int)(t % d.G);
  int b = (int)(t / d.G);
  int h = oh / d.s, i = oh - h * d.s;
  const long long HW = (long long)d.H * d.W, OHW = (long long)d.OH * d.OW;
  const float* offb = offset + (size_t)b * d.NOFF * HW + (size_t)h * d.W;
  const int ss = d.s * d.s;
  int o00[VEC];                    // offset of the top-left tap inside a channel plane
  float w00[VEC], w01[VEC], w10[VEC], w11[VEC];
  int dxo[VEC], dyo[VEC];          // +1 column / +1 row step, 0 when that neighbour is outside (its weight is zeroed too)
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    int ow = ow0 + e;
    int w = ow / d.s, j = ow - w * d.s;
    int ch = (g * d.s + i) * d.s + j;
    float ox = __ldg(offb + (size_t)ch * HW + w);
    float oy = __ldg(offb + (size_t)(d.G * ss + ch) * HW + w);
    Coord q = make_coord((float)w + ox, (float)h + oy, d.W, d.H);
    o00[e] = q.y0 * d.W + q.x0;
    dxo[e] = q.x1ok ? 1 : 0;
    dyo[e] = q.y1ok ? d.W : 0;
    const float wx1 = q.x1ok ? q.fx : 0.f, wy1 = q.y1ok ? q.fy : 0.f;
    const float wx0 = 1.f - q.fx, wy0 = 1.f - q.fy;
    w00[e] = wx0 * wy0;
    w01[e] = wx1 * wy0;
    w10[e] = wx0 * wy1;
    w11[e] = wx1 * wy1;
  }
  const size_t c0 = (size_t)b * d.C + (size_t)g * d.Cg + (size_t)cq * cpt;
  const float* xp = x + c0 * HW;
  float* op = out + c0 * OHW + (size_t)oh * d.OW + ow0;
#pragma unroll 2
  for (int c = 0; c < cpt; ++c) {
    const float* xc = xp + (size_t)c * HW;
    float res[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float* r0 = xc + o00[e];
      res[e] = __ldg(r0) * w00[e] + __ldg(r0 + dxo[e]) * w01[e] + __ldg(r0 + dyo[e]) * w10[e] + __ldg(r0 + dyo[e] + dxo[e]) * w11[e];
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(op + (size_t)c * OHW) = make_float4(res[0], res[1], res[2], res[3]);
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) op[(size_t)c * OHW + e] = res[e];
    }
  }
}

// ---------------------------------------------------------------------------------------------- fused forward
// ONE kernel for DySample.forward_lp (scale 2, 4 groups): CTA = (b, band of TH input rows, full width).
//   1. the band of x (+ one halo row above and below, all C channels) is loaded ONCE with 128-bit loads into shared memory;
//   2. the 1x1 offset convolution (C -> 32) of the band's pixels is computed from that tile (no second pass over x, and the
//      offset tensor is only written to HBM when the caller wants it saved for backward);
//   3. every output pixel (2 TH rows x 2 W columns x C channels) takes its four bilinear taps from shared memory (conflict-free:
//      a warp's taps span 17 consecutive floats) and is written with 128-byte coalesced streaming stores.  Samples whose taps leave the staged rows (|offset| > 1 row) read global memory.
// HBM traffic = x once (+ 50 % halo rows) + out once: the algorithmic 20 B / input element + 2 B.
constexpr int FT = 256;   // threads of the fused kernel
// W (= H: square maps) and C are template parameters: every stride of the sampling loop is then an immediate (the first version spent
// ~20 integer instructions per output element on address arithmetic and was issue-bound at 69 % with the LSU half idle).
template <int TH, int W, int C>
__global__ void __launch_bounds__(FT, 2) dys_fused_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, const float* __restrict__ init_pos,
                                                              float* __restrict__ offset, float* __restrict__ out, int B) {
  extern __shared__ __align__(16) float sm[];
  constexpr int H = W, NP = TH * W, ROWS = TH + 2;
  float* xs = sm;                                 // [C][ROWS][W]
  float* ws = xs + (size_t)C * ROWS * W;          // [C][32]
  float* offs = ws + (size_t)C * 32;              // [32][NP]
  const int tid = threadIdx.x, b = blockIdx.y, h0 = blockIdx.x * TH;
  constexpr size_t HW = (size_t)H * W;
  const float* xb = x + (size_t)b * C * HW;
  (void)B;
  // ---- 1. stage
  for (int i = tid; i < C * 32; i += FT) {
    const int c = i >> 5, j = i & 31;
    ws[i] = __ldg(w + (size_t)j * C + c);
  }
  {
    constexpr int W4 = W >> 2, lw = W == 16 ? 2 : W == 32 ? 3 : W == 64 ? 4 : 5;
    constexpr int n4 = C * ROWS * W4;
    for (int i = tid; i < n4; i += FT) {
      const int q = i & (W4 - 1), cr = i >> lw;                      // cr = c * ROWS + r
      const int c = cr / ROWS, r = cr - c * ROWS;
      const int gy = h0 - 1 + r;
      if (gy >= 0 && gy < H)
        *reinterpret_cast<float4*>(xs + (size_t)cr * W + 4 * q) = __ldg(reinterpret_cast<const float4*>(xb + (size_t)c * HW + (size_t)gy * W) + q);
    }
  }
  __syncthreads();
  // ---- 2. offsets: thread = (pixel, block of OCB offset channels), weights as broadcast 128-bit loads
  {
    constexpr int TPP = FT / NP, OCB = 32 / TPP;  // threads per pixel, offset channels per thread: 4, 8, 16 or 32
    const int p = tid % NP, oc0 = (tid / NP) * OCB;
    const int hl = p / W, wv = p - hl * W;
    if (h0 + hl < H) {
      float4 acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* xp = xs + (size_t)(hl + 1) * W + wv;
      for (int c = 0; c < C; ++c) {
        const float xv = xp[(size_t)c * ROWS * W];
        const float4* wr = reinterpret_cast<const float4*>(ws + c * 32 + oc0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (4 * j < OCB) {
            const float4 ww = wr[j];
            acc[j].x = fmaf(xv, ww.x, acc[j].x); acc[j].y = fmaf(xv, ww.y, acc[j].y);
            acc[j].z = fmaf(xv, ww.z, acc[j].z); acc[j].w = fmaf(xv, ww.w, acc[j].w);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (4 * j < OCB) {
          const float a4[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int oc = oc0 + 4 * j + e;
            const float v = (a4[e] + __ldg(bias + oc)) * 0.25f + __ldg(init_pos + oc);
            offs[oc * NP + p] = v;
            if (offset) offset[((size_t)b * 32 + oc) * HW + (size_t)(h0 + hl) * W + wv] = v;
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- 3. sample: thread = one output pixel of one group (lanes = consecutive output columns: pairs of lanes share their taps'
  //         columns, a warp's taps span 16 + 1 consecutive floats of a row -> conflict-free LDS, 128-byte coalesced stores)
  constexpr int OW = 2 * W, Cg = C >> 2;
  constexpr size_t OHW = (size_t)4 * HW;
  constexpr int items = 4 * (2 * TH) * OW;
  for (int it = tid; it < items; it += FT) {
    const int ow = it % OW, orl = (it / OW) % (2 * TH), g = it / (OW * 2 * TH);
    const int hl = orl >> 1, i = orl & 1, h = h0 + hl;
    if (h >= H) continue;
    const int wv = ow >> 1, j = ow & 1;
    const int ch = (g * 2 + i) * 2 + j;
    const float ox = offs[ch * NP + hl * W + wv], oy = offs[(16 + ch) * NP + hl * W + wv];
    const Coord c = make_coord((float)wv + ox, (float)h + oy, W, H);
    const int yl = c.y0 - (h0 - 1);                         // row inside the staged tile
    const bool inwin = yl >= 0 && yl + (c.y1ok ? 1 : 0) < ROWS;
    const int o00 = yl * W + c.x0, dxo = c.x1ok ? 1 : 0, dyo = c.y1ok ? W : 0;
    const float wx1 = c.x1ok ? c.fx : 0.f, wy1 = c.y1ok ? c.fy : 0.f;
    const float wx0 = 1.f - c.fx, wy0 = 1.f - c.fy;
    const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;
    float* op = out + ((size_t)b * C + (size_t)g * Cg) * OHW + (size_t)(2 * h + i) * OW + ow;
    if (inwin) {
      const float* t0 = xs + (size_t)g * Cg * ROWS * W + o00;
      const float *t1 = t0 + dxo, *t2 = t0 + dyo, *t3 = t0 + dyo + dxo;
#pragma unroll
      for (int cc = 0; cc < Cg; ++cc)
        __stcs(op + (size_t)cc * OHW, t0[cc * ROWS * W] * w00 + t1[cc * ROWS * W] * w01 + t2[cc * ROWS * W] * w10 + t3[cc * ROWS * W] * w11);
    } else {                                                // a tap left the staged rows: same arithmetic from global memory
      const float* t0 = xb + (size_t)g * Cg * HW + (ptrdiff_t)(h0 - 1) * W + o00;
      for (int cc = 0; cc < Cg; ++cc) {
        const float* t = t0 + (size_t)cc * HW;
        __stcs(op + (size_t)cc * OHW, __ldg(t) * w00 + __ldg(t + dxo) * w01 + __ldg(t + dyo) * w10 + __ldg(t + dyo + dxo) * w11);
      }
    }
  }
}

static bool fused_fwd_ok(const Dims& d, int* th, size_t* smem) {
  if (d.s != 2 || d.G != 4 || d.C != 64 || d.H != d.W) return false;
  if (d.W != 16 && d.W != 32 && d.W != 64 && d.W != 128) return false;
  const int TH = 2;
  if (TH * d.W > FT) return false;
  *th = TH;
  *smem = ((size_t)d.C * (TH + 2) * d.W + (size_t)d.C * 32 + (size_t)32 * TH * d.W) * 4;
  return *smem <= 200 * 1024;
}

// ---------------------------------------------------------------------------------------------- sample (bwd)
// thread = one output pixel of one (b, group): loops over the group's channels, scatters dX with fp32 atomics (like
// PyTorch's grid_sampler backward) and keeps the offset gradient in registers (no atomics: one owner per offset).
// FIX: dX is accumulated as 64-bit fixed point (value * 2^shift, integer atomics: the sum does not depend on the order of the adds ->
// bit-reproducible); `fixscale` holds 2^shift chosen from max|dout| by dys_amax / dys_fixscale so that 2^20 terms cannot overflow.
template <bool FIX>
__global__ void __launch_bounds__(256) dys_sample_bwd_kernel(const float* __restrict__ x, const float* __restrict__ offset,
                                                             const float* __restrict__ dout, float* __restrict__ dx,
                                                             float* __restrict__ doffset, Dims d, unsigned long long* __restrict__ acc,
                                                             const float* __restrict__ fixscale) {
  long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  long long total = (long long)d.B * d.G * d.OH * d.OW;
  if (idx >= total) return;
  int ow = (int)(idx % d.OW);
  long long t = idx / d.OW;
  int oh = (int)(t % d.OH);
  t /= d.OH;
  int g = (int)(t % d.G);
  int b = (int)(t / d.G);
  int h = oh / d.s, i = oh - h * d.s;
  int w = ow / d.s, j = ow - w * d.s;
  const long long HW = (long long)d.H * d.W, OHW = (long long)d.OH * d.OW;
  const int ss = d.s * d.s;
  int ch = (g * d.s + i) * d.s + j;
  size_t offi = (size_t)b * d.NOFF * HW + (size_t)ch * HW + (size_t)h * d.W + w;
  float ox = offset[offi], oy = offset[offi + (size_t)d.G * ss * HW];
  Coord q = make_coord((float)w + ox, (float)h + oy, d.W, d.H);
  float wx1 = q.fx, wx0 = 1.f - q.fx, wy1 = q.fy, wy0 = 1.f - q.fy;
  size_t base = ((size_t)b * d.C + (size_t)g * d.Cg) * HW + (size_t)q.y0 * d.W + q.x0;
  const float* dop = dout + ((size_t)b * d.C + (size_t)g * d.Cg) * OHW + (size_t)oh * d.OW + ow;
  float gx = 0.f, gy = 0.f;
  const float fs = FIX ? __ldg(fixscale) : 1.f;
  for (int c = 0; c < d.Cg; ++c) {
    float go = __ldg(dop + (size_t)c * OHW);
    const float* r0 = x + base + (size_t)c * HW;
    float v00 = __ldg(r0);
    float v01 = q.x1ok ? __ldg(r0 + 1) : 0.f;
    float v10 = q.y1ok ? __ldg(r0 + d.W) : 0.f;
    float v11 = (q.x1ok && q.y1ok) ? __ldg(r0 + d.W + 1) : 0.f;
    if (FIX) {
      unsigned long long* a0 = acc + base + (size_t)c * HW;
      const float gs = go * fs;
      atomicAdd(a0, (unsigned long long)__float2ll_rn(gs * (wx0 * wy0)));
      if (q.x1ok) atomicAdd(a0 + 1, (unsigned long long)__float2ll_rn(gs * (wx1 * wy0)));
      if (q.y1ok) atomicAdd(a0 + d.W, (unsigned long long)__float2ll_rn(gs * (wx0 * wy1)));
      if (q.x1ok && q.y1ok) atomicAdd(a0 + d.W + 1, (unsigned long long)__float2ll_rn(gs * (wx1 * wy1)));
    } else {
      float* g0 = dx + base + (size_t)c * HW;
      atomicAdd(g0, go * (wx0 * wy0));
      if (q.x1ok) atomicAdd(g0 + 1, go * (wx1 * wy0));
      if (q.y1ok) atomicAdd(g0 + d.W, go * (wx0 * wy1));
      if (q.x1ok && q.y1ok) atomicAdd(g0 + d.W + 1, go * (wx1 * wy1));
    }
    gx = fmaf(go, (v01 - v00) * wy0 + (v11 - v10) * wy1, gx);
    gy = fmaf(go, (v10 - v00) * wx0 + (v11 - v01) * wx1, gy);
  }
  doffset[offi] = q.gx ? gx : 0.f;
  doffset[offi + (size_t)d.G * ss * HW] = q.gy ? gy : 0.f;
}

// max|dout| (order-independent: integer max of the float bit patterns of |v|) and the fixed-point scale derived from it
__global__ void __launch_bounds__(256) dys_amax_kernel(const float* __restrict__ v, long long n4, unsigned int* __restrict__ out) {
  unsigned int m = 0u;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(v) + i);
    m = max(max(m, __float_as_uint(fabsf(t.x))), max(__float_as_uint(fabsf(t.y)), max(__float_as_uint(fabsf(t.z)), __float_as_uint(fabsf(t.w)))));
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}
__global__ void dys_fixscale_kernel(const unsigned int* __restrict__ amax_bits, float* __restrict__ scale) {
  // 2^shift with max|dout| * 2^shift in [2^40, 2^41): a position receives far fewer than 2^20 contributions of at most that size
  const float a = __uint_as_float(*amax_bits);
  int e = 0;
  if (a > 0.f && isfinite(a)) frexpf(a, &e);              // a = f * 2^e, f in [0.5, 1)
  scale[0] = ldexpf(1.f, 41 - e);
  scale[1] = ldexpf(1.f, e - 41);
}

// ---------------------------------------------------------------------------------------------- offset conv (bwd)
// thread = input pixel, 32 offset channels per pass.  dX += W^T (0.25 dOff) with atomics (several passes / the sampler's
// scatter share dX); per-block partials of dW = (0.25 dOff) x^T and db go to the workspace, reduced in fixed order.
__global__ void __launch_bounds__(128) dys_offset_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ doffset, float* __restrict__ dx,
                                                             float* __restrict__ partial, Dims d, const unsigned long long* __restrict__ acc,
                                                             const float* __restrict__ fixscale) {
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;                    // [C][32]
  float* g_s = w_s + d.C * 32;          // [128][33]
  float* x_s = g_s + 128 * 33;          // [128][36]
  const int tid = threadIdx.x;
  const int j0 = blockIdx.y * 32;
  for (int i = tid; i < d.C * 32; i += 128) {
    int c = i >> 5, j = i & 31;
    w_s[i] = (j0 + j < d.NOFF) ? w[(size_t)(j0 + j) * d.C + c] : 0.f;
  }
  const long long HW = (long long)d.H * d.W;
  long long p = (long long)blockIdx.x * 128 + tid;
  const bool valid = p < d.B * HW;
  int b = valid ? (int)(p / HW) : 0;
  long long r = valid ? p - b * HW : 0;
  float g[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    g[j] = (valid && j0 + j < d.NOFF) ? 0.25f * doffset[(size_t)b * d.NOFF * HW + (size_t)(j0 + j) * HW + r] : 0.f;
    g_s[tid * 33 + j] = g[j];
  }
  __syncthreads();
  // dX
  if (valid) {
    float* dxp = dx + (size_t)b * d.C * HW + r;
    for (int c = 0; c < d.C; ++c) {
      const float4* wr = reinterpret_cast<const float4*>(w_s + c * 32);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        float4 ww = wr[j4];
        s0 = fmaf(g[j4 * 4 + 0], ww.x, s0);
        s1 = fmaf(g[j4 * 4 + 1], ww.y, s1);
        s2 = fmaf(g[j4 * 4 + 2], ww.z, s2);
        s3 = fmaf(g[j4 * 4 + 3], ww.w, s3);
      }
      if (acc)    // fixed-point sampler sums -> float, plus this (single-pass) projection term: plain store, no atomics
        dxp[(size_t)c * HW] = (float)(long long)acc[(size_t)b * d.C * HW + (size_t)c * HW + r] * __ldg(fixscale + 1) + ((s0 + s1) + (s2 + s3));
      else
        atomicAdd(dxp + (size_t)c * HW, (s0 + s1) + (s2 + s3));
    }
  }
  // dW partial: thread owns offset channel j = lane, channels cq*8..cq*8+7 of each 32-channel chunk (cq = warp id)
  const int j = tid & 31, cq = tid >> 5;
  float* pb = partial + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * (32 * d.C + 32);
  for (int c0 = 0; c0 < d.C; c0 += 32) {
    __syncthreads();
    for (int cc = 0; cc < 32; ++cc) {
      int c = c0 + cc;
      x_s[tid * 36 + cc] = (valid && c < d.C) ? __ldg(x + (size_t)b * d.C * HW + (size_t)c * HW + r) : 0.f;
    }
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int px = 0; px < 128; ++px) {
      float gv = g_s[px * 33 + j];
      const float4* xr = reinterpret_cast<const float4*>(x_s + px * 36 + cq * 8);
      float4 a = xr[0], c4 = xr[1];
      acc[0] = fmaf(gv, a.x, acc[0]);
      acc[1] = fmaf(gv, a.y, acc[1]);
      acc[2] = fmaf(gv, a.z, acc[2]);
      acc[3] = fmaf(gv, a.w, acc[3]);
      acc[4] = fmaf(gv, c4.x, acc[4]);
      acc[5] = fmaf(gv, c4.y, acc[5]);
      acc[6] = fmaf(gv, c4.z, acc[6]);
      acc[7] = fmaf(gv, c4.w, acc[7]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c = c0 + cq * 8 + i;
      if (c < d.C) pb[(size_t)j * d.C + c] = acc[i];
    }
  }
  if (cq == 0) {
    float sb = 0.f;
    for (int px = 0; px < 128; ++px) sb += g_s[px * 33 + j];
    pb[(size_t)32 * d.C + j] = sb;
  }
}

// fixed-order sum of the per-block partials: CTA = 32 outputs x 8 interleaved slices of the block list
__global__ void __launch_bounds__(256) dys_offset_bwd_reduce_kernel(const float* __restrict__ partial, int nblk, int npass,
                                                                    float* __restrict__ dw, float* __restrict__ db, Dims d) {
  __shared__ float red[8][33];
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + o;
  const int per = 32 * d.C + 32;
  float s = 0.f;
  int pass = 0, e = 0;
  if (idx < npass * per) {
    pass = idx / per;
    e = idx - pass * per;
    for (int k = sl; k < nblk; k += 8) s += partial[((size_t)pass * nblk + k) * per + e];
  }
  red[sl][o] = s;
  __syncthreads();
  if (sl == 0 && idx < npass * per) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][o];
    if (e < 32 * d.C) {
      int j = pass * 32 + e / d.C, c = e % d.C;
      if (j < d.NOFF) dw[(size_t)j * d.C + c] = t;
    } else {
      int j = pass * 32 + (e - 32 * d.C);
      if (j < d.NOFF) db[j] = t;
    }
  }
}

__global__ void zero_kernel(float* __restrict__ p, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = 0.f;
}

static int check(const kmu_dysample_desc* d, const char* who) {
  KMU_REQUIRE(d != nullptr, KMU_ERR_BAD_ARG, "%s: null descriptor", who);
  KMU_REQUIRE(d->B > 0 && d->C > 0 && d->H > 0 && d->W > 0 && d->scale > 0 && d->groups > 0, KMU_ERR_BAD_ARG,
              "%s: non-positive shape", who);
  KMU_REQUIRE(d->C % d->groups == 0, KMU_ERR_BAD_ARG, "%s: channels %d not divisible by groups %d", who, d->C, d->groups);
  KMU_REQUIRE(d->C <= 1024, KMU_ERR_UNSUPPORTED, "%s: C=%d > 1024 not supported", who, d->C);
  return KMU_OK;
}

static int launch_sample_fwd(const Dims& d, const float* x, const float* offset, float* out, cudaStream_t st) {
  // threads: one per (b, group, channel quarter, output row, VEC output columns)
  // measured (B=32, C=64): all 16 channels of a group per thread wins from 32x32 inputs on (25.9 vs 30.0 us, 87 vs 106 us at 64x64),
  // 4 channels per thread wins while the grid would otherwise not fill the machine (10.2 vs 14.5 us at 16x16)
  const long long group_ctas = cdiv((long long)d.B * d.G * d.OH * (d.OW / 4), 256);
  if (d.OW % 4 == 0 && d.Cg % 4 == 0 && group_ctas < 2 * 148) {
    long long total = (long long)d.B * d.G * (d.Cg / 4) * d.OH * (d.OW / 4);
    dys_sample_fwd_kernel<4, 4><<<cdiv(total, 256), 256, 0, st>>>(x, offset, out, d);
  } else if (d.OW % 4 == 0) {
    long long total = (long long)d.B * d.G * d.OH * (d.OW / 4);
    dys_sample_fwd_kernel<4, 0><<<cdiv(total, 256), 256, 0, st>>>(x, offset, out, d);
  } else {
    long long total = (long long)d.B * d.G * d.OH * d.OW;
    dys_sample_fwd_kernel<1, 0><<<cdiv(total, 256), 256, 0, st>>>(x, offset, out, d);
  }
  KMU_LAUNCH_CHECK("dys_sample_fwd");
  return KMU_OK;
}

static int launch_sample_bwd(const Dims& d, const float* x, const float* offset, const float* dout, float* dx, float* doffset,
                             cudaStream_t st, unsigned long long* acc = nullptr, const float* fixscale = nullptr) {
  long long total = (long long)d.B * d.G * d.OH * d.OW;
  if (acc)
    dys_sample_bwd_kernel<true><<<cdiv(total, 256), 256, 0, st>>>(x, offset, dout, dx, doffset, d, acc, fixscale);
  else
    dys_sample_bwd_kernel<false><<<cdiv(total, 256), 256, 0, st>>>(x, offset, dout, dx, doffset, d, nullptr, nullptr);
  KMU_LAUNCH_CHECK("dys_sample_bwd");
  return KMU_OK;
}
// deterministic path: one pass of the offset projection (NOFF <= 32) and 16-byte aligned dout
static bool fix_ok(const Dims& d, const float* dout) { return d.NOFF <= 32 && ((uintptr_t)dout & 15) == 0 && ((long long)d.B * d.C * d.OH * d.OW) % 4 == 0; }

}  // namespace dys
}  // namespace kmu

using namespace kmu;
using namespace kmu::dys;

extern "C" {

size_t kmu_dysample_bwd_workspace_bytes(const kmu_dysample_desc* dd) {
  if (check(dd, "dysample_bwd_workspace_bytes") != KMU_OK) return 0;
  Dims d = make_dims(*dd);
  long long npix = (long long)d.B * d.H * d.W;
  size_t doff = align_up((size_t)d.B * d.NOFF * d.H * d.W * 4, 256);
  size_t part = (size_t)cdiv(npix, 128) * cdiv(d.NOFF, 32) * (32 * d.C + 32) * 4;
  return doff + align_up(part, 256) + align_up((size_t)npix * d.C * 8, 256) + 256;      // + fixed-point dX accumulator + scale
}

int kmu_dysample_fwd(const kmu_dysample_fwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "dysample_fwd: null args");
  int st_ = check(&a->d, "dysample_fwd");
  if (st_ != KMU_OK) return st_;
  KMU_REQUIRE(a->x && a->w_offset && a->b_offset && a->init_pos && a->out, KMU_ERR_BAD_ARG, "dysample_fwd: null tensor");
  Dims d = make_dims(a->d);
  cudaStream_t st = (cudaStream_t)stream;
  {
    int th = 0;
    size_t fsm = 0;
    static const bool fused_on = !(getenv("KMU_DYS_FUSED") && getenv("KMU_DYS_FUSED")[0] == '0');
    if (fused_on && fused_fwd_ok(d, &th, &fsm) && ((uintptr_t)a->x & 15) == 0 && ((uintptr_t)a->out & 15) == 0) {
#define KMU_DYS_FUSED(WW)                                                                                                          \
  do {                                                                                                                             \
    cudaError_t e = cudaFuncSetAttribute(dys_fused_fwd_kernel<2, WW, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);   \
    KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "dysample_fwd: cannot opt in to %zu B shared memory: %s", fsm,                  \
                cudaGetErrorString(e));                                                                                            \
    dys_fused_fwd_kernel<2, WW, 64><<<dim3(cdiv(d.H, th), d.B), FT, fsm, st>>>(a->x, a->w_offset, a->b_offset, a->init_pos,       \
                                                                                a->offset, a->out, d.B);                           \
  } while (0)
      switch (d.W) {
        case 16: KMU_DYS_FUSED(16); break;
        case 32: KMU_DYS_FUSED(32); break;
        case 64: KMU_DYS_FUSED(64); break;
        default: KMU_DYS_FUSED(128); break;
      }
#undef KMU_DYS_FUSED
      KMU_LAUNCH_CHECK("dys_fused_fwd");
      return KMU_OK;
    }
  }
  KMU_REQUIRE(a->offset, KMU_ERR_BAD_ARG, "dysample_fwd: the two-kernel path needs the offset buffer");
  long long npix = (long long)d.B * d.H * d.W;
  size_t smem = (size_t)d.C * 32 * 4;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(dys_offset_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dys_offset_fwd_kernel<<<dim3(cdiv(npix, 128), cdiv(d.NOFF, 32)), 128, smem, st>>>(a->x, a->w_offset, a->b_offset, a->init_pos,
                                                                                   a->offset, d);
  KMU_LAUNCH_CHECK("dys_offset_fwd");
  return launch_sample_fwd(d, a->x, a->offset, a->out, st);
}

int kmu_dysample_sample_fwd(const kmu_dysample_desc* dd, const float* x, const float* offset, float* out, kmu_stream stream) {
  int st_ = check(dd, "dysample_sample_fwd");
  if (st_ != KMU_OK) return st_;
  KMU_REQUIRE(x && offset && out, KMU_ERR_BAD_ARG, "dysample_sample_fwd: null tensor");
  return launch_sample_fwd(make_dims(*dd), x, offset, out, (cudaStream_t)stream);
}

int kmu_dysample_sample_bwd(const kmu_dysample_desc* dd, const float* x, const float* offset, const float* dout, float* dx,
                            float* doffset, kmu_stream stream) {
  int st_ = check(dd, "dysample_sample_bwd");
  if (st_ != KMU_OK) return st_;
  KMU_REQUIRE(x && offset && dout && dx && doffset, KMU_ERR_BAD_ARG, "dysample_sample_bwd: null tensor");
  return launch_sample_bwd(make_dims(*dd), x, offset, dout, dx, doffset, (cudaStream_t)stream);
}

int kmu_dysample_bwd(const kmu_dysample_bwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "dysample_bwd: null args");
  int st_ = check(&a->d, "dysample_bwd");
  if (st_ != KMU_OK) return st_;
  KMU_REQUIRE(a->x && a->w_offset && a->offset && a->dout && a->dx && a->d_w_offset && a->d_b_offset, KMU_ERR_BAD_ARG,
              "dysample_bwd: null tensor");
  size_t need = kmu_dysample_bwd_workspace_bytes(&a->d);
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= need, KMU_ERR_WORKSPACE, "dysample_bwd: workspace %zu < %zu",
              a->workspace_bytes, need);
  Dims d = make_dims(a->d);
  cudaStream_t st = (cudaStream_t)stream;
  long long npix = (long long)d.B * d.H * d.W;
  float* doff = (float*)a->workspace;
  float* partial = (float*)((char*)a->workspace + align_up((size_t)d.B * d.NOFF * d.H * d.W * 4, 256));
  long long nx = npix * d.C;
  size_t part_bytes = align_up((size_t)cdiv(npix, 128) * cdiv(d.NOFF, 32) * (32 * d.C + 32) * 4, 256);
  unsigned long long* acc = nullptr;
  float* fixscale = nullptr;
  if (deterministic() && fix_ok(d, a->dout)) {      // bit-reproducible: fixed-point dX accumulation scaled by max|dout| (64-bit integer
                                                    // atomics cost 1.75x the fp32 ones: 253 -> 446 us at (32,64,64x64), hence opt-in)
    acc = (unsigned long long*)((char*)partial + part_bytes);
    fixscale = (float*)((char*)acc + align_up((size_t)nx * 8, 256));
    cudaMemsetAsync(acc, 0, (size_t)nx * 8, st);
    cudaMemsetAsync(fixscale, 0, 16, st);
    const long long n4 = (long long)d.B * d.C * d.OH * d.OW / 4;
    dys_amax_kernel<<<(int)std::min<long long>(cdiv(n4, 256), 148 * 8), 256, 0, st>>>(a->dout, n4, (unsigned int*)(fixscale + 2));
    KMU_LAUNCH_CHECK("dys_amax");
    dys_fixscale_kernel<<<1, 1, 0, st>>>((const unsigned int*)(fixscale + 2), fixscale);
    KMU_LAUNCH_CHECK("dys_fixscale");
  } else {
    zero_kernel<<<(int)std::min<long long>(cdiv(nx, 1024), 148 * 8), 256, 0, st>>>(a->dx, nx);
    KMU_LAUNCH_CHECK("dys_zero");
  }
  st_ = launch_sample_bwd(d, a->x, a->offset, a->dout, a->dx, doff, st, acc, fixscale);
  if (st_ != KMU_OK) return st_;
  int nblk = cdiv(npix, 128), npass = cdiv(d.NOFF, 32);
  size_t smem = ((size_t)d.C * 32 + 128 * 33 + 128 * 36) * 4;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(dys_offset_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dys_offset_bwd_kernel<<<dim3(nblk, npass), 128, smem, st>>>(a->x, a->w_offset, doff, a->dx, partial, d, acc, fixscale);
  KMU_LAUNCH_CHECK("dys_offset_bwd");
  int n = npass * (32 * d.C + 32);
  dys_offset_bwd_reduce_kernel<<<cdiv(n, 32), 256, 0, st>>>(partial, nblk, npass, a->d_w_offset, a->d_b_offset, d);
  KMU_LAUNCH_CHECK("dys_offset_bwd_reduce");
  return KMU_OK;
}

}  // extern "C"
