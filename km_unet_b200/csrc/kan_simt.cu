// kan_simt.cu -- fp32 CUDA-core family of KANConv2d / KANLinear (KMU_PREC_FP32).
//
// Replaces convKAN/KANConv2Dlayers.py:15-37 + convKAN/KANlayers.py:577-610,644-660 (forward) and the autograd graph
// PyTorch builds for them (backward).  General in kernel size / stride / padding and in the per-feature knot table
// (so it stays correct after KANLinear.update_grid); cubic splines with 8 basis functions only.
//
// Never materialises the im2col matrix nor the (M, in, 8) basis tensor: Phi(x) = [SiLU(x), B_0..B_7(x)] is evaluated
// in registers per (output pixel, feature) and contracted immediately against the packed weights
//     Weff[f][q][o] = q == 0 ? base_weight[o,f] : spline_weight[o,f,q-1] * spline_scaler[o,f].
// This is the exact-parity path (1e-4 gate) and the fallback for shapes the tcgen05 path does not take.
#include "common.cuh"
#include "kan_common.cuh"

namespace kmu {
namespace kan {

// ------------------------------------------------------------------------------------------------ pack
__global__ void kan_pack_kernel(const float* __restrict__ base_w, const float* __restrict__ spline_w,
                                const float* __restrict__ scaler, const float* __restrict__ grid, float* __restrict__ weff,
                                float* __restrict__ ktab, int F, int Cout) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < F * Cout) {
    int o = idx / F, f = idx - o * F;
    float s = scaler ? scaler[idx] : 1.0f;
    weff[(size_t)(f * NPHI) * Cout + o] = base_w[idx];
#pragma unroll
    for (int j = 0; j < NB; ++j) weff[(size_t)(f * NPHI + 1 + j) * Cout + o] = spline_w[(size_t)idx * NB + j] * s;
  }
  if (idx < F) {
    float t[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) t[j] = grid[idx * NK + j];
    float* k = ktab + (size_t)idx * KT;
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = t[j];
#pragma unroll
    for (int j = 0; j < 11; ++j) k[KT_R1 + j] = 1.0f / (t[j + 1] - t[j]);
#pragma unroll
    for (int j = 0; j < 10; ++j) k[KT_R2 + j] = 1.0f / (t[j + 2] - t[j]);
#pragma unroll
    for (int j = 0; j < 9; ++j) k[KT_R3 + j] = 1.0f / (t[j + 3] - t[j]);
    k[42] = 0.f;
    k[43] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <int COUT_T>
__global__ void __launch_bounds__(128) kan_fwd_simt_kernel(const float* __restrict__ x, const float* __restrict__ weff,
                                                           const float* __restrict__ ktab, float* __restrict__ y, Dims d) {
  constexpr int FC = 8;
  __shared__ __align__(16) float w_s[FC][NPHI][COUT_T];
  __shared__ __align__(16) float k_s[FC][KT];
  const int tid = threadIdx.x;
  const int o0 = blockIdx.y * COUT_T;
  const long long m = (long long)blockIdx.x * 128 + tid;
  const bool valid = m < d.M;
  int b = 0, ho = 0, wo = 0;
  if (valid) {
    b = (int)(m / (d.Ho * d.Wo));
    int r = (int)(m - (long long)b * d.Ho * d.Wo);
    ho = r / d.Wo;
    wo = r - ho * d.Wo;
  }
  float acc[COUT_T];
#pragma unroll
  for (int i = 0; i < COUT_T; ++i) acc[i] = 0.f;
  const int kk = d.k * d.k;
  const float* xb = x + (size_t)b * d.Cin * d.H * d.W;

  for (int f0 = 0; f0 < d.F; f0 += FC) {
    __syncthreads();
    for (int i = tid; i < FC * NPHI * COUT_T; i += 128) {
      int fl = i / (NPHI * COUT_T), r = i - fl * (NPHI * COUT_T);
      int q = r / COUT_T, oo = r - q * COUT_T;
      int f = f0 + fl, o = o0 + oo;
      w_s[fl][q][oo] = (f < d.F && o < d.Cout) ? weff[(size_t)(f * NPHI + q) * d.Cout + o] : 0.f;
    }
    for (int i = tid; i < FC * KT; i += 128) {
      int fl = i / KT, f = f0 + fl;
      k_s[fl][i - fl * KT] = f < d.F ? ktab[(size_t)f * KT + (i - fl * KT)] : 0.f;
    }
    __syncthreads();
    if (!valid) continue;
    float xv[FC];
#pragma unroll
    for (int fl = 0; fl < FC; ++fl) {
      int f = f0 + fl;
      xv[fl] = 0.f;
      if (f < d.F) {
        int c = f / kk, t = f - c * kk;
        int ki = t / d.k, kj = t - ki * d.k;
        int hi = ho * d.stride - d.pad + ki, wi = wo * d.stride - d.pad + kj;
        if (hi >= 0 && hi < d.H && wi >= 0 && wi < d.W) xv[fl] = __ldg(xb + ((size_t)c * d.H + hi) * d.W + wi);
      }
    }
#pragma unroll
    for (int fl = 0; fl < FC; ++fl) {
      if (f0 + fl >= d.F) break;
      float phi[NPHI];
      eval_phi(xv[fl], k_s[fl], phi);
#pragma unroll
      for (int q = 0; q < NPHI; ++q) {
#pragma unroll
        for (int oo = 0; oo < COUT_T; oo += 4) {
          float4 w = *reinterpret_cast<const float4*>(&w_s[fl][q][oo]);
          acc[oo + 0] = fmaf(phi[q], w.x, acc[oo + 0]);
          acc[oo + 1] = fmaf(phi[q], w.y, acc[oo + 1]);
          acc[oo + 2] = fmaf(phi[q], w.z, acc[oo + 2]);
          acc[oo + 3] = fmaf(phi[q], w.w, acc[oo + 3]);
        }
      }
    }
  }
  if (valid) {
    float* yb = y + ((size_t)b * d.Cout) * d.Ho * d.Wo + (size_t)ho * d.Wo + wo;
#pragma unroll
    for (int oo = 0; oo < COUT_T; ++oo)
      if (o0 + oo < d.Cout) yb[(size_t)(o0 + oo) * d.Ho * d.Wo] = acc[oo];
  }
}

// ------------------------------------------------------------------------------------------------ backward: dX
// Gather form (deterministic, no atomics): one thread per INPUT pixel and CT channels; for every tap it finds the
// output pixel that read this input through that tap, forms dPhi = dY . Weff and contracts with Phi'(x).
template <int CT, int OT>
__global__ void __launch_bounds__(128) kan_bwd_dx_simt_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              const float* __restrict__ weff, const float* __restrict__ ktab,
                                                              float* __restrict__ dx, Dims d) {
  __shared__ __align__(16) float w_s[CT][NPHI][OT];
  __shared__ __align__(16) float k_s[CT][KT];
  const int tid = threadIdx.x;
  const int c0 = blockIdx.y * CT;
  const long long p = (long long)blockIdx.x * 128 + tid;
  const long long NP = (long long)d.B * d.H * d.W;
  const bool valid = p < NP;
  int b = 0, hi = 0, wi = 0;
  if (valid) {
    b = (int)(p / (d.H * d.W));
    int r = (int)(p - (long long)b * d.H * d.W);
    hi = r / d.W;
    wi = r - hi * d.W;
  }
  float xs[CT], acc[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) {
    acc[c] = 0.f;
    xs[c] = (valid && c0 + c < d.Cin) ? x[(((size_t)b * d.Cin + c0 + c) * d.H + hi) * d.W + wi] : 0.f;
  }
  const int kk = d.k * d.k;
  for (int t = 0; t < kk; ++t) {
    int ki = t / d.k, kj = t - ki * d.k;
    int th = hi + d.pad - ki, tw = wi + d.pad - kj;
    int ho = th / d.stride, wo = tw / d.stride;
    bool ok = valid && th >= 0 && tw >= 0 && ho * d.stride == th && wo * d.stride == tw && ho < d.Ho && wo < d.Wo;
    for (int o0 = 0; o0 < d.Cout; o0 += OT) {
      __syncthreads();
      for (int i = tid; i < CT * NPHI * OT; i += 128) {
        int c = i / (NPHI * OT), r = i - c * (NPHI * OT);
        int q = r / OT, oo = r - q * OT;
        int f = (c0 + c) * kk + t, o = o0 + oo;
        w_s[c][q][oo] = (c0 + c < d.Cin && o < d.Cout) ? weff[(size_t)(f * NPHI + q) * d.Cout + o] : 0.f;
      }
      for (int i = tid; i < CT * KT; i += 128) {
        int c = i / KT;
        k_s[c][i - c * KT] = (c0 + c < d.Cin) ? ktab[(size_t)((c0 + c) * kk + t) * KT + (i - c * KT)] : 0.f;
      }
      __syncthreads();
      if (!ok) continue;
      float dyv[OT];
      const float* dyp = dy + (((size_t)b * d.Cout + o0) * d.Ho + ho) * d.Wo + wo;
#pragma unroll
      for (int oo = 0; oo < OT; ++oo) dyv[oo] = (o0 + oo < d.Cout) ? __ldg(dyp + (size_t)oo * d.Ho * d.Wo) : 0.f;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        float dphi[NPHI];
        eval_dphi(xs[c], k_s[c], dphi);
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < NPHI; ++q) {
          float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
          for (int oo = 0; oo < OT; oo += 4) {
            float4 w = *reinterpret_cast<const float4*>(&w_s[c][q][oo]);
            g0 = fmaf(dyv[oo + 0], w.x, g0);
            g1 = fmaf(dyv[oo + 1], w.y, g1);
            g2 = fmaf(dyv[oo + 2], w.z, g2);
            g3 = fmaf(dyv[oo + 3], w.w, g3);
          }
          s = fmaf((g0 + g1) + (g2 + g3), dphi[q], s);
        }
        acc[c] += s;
      }
    }
  }
  if (valid) {
#pragma unroll
    for (int c = 0; c < CT; ++c)
      if (c0 + c < d.Cin) dx[(((size_t)b * d.Cin + c0 + c) * d.H + hi) * d.W + wi] = acc[c];
  }
}

// ------------------------------------------------------------------------------------------------ backward: dW
// dWeff[f][q][o] = sum_m Phi_q(x[m,f]) dY[m,o].  Block = (pixel split, 16 features, 64 outputs); per-split partials go
// to the workspace and are reduced in a fixed order by kan_bwd_dw_reduce_kernel (deterministic).
constexpr int DW_FT = 16, DW_OT = 64, DW_PC = 32;

__global__ void __launch_bounds__(256) kan_bwd_dw_simt_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              const float* __restrict__ ktab, float* __restrict__ partial,
                                                              Dims d, long long pix_per_split) {
  __shared__ __align__(16) float phi_s[DW_FT][DW_PC][12];
  __shared__ __align__(16) float dy_s[DW_PC][DW_OT + 4];
  __shared__ __align__(16) float k_s[DW_FT][KT];
  const int tid = threadIdx.x;
  const int f0 = blockIdx.y * DW_FT, o0 = blockIdx.z * DW_OT;
  const int fl_own = tid >> 4, og = tid & 15;
  const long long m_begin = (long long)blockIdx.x * pix_per_split;
  const long long m_end = min(m_begin + pix_per_split, d.M);
  const int kk = d.k * d.k;
  const int HoWo = d.Ho * d.Wo;
  for (int i = tid; i < DW_FT * KT; i += 256) {
    int fl = i / KT;
    k_s[fl][i - fl * KT] = (f0 + fl < d.F) ? ktab[(size_t)(f0 + fl) * KT + (i - fl * KT)] : 0.f;
  }
  float acc[NPHI][4];
#pragma unroll
  for (int q = 0; q < NPHI; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;

  for (long long p0 = m_begin; p0 < m_end; p0 += DW_PC) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < (DW_FT * DW_PC) / 256; ++r) {
      int idx = tid + r * 256;
      int fl = idx / DW_PC, px = idx - fl * DW_PC;
      long long m = p0 + px;
      int f = f0 + fl;
      float phi[12];
#pragma unroll
      for (int q = 0; q < 12; ++q) phi[q] = 0.f;
      if (m < m_end && f < d.F) {
        int b = (int)(m / HoWo);
        int rr = (int)(m - (long long)b * HoWo);
        int ho = rr / d.Wo, wo = rr - ho * d.Wo;
        int c = f / kk, t = f - c * kk;
        int ki = t / d.k, kj = t - ki * d.k;
        int hi = ho * d.stride - d.pad + ki, wi = wo * d.stride - d.pad + kj;
        float v = 0.f;
        if (hi >= 0 && hi < d.H && wi >= 0 && wi < d.W) v = __ldg(x + (((size_t)b * d.Cin + c) * d.H + hi) * d.W + wi);
        eval_phi(v, k_s[fl], phi);
      }
      float4* dst = reinterpret_cast<float4*>(&phi_s[fl][px][0]);
      dst[0] = make_float4(phi[0], phi[1], phi[2], phi[3]);
      dst[1] = make_float4(phi[4], phi[5], phi[6], phi[7]);
      dst[2] = make_float4(phi[8], 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < (DW_OT * DW_PC) / 256; ++r) {
      int idx = tid + r * 256;
      int oo = idx / DW_PC, px = idx - oo * DW_PC;
      long long m = p0 + px;
      float v = 0.f;
      if (m < m_end && o0 + oo < d.Cout) {
        int b = (int)(m / HoWo);
        int rr = (int)(m - (long long)b * HoWo);
        v = __ldg(dy + ((size_t)b * d.Cout + o0 + oo) * HoWo + rr);
      }
      dy_s[px][oo] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int px = 0; px < DW_PC; ++px) {
      const float4* ph = reinterpret_cast<const float4*>(&phi_s[fl_own][px][0]);
      float4 a0 = ph[0], a1 = ph[1], a2 = ph[2];
      float4 g = *reinterpret_cast<const float4*>(&dy_s[px][og * 4]);
      float a[NPHI] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x};
#pragma unroll
      for (int q = 0; q < NPHI; ++q) {
        acc[q][0] = fmaf(a[q], g.x, acc[q][0]);
        acc[q][1] = fmaf(a[q], g.y, acc[q][1]);
        acc[q][2] = fmaf(a[q], g.z, acc[q][2]);
        acc[q][3] = fmaf(a[q], g.w, acc[q][3]);
      }
    }
  }
  int f = f0 + fl_own;
  if (f < d.F) {
    float* dst = partial + ((size_t)blockIdx.x * d.F + f) * NPHI * d.Cout;
#pragma unroll
    for (int q = 0; q < NPHI; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int o = o0 + og * 4 + i;
        if (o < d.Cout) dst[(size_t)q * d.Cout + o] = acc[q][i];
      }
  }
}

// Fixed-order reduction over splits + chain rule into the three parameter tensors
// (d spline_weight = dWeff * s ; d spline_scaler = sum_j dWeff_j * spline_weight_j ; convKAN/KANlayers.py:644-650).
__global__ void kan_bwd_dw_reduce_kernel(const float* __restrict__ partial, int splits, const float* __restrict__ spline_w,
                                         const float* __restrict__ scaler, float* __restrict__ d_base,
                                         float* __restrict__ d_spline, float* __restrict__ d_scaler, int F, int Cout) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= F * Cout) return;
  int f = idx / Cout, o = idx - f * Cout;
  float s[NPHI];
#pragma unroll
  for (int q = 0; q < NPHI; ++q) s[q] = 0.f;
  for (int sp = 0; sp < splits; ++sp) {
    const float* src = partial + ((size_t)sp * F + f) * NPHI * Cout + o;
#pragma unroll
    for (int q = 0; q < NPHI; ++q) s[q] += src[(size_t)q * Cout];
  }
  size_t of = (size_t)o * F + f;
  d_base[of] = s[0];
  float sc = scaler ? scaler[of] : 1.0f;
  float dsc = 0.f;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    d_spline[of * NB + j] = s[1 + j] * sc;
    dsc = fmaf(s[1 + j], spline_w[of * NB + j], dsc);
  }
  if (d_scaler) d_scaler[of] = dsc;
}

// ------------------------------------------------------------------------------------------------ host side
static int dw_splits(const Dims& d) {
  long long tiles = (long long)cdiv(d.F, DW_FT) * cdiv(d.Cout, DW_OT);
  long long want = (4LL * 148 + tiles - 1) / tiles;
  long long max_splits = (d.M + DW_PC * 8 - 1) / (DW_PC * 8);
  long long s = want < max_splits ? want : max_splits;
  return (int)(s < 1 ? 1 : s);
}

size_t simt_fwd_workspace(const Dims& d) {
  return align_up((size_t)d.F * NPHI * d.Cout * 4, 256) + align_up((size_t)d.F * KT * 4, 256);
}
size_t simt_bwd_workspace(const Dims& d) {
  return simt_fwd_workspace(d) + align_up((size_t)dw_splits(d) * d.F * NPHI * d.Cout * 4, 256);
}

static int pack(const float* base_w, const float* spline_w, const float* scaler, const float* grid, float* weff, float* ktab,
                const Dims& d, cudaStream_t st) {
  int n = d.F * d.Cout;
  kan_pack_kernel<<<cdiv(n, 256), 256, 0, st>>>(base_w, spline_w, scaler, grid, weff, ktab, d.F, d.Cout);
  KMU_LAUNCH_CHECK("kan_pack");
  return KMU_OK;
}

int simt_forward(const kmu_kanconv2d_fwd_args* a, const Dims& d, cudaStream_t st) {
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= simt_fwd_workspace(d), KMU_ERR_WORKSPACE,
              "kanconv2d_fwd: workspace %zu < %zu", a->workspace_bytes, simt_fwd_workspace(d));
  float* weff = (float*)a->workspace;
  float* ktab = (float*)((char*)a->workspace + align_up((size_t)d.F * NPHI * d.Cout * 4, 256));
  int st_ = pack(a->base_weight, a->spline_weight, a->spline_scaler, a->grid, weff, ktab, d, st);
  if (st_ != KMU_OK) return st_;
  dim3 block(128);
  if (d.Cout <= 16) {
    kan_fwd_simt_kernel<16><<<dim3(cdiv(d.M, 128), 1), block, 0, st>>>(a->x, weff, ktab, a->y, d);
  } else if (d.Cout <= 32) {
    kan_fwd_simt_kernel<32><<<dim3(cdiv(d.M, 128), 1), block, 0, st>>>(a->x, weff, ktab, a->y, d);
  } else {
    kan_fwd_simt_kernel<64><<<dim3(cdiv(d.M, 128), cdiv(d.Cout, 64)), block, 0, st>>>(a->x, weff, ktab, a->y, d);
  }
  KMU_LAUNCH_CHECK("kan_fwd_simt");
  return KMU_OK;
}

int simt_backward(const kmu_kanconv2d_bwd_args* a, const Dims& d, cudaStream_t st) {
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= simt_bwd_workspace(d), KMU_ERR_WORKSPACE,
              "kanconv2d_bwd: workspace %zu < %zu", a->workspace_bytes, simt_bwd_workspace(d));
  char* ws = (char*)a->workspace;
  float* weff = (float*)ws;
  float* ktab = (float*)(ws + align_up((size_t)d.F * NPHI * d.Cout * 4, 256));
  float* partial = (float*)(ws + simt_fwd_workspace(d));
  int st_ = pack(a->base_weight, a->spline_weight, a->spline_scaler, a->grid, weff, ktab, d, st);
  if (st_ != KMU_OK) return st_;
  if (a->dx) {
    long long NP = (long long)d.B * d.H * d.W;
    kan_bwd_dx_simt_kernel<4, 32><<<dim3(cdiv(NP, 128), cdiv(d.Cin, 4)), 128, 0, st>>>(a->x, a->dy, weff, ktab, a->dx, d);
    KMU_LAUNCH_CHECK("kan_bwd_dx_simt");
  }
  if (a->d_base_weight) {
    int splits = dw_splits(d);
    long long pps = (d.M + splits - 1) / splits;
    pps = (pps + DW_PC - 1) / DW_PC * DW_PC;
    splits = cdiv(d.M, pps);
    kan_bwd_dw_simt_kernel<<<dim3(splits, cdiv(d.F, DW_FT), cdiv(d.Cout, DW_OT)), 256, 0, st>>>(a->x, a->dy, ktab, partial, d,
                                                                                                  pps);
    KMU_LAUNCH_CHECK("kan_bwd_dw_simt");
    int n = d.F * d.Cout;
    kan_bwd_dw_reduce_kernel<<<cdiv(n, 256), 256, 0, st>>>(partial, splits, a->spline_weight, a->spline_scaler,
                                                           a->d_base_weight, a->d_spline_weight, a->d_spline_scaler, d.F, d.Cout);
    KMU_LAUNCH_CHECK("kan_bwd_dw_reduce");
  }
  return KMU_OK;
}

}  // namespace kan
}  // namespace kmu
