// hsmssd.cu -- EfficientViM HSM-SSD hidden-state mixer, forward and backward, plus LayerNorm1D.
//
// Replaces vim_block_init/efficient_vim_init.py:33-61 (HSMSSD.forward) and vim_utils_init.py:50-59 (LayerNorm1D.forward)
// and the autograd graph PyTorch builds for them.  Arithmetic, x (B,C,L), L = H*H, N = 64 states:
//     Q = Wp x ; P = dw3x3(Q) ; [Bm; Cm; dt] = P ; A = softmax_L(dt)   (the per-state shift A_param cancels under the
//     over-L softmax, efficient_vim_init.py:46, so it is neither read nor given a gradient)
//     hs = x (A.Bm)^T ; [hh; z] = Whz hs ; ho = Wo (hh.SiLU(z) + hh D) ; y = ho Cm
// Forward kernels
//     hsm_proj_dw      x -> P          1x1 projection recomputed on a 1-pixel halo, depthwise 3x3 from shared memory
//     hsm_softmax_hs   P,x -> per-tile online-softmax partials (max, sum, unnormalised hs): one sweep over L
//     hsm_combine_gate partials -> stats, hs, hz, ho (one CTA per batch element; all (C,64)-sized)
//     hsm_out          ho,P -> y
// Backward kernels (three sweeps over L; the softmax-backward reduction sum_L(dA.A) equals sum_c dhs.hs, so it needs none)
//     hsm_contract     dy,Cm -> dho partials
//     hsm_gate_bwd     -> dhs, r, per-batch dWo / dWhz / dD
//     hsm_dp           -> dP = [dBm; dCm; ddt] and the direct part of dx
//     hsm_proj_dw_bwd  dP -> dx += Wp^T dQ, per-CTA partials of dWp and dWd ; hsm_wgrad_reduce sums them in fixed order.
// P is materialised in fp32 (B,192,L): with 180 GB of HBM3e that is cheaper than recomputing the projection in the
// three backward sweeps on CUDA cores.  All reductions are deterministic (no atomics).
#include <algorithm>
#include <cstdlib>
#include <cuda_bf16.h>

#include "common.cuh"

namespace kmu {
namespace hsm {

namespace tc {  // hsm_tc.cu: tcgen05 path of the projection
size_t pack_bytes(int C);
size_t weff_bytes(int B, int C);
int project(const float* x, const float* wp, const float* wd, float* P, int B, int C, int H, int slices, void* workspace,
            cudaStream_t st);
int out(const float* x, const float* ho, const float* wp, const float* wd, float* y, int B, int C, int H, void* workspace,
        cudaStream_t st);
}  // namespace tc
namespace tcb {  // hsm_tc_bwd.cu: tcgen05 / TMA path of the projection backward
size_t dpp_bytes(int B, int L);
size_t workspace_bytes(int B, int C, int H);
int contract(const float* x, const float* dy, const float* wp, const float* wd, float* dho, int B, int C, int H, void* workspace,
             cudaStream_t st);
int backward(const float* x, const float* wp, const float* wd, const void* dPp, float* dx, float* dwp, float* dwd, int B, int C, int H,
             void* workspace, int x_packed, cudaStream_t st);
}  // namespace tcb
namespace fz {   // hsm_fused.cu: sweeps over L with P = dw3x3(Wp x) kept in TMEM (never written to HBM)
size_t pack_bytes(int C);
int tiles_per_image(int H);
size_t merge_bytes(int B, int C, int H);
int merge(const float* part_m, const float* part_s, const float* part_hs, int B, int C, int H, void* workspace, float** m1, float** s1,
          float** hs1, cudaStream_t st);
int forward_hs(const float* x, const float* wp, const float* wd, float* part_m, float* part_s, float* part_hs, int B, int C, int H,
               void* workspace, cudaStream_t st);
int dp(const float* x, const float* dy, const float* wp, const float* wd, const float* stats, const float* dhs, const float* ho,
       const float* r, void* dPp, float* dx, int B, int C, int H, void* workspace, cudaStream_t st);
}  // namespace fz

constexpr int N = 64;         // states
constexpr int N3 = 192;       // projected channels
constexpr int TH = 8, TW = 32, HW_ = TW + 2, HH_ = TH + 2, NHALO = HW_ * HH_;  // spatial tile + 1-pixel halo (340)
constexpr int SUB = 256;      // positions per sub-tile in the over-L sweeps
constexpr int SUBS_PER_CTA = 2;
constexpr int GT = 256;      // threads of the per-batch (C,64) gate kernels
constexpr int GNS = 4;       // ... which split the 64 state columns over GNS CTAs per batch element (the gate is column-wise
constexpr int GCW = 64 / GNS; //     independent; sums over columns become per-CTA partials): 4x the CTAs on a latency-bound kernel

struct Dims {
  int B, C, L, H;
  int tiles_x, tiles_y;  // spatial tiling (TH x TW)
  int T;                 // CTAs per batch element in the over-L sweeps
};

static Dims make_dims(const kmu_hsmssd_desc& s) {
  Dims d;
  d.B = s.B; d.C = s.C; d.L = s.L; d.H = s.H;
  d.tiles_x = cdiv(s.H, TW);
  d.tiles_y = cdiv(s.H, TH);
  d.T = cdiv(s.L, SUB * SUBS_PER_CTA);
  return d;
}

// ================================================================================================ forward
// ---- P = dw3x3(Wp x).  grid (tiles, B), 256 threads, tile = 8 rows x 32 columns.  The x halo tile is staged once in
//      shared memory (10 rows x 36 columns: halo columns -2..33 so that every group of 4 positions is 16-byte aligned); the CTA
//      then walks the 192 projected channels in passes of PCH:
//        projection  task = (4 consecutive halo positions, 4 channels): per input channel one LDS.128 of x and one of
//                    weights feed 16 FMAs (register tile; shared-memory bandwidth was the limiter of the 1x16 version)
//        depthwise   thread = (tile column, channel): 3x3 window slides down the 8 rows in registers, taps in registers,
//                    3 LDS + 9 FMA per output, one coalesced 128-byte store per warp and row.
constexpr int PCH = 16;
constexpr int HPX = 36, HPR = 10, HPN = HPX * HPR;  // padded halo tile: 360 positions
template <int C>
__global__ void __launch_bounds__(256) hsm_proj_dw_kernel(const float* __restrict__ x, const float* __restrict__ wp,
                                                          const float* __restrict__ wd, float* __restrict__ P, Dims d) {
  extern __shared__ __align__(16) float smem[];
  float* x_s = smem;                    // [C][HPN]
  float* q_s = x_s + C * HPN;           // [PCH][HPN]
  float* wp_s = q_s + PCH * HPN;        // [C][PCH]
  float* wd_s = wp_s + C * PCH;         // [PCH][12]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ty0 = (blockIdx.x / d.tiles_x) * TH, tx0 = (blockIdx.x % d.tiles_x) * TW;
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * C * d.L;
  for (int i = tid; i < C * HPN; i += 256) {
    int c = i / HPN, pos = i - c * HPN;
    int hy = pos / HPX, hx = pos - hy * HPX;
    int gy = ty0 + hy - 1, gx = tx0 + hx - 2;
    x_s[i] = (gy >= 0 && gy < d.H && gx >= 0 && gx < d.H) ? __ldg(xb + (size_t)c * d.L + (size_t)gy * d.H + gx) : 0.f;
  }
  const int gx = tx0 + lane;
  for (int n0 = 0; n0 < N3; n0 += PCH) {
    __syncthreads();
    for (int i = tid; i < C * PCH; i += 256) {
      int c = i / PCH, nn = i - c * PCH;
      wp_s[i] = wp[(size_t)(n0 + nn) * C + c];
    }
    for (int i = tid; i < PCH * 9; i += 256) wd_s[(i / 9) * 12 + (i % 9)] = wd[(size_t)n0 * 9 + i];
    __syncthreads();
    // projection on the halo: 90 position quads x 4 channel groups = 360 tasks
    for (int task = tid; task < (HPN / 4) * (PCH / 4); task += 256) {
      const int quad = task % (HPN / 4), g = task / (HPN / 4);
      float q[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) q[k][e] = 0.f;
#pragma unroll 8
      for (int c = 0; c < C; ++c) {
        const float4 xv = *reinterpret_cast<const float4*>(x_s + c * HPN + 4 * quad);
        const float4 w = *reinterpret_cast<const float4*>(wp_s + c * PCH + 4 * g);
        q[0][0] = fmaf(w.x, xv.x, q[0][0]); q[0][1] = fmaf(w.x, xv.y, q[0][1]); q[0][2] = fmaf(w.x, xv.z, q[0][2]); q[0][3] = fmaf(w.x, xv.w, q[0][3]);
        q[1][0] = fmaf(w.y, xv.x, q[1][0]); q[1][1] = fmaf(w.y, xv.y, q[1][1]); q[1][2] = fmaf(w.y, xv.z, q[1][2]); q[1][3] = fmaf(w.y, xv.w, q[1][3]);
        q[2][0] = fmaf(w.z, xv.x, q[2][0]); q[2][1] = fmaf(w.z, xv.y, q[2][1]); q[2][2] = fmaf(w.z, xv.z, q[2][2]); q[2][3] = fmaf(w.z, xv.w, q[2][3]);
        q[3][0] = fmaf(w.w, xv.x, q[3][0]); q[3][1] = fmaf(w.w, xv.y, q[3][1]); q[3][2] = fmaf(w.w, xv.z, q[3][2]); q[3][3] = fmaf(w.w, xv.w, q[3][3]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        *reinterpret_cast<float4*>(q_s + (4 * g + k) * HPN + 4 * quad) = make_float4(q[k][0], q[k][1], q[k][2], q[k][3]);
    }
    __syncthreads();
    // depthwise 3x3: warp handles channels 2*wid, 2*wid+1; lane = tile column (halo column lane + 2)
    if (gx < d.H) {
#pragma unroll
      for (int j = 0; j < PCH / 8; ++j) {
        const int nn = wid * (PCH / 8) + j;
        const float4 wa = *reinterpret_cast<const float4*>(wd_s + nn * 12);
        const float4 wb = *reinterpret_cast<const float4*>(wd_s + nn * 12 + 4);
        const float w8 = wd_s[nn * 12 + 8];
        const float* qp = q_s + nn * HPN + lane + 1;   // halo column of tap kx = 0
        float r0[3], r1[3], r2[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) { r0[cc] = qp[cc]; r1[cc] = qp[HPX + cc]; }
        float* pp = P + ((size_t)b * N3 + n0 + nn) * d.L + (size_t)ty0 * d.H + gx;
#pragma unroll
        for (int ly = 0; ly < TH; ++ly) {
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) r2[cc] = qp[(ly + 2) * HPX + cc];
          float sacc = r0[0] * wa.x + r0[1] * wa.y + r0[2] * wa.z;
          sacc += r1[0] * wa.w + r1[1] * wb.x + r1[2] * wb.y;
          sacc += r2[0] * wb.z + r2[1] * wb.w + r2[2] * w8;
          if (ty0 + ly < d.H) pp[(size_t)ly * d.H] = sacc;
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) { r0[cc] = r1[cc]; r1[cc] = r2[cc]; }
        }
      }
    }
  }
}

// ---- shared inner product over a sub-tile: acc[k] += sum_l u_s[c_k][l] * v_s[l][n]  (u_s [C][SUB], v_s [SUB][65])
template <int CPT>
__device__ __forceinline__ void contract_subtile(const float* __restrict__ u_s, const float* __restrict__ v_s, int n, int cg,
                                                 int C, float* acc) {
  for (int l4 = 0; l4 < SUB; l4 += 4) {
    float v0 = v_s[(l4 + 0) * 65 + n], v1 = v_s[(l4 + 1) * 65 + n], v2 = v_s[(l4 + 2) * 65 + n], v3 = v_s[(l4 + 3) * 65 + n];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      int c = cg + 4 * k;
      if (c < C) {
        float4 u = *reinterpret_cast<const float4*>(u_s + c * SUB + l4);
        acc[k] = fmaf(u.x, v0, fmaf(u.y, v1, fmaf(u.z, v2, fmaf(u.w, v3, acc[k]))));
      }
    }
  }
}

// ---- one sweep over L: online softmax of dt along L fused with hs_partial = x (e.Bm)^T.  grid (T, B), 256 threads.
template <int CPT>
__global__ void __launch_bounds__(256) hsm_softmax_hs_kernel(const float* __restrict__ x, const float* __restrict__ P,
                                                             float* __restrict__ part_m, float* __restrict__ part_s,
                                                             float* __restrict__ part_hs, Dims d) {
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;               // [SUB][65]
  float* x_s = w_s + SUB * 65;     // [C][SUB]
  float* m_run = x_s + d.C * SUB;  // [64]
  float* s_run = m_run + 64;
  float* scale_s = s_run + 64;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.y, t = blockIdx.x;
  const int n = tid & 63, cg = tid >> 6;
  if (tid < 64) { m_run[tid] = -INFINITY; s_run[tid] = 0.f; }
  float acc[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) acc[k] = 0.f;
  const float* xb = x + (size_t)b * d.C * d.L;
  const float* Pb = P + (size_t)b * N3 * d.L;
  for (int sub = 0; sub < SUBS_PER_CTA; ++sub) {
    const int l0 = (t * SUBS_PER_CTA + sub) * SUB;
    if (l0 >= d.L) break;
    __syncthreads();
    for (int i = tid; i < d.C * SUB; i += 256) {
      int c = i / SUB, li = i - c * SUB;
      x_s[i] = (l0 + li < d.L) ? __ldg(xb + (size_t)c * d.L + l0 + li) : 0.f;
    }
    for (int j = 0; j < 8; ++j) {
      const int ns = wid * 8 + j;
      float dtv[8], mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int l = l0 + lane + 32 * k;
        dtv[k] = (l < d.L) ? __ldg(Pb + (size_t)(2 * N + ns) * d.L + l) : -INFINITY;
        mx = fmaxf(mx, dtv[k]);
      }
      mx = warp_max(mx);
      const float m_old = m_run[ns];
      const float m_new = fmaxf(m_old, mx);
      float ssum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int l = l0 + lane + 32 * k;
        float e = (l < d.L) ? expf(dtv[k] - m_new) : 0.f;
        ssum += e;
        float bm = (l < d.L) ? __ldg(Pb + (size_t)ns * d.L + l) : 0.f;
        w_s[(lane + 32 * k) * 65 + ns] = e * bm;
      }
      ssum = warp_sum(ssum);
      __syncwarp();
      if (lane == 0) {
        float sc = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
        scale_s[ns] = sc;
        s_run[ns] = s_run[ns] * sc + ssum;
        m_run[ns] = m_new;
      }
    }
    __syncthreads();
    const float sc = scale_s[n];
#pragma unroll
    for (int k = 0; k < CPT; ++k) acc[k] *= sc;
    contract_subtile<CPT>(x_s, w_s, n, cg, d.C, acc);
  }
  __syncthreads();
  const size_t bt = (size_t)b * d.T + t;
  if (cg == 0) { part_m[bt * 64 + n] = m_run[n]; part_s[bt * 64 + n] = s_run[n]; }
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    int c = cg + 4 * k;
    if (c < d.C) part_hs[(bt * d.C + c) * 64 + n] = acc[k];
  }
}

// ---- combine the T partials of one batch element, then the (C,64)-sized gate.  grid B, GT threads; weights in smem.
__global__ void __launch_bounds__(GT) hsm_combine_gate_kernel(const float* __restrict__ part_m, const float* __restrict__ part_s,
                                                              const float* __restrict__ part_hs, const float* __restrict__ whz,
                                                              const float* __restrict__ wo, const float* __restrict__ Dp,
                                                              float* __restrict__ stats, float* __restrict__ hs_out,
                                                              float* __restrict__ hz_out, float* __restrict__ ho_out, Dims d) {
  extern __shared__ __align__(16) float smem[];
  const int C = d.C;
  constexpr int CG = GT / GCW;
  float* hs_s = smem;                  // [C][64]
  float* hz_s = hs_s + C * 64;         // [2C][64]
  float* v_s = hz_s + 2 * C * 64;      // [C][64]
  float* whz_s = v_s + C * 64;         // [2C][C]
  float* wo_s = whz_s + 2 * C * C;     // [C][C]
  float* sc_s = wo_s + C * C;          // [T][64] exp(m_t - m)
  const int tid = threadIdx.x, n = blockIdx.y * GCW + (tid & (GCW - 1)), cg = tid / GCW, b = blockIdx.x;
  for (int i = tid; i < 2 * C * C; i += GT) whz_s[i] = whz[i];
  for (int i = tid; i < C * C; i += GT) wo_s[i] = wo[i];
  const float* pm = part_m + (size_t)b * d.T * 64 + n;
  const float* ps = part_s + (size_t)b * d.T * 64 + n;
  float m = -INFINITY;
  for (int t = 0; t < d.T; ++t) m = fmaxf(m, pm[t * 64]);
  float ssum = 0.f;
  for (int t = 0; t < d.T; ++t) {
    float e = expf(pm[t * 64] - m);
    ssum += ps[t * 64] * e;
    if (cg == 0) sc_s[t * 64 + n] = e;
  }
  const float inv_s = 1.0f / ssum;
  if (cg == 0 && stats) { stats[(size_t)b * 128 + n] = m; stats[(size_t)b * 128 + 64 + n] = ssum; }
  __syncthreads();
  for (int c = cg; c < C; c += CG) {
    float a = 0.f;
    for (int t = 0; t < d.T; ++t) a = fmaf(part_hs[(((size_t)b * d.T + t) * C + c) * 64 + n], sc_s[t * 64 + n], a);
    a *= inv_s;
    hs_s[c * 64 + n] = a;
    if (hs_out) hs_out[((size_t)b * C + c) * 64 + n] = a;
  }
  __syncthreads();
  for (int dd = cg; dd < 2 * C; dd += CG) {
    float a = 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(whz_s[dd * C + c], hs_s[c * 64 + n], a);
    hz_s[dd * 64 + n] = a;
    if (hz_out) hz_out[((size_t)b * 2 * C + dd) * 64 + n] = a;
  }
  __syncthreads();
  const float Dv = Dp[0];
  for (int c = cg; c < C; c += CG) {
    float hh = hz_s[c * 64 + n], z = hz_s[(C + c) * 64 + n];
    v_s[c * 64 + n] = hh * siluf_(z) + hh * Dv;
  }
  __syncthreads();
  for (int dd = cg; dd < C; dd += CG) {
    float a = 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(wo_s[dd * C + c], v_s[c * 64 + n], a);
    ho_out[((size_t)b * C + dd) * 64 + n] = a;
  }
}

// ---- y[b,c,l] = sum_n ho[b,c,n] Cm[b,n,l].  thread = position, 64 states in registers.  grid (ceil(L/256), B).
__global__ void __launch_bounds__(256) hsm_out_kernel(const float* __restrict__ ho, const float* __restrict__ P,
                                                      float* __restrict__ y, Dims d) {
  extern __shared__ __align__(16) float smem[];
  float* ho_s = smem;  // [C][64]
  const int tid = threadIdx.x, b = blockIdx.y;
  for (int i = tid; i < d.C * 64; i += 256) ho_s[i] = ho[(size_t)b * d.C * 64 + i];
  __syncthreads();
  const int l = blockIdx.x * 256 + tid;
  if (l >= d.L) return;
  float cm[64];
  const float* pc = P + ((size_t)b * N3 + N) * d.L + l;
#pragma unroll
  for (int nn = 0; nn < 64; ++nn) cm[nn] = __ldg(pc + (size_t)nn * d.L);
  float* yb = y + (size_t)b * d.C * d.L + l;
  for (int c = 0; c < d.C; ++c) {
    const float4* h4 = reinterpret_cast<const float4*>(ho_s + c * 64);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float4 h = h4[i];
      s0 = fmaf(h.x, cm[4 * i + 0], s0);
      s1 = fmaf(h.y, cm[4 * i + 1], s1);
      s2 = fmaf(h.z, cm[4 * i + 2], s2);
      s3 = fmaf(h.w, cm[4 * i + 3], s3);
    }
    yb[(size_t)c * d.L] = (s0 + s1) + (s2 + s3);
  }
}

// ================================================================================================ backward
// ---- dho partials: part[b,t,c,n] = sum_{l in CTA range} dy[b,c,l] Cm[b,n,l].  grid (T, B).
template <int CPT>
__global__ void __launch_bounds__(256) hsm_contract_kernel(const float* __restrict__ u, const float* __restrict__ P,
                                                           float* __restrict__ part, Dims d) {
  extern __shared__ __align__(16) float smem[];
  float* v_s = smem;            // [SUB][65]
  float* u_s = v_s + SUB * 65;  // [C][SUB]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.y, t = blockIdx.x, n = tid & 63, cg = tid >> 6;
  float acc[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) acc[k] = 0.f;
  const float* ub = u + (size_t)b * d.C * d.L;
  const float* Pc = P + ((size_t)b * N3 + N) * d.L;
  for (int sub = 0; sub < SUBS_PER_CTA; ++sub) {
    const int l0 = (t * SUBS_PER_CTA + sub) * SUB;
    if (l0 >= d.L) break;
    __syncthreads();
    for (int i = tid; i < d.C * SUB; i += 256) {
      int c = i / SUB, li = i - c * SUB;
      u_s[i] = (l0 + li < d.L) ? __ldg(ub + (size_t)c * d.L + l0 + li) : 0.f;
    }
    for (int j = 0; j < 8; ++j) {
      const int ns = wid * 8 + j;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int l = l0 + lane + 32 * k;
        v_s[(lane + 32 * k) * 65 + ns] = (l < d.L) ? __ldg(Pc + (size_t)ns * d.L + l) : 0.f;
      }
    }
    __syncthreads();
    contract_subtile<CPT>(u_s, v_s, n, cg, d.C, acc);
  }
  const size_t bt = (size_t)b * d.T + t;
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    int c = cg + 4 * k;
    if (c < d.C) part[(bt * d.C + c) * 64 + n] = acc[k];
  }
}

// ---- gate backward on (C,64) tiles.  grid B, GT threads; weights in smem.  Per-batch weight-gradient partials ->
//      wpart[b][...], layout of wpart[b]: dWo (C*C) | dWhz (2C*C) | dD (1)
__global__ void __launch_bounds__(GT) hsm_gate_bwd_kernel(const float* __restrict__ part_dho, const float* __restrict__ dh,
                                                          const float* __restrict__ hs,
                                                          const float* __restrict__ hz, const float* __restrict__ whz,
                                                          const float* __restrict__ wo, const float* __restrict__ Dp,
                                                          float* __restrict__ dhs_out, float* __restrict__ r_out,
                                                          float* __restrict__ wpart, Dims d) {
  extern __shared__ __align__(16) float smem[];
  const int C = d.C;
  constexpr int CG = GT / GCW;
  float* dho_s = smem;               // [C][64]
  float* v_s = dho_s + C * 64;       // [C][64]
  float* hs_s = v_s + C * 64;        // [C][64]
  float* dhz_s = hs_s + C * 64;      // [2C][64]
  float* whz_s = dhz_s + 2 * C * 64; // [2C][C]
  float* wo_s = whz_s + 2 * C * C;   // [C][C]
  float* red_s = wo_s + C * C;       // [GT]
  const int tid = threadIdx.x, nl = tid & (GCW - 1), n0 = blockIdx.y * GCW, n = n0 + nl, cg = tid / GCW, b = blockIdx.x;
  const float Dv = Dp[0];
  for (int i = tid; i < 2 * C * C; i += GT) whz_s[i] = whz[i];
  for (int i = tid; i < C * C; i += GT) wo_s[i] = wo[i];
  for (int c = cg; c < C; c += CG) {
    float a = 0.f;
    for (int t = 0; t < d.T; ++t) a += part_dho[(((size_t)b * d.T + t) * C + c) * 64 + n];
    if (dh) a += dh[((size_t)b * C + c) * 64 + n];
    dho_s[c * 64 + n] = a;
    hs_s[c * 64 + n] = hs[((size_t)b * C + c) * 64 + n];
    float hh = hz[((size_t)b * 2 * C + c) * 64 + n], z = hz[((size_t)b * 2 * C + C + c) * 64 + n];
    v_s[c * 64 + n] = hh * (siluf_(z) + Dv);
  }
  __syncthreads();
  float dD_local = 0.f;
  for (int c = cg; c < C; c += CG) {
    float dv = 0.f;
    for (int dd = 0; dd < C; ++dd) dv = fmaf(wo_s[dd * C + c], dho_s[dd * 64 + n], dv);
    float hh = hz[((size_t)b * 2 * C + c) * 64 + n], z = hz[((size_t)b * 2 * C + C + c) * 64 + n];
    dhz_s[c * 64 + n] = dv * (siluf_(z) + Dv);
    dhz_s[(C + c) * 64 + n] = dv * hh * silu_gradf_(z);
    dD_local = fmaf(dv, hh, dD_local);
  }
  red_s[tid] = dD_local;
  __syncthreads();
  float rn = 0.f;
  for (int c = cg; c < C; c += CG) {
    float a = 0.f;
    for (int dd = 0; dd < 2 * C; ++dd) a = fmaf(whz_s[dd * C + c], dhz_s[dd * 64 + n], a);
    dhs_out[((size_t)b * C + c) * 64 + n] = a;
    rn = fmaf(a, hs_s[c * 64 + n], rn);
  }
  // r[n] = sum_c dhs[c][n] hs[c][n]: the CG channel groups hold partial sums for the same n
  float dD_tot = 0.f;
  if (tid == 0) {
    for (int i = 0; i < GT; ++i) dD_tot += red_s[i];
  }
  __syncthreads();
  red_s[tid] = rn;
  __syncthreads();
  if (cg == 0) {
    float a = 0.f;
#pragma unroll
    for (int g = 0; g < CG; ++g) a += red_s[g * GCW + nl];
    r_out[(size_t)b * 64 + n] = a;
  }
  float* wb = wpart + ((size_t)b * GNS + blockIdx.y) * (3 * C * C + 1);   // partial over this CTA's GCW columns
  if (tid == 0) wb[3 * C * C] = dD_tot;
  // dWo[dd][c] = sum_n dho[dd][n] v[c][n] ; dWhz[dd][c] = sum_n dhz[dd][n] hs[c][n]
  for (int i = tid; i < 3 * C * C; i += GT) {
    const float *ar, *br;
    if (i < C * C) { ar = dho_s + (i / C) * 64; br = v_s + (i % C) * 64; }
    else { int k = i - C * C; ar = dhz_s + (k / C) * 64; br = hs_s + (k % C) * 64; }
    float a = 0.f;
    for (int k = 0; k < GCW; ++k) {
      int nn = n0 + ((k + tid) & (GCW - 1));  // rotate the start so that neighbouring lanes hit different banks
      a = fmaf(ar[nn], br[nn], a);
    }
    wb[i] = a;
  }
}

// ---- dP = [dBm; dCm; ddt] and the direct part of dx.  thread = position.  grid (ceil(L/256), B).
//      PACK: dP is written as bf16 K-group planes dPp[b][n / 8][l][n % 8] (16 bytes per thread and group, fully coalesced) --
//      the operand image of the tcgen05 backward (hsm_tc_bwd.cu) -- instead of fp32 (B,192,L).
__device__ __forceinline__ uint32_t dp_pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <int C, bool PACK>
__global__ void __launch_bounds__(256, 2) hsm_dp_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                     const float* __restrict__ P, const float* __restrict__ stats,
                                                     const float* __restrict__ dhs, const float* __restrict__ ho,
                                                     const float* __restrict__ r, float* __restrict__ dP, float* __restrict__ dx,
                                                     Dims d) {
  extern __shared__ __align__(16) float smem[];
  float* dhs_s = smem;            // [C][64]
  float* ho_s = dhs_s + C * 64;   // [C][64]
  float* m_s = ho_s + C * 64;     // [64] max, [64] 1/sum, [64] r
  const int tid = threadIdx.x, b = blockIdx.y;
  for (int i = tid; i < C * 64; i += 256) {
    dhs_s[i] = dhs[(size_t)b * C * 64 + i];
    ho_s[i] = ho[(size_t)b * C * 64 + i];
  }
  if (tid < 64) {
    m_s[tid] = stats[(size_t)b * 128 + tid];
    m_s[64 + tid] = 1.0f / stats[(size_t)b * 128 + 64 + tid];
    m_s[128 + tid] = r[(size_t)b * 64 + tid];
  }
  __syncthreads();
  const int l = blockIdx.x * 256 + tid;
  if (l >= d.L) return;
  const float* xb = x + (size_t)b * C * d.L + l;
  const float* dyb = dy + (size_t)b * C * d.L + l;
  const float* Pb = P + (size_t)b * N3 * d.L + l;
  float* dPb = dP + (size_t)b * N3 * d.L + l;
  uint4* dPq = reinterpret_cast<uint4*>(dP) + (size_t)b * 24 * d.L + l;  // PACK: plane g at dPq[g * L]
  float t[64];
#pragma unroll
  for (int nn = 0; nn < 64; ++nn) t[nn] = 0.f;
  // dG[n] = sum_c dhs[c][n] x[c]
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    float xv = __ldg(xb + (size_t)c * d.L);
    const float4* g4 = reinterpret_cast<const float4*>(dhs_s + c * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float4 g = g4[i];
      t[4 * i + 0] = fmaf(g.x, xv, t[4 * i + 0]);
      t[4 * i + 1] = fmaf(g.y, xv, t[4 * i + 1]);
      t[4 * i + 2] = fmaf(g.z, xv, t[4 * i + 2]);
      t[4 * i + 3] = fmaf(g.w, xv, t[4 * i + 3]);
    }
  }
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    float vb[8], vt[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int nn = g * 8 + e;
      float dtv = __ldg(Pb + (size_t)(2 * N + nn) * d.L);
      float bm = __ldg(Pb + (size_t)nn * d.L);
      float a = expf(dtv - m_s[nn]) * m_s[64 + nn];
      float dG = t[nn];
      vb[e] = dG * a;                             // dBm
      vt[e] = a * (dG * bm - m_s[128 + nn]);      // ddt
      t[nn] = a * bm;
    }
    if (PACK) {
      dPq[(size_t)g * d.L] = make_uint4(dp_pack2(vb[0], vb[1]), dp_pack2(vb[2], vb[3]), dp_pack2(vb[4], vb[5]), dp_pack2(vb[6], vb[7]));
      dPq[(size_t)(16 + g) * d.L] = make_uint4(dp_pack2(vt[0], vt[1]), dp_pack2(vt[2], vt[3]), dp_pack2(vt[4], vt[5]), dp_pack2(vt[6], vt[7]));
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dPb[(size_t)(g * 8 + e) * d.L] = vb[e];
        dPb[(size_t)(2 * N + g * 8 + e) * d.L] = vt[e];
      }
    }
  }
  // direct dx[c] = sum_n dhs[c][n] A[n] Bm[n]
  float* dxb = dx + (size_t)b * C * d.L + l;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float4* g4 = reinterpret_cast<const float4*>(dhs_s + c * 64);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float4 g = g4[i];
      s0 = fmaf(g.x, t[4 * i + 0], s0);
      s1 = fmaf(g.y, t[4 * i + 1], s1);
      s2 = fmaf(g.z, t[4 * i + 2], s2);
      s3 = fmaf(g.w, t[4 * i + 3], s3);
    }
    dxb[(size_t)c * d.L] = (s0 + s1) + (s2 + s3);
  }
  // dCm[n] = sum_c ho[c][n] dy[c]
#pragma unroll
  for (int nn = 0; nn < 64; ++nn) t[nn] = 0.f;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    float gv = __ldg(dyb + (size_t)c * d.L);
    const float4* h4 = reinterpret_cast<const float4*>(ho_s + c * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float4 h = h4[i];
      t[4 * i + 0] = fmaf(h.x, gv, t[4 * i + 0]);
      t[4 * i + 1] = fmaf(h.y, gv, t[4 * i + 1]);
      t[4 * i + 2] = fmaf(h.z, gv, t[4 * i + 2]);
      t[4 * i + 3] = fmaf(h.w, gv, t[4 * i + 3]);
    }
  }
  if (PACK) {
#pragma unroll
    for (int g = 0; g < 8; ++g)
      dPq[(size_t)(8 + g) * d.L] = make_uint4(dp_pack2(t[g * 8], t[g * 8 + 1]), dp_pack2(t[g * 8 + 2], t[g * 8 + 3]),
                                              dp_pack2(t[g * 8 + 4], t[g * 8 + 5]), dp_pack2(t[g * 8 + 6], t[g * 8 + 7]));
  } else {
#pragma unroll
    for (int nn = 0; nn < 64; ++nn) dPb[(size_t)(N + nn) * d.L] = t[nn];
  }
}

// ---- projection / depthwise backward on a spatial tile.  grid (tiles, B), 256 threads, tile = 8 rows x 32 columns.
//      dQ = dw3x3^T(dP) ; dx += Wp^T dQ ; per-CTA partials of dWp[n][c] = sum dQ x and dWd[n][tap] = sum Q(q) dP(q - tap)
//      (Q only at the CTA's own pixels: every pixel q belongs to exactly one tile and dP is zero outside the image).
//      partial layout per CTA: dWp (192*C) | dWd (192*9).  The x tile is staged once; the 192 channels go in passes of BN:
//        Q        task = (4 consecutive pixels, 4 channels): LDS.128 x + LDS.128 weights -> 16 FMA
//        stencil  thread = (tile column, channel): one 3x3 window of dP slides down the rows in registers and feeds BOTH
//                 dQ (taps in registers) and the dWd partial (9 register accumulators, one warp reduction per channel)
//        dx       task = (4 consecutive pixels, C/4 input channels), accumulators live in registers across all passes
//        dWp      4x4 register blocks over interleaved pixel slices (conflict-free LDS.128), shuffle-reduced over slices
constexpr int BN = 16;   // projected channels per pass
constexpr int QP = 260;  // row pitch of the [channel][256 pixel] tiles
template <int C>
__global__ void __launch_bounds__(256) hsm_proj_dw_bwd_kernel(const float* __restrict__ x, const float* __restrict__ wp,
                                                              const float* __restrict__ wd, const float* __restrict__ dP,
                                                              float* __restrict__ dx, float* __restrict__ partial, Dims d) {
  constexpr int NSL = 256 / C;  // pixel slices of the dWp contraction (C blocks of 4x4 outputs x NSL slices = 256 threads)
  constexpr int CQ = C / 4;     // input channels per dx task
  extern __shared__ __align__(16) float smem[];
  float* dp_s = smem;                 // [BN][NHALO]
  float* q_s = dp_s + BN * NHALO;     // [BN][QP]
  float* dq_s = q_s + BN * QP;        // [BN][QP]
  float* x_s = dq_s + BN * QP;        // [C][QP]
  float* wp_s = x_s + C * QP;         // [C][BN]  (c-major for the Q recompute)
  float* wpt_s = wp_s + C * BN;       // [BN][C]  (n-major for dx)
  float* wd_s = wpt_s + BN * C;       // [BN][12]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ty0 = (blockIdx.x / d.tiles_x) * TH, tx0 = (blockIdx.x % d.tiles_x) * TW;
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * C * d.L;
  {
    const int gy = ty0 + (tid >> 5), gx = tx0 + (tid & 31);
    const bool inside = gy < d.H && gx < d.H;
#pragma unroll 4
    for (int c = 0; c < C; ++c) x_s[c * QP + tid] = inside ? __ldg(xb + (size_t)c * d.L + (size_t)gy * d.H + gx) : 0.f;
  }
  float* pb = partial + ((size_t)b * gridDim.x + blockIdx.x) * (size_t)(N3 * C + N3 * 9);
  const int quad = tid & 63, grp = tid >> 6;          // Q / dx task: pixels 4*quad .. 4*quad+3 (one tile row), group 0..3
  float dxa[CQ][4];
#pragma unroll
  for (int c = 0; c < CQ; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) dxa[c][e] = 0.f;
  const bool col_in = tx0 + lane < d.H;

  for (int n0 = 0; n0 < N3; n0 += BN) {
    __syncthreads();
    for (int i = tid; i < C * BN; i += 256) {
      int c = i / BN, nn = i - c * BN;
      float w = wp[(size_t)(n0 + nn) * C + c];
      wp_s[i] = w;
      wpt_s[nn * C + c] = w;
    }
    for (int i = tid; i < BN * 9; i += 256) wd_s[(i / 9) * 12 + (i % 9)] = wd[(size_t)n0 * 9 + i];
    // dP tile with halo (zero outside the image).  One warp per halo row, lanes along the row: no per-element div/mod, and
    // the 20 row loads of a warp are issued back to back before any of them is stored.
    {
      constexpr int ROWS = BN * HH_ / 8;   // 20 halo rows per warp
      float v[ROWS];
      const int xx = tx0 + lane - 1;
      const bool colok = xx >= 0 && xx < d.H;
      const float* src0 = dP + ((size_t)b * N3 + n0) * d.L + xx;
#pragma unroll
      for (int k = 0; k < ROWS; ++k) {
        const int row = wid + 8 * k;
        const int nn = row / HH_, hy = row - nn * HH_;
        const int yy = ty0 + hy - 1;
        v[k] = (colok && yy >= 0 && yy < d.H) ? __ldg(src0 + (size_t)nn * d.L + (size_t)yy * d.H) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < ROWS; ++k) {
        const int row = wid + 8 * k;
        const int nn = row / HH_, hy = row - nn * HH_;
        dp_s[nn * NHALO + hy * HW_ + lane] = v[k];
      }
      // the two right-most halo columns (32, 33) of the 160 rows
      for (int i = tid; i < BN * HH_ * 2; i += 256) {
        const int row = i >> 1, hx = 32 + (i & 1);
        const int nn = row / HH_, hy = row - nn * HH_;
        const int yy = ty0 + hy - 1, x2 = tx0 + hx - 1;
        dp_s[nn * NHALO + hy * HW_ + hx] = (yy >= 0 && yy < d.H && x2 < d.H)
                                               ? __ldg(dP + ((size_t)b * N3 + n0 + nn) * d.L + (size_t)yy * d.H + x2) : 0.f;
      }
    }
    __syncthreads();
    {
      // Q at the CTA's own pixels (0 outside the image because x_s is 0 there): channels 4*grp .. 4*grp+3
      float q[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) q[k][e] = 0.f;
#pragma unroll 8
      for (int c = 0; c < C; ++c) {
        const float4 xv = *reinterpret_cast<const float4*>(x_s + c * QP + 4 * quad);
        const float4 w = *reinterpret_cast<const float4*>(wp_s + c * BN + 4 * grp);
        q[0][0] = fmaf(w.x, xv.x, q[0][0]); q[0][1] = fmaf(w.x, xv.y, q[0][1]); q[0][2] = fmaf(w.x, xv.z, q[0][2]); q[0][3] = fmaf(w.x, xv.w, q[0][3]);
        q[1][0] = fmaf(w.y, xv.x, q[1][0]); q[1][1] = fmaf(w.y, xv.y, q[1][1]); q[1][2] = fmaf(w.y, xv.z, q[1][2]); q[1][3] = fmaf(w.y, xv.w, q[1][3]);
        q[2][0] = fmaf(w.z, xv.x, q[2][0]); q[2][1] = fmaf(w.z, xv.y, q[2][1]); q[2][2] = fmaf(w.z, xv.z, q[2][2]); q[2][3] = fmaf(w.z, xv.w, q[2][3]);
        q[3][0] = fmaf(w.w, xv.x, q[3][0]); q[3][1] = fmaf(w.w, xv.y, q[3][1]); q[3][2] = fmaf(w.w, xv.z, q[3][2]); q[3][3] = fmaf(w.w, xv.w, q[3][3]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        *reinterpret_cast<float4*>(q_s + (4 * grp + k) * QP + 4 * quad) = make_float4(q[k][0], q[k][1], q[k][2], q[k][3]);
    }
    __syncthreads();
    // stencil: warp handles channels 2*wid, 2*wid+1; lane = tile column; the dP window (halo rows qy..qy+2, halo columns
    // lane..lane+2) gives dQ(qy, lane) = sum w[ky][kx] dP(qy+2-ky, lane+2-kx) and dWd[ky][kx] += Q(qy, lane) dP(same)
#pragma unroll
    for (int j = 0; j < BN / 8; ++j) {
      const int nn = wid * (BN / 8) + j;
      const float4 wa = *reinterpret_cast<const float4*>(wd_s + nn * 12);
      const float4 wb = *reinterpret_cast<const float4*>(wd_s + nn * 12 + 4);
      const float w8 = wd_s[nn * 12 + 8];
      float a[9];
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) a[tp] = 0.f;
      const float* gp = dp_s + nn * NHALO + lane;
      float r0[3], r1[3], r2[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) { r0[cc] = gp[cc]; r1[cc] = gp[HW_ + cc]; }
#pragma unroll
      for (int qy = 0; qy < TH; ++qy) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) r2[cc] = gp[(qy + 2) * HW_ + cc];
        const float qv = q_s[nn * QP + qy * 32 + lane];
        float dq = wa.x * r2[2] + wa.y * r2[1] + wa.z * r2[0];
        dq += wa.w * r1[2] + wb.x * r1[1] + wb.y * r1[0];
        dq += wb.z * r0[2] + wb.w * r0[1] + w8 * r0[0];
        dq_s[nn * QP + qy * 32 + lane] = (col_in && ty0 + qy < d.H) ? dq : 0.f;
        a[0] = fmaf(qv, r2[2], a[0]); a[1] = fmaf(qv, r2[1], a[1]); a[2] = fmaf(qv, r2[0], a[2]);
        a[3] = fmaf(qv, r1[2], a[3]); a[4] = fmaf(qv, r1[1], a[4]); a[5] = fmaf(qv, r1[0], a[5]);
        a[6] = fmaf(qv, r0[2], a[6]); a[7] = fmaf(qv, r0[1], a[7]); a[8] = fmaf(qv, r0[0], a[8]);
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) { r0[cc] = r1[cc]; r1[cc] = r2[cc]; }
      }
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        float sacc = warp_sum(a[tp]);
        if (lane == 0) pb[(size_t)N3 * C + (size_t)(n0 + nn) * 9 + tp] = sacc;
      }
    }
    __syncthreads();
    // dx accumulation: pixels 4*quad.., input channels grp*CQ .. grp*CQ + CQ - 1
#pragma unroll 4
    for (int nn = 0; nn < BN; ++nn) {
      const float4 g = *reinterpret_cast<const float4*>(dq_s + nn * QP + 4 * quad);
      const float4* wr = reinterpret_cast<const float4*>(wpt_s + nn * C + grp * CQ);
#pragma unroll
      for (int c4 = 0; c4 < CQ / 4; ++c4) {
        const float4 w = wr[c4];
        dxa[4 * c4 + 0][0] = fmaf(w.x, g.x, dxa[4 * c4 + 0][0]); dxa[4 * c4 + 0][1] = fmaf(w.x, g.y, dxa[4 * c4 + 0][1]);
        dxa[4 * c4 + 0][2] = fmaf(w.x, g.z, dxa[4 * c4 + 0][2]); dxa[4 * c4 + 0][3] = fmaf(w.x, g.w, dxa[4 * c4 + 0][3]);
        dxa[4 * c4 + 1][0] = fmaf(w.y, g.x, dxa[4 * c4 + 1][0]); dxa[4 * c4 + 1][1] = fmaf(w.y, g.y, dxa[4 * c4 + 1][1]);
        dxa[4 * c4 + 1][2] = fmaf(w.y, g.z, dxa[4 * c4 + 1][2]); dxa[4 * c4 + 1][3] = fmaf(w.y, g.w, dxa[4 * c4 + 1][3]);
        dxa[4 * c4 + 2][0] = fmaf(w.z, g.x, dxa[4 * c4 + 2][0]); dxa[4 * c4 + 2][1] = fmaf(w.z, g.y, dxa[4 * c4 + 2][1]);
        dxa[4 * c4 + 2][2] = fmaf(w.z, g.z, dxa[4 * c4 + 2][2]); dxa[4 * c4 + 2][3] = fmaf(w.z, g.w, dxa[4 * c4 + 2][3]);
        dxa[4 * c4 + 3][0] = fmaf(w.w, g.x, dxa[4 * c4 + 3][0]); dxa[4 * c4 + 3][1] = fmaf(w.w, g.y, dxa[4 * c4 + 3][1]);
        dxa[4 * c4 + 3][2] = fmaf(w.w, g.z, dxa[4 * c4 + 3][2]); dxa[4 * c4 + 3][3] = fmaf(w.w, g.w, dxa[4 * c4 + 3][3]);
      }
    }
    // dWp partial: thread = (4x4 block of (n,c) outputs, pixel slice); slice sl owns the float4 groups sl, sl+NSL, ... so the
    // lanes of a warp read consecutive 16-byte groups (no bank conflicts)
    {
      const int sl = tid % NSL, blk = tid / NSL;
      const int i4 = (blk & 3) * 4, j4 = (blk >> 2) * 4;
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[r][t] = 0.f;
#pragma unroll 2
      for (int k = 0; k < 64 / NSL; ++k) {
        const int px = (k * NSL + sl) * 4;
        float4 av[4], bv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          av[r] = *reinterpret_cast<const float4*>(dq_s + (i4 + r) * QP + px);
          bv[r] = *reinterpret_cast<const float4*>(x_s + (j4 + r) * QP + px);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int t = 0; t < 4; ++t)
            acc[r][t] = fmaf(av[r].x, bv[t].x, fmaf(av[r].y, bv[t].y, fmaf(av[r].z, bv[t].z, fmaf(av[r].w, bv[t].w, acc[r][t]))));
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float v = acc[r][t];
#pragma unroll
          for (int o = NSL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (sl == 0) pb[(size_t)(n0 + i4 + r) * C + j4 + t] = v;
        }
    }
  }
  // dx += the accumulated projection-path gradient (the direct part was written by hsm_dp_kernel)
  {
    const int gy = ty0 + (quad >> 3), gx0 = tx0 + (quad & 7) * 4;
    if (gy < d.H) {
#pragma unroll
      for (int c = 0; c < CQ; ++c) {
        float* dxp = dx + ((size_t)b * C + grp * CQ + c) * d.L + (size_t)gy * d.H + gx0;
        if ((d.H & 3) == 0 && gx0 + 3 < d.H) {
          float4 t = *reinterpret_cast<float4*>(dxp);
          t.x += dxa[c][0]; t.y += dxa[c][1]; t.z += dxa[c][2]; t.w += dxa[c][3];
          *reinterpret_cast<float4*>(dxp) = t;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (gx0 + e < d.H) dxp[e] += dxa[c][e];
        }
      }
    }
  }
}

// ---- fixed-order reductions of the weight-gradient partials: CTA = 32 outputs x 8 interleaved part slices (each thread
//      sums parts k = slice, slice+8, ...; the 8 slice sums are then added in order), so many loads are in flight per output
__global__ void __launch_bounds__(256) hsm_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, int per,
                                                               float* __restrict__ out0, int n0, float* __restrict__ out1, int n1,
                                                               float* __restrict__ out2, int n2) {
  __shared__ float red[8][33];
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + o;
  float s = 0.f;
  if (idx < n0 + n1 + n2)
    for (int k = sl; k < nparts; k += 8) s += partial[(size_t)k * per + idx];
  red[sl][o] = s;
  __syncthreads();
  if (sl == 0 && idx < n0 + n1 + n2) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][o];
    if (idx < n0) out0[idx] = t;
    else if (idx < n0 + n1) out1[idx - n0] = t;
    else out2[idx - n0 - n1] = t;
  }
}

__global__ void fill_kernel(float* __restrict__ p, int n, float v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ================================================================================================ LayerNorm1D
__global__ void __launch_bounds__(256) ln1d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ y,
                                                       float* __restrict__ rstd_out, int B, int C, int L, float eps) {
  long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= (long long)B * L) return;
  int b = (int)(p / L), l = (int)(p - (long long)b * L);
  const float* xp = x + (size_t)b * C * L + l;
  float mean = 0.f;
  for (int c = 0; c < C; ++c) mean += __ldg(xp + (size_t)c * L);
  mean /= (float)C;
  float var = 0.f;
  for (int c = 0; c < C; ++c) {
    float dlt = __ldg(xp + (size_t)c * L) - mean;
    var = fmaf(dlt, dlt, var);
  }
  var /= (float)C;
  float sd = sqrtf(var + eps);
  float* yp = y + (size_t)b * C * L + l;
  for (int c = 0; c < C; ++c) yp[(size_t)c * L] = (__ldg(xp + (size_t)c * L) - mean) / sd * w[c] + bias[c];
  if (rstd_out) rstd_out[p] = 1.0f / sd;
}

__global__ void __launch_bounds__(256) ln1d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ dy, float* __restrict__ dx,
                                                       float* __restrict__ part, int B, int C, int L, float eps) {
  extern __shared__ float red[];  // [8 warps][2 C]: per-warp sums, combined in warp order (no atomics: bit-reproducible)
  long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  const bool valid = p < (long long)B * L;
  int b = valid ? (int)(p / L) : 0, l = valid ? (int)(p - (long long)b * L) : 0;
  const float* xp = x + (size_t)b * C * L + l;
  const float* gp = dy + (size_t)b * C * L + l;
  float mean = 0.f, var = 0.f, rstd = 0.f, m1 = 0.f, m2 = 0.f;
  if (valid) {
    for (int c = 0; c < C; ++c) mean += __ldg(xp + (size_t)c * L);
    mean /= (float)C;
    for (int c = 0; c < C; ++c) {
      float dlt = __ldg(xp + (size_t)c * L) - mean;
      var = fmaf(dlt, dlt, var);
    }
    var /= (float)C;
    rstd = 1.0f / sqrtf(var + eps);
    for (int c = 0; c < C; ++c) {
      float xh = (__ldg(xp + (size_t)c * L) - mean) * rstd;
      float g = __ldg(gp + (size_t)c * L) * w[c];
      m1 += g;
      m2 = fmaf(g, xh, m2);
    }
    m1 /= (float)C;
    m2 /= (float)C;
  }
  const int lane = threadIdx.x & 31;
  for (int c = 0; c < C; ++c) {
    float xh = 0.f, gy = 0.f;
    if (valid) {
      xh = (__ldg(xp + (size_t)c * L) - mean) * rstd;
      gy = __ldg(gp + (size_t)c * L);
      dx[(size_t)b * C * L + (size_t)c * L + l] = rstd * (gy * w[c] - m1 - xh * m2);
    }
    float sw = warp_sum(gy * xh), sb = warp_sum(gy);
    if (lane == 0) { red[(threadIdx.x >> 5) * 2 * C + c] = sw; red[(threadIdx.x >> 5) * 2 * C + C + c] = sb; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float a = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) a += red[w8 * 2 * C + i];
    part[(size_t)blockIdx.x * 2 * C + i] = a;
  }
}

// dw[c] = sum over blocks of part[blk][c], db[c] = ... part[blk][C + c]: 32 outputs x 8 interleaved slices of the block list per CTA
__global__ void __launch_bounds__(256) ln1d_bwd_reduce_kernel(const float* __restrict__ part, int nblk, int C, float* __restrict__ dw,
                                                              float* __restrict__ db) {
  __shared__ float red[8][33];
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + o;
  float s = 0.f;
  if (idx < 2 * C)
    for (int k = sl; k < nblk; k += 8) s += part[(size_t)k * 2 * C + idx];
  red[sl][o] = s;
  __syncthreads();
  if (sl == 0 && idx < 2 * C) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][o];
    if (idx < C) dw[idx] = t; else db[idx - C] = t;
  }
}

// Register-resident forms for the widths KM-UNet uses (C = 16, 32, 64): thread = position, the C values of x (and dy) are loaded ONCE
// (coalesced across the warp) and every later pass runs on registers -- the generic kernels above re-read x three times and dy twice.
template <int C>
__global__ void __launch_bounds__(256) ln1d_fwd_reg_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ y,
                                                           float* __restrict__ rstd_out, int B, int L, float eps) {
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= (long long)B * L) return;
  const int b = (int)(p / L), l = (int)(p - (long long)b * L);
  const float* xp = x + (size_t)b * C * L + l;
  float v[C];
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = __ldg(xp + (size_t)c * L);
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) mean += v[c];
  mean /= (float)C;
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float dlt = v[c] - mean;
    var = fmaf(dlt, dlt, var);
  }
  var /= (float)C;
  const float sd = sqrtf(var + eps);
  float* yp = y + (size_t)b * C * L + l;
#pragma unroll
  for (int c = 0; c < C; ++c) yp[(size_t)c * L] = (v[c] - mean) / sd * __ldg(w + c) + __ldg(bias + c);
  if (rstd_out) rstd_out[p] = 1.0f / sd;
}

template <int C>
__global__ void __launch_bounds__(256) ln1d_bwd_reg_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ dy, float* __restrict__ dx,
                                                           float* __restrict__ part, int B, int L, float eps) {
  __shared__ float red[8][2 * C];
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  const bool valid = p < (long long)B * L;
  const int b = valid ? (int)(p / L) : 0, l = valid ? (int)(p - (long long)b * L) : 0;
  const float* xp = x + (size_t)b * C * L + l;
  const float* gp = dy + (size_t)b * C * L + l;
  float xh[C], gy[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    xh[c] = valid ? __ldg(xp + (size_t)c * L) : 0.f;
    gy[c] = valid ? __ldg(gp + (size_t)c * L) : 0.f;
  }
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) mean += xh[c];
  mean /= (float)C;
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    xh[c] -= mean;
    var = fmaf(xh[c], xh[c], var);
  }
  var /= (float)C;
  const float rstd = 1.0f / sqrtf(var + eps);
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    xh[c] *= rstd;
    const float g = gy[c] * __ldg(w + c);
    m1 += g;
    m2 = fmaf(g, xh[c], m2);
  }
  m1 /= (float)C;
  m2 /= (float)C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    if (valid) dx[(size_t)b * C * L + (size_t)c * L + l] = rstd * (gy[c] * __ldg(w + c) - m1 - xh[c] * m2);
    const float sw = warp_sum(gy[c] * xh[c]), sb = warp_sum(gy[c]);
    if (lane == 0) { red[warp][c] = sw; red[warp][C + c] = sb; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float a = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) a += red[w8][i];
    part[(size_t)blockIdx.x * 2 * C + i] = a;
  }
}

static int check(const kmu_hsmssd_desc* d, const char* who) {
  KMU_REQUIRE(d != nullptr, KMU_ERR_BAD_ARG, "%s: null descriptor", who);
  KMU_REQUIRE(d->B > 0 && d->C > 0 && d->L > 0 && d->H > 0, KMU_ERR_BAD_ARG, "%s: non-positive shape", who);
  KMU_REQUIRE(d->H * d->H == d->L, KMU_ERR_BAD_ARG, "%s: L=%d is not H*H with H=%d (the reference requires square maps)", who,
              d->L, d->H);
  KMU_REQUIRE(d->N == N, KMU_ERR_UNSUPPORTED, "%s: state_dim %d != 64", who, d->N);
  KMU_REQUIRE(d->C == 16 || d->C == 32 || d->C == 64, KMU_ERR_UNSUPPORTED, "%s: C=%d not in {16,32,64} (the widths KM-UNet uses)", who,
              d->C);
  KMU_REQUIRE(d->B <= 65535, KMU_ERR_UNSUPPORTED, "%s: B=%d > 65535", who, d->B);
  return KMU_OK;
}

// KMU_HSM_FUSED=0 selects the round-1 tensor-core path (P written to HBM) for A/B measurements
static bool fused_sweeps() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KMU_HSM_FUSED");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

struct FwdWs { size_t part_m, part_s, part_hs, wpack, weff, merge, total; };
static FwdWs fwd_ws(const Dims& d) {
  FwdWs w;
  size_t o = 0;
  const size_t T = (size_t)std::max(d.T, fz::tiles_per_image(d.H));    // per-tile partials of the fused sweep
  w.wpack = o; o += std::max(tc::pack_bytes(d.C), fz::pack_bytes(d.C));
  w.weff = o; o += tc::weff_bytes(d.B, d.C);
  w.part_m = o; o += align_up((size_t)d.B * T * 64 * 4, 256);
  w.part_s = o; o += align_up((size_t)d.B * T * 64 * 4, 256);
  w.part_hs = o; o += align_up((size_t)d.B * T * d.C * 64 * 4, 256);
  w.merge = o; o += fz::merge_bytes(d.B, d.C, d.H);
  w.total = o;
  return w;
}
struct BwdWs { size_t part_dho, dhs, r, wpart, dP, tpart, wpk, total; };
static BwdWs bwd_ws(const Dims& d, int precision) {
  BwdWs w;
  size_t o = 0;
  w.part_dho = o; o += align_up((size_t)d.B * d.T * d.C * 64 * 4, 256);
  w.dhs = o; o += align_up((size_t)d.B * d.C * 64 * 4, 256);
  w.r = o; o += align_up((size_t)d.B * 64 * 4, 256);
  w.wpart = o; o += align_up((size_t)d.B * GNS * (3 * d.C * d.C + 1) * 4, 256);
  if (precision == KMU_PREC_BF16) {  // dP as bf16 planes; tpart = the tcgen05 backward's own workspace
    w.dP = o; o += tcb::dpp_bytes(d.B, d.L);
    w.tpart = o; o += tcb::workspace_bytes(d.B, d.C, d.H);
    w.wpk = o; o += fz::pack_bytes(d.C);
  } else {
    w.dP = o; o += align_up((size_t)d.B * N3 * d.L * 4, 256);
    w.tpart = o; o += align_up((size_t)d.B * d.tiles_x * d.tiles_y * (N3 * d.C + N3 * 9) * 4, 256);
  }
  w.total = o;
  return w;
}

template <typename K>
static void opt_in_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace hsm
}  // namespace kmu

using namespace kmu;
using namespace kmu::hsm;

extern "C" {

size_t kmu_hsmssd_fwd_workspace_bytes(const kmu_hsmssd_desc* dd) {
  if (check(dd, "hsmssd_fwd_workspace_bytes") != KMU_OK) return 0;
  return fwd_ws(make_dims(*dd)).total;
}
size_t kmu_hsmssd_bwd_workspace_bytes(const kmu_hsmssd_desc* dd) {
  if (check(dd, "hsmssd_bwd_workspace_bytes") != KMU_OK) return 0;
  return bwd_ws(make_dims(*dd), dd->precision).total;
}

int kmu_hsmssd_fwd(const kmu_hsmssd_fwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "hsmssd_fwd: null args");
  int st_ = check(&a->d, "hsmssd_fwd");
  if (st_ != KMU_OK) return st_;
  const bool fused = a->d.precision == KMU_PREC_BF16 && fused_sweeps();
  KMU_REQUIRE(a->x && a->w_bcdt && a->w_dw && a->w_hz && a->w_out && a->D && a->y && a->h && (a->P || fused), KMU_ERR_BAD_ARG,
              "hsmssd_fwd: null tensor (the P scratch is required outside the fused bf16 path)");
  Dims d = make_dims(a->d);
  FwdWs w = fwd_ws(d);
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= w.total, KMU_ERR_WORKSPACE, "hsmssd_fwd: workspace %zu < %zu",
              a->workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)a->workspace;
  float* part_m = (float*)(ws + w.part_m);
  float* part_s = (float*)(ws + w.part_s);
  float* part_hs = (float*)(ws + w.part_hs);
  Dims dc = d;                    // combine: number of partials per batch element
  if (fused) {
    // projection, tile-local softmax and the x (e . Bm)^T contraction in ONE kernel; P lives in TMEM only (hsm_fused.cu)
    int rc = fz::forward_hs(a->x, a->w_bcdt, a->w_dw, part_m, part_s, part_hs, d.B, d.C, d.H, ws + w.wpack, st);
    if (rc != KMU_OK) return rc;
    rc = fz::merge(part_m, part_s, part_hs, d.B, d.C, d.H, ws + w.merge, &part_m, &part_s, &part_hs, st);
    if (rc != KMU_OK) return rc;
    dc.T = 1;
  } else if (a->d.precision == KMU_PREC_BF16) {
    // Bm and dt slices only: y = ho Cm is folded into one more convolution of x (tc::out), d(ho) into a correlation (tcb::contract)
    int rc = tc::project(a->x, a->w_bcdt, a->w_dw, a->P, d.B, d.C, d.H, 2, ws + w.wpack, st);
    if (rc != KMU_OK) return rc;
  } else
  {
    size_t smem = ((size_t)d.C * HPN + (size_t)PCH * HPN + (size_t)d.C * PCH + PCH * 12) * 4;
    dim3 grid(d.tiles_x * d.tiles_y, d.B);
#define KMU_HSM_PROJ(CC)                                                                     \
  do {                                                                                       \
    opt_in_smem(hsm_proj_dw_kernel<CC>, smem);                                               \
    hsm_proj_dw_kernel<CC><<<grid, 256, smem, st>>>(a->x, a->w_bcdt, a->w_dw, a->P, d);      \
  } while (0)
    switch (d.C) {
      case 16: KMU_HSM_PROJ(16); break;
      case 32: KMU_HSM_PROJ(32); break;
      default: KMU_HSM_PROJ(64); break;
    }
#undef KMU_HSM_PROJ
    KMU_LAUNCH_CHECK("hsm_proj_dw");
  }
  if (!fused) {
    size_t smem = ((size_t)SUB * 65 + (size_t)d.C * SUB + 192) * 4;
    dim3 grid(d.T, d.B);
    if (d.C <= 16) {
      opt_in_smem(hsm_softmax_hs_kernel<4>, smem);
      hsm_softmax_hs_kernel<4><<<grid, 256, smem, st>>>(a->x, a->P, part_m, part_s, part_hs, d);
    } else if (d.C <= 32) {
      opt_in_smem(hsm_softmax_hs_kernel<8>, smem);
      hsm_softmax_hs_kernel<8><<<grid, 256, smem, st>>>(a->x, a->P, part_m, part_s, part_hs, d);
    } else {
      opt_in_smem(hsm_softmax_hs_kernel<16>, smem);
      hsm_softmax_hs_kernel<16><<<grid, 256, smem, st>>>(a->x, a->P, part_m, part_s, part_hs, d);
    }
    KMU_LAUNCH_CHECK("hsm_softmax_hs");
  }
  {
    size_t smem = ((size_t)4 * d.C * 64 + (size_t)3 * d.C * d.C + (size_t)dc.T * 64) * 4;
    KMU_REQUIRE(smem <= 220 * 1024, KMU_ERR_UNSUPPORTED, "hsmssd_fwd: L=%d too long for the per-batch combine (T=%d)", d.L, dc.T);
    opt_in_smem(hsm_combine_gate_kernel, smem);
    hsm_combine_gate_kernel<<<dim3(d.B, GNS), GT, smem, st>>>(part_m, part_s, part_hs, a->w_hz, a->w_out, a->D, a->stats, a->hs, a->hz,
                                                    a->h, dc);
    KMU_LAUNCH_CHECK("hsm_combine_gate");
  }
  if (a->d.precision == KMU_PREC_BF16) {
    int rc = tc::out(a->x, a->h, a->w_bcdt, a->w_dw, a->y, d.B, d.C, d.H, ws + w.weff, st);
    if (rc != KMU_OK) return rc;
  } else {
    size_t smem = (size_t)d.C * 64 * 4;
    hsm_out_kernel<<<dim3(cdiv(d.L, 256), d.B), 256, smem, st>>>(a->h, a->P, a->y, d);
    KMU_LAUNCH_CHECK("hsm_out");
  }
  return KMU_OK;
}

int kmu_hsmssd_bwd(const kmu_hsmssd_bwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "hsmssd_bwd: null args");
  int st_ = check(&a->d, "hsmssd_bwd");
  if (st_ != KMU_OK) return st_;
  const bool fused = a->d.precision == KMU_PREC_BF16 && fused_sweeps();
  KMU_REQUIRE(a->x && a->dy && a->w_bcdt && a->w_dw && a->w_hz && a->w_out && a->D && (a->P || fused) && a->stats && a->hs && a->hz && a->h,
              KMU_ERR_BAD_ARG, "hsmssd_bwd: null input tensor");
  KMU_REQUIRE(a->dx && a->d_w_bcdt && a->d_w_dw && a->d_w_hz && a->d_w_out && a->d_D, KMU_ERR_BAD_ARG,
              "hsmssd_bwd: null output tensor");
  Dims d = make_dims(a->d);
  const bool tcbwd = a->d.precision == KMU_PREC_BF16;
  BwdWs w = bwd_ws(d, a->d.precision);
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= w.total, KMU_ERR_WORKSPACE, "hsmssd_bwd: workspace %zu < %zu",
              a->workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)a->workspace;
  float* part_dho = (float*)(ws + w.part_dho);
  float* dhs = (float*)(ws + w.dhs);
  float* r = (float*)(ws + w.r);
  float* wpart = (float*)(ws + w.wpart);
  float* dP = (float*)(ws + w.dP);
  float* tpart = (float*)(ws + w.tpart);
  const int C = d.C;
  Dims dg = d;                    // gate backward: number of d(ho) partials per batch element
  if (tcbwd) {
    int rc = tcb::contract(a->x, a->dy, a->w_bcdt, a->w_dw, part_dho, d.B, C, d.H, tpart, st);
    if (rc != KMU_OK) return rc;
    dg.T = 1;
  } else {
    size_t smem = ((size_t)SUB * 65 + (size_t)C * SUB) * 4;
    dim3 grid(d.T, d.B);
    if (C <= 16) {
      opt_in_smem(hsm_contract_kernel<4>, smem);
      hsm_contract_kernel<4><<<grid, 256, smem, st>>>(a->dy, a->P, part_dho, d);
    } else if (C <= 32) {
      opt_in_smem(hsm_contract_kernel<8>, smem);
      hsm_contract_kernel<8><<<grid, 256, smem, st>>>(a->dy, a->P, part_dho, d);
    } else {
      opt_in_smem(hsm_contract_kernel<16>, smem);
      hsm_contract_kernel<16><<<grid, 256, smem, st>>>(a->dy, a->P, part_dho, d);
    }
    KMU_LAUNCH_CHECK("hsm_contract");
  }
  {
    size_t smem = ((size_t)5 * C * 64 + (size_t)3 * C * C + GT) * 4;
    opt_in_smem(hsm_gate_bwd_kernel, smem);
    hsm_gate_bwd_kernel<<<dim3(d.B, GNS), GT, smem, st>>>(part_dho, a->dh, a->hs, a->hz, a->w_hz, a->w_out, a->D, dhs, r, wpart, dg);
    KMU_LAUNCH_CHECK("hsm_gate_bwd");
    int n = 3 * C * C + 1;
    hsm_wgrad_reduce_kernel<<<cdiv(n, 32), 256, 0, st>>>(wpart, d.B * GNS, n, a->d_w_out, C * C, a->d_w_hz, 2 * C * C, a->d_D, 1);
    KMU_LAUNCH_CHECK("hsm_wgrad_reduce(gate)");
  }
  if (fused) {
    // P recomputed into TMEM (bit-identical to the forward's), dP planes + the direct part of dx from the same CTA (hsm_fused.cu)
    int rc = fz::dp(a->x, a->dy, a->w_bcdt, a->w_dw, a->stats, dhs, a->h, r, dP, a->dx, d.B, C, d.H, ws + w.wpk, st);
    if (rc != KMU_OK) return rc;
  } else {
    size_t smem = ((size_t)2 * C * 64 + 192) * 4;
    dim3 grid(cdiv(d.L, 256), d.B);
#define KMU_HSM_DP(CC)                                                                                              \
  do {                                                                                                              \
    if (tcbwd) {                                                                                                    \
      opt_in_smem(hsm_dp_kernel<CC, true>, smem);                                                                   \
      hsm_dp_kernel<CC, true><<<grid, 256, smem, st>>>(a->x, a->dy, a->P, a->stats, dhs, a->h, r, dP, a->dx, d);     \
    } else {                                                                                                        \
      opt_in_smem(hsm_dp_kernel<CC, false>, smem);                                                                  \
      hsm_dp_kernel<CC, false><<<grid, 256, smem, st>>>(a->x, a->dy, a->P, a->stats, dhs, a->h, r, dP, a->dx, d);    \
    }                                                                                                               \
  } while (0)
    switch (C) {
      case 16: KMU_HSM_DP(16); break;
      case 32: KMU_HSM_DP(32); break;
      default: KMU_HSM_DP(64); break;
    }
#undef KMU_HSM_DP
    KMU_LAUNCH_CHECK("hsm_dp");
  }
  if (tcbwd) {
    int rc = tcb::backward(a->x, a->w_bcdt, a->w_dw, dP, a->dx, a->d_w_bcdt, a->d_w_dw, d.B, C, d.H, tpart, 1, st);
    if (rc != KMU_OK) return rc;
  } else {
    size_t smem = ((size_t)BN * NHALO + 2 * (size_t)BN * QP + (size_t)C * QP + 2 * (size_t)C * BN + BN * 12) * 4;
    int tiles = d.tiles_x * d.tiles_y;
    dim3 grid(tiles, d.B);
#define KMU_HSM_PBWD(CC)                                                                                            \
  do {                                                                                                              \
    opt_in_smem(hsm_proj_dw_bwd_kernel<CC>, smem);                                                                  \
    hsm_proj_dw_bwd_kernel<CC><<<grid, 256, smem, st>>>(a->x, a->w_bcdt, a->w_dw, dP, a->dx, tpart, d);              \
  } while (0)
    switch (C) {
      case 16: KMU_HSM_PBWD(16); break;
      case 32: KMU_HSM_PBWD(32); break;
      default: KMU_HSM_PBWD(64); break;
    }
#undef KMU_HSM_PBWD
    KMU_LAUNCH_CHECK("hsm_proj_dw_bwd");
    int per = N3 * C + N3 * 9;
    hsm_wgrad_reduce_kernel<<<cdiv(per, 32), 256, 0, st>>>(tpart, tiles * d.B, per, a->d_w_bcdt, N3 * C, a->d_w_dw, N3 * 9,
                                                            nullptr, 0);
    KMU_LAUNCH_CHECK("hsm_wgrad_reduce(proj)");
  }
  if (a->d_A) {
    fill_kernel<<<1, 64, 0, st>>>(a->d_A, N, 0.f);
    KMU_LAUNCH_CHECK("hsm_fill_dA");
  }
  return KMU_OK;
}

int kmu_layernorm1d_fwd(const float* x, const float* weight, const float* bias, float* y, float* rstd, int32_t B, int32_t C,
                        int32_t L, float eps, kmu_stream stream) {
  KMU_REQUIRE(x && weight && bias && y && B > 0 && C > 0 && L > 0, KMU_ERR_BAD_ARG, "layernorm1d_fwd: bad argument");
  const int nb = cdiv((long long)B * L, 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 16: ln1d_fwd_reg_kernel<16><<<nb, 256, 0, st>>>(x, weight, bias, y, rstd, B, L, eps); break;
    case 32: ln1d_fwd_reg_kernel<32><<<nb, 256, 0, st>>>(x, weight, bias, y, rstd, B, L, eps); break;
    case 64: ln1d_fwd_reg_kernel<64><<<nb, 256, 0, st>>>(x, weight, bias, y, rstd, B, L, eps); break;
    default: ln1d_fwd_kernel<<<nb, 256, 0, st>>>(x, weight, bias, y, rstd, B, C, L, eps);
  }
  KMU_LAUNCH_CHECK("ln1d_fwd");
  return KMU_OK;
}

size_t kmu_layernorm1d_bwd_workspace_bytes(int32_t B, int32_t C, int32_t L) {
  if (B <= 0 || C <= 0 || L <= 0) return 0;
  return align_up((size_t)cdiv((long long)B * L, 256) * 2 * C * 4, 256);
}

int kmu_layernorm1d_bwd(const float* x, const float* weight, const float* dy, float* dx, float* dweight, float* dbias, int32_t B,
                        int32_t C, int32_t L, float eps, void* workspace, size_t workspace_bytes, kmu_stream stream) {
  KMU_REQUIRE(x && weight && dy && dx && dweight && dbias && B > 0 && C > 0 && L > 0, KMU_ERR_BAD_ARG,
              "layernorm1d_bwd: bad argument");
  KMU_REQUIRE(C <= 1024, KMU_ERR_UNSUPPORTED, "layernorm1d_bwd: C=%d > 1024", C);
  KMU_REQUIRE(workspace && workspace_bytes >= kmu_layernorm1d_bwd_workspace_bytes(B, C, L), KMU_ERR_WORKSPACE,
              "layernorm1d_bwd: workspace too small");
  const int nblk = cdiv((long long)B * L, 256);
  float* part = (float*)workspace;
  const size_t smem = (size_t)16 * C * 4;
  if (smem > 48 * 1024) cudaFuncSetAttribute(ln1d_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  switch (C) {
    case 16: ln1d_bwd_reg_kernel<16><<<nblk, 256, 0, (cudaStream_t)stream>>>(x, weight, dy, dx, part, B, L, eps); break;
    case 32: ln1d_bwd_reg_kernel<32><<<nblk, 256, 0, (cudaStream_t)stream>>>(x, weight, dy, dx, part, B, L, eps); break;
    case 64: ln1d_bwd_reg_kernel<64><<<nblk, 256, 0, (cudaStream_t)stream>>>(x, weight, dy, dx, part, B, L, eps); break;
    default: ln1d_bwd_kernel<<<nblk, 256, smem, (cudaStream_t)stream>>>(x, weight, dy, dx, part, B, C, L, eps);
  }
  KMU_LAUNCH_CHECK("ln1d_bwd");
  ln1d_bwd_reduce_kernel<<<cdiv(2 * C, 32), 256, 0, (cudaStream_t)stream>>>(part, nblk, C, dweight, dbias);
  KMU_LAUNCH_CHECK("ln1d_bwd_reduce");
  return KMU_OK;
}

}  // extern "C"
