// dagem.cu -- DAGEM attention-gated fusion (edge/vertex gating MLPs with train-mode BatchNorm + final 1x1 fusion),
// forward and backward.
//
// Replaces DAGEM_md.py:62-92,104-110 and the autograd graph PyTorch builds for it (38 fwd / 98 fwd+bwd ATen ops on a
// 2 MB tensor -> 7 / 13 launches).  Algebra, nb_k = circular neighbours x[h-1], x[h+1], x[w-1], x[w+1]:
//     s    = x . sum_k a_k nb_k + a_b                 agg = ReLU(BN0(s))          BN0: one channel over B*C*H*W
//     ue_k = We[:, :C] x + We[:, C:] (x . nb_k) + e_b                             BN1: Ch channels over B*H*W*4
//     uvp  = Wv [x; agg] + v_b                        uv  = ReLU(BN2(uvp))        BN2: Ch channels over B*H*W
//     r    = sum_k r_k ReLU(BN1(ue_k)) + r_b          ur  = ReLU(BN3(r))          BN3: one channel over B*Ch*H*W
//     z    = Wf [D; uv . ur]                          out = ReLU(BN4(z))          BN4: C channels over B*H*W
// D = deform_conv3x3(x, offset_conv(x)) + x is the caller's tensor (DAGEM_md.py:95-104).
// Every BatchNorm is a grid-wide reduction, so the chain is cut exactly there: each kernel writes per-CTA (sum, sumsq)
// partials, a finalize kernel folds them in double precision into (mean, rstd, scale, shift) and updates the running
// statistics.  All reductions use fixed orders (no atomics): results are bit-reproducible.
// Threads: CTA = 64 consecutive pixels x 4 "parts"; a part owns a slice of the output channels (or one neighbour k), so the
// 32 lanes of a warp always share channels and the BN partial sums are plain warp-shuffle reductions.
#include "common.cuh"

namespace kmu {
namespace dagem {

constexpr int TP = 64;    // pixels per CTA
constexpr int NT = 256;   // threads per CTA
constexpr int PAD = 65;   // row pitch of [channel][pixel] tiles in shared memory
constexpr int NEX = 6;    // extra "stat channels" carrying the scalar-parameter gradients (edge_aggregation / reduce Linear(4,1))

struct Dims {
  int B, H, W, HW, NPIX, nblk;
};

template <int C>
struct Lay {
  static constexpr int Ch = C / 2;
  static constexpr int ST_S = 0, ST_UE = 1, ST_UV = 1 + Ch, ST_R = 1 + 2 * Ch, ST_Z = 2 + 2 * Ch, NSTAT = 2 + 2 * Ch + C;
  static constexpr int EX = NSTAT;  // EX+0: d r_w[0,1]; EX+1: d r_w[2,3]; EX+2: d r_b; EX+3: d a_w[0,1]; EX+4: d a_w[2,3]; EX+5: d a_b
  static constexpr int NCHAN = NSTAT + NEX;
};

// saved-for-backward buffer (floats): s | ue | uvp | r | z | stat (NSTAT x 4)
struct Saved {
  size_t s, ue, uvp, r, z, stat, total;
};
static Saved saved_layout(int B, int C, int HW) {
  Saved v;
  size_t n = (size_t)B * HW, Ch = C / 2, o = 0;
  v.s = o; o += n * C;
  v.ue = o; o += n * 4 * Ch;
  v.uvp = o; o += n * Ch;
  v.r = o; o += n * Ch;
  v.z = o; o += n * C;
  v.stat = o; o += (size_t)(2 + 2 * Ch + C) * 4;
  v.total = o;
  return v;
}

struct Pix {
  int b, hw, h, w;
  bool valid;
};
__device__ __forceinline__ Pix decode(int p, const Dims& d) {
  Pix q;
  q.valid = p < d.NPIX;
  int pp = q.valid ? p : 0;
  q.b = pp / d.HW;
  q.hw = pp - q.b * d.HW;
  q.h = q.hw / d.W;
  q.w = q.hw - q.h * d.W;
  return q;
}
// neighbour k of (h,w): 0 = (h-1,w), 1 = (h+1,w), 2 = (h,w-1), 3 = (h,w+1), circular (DAGEM_md.py:64-67)
__device__ __forceinline__ int nbr(int k, int h, int w, const Dims& d) {
  if (k == 0) return (h == 0 ? d.H - 1 : h - 1) * d.W + w;
  if (k == 1) return (h == d.H - 1 ? 0 : h + 1) * d.W + w;
  if (k == 2) return h * d.W + (w == 0 ? d.W - 1 : w - 1);
  return h * d.W + (w == d.W - 1 ? 0 : w + 1);
}
// pixel whose neighbour k is (h,w)
__device__ __forceinline__ int nbr_inv(int k, int h, int w, const Dims& d) { return nbr(k ^ 1, h, w, d); }

// per-warp partial sums of one stat channel; every warp owns its own slot, so there is no race and the order is fixed
__device__ __forceinline__ void wstat(float* s_w, int nchan, int ch, float a, float b) {
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) {
    float* p = s_w + ((size_t)(threadIdx.x >> 5) * nchan + ch) * 2;
    p[0] += a;
    p[1] += b;
  }
}
__device__ __forceinline__ void wstat_zero(float* s_w, int nchan) {
  for (int i = threadIdx.x; i < 8 * nchan * 2; i += NT) s_w[i] = 0.f;
}
// part[(ch*2+j)*nblk + blk] = sum over the 8 warps (call after __syncthreads)
__device__ __forceinline__ void wstat_flush(const float* s_w, int nchan, int ch0, int nch, float* __restrict__ part, int nblk) {
  for (int i = threadIdx.x; i < nch * 2; i += NT) {
    int ch = ch0 + (i >> 1), j = i & 1;
    float a = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) a += s_w[((size_t)wv * nchan + ch) * 2 + j];
    part[(size_t)(ch * 2 + j) * nblk + blockIdx.x] = a;
  }
}

// out[ia*NB + ib] = sum_px a_s[ia][px] * b_s[ib][px] for this CTA's 64 pixels (weight-gradient partial)
__device__ __forceinline__ void outer_partial(const float* a_s, int NA, const float* b_s, int NB, float* __restrict__ out) {
  for (int idx = threadIdx.x; idx < NA * NB; idx += NT) {
    int ia = idx / NB, ib = idx - ia * NB;
    const float* ar = a_s + ia * PAD;
    const float* br = b_s + ib * PAD;
    float acc = 0.f;
#pragma unroll 8
    for (int p = 0; p < TP; ++p) acc = fmaf(ar[p], br[p], acc);
    out[idx] = acc;
  }
}

// ================================================================================================ forward
// ---- K1: s and the four pre-BN edge updates ue_k.  part = neighbour k.
template <int C>
__global__ void __launch_bounds__(NT) edge_fwd_kernel(const float* __restrict__ x, const float* __restrict__ ea_w,
                                                      const float* __restrict__ ea_b, const float* __restrict__ eu_w,
                                                      const float* __restrict__ eu_b, float* __restrict__ s_out,
                                                      float* __restrict__ ue_out, float* __restrict__ part, Dims d) {
  using L = Lay<C>;
  constexpr int Ch = L::Ch;
  extern __shared__ __align__(16) float smem[];
  float* wa_s = smem;               // [C][Ch]  We[:, :C] transposed
  float* wb_s = wa_s + C * Ch;      // [C][Ch]  We[:, C:] transposed
  float* s_w = wb_s + C * Ch;       // [8][NCHAN][2]
  const int tid = threadIdx.x, px = tid & 63, k = tid >> 6;
  for (int i = tid; i < C * Ch; i += NT) {
    int c = i / Ch, o = i - c * Ch;
    wa_s[i] = eu_w[(size_t)o * 2 * C + c];
    wb_s[i] = eu_w[(size_t)o * 2 * C + C + c];
  }
  wstat_zero(s_w, L::NCHAN);
  __syncthreads();
  const Pix q = decode(blockIdx.x * TP + px, d);
  const int nk = nbr(k, q.h, q.w, d);
  float acc[Ch];
#pragma unroll
  for (int o = 0; o < Ch; ++o) acc[o] = eu_b[o];
  const float* xb = x + (size_t)q.b * C * d.HW;
  for (int c = 0; c < C; ++c) {
    float xv = q.valid ? __ldg(xb + (size_t)c * d.HW + q.hw) : 0.f;
    float nv = q.valid ? __ldg(xb + (size_t)c * d.HW + nk) : 0.f;
    float e = xv * nv;
    const float4* a4 = reinterpret_cast<const float4*>(wa_s + c * Ch);
    const float4* b4 = reinterpret_cast<const float4*>(wb_s + c * Ch);
#pragma unroll
    for (int o4 = 0; o4 < Ch / 4; ++o4) {
      float4 a = a4[o4], bb = b4[o4];
      acc[4 * o4 + 0] = fmaf(a.x, xv, fmaf(bb.x, e, acc[4 * o4 + 0]));
      acc[4 * o4 + 1] = fmaf(a.y, xv, fmaf(bb.y, e, acc[4 * o4 + 1]));
      acc[4 * o4 + 2] = fmaf(a.z, xv, fmaf(bb.z, e, acc[4 * o4 + 2]));
      acc[4 * o4 + 3] = fmaf(a.w, xv, fmaf(bb.w, e, acc[4 * o4 + 3]));
    }
  }
  float* uo = ue_out + ((size_t)(q.b * 4 + k) * Ch) * d.HW + q.hw;
#pragma unroll
  for (int o = 0; o < Ch; ++o) {
    float v = q.valid ? acc[o] : 0.f;
    if (q.valid) uo[(size_t)o * d.HW] = v;
    wstat(s_w, L::NCHAN, L::ST_UE + o, v, v * v);
  }
  // s over the CTA's 64 pixels x C channels
  const float w0 = ea_w[0], w1 = ea_w[1], w2 = ea_w[2], w3 = ea_w[3], bb = ea_b[0];
  float ssum = 0.f, ssq = 0.f;
  for (int c = k; c < C; c += 4) {
    float v = 0.f;
    if (q.valid) {
      const float* xc = xb + (size_t)c * d.HW;
      float xv = __ldg(xc + q.hw);
      float t = w0 * __ldg(xc + nbr(0, q.h, q.w, d)) + w1 * __ldg(xc + nbr(1, q.h, q.w, d)) +
                w2 * __ldg(xc + nbr(2, q.h, q.w, d)) + w3 * __ldg(xc + nbr(3, q.h, q.w, d));
      v = fmaf(xv, t, bb);
      s_out[((size_t)q.b * C + c) * d.HW + q.hw] = v;
      ssum += v;
      ssq = fmaf(v, v, ssq);
    }
  }
  wstat(s_w, L::NCHAN, L::ST_S, ssum, ssq);
  __syncthreads();
  wstat_flush(s_w, L::NCHAN, L::ST_S, 1 + Ch, part, d.nblk);
}

// ---- K2: pre-BN vertex update uvp and reduced edge gate r.  part = quarter of the Ch outputs.
template <int C>
__global__ void __launch_bounds__(NT) gate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ s_in,
                                                      const float* __restrict__ ue_in, const float* __restrict__ vu_w,
                                                      const float* __restrict__ vu_b, const float* __restrict__ er_w,
                                                      const float* __restrict__ er_b, const float4* __restrict__ stat,
                                                      float* __restrict__ uvp_out, float* __restrict__ r_out,
                                                      float* __restrict__ part, Dims d) {
  using L = Lay<C>;
  constexpr int Ch = L::Ch, OPP = Ch / 4;
  extern __shared__ __align__(16) float smem[];
  float* wv_s = smem;                // [2C][Ch] transposed
  float* s_w = wv_s + 2 * C * Ch;    // [8][NCHAN][2]
  const int tid = threadIdx.x, px = tid & 63, part_id = tid >> 6, o0 = part_id * OPP;
  for (int i = tid; i < 2 * C * Ch; i += NT) {
    int c = i / Ch, o = i - c * Ch;
    wv_s[i] = vu_w[(size_t)o * 2 * C + c];
  }
  wstat_zero(s_w, L::NCHAN);
  __syncthreads();
  const Pix q = decode(blockIdx.x * TP + px, d);
  const float4 st_s = stat[L::ST_S];
  float acc[OPP];
#pragma unroll
  for (int j = 0; j < OPP; ++j) acc[j] = vu_b[o0 + j];
  const float* xb = x + (size_t)q.b * C * d.HW + q.hw;
  const float* sb = s_in + (size_t)q.b * C * d.HW + q.hw;
  for (int c = 0; c < C; ++c) {
    float xv = q.valid ? __ldg(xb + (size_t)c * d.HW) : 0.f;
    float sv = q.valid ? __ldg(sb + (size_t)c * d.HW) : 0.f;
    float ag = fmaxf(fmaf(sv, st_s.z, st_s.w), 0.f);
    const float* wx = wv_s + c * Ch + o0;
    const float* wa = wv_s + (C + c) * Ch + o0;
#pragma unroll
    for (int j = 0; j < OPP; ++j) acc[j] = fmaf(wx[j], xv, fmaf(wa[j], ag, acc[j]));
  }
  float rs = 0.f, rq = 0.f;
  const float rw0 = er_w[0], rw1 = er_w[1], rw2 = er_w[2], rw3 = er_w[3], rb = er_b[0];
#pragma unroll
  for (int j = 0; j < OPP; ++j) {
    const int o = o0 + j;
    float v = q.valid ? acc[j] : 0.f;
    if (q.valid) uvp_out[((size_t)q.b * Ch + o) * d.HW + q.hw] = v;
    wstat(s_w, L::NCHAN, L::ST_UV + o, v, v * v);
    float rr = 0.f;
    if (q.valid) {
      const float4 st = stat[L::ST_UE + o];
      const float* up = ue_in + ((size_t)(q.b * 4) * Ch + o) * d.HW + q.hw;
      const size_t ks = (size_t)Ch * d.HW;
      rr = rb;
      rr = fmaf(rw0, fmaxf(fmaf(__ldg(up), st.z, st.w), 0.f), rr);
      rr = fmaf(rw1, fmaxf(fmaf(__ldg(up + ks), st.z, st.w), 0.f), rr);
      rr = fmaf(rw2, fmaxf(fmaf(__ldg(up + 2 * ks), st.z, st.w), 0.f), rr);
      rr = fmaf(rw3, fmaxf(fmaf(__ldg(up + 3 * ks), st.z, st.w), 0.f), rr);
      r_out[((size_t)q.b * Ch + o) * d.HW + q.hw] = rr;
    }
    rs += rr;
    rq = fmaf(rr, rr, rq);
  }
  wstat(s_w, L::NCHAN, L::ST_R, rs, rq);
  __syncthreads();
  wstat_flush(s_w, L::NCHAN, L::ST_UV, Ch + 1, part, d.nblk);
}

// ---- K3: z = Wf [D; uv.ur].  part = quarter of the C outputs.
template <int C>
__global__ void __launch_bounds__(NT) fuse_fwd_kernel(const float* __restrict__ dfm, const float* __restrict__ uvp_in,
                                                      const float* __restrict__ r_in, const float* __restrict__ wf,
                                                      const float4* __restrict__ stat, float* __restrict__ z_out,
                                                      float* __restrict__ part, Dims d) {
  using L = Lay<C>;
  constexpr int Ch = L::Ch, OPP = C / 4, NI = C + Ch;
  extern __shared__ __align__(16) float smem[];
  float* wf_s = smem;             // [NI][C] transposed
  float* s_w = wf_s + NI * C;     // [8][NCHAN][2]
  const int tid = threadIdx.x, px = tid & 63, part_id = tid >> 6, o0 = part_id * OPP;
  for (int i = tid; i < NI * C; i += NT) {
    int c = i / C, o = i - c * C;
    wf_s[i] = wf[(size_t)o * NI + c];
  }
  wstat_zero(s_w, L::NCHAN);
  __syncthreads();
  const Pix q = decode(blockIdx.x * TP + px, d);
  float acc[OPP];
#pragma unroll
  for (int j = 0; j < OPP; ++j) acc[j] = 0.f;
  const float* db = dfm + (size_t)q.b * C * d.HW + q.hw;
  for (int c = 0; c < C; ++c) {
    float dv = q.valid ? __ldg(db + (size_t)c * d.HW) : 0.f;
    const float* wr = wf_s + c * C + o0;
#pragma unroll
    for (int j = 0; j < OPP; ++j) acc[j] = fmaf(wr[j], dv, acc[j]);
  }
  const float4 st_r = stat[L::ST_R];
  const float* ub = uvp_in + (size_t)q.b * Ch * d.HW + q.hw;
  const float* rb = r_in + (size_t)q.b * Ch * d.HW + q.hw;
  for (int jj = 0; jj < Ch; ++jj) {
    float f = 0.f;
    if (q.valid) {
      const float4 st = stat[L::ST_UV + jj];
      float uv = fmaxf(fmaf(__ldg(ub + (size_t)jj * d.HW), st.z, st.w), 0.f);
      float ur = fmaxf(fmaf(__ldg(rb + (size_t)jj * d.HW), st_r.z, st_r.w), 0.f);
      f = uv * ur;
    }
    const float* wr = wf_s + (C + jj) * C + o0;
#pragma unroll
    for (int j = 0; j < OPP; ++j) acc[j] = fmaf(wr[j], f, acc[j]);
  }
#pragma unroll
  for (int j = 0; j < OPP; ++j) {
    float v = q.valid ? acc[j] : 0.f;
    if (q.valid) z_out[((size_t)q.b * C + o0 + j) * d.HW + q.hw] = v;
    wstat(s_w, L::NCHAN, L::ST_Z + o0 + j, v, v * v);
  }
  __syncthreads();
  wstat_flush(s_w, L::NCHAN, L::ST_Z, C, part, d.nblk);
}

// ---- K4: out = ReLU(BN4(z))
__global__ void __launch_bounds__(256) out_fwd_kernel(const float* __restrict__ z, const float4* __restrict__ stat, int st_z, int C,
                                                      int HW, long long total, float* __restrict__ out) {
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  int c = (int)((i / HW) % C);
  const float4 st = stat[st_z + c];
  out[i] = fmaxf(fmaf(z[i], st.z, st.w), 0.f);
}

// ---- finalize the BN statistics of up to two groups of stat channels.  One CTA (128 threads) per channel.
struct FinGroup {
  int ch0, nch;
  float count;
  const float* gamma;
  const float* beta;
  float* rmean;
  float* rvar;
};
__global__ void __launch_bounds__(128) fin_fwd_kernel(const float* __restrict__ part, int nblk, float4* __restrict__ stat,
                                                      FinGroup g0, FinGroup g1, int training, float momentum, float eps) {
  __shared__ double red[2][128];
  const bool first = (int)blockIdx.x < g0.nch;
  const FinGroup& g = first ? g0 : g1;
  const int local = first ? blockIdx.x : blockIdx.x - g0.nch;
  const int ch = g.ch0 + local;
  double mean, var;
  if (training) {
    double s = 0.0, qq = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 128) {
      s += (double)part[(size_t)(ch * 2) * nblk + i];
      qq += (double)part[(size_t)(ch * 2 + 1) * nblk + i];
    }
    red[0][threadIdx.x] = s;
    red[1][threadIdx.x] = qq;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) {
        red[0][threadIdx.x] += red[0][threadIdx.x + o];
        red[1][threadIdx.x] += red[1][threadIdx.x + o];
      }
      __syncthreads();
    }
    mean = red[0][0] / (double)g.count;
    var = red[1][0] / (double)g.count - mean * mean;
    if (var < 0.0) var = 0.0;
  } else {
    mean = (double)g.rmean[local];
    var = (double)g.rvar[local];
  }
  if (threadIdx.x == 0) {
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float scale = g.gamma[local] * rstd;
    float shift = g.beta[local] - (float)mean * scale;
    stat[ch] = make_float4((float)mean, rstd, scale, shift);
    if (training && g.rmean && g.rvar) {
      double unb = g.count > 1.f ? var * (double)g.count / ((double)g.count - 1.0) : var;
      g.rmean[local] = (1.f - momentum) * g.rmean[local] + momentum * (float)mean;
      g.rvar[local] = (1.f - momentum) * g.rvar[local] + momentum * (float)unb;
    }
  }
}

// ================================================================================================ backward
// BatchNorm backward of v -> y = gamma xhat + beta with upstream g:  dv = gamma rstd (g - mean(g) - xhat mean(g xhat)),
// dgamma = sum g xhat, dbeta = sum g  (eval mode: the two means are 0).  bstat[ch] = (mean g, mean g xhat).
__global__ void __launch_bounds__(128) fin_bwd_kernel(const float* __restrict__ part, int nblk, float2* __restrict__ bstat,
                                                      int ch0a, int ncha, float counta, float* dga, float* dba, int ch0b,
                                                      int nchb, float countb, float* dgb, float* dbb, int training) {
  __shared__ double red[2][128];
  const bool first = (int)blockIdx.x < ncha;
  const int local = first ? blockIdx.x : blockIdx.x - ncha;
  const int ch = (first ? ch0a : ch0b) + local;
  const float count = first ? counta : countb;
  float* dg = first ? dga : dgb;
  float* db = first ? dba : dbb;
  double s = 0.0, qq = 0.0;
  for (int i = threadIdx.x; i < nblk; i += 128) {
    s += (double)part[(size_t)(ch * 2) * nblk + i];
    qq += (double)part[(size_t)(ch * 2 + 1) * nblk + i];
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = qq;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    db[local] = (float)red[0][0];
    dg[local] = (float)red[1][0];
    bstat[ch] = training ? make_float2((float)(red[0][0] / (double)count), (float)(red[1][0] / (double)count)) : make_float2(0.f, 0.f);
  }
}

// raw sums of `n` extra channels -> out[2*i], out[2*i+1]
__global__ void __launch_bounds__(128) fin_sum_kernel(const float* __restrict__ part, int nblk, int ch0, float* __restrict__ out) {
  __shared__ double red[2][128];
  const int ch = ch0 + blockIdx.x;
  double s = 0.0, qq = 0.0;
  for (int i = threadIdx.x; i < nblk; i += 128) {
    s += (double)part[(size_t)(ch * 2) * nblk + i];
    qq += (double)part[(size_t)(ch * 2 + 1) * nblk + i];
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = qq;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x] = (float)red[0][0];
    out[2 * blockIdx.x + 1] = (float)red[1][0];
  }
}

// ---- B1: statistics of g = dout . 1[out > 0] against zhat
template <int C>
__global__ void __launch_bounds__(NT) out_bwd_stats_kernel(const float* __restrict__ dout, const float* __restrict__ z,
                                                           const float4* __restrict__ stat, float* __restrict__ part, Dims d) {
  using L = Lay<C>;
  constexpr int OPP = C / 4;
  extern __shared__ __align__(16) float smem[];
  float* s_w = smem;
  const int tid = threadIdx.x, px = tid & 63, o0 = (tid >> 6) * OPP;
  wstat_zero(s_w, L::NCHAN);
  __syncthreads();
  const Pix q = decode(blockIdx.x * TP + px, d);
#pragma unroll
  for (int j = 0; j < OPP; ++j) {
    float g = 0.f, gx = 0.f;
    if (q.valid) {
      const float4 st = stat[L::ST_Z + o0 + j];
      size_t off = ((size_t)q.b * C + o0 + j) * d.HW + q.hw;
      float zv = __ldg(z + off);
      g = fmaf(zv, st.z, st.w) > 0.f ? __ldg(dout + off) : 0.f;
      gx = g * (zv - st.x) * st.y;
    }
    wstat(s_w, L::NCHAN, L::ST_Z + o0 + j, g, gx);
  }
  __syncthreads();
  wstat_flush(s_w, L::NCHAN, L::ST_Z, C, part, d.nblk);
}

// ---- B2: dz -> dD, dWf partial, g2 = d uv . 1[uv>0], g4 = d ur . 1[ur>0] and their statistics
template <int C>
__global__ void __launch_bounds__(NT) fuse_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ z,
                                                      const float* __restrict__ dfm, const float* __restrict__ uvp_in,
                                                      const float* __restrict__ r_in, const float* __restrict__ wf,
                                                      const float4* __restrict__ stat, const float2* __restrict__ bstat,
                                                      float* __restrict__ d_dfm, float* __restrict__ g2_out,
                                                      float* __restrict__ g4_out, float* __restrict__ wpart,
                                                      float* __restrict__ part, Dims d) {
  using L = Lay<C>;
  constexpr int Ch = L::Ch, OPP = C / 4, JPP = Ch / 4, NI = C + Ch;
  extern __shared__ __align__(16) float smem[];
  float* dz_s = smem;                 // [C][PAD]
  float* in_s = dz_s + C * PAD;       // [NI][PAD]
  float* wf_s = in_s + NI * PAD;      // [C][NI] natural
  float* s_w = wf_s + C * NI;
  const int tid = threadIdx.x, px = tid & 63, part_id = tid >> 6, o0 = part_id * OPP, j0 = part_id * JPP;
  for (int i = tid; i < C * NI; i += NT) wf_s[i] = wf[i];
  wstat_zero(s_w, L::NCHAN);
  const Pix q = decode(blockIdx.x * TP + px, d);
#pragma unroll
  for (int j = 0; j < OPP; ++j) {
    const int o = o0 + j;
    float dz = 0.f, dv = 0.f;
    if (q.valid) {
      const float4 st = stat[L::ST_Z + o];
      const float2 bs = bstat[L::ST_Z + o];
      size_t off = ((size_t)q.b * C + o) * d.HW + q.hw;
      float zv = __ldg(z + off);
      float g = fmaf(zv, st.z, st.w) > 0.f ? __ldg(dout + off) : 0.f;
      float zh = (zv - st.x) * st.y;
      dz = st.z * (g - bs.x - zh * bs.y);
      dv = __ldg(dfm + off);
    }
    dz_s[o * PAD + px] = dz;
    in_s[o * PAD + px] = dv;
  }
  float uv[JPP], ur[JPP], uvh[JPP], rh[JPP];
  const float4 st_r = stat[L::ST_R];
#pragma unroll
  for (int j = 0; j < JPP; ++j) {
    const int jj = j0 + j;
    uv[j] = ur[j] = uvh[j] = rh[j] = 0.f;
    if (q.valid) {
      const float4 st = stat[L::ST_UV + jj];
      size_t off = ((size_t)q.b * Ch + jj) * d.HW + q.hw;
      float a = __ldg(uvp_in + off), rr = __ldg(r_in + off);
      uv[j] = fmaxf(fmaf(a, st.z, st.w), 0.f);
      ur[j] = fmaxf(fmaf(rr, st_r.z, st_r.w), 0.f);
      uvh[j] = (a - st.x) * st.y;
      rh[j] = (rr - st_r.x) * st_r.y;
    }
    in_s[(C + jj) * PAD + px] = uv[j] * ur[j];
  }
  __syncthreads();
  // dD
#pragma unroll
  for (int j = 0; j < OPP; ++j) {
    const int c = o0 + j;
    float a = 0.f;
    for (int o = 0; o < C; ++o) a = fmaf(wf_s[o * NI + c], dz_s[o * PAD + px], a);
    if (q.valid) d_dfm[((size_t)q.b * C + c) * d.HW + q.hw] = a;
  }
  float s4 = 0.f, s4x = 0.f;
#pragma unroll
  for (int j = 0; j < JPP; ++j) {
    const int jj = j0 + j;
    float df = 0.f;
    for (int o = 0; o < C; ++o) df = fmaf(wf_s[o * NI + C + jj], dz_s[o * PAD + px], df);
    float g2 = uv[j] > 0.f ? df * ur[j] : 0.f;
    float g4 = ur[j] > 0.f ? df * uv[j] : 0.f;
    if (q.valid) {
      size_t off = ((size_t)q.b * Ch + jj) * d.HW + q.hw;
      g2_out[off] = g2;
      g4_out[off] = g4;
    } else {
      g2 = g4 = 0.f;
    }
    wstat(s_w, L::NCHAN, L::ST_UV + jj, g2, g2 * uvh[j]);
    s4 += g4;
    s4x = fmaf(g4, rh[j], s4x);
  }
  wstat(s_w, L::NCHAN, L::ST_R, s4, s4x);
  outer_partial(dz_s, C, in_s, NI, wpart + (size_t)blockIdx.x * C * NI);
  __syncthreads();
  wstat_flush(s_w, L::NCHAN, L::ST_UV, Ch + 1, part, d.nblk);
}

// ---- B3: d uvp, dr -> dagg (g1), dx (vertex part), dWv partial, reduce-Linear gradients, statistics of g1 and g3
template <int C>
__global__ void __launch_bounds__(NT) gate_bwd_kernel(const float* __restrict__ x, const float* __restrict__ s_in,
                                                      const float* __restrict__ ue_in, const float* __restrict__ uvp_in,
                                                      const float* __restrict__ r_in, const float* __restrict__ g2_in,
                                                      const float* __restrict__ g4_in, const float* __restrict__ vu_w,
                                                      const float* __restrict__ er_w, const float4* __restrict__ stat,
                                                      const float2* __restrict__ bstat, float* __restrict__ dr_out,
                                                      float* __restrict__ g1_out, float* __restrict__ dxa, float* __restrict__ wpart,
                                                      float* __restrict__ part, Dims d) {
  using L = Lay<C>;
  constexpr int Ch = L::Ch, OPP = C / 4, JPP = Ch / 4, NI = 2 * C + 1;
  extern __shared__ __align__(16) float smem[];
  float* du_s = smem;                 // [Ch][PAD]
  float* in_s = du_s + Ch * PAD;      // [NI][PAD]  x | agg | ones
  float* wv_s = in_s + NI * PAD;      // [Ch][2C] natural
  float* s_w = wv_s + Ch * 2 * C;
  const int tid = threadIdx.x, px = tid & 63, part_id = tid >> 6, o0 = part_id * OPP, j0 = part_id * JPP;
  for (int i = tid; i < Ch * 2 * C; i += NT) wv_s[i] = vu_w[i];
  wstat_zero(s_w, L::NCHAN);
  const Pix q = decode(blockIdx.x * TP + px, d);
  const float4 st_r = stat[L::ST_R], st_s = stat[L::ST_S];
  const float2 bs_r = bstat[L::ST_R];
  const float rw[4] = {er_w[0], er_w[1], er_w[2], er_w[3]};
  float ew[4] = {0.f, 0.f, 0.f, 0.f}, eb = 0.f;
#pragma unroll
  for (int j = 0; j < JPP; ++j) {
    const int jj = j0 + j;
    float du = 0.f, dr = 0.f;
    float g3s = 0.f, g3x = 0.f;
    if (q.valid) {
      size_t off = ((size_t)q.b * Ch + jj) * d.HW + q.hw;
      const float4 st = stat[L::ST_UV + jj];
      const float2 bs = bstat[L::ST_UV + jj];
      float uvh = (__ldg(uvp_in + off) - st.x) * st.y;
      du = st.z * (__ldg(g2_in + off) - bs.x - uvh * bs.y);
      float rh = (__ldg(r_in + off) - st_r.x) * st_r.y;
      dr = st_r.z * (__ldg(g4_in + off) - bs_r.x - rh * bs_r.y);
      dr_out[off] = dr;
      const float4 se = stat[L::ST_UE + jj];
      const float* up = ue_in + ((size_t)(q.b * 4) * Ch + jj) * d.HW + q.hw;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float upre = __ldg(up + (size_t)k * Ch * d.HW);
        float u = fmaxf(fmaf(upre, se.z, se.w), 0.f);
        ew[k] = fmaf(dr, u, ew[k]);
        float g3 = u > 0.f ? rw[k] * dr : 0.f;
        g3s += g3;
        g3x = fmaf(g3, (upre - se.x) * se.y, g3x);
      }
      eb += dr;
    }
    du_s[jj * PAD + px] = du;
    wstat(s_w, L::NCHAN, L::ST_UE + jj, g3s, g3x);
  }
  wstat(s_w, L::NCHAN, L::EX + 0, ew[0], ew[1]);
  wstat(s_w, L::NCHAN, L::EX + 1, ew[2], ew[3]);
  wstat(s_w, L::NCHAN, L::EX + 2, eb, 0.f);
  float agg[OPP], sh[OPP];
#pragma unroll
  for (int j = 0; j < OPP; ++j) {
    const int c = o0 + j;
    float xv = 0.f;
    agg[j] = sh[j] = 0.f;
    if (q.valid) {
      size_t off = ((size_t)q.b * C + c) * d.HW + q.hw;
      xv = __ldg(x + off);
      float sv = __ldg(s_in + off);
      agg[j] = fmaxf(fmaf(sv, st_s.z, st_s.w), 0.f);
      sh[j] = (sv - st_s.x) * st_s.y;
    }
    in_s[c * PAD + px] = xv;
    in_s[(C + c) * PAD + px] = agg[j];
  }
  if (part_id == 0) in_s[2 * C * PAD + px] = q.valid ? 1.f : 0.f;
  __syncthreads();
  float s1 = 0.f, s1x = 0.f;
#pragma unroll
  for (int j = 0; j < OPP; ++j) {
    const int c = o0 + j;
    float dxv = 0.f, dag = 0.f;
    for (int jj = 0; jj < Ch; ++jj) {
      float du = du_s[jj * PAD + px];
      dxv = fmaf(wv_s[jj * 2 * C + c], du, dxv);
      dag = fmaf(wv_s[jj * 2 * C + C + c], du, dag);
    }
    float g1 = agg[j] > 0.f ? dag : 0.f;
    if (q.valid) {
      size_t off = ((size_t)q.b * C + c) * d.HW + q.hw;
      dxa[off] = dxv;
      g1_out[off] = g1;
    } else {
      g1 = 0.f;
    }
    s1 += g1;
    s1x = fmaf(g1, sh[j], s1x);
  }
  wstat(s_w, L::NCHAN, L::ST_S, s1, s1x);
  outer_partial(du_s, Ch, in_s, NI, wpart + (size_t)blockIdx.x * Ch * NI);
  __syncthreads();
  wstat_flush(s_w, L::NCHAN, L::ST_S, 1 + Ch, part, d.nblk);
  wstat_flush(s_w, L::NCHAN, L::EX, 3, part, d.nblk);
}

// ---- B4: d ue_k and ds -> dx (local part), neighbour products t_k / t_s, dWe partial, aggregation-Linear gradients
template <int C>
__global__ void __launch_bounds__(NT) edge_bwd_kernel(const float* __restrict__ x, const float* __restrict__ s_in,
                                                      const float* __restrict__ ue_in, const float* __restrict__ dr_in,
                                                      const float* __restrict__ g1_in, const float* __restrict__ ea_w,
                                                      const float* __restrict__ eu_w, const float* __restrict__ er_w,
                                                      const float4* __restrict__ stat, const float2* __restrict__ bstat,
                                                      float* __restrict__ dxa, float* __restrict__ ts_out, float* __restrict__ tk_out,
                                                      float* __restrict__ wpart, float* __restrict__ part, Dims d) {
  using L = Lay<C>;
  constexpr int Ch = L::Ch, OPP = C / 4, JPP = Ch / 4;
  constexpr int NW = (Ch * C + NT - 1) / NT;   // dWea / dWeb outputs per thread
  extern __shared__ __align__(16) float smem[];
  float* du_s = smem;                 // [Ch][PAD]
  float* x_s = du_s + Ch * PAD;       // [C][PAD]
  float* e_s = x_s + C * PAD;         // [C][PAD]
  float* we_s = e_s + C * PAD;        // [Ch][2C] natural
  float* s_w = we_s + Ch * 2 * C;
  const int tid = threadIdx.x, px = tid & 63, part_id = tid >> 6, o0 = part_id * OPP, j0 = part_id * JPP;
  for (int i = tid; i < Ch * 2 * C; i += NT) we_s[i] = eu_w[i];
  wstat_zero(s_w, L::NCHAN);
  const Pix q = decode(blockIdx.x * TP + px, d);
  const float4 st_s = stat[L::ST_S];
  const float2 bs_s = bstat[L::ST_S];
  const float aw[4] = {ea_w[0], ea_w[1], ea_w[2], ea_w[3]};
  int nk[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) nk[k] = nbr(k, q.h, q.w, d);
  // ds path and x tile
  float xv[OPP], dxl[OPP];
  float eaw[4] = {0.f, 0.f, 0.f, 0.f}, eab = 0.f;
#pragma unroll
  for (int j = 0; j < OPP; ++j) {
    const int c = o0 + j;
    xv[j] = dxl[j] = 0.f;
    if (q.valid) {
      const float* xc = x + ((size_t)q.b * C + c) * d.HW;
      size_t off = ((size_t)q.b * C + c) * d.HW + q.hw;
      xv[j] = __ldg(xc + q.hw);
      float sv = __ldg(s_in + off);
      float ds = st_s.z * (__ldg(g1_in + off) - bs_s.x - (sv - st_s.x) * st_s.y * bs_s.y);
      float nsum = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float nv = __ldg(xc + nk[k]);
        nsum = fmaf(aw[k], nv, nsum);
        eaw[k] = fmaf(ds * xv[j], nv, eaw[k]);
      }
      eab += ds;
      dxl[j] = dxa[off] + ds * nsum;
      ts_out[off] = ds * xv[j];
    }
    x_s[c * PAD + px] = xv[j];
  }
  wstat(s_w, L::NCHAN, L::EX + 3, eaw[0], eaw[1]);
  wstat(s_w, L::NCHAN, L::EX + 4, eaw[2], eaw[3]);
  wstat(s_w, L::NCHAN, L::EX + 5, eab, 0.f);
  float wa_acc[NW], wb_acc[NW], bias_acc = 0.f;
#pragma unroll
  for (int r = 0; r < NW; ++r) wa_acc[r] = wb_acc[r] = 0.f;
  for (int k = 0; k < 4; ++k) {
    __syncthreads();  // previous round's readers of du_s / e_s are done
    const float rwk = er_w[k];
#pragma unroll
    for (int j = 0; j < JPP; ++j) {
      const int jj = j0 + j;
      float du = 0.f;
      if (q.valid) {
        const float4 se = stat[L::ST_UE + jj];
        const float2 be = bstat[L::ST_UE + jj];
        float upre = __ldg(ue_in + ((size_t)(q.b * 4 + k) * Ch + jj) * d.HW + q.hw);
        float dr = __ldg(dr_in + ((size_t)q.b * Ch + jj) * d.HW + q.hw);
        float g3 = fmaf(upre, se.z, se.w) > 0.f ? rwk * dr : 0.f;
        du = se.z * (g3 - be.x - (upre - se.x) * se.y * be.y);
      }
      du_s[jj * PAD + px] = du;
    }
    float nv[OPP];
#pragma unroll
    for (int j = 0; j < OPP; ++j) {
      const int c = o0 + j;
      nv[j] = q.valid ? __ldg(x + ((size_t)q.b * C + c) * d.HW + nk[k]) : 0.f;
      e_s[c * PAD + px] = xv[j] * nv[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < OPP; ++j) {
      const int c = o0 + j;
      float da = 0.f, de = 0.f;
      for (int o = 0; o < Ch; ++o) {
        float du = du_s[o * PAD + px];
        da = fmaf(we_s[o * 2 * C + c], du, da);
        de = fmaf(we_s[o * 2 * C + C + c], du, de);
      }
      dxl[j] += da + de * nv[j];
      if (q.valid) tk_out[((size_t)(q.b * 4 + k) * C + c) * d.HW + q.hw] = de * xv[j];
    }
#pragma unroll
    for (int r = 0; r < NW; ++r) {
      int idx = tid + r * NT;
      if (idx < Ch * C) {
        int o = idx / C, c = idx - o * C;
        const float* ar = du_s + o * PAD;
        const float* xr = x_s + c * PAD;
        const float* er = e_s + c * PAD;
        float a = 0.f, bq = 0.f;
#pragma unroll 8
        for (int p = 0; p < TP; ++p) {
          float du = ar[p];
          a = fmaf(du, xr[p], a);
          bq = fmaf(du, er[p], bq);
        }
        wa_acc[r] += a;
        wb_acc[r] += bq;
      }
    }
    if (tid < Ch) {
      float a = 0.f;
      for (int p = 0; p < TP; ++p) a += du_s[tid * PAD + p];
      bias_acc += a;
    }
  }
#pragma unroll
  for (int j = 0; j < OPP; ++j)
    if (q.valid) dxa[((size_t)q.b * C + o0 + j) * d.HW + q.hw] = dxl[j];
  // partial layout per CTA: dWe [Ch][2C] | d e_b [Ch]
  float* wp = wpart + (size_t)blockIdx.x * (Ch * 2 * C + Ch);
#pragma unroll
  for (int r = 0; r < NW; ++r) {
    int idx = tid + r * NT;
    if (idx < Ch * C) {
      int o = idx / C, c = idx - o * C;
      wp[o * 2 * C + c] = wa_acc[r];
      wp[o * 2 * C + C + c] = wb_acc[r];
    }
  }
  if (tid < Ch) wp[Ch * 2 * C + tid] = bias_acc;
  __syncthreads();
  wstat_flush(s_w, L::NCHAN, L::EX + 3, 3, part, d.nblk);
}

// ---- B5: dx = local part + the products routed through the circular neighbours
__global__ void __launch_bounds__(256) dx_gather_kernel(const float* __restrict__ dxa, const float* __restrict__ ts,
                                                        const float* __restrict__ tk, const float* __restrict__ ea_w, int C,
                                                        Dims d, long long total, float* __restrict__ dx) {
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  int hw = (int)(i % d.HW);
  long long bc = i / d.HW;
  int c = (int)(bc % C);
  int b = (int)(bc / C);
  int h = hw / d.W, w = hw - h * d.W;
  float a = dxa[i];
  const float* tsp = ts + (size_t)bc * d.HW;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int qk = nbr_inv(k, h, w, d);
    a = fmaf(ea_w[k], tsp[qk], a);
    a += tk[((size_t)(b * 4 + k) * C + c) * d.HW + qk];
  }
  dx[i] = a;
}

// ---- fixed-order sum of per-CTA weight-gradient partials: out0[0..n0) | out1[0..n1) laid out as rows of `row` floats where the
//      first `split` go to out0 and the rest to out1 (split == row -> everything to out0); tail0 = extra trailing block to out2.
__global__ void __launch_bounds__(128) wreduce_kernel(const float* __restrict__ wpart, int nblk, int per, int rows, int row,
                                                      int split, float* __restrict__ out0, float* __restrict__ out1,
                                                      float* __restrict__ out2) {
  int idx = blockIdx.x * 128 + threadIdx.x;
  if (idx >= per) return;
  float a = 0.f;
  for (int k = 0; k < nblk; ++k) a += wpart[(size_t)k * per + idx];
  if (idx < rows * row) {
    int r = idx / row, cidx = idx - r * row;
    if (cidx < split) out0[r * split + cidx] = a;
    else out1[r * (row - split) + (cidx - split)] = a;
  } else {
    out2[idx - rows * row] = a;
  }
}

static int check(const kmu_dagem_desc* d, const char* who) {
  KMU_REQUIRE(d != nullptr, KMU_ERR_BAD_ARG, "%s: null descriptor", who);
  KMU_REQUIRE(d->B > 0 && d->C > 0 && d->H > 0 && d->W > 0, KMU_ERR_BAD_ARG, "%s: non-positive shape", who);
  KMU_REQUIRE(d->C == 8 || d->C == 16 || d->C == 32 || d->C == 64, KMU_ERR_UNSUPPORTED, "%s: input_channels %d not in {8,16,32,64}",
              who, d->C);
  KMU_REQUIRE((long long)d->B * d->H * d->W < (1LL << 30), KMU_ERR_UNSUPPORTED, "%s: too many pixels", who);
  return KMU_OK;
}

static Dims make_dims(const kmu_dagem_desc& s) {
  Dims d;
  d.B = s.B; d.H = s.H; d.W = s.W; d.HW = s.H * s.W;
  d.NPIX = s.B * d.HW;
  d.nblk = cdiv(d.NPIX, TP);
  return d;
}

struct FwdWs { size_t part, total; };
static FwdWs fwd_ws(const kmu_dagem_desc& s) {
  Dims d = make_dims(s);
  FwdWs w;
  size_t nchan = 2 + 2 * (s.C / 2) + s.C + NEX;
  w.part = 0;
  w.total = align_up(nchan * 2 * d.nblk * 4, 256);
  return w;
}
struct BwdWs { size_t part, bstat, ex, g2, g4, dr, g1, dxa, ts, tk, wpart, total; };
static BwdWs bwd_ws(const kmu_dagem_desc& s) {
  Dims d = make_dims(s);
  const size_t C = s.C, Ch = s.C / 2, n = (size_t)d.NPIX;
  size_t nchan = 2 + 2 * Ch + C + NEX;
  BwdWs w;
  size_t o = 0;
  w.part = o; o += align_up(nchan * 2 * d.nblk * 4, 256);
  w.bstat = o; o += align_up(nchan * 8, 256);
  w.ex = o; o += align_up(NEX * 2 * 4, 256);
  w.g2 = o; o += align_up(n * Ch * 4, 256);
  w.g4 = o; o += align_up(n * Ch * 4, 256);
  w.dr = o; o += align_up(n * Ch * 4, 256);
  w.g1 = o; o += align_up(n * C * 4, 256);
  w.dxa = o; o += align_up(n * C * 4, 256);
  w.ts = o; o += align_up(n * C * 4, 256);
  w.tk = o; o += align_up(n * 4 * C * 4, 256);
  size_t per = C * (C + Ch);                       // dWf
  if (Ch * (2 * C + 1) > per) per = Ch * (2 * C + 1);
  w.wpart = o; o += align_up(per * d.nblk * 4, 256);
  w.total = o;
  return w;
}

template <typename K>
static void opt_in(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

static FinGroup group(const kmu_dagem_bn& bn, int ch0, int nch, double count) {
  FinGroup g;
  g.ch0 = ch0; g.nch = nch; g.count = (float)count;
  g.gamma = bn.weight; g.beta = bn.bias; g.rmean = bn.running_mean; g.rvar = bn.running_var;
  return g;
}

template <int C>
static int forward(const kmu_dagem_fwd_args* a, cudaStream_t st) {
  using L = Lay<C>;
  constexpr int Ch = L::Ch;
  const Dims d = make_dims(a->d);
  const Saved sv = saved_layout(d.B, C, d.HW);
  float* S = a->saved;
  float4* stat = reinterpret_cast<float4*>(S + sv.stat);
  float* part = (float*)a->workspace;
  const int training = a->d.training;
  const float mom = a->d.momentum, eps = a->d.eps;
  const double npx = (double)d.NPIX;
  const FinGroup g_s = group(a->bn[0], L::ST_S, 1, npx * C), g_ue = group(a->bn[1], L::ST_UE, Ch, npx * 4),
                 g_uv = group(a->bn[2], L::ST_UV, Ch, npx), g_r = group(a->bn[3], L::ST_R, 1, npx * Ch),
                 g_z = group(a->bn[4], L::ST_Z, C, npx);
  const size_t sw = (size_t)8 * L::NCHAN * 2 * 4;
  {
    size_t smem = (size_t)2 * C * Ch * 4 + sw;
    opt_in(edge_fwd_kernel<C>, smem);
    edge_fwd_kernel<C><<<d.nblk, NT, smem, st>>>(a->x, a->ea_w, a->ea_b, a->eu_w, a->eu_b, S + sv.s, S + sv.ue, part, d);
    KMU_LAUNCH_CHECK("dagem_edge_fwd");
    fin_fwd_kernel<<<1 + Ch, 128, 0, st>>>(part, d.nblk, stat, g_s, g_ue, training, mom, eps);
    KMU_LAUNCH_CHECK("dagem_fin_fwd(0,1)");
  }
  {
    size_t smem = (size_t)2 * C * Ch * 4 + sw;
    opt_in(gate_fwd_kernel<C>, smem);
    gate_fwd_kernel<C><<<d.nblk, NT, smem, st>>>(a->x, S + sv.s, S + sv.ue, a->vu_w, a->vu_b, a->er_w, a->er_b, stat, S + sv.uvp,
                                                 S + sv.r, part, d);
    KMU_LAUNCH_CHECK("dagem_gate_fwd");
    fin_fwd_kernel<<<Ch + 1, 128, 0, st>>>(part, d.nblk, stat, g_uv, g_r, training, mom, eps);
    KMU_LAUNCH_CHECK("dagem_fin_fwd(2,3)");
  }
  {
    size_t smem = (size_t)(C + Ch) * C * 4 + sw;
    opt_in(fuse_fwd_kernel<C>, smem);
    fuse_fwd_kernel<C><<<d.nblk, NT, smem, st>>>(a->deformed, S + sv.uvp, S + sv.r, a->wf, stat, S + sv.z, part, d);
    KMU_LAUNCH_CHECK("dagem_fuse_fwd");
    FinGroup none = g_z;
    none.nch = 0;
    fin_fwd_kernel<<<C, 128, 0, st>>>(part, d.nblk, stat, g_z, none, training, mom, eps);
    KMU_LAUNCH_CHECK("dagem_fin_fwd(4)");
  }
  {
    long long total = (long long)d.NPIX * C;
    out_fwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(S + sv.z, stat, L::ST_Z, C, d.HW, total, a->out);
    KMU_LAUNCH_CHECK("dagem_out_fwd");
  }
  return KMU_OK;
}

template <int C>
static int backward(const kmu_dagem_bwd_args* a, cudaStream_t st) {
  using L = Lay<C>;
  constexpr int Ch = L::Ch;
  const Dims d = make_dims(a->d);
  const Saved sv = saved_layout(d.B, C, d.HW);
  const float* S = a->saved;
  const float4* stat = reinterpret_cast<const float4*>(S + sv.stat);
  const BwdWs w = bwd_ws(a->d);
  char* ws = (char*)a->workspace;
  float* part = (float*)(ws + w.part);
  float2* bstat = (float2*)(ws + w.bstat);
  float* ex = (float*)(ws + w.ex);
  float *g2 = (float*)(ws + w.g2), *g4 = (float*)(ws + w.g4), *dr = (float*)(ws + w.dr), *g1 = (float*)(ws + w.g1);
  float *dxa = (float*)(ws + w.dxa), *ts = (float*)(ws + w.ts), *tk = (float*)(ws + w.tk), *wpart = (float*)(ws + w.wpart);
  const int training = a->d.training;
  const float npx = (float)d.NPIX;
  const size_t sw = (size_t)8 * L::NCHAN * 2 * 4;
  {
    out_bwd_stats_kernel<C><<<d.nblk, NT, sw, st>>>(a->dout, S + sv.z, stat, part, d);
    KMU_LAUNCH_CHECK("dagem_out_bwd_stats");
    fin_bwd_kernel<<<C, 128, 0, st>>>(part, d.nblk, bstat, L::ST_Z, C, npx, a->d_bn_weight[4], a->d_bn_bias[4], 0, 0, 1.f, nullptr,
                                      nullptr, training);
    KMU_LAUNCH_CHECK("dagem_fin_bwd(4)");
  }
  {
    size_t smem = ((size_t)C * PAD + (size_t)(C + Ch) * PAD + (size_t)C * (C + Ch)) * 4 + sw;
    opt_in(fuse_bwd_kernel<C>, smem);
    fuse_bwd_kernel<C><<<d.nblk, NT, smem, st>>>(a->dout, S + sv.z, a->deformed, S + sv.uvp, S + sv.r, a->wf, stat, bstat,
                                                 a->d_deformed, g2, g4, wpart, part, d);
    KMU_LAUNCH_CHECK("dagem_fuse_bwd");
    int per = C * (C + Ch);
    wreduce_kernel<<<cdiv(per, 128), 128, 0, st>>>(wpart, d.nblk, per, C, C + Ch, C + Ch, a->d_wf, nullptr, nullptr);
    KMU_LAUNCH_CHECK("dagem_wreduce(f)");
    fin_bwd_kernel<<<Ch + 1, 128, 0, st>>>(part, d.nblk, bstat, L::ST_UV, Ch, npx, a->d_bn_weight[2], a->d_bn_bias[2], L::ST_R, 1,
                                           npx * Ch, a->d_bn_weight[3], a->d_bn_bias[3], training);
    KMU_LAUNCH_CHECK("dagem_fin_bwd(2,3)");
  }
  {
    size_t smem = ((size_t)Ch * PAD + (size_t)(2 * C + 1) * PAD + (size_t)Ch * 2 * C) * 4 + sw;
    opt_in(gate_bwd_kernel<C>, smem);
    gate_bwd_kernel<C><<<d.nblk, NT, smem, st>>>(a->x, S + sv.s, S + sv.ue, S + sv.uvp, S + sv.r, g2, g4, a->vu_w, a->er_w, stat,
                                                 bstat, dr, g1, dxa, wpart, part, d);
    KMU_LAUNCH_CHECK("dagem_gate_bwd");
    int per = Ch * (2 * C + 1);
    wreduce_kernel<<<cdiv(per, 128), 128, 0, st>>>(wpart, d.nblk, per, Ch, 2 * C + 1, 2 * C, a->d_vu_w, a->d_vu_b, nullptr);
    KMU_LAUNCH_CHECK("dagem_wreduce(v)");
    fin_bwd_kernel<<<1 + Ch, 128, 0, st>>>(part, d.nblk, bstat, L::ST_S, 1, npx * C, a->d_bn_weight[0], a->d_bn_bias[0], L::ST_UE,
                                           Ch, npx * 4, a->d_bn_weight[1], a->d_bn_bias[1], training);
    KMU_LAUNCH_CHECK("dagem_fin_bwd(0,1)");
  }
  {
    size_t smem = ((size_t)Ch * PAD + (size_t)2 * C * PAD + (size_t)Ch * 2 * C) * 4 + sw;
    opt_in(edge_bwd_kernel<C>, smem);
    edge_bwd_kernel<C><<<d.nblk, NT, smem, st>>>(a->x, S + sv.s, S + sv.ue, dr, g1, a->ea_w, a->eu_w, a->er_w, stat, bstat, dxa, ts,
                                                 tk, wpart, part, d);
    KMU_LAUNCH_CHECK("dagem_edge_bwd");
    int per = Ch * 2 * C + Ch;
    wreduce_kernel<<<cdiv(per, 128), 128, 0, st>>>(wpart, d.nblk, per, Ch, 2 * C, 2 * C, a->d_eu_w, nullptr, a->d_eu_b);
    KMU_LAUNCH_CHECK("dagem_wreduce(e)");
    fin_sum_kernel<<<NEX, 128, 0, st>>>(part, d.nblk, L::EX, ex);
    KMU_LAUNCH_CHECK("dagem_fin_sum");
  }
  {
    long long total = (long long)d.NPIX * C;
    dx_gather_kernel<<<cdiv(total, 256), 256, 0, st>>>(dxa, ts, tk, a->ea_w, C, d, total, a->dx);
    KMU_LAUNCH_CHECK("dagem_dx_gather");
  }
  // ex = [d r_w0, d r_w1, d r_w2, d r_w3, d r_b, -, d a_w0, d a_w1, d a_w2, d a_w3, d a_b, -]
  cudaMemcpyAsync(a->d_er_w, ex, 4 * sizeof(float), cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(a->d_er_b, ex + 4, sizeof(float), cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(a->d_ea_w, ex + 6, 4 * sizeof(float), cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(a->d_ea_b, ex + 10, sizeof(float), cudaMemcpyDeviceToDevice, st);
  count_launches(4);
  return KMU_OK;
}

}  // namespace dagem
}  // namespace kmu

using namespace kmu;
using namespace kmu::dagem;

extern "C" {

size_t kmu_dagem_saved_bytes(const kmu_dagem_desc* d) {
  if (check(d, "dagem_saved_bytes") != KMU_OK) return 0;
  return saved_layout(d->B, d->C, d->H * d->W).total * sizeof(float);
}
size_t kmu_dagem_fwd_workspace_bytes(const kmu_dagem_desc* d) {
  if (check(d, "dagem_fwd_workspace_bytes") != KMU_OK) return 0;
  return fwd_ws(*d).total;
}
size_t kmu_dagem_bwd_workspace_bytes(const kmu_dagem_desc* d) {
  if (check(d, "dagem_bwd_workspace_bytes") != KMU_OK) return 0;
  return bwd_ws(*d).total;
}

int kmu_dagem_fwd(const kmu_dagem_fwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "dagem_fwd: null args");
  int rc = check(&a->d, "dagem_fwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(a->x && a->deformed && a->ea_w && a->ea_b && a->vu_w && a->vu_b && a->eu_w && a->eu_b && a->er_w && a->er_b && a->wf &&
                  a->out && a->saved,
              KMU_ERR_BAD_ARG, "dagem_fwd: null tensor");
  for (int i = 0; i < 5; ++i) {
    KMU_REQUIRE(a->bn[i].weight && a->bn[i].bias, KMU_ERR_BAD_ARG, "dagem_fwd: BatchNorm %d has no affine parameters", i);
    KMU_REQUIRE(a->d.training || (a->bn[i].running_mean && a->bn[i].running_var), KMU_ERR_BAD_ARG,
                "dagem_fwd: eval mode needs the running statistics of BatchNorm %d", i);
  }
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= fwd_ws(a->d).total, KMU_ERR_WORKSPACE, "dagem_fwd: workspace %zu < %zu",
              a->workspace_bytes, fwd_ws(a->d).total);
  cudaStream_t st = (cudaStream_t)stream;
  switch (a->d.C) {
    case 8: return forward<8>(a, st);
    case 16: return forward<16>(a, st);
    case 32: return forward<32>(a, st);
    case 64: return forward<64>(a, st);
  }
  return KMU_ERR_UNSUPPORTED;
}

int kmu_dagem_bwd(const kmu_dagem_bwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "dagem_bwd: null args");
  int rc = check(&a->d, "dagem_bwd");
  if (rc != KMU_OK) return rc;
  KMU_REQUIRE(a->x && a->deformed && a->dout && a->saved && a->ea_w && a->vu_w && a->eu_w && a->er_w && a->wf, KMU_ERR_BAD_ARG,
              "dagem_bwd: null input tensor");
  KMU_REQUIRE(a->dx && a->d_deformed && a->d_ea_w && a->d_ea_b && a->d_vu_w && a->d_vu_b && a->d_eu_w && a->d_eu_b && a->d_er_w &&
                  a->d_er_b && a->d_wf,
              KMU_ERR_BAD_ARG, "dagem_bwd: null output tensor");
  for (int i = 0; i < 5; ++i)
    KMU_REQUIRE(a->d_bn_weight[i] && a->d_bn_bias[i], KMU_ERR_BAD_ARG, "dagem_bwd: null BatchNorm gradient %d", i);
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= bwd_ws(a->d).total, KMU_ERR_WORKSPACE, "dagem_bwd: workspace %zu < %zu",
              a->workspace_bytes, bwd_ws(a->d).total);
  cudaStream_t st = (cudaStream_t)stream;
  switch (a->d.C) {
    case 8: return backward<8>(a, st);
    case 16: return backward<16>(a, st);
    case 32: return backward<32>(a, st);
    case 64: return backward<64>(a, st);
  }
  return KMU_ERR_UNSUPPORTED;
}

}  // extern "C"
