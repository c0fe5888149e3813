// kan_tc_bwd.cu -- tcgen05 / TMEM family of the KANConv2d BACKWARD (KMU_PREC_BF16), 3x3 / stride 1 / padding 1.
//
// Replaces the autograd graph PyTorch builds for convKAN/KANConv2Dlayers.py:15-37 + convKAN/KANlayers.py:577-610,644-660.
// With Phi(x) = [SiLU(x), B_0(x)..B_7(x)] per input pixel (see kan_tc.cu) the layer is Y = conv3x3(Phi(X), Wfull), so
//     dPhi[p, c, q]   = sum_{tap, o} dY[p + 1 - tap, o] * Wfull[o, c, tap, q]          dX[p, c] = sum_q dPhi[p,c,q] Phi'_q(x[p,c])
//     dWfull[o,c,tap,q] = sum_p dY[p, o] * Phi_q(x[c, p - 1 + tap])                     (Phi(0) at the zero-padded border)
// Two persistent warp-specialised kernels, one tensor-core GEMM each, bf16 operands / fp32 TMEM accumulators:
//
//  kan_bwd_dx_tc_kernel   WEIGHT-STATIONARY.  A CTA owns one block of 16 input channels: its slice of the transposed
//     weights (9 shifts x Cout x 144 columns, <= 166 KB bf16) is bulk-copied into shared memory ONCE and stays there.
//     The CTA then walks 8x16-pixel tiles: stager warps convert the dY halo tile to bf16 K-major planes (the nine taps
//     are nine shifted views of the same planes, as in the forward), one thread issues 9*Cout/16 MMAs of
//     M=128 pixels x N=144 (16 channels x 9 Phi components) x K=16, and eight epilogue warps read dPhi from TMEM,
//     contract it with the closed-form Phi'(x) (SiLU' and the four non-zero cubic B-spline derivatives, placed with a
//     select network -- no shared memory, the MMA already uses most of its bandwidth) and store dX.  dPhi never exists
//     in HBM.  Accumulators are double buffered (2 x 144 TMEM columns) so tile t's epilogue overlaps tile t+1's MMAs.
//
//  kan_bwd_dw_tc_kernel   OUTPUT-STATIONARY over a split of the pixels.  The reduction dimension is the PIXEL, so both
//     operands are MN-major: A = dY rows (tap-row slot, o) read from [row][o-group][8 px][8 o] planes where "next tap
//     row" is just 'Cout/8 core matrices further' (one descriptor covers 128/Cout vertical taps), B = Phi planes
//     (8 spline values per channel + the SiLU plane) produced on the fly exactly as in the forward, the kj shift being
//     a 16-byte start offset.  A CTA owns (channel block, pixel split) and keeps ALL its dWfull partial sums
//     (128 rows x up to 480 columns) in TMEM for its whole life; one epilogue writes them to the workspace and
//     kan_bwd_dw_tc_reduce_kernel adds the splits in a fixed order (deterministic) and applies the chain rule into
//     base_weight / spline_weight / spline_scaler gradients (convKAN/KANlayers.py:644-650).
#include "common.cuh"
#include "kan_common.cuh"
#include "tc_common.cuh"

namespace kmu {
namespace kan {
namespace tc {

using namespace kmu::tcx;

constexpr int BP = 10;                 // halo pitch: 8-wide strip + 1 pixel each side
constexpr int BR = 16;                 // tile rows
constexpr int BNPOS = BP * (BR + 2);   // 180 halo positions
constexpr int BPLANE = BNPOS * 16;     // bytes of one K-group plane (16 B per position)

static int g_debug_flags = 0;          // bit0: swap LBO/SBO of the MN-major descriptors (bring-up aid)
void set_debug_flags(int f) { g_debug_flags = f; }

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// ================================================================================================ dX
constexpr int DX_N = 144;              // 16 channels x 9 Phi components
constexpr int DX_STAGERS = 8, DX_WARP_MMA = 8, DX_WARP_EPI0 = 9, DX_EPI_WARPS = 8;
constexpr int DX_THREADS = (DX_WARP_EPI0 + DX_EPI_WARPS) * 32;

struct BwdDims {
  int B, Cin, H, W;
  int tiles_x, tiles_y, num_tiles;
  int nblk;        // channel blocks
  int splits;      // CTAs per channel block
  float t0, inv_h;
};

// w2pack[cb][shift][ks][gi][n][e] (bf16): shared-memory image of the B operand of (shift, K-step): K-major no-swizzle,
// row n = c_local*9 + q, 16 B = 8 consecutive output channels o = ks*16 + gi*8 + e.  shift (si,sj) reads dY at halo
// offset (si,sj), i.e. it carries the weights of tap (2-si, 2-sj).
__global__ void kan_tc_pack_dx_kernel(const float* __restrict__ base_w, const float* __restrict__ spline_w,
                                      const float* __restrict__ scaler, __nv_bfloat16* __restrict__ w2pack, int Cin, int Cout) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)81 * Cin * Cout;
  if (idx >= total) return;
  const int KS = Cout / 16;
  int e = (int)(idx & 7);
  long long r = idx >> 3;
  int n = (int)(r % DX_N);
  r /= DX_N;
  int gi = (int)(r & 1);
  r >>= 1;
  int ks = (int)(r % KS);
  r /= KS;
  int t = (int)(r % 9);
  int cb = (int)(r / 9);
  int o = ks * 16 + gi * 8 + e;
  int c = cb * 16 + n / 9, q = n % 9;
  int si = t / 3, sj = t - si * 3;
  int tap = (2 - si) * 3 + (2 - sj);
  size_t of = (size_t)o * (Cin * 9) + c * 9 + tap;
  float v = q == 0 ? base_w[of] : spline_w[of * NB + (q - 1)] * (scaler ? scaler[of] : 1.0f);
  w2pack[idx] = __float2bfloat16_rn(v);
}

// dX[p,c] from the nine dPhi values of (p,c): SiLU'(x) v0 + sum_k d_k v[1 + i-3+k] with i the knot span of x and d the
// four non-zero cubic B-spline derivatives [-3(1-u)^2, 9u^2-12u, -9u^2+6u+3, 3u^2] / (6h)   (SURVEY appendix A.1).
// The data-dependent 4-of-8 pick is a 4-level select network on the bits of i (registers only).
__device__ __forceinline__ float dx_contract(float x, const float* v, float t0, float inv_h) {
  const float sg = __fdividef(1.0f, 1.0f + __expf(-x));
  float acc = sg * (1.0f + x * (1.0f - sg)) * v[0];
  const float s = (x - t0) * inv_h;
  const bool in = s >= 0.f && s < 11.f;
  const float fi = floorf(s);
  const float u = s - fi;
  const int idx = in ? (int)fi : 0;
  const float sc = in ? inv_h : 0.f;
  const float om = 1.f - u, u2 = u * u;
  const float d0 = -0.5f * om * om * sc;
  const float d1 = (1.5f * u2 - 2.f * u) * sc;
  const float d2 = (-1.5f * u2 + u + 0.5f) * sc;
  const float d3 = 0.5f * u2 * sc;
  const bool b3 = (idx & 8) != 0, b2 = (idx & 4) != 0, b1 = (idx & 2) != 0, b0 = (idx & 1) != 0;
  // P[i] = v[i-2] for i in [3,10], zero elsewhere (i in [0,13]); wanted: P[idx + k], k = 0..3
  float P[14];
#pragma unroll
  for (int i = 0; i < 14; ++i) P[i] = (i >= 3 && i <= 10) ? v[i - 2] : 0.f;
  float Q[11];
#pragma unroll
  for (int i = 0; i < 11; ++i) Q[i] = b3 ? ((i + 8 < 14) ? P[i + 8] : 0.f) : P[i];
  float U[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) U[i] = b2 ? Q[i + 4] : Q[i];
  float T[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) T[i] = b1 ? U[i + 2] : U[i];
  const float g0 = b0 ? T[1] : T[0], g1 = b0 ? T[2] : T[1], g2 = b0 ? T[3] : T[2], g3 = b0 ? T[4] : T[3];
  acc = fmaf(d0, g0, acc);
  acc = fmaf(d1, g1, acc);
  acc = fmaf(d2, g2, acc);
  acc = fmaf(d3, g3, acc);
  return acc;
}

template <int COUT>
__global__ void __launch_bounds__(DX_THREADS, 1) kan_bwd_dx_tc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                       const __nv_bfloat16* __restrict__ w2pack,
                                                                       float* __restrict__ dx, BwdDims d) {
  constexpr int OG = COUT / 8, KS = COUT / 16;
  constexpr int WBLK = 2 * DX_N * 16;        // bytes of one (shift, K-step) B block
  constexpr int WBYTES = 9 * KS * WBLK;
  constexpr int DYSTAGE = OG * BPLANE;
  constexpr uint32_t IDESC = make_idesc_bf16(128, DX_N);
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* w_base = smem;
  uint8_t* dy_base = smem + WBYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dy_base + 2 * DYSTAGE);
  uint64_t* w_full = bars;
  uint64_t* dy_full = bars + 1;
  uint64_t* dy_empty = bars + 3;
  uint64_t* acc_full = bars + 5;
  uint64_t* acc_empty = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(smem_u32(w_full), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&dy_full[i]), DX_STAGERS);
      mbar_init(smem_u32(&dy_empty[i]), 1);
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), DX_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == DX_WARP_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const size_t HW = (size_t)d.H * d.W;
  const int tiles_per_img = d.tiles_x * d.tiles_y;
  const int cb = blockIdx.x % d.nblk;
  const int first = blockIdx.x / d.nblk, step = d.splits;

  if (warp < DX_STAGERS) {
    // ===================================================================== dY stagers: fp32 NCHW -> bf16 [og][pos][8 o]
    // Every thread owns UPT (o-group, halo position) units; the 8*UPT global loads of the NEXT tile are issued right after
    // the current tile has been written to shared memory, so their latency hides behind that tile's MMAs.
    constexpr int NST = DX_STAGERS * 32;
    constexpr int UPT = (OG * BNPOS + NST - 1) / NST;
    float v[UPT][8];
    auto load_tile = [&](int tile) {
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty0 = (tr / d.tiles_x) * BR, tx0 = (tr % d.tiles_x) * 8;
      const float* dyb = dy + (size_t)b * COUT * HW;
#pragma unroll
      for (int k = 0; k < UPT; ++k) {
        const int u = tid + k * NST;
        const int og = u / BNPOS, pos = u - og * BNPOS;
        const int py = pos / BP, px = pos - py * BP;
        const int gy = ty0 - 1 + py, gx = tx0 - 1 + px;
        const bool ok = u < OG * BNPOS && gy >= 0 && gy < d.H && gx >= 0 && gx < d.W;
        const float* p = dyb + (size_t)(og * 8) * HW + (size_t)gy * d.W + gx;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[k][e] = ok ? __ldg(p + (size_t)e * HW) : 0.f;
      }
    };
    uint32_t it = 0;
    if (first < d.num_tiles) load_tile(first);
    for (int tile = first; tile < d.num_tiles; tile += step, ++it) {
      const uint32_t s = it & 1u;
      mbar_wait(smem_u32(&dy_empty[s]), ((it >> 1) & 1u) ^ 1u);
      uint8_t* stage = dy_base + (size_t)s * DYSTAGE;
#pragma unroll
      for (int k = 0; k < UPT; ++k) {
        const int u = tid + k * NST;
        if (u < OG * BNPOS) {
          const int og = u / BNPOS, pos = u - og * BNPOS;
          *reinterpret_cast<uint4*>(stage + (size_t)og * BPLANE + (size_t)pos * 16) = pack8_bf16(v[k]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&dy_full[s]));
      if (tile + step < d.num_tiles) load_tile(tile + step);
    }
  } else if (warp == DX_WARP_MMA) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      {  // one-shot weight load: this channel block's transposed weights stay resident for the CTA's whole life
        const uint32_t bar = smem_u32(w_full);
        mbar_expect_tx(bar, WBYTES);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(w2pack) + (size_t)cb * WBYTES;
#pragma unroll 1
        for (int t = 0; t < 9; ++t) bulk_g2s(smem_u32(w_base + (size_t)t * KS * WBLK), src + (size_t)t * KS * WBLK, KS * WBLK, bar);
      }
      mbar_wait_hot(smem_u32(w_full), 0);
      uint32_t it = 0;
      const uint32_t w0 = smem_u32(w_base);
      for (int tile = first; tile < d.num_tiles; tile += step, ++it) {
        const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
        mbar_wait_hot(smem_u32(&acc_empty[s]), ph ^ 1u);
        mbar_wait_hot(smem_u32(&dy_full[s]), ph);
        tc_fence_after();
        const uint64_t adesc0 = make_smem_desc(smem_u32(dy_base + (size_t)s * DYSTAGE), BPLANE, BP * 16);
        const uint64_t bdesc0 = make_smem_desc(w0, DX_N * 16, 128);
        const uint32_t d_tmem = tmem_base + s * 256u;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int si = t / 3, sj = t - si * 3;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t adesc = desc_advance(adesc0, (uint32_t)(ks * 2 * BPLANE + (si * BP + sj) * 16));
            const uint64_t bdesc = desc_advance(bdesc0, (uint32_t)((t * KS + ks) * WBLK));
            umma_bf16(d_tmem, adesc, bdesc, IDESC, (t > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&dy_empty[s]));
        umma_commit(smem_u32(&acc_full[s]));
      }
    }
  } else {
    // ===================================================================== epilogue: dPhi (TMEM) x Phi'(x) -> dX
    const int ew = warp - DX_WARP_EPI0;
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int half = ew >> 2;          // which 8 of the block's 16 channels
    uint32_t it = 0;
    for (int tile = first; tile < d.num_tiles; tile += step, ++it) {
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty0 = (tr / d.tiles_x) * BR, tx0 = (tr % d.tiles_x) * 8;
      const int m = q * 32 + lane;
      const int gy = ty0 + (m >> 3), gx = tx0 + (m & 7);
      const bool ok = gy < d.H && gx < d.W;
      const size_t off = ((size_t)b * d.Cin + cb * 16 + half * 8) * HW + (size_t)gy * d.W + gx;
      float xv[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) xv[c] = ok ? __ldg(x + off + (size_t)c * HW) : 0.f;
      const uint32_t s = it & 1u;
      mbar_wait(smem_u32(&acc_full[s]), (it >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + s * 256u + (uint32_t)(half * 72);
      uint32_t r[72];
#pragma unroll
      for (int g = 0; g < 9; ++g) tmem_ld8(taddr + (uint32_t)(g * 8), r + g * 8);   // 8-column aligned loads
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = __uint_as_float(r[c * 9 + i]);
        const float g = dx_contract(xv[c], v, d.t0, d.inv_h);
        if (ok) dx[off + (size_t)c * HW] = g;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[s]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == DX_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================ dW
constexpr int DW_PRODUCERS = 8, DW_STAGERS = 4, DW_WARP_MMA = 12;
constexpr int DW_THREADS = 13 * 32;
constexpr int DW_STAGES = 3;

template <int COUT, int CH>
struct DwCfg {
  static constexpr int OG = COUT / 8;
  static constexpr int SLOTS = 128 / COUT;                       // vertical taps one A descriptor covers
  static constexpr int SETS = (3 + SLOTS - 1) / SLOTS;           // accumulator sets (tap rows 2,1 | 0,pad for Cout=64)
  static constexpr int NPL = ((CH + CH / 8 + 1) / 2) * 2;        // B planes: CH spline + CH/8 SiLU (+1 zero pad -> N % 16 == 0)
  static constexpr int N = NPL * 8;
  static constexpr int NACC = SETS * 3;
  static constexpr int PHI_STAGE = NPL * BPLANE;
  static constexpr int DY_ROWB = OG * 128;                       // one output row: OG core matrices of 8 px x 8 o
  static constexpr int DY_ROWS = BR + 2 + SETS * SLOTS - 1;      // smem row ry = (output row - ty0) + 2; rows outside [2, BR+2) stay 0
  static constexpr int DY_STAGE = DY_ROWS * DY_ROWB;
  static constexpr int STAGE = PHI_STAGE + DY_STAGE;
  static_assert(NACC * N <= 512, "dW accumulators exceed TMEM");
  static_assert(N % 16 == 0 && N <= 256, "bad UMMA N");
};

template <int COUT, int CH>
__global__ void __launch_bounds__(DW_THREADS, 1) kan_bwd_dw_tc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                       float* __restrict__ partial, BwdDims d, int mn_swap) {
  using C = DwCfg<COUT, CH>;
  constexpr uint32_t IDESC = make_idesc_bf16_mn(128, C::N);
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)DW_STAGES * C::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + DW_STAGES;
  uint64_t* done = bars + 2 * DW_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DW_STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // zero every stage once: the dY halo rows, the pad plane and the SiLU lanes of absent channels are never written again
  for (int i = tid; i < DW_STAGES * C::STAGE / 16; i += DW_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < DW_STAGES; ++i) {
      mbar_init(smem_u32(&full[i]), DW_PRODUCERS + DW_STAGERS);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    mbar_init(smem_u32(done), 1);
    fence_barrier_init();
  }
  if (warp == DW_WARP_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const size_t HW = (size_t)d.H * d.W;
  const int tiles_per_img = d.tiles_x * d.tiles_y;
  const int cb = blockIdx.x % d.nblk;
  const int split = blockIdx.x / d.nblk, step = d.splits;
  const bool has_work = split < d.num_tiles;

  if (warp < DW_PRODUCERS) {
    // ===================================================================== Phi producers (channels cb*CH .. +CH)
    // x of the NEXT tile is fetched into registers while the current one is evaluated (latency behind the math).
    constexpr int NPR = DW_PRODUCERS * 32;
    constexpr int UPT = (CH * BNPOS + NPR - 1) / NPR;
    float xr[UPT];
    auto load_tile = [&](int tile) {
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty0 = (tr / d.tiles_x) * BR, tx0 = (tr % d.tiles_x) * 8;
      const float* xb = x + ((size_t)b * d.Cin + cb * CH) * HW;
#pragma unroll
      for (int k = 0; k < UPT; ++k) {
        const int u = tid + k * NPR;
        const int cl = u / BNPOS, pos = u - cl * BNPOS;
        const int py = pos / BP, px = pos - py * BP;
        const int gy = ty0 - 1 + py, gx = tx0 - 1 + px;
        const bool ok = u < CH * BNPOS && gy >= 0 && gy < d.H && gx >= 0 && gx < d.W;
        xr[k] = ok ? __ldg(xb + (size_t)cl * HW + (size_t)gy * d.W + gx) : 0.f;
      }
    };
    uint32_t it = 0;
    if (split < d.num_tiles) load_tile(split);
    for (int tile = split; tile < d.num_tiles; tile += step, ++it) {
      float xc[UPT];
#pragma unroll
      for (int k = 0; k < UPT; ++k) xc[k] = xr[k];
      if (tile + step < d.num_tiles) load_tile(tile + step);
      const int st = it % DW_STAGES;
      mbar_wait(smem_u32(&empty[st]), ((it / DW_STAGES) & 1u) ^ 1u);
      uint8_t* stage = smem + (size_t)st * C::STAGE;
#pragma unroll
      for (int k = 0; k < UPT; ++k) {
        const int u = tid + k * NPR;
        if (u < CH * BNPOS) {
          const int cl = u / BNPOS, pos = u - cl * BNPOS;
          *reinterpret_cast<uint4*>(stage + (size_t)cl * BPLANE + (size_t)pos * 16) = spline_group(xc[k], d.t0, d.inv_h);
          *reinterpret_cast<__nv_bfloat16*>(stage + (size_t)(CH + (cl >> 3)) * BPLANE + (size_t)pos * 16 + (cl & 7) * 2) =
              __float2bfloat16_rn(silu_fast(xc[k]));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&full[st]));
    }
  } else if (warp < DW_PRODUCERS + DW_STAGERS) {
    // ===================================================================== dY stagers: fp32 NCHW -> bf16 [row][og][8 px][8 o]
    constexpr int NST = DW_STAGERS * 32;
    constexpr int UNITS = BR * C::OG * 8;
    constexpr int UPT = (UNITS + NST - 1) / NST;
    const int stid = tid - DW_PRODUCERS * 32;
    float v[UPT][8];
    auto load_tile = [&](int tile) {
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty0 = (tr / d.tiles_x) * BR, tx0 = (tr % d.tiles_x) * 8;
      const float* dyb = dy + (size_t)b * COUT * HW;
#pragma unroll
      for (int k = 0; k < UPT; ++k) {
        const int u = stid + k * NST;
        const int cx = u & 7, r = (u >> 3) % BR, og = u / (8 * BR);
        const int gy = ty0 + r, gx = tx0 + cx;
        const bool ok = u < UNITS && gy < d.H && gx < d.W;
        const float* p = dyb + (size_t)(og * 8) * HW + (size_t)gy * d.W + gx;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[k][e] = ok ? __ldg(p + (size_t)e * HW) : 0.f;
      }
    };
    uint32_t it = 0;
    if (split < d.num_tiles) load_tile(split);
    for (int tile = split; tile < d.num_tiles; tile += step, ++it) {
      const int st = it % DW_STAGES;
      mbar_wait(smem_u32(&empty[st]), ((it / DW_STAGES) & 1u) ^ 1u);
      uint8_t* stage = smem + (size_t)st * C::STAGE + C::PHI_STAGE;
#pragma unroll
      for (int k = 0; k < UPT; ++k) {
        const int u = stid + k * NST;
        if (u < UNITS) {
          const int cx = u & 7, r = (u >> 3) % BR, og = u / (8 * BR);
          *reinterpret_cast<uint4*>(stage + (size_t)(r + 2) * C::DY_ROWB + (size_t)og * 128 + (size_t)cx * 16) = pack8_bf16(v[k]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&full[st]));
      if (tile + step < d.num_tiles) load_tile(tile + step);
    }
  } else {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = split; tile < d.num_tiles; tile += step, ++it) {
        const int st = it % DW_STAGES;
        mbar_wait_hot(smem_u32(&full[st]), (it / DW_STAGES) & 1u);
        tc_fence_after();
        const uint32_t phi0 = smem_u32(smem + (size_t)st * C::STAGE);
        const uint32_t dy0 = phi0 + C::PHI_STAGE;
        const uint64_t bdesc0 = mn_swap ? make_smem_desc(phi0, BPLANE, BP * 16) : make_smem_desc(phi0, BP * 16, BPLANE);
        const uint64_t adesc0 = mn_swap ? make_smem_desc(dy0, 128, C::DY_ROWB) : make_smem_desc(dy0, C::DY_ROWB, 128);
#pragma unroll 1
        for (int hk = 0; hk < (BR + 2) / 2; ++hk) {
          const int h = 2 * hk;
          const uint64_t bdesc_h = desc_advance(bdesc0, (uint32_t)(h * BP * 16));
          const uint64_t adesc_h = desc_advance(adesc0, (uint32_t)(h * C::DY_ROWB));
#pragma unroll
          for (int kj = 0; kj < 3; ++kj) {
            const uint64_t bdesc = desc_advance(bdesc_h, (uint32_t)(kj * 16));
#pragma unroll
            for (int s = 0; s < C::SETS; ++s) {
              const uint64_t adesc = desc_advance(adesc_h, (uint32_t)(s * C::SLOTS * C::DY_ROWB));
              umma_bf16(tmem_base + (uint32_t)((s * 3 + kj) * C::N), adesc, bdesc, IDESC, (it > 0 || hk > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(smem_u32(&empty[st]));
      }
      umma_commit(smem_u32(done));
    }
  }
  // ======================================================================= epilogue: TMEM -> partial[split][cb][acc][col][row]
  if (warp < 12) {
    if (has_work) {
      mbar_wait(smem_u32(done), 0);
      tc_fence_after();
    }
    const int q = warp & 3;
    float* dst = partial + ((size_t)(split * d.nblk + cb) * C::NACC) * C::N * 128 + q * 32 + lane;
    for (int acc = warp >> 2; acc < C::NACC; acc += 3) {
#pragma unroll 1
      for (int c0 = 0; c0 < C::N; c0 += 16) {
        uint32_t v[16];
        if (has_work) {
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C::N + c0), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 0u;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) dst[((size_t)acc * C::N + c0 + e) * 128] = __uint_as_float(v[e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == DW_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Fixed-order sum over the pixel splits + chain rule (convKAN/KANlayers.py:644-650):
//   d base_weight = dWfull[q=0],  d spline_weight = dWfull[q>=1] * scaler,  d spline_scaler = sum_j dWfull[1+j] * spline_weight[j]
__global__ void kan_bwd_dw_tc_reduce_kernel(const float* __restrict__ partial, int splits, int nblk, int ch, int slots, int nacc,
                                            int n, const float* __restrict__ spline_w, const float* __restrict__ scaler,
                                            float* __restrict__ d_base, float* __restrict__ d_spline,
                                            float* __restrict__ d_scaler, int Cin, int Cout) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int F = Cin * 9;
  if (idx >= F * Cout) return;
  int o = idx % Cout, f = idx / Cout;   // o fastest: consecutive threads read consecutive accumulator rows
  int c = f / 9, tap = f - c * 9;
  int ki = tap / 3, kj = tap - ki * 3;
  int sg = 2 - ki;
  int set = sg / slots, slot = sg - set * slots;
  int row = slot * Cout + o;
  int cb = c / ch, cl = c - cb * ch;
  int acc = set * 3 + kj;
  float s[NPHI];
#pragma unroll
  for (int q = 0; q < NPHI; ++q) s[q] = 0.f;
  for (int sp = 0; sp < splits; ++sp) {
    const float* src = partial + (((size_t)(sp * nblk + cb) * nacc + acc) * n) * 128 + row;
    s[0] += src[(size_t)((ch + (cl >> 3)) * 8 + (cl & 7)) * 128];
#pragma unroll
    for (int j = 0; j < NB; ++j) s[1 + j] += src[(size_t)(cl * 8 + j) * 128];
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

// ================================================================================================ host side
static int num_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

static int dw_ch(const Dims& d) { return d.Cout == 64 ? 8 : 16; }

struct DwPlan {
  int ch, nblk, splits, nacc, n, slots;
  size_t partial_bytes;
};

static DwPlan dw_plan(const Dims& d, int sms) {
  DwPlan p;
  p.ch = dw_ch(d);
  p.nblk = d.Cin / p.ch;
  p.slots = 128 / d.Cout;
  int sets = (3 + p.slots - 1) / p.slots;
  p.nacc = sets * 3;
  int npl = ((p.ch + p.ch / 8 + 1) / 2) * 2;
  p.n = npl * 8;
  long long tiles = (long long)d.B * cdiv(d.W, 8) * cdiv(d.H, BR);
  int s = sms / p.nblk;
  if (s < 1) s = 1;
  if (s > tiles) s = (int)tiles;
  p.splits = s;
  p.partial_bytes = (size_t)p.splits * p.nblk * p.nacc * p.n * 128 * 4;
  return p;
}

size_t bwd_workspace(const Dims& d) {
  // sized for the largest grid any sm_100 part can ask for (<= 160 SMs), so the call never depends on the device
  DwPlan p = dw_plan(d, 160);
  return align_up((size_t)81 * d.Cin * d.Cout * 2, 256) + align_up(p.partial_bytes, 256);
}

template <int COUT>
static int launch_dx(const kmu_kanconv2d_bwd_args* a, const Dims& d, const __nv_bfloat16* w2pack, cudaStream_t st) {
  constexpr int OG = COUT / 8, KS = COUT / 16;
  BwdDims t;
  t.B = d.B; t.Cin = d.Cin; t.H = d.H; t.W = d.W;
  t.tiles_x = cdiv(d.W, 8);
  t.tiles_y = cdiv(d.H, BR);
  t.num_tiles = d.B * t.tiles_x * t.tiles_y;
  t.nblk = d.Cin / 16;
  int s = num_sms() / t.nblk;
  if (s < 1) s = 1;
  if (s > t.num_tiles) s = t.num_tiles;
  t.splits = s;
  t.t0 = a->d.grid_t0;
  t.inv_h = 1.0f / a->d.grid_h;
  size_t smem = (size_t)9 * KS * 2 * DX_N * 16 + 2 * (size_t)OG * BPLANE + 9 * 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(kan_bwd_dx_tc_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "kan_bwd_dx_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
  kan_bwd_dx_tc_kernel<COUT><<<t.splits * t.nblk, DX_THREADS, smem, st>>>(a->x, a->dy, w2pack, a->dx, t);
  KMU_LAUNCH_CHECK("kan_bwd_dx_tc");
  return KMU_OK;
}

template <int COUT, int CH>
static int launch_dw(const kmu_kanconv2d_bwd_args* a, const Dims& d, float* partial, const DwPlan& p, cudaStream_t st) {
  using C = DwCfg<COUT, CH>;
  BwdDims t;
  t.B = d.B; t.Cin = d.Cin; t.H = d.H; t.W = d.W;
  t.tiles_x = cdiv(d.W, 8);
  t.tiles_y = cdiv(d.H, BR);
  t.num_tiles = d.B * t.tiles_x * t.tiles_y;
  t.nblk = p.nblk;
  t.splits = p.splits;
  t.t0 = a->d.grid_t0;
  t.inv_h = 1.0f / a->d.grid_h;
  size_t smem = (size_t)DW_STAGES * C::STAGE + (2 * DW_STAGES + 1) * 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(kan_bwd_dw_tc_kernel<COUT, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "kan_bwd_dw_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
  kan_bwd_dw_tc_kernel<COUT, CH><<<p.splits * p.nblk, DW_THREADS, smem, st>>>(a->x, a->dy, partial, t, g_debug_flags & 1);
  KMU_LAUNCH_CHECK("kan_bwd_dw_tc");
  return KMU_OK;
}

int backward(const kmu_kanconv2d_bwd_args* a, const Dims& d, cudaStream_t st) {
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= bwd_workspace(d), KMU_ERR_WORKSPACE, "kanconv2d_bwd(tc): workspace %zu < %zu",
              a->workspace_bytes, bwd_workspace(d));
  KMU_REQUIRE(a->d.grid_h > 0.f, KMU_ERR_BAD_ARG, "kanconv2d_bwd(tc): grid_h must be positive");
  char* ws = (char*)a->workspace;
  __nv_bfloat16* w2pack = (__nv_bfloat16*)ws;
  float* partial = (float*)(ws + align_up((size_t)81 * d.Cin * d.Cout * 2, 256));
  int rc = KMU_OK;
  if (a->dx) {
    long long total = (long long)81 * d.Cin * d.Cout;
    kan_tc_pack_dx_kernel<<<cdiv(total, 256), 256, 0, st>>>(a->base_weight, a->spline_weight, a->spline_scaler, w2pack, d.Cin, d.Cout);
    KMU_LAUNCH_CHECK("kan_tc_pack_dx");
    switch (d.Cout) {
      case 16: rc = launch_dx<16>(a, d, w2pack, st); break;
      case 32: rc = launch_dx<32>(a, d, w2pack, st); break;
      case 64: rc = launch_dx<64>(a, d, w2pack, st); break;
      default: set_error("kanconv2d_bwd(tc): unsupported Cout %d", d.Cout); return KMU_ERR_UNSUPPORTED;
    }
    if (rc != KMU_OK) return rc;
  }
  if (a->d_base_weight) {
    DwPlan p = dw_plan(d, num_sms());
    switch (d.Cout) {
      case 16: rc = launch_dw<16, 16>(a, d, partial, p, st); break;
      case 32: rc = launch_dw<32, 16>(a, d, partial, p, st); break;
      case 64: rc = launch_dw<64, 8>(a, d, partial, p, st); break;
      default: set_error("kanconv2d_bwd(tc): unsupported Cout %d", d.Cout); return KMU_ERR_UNSUPPORTED;
    }
    if (rc != KMU_OK) return rc;
    int n = d.F * d.Cout;
    kan_bwd_dw_tc_reduce_kernel<<<cdiv(n, 256), 256, 0, st>>>(partial, p.splits, p.nblk, p.ch, p.slots, p.nacc, p.n, a->spline_weight,
                                                              a->spline_scaler, a->d_base_weight, a->d_spline_weight,
                                                              a->d_spline_scaler, d.Cin, d.Cout);
    KMU_LAUNCH_CHECK("kan_bwd_dw_tc_reduce");
  }
  return KMU_OK;
}

}  // namespace tc
}  // namespace kan
}  // namespace kmu
