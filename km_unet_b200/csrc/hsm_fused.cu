// hsm_fused.cu -- HSM-SSD sweeps over L with the projected tensor P = dw3x3(Wp x) living only in TMEM (KMU_PREC_BF16 family).
//
// vim_block_init/efficient_vim_init.py:39-52:  BCdt = dw(BCdt_proj(x)); A = softmax_L(dt); h = x @ (A * B)^T ...
// The round-1 tensor-core path wrote the Bm / dt slices of P (B,192,L fp32) to HBM in one kernel and re-read them in the softmax
// sweep (forward) and in the dP kernel (backward): 8.5x the algorithmic traffic of the mixer.  Here the same merged 3x3 convolution
// (hsm_tc.cu: nine descriptor offsets into bf16 K-group planes of the x tile) leaves Bm | dt as 128 fp32 TMEM columns and the
// consumer runs in the epilogue of the same CTA:
//
//   forward   hsm_hs_kernel<C>   per 8x16-pixel tile: D1[px, Bm|dt] = conv3x3(x)                                   (tcgen05, N = 128)
//                                 m_t[n] = max_px dt ; Ea = exp(dt - m_t) ; Eb = Ea * Bm      -> bf16 MN-major operand image in smem
//                                 D2[(Eb | Ea) rows, (x channels | ones)] = E^T x_tile                               (tcgen05, K = 128 pixels)
//                                 -> per-tile partials (m_t, sum Ea, x (Ea Bm)^T): the split-softmax that hsm_combine_gate merges.
//   backward  hsm_dp_kernel<C>   D1 = conv3x3(x) again (bit-identical to the forward's), D3[px, dG | dCm] = x^T dhs | dy^T ho   (K = C)
//                                 A = exp(dt - m) / s ; dBm = dG A ; ddt = A (dG Bm - r) ; dCm            -> dP planes (bf16) to HBM
//                                 D4[px, c] = (A Bm) dhs^T   (K = 64): the direct part of dx, through a K-major smem image.
// P is never written; the forward reads x once (+ halo) and writes 68 floats per tile and state column.
// CTA = 256 threads: warps w and w + 4 share TMEM lane quarter w & 3 (lane m <-> pixel (m >> 3, m & 7) of the tile) and split the
// columns between them; persistent over tiles, weights resident.
#include "common.cuh"
#include "tc_common.cuh"

namespace kmu {
namespace hsm {
namespace fz {

using namespace kmu::tcx;

constexpr int PITCH = 10, ROWS = 18, NPOS = PITCH * ROWS;  // 8x16 tile + 1-pixel halo
constexpr int PLANE = NPOS * 16;                           // bytes of one K-group plane (8 channels x bf16 per position)
constexpr int WBLK = 2 * 128 * 16;                         // bytes of one (tap, k-step) weight block [gi][n 128][8]
constexpr int EGRP = 128 * 16;
constexpr int NT = 256;                                    // threads per CTA                             // bytes of one 8-row group plane of a 128-pixel operand image

// wpk[tap][ks][gi][n][e] = bf16( wd[ng][tap] * Wp[ng][ks*16 + gi*8 + e] ),  ng = n < 64 ? n (Bm) : 64 + n (dt rows 128..191)
__global__ void hsm_fz_pack_kernel(const float* __restrict__ wp, const float* __restrict__ wd, __nv_bfloat16* __restrict__ wpk, int C) {
  const int KS = C / 16;
  const int total = 9 * KS * 2 * 128 * 8;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int e = idx & 7;
  int r = idx >> 3;
  int n = r & 127; r >>= 7;
  int gi = r & 1; r >>= 1;
  int ks = r % KS;
  int tap = r / KS;
  const int ng = n < 64 ? n : 64 + n, c = ks * 16 + gi * 8 + e;
  wpk[idx] = __float2bfloat16_rn(wd[ng * 9 + tap] * wp[(size_t)ng * C + c]);
}

__device__ __forceinline__ int f2ord(float f) {   // monotone float -> int map (for redux.sync.max)
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// x tile (+ halo, zero outside the image) -> bf16 K-group planes [g][NPOS][16 B]
template <int C>
__device__ __forceinline__ void stage_planes(const float* __restrict__ xb, uint8_t* planes, int H, int L, int ty0, int tx0, int tid,
                                             bool center_only) {
  constexpr int G = C / 8;
  constexpr int ITEMS = G * NPOS, BATCH = 3;     // BATCH items (24 loads) in flight per thread before the first conversion
  for (int u0 = 0; u0 < ITEMS; u0 += BATCH * NT) {
    float f[BATCH][8];
    bool in[BATCH];
#pragma unroll
    for (int k = 0; k < BATCH; ++k) {
      const int u = u0 + k * NT + tid;
      const int g = u / NPOS, pos = u - g * NPOS;
      const int py = pos / PITCH, px = pos - py * PITCH;
      const int gy = ty0 - 1 + py, gx = tx0 - 1 + px;
      in[k] = u < ITEMS && gy >= 0 && gy < H && gx >= 0 && gx < H &&
              !(center_only && (py == 0 || py == ROWS - 1 || px == 0 || px == PITCH - 1));
      const float* xp = xb + (size_t)(g * 8) * L + (size_t)gy * H + gx;
#pragma unroll
      for (int e = 0; e < 8; ++e) f[k][e] = in[k] ? __ldg(xp + (size_t)e * L) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < BATCH; ++k) {
      const int u = u0 + k * NT + tid;
      if (u < ITEMS) {
        const int g = u / NPOS, pos = u - g * NPOS;
        *reinterpret_cast<uint4*>(planes + (size_t)g * PLANE + (size_t)pos * 16) =
            make_uint4(pack_bf16x2(f[k][0], f[k][1]), pack_bf16x2(f[k][2], f[k][3]), pack_bf16x2(f[k][4], f[k][5]), pack_bf16x2(f[k][6], f[k][7]));
      }
    }
  }
}

// D1[px, 0:128] = conv3x3(x tile): 9 taps x KS k-steps, A = planes (K-major, tap = start offset), B = resident weights
template <int C>
__device__ __forceinline__ void issue_proj(uint32_t d_tmem, const uint8_t* planes, const uint8_t* w_base) {
  constexpr int KS = C / 16;
  constexpr uint32_t IDESC = make_idesc_bf16(128, 128);
  const uint64_t adesc0 = make_smem_desc(smem_u32(planes), PLANE, PITCH * 16);
  const uint64_t bdesc0 = make_smem_desc(smem_u32(w_base), 128 * 16, 128);
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int ki = t / 3, kj = t - ki * 3;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const uint64_t adesc = desc_advance(adesc0, (uint32_t)(ks * 2 * PLANE + (ki * PITCH + kj) * 16));
      const uint64_t bdesc = desc_advance(bdesc0, (uint32_t)((t * KS + ks) * WBLK));
      umma_bf16(d_tmem, adesc, bdesc, IDESC, (t > 0 || ks > 0) ? 1u : 0u);
    }
  }
}

// ================================================================================================ forward
template <int C>
__global__ void __launch_bounds__(NT, 2) hsm_hs_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ wpk,
                                                     float* __restrict__ part_m, float* __restrict__ part_s,
                                                     float* __restrict__ part_hs, int H, int tiles_x, int tiles_per_img, int ntiles) {
  constexpr int G = C / 8, KS = C / 16, NB = C + 16;       // NB: columns of the second product (x channels | ones | zeros)
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* planes = smem;                                      // [G + 2][NPOS][16 B]: x planes, the ones plane, a zero plane
  uint8_t* w_base = planes + (size_t)(G + 2) * PLANE;          // [9][KS][WBLK]
  uint8_t* e_img = w_base + (size_t)9 * KS * WBLK;             // [16 groups: Eb n/8, Ea n/8][128 px][16 B]   (MN-major A operand)
  int* wmax = reinterpret_cast<int*>(e_img + 16 * EGRP);       // [4 warps][64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(wmax + 4 * 64);  // 2 barriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = H * H;

  if (tid == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 128);
  {
    const uint4* src = reinterpret_cast<const uint4*>(wpk);
    uint4* dst = reinterpret_cast<uint4*>(w_base);
    for (int i = tid; i < 9 * KS * WBLK / 16; i += NT) dst[i] = __ldg(src + i);
    // constant planes: ones (element 0 of every position = 1.0 bf16) and zeros
    uint4* op = reinterpret_cast<uint4*>(planes + (size_t)G * PLANE);
    for (int i = tid; i < NPOS; i += NT) { op[i] = make_uint4(0x00003F80u, 0u, 0u, 0u); op[NPOS + i] = make_uint4(0u, 0u, 0u, 0u); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int lq = warp & 3, half = warp >> 2;                 // TMEM lane quarter, column half
  const uint32_t lane_base = tmem_base + ((uint32_t)(lq * 32) << 16);
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img, t_in = tile - b * tiles_per_img;
    const int ty0 = (t_in / tiles_x) * 16, tx0 = (t_in % tiles_x) * 8;
    stage_planes<C>(x + (size_t)b * C * L, planes, H, L, ty0, tx0, tid, false);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_proj<C>(tmem_base, planes, w_base);
      umma_commit(smem_u32(&bar[0]));
    }
    mbar_wait(smem_u32(&bar[0]), phase);
    tc_fence_after();
    // ---- epilogue 1: tile-local softmax reference, E image
    const int m = lq * 32 + lane;
    const bool ok = (ty0 + (m >> 3)) < H && (tx0 + (m & 7)) < H;
    float dt[32];                                             // this warp's half of the dt columns: n = half * 32 + [0, 32)
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(lane_base + 64u + (uint32_t)(half * 32 + c0), v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) dt[c0 + e] = __uint_as_float(v[e]);
    }
#pragma unroll
    for (int n = 0; n < 32; ++n) {
      const int mx = __reduce_max_sync(0xffffffffu, ok ? f2ord(dt[n]) : (int)0x80000000);
      if (lane == n) wmax[lq * 64 + half * 32 + n] = mx;
    }
    __syncthreads();
    const size_t bt = (size_t)b * tiles_per_img + t_in;
    if (tid < 64) {
      const int mx = max(max(wmax[tid], wmax[64 + tid]), max(wmax[128 + tid], wmax[192 + tid]));
      wmax[tid] = mx;
      part_m[bt * 64 + tid] = ord2f(mx);
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = half * 4 + jj;
      uint32_t v[8];
      tmem_ld8(lane_base + (uint32_t)(8 * j), v);       // Bm columns 8j .. 8j+7
      tmem_ld_wait();
      float ea[8], eb[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float mref = ord2f(wmax[8 * j + e]);
        ea[e] = ok ? __expf(dt[8 * jj + e] - mref) : 0.f;
        eb[e] = ea[e] * __uint_as_float(v[e]);
      }
      *reinterpret_cast<uint4*>(e_img + (size_t)j * EGRP + (size_t)m * 16) =
          make_uint4(pack_bf16x2(eb[0], eb[1]), pack_bf16x2(eb[2], eb[3]), pack_bf16x2(eb[4], eb[5]), pack_bf16x2(eb[6], eb[7]));
      *reinterpret_cast<uint4*>(e_img + (size_t)(8 + j) * EGRP + (size_t)m * 16) =
          make_uint4(pack_bf16x2(ea[0], ea[1]), pack_bf16x2(ea[2], ea[3]), pack_bf16x2(ea[4], ea[5]), pack_bf16x2(ea[6], ea[7]));
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- D2[(Eb | Ea) rows, NB] = sum over the tile's 128 pixels: both operands MN-major, K step = two tile rows
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t IDESC2 = make_idesc_bf16_mn(128, NB);
      const uint64_t adesc0 = make_smem_desc(smem_u32(e_img), 128, EGRP);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(planes + (PITCH + 1) * 16), PITCH * 16, PLANE);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        umma_bf16(tmem_base, desc_advance(adesc0, (uint32_t)(ks * 256)), desc_advance(bdesc0, (uint32_t)(ks * 2 * PITCH * 16)), IDESC2,
                  ks > 0 ? 1u : 0u);
      umma_commit(smem_u32(&bar[1]));
    }
    mbar_wait(smem_u32(&bar[1]), phase);
    tc_fence_after();
    // ---- epilogue 2: rows 0..63 = x (Ea Bm)^T columns 0..C-1, rows 64..127 column C = sum Ea
    if (m < 64) {
#pragma unroll
      for (int c0 = half * (C / 2); c0 < (half + 1) * (C / 2); c0 += 8) {
        uint32_t v[8];
        tmem_ld8(lane_base + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) part_hs[(bt * C + c0 + e) * 64 + m] = __uint_as_float(v[e]);
      }
    } else if (half == 0) {
      uint32_t v[1];
      tmem_ld1(lane_base + (uint32_t)C, v);
      tmem_ld_wait();
      part_s[bt * 64 + (m - 64)] = __uint_as_float(v[0]);
    }
    tc_fence_before();
    __syncthreads();
    phase ^= 1u;
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ================================================================================================ backward: dP, direct dx
template <int C>
__global__ void __launch_bounds__(NT, 2) hsm_dp_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                     const __nv_bfloat16* __restrict__ wpk, const float* __restrict__ stats,
                                                     const float* __restrict__ dhs, const float* __restrict__ ho,
                                                     const float* __restrict__ r, uint4* __restrict__ dPq, float* __restrict__ dx, int H,
                                                     int tiles_x, int tiles_per_img, int ntiles) {
  constexpr int G = C / 8, KS = C / 16;
  constexpr int PKB = KS * 2 * 64 * 16;                       // bytes of a [K = C][N = 64] K-major B image ( = C * 128 )
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* planes = smem;                                      // [2 G][NPOS][16 B]: x planes (with halo), dy planes (centre only)
  uint8_t* ab_img = planes;                                    // [8 groups n/8][128 px][16 B]: reuses the planes once D1 / D3 are complete
  uint8_t* w_base = planes + (size_t)(2 * G > 6 ? 2 * G : 6) * PLANE;   // >= 16 KB below for the A Bm image
  uint8_t* pk_dhs = w_base + (size_t)9 * KS * WBLK;            // [ks][gi][n 64][8 c]  = dhs[b, c, n]
  uint8_t* pk_ho = pk_dhs + PKB;                               // [ks][gi][n 64][8 c]  = ho[b, c, n]
  uint8_t* pk_dx = pk_ho + PKB;                                // [ks 4][gi][c][8 n]   = dhs[b, c, n]
  float* m_s = reinterpret_cast<float*>(pk_dx + PKB);          // [64] max, [64] 1 / sum, [64] r
  uint64_t* bar = reinterpret_cast<uint64_t*>(m_s + 192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int L = H * H;

  if (tid == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
  {
    const uint4* src = reinterpret_cast<const uint4*>(wpk);
    uint4* dst = reinterpret_cast<uint4*>(w_base);
    for (int i = tid; i < 9 * KS * WBLK / 16; i += NT) dst[i] = __ldg(src + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int lq = warp & 3, half = warp >> 2, lane = tid & 31;
  const uint32_t lane_base = tmem_base + ((uint32_t)(lq * 32) << 16);
  uint32_t phase = 0;
  int cur_b = -1;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img, t_in = tile - b * tiles_per_img;
    const int ty0 = (t_in / tiles_x) * 16, tx0 = (t_in % tiles_x) * 8;
    if (b != cur_b) {                                          // per-image operands of the two small products
      cur_b = b;
      const float* dh = dhs + (size_t)b * C * 64;
      const float* hb = ho + (size_t)b * C * 64;
      for (int i = tid; i < C * 64; i += NT) {
        const int c = i >> 6, n = i & 63;
        const __nv_bfloat16 vd = __float2bfloat16_rn(dh[i]), vh = __float2bfloat16_rn(hb[i]);
        const int o1 = (((c >> 4) * 2 + ((c >> 3) & 1)) * 64 + n) * 8 + (c & 7);      // [ks][gi][n][c % 8]
        reinterpret_cast<__nv_bfloat16*>(pk_dhs)[o1] = vd;
        reinterpret_cast<__nv_bfloat16*>(pk_ho)[o1] = vh;
        const int o2 = (((n >> 4) * 2 + ((n >> 3) & 1)) * C + c) * 8 + (n & 7);       // [ks][gi][c][n % 8]
        reinterpret_cast<__nv_bfloat16*>(pk_dx)[o2] = vd;
      }
      if (tid < 64) {
        m_s[tid] = stats[(size_t)b * 128 + tid];
        m_s[64 + tid] = 1.0f / stats[(size_t)b * 128 + 64 + tid];
        m_s[128 + tid] = r[(size_t)b * 64 + tid];
      }
    }
    stage_planes<C>(x + (size_t)b * C * L, planes, H, L, ty0, tx0, tid, false);
    stage_planes<C>(dy + (size_t)b * C * L, planes + (size_t)G * PLANE, H, L, ty0, tx0, tid, true);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_proj<C>(tmem_base, planes, w_base);                // D1: Bm | dt
      constexpr uint32_t IDESC3 = make_idesc_bf16(128, 64);
      const uint64_t ax0 = make_smem_desc(smem_u32(planes + (PITCH + 1) * 16), PLANE, PITCH * 16);               // centre pixels of x
      const uint64_t ay0 = make_smem_desc(smem_u32(planes + (size_t)G * PLANE + (PITCH + 1) * 16), PLANE, PITCH * 16);
      const uint64_t bd0 = make_smem_desc(smem_u32(pk_dhs), 64 * 16, 128);
      const uint64_t bh0 = make_smem_desc(smem_u32(pk_ho), 64 * 16, 128);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        umma_bf16(tmem_base + 128u, desc_advance(ax0, (uint32_t)(ks * 2 * PLANE)), desc_advance(bd0, (uint32_t)(ks * 2 * 64 * 16)), IDESC3,
                  ks > 0 ? 1u : 0u);                           // D3[:, 0:64]   = dG  = x^T dhs
        umma_bf16(tmem_base + 192u, desc_advance(ay0, (uint32_t)(ks * 2 * PLANE)), desc_advance(bh0, (uint32_t)(ks * 2 * 64 * 16)), IDESC3,
                  ks > 0 ? 1u : 0u);                           // D3[:, 64:128] = dCm = dy^T ho
      }
      umma_commit(smem_u32(&bar[0]));
    }
    mbar_wait(smem_u32(&bar[0]), phase);
    tc_fence_after();
    const int m = lq * 32 + lane;
    const int gy = ty0 + (m >> 3), gx = tx0 + (m & 7);
    const bool ok = gy < H && gx < H;
    uint4* dq = dPq + (size_t)b * 24 * L + (size_t)gy * H + gx;       // plane g at dq[g * L]
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = half * 4 + jj;
      uint32_t vb[8], vt[8], vg[8], vc[8];
      tmem_ld8(lane_base + (uint32_t)(8 * j), vb);             // Bm
      tmem_ld8(lane_base + 64u + (uint32_t)(8 * j), vt);       // dt
      tmem_ld8(lane_base + 128u + (uint32_t)(8 * j), vg);      // dG
      tmem_ld8(lane_base + 192u + (uint32_t)(8 * j), vc);      // dCm
      tmem_ld_wait();
      float dbm[8], ddt[8], ab[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int nn = 8 * j + e;
        const float bm = __uint_as_float(vb[e]), dG = __uint_as_float(vg[e]);
        const float a = __expf(__uint_as_float(vt[e]) - m_s[nn]) * m_s[64 + nn];
        dbm[e] = dG * a;
        ddt[e] = a * (dG * bm - m_s[128 + nn]);
        ab[e] = ok ? a * bm : 0.f;
      }
      if (ok) {
        dq[(size_t)j * L] = make_uint4(pack_bf16x2(dbm[0], dbm[1]), pack_bf16x2(dbm[2], dbm[3]), pack_bf16x2(dbm[4], dbm[5]),
                                       pack_bf16x2(dbm[6], dbm[7]));
        dq[(size_t)(8 + j) * L] = make_uint4(pack_bf16x2(__uint_as_float(vc[0]), __uint_as_float(vc[1])),
                                             pack_bf16x2(__uint_as_float(vc[2]), __uint_as_float(vc[3])),
                                             pack_bf16x2(__uint_as_float(vc[4]), __uint_as_float(vc[5])),
                                             pack_bf16x2(__uint_as_float(vc[6]), __uint_as_float(vc[7])));
        dq[(size_t)(16 + j) * L] = make_uint4(pack_bf16x2(ddt[0], ddt[1]), pack_bf16x2(ddt[2], ddt[3]), pack_bf16x2(ddt[4], ddt[5]),
                                              pack_bf16x2(ddt[6], ddt[7]));
      }
      *reinterpret_cast<uint4*>(ab_img + (size_t)j * EGRP + (size_t)m * 16) =
          make_uint4(pack_bf16x2(ab[0], ab[1]), pack_bf16x2(ab[2], ab[3]), pack_bf16x2(ab[4], ab[5]), pack_bf16x2(ab[6], ab[7]));
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- D4[px, c] = sum_n (A Bm)[px, n] dhs[c, n]: K-major image, K = 64
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t IDESC4 = make_idesc_bf16(128, C);
      const uint64_t a0 = make_smem_desc(smem_u32(ab_img), EGRP, 128);
      const uint64_t b0 = make_smem_desc(smem_u32(pk_dx), C * 16, 128);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem_base, desc_advance(a0, (uint32_t)(ks * 2 * EGRP)), desc_advance(b0, (uint32_t)(ks * 2 * C * 16)), IDESC4,
                  ks > 0 ? 1u : 0u);
      umma_commit(smem_u32(&bar[1]));
    }
    mbar_wait(smem_u32(&bar[1]), phase);
    tc_fence_after();
    {
      float* dxp = dx + (size_t)b * C * L + (size_t)gy * H + gx;
#pragma unroll
      for (int c0 = half * (C / 2); c0 < (half + 1) * (C / 2); c0 += 8) {
        uint32_t v[8];
        tmem_ld8(lane_base + (uint32_t)c0, v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int e = 0; e < 8; ++e) dxp[(size_t)(c0 + e) * L] = __uint_as_float(v[e]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    phase ^= 1u;
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---- merge the per-tile partials of one image (split softmax): two small kernels with fixed summation orders
// R1: grid B, 256 threads = 64 columns x 4 interleaved slices of the tile list.
//     m1[b][n] = max_t m_t ; sc[b][t][n] = exp(m_t - m1) ; s1[b][n] = sum_t s_t sc   (slice partials combined in slice order)
__global__ void __launch_bounds__(256) hsm_merge_stats_kernel(const float* __restrict__ part_m, const float* __restrict__ part_s,
                                                              float* __restrict__ m1, float* __restrict__ s1, float* __restrict__ sc,
                                                              int T) {
  __shared__ float red[4][64];
  const int b = blockIdx.x, n = threadIdx.x & 63, sl = threadIdx.x >> 6;
  const float* pm = part_m + (size_t)b * T * 64 + n;
  const float* ps = part_s + (size_t)b * T * 64 + n;
  float m = -INFINITY;
  for (int t = sl; t < T; t += 4) m = fmaxf(m, __ldg(pm + (size_t)t * 64));
  red[sl][n] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0][n], red[1][n]), fmaxf(red[2][n], red[3][n]));
  __syncthreads();
  float ssum = 0.f;
  for (int t = sl; t < T; t += 4) {
    const float e = __expf(__ldg(pm + (size_t)t * 64) - m);
    sc[((size_t)b * T + t) * 64 + n] = e;
    ssum = fmaf(__ldg(ps + (size_t)t * 64), e, ssum);
  }
  red[sl][n] = ssum;
  __syncthreads();
  if (sl == 0) {
    m1[(size_t)b * 64 + n] = m;
    s1[(size_t)b * 64 + n] = (red[0][n] + red[1][n]) + (red[2][n] + red[3][n]);
  }
}
// R2: grid (C, B), 256 threads = 64 columns x 4 interleaved slices of the tile list.  hs1[b][c][n] = sum_t part_hs[b][t][c][n] sc[b][t][n]
__global__ void __launch_bounds__(256) hsm_merge_hs_kernel(const float* __restrict__ part_hs, const float* __restrict__ sc,
                                                           float* __restrict__ hs1, int T, int C) {
  __shared__ float red[4][64];
  const int c = blockIdx.x, b = blockIdx.y, n = threadIdx.x & 63, sl = threadIdx.x >> 6;
  float a = 0.f;
#pragma unroll 4
  for (int t = sl; t < T; t += 4)
    a = fmaf(__ldg(part_hs + (((size_t)b * T + t) * C + c) * 64 + n), __ldg(sc + ((size_t)b * T + t) * 64 + n), a);
  red[sl][n] = a;
  __syncthreads();
  if (sl == 0) hs1[((size_t)b * C + c) * 64 + n] = (red[0][n] + red[1][n]) + (red[2][n] + red[3][n]);
}

size_t merge_bytes(int B, int C, int H) {
  const size_t T = (size_t)cdiv(H, 8) * cdiv(H, 16);
  return align_up((size_t)B * T * 64 * 4, 256) + 2 * align_up((size_t)B * 64 * 4, 256) + align_up((size_t)B * C * 64 * 4, 256);
}
// part_* (T partials per image) -> m1, s1, hs1 (ONE partial per image, the layout hsm_combine_gate reads with T = 1)
int merge(const float* part_m, const float* part_s, const float* part_hs, int B, int C, int H, void* workspace, float** m1, float** s1,
          float** hs1, cudaStream_t st) {
  const int T = cdiv(H, 8) * cdiv(H, 16);
  char* ws = (char*)workspace;
  float* sc = (float*)ws; ws += align_up((size_t)B * T * 64 * 4, 256);
  *m1 = (float*)ws; ws += align_up((size_t)B * 64 * 4, 256);
  *s1 = (float*)ws; ws += align_up((size_t)B * 64 * 4, 256);
  *hs1 = (float*)ws;
  hsm_merge_stats_kernel<<<B, 256, 0, st>>>(part_m, part_s, *m1, *s1, sc, T);
  KMU_LAUNCH_CHECK("hsm_merge_stats");
  hsm_merge_hs_kernel<<<dim3(C, B), 256, 0, st>>>(part_hs, sc, *hs1, T, C);
  KMU_LAUNCH_CHECK("hsm_merge_hs");
  return KMU_OK;
}

size_t pack_bytes(int C) { return align_up((size_t)9 * (C / 16) * WBLK, 256); }
int tiles_per_image(int H) { return cdiv(H, 8) * cdiv(H, 16); }

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int C>
static int launch_hs(const float* x, const __nv_bfloat16* wpk, float* part_m, float* part_s, float* part_hs, int B, int H,
                     cudaStream_t st) {
  constexpr int G = C / 8, KS = C / 16;
  const size_t smem = (size_t)(G + 2) * PLANE + (size_t)9 * KS * WBLK + 16 * EGRP + 4 * 64 * 4 + 64;
  cudaError_t e = cudaFuncSetAttribute(hsm_hs_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "hsm_hs: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
  const int tiles_x = cdiv(H, 8), tpi = tiles_per_image(H), ntiles = tpi * B;
  const int per_sm = (int)((227 * 1024) / (smem + 1024)) < 4 ? (int)((227 * 1024) / (smem + 1024)) : 4;
  int grid = sm_count() * (per_sm < 1 ? 1 : per_sm);
  if (grid > ntiles) grid = ntiles;
  hsm_hs_kernel<C><<<grid, NT, smem, st>>>(x, wpk, part_m, part_s, part_hs, H, tiles_x, tpi, ntiles);
  KMU_LAUNCH_CHECK("hsm_hs");
  return KMU_OK;
}

template <int C>
static int launch_dp(const float* x, const float* dy, const __nv_bfloat16* wpk, const float* stats, const float* dhs, const float* ho,
                     const float* r, void* dPp, float* dx, int B, int H, cudaStream_t st) {
  constexpr int G = C / 8, KS = C / 16;
  const size_t smem = (size_t)(2 * G > 6 ? 2 * G : 6) * PLANE + (size_t)9 * KS * WBLK + 3 * (size_t)C * 128 + 192 * 4 + 64;
  cudaError_t e = cudaFuncSetAttribute(hsm_dp_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "hsm_dp: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
  const int tiles_x = cdiv(H, 8), tpi = tiles_per_image(H), ntiles = tpi * B;
  const int per_sm = (int)((227 * 1024) / (smem + 1024)) < 2 ? (int)((227 * 1024) / (smem + 1024)) : 2;   // 256 TMEM columns each
  int grid = sm_count() * (per_sm < 1 ? 1 : per_sm);
  if (grid > ntiles) grid = ntiles;
  hsm_dp_kernel<C><<<grid, NT, smem, st>>>(x, dy, wpk, stats, dhs, ho, r, (uint4*)dPp, dx, H, tiles_x, tpi, ntiles);
  KMU_LAUNCH_CHECK("hsm_dp_tc");
  return KMU_OK;
}

static int pack(const float* wp, const float* wd, __nv_bfloat16* wpk, int C, cudaStream_t st) {
  const int total = 9 * (C / 16) * 2 * 128 * 8;
  hsm_fz_pack_kernel<<<cdiv(total, 256), 256, 0, st>>>(wp, wd, wpk, C);
  KMU_LAUNCH_CHECK("hsm_fz_pack");
  return KMU_OK;
}

// forward sweep: per-tile partials (reference max, sum of exponentials, x (e . Bm)^T) for hsm_combine_gate.  workspace >= pack_bytes(C)
int forward_hs(const float* x, const float* wp, const float* wd, float* part_m, float* part_s, float* part_hs, int B, int C, int H,
               void* workspace, cudaStream_t st) {
  __nv_bfloat16* wpk = (__nv_bfloat16*)workspace;
  int rc = pack(wp, wd, wpk, C, st);
  if (rc != KMU_OK) return rc;
  switch (C) {
    case 16: return launch_hs<16>(x, wpk, part_m, part_s, part_hs, B, H, st);
    case 32: return launch_hs<32>(x, wpk, part_m, part_s, part_hs, B, H, st);
    case 64: return launch_hs<64>(x, wpk, part_m, part_s, part_hs, B, H, st);
  }
  set_error("hsm_hs: unsupported C=%d", C);
  return KMU_ERR_UNSUPPORTED;
}

// backward: dP planes (bf16, [b][n / 8][l][n % 8]) and the direct part of dx.  workspace >= pack_bytes(C)
int dp(const float* x, const float* dy, const float* wp, const float* wd, const float* stats, const float* dhs, const float* ho,
       const float* r, void* dPp, float* dx, int B, int C, int H, void* workspace, cudaStream_t st) {
  __nv_bfloat16* wpk = (__nv_bfloat16*)workspace;
  int rc = pack(wp, wd, wpk, C, st);
  if (rc != KMU_OK) return rc;
  switch (C) {
    case 16: return launch_dp<16>(x, dy, wpk, stats, dhs, ho, r, dPp, dx, B, H, st);
    case 32: return launch_dp<32>(x, dy, wpk, stats, dhs, ho, r, dPp, dx, B, H, st);
    case 64: return launch_dp<64>(x, dy, wpk, stats, dhs, ho, r, dPp, dx, B, H, st);
  }
  set_error("hsm_dp_tc: unsupported C=%d", C);
  return KMU_ERR_UNSUPPORTED;
}

}  // namespace fz
}  // namespace hsm
}  // namespace kmu
