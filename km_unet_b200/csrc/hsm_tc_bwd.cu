// hsm_tc_bwd.cu -- tcgen05 / TMEM / TMA path of the HSM-SSD projection backward (KMU_PREC_BF16).
//
// Backward of P = dw3x3(Wp x) (vim_block_init/efficient_vim_init.py:39 and the autograd graph PyTorch builds for it), with
// the pair treated as ONE dense 3x3 convolution W[n, tap, c] = wd[n, tap] Wp[n, c] exactly like the forward (hsm_tc.cu):
//     dgrad   dx[c, q]          += sum_{tap, n} W[n, tap, c] dP[n, q + 1 - tap]                    M = pixels, N = c, K = 9 * 192
//     wgrad   dWfull[n, tap, c]  = sum_p dP[n, p] x[c, p + tap - 1]                                M = n, N = (tap, c), K = pixels
//             dWp[n, c] = sum_tap wd[n, tap] dWfull[n, tap, c] ;  dWd[n, tap] = sum_c Wp[n, c] dWfull[n, tap, c]     (chain rule)
// hsm_dp (hsmssd.cu) writes dP directly as bf16 "K-group planes" dPp[b][n / 8][l][n % 8] and hsm_xpack does the same for x, so
// a spatial tile with its 1-pixel halo ([plane][18 rows][10 positions][16 B]) is ONE 3-D TMA box per 8 planes: out-of-image
// halo positions are zero-filled by the TMA unit, no thread touches the operands.  That single shared-memory image is
//   * the UMMA K-major A operand of dgrad  (rows = pixels, 16 B = 8 n;   tap = descriptor start offset, as in the forward) and
//   * the UMMA MN-major A / B operands of wgrad (K rows = 8 consecutive pixels, 16 B = 8 n resp. 8 c; tap = start offset of x).
// Both kernels are persistent warp-specialised pipelines: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue.
#include <cuda.h>  // CUtensorMap and enums only: cuTensorMapEncodeTiled is fetched at run time (cudaGetDriverEntryPoint)

#include "common.cuh"
#include "tc_common.cuh"

namespace kmu {
namespace hsm {
namespace tcb {

using namespace kmu::tcx;

constexpr int PITCH = 10, ROWS = 18, NPOS = PITCH * ROWS;  // 8 x 16 pixel tile + 1-pixel halo
constexpr int PLANE = NPOS * 16;                           // bytes of one 8-channel plane of a tile (2880)
constexpr int BOXP = 8;                                    // planes per TMA box of dP
constexpr int BOX_BYTES = BOXP * PLANE;                    // 23040
constexpr int NTHREADS = 192;

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ------------------------------------------------------------------------------------------------ operand packing
// xp[b][c / 8][l][c % 8] = bf16(x[b][c][l])
__global__ void __launch_bounds__(256) hsm_xpack_kernel(const float* __restrict__ x, uint4* __restrict__ xp, int C, int L, long long total) {
  long long idx = (long long)blockIdx.x * 256 + threadIdx.x;  // over (b, g, l)
  if (idx >= total) return;
  const int l = (int)(idx % L);
  const long long bg = idx / L;
  const int G = C / 8;
  const int g = (int)(bg % G);
  const long long b = bg / G;
  const float* sp = x + ((size_t)b * C + (size_t)g * 8) * L + l;
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = __ldg(sp + (size_t)e * L);
  xp[idx] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// w2pack[cblk][tap][ks][gi][c16][e] = bf16( wd[n][tap] * Wp[n][cblk*16 + c16] ),  n = ks*16 + gi*8 + e   (K-major B operand of dgrad)
__global__ void hsm_w2pack_kernel(const float* __restrict__ wp, const float* __restrict__ wd, __nv_bfloat16* __restrict__ w2pack, int C) {
  const int total = 9 * 192 * C;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int e = idx & 7;
  int r = idx >> 3;
  int c16 = r & 15; r >>= 4;
  int gi = r & 1; r >>= 1;
  int ks = r % 12; r /= 12;
  int tap = r % 9;
  int cblk = r / 9;
  const int n = ks * 16 + gi * 8 + e, c = cblk * 16 + c16;
  w2pack[idx] = __float2bfloat16_rn(wd[n * 9 + tap] * wp[(size_t)n * C + c]);
}

// ------------------------------------------------------------------------------------------------ dgrad
constexpr int DG_STAGES = 2;
constexpr int DG_STAGE = 24 * PLANE;       // all 192 channels of dP for one tile (69120 B)
constexpr int DG_WBLK = 2 * 16 * 16;       // one (tap, k-step) block of the weights [gi][c16][8] (512 B)
constexpr int DG_W = 9 * 12 * DG_WBLK;     // 55296 B
constexpr size_t DG_SMEM = (size_t)DG_STAGES * DG_STAGE + DG_W + 128;

// grid (persistent CTAs, C / 16): CTA = one 16-channel block of dx, walks its share of the 8 x 16 pixel tiles.
__global__ void __launch_bounds__(NTHREADS, 1) hsm_dgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dp,
                                                                   const __nv_bfloat16* __restrict__ w2pack, float* __restrict__ dx,
                                                                   int C, int H, int tiles_x, int tiles_per_img, int ntiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* a_base = smem;                                  // [DG_STAGES][24][NPOS][16 B]
  uint8_t* w_base = smem + (size_t)DG_STAGES * DG_STAGE;   // [9][12][DG_WBLK]
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + DG_W);
  uint64_t* full = bars;            // [2] TMA -> MMA
  uint64_t* empty = bars + 2;       // [2] MMA -> TMA
  uint64_t* acc_full = bars + 4;    // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 6;   // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cblk = blockIdx.y;
  const int L = H * H;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_dp);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 32);
  {  // this channel block's weights: resident for the CTA's whole life
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(w2pack) + (size_t)cblk * DG_W);
    uint4* dst = reinterpret_cast<uint4*>(w_base);
    for (int i = tid; i < DG_W / 16; i += NTHREADS) dst[i] = __ldg(src + i);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
        const int ty0 = (tr / tiles_x) * 16, tx0 = (tr % tiles_x) * 8;
        const uint32_t s = it & 1u;
        mbar_wait(smem_u32(&empty[s]), ((it >> 1) & 1u) ^ 1u);
        const uint32_t bar = smem_u32(&full[s]);
        mbar_expect_tx(bar, DG_STAGE);
        const uint32_t dst = smem_u32(a_base + (size_t)s * DG_STAGE);
#pragma unroll
        for (int k = 0; k < 3; ++k) tma_load_3d(dst + (uint32_t)(k * BOX_BYTES), &map_dp, (tx0 - 1) * 8, ty0 - 1, b * 24 + k * BOXP, bar);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc_bf16(128, 16);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(w_base), 16 * 16, 128);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t s = it & 1u;
        mbar_wait_hot(smem_u32(&acc_empty[s]), ((it >> 1) & 1u) ^ 1u);
        mbar_wait_hot(smem_u32(&full[s]), (it >> 1) & 1u);
        tc_fence_after();
        const uint64_t adesc0 = make_smem_desc(smem_u32(a_base + (size_t)s * DG_STAGE), PLANE, PITCH * 16);
        const uint32_t d_tmem = tmem_base + s * 16u;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ki = t / 3, kj = t - ki * 3;
          const uint32_t ashift = (uint32_t)(((2 - ki) * PITCH + (2 - kj)) * 16);
#pragma unroll
          for (int ks = 0; ks < 12; ++ks) {
            const uint64_t adesc = desc_advance(adesc0, (uint32_t)(ks * 2 * PLANE) + ashift);
            const uint64_t bdesc = desc_advance(bdesc0, (uint32_t)((t * 12 + ks) * DG_WBLK));
            umma_bf16(d_tmem, adesc, bdesc, IDESC, (t > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&empty[s]));
        umma_commit(smem_u32(&acc_full[s]));
      }
    }
  } else {
    // ===================================================================== epilogue: dx += accumulator
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
      const int ty0 = (tr / tiles_x) * 16, tx0 = (tr % tiles_x) * 8;
      const uint32_t s = it & 1u;
      const int m = q * 32 + lane;
      const int gy = ty0 + (m >> 3), gx = tx0 + (m & 7);
      const bool ok = gy < H && gx < H;
      float* dp = dx + ((size_t)b * C + (size_t)cblk * 16) * L + (size_t)gy * H + gx;
      float old[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) old[e] = ok ? dp[(size_t)e * L] : 0.f;  // issued before the wait: overlaps the MMAs
      mbar_wait(smem_u32(&acc_full[s]), (it >> 1) & 1u);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + s * 16u, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[s]));
      if (ok) {
#pragma unroll
        for (int e = 0; e < 16; ++e) dp[(size_t)e * L] = old[e] + __uint_as_float(v[e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

// ------------------------------------------------------------------------------------------------ wgrad
// One tcgen05.mma of M = 128 takes ~45-64 cycles whatever N <= 128 is (tools/probes/mma_align_probe.cu), so nine N = 16 MMAs per K
// step (one per tap, the tap being a start offset into ONE halo tile of x) cost nine times the tensor-pipe time of a single
// N = 144 MMA.  The taps are therefore moved into the N dimension: the TMA unit builds the im2col image of x itself -- nine boxes
// of the same [NCW/8 planes][16 rows][8 pixels] shape, each shifted by its tap (out-of-image parts zero-filled) -- which lands as
// [tap][channel group][128 pixels][16 B], i.e. an MN-major B operand with N = 9 * NCW whose groups sit at one uniform stride.
constexpr int WPLANE = 128 * 16;     // one 8-channel plane of a 16 x 8 pixel tile WITHOUT halo (2048 B)
constexpr int WG_A = 16 * WPLANE;    // dP planes of one 128-row half (32768 B)

// grid (persistent CTAs, 2 halves of n, C / NCW): the CTA keeps dWfull[half rows][9 taps][NCW channels] in TMEM (9 * NCW columns)
// over all its tiles and writes it once: partial[cta][half][cz][128][9 * NCW].  Rows 64..127 of half 1 do not exist (n < 192):
// their A planes stay zero and their accumulator rows are never read back.
//
// per_batch != 0 turns the same pipeline into the per-image correlation G[b][c'][tap][c] = sum_p dy[b, c', p] x[b, c, p + tap - 1]
// (A = dy planes, a_planes_per_b = C / 8 of them per image in ONE box; grid (CTAs per image, 1, Z * B); rows >= C are not
// written): d(ho)[b] = G[b] . W_Cm replaces the backward's sweep over the materialised Cm slice (hsm_dho_kernel below).
template <int NCW>
__global__ void __launch_bounds__(NTHREADS, 1) hsm_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dp,
                                                                   const __grid_constant__ CUtensorMap map_x,
                                                                   float* __restrict__ partial, int C, int tiles_x, int tiles_per_img,
                                                                   int ntiles, int per_batch, int a_planes_per_b, int a_box_planes,
                                                                   int Z) {
  constexpr int XT = (NCW / 8) * WPLANE;         // x planes of one tap
  constexpr int XB = 9 * XT;                     // im2col image of x for one tile
  constexpr int STAGE = WG_A + XB;
  constexpr int STAGES = NCW == 16 ? 3 : 2;
  constexpr int NCOL = 9 * NCW;
  constexpr int TCOLS = NCOL <= 256 ? 256 : 512;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + STAGES;        // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = blockIdx.y, cz = blockIdx.z % Z;
  // tile range of this CTA: every tile of the batch, or the tiles of image blockIdx.z / Z
  const int tile_begin = per_batch ? (int)(blockIdx.z / Z) * tiles_per_img + (int)blockIdx.x : (int)blockIdx.x;
  const int tile_end = per_batch ? (int)(blockIdx.z / Z + 1) * tiles_per_img : ntiles;

  for (int i = tid; i < STAGES * STAGE / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    mbar_init(smem_u32(acc_full), 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_dp);
    tma_prefetch_desc(&map_x);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TCOLS);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      const int nbox = per_batch ? 1 : (half == 0 ? 2 : 1);
      const int box_bytes = a_box_planes * WPLANE;
      uint32_t it = 0;
      for (int tile = tile_begin; tile < tile_end; tile += gridDim.x, ++it) {
        const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
        const int ty0 = (tr / tiles_x) * 16, tx0 = (tr % tiles_x) * 8;
        const uint32_t s = it % STAGES;
        mbar_wait(smem_u32(&empty[s]), ((it / STAGES) & 1u) ^ 1u);
        const uint32_t bar = smem_u32(&full[s]);
        mbar_expect_tx(bar, (uint32_t)(nbox * box_bytes + XB));
        const uint32_t dst = smem_u32(smem + (size_t)s * STAGE);
        for (int k = 0; k < nbox; ++k)
          tma_load_3d(dst + (uint32_t)(k * box_bytes), &map_dp, tx0 * 8, ty0, b * a_planes_per_b + half * 16 + k * a_box_planes, bar);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ki = t / 3, kj = t - ki * 3;
          tma_load_3d(dst + (uint32_t)(WG_A + t * XT), &map_x, (tx0 + kj - 1) * 8, ty0 + ki - 1, b * (C / 8) + cz * (NCW / 8), bar);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = tile_begin; tile < tile_end; tile += gridDim.x, ++it) {
        const uint32_t s = it % STAGES;
        mbar_wait_hot(smem_u32(&full[s]), (it / STAGES) & 1u);
        tc_fence_after();
        // MN-major: LBO = byte step between the two 8-pixel K groups (next tile row, 128 B), SBO = byte step between 8-channel groups
        const uint64_t adesc0 = make_smem_desc(smem_u32(smem + (size_t)s * STAGE), 128, WPLANE);
        const uint64_t bdesc0 = desc_advance(adesc0, (uint32_t)WG_A);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // K step = two tile rows of 8 pixels
          const uint64_t adesc = desc_advance(adesc0, (uint32_t)(ks * 256));
          const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
          if (NCW == 16) {
            umma_bf16(tmem_base, adesc, desc_advance(bdesc0, (uint32_t)(ks * 256)), make_idesc_bf16_mn(128, 144), acc);
          } else {  // N = 288 > 256: taps 0..3 (128 columns) and 4..8 (160 columns)
            umma_bf16(tmem_base, adesc, desc_advance(bdesc0, (uint32_t)(ks * 256)), make_idesc_bf16_mn(128, 128), acc);
            umma_bf16(tmem_base + 128u, adesc, desc_advance(bdesc0, (uint32_t)(4 * XT + ks * 256)), make_idesc_bf16_mn(128, 160), acc);
          }
        }
        umma_commit(smem_u32(&empty[s]));
      }
      umma_commit(smem_u32(acc_full));
    }
  } else {
    // ===================================================================== epilogue (once)
    const int q = warp & 3;
    mbar_wait(smem_u32(acc_full), 0);
    tc_fence_after();
    const int m = q * 32 + lane;
    float* pp = partial + ((((size_t)blockIdx.x * gridDim.y + half) * gridDim.z + blockIdx.z) * 128 + m) * NCOL;
    const bool real = per_batch ? (m < a_planes_per_b * 8) : (half == 0 || m < 64);
#pragma unroll 1
    for (int c0 = 0; c0 < NCOL; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (real) {
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          *reinterpret_cast<float4*>(pp + c0 + e) =
              make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

static size_t wgrad_smem(int NCW) { return (size_t)(NCW == 16 ? 3 : 2) * (WG_A + 9 * (NCW / 8) * WPLANE) + 128; }

// dWfull[n][tap][c] = sum over CTAs, fixed order.  CTA = 32 outputs x 8 interleaved slices of the CTA list.
__global__ void __launch_bounds__(256) hsm_wgrad_tc_reduce_kernel(const float* __restrict__ partial, int nctas, int Z, int NCW, int C,
                                                                  float* __restrict__ dwfull) {
  __shared__ float red[8][33];
  const int o_l = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + o_l;  // over (n, tap, c)
  const int total = 192 * 9 * C;
  float s = 0.f;
  if (idx < total) {
    const int c = idx % C, r = idx / C;
    const int tap = r % 9, n = r / 9;
    const int half = n >> 7, m = n & 127, cz = c / NCW, cc = c - cz * NCW;
    const size_t ncol = (size_t)9 * NCW;
    for (int k = sl; k < nctas; k += 8) s += partial[((((size_t)k * 2 + half) * Z + cz) * 128 + m) * ncol + (size_t)tap * NCW + cc];
  }
  red[sl][o_l] = s;
  __syncthreads();
  if (sl == 0 && idx < total) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][o_l];
    dwfull[idx] = t;
  }
}

// chain rule of W[n, tap, c] = wd[n, tap] Wp[n, c]
__global__ void __launch_bounds__(256) hsm_wgrad_tc_chain_kernel(const float* __restrict__ dwfull, const float* __restrict__ wp,
                                                                 const float* __restrict__ wd, float* __restrict__ dwp,
                                                                 float* __restrict__ dwd, int C) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int n_wp = 192 * C;
  if (idx < n_wp) {
    const int n = idx / C, c = idx - n * C;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) s = fmaf(wd[n * 9 + t], dwfull[((size_t)n * 9 + t) * C + c], s);
    dwp[idx] = s;
  } else if (idx < n_wp + 192 * 9) {
    const int j = idx - n_wp;
    const int n = j / 9;
    const float* src = dwfull + (size_t)j * C;
    const float* w = wp + (size_t)n * C;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(w[c], src[c], s);
    dwd[j] = s;
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;  // same value from every thread: a benign race
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// bf16 planes [planes][H][H * 8 elements], box = [box_planes][18 rows][10 pixels] (tile + halo) or [box_planes][16 rows][8 pixels]
static int make_plane_map(CUtensorMap* m, const void* base, long long planes, int H, int box_planes, bool halo = true) {
  EncodeTiledFn fn = encode_fn();
  KMU_REQUIRE(fn != nullptr, KMU_ERR_LAUNCH, "hsm_tc_bwd: cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {(cuuint64_t)H * 8, (cuuint64_t)H, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)H * 16, (cuuint64_t)H * H * 16};
  cuuint32_t box[3] = {(cuuint32_t)((halo ? PITCH : 8) * 8), (cuuint32_t)(halo ? ROWS : 16), (cuuint32_t)box_planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KMU_REQUIRE(r == CUDA_SUCCESS, KMU_ERR_LAUNCH, "hsm_tc_bwd: cuTensorMapEncodeTiled failed (%d) for %lld planes of %d x %d", (int)r,
              planes, H, H);
  return KMU_OK;
}

static int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

static int wgrad_ncw(int C) { return C < 32 ? C : 32; }
static int wgrad_ctas(int B, int C, int H) {
  const int Z = C / wgrad_ncw(C);
  long long ntiles = (long long)B * cdiv(H, 8) * cdiv(H, 16);
  int g = 148 / (2 * Z);
  if (g < 1) g = 1;
  return (int)(ntiles < g ? ntiles : g);
}

// wt[tap][c][m] = wd[64 + m, tap] * Wp[64 + m, c]: the Cm rows of the merged convolution, state index m fastest (coalesced reads)
__global__ void hsm_wt_kernel(const float* __restrict__ wp, const float* __restrict__ wd, float* __restrict__ wt, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 9 * C * 64) return;
  const int m = idx & 63, c = (idx >> 6) % C, t = idx / (64 * C);
  wt[idx] = wd[(64 + m) * 9 + t] * wp[(size_t)(64 + m) * C + c];
}

// d(ho)[b][c'][m] = sum_{tap, c} G[b][c'][tap][c] * wt[tap][c][m],  G = fixed-order sum of the CTAs' partials.
// CTA = (b, c'): the G row is reduced into shared memory once, thread m contracts it with column m of wt.
__global__ void __launch_bounds__(64) hsm_dho_kernel(const float* __restrict__ gpart, const float* __restrict__ wt,
                                                     float* __restrict__ dho, int B, int C, int nct, int Z, int NCW) {
  extern __shared__ float g_s[];  // [9][C]
  const int b = blockIdx.x / C, cp = blockIdx.x - b * C, m = threadIdx.x;
  const size_t ncol = (size_t)9 * NCW;
  for (int j = m; j < 9 * C; j += 64) {
    const int t = j / C, c = j - t * C;
    const int cz = c / NCW, cc = c - cz * NCW;
    float g = 0.f;
    for (int k = 0; k < nct; ++k) g += gpart[((((size_t)k * B + b) * Z + cz) * 128 + cp) * ncol + (size_t)t * NCW + cc];
    g_s[j] = g;
  }
  __syncthreads();
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const int n = 9 * C;   // multiple of 4
  for (int j = 0; j < n; j += 4) {
    s0 = fmaf(g_s[j], __ldg(wt + (size_t)j * 64 + m), s0);
    s1 = fmaf(g_s[j + 1], __ldg(wt + (size_t)(j + 1) * 64 + m), s1);
    s2 = fmaf(g_s[j + 2], __ldg(wt + (size_t)(j + 2) * 64 + m), s2);
    s3 = fmaf(g_s[j + 3], __ldg(wt + (size_t)(j + 3) * 64 + m), s3);
  }
  dho[((size_t)b * C + cp) * 64 + m] = (s0 + s1) + (s2 + s3);
}

static int corr_ctas(int B, int C, int H) {
  const int Z = C / wgrad_ncw(C);
  int per = 148 / (B * Z);
  if (per < 1) per = 1;
  const int tiles = cdiv(H, 8) * cdiv(H, 16);
  return per < tiles ? per : tiles;
}

size_t dpp_bytes(int B, int L) { return align_up((size_t)B * 24 * L * 16, 256); }

struct Ws { size_t xp, dyp, w2, partial, dwfull, total; };
static Ws ws_layout(int B, int C, int H) {
  Ws w;
  size_t o = 0;
  const int L = H * H;
  w.xp = o; o += align_up((size_t)B * (C / 8) * L * 16, 256);
  w.dyp = o; o += align_up((size_t)B * (C / 8) * L * 16, 256);
  w.w2 = o; o += align_up((size_t)9 * 192 * C * 2, 256);
  const int NCW = wgrad_ncw(C), Z = C / NCW;
  size_t part = (size_t)wgrad_ctas(B, C, H) * 2 * Z * 128 * 9 * NCW * 4;
  const size_t gpart = (size_t)corr_ctas(B, C, H) * B * Z * 128 * 9 * NCW * 4;    // the correlation runs first, same buffer
  if (gpart > part) part = gpart;
  w.partial = o; o += align_up(part, 256);
  w.dwfull = o; o += align_up((size_t)192 * 9 * C * 4, 256);
  w.total = o;
  return w;
}
size_t workspace_bytes(int B, int C, int H) { return ws_layout(B, C, H).total; }

// part_dho[b][c'][m] = sum_p dy[b, c', p] Cm[b, m, p] without Cm: correlation of dy and x on tensor cores, then . W_Cm.
// Also leaves x packed (xp) in the workspace for backward().
int contract(const float* x, const float* dy, const float* wp, const float* wd, float* dho, int B, int C, int H, void* workspace,
             cudaStream_t st) {
  KMU_REQUIRE(C == 16 || C == 32 || C == 64, KMU_ERR_UNSUPPORTED, "hsm_tc_bwd: unsupported C=%d", C);
  const int L = H * H;
  const Ws wl = ws_layout(B, C, H);
  char* ws = (char*)workspace;
  uint4* xp = (uint4*)(ws + wl.xp);
  uint4* dyp = (uint4*)(ws + wl.dyp);
  float* gpart = (float*)(ws + wl.partial);
  const int tiles_x = cdiv(H, 8), tiles_per_img = tiles_x * cdiv(H, 16), ntiles = B * tiles_per_img;
  const int NCW = wgrad_ncw(C), Z = C / NCW;
  CUtensorMap map_dy, map_x;
  int rc = make_plane_map(&map_dy, dyp, (long long)B * (C / 8), H, C / 8, false);
  if (rc != KMU_OK) return rc;
  rc = make_plane_map(&map_x, xp, (long long)B * (C / 8), H, NCW / 8, false);
  if (rc != KMU_OK) return rc;
  const long long total = (long long)B * (C / 8) * L;
  hsm_xpack_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(x, xp, C, L, total);
  KMU_LAUNCH_CHECK("hsm_xpack");
  hsm_xpack_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(dy, dyp, C, L, total);
  KMU_LAUNCH_CHECK("hsm_xpack(dy)");
  const int nct = corr_ctas(B, C, H);
  const size_t smem = wgrad_smem(NCW);
  cudaError_t e;
  if (NCW == 16) {
    e = cudaFuncSetAttribute(hsm_wgrad_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "hsm_corr_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
    hsm_wgrad_tc_kernel<16><<<dim3(nct, 1, Z * B), NTHREADS, smem, st>>>(map_dy, map_x, gpart, C, tiles_x, tiles_per_img, ntiles, 1, C / 8, C / 8, Z);
  } else {
    e = cudaFuncSetAttribute(hsm_wgrad_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "hsm_corr_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
    hsm_wgrad_tc_kernel<32><<<dim3(nct, 1, Z * B), NTHREADS, smem, st>>>(map_dy, map_x, gpart, C, tiles_x, tiles_per_img, ntiles, 1, C / 8, C / 8, Z);
  }
  KMU_LAUNCH_CHECK("hsm_corr_tc");
  float* wt = (float*)(ws + wl.dwfull);   // 9 * C * 64 floats: fits the (later) dWfull buffer of 192 * 9 * C floats
  hsm_wt_kernel<<<cdiv(9 * C * 64, 256), 256, 0, st>>>(wp, wd, wt, C);
  KMU_LAUNCH_CHECK("hsm_wt");
  hsm_dho_kernel<<<B * C, 64, 9 * C * sizeof(float), st>>>(gpart, wt, dho, B, C, nct, Z, NCW);
  KMU_LAUNCH_CHECK("hsm_dho");
  return KMU_OK;
}

// dx += dgrad(dPp) ; dWp, dWd = chain(wgrad(dPp, x)).  dPp = bf16 planes written by hsm_dp; workspace >= workspace_bytes().
int backward(const float* x, const float* wp, const float* wd, const void* dPp, float* dx, float* dwp, float* dwd, int B, int C, int H,
             void* workspace, int x_packed, cudaStream_t st) {
  KMU_REQUIRE(C == 16 || C == 32 || C == 64, KMU_ERR_UNSUPPORTED, "hsm_tc_bwd: unsupported C=%d", C);
  const int L = H * H;
  const Ws wl = ws_layout(B, C, H);
  char* ws = (char*)workspace;
  uint4* xp = (uint4*)(ws + wl.xp);
  __nv_bfloat16* w2 = (__nv_bfloat16*)(ws + wl.w2);
  float* partial = (float*)(ws + wl.partial);
  float* dwfull = (float*)(ws + wl.dwfull);
  const int tiles_x = cdiv(H, 8), tiles_per_img = tiles_x * cdiv(H, 16), ntiles = B * tiles_per_img;
  const int NCW = wgrad_ncw(C), Z = C / NCW;

  CUtensorMap map_dp, map_dpw, map_x;
  int rc = make_plane_map(&map_dp, dPp, (long long)B * 24, H, BOXP);            // dgrad: tile + halo
  if (rc != KMU_OK) return rc;
  rc = make_plane_map(&map_dpw, dPp, (long long)B * 24, H, BOXP, false);         // wgrad: the tile itself
  if (rc != KMU_OK) return rc;
  rc = make_plane_map(&map_x, xp, (long long)B * (C / 8), H, NCW / 8, false);    // wgrad: nine tap-shifted boxes of x
  if (rc != KMU_OK) return rc;

  {
    const long long total = (long long)B * (C / 8) * L;
    if (!x_packed) {
      hsm_xpack_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(x, xp, C, L, total);
      KMU_LAUNCH_CHECK("hsm_xpack");
    }
    hsm_w2pack_kernel<<<cdiv(9 * 192 * C, 256), 256, 0, st>>>(wp, wd, w2, C);
    KMU_LAUNCH_CHECK("hsm_w2pack");
  }
  {
    const int cb = C / 16;
    int gx = sm_count() / cb;
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    cudaError_t e = cudaFuncSetAttribute(hsm_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DG_SMEM);
    KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "hsm_dgrad_tc: cannot opt in to %zu B shared memory: %s", DG_SMEM, cudaGetErrorString(e));
    hsm_dgrad_tc_kernel<<<dim3(gx, cb), NTHREADS, DG_SMEM, st>>>(map_dp, w2, dx, C, H, tiles_x, tiles_per_img, ntiles);
    KMU_LAUNCH_CHECK("hsm_dgrad_tc");
  }
  {
    const int ctas = wgrad_ctas(B, C, H);
    const size_t smem = wgrad_smem(NCW);
    cudaError_t e;
    if (NCW == 16) {
      e = cudaFuncSetAttribute(hsm_wgrad_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "hsm_wgrad_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
      hsm_wgrad_tc_kernel<16><<<dim3(ctas, 2, Z), NTHREADS, smem, st>>>(map_dpw, map_x, partial, C, tiles_x, tiles_per_img, ntiles, 0, 24, BOXP, Z);
    } else {
      e = cudaFuncSetAttribute(hsm_wgrad_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "hsm_wgrad_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
      hsm_wgrad_tc_kernel<32><<<dim3(ctas, 2, Z), NTHREADS, smem, st>>>(map_dpw, map_x, partial, C, tiles_x, tiles_per_img, ntiles, 0, 24, BOXP, Z);
    }
    KMU_LAUNCH_CHECK("hsm_wgrad_tc");
    hsm_wgrad_tc_reduce_kernel<<<cdiv(192 * 9 * C, 32), 256, 0, st>>>(partial, ctas, Z, NCW, C, dwfull);
    KMU_LAUNCH_CHECK("hsm_wgrad_tc_reduce");
    hsm_wgrad_tc_chain_kernel<<<cdiv(192 * C + 192 * 9, 256), 256, 0, st>>>(dwfull, wp, wd, dwp, dwd, C);
    KMU_LAUNCH_CHECK("hsm_wgrad_tc_chain");
  }
  return KMU_OK;
}

}  // namespace tcb
}  // namespace hsm
}  // namespace kmu
