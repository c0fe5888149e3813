// pwconv_bwd_tc.cu -- fused backward of the pointwise (1x1) convolution: ONE persistent kernel reads x and dy once (TMA, fp32,
// straight from the NCHW planes) and produces dx, dW and db.
//
// Same operator as pwconv.cu / pwconv_tc.cu (vim_utils_init.py:122-130 FFN, KM_UNetV3_SH.py:59,118-122,178,221); this file
// replaces the autograd pair (dgrad kernel + wgrad kernel + reduce), each of which streamed dy from HBM again:
//     dx[c, p]  = sum_o W[o, c] dy[o, p]          M = 128 pixels, N = Cin,      K = Cout     (A = dy, K-major)
//     dW[o, c]  = sum_p dy[o, p] x[c, p]          M = Cout (128-row halves), N = Cin + 16, K = pixels (A = dy, B = x, MN-major)
//     db[o]     = column Cin of the same GEMM (a constant-one plane in B)
// Pipeline per CTA (one per SM, persistent over 128-pixel tiles):
//     warp 0      TMA producer: two 2-D boxes per tile, [Cin][128] and [Cout][128] fp32, into a ring of raw stages
//     warps 2..9  convert a raw stage to the bf16 [channel group][pixel][8] plane image (the one image that is both the K-major
//                 A operand of dgrad and the MN-major A / B operands of wgrad, see pwconv_tc.cu), then run the dx epilogue of the
//                 PREVIOUS tile (TMEM -> coalesced NCHW stores) while the tensor core works on the current one
//     warp 1      MMA issuer: Cout/16 dgrad MMAs into a double-buffered accumulator + 8 x ceil(Cout/128) wgrad MMAs into
//                 accumulators that stay in TMEM for the CTA's whole life (written once: partial[cta][half][128][Cin + 16])
// bf16 operands, fp32 accumulation (2e-2 gate, the precision class of the reference's fp16-autocast convolutions).
#include <cuda.h>  // CUtensorMap and enums only

#include "common.cuh"
#include "tc_common.cuh"

namespace kmu {
namespace pwtc {

// pwconv_tc.cu
void launch_pack(const float* w, __nv_bfloat16* wpack, int NI, int NJ, int NJp, int dgrad, cudaStream_t st);
void launch_wreduce(const float* partial, int nparts, int MH, int Np, int Cin, int Cout, float* dw, float* db, cudaStream_t st);

namespace fused {

using namespace kmu::tcx;

constexpr int TPX = 128;
constexpr int PLANE = TPX * 16;          // bytes of one 8-channel bf16 plane of a tile
constexpr int NCONV = 8;                 // converter / epilogue warps
constexpr int NTHREADS = (2 + NCONV) * 32;
constexpr int MAXST = 4;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

// weights fp32 -> the bf16 K-major B-operand image [ks][gi][n][8] straight into shared memory (every CTA converts its own copy:
// at most 64 K values from L2, cheaper than a separate pack launch per call).  transposed: element (k, n) = w[k * NJ + n] (dgrad of
// W (Cout, Cin) with K = Cout), else w[n * NI + k] (forward, K = Cin).
__device__ __forceinline__ void pack_weights_smem(const float* __restrict__ w, uint8_t* w_base, int NI, int NJ, bool transposed) {
  const int total = NI * NJ / 8;                      // 16-byte units: (ks, gi, n)
  for (int u = threadIdx.x; u < total; u += blockDim.x) {
    const int n = u % NJ, r = u / NJ;                 // r = ks * 2 + gi
    const int k0 = r * 8;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = transposed ? __ldg(w + (size_t)(k0 + e) * NJ + n) : __ldg(w + (size_t)n * NI + k0 + e);
    *reinterpret_cast<uint4*>(w_base + (size_t)u * 16) =
        make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
}

// residual image of the same weights: bf16(w - bf16(w)) -- the low half of the split-bf16 forward (x = hi + lo to ~16 mantissa bits)
__device__ __forceinline__ void pack_weights_lo_smem(const float* __restrict__ w, uint8_t* w_base, int NI, int NJ) {
  const int total = NI * NJ / 8;
  for (int u = threadIdx.x; u < total; u += blockDim.x) {
    const int n = u % NJ, r = u / NJ;
    const int k0 = r * 8;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float v = __ldg(w + (size_t)n * NI + k0 + e);
      f[e] = v - __bfloat162float(__float2bfloat16_rn(v));
    }
    *reinterpret_cast<uint4*>(w_base + (size_t)u * 16) =
        make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
}

struct Plan {
  int Cin, Cout, HW, B;
  int MH, Np;            // 128-row halves of Cout; UMMA N of the wgrad GEMM (Cin + ones group + zero group)
  int GA, GB;            // plane groups of one plane buffer: dy (Cout / 8), x (+ ones + zero)
  int NST, NPL, NACC;    // raw stages, plane buffers, dgrad accumulators
  int raw_stage;         // bytes of one raw stage: x rows then dy rows, 512 B each
  int plane_buf;         // bytes of one plane buffer
  int wbytes;            // dgrad weights (bf16 K-major B operand)
  int tmem_cols, wacc_col;
  int tiles_per_img, ntiles;
  size_t smem;
};

__global__ void __launch_bounds__(NTHREADS, 1) pw_bwd_fused_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                   const __grid_constant__ CUtensorMap map_dy,
                                                                   const float* __restrict__ w, float* __restrict__ dx,
                                                                   float* __restrict__ partial, const Plan pl) {
  extern __shared__ __align__(128) uint8_t smem[];
  // The wgrad A descriptor always spans 16 channel groups per 128-row half; groups >= Cout/8 are whatever follows the dy planes
  // (x planes, then the raw stages): they only feed accumulator rows >= Cout, which nobody reads.  The plane buffers therefore
  // come FIRST so that this window stays inside the CTA's shared memory (make_plan checks it).
  uint8_t* pl_base = smem;                                                 // [NPL][GA + GB][128][16 B]
  uint8_t* raw_base = pl_base + (size_t)pl.NPL * pl.plane_buf;             // [NST][(Cin + Cout)][128] fp32
  uint8_t* w_base = raw_base + (size_t)pl.NST * pl.raw_stage;              // [Cout/16][2][Cin][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + pl.wbytes);
  uint64_t* raw_full = bars;                 // [MAXST]
  uint64_t* raw_empty = bars + MAXST;        // [MAXST]
  uint64_t* pl_full = bars + 2 * MAXST;      // [2]
  uint64_t* pl_empty = pl_full + 2;          // [2]
  uint64_t* acc_full = pl_empty + 2;         // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint64_t* wacc_full = acc_empty + 2;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wacc_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Cin = pl.Cin, Cout = pl.Cout, HW = pl.HW;

  // zero the plane buffers once: the zero group of B (columns Cin + 8 .. Cin + 15 of the wgrad GEMM) stays zero
  for (int i = tid; i < pl.NPL * pl.plane_buf / 16; i += NTHREADS) reinterpret_cast<uint4*>(pl_base)[i] = make_uint4(0u, 0u, 0u, 0u);
  pack_weights_smem(w, w_base, pl.Cout, pl.Cin, true);      // dgrad: K = Cout, N = Cin
  if (tid == 0) {
    for (int i = 0; i < MAXST; ++i) {
      mbar_init(smem_u32(&raw_full[i]), 1);
      mbar_init(smem_u32(&raw_empty[i]), NCONV);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&pl_full[i]), NCONV);
      mbar_init(smem_u32(&pl_empty[i]), 1);
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), NCONV);
    }
    mbar_init(smem_u32(wacc_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)pl.tmem_cols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < pl.ntiles; tile += gridDim.x, ++it) {
        const int b = tile / pl.tiles_per_img, p0 = (tile - b * pl.tiles_per_img) * TPX;
        const uint32_t s = it % (uint32_t)pl.NST;
        mbar_wait(smem_u32(&raw_empty[s]), ((it / (uint32_t)pl.NST) & 1u) ^ 1u);
        const uint32_t bar = smem_u32(&raw_full[s]);
        mbar_expect_tx(bar, (uint32_t)pl.raw_stage);
        const uint32_t dst = smem_u32(raw_base + (size_t)s * pl.raw_stage);
        tma_load_2d(dst, &map_x, p0, b * Cin, bar);
        tma_load_2d(dst + (uint32_t)(Cin * 512), &map_dy, p0, b * Cout, bar);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc_dg = make_idesc_bf16(128, Cin);
      const uint32_t idesc_wg = make_idesc_bf16_mn(128, pl.Np);
      const uint64_t wdesc0 = make_smem_desc(smem_u32(w_base), (uint32_t)(Cin * 16), 128);
      const int KS = Cout / 16;
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < pl.ntiles; tile += gridDim.x, ++it) {
        const uint32_t pb = it % (uint32_t)pl.NPL, a = it % (uint32_t)pl.NACC;
        mbar_wait_hot(smem_u32(&acc_empty[a]), ((it / (uint32_t)pl.NACC) & 1u) ^ 1u);
        mbar_wait_hot(smem_u32(&pl_full[pb]), (it / (uint32_t)pl.NPL) & 1u);
        tc_fence_after();
        const uint32_t a0 = smem_u32(pl_base + (size_t)pb * pl.plane_buf);
        // dgrad: A = dy planes K-major (LBO = plane step between the two K halves, SBO = 128 B between 8-pixel row groups)
        const uint64_t adesc_k = make_smem_desc(a0, PLANE, 128);
        const uint32_t d_acc = tmem_base + a * (uint32_t)Cin;
        for (int ks = 0; ks < KS; ++ks)
          umma_bf16(d_acc, desc_advance(adesc_k, (uint32_t)(ks * 2 * PLANE)), desc_advance(wdesc0, (uint32_t)(ks * 2 * Cin * 16)), idesc_dg,
                    ks > 0 ? 1u : 0u);
        // wgrad: both operands MN-major (LBO = 128 B between the two 8-pixel K groups, SBO = plane step between channel groups)
        const uint64_t adesc_m = make_smem_desc(a0, 128, PLANE);
        const uint64_t bdesc_m = desc_advance(adesc_m, (uint32_t)(pl.GA * PLANE));
        for (int h = 0; h < pl.MH; ++h) {
#pragma unroll
          for (int ks = 0; ks < TPX / 16; ++ks)
            umma_bf16(tmem_base + (uint32_t)(pl.wacc_col + h * pl.Np), desc_advance(adesc_m, (uint32_t)(h * 16 * PLANE + ks * 256)),
                      desc_advance(bdesc_m, (uint32_t)(ks * 256)), idesc_wg, (it > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&pl_empty[pb]));
        umma_commit(smem_u32(&acc_full[a]));
      }
      umma_commit(smem_u32(wacc_full));
    }
  } else {
    // ===================================================================== converters + epilogue (warps 2..9)
    const int cw = warp - 2;                 // 0..7
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int chalf = cw >> 2;               // which half of the accumulator columns this warp stores
    const int ct = cw * 32 + lane;           // 0..255
    const int px = ct & 127, gpar = ct >> 7; // pixel of the tile, group parity
    const int gx = Cin / 8, gy = Cout / 8;
    auto epilogue = [&](uint32_t jt, int tile) {
      const int b = tile / pl.tiles_per_img, p0 = (tile - b * pl.tiles_per_img) * TPX;
      const uint32_t a = jt % (uint32_t)pl.NACC;
      mbar_wait(smem_u32(&acc_full[a]), (jt / (uint32_t)pl.NACC) & 1u);
      tc_fence_after();
      const int p = p0 + q * 32 + lane;
      const int ncol = Cin / 2, c_lo = chalf * ncol;
      float* op = dx + ((size_t)b * Cin + c_lo) * HW + p;
      for (int c0 = 0; c0 < ncol; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + a * (uint32_t)Cin + (uint32_t)(c_lo + c0), v);
        tmem_ld_wait();
        if (p < HW) {
#pragma unroll
          for (int e = 0; e < 8; ++e) op[(size_t)(c0 + e) * HW] = __uint_as_float(v[e]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[a]));
    };
    uint32_t it = 0;
    int prev_tile = -1;
    for (int tile = blockIdx.x; tile < pl.ntiles; tile += gridDim.x, ++it) {
      const int b = tile / pl.tiles_per_img, p0 = (tile - b * pl.tiles_per_img) * TPX;
      const uint32_t s = it % (uint32_t)pl.NST, pb = it % (uint32_t)pl.NPL;
      mbar_wait(smem_u32(&raw_full[s]), (it / (uint32_t)pl.NST) & 1u);
      mbar_wait(smem_u32(&pl_empty[pb]), ((it / (uint32_t)pl.NPL) & 1u) ^ 1u);
      const float* rx = reinterpret_cast<const float*>(raw_base + (size_t)s * pl.raw_stage) + px;
      const float* ry = rx + Cin * TPX;
      uint8_t* pa = pl_base + (size_t)pb * pl.plane_buf + (size_t)px * 16;
      uint8_t* pbx = pa + (size_t)pl.GA * PLANE;
      for (int g = gpar; g < gy; g += 2) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = ry[(g * 8 + e) * TPX];
        *reinterpret_cast<uint4*>(pa + (size_t)g * PLANE) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      }
      for (int g = gpar; g < gx; g += 2) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = rx[(g * 8 + e) * TPX];
        *reinterpret_cast<uint4*>(pbx + (size_t)g * PLANE) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      }
      if (gpar == 0)  // ones plane: element 0 of group gx is 1 for the pixels that exist
        *reinterpret_cast<uint4*>(pbx + (size_t)gx * PLANE) = make_uint4(p0 + px < HW ? 0x00003F80u : 0u, 0u, 0u, 0u);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&pl_full[pb]));
        mbar_arrive(smem_u32(&raw_empty[s]));
      }
      if (prev_tile >= 0) epilogue(it - 1, prev_tile);
      prev_tile = tile;
    }
    if (prev_tile >= 0) epilogue(it - 1, prev_tile);
    // the weight-gradient accumulators, once
    mbar_wait(smem_u32(wacc_full), 0);
    tc_fence_after();
    {
      const int m = q * 32 + lane;
      const int ncol = pl.Np / 2, c_lo = chalf * ncol;
      for (int h = 0; h < pl.MH; ++h) {
        float* pp = partial + (((size_t)blockIdx.x * pl.MH + h) * 128 + m) * pl.Np + c_lo;
        for (int c0 = 0; c0 < ncol; c0 += 8) {
          uint32_t v[8];
          tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pl.wacc_col + h * pl.Np + c_lo + c0), v);
          tmem_ld_wait();
          *reinterpret_cast<float4*>(pp + c0) = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          *reinterpret_cast<float4*>(pp + c0 + 4) = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)pl.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ forward (same pipeline)
// y[b, o, p] = bias[o] + sum_c W[o, c] x[b, c, p]: TMA box [Cin][128] fp32 -> converter warps -> bf16 planes -> Cin/16 MMAs of
// M = 128 pixels x N = Cout -> epilogue warps store NCHW rows (128 B per warp and channel).  Also the input gradient of the pairs
// the fused backward does not take (called with the transposed weights).
struct FwdPlan {
  int Cin, Cout, HW, B;
  int NST;               // raw stages
  int raw_stage, plane_buf, wbytes, tmem_cols, NPL;
  int tiles_per_img, ntiles, ctas;
  size_t smem;
};

__global__ void __launch_bounds__(NTHREADS) pw_fwd_tma_kernel(const __grid_constant__ CUtensorMap map_x,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              float* __restrict__ y, const FwdPlan pl) {
  extern __shared__ __align__(128) uint8_t smem[];
  // Split-bf16 operands: every fp32 value is fed to the tensor core as hi + lo (hi = bf16(v), lo = bf16(v - hi)) and the product as
  // hi*hi + lo*hi + hi*lo -- ~16 mantissa bits instead of 8 at three times the MMA work, which is free here (the kernel is HBM-bound).
  // The bf16-only version was the largest single source of end-to-end error of the tensor-core precision class (tools/bf16_ablation.py).
  uint8_t* pl_base = smem;                                                 // [2][hi, lo][Cin/8][128][16 B]
  uint8_t* raw_base = pl_base + (size_t)pl.NPL * pl.plane_buf;                  // [NST][Cin][128] fp32
  uint8_t* w_base = raw_base + (size_t)pl.NST * pl.raw_stage;              // [hi, lo][Cin/16][2][Cout][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + pl.wbytes);
  uint64_t* raw_full = bars;                 // [MAXST]
  uint64_t* raw_empty = bars + MAXST;        // [MAXST]
  uint64_t* pl_full = bars + 2 * MAXST;      // [2]
  uint64_t* pl_empty = pl_full + 2;          // [2]
  uint64_t* acc_full = pl_empty + 2;         // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Cin = pl.Cin, Cout = pl.Cout, HW = pl.HW;

  pack_weights_smem(w, w_base, pl.Cin, pl.Cout, false);     // forward: K = Cin, N = Cout
  pack_weights_lo_smem(w, w_base + pl.wbytes / 2, pl.Cin, pl.Cout);
  if (tid == 0) {
    for (int i = 0; i < MAXST; ++i) {
      mbar_init(smem_u32(&raw_full[i]), 1);
      mbar_init(smem_u32(&raw_empty[i]), NCONV);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&pl_full[i]), NCONV);
      mbar_init(smem_u32(&pl_empty[i]), 1);
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), NCONV);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)pl.tmem_cols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < pl.ntiles; tile += gridDim.x, ++it) {
        const int b = tile / pl.tiles_per_img, p0 = (tile - b * pl.tiles_per_img) * TPX;
        const uint32_t s = it % (uint32_t)pl.NST;
        mbar_wait(smem_u32(&raw_empty[s]), ((it / (uint32_t)pl.NST) & 1u) ^ 1u);
        const uint32_t bar = smem_u32(&raw_full[s]);
        mbar_expect_tx(bar, (uint32_t)pl.raw_stage);
        tma_load_2d(smem_u32(raw_base + (size_t)s * pl.raw_stage), &map_x, p0, b * Cin, bar);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, Cout);
      const uint64_t wdesc0 = make_smem_desc(smem_u32(w_base), (uint32_t)(Cout * 16), 128);
      const int KS = Cin / 16;
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < pl.ntiles; tile += gridDim.x, ++it) {
        const uint32_t ab = it & 1u, ah = (it >> 1) & 1u;                               // accumulator buffer / phase
        const uint32_t pb = it % (uint32_t)pl.NPL, ph = (it / (uint32_t)pl.NPL) & 1u;   // plane buffer / phase
        mbar_wait_hot(smem_u32(&acc_empty[ab]), ah ^ 1u);
        mbar_wait_hot(smem_u32(&pl_full[pb]), ph);
        tc_fence_after();
        const uint64_t adesc0 = make_smem_desc(smem_u32(pl_base + (size_t)pb * pl.plane_buf), PLANE, 128);
        const uint32_t d_acc = tmem_base + ab * (uint32_t)Cout;
        const uint32_t a_lo = (uint32_t)(pl.plane_buf / 2), w_lo = (uint32_t)(pl.wbytes / 2);
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t ao = (uint32_t)(ks * 2 * PLANE), wo = (uint32_t)(ks * 2 * Cout * 16);
          umma_bf16(d_acc, desc_advance(adesc0, ao + a_lo), desc_advance(wdesc0, wo), idesc, ks > 0 ? 1u : 0u);   // lo * hi (small terms first)
          umma_bf16(d_acc, desc_advance(adesc0, ao), desc_advance(wdesc0, wo + w_lo), idesc, 1u);                 // hi * lo
          umma_bf16(d_acc, desc_advance(adesc0, ao), desc_advance(wdesc0, wo), idesc, 1u);                        // hi * hi
        }
        umma_commit(smem_u32(&pl_empty[pb]));
        umma_commit(smem_u32(&acc_full[ab]));
      }
    }
  } else {
    const int cw = warp - 2, q = warp & 3, chalf = cw >> 2;
    const int ct = cw * 32 + lane, px = ct & 127, gpar = ct >> 7;
    const int gx = Cin / 8;
    auto epilogue = [&](uint32_t jt, int tile) {
      const int b = tile / pl.tiles_per_img, p0 = (tile - b * pl.tiles_per_img) * TPX;
      const uint32_t a = jt & 1u;
      mbar_wait(smem_u32(&acc_full[a]), (jt >> 1) & 1u);
      tc_fence_after();
      const int p = p0 + q * 32 + lane;
      const int ncol = Cout / 2, c_lo = chalf * ncol;
      float* op = y + ((size_t)b * Cout + c_lo) * HW + p;
      for (int c0 = 0; c0 < ncol; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + a * (uint32_t)Cout + (uint32_t)(c_lo + c0), v);
        tmem_ld_wait();
        if (p < HW) {
#pragma unroll
          for (int e = 0; e < 8; ++e) op[(size_t)(c0 + e) * HW] = __uint_as_float(v[e]) + (bias ? __ldg(bias + c_lo + c0 + e) : 0.f);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[a]));
    };
    uint32_t it = 0;
    int prev_tile = -1;
    for (int tile = blockIdx.x; tile < pl.ntiles; tile += gridDim.x, ++it) {
      const uint32_t s = it % (uint32_t)pl.NST, pb = it % (uint32_t)pl.NPL;
      mbar_wait(smem_u32(&raw_full[s]), (it / (uint32_t)pl.NST) & 1u);
      mbar_wait(smem_u32(&pl_empty[pb]), ((it / (uint32_t)pl.NPL) & 1u) ^ 1u);
      const float* rx = reinterpret_cast<const float*>(raw_base + (size_t)s * pl.raw_stage) + px;
      uint8_t* pa = pl_base + (size_t)pb * pl.plane_buf + (size_t)px * 16;
      for (int g = gpar; g < gx; g += 2) {
        float f[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          f[e] = rx[(g * 8 + e) * TPX];
          l[e] = f[e] - __bfloat162float(__float2bfloat16_rn(f[e]));
        }
        *reinterpret_cast<uint4*>(pa + (size_t)g * PLANE) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        *reinterpret_cast<uint4*>(pa + (size_t)(pl.plane_buf / 2) + (size_t)g * PLANE) =
            make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&pl_full[pb]));
        mbar_arrive(smem_u32(&raw_empty[s]));
      }
      if (prev_tile >= 0) epilogue(it - 1, prev_tile);
      prev_tile = tile;
    }
    if (prev_tile >= 0) epilogue(it - 1, prev_tile);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)pl.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// fp32 rows [rows][HW], box = [box_rows][128 pixels]
static int make_row_map(CUtensorMap* m, const void* base, long long rows, int HW, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  KMU_REQUIRE(fn != nullptr, KMU_ERR_LAUNCH, "pwconv_fused_bwd: cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)HW * 4};
  cuuint32_t box[2] = {(cuuint32_t)TPX, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KMU_REQUIRE(r == CUDA_SUCCESS, KMU_ERR_LAUNCH, "pwconv_fused_bwd: cuTensorMapEncodeTiled failed (%d) for %lld rows of %d", (int)r, rows, HW);
  return KMU_OK;
}

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

static int ctas_for(const kmu_pwconv_desc& d) {
  long long tiles = (long long)d.B * cdiv(d.HW, TPX);
  return (int)(tiles < 148 ? tiles : 148);
}

// false when the shape does not fit (channel counts, TMA alignment, shared memory or TMEM budget)
static bool make_plan(const kmu_pwconv_desc& d, Plan* out) {
  if (d.Cin % 16 || d.Cout % 16 || d.Cin < 16 || d.Cout < 16 || d.Cin > 240 || d.Cout > 256) return false;
  if (d.HW % 4 || d.HW < TPX || d.B <= 0) return false;    // TMA: 16-byte row pitch; tiny maps stay on the streaming kernels
  Plan p;
  p.Cin = d.Cin; p.Cout = d.Cout; p.HW = d.HW; p.B = d.B;
  p.MH = cdiv(d.Cout, 128);
  p.Np = d.Cin + 16;
  p.GA = d.Cout / 8;
  p.GB = p.Np / 8;
  p.raw_stage = (d.Cin + d.Cout) * 512;
  p.plane_buf = (p.GA + p.GB) * PLANE;
  p.wbytes = d.Cin * d.Cout * 2;
  const int budget = 220 * 1024 - p.wbytes - 512;
  p.NPL = 2;
  p.NST = (budget - 2 * p.plane_buf) / p.raw_stage;
  if (p.NST < 2) {
    p.NPL = 1;
    p.NST = (budget - p.plane_buf) / p.raw_stage;
  }
  if (p.NST < 2) return false;
  if (p.NST > MAXST) p.NST = MAXST;
  if (p.plane_buf + p.NST * p.raw_stage < p.MH * 16 * PLANE) return false;   // the 16-group window of the last plane buffer
  p.NACC = 2;
  if (2 * d.Cin + p.MH * p.Np > 512) p.NACC = 1;
  if (p.NACC * d.Cin + p.MH * p.Np > 512) return false;
  p.wacc_col = p.NACC * d.Cin;
  p.tmem_cols = pow2_cols(p.wacc_col + p.MH * p.Np);
  p.tiles_per_img = cdiv(d.HW, TPX);
  p.ntiles = d.B * p.tiles_per_img;
  p.smem = (size_t)p.NST * p.raw_stage + (size_t)p.NPL * p.plane_buf + p.wbytes + 256;
  *out = p;
  return true;
}

struct Ws { size_t wpack, partial, total; };
static Ws ws_layout(const kmu_pwconv_desc& d, const Plan& p) {
  Ws w;
  size_t o = 0;
  w.wpack = o; o += align_up((size_t)p.wbytes, 256);
  w.partial = o; o += align_up((size_t)ctas_for(d) * p.MH * 128 * p.Np * 4, 256);
  w.total = o;
  return w;
}

static bool make_fwd_plan(const kmu_pwconv_desc& d, FwdPlan* out) {
  if (d.Cin % 16 || d.Cout % 16 || d.Cin < 16 || d.Cout < 16 || d.Cin > 256 || d.Cout > 256) return false;
  if (d.HW % 4 || d.HW < TPX || d.B <= 0) return false;
  FwdPlan p;
  p.Cin = d.Cin; p.Cout = d.Cout; p.HW = d.HW; p.B = d.B;
  p.raw_stage = d.Cin * 512;
  p.plane_buf = 2 * (d.Cin / 8) * PLANE;       // hi and lo planes
  p.wbytes = 2 * d.Cin * d.Cout * 2;           // hi and lo weights
  p.tmem_cols = pow2_cols(2 * d.Cout);
  // two CTAs per SM when shared memory and TMEM allow it: the store-heavy epilogue of one overlaps the other's loads
  // in order of preference: two CTAs per SM (the store-heavy epilogue of one overlaps the other's loads) with two plane buffers,
  // two CTAs with one plane buffer (the conversion of tile i + 1 then waits for the MMAs of tile i), one CTA with two / one buffers
  int per_sm = 1;
  p.NST = 0;
  for (int choice = 0; choice < 4 && p.NST < 2; ++choice) {
    per_sm = choice < 2 ? 2 : 1;
    p.NPL = (choice & 1) ? 1 : 2;
    if (per_sm == 2 && p.tmem_cols > 256) continue;
    const int budget = (per_sm == 2 ? 108 : 220) * 1024 - p.NPL * p.plane_buf - p.wbytes - 512;
    p.NST = budget > 0 ? budget / p.raw_stage : 0;
  }
  if (p.NST < 2) return false;
  if (p.NST > MAXST) p.NST = MAXST;
  p.tiles_per_img = cdiv(d.HW, TPX);
  p.ntiles = d.B * p.tiles_per_img;
  p.ctas = p.ntiles < 148 * per_sm ? p.ntiles : 148 * per_sm;
  p.smem = (size_t)p.NST * p.raw_stage + (size_t)p.NPL * p.plane_buf + p.wbytes + 256;
  *out = p;
  return true;
}

}  // namespace fused
}  // namespace pwtc
}  // namespace kmu

using namespace kmu;
using namespace kmu::pwtc;
using namespace kmu::pwtc::fused;

extern "C" {

int kmu_pwconv_tma_fwd_supported(const kmu_pwconv_desc* d) {
  FwdPlan p;
  return (d && make_fwd_plan(*d, &p)) ? 1 : 0;
}

size_t kmu_pwconv_tma_fwd_workspace_bytes(const kmu_pwconv_desc* d) {
  FwdPlan p;
  if (!d || !make_fwd_plan(*d, &p)) return 0;
  return align_up((size_t)p.wbytes, 256);
}

int kmu_pwconv_tma_fwd(const kmu_pwconv_desc* d, const float* x, const float* w, const float* bias, float* y, void* workspace,
                       size_t workspace_bytes, kmu_stream stream) {
  FwdPlan p;
  KMU_REQUIRE(d && make_fwd_plan(*d, &p) && d->B <= 65535, KMU_ERR_UNSUPPORTED, "pwconv_tma_fwd: unsupported shape");
  KMU_REQUIRE(x && w && y, KMU_ERR_BAD_ARG, "pwconv_tma_fwd: null tensor");
  KMU_REQUIRE(((uintptr_t)x & 15) == 0, KMU_ERR_BAD_ARG, "pwconv_tma_fwd: x must be 16-byte aligned");
  KMU_REQUIRE(workspace && workspace_bytes >= align_up((size_t)p.wbytes, 256), KMU_ERR_WORKSPACE, "pwconv_tma_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* wpack = (__nv_bfloat16*)workspace;
  CUtensorMap map_x;
  int rc = make_row_map(&map_x, x, (long long)d->B * d->Cin, d->HW, d->Cin);
  if (rc != KMU_OK) return rc;
  cudaError_t e = cudaFuncSetAttribute(pw_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "pwconv_tma_fwd: cannot opt in to %zu B shared memory: %s", p.smem, cudaGetErrorString(e));
  pw_fwd_tma_kernel<<<p.ctas, NTHREADS, p.smem, st>>>(map_x, w, bias, y, p);
  KMU_LAUNCH_CHECK("pw_fwd_tma");
  return KMU_OK;
}

int kmu_pwconv_fused_bwd_supported(const kmu_pwconv_desc* d) {
  Plan p;
  return (d && make_plan(*d, &p)) ? 1 : 0;
}

size_t kmu_pwconv_fused_bwd_workspace_bytes(const kmu_pwconv_desc* d) {
  Plan p;
  if (!d || !make_plan(*d, &p)) return 0;
  return ws_layout(*d, p).total;
}

int kmu_pwconv_fused_bwd(const kmu_pwconv_desc* d, const float* x, const float* dy, const float* w, float* dx, float* dw, float* dbias,
                         void* workspace, size_t workspace_bytes, kmu_stream stream) {
  Plan p;
  KMU_REQUIRE(d && make_plan(*d, &p) && d->B <= 65535, KMU_ERR_UNSUPPORTED, "pwconv_fused_bwd: unsupported shape");
  KMU_REQUIRE(x && dy && w && dx && dw, KMU_ERR_BAD_ARG, "pwconv_fused_bwd: null tensor (dx and dw are both produced)");
  KMU_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0, KMU_ERR_BAD_ARG, "pwconv_fused_bwd: x / dy must be 16-byte aligned");
  const Ws wl = ws_layout(*d, p);
  KMU_REQUIRE(workspace && workspace_bytes >= wl.total, KMU_ERR_WORKSPACE, "pwconv_fused_bwd: workspace %zu < %zu", workspace_bytes, wl.total);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* wpack = (__nv_bfloat16*)((char*)workspace + wl.wpack);
  float* partial = (float*)((char*)workspace + wl.partial);
  CUtensorMap map_x, map_dy;
  int rc = make_row_map(&map_x, x, (long long)d->B * d->Cin, d->HW, d->Cin);
  if (rc != KMU_OK) return rc;
  rc = make_row_map(&map_dy, dy, (long long)d->B * d->Cout, d->HW, d->Cout);
  if (rc != KMU_OK) return rc;
  cudaError_t e = cudaFuncSetAttribute(pw_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "pwconv_fused_bwd: cannot opt in to %zu B shared memory: %s", p.smem, cudaGetErrorString(e));
  const int ctas = ctas_for(*d);
  pw_bwd_fused_kernel<<<ctas, NTHREADS, p.smem, st>>>(map_x, map_dy, w, dx, partial, p);
  KMU_LAUNCH_CHECK("pw_bwd_fused");
  launch_wreduce(partial, ctas, p.MH, p.Np, d->Cin, d->Cout, dw, dbias, st);
  KMU_LAUNCH_CHECK("pw_wreduce_tc");
  return KMU_OK;
}

}  // extern "C"
