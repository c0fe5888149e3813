// pwconv_tc.cu -- tcgen05 / TMEM path of the pointwise (1x1) convolution (KMU_PREC_BF16): forward, input gradient and
// weight / bias gradient as small tensor-core GEMMs on NCHW tensors.
//
// Same operator as pwconv.cu (vim_utils_init.py:122-130 FFN, KM_UNetV3_SH.py:59,118-122,178,221), bf16 operands with fp32
// TMEM accumulation -- the precision class of the reference's own GPU run (fp16 autocast convolutions), 2e-2 gate.
// Layout trick: a pixel-contiguous NCHW plane group, read as 8 channels x 1 pixel per thread, lands in shared memory as
// [channel group][pixel][8 x bf16] = 16-byte rows.  That one image is
//   * the UMMA K-major A operand of the forward / dgrad GEMM  (M = 128 pixels, K = channels), and
//   * the UMMA MN-major operand of the weight-gradient GEMM   (K = pixels, MN = channels)
// so all three directions share the same stage-in code.  CTAs are small (128 threads, 16..100 KB of shared memory, <= 512 TMEM
// columns) and several are resident per SM; there is no warp specialisation: stage -> one thread issues the MMAs -> commit ->
// everybody waits on the mbarrier -> epilogue.
#include "common.cuh"
#include "tc_common.cuh"

namespace kmu {
namespace pwtc {

using namespace kmu::tcx;

constexpr int TPX = 128;            // pixels per tile (= UMMA M of the forward GEMM)
constexpr int PLANE = TPX * 16;     // bytes of one 8-channel group of a tile

// stage `ngroups` channel groups of one 128-pixel tile: src (B, NC, HW) fp32 -> dst [group][pixel][8 x bf16]
__device__ __forceinline__ void stage_planes(const float* __restrict__ src, uint8_t* dst, int ngroups, int NC, int HW, int b, int p0) {
  const float* sb = src + (size_t)b * NC * HW;
  for (int u = threadIdx.x; u < ngroups * TPX; u += 128) {
    const int g = u / TPX, pos = u - g * TPX;
    const int p = p0 + pos;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (p < HW) {
      const float* sp = sb + (size_t)(g * 8) * HW + p;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __ldg(sp + (size_t)e * HW);
      v = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
    *reinterpret_cast<uint4*>(dst + (size_t)g * PLANE + (size_t)pos * 16) = v;
  }
}

// wpack[ks][gi][n][e] (bf16, K-major B operand): forward W[n][k], dgrad W[k][n] with k = ks*16 + gi*8 + e; rows n >= NJ are zero
__global__ void pw_tc_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wpack, int NI, int NJ, int NJp, int dgrad) {
  const int total = NI * NJp;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int e = idx & 7;
  int r = idx >> 3;
  int n = r % NJp; r /= NJp;
  int gi = r & 1;
  int ks = r >> 1;
  const int k = ks * 16 + gi * 8 + e;
  float v = 0.f;
  if (n < NJ) v = dgrad ? w[(size_t)k * NJ + n] : w[(size_t)n * NI + k];
  wpack[idx] = __float2bfloat16_rn(v);
}

// out[b, n, p] = bias[n] + sum_k Wt[n][k] in[b, k, p]      grid (ceil(HW/128), B), 128 threads
__global__ void __launch_bounds__(128) pw_tc_kernel(const float* __restrict__ in, const __nv_bfloat16* __restrict__ wpack,
                                                    const float* __restrict__ bias, float* __restrict__ out, int NI, int NJ, int NJp,
                                                    int HW, int tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int G = NI / 8, KS = NI / 16;
  uint8_t* a_base = smem;                                   // [G][128][16 B]
  uint8_t* w_base = smem + (size_t)G * PLANE;               // [KS][2][NJp][16 B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(w_base + (size_t)NI * NJp * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, p0 = blockIdx.x * TPX;
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)tmem_cols);
  {
    const uint4* src = reinterpret_cast<const uint4*>(wpack);
    uint4* dst = reinterpret_cast<uint4*>(w_base);
    for (int i = tid; i < NI * NJp * 2 / 16; i += 128) dst[i] = __ldg(src + i);
  }
  stage_planes(in, a_base, G, NI, HW, b, p0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, NJp);
    const uint32_t a0 = smem_u32(a_base), w0 = smem_u32(w_base);
    for (int ks = 0; ks < KS; ++ks) {
      const uint64_t adesc = make_smem_desc(a0 + (uint32_t)(ks * 2 * PLANE), PLANE, 128);
      const uint64_t bdesc = make_smem_desc(w0 + (uint32_t)(ks * 2 * NJp * 16), NJp * 16, 128);
      umma_bf16(tmem_base, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after();
  {
    const int p = p0 + warp * 32 + lane;
    float* op = out + (size_t)b * NJ * HW + p;
    for (int n0 = 0; n0 < NJp; n0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, v);
      tmem_ld_wait();
      if (p < HW) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int n = n0 + e;
          if (n < NJ) op[(size_t)n * HW] = __uint_as_float(v[e]) + (bias ? __ldg(bias + n) : 0.f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

// ---- persistent, software-pipelined variant of pw_tc_kernel for NI <= 64 (G = NI/8 in {2,4,8}): a CTA walks many tiles with
// the weights resident, two A buffers and two TMEM accumulators; the next tile's global loads are issued into registers before
// the current tile's epilogue stores, so load latency hides behind the store stream.
template <int G>
__global__ void __launch_bounds__(128) pw_tc_persistent_kernel(const float* __restrict__ in, const __nv_bfloat16* __restrict__ wpack,
                                                               const float* __restrict__ bias, float* __restrict__ out, int NJ, int NJp,
                                                               int HW, int ntiles, int tmem_cols) {
  constexpr int NI = G * 8, KS = G / 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* a_base = smem;                                   // [2][G][128][16 B]
  uint8_t* w_base = smem + (size_t)2 * G * PLANE;           // [KS][2][NJp][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + (size_t)NI * NJp * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)tmem_cols);
  {
    const uint4* src = reinterpret_cast<const uint4*>(wpack);
    uint4* dst = reinterpret_cast<uint4*>(w_base);
    for (int i = tid; i < NI * NJp * 2 / 16; i += 128) dst[i] = __ldg(src + i);
  }
  const int tiles_per_img = (HW + TPX - 1) / TPX;
  float f[G][8];
  auto fetch = [&](int tile) {
    const int b = tile / tiles_per_img, p = (tile - b * tiles_per_img) * TPX + tid;
    const float* sp = in + (size_t)b * NI * HW + p;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int e = 0; e < 8; ++e) f[g][e] = p < HW ? __ldg(sp + (size_t)(g * 8 + e) * HW) : 0.f;
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int g = 0; g < G; ++g)
      *reinterpret_cast<uint4*>(a_base + (size_t)(buf * G + g) * PLANE + (size_t)tid * 16) =
          make_uint4(pack_bf16x2(f[g][0], f[g][1]), pack_bf16x2(f[g][2], f[g][3]), pack_bf16x2(f[g][4], f[g][5]), pack_bf16x2(f[g][6], f[g][7]));
  };
  const uint32_t idesc = make_idesc_bf16(128, NJp);
  auto issue = [&](uint32_t tmem_base, int buf) {          // one thread
    const uint32_t a0 = smem_u32(a_base + (size_t)buf * G * PLANE), w0 = smem_u32(w_base);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const uint64_t adesc = make_smem_desc(a0 + (uint32_t)(ks * 2 * PLANE), PLANE, 128);
      const uint64_t bdesc = make_smem_desc(w0 + (uint32_t)(ks * 2 * NJp * 16), NJp * 16, 128);
      umma_bf16(tmem_base + (uint32_t)(buf * NJp), adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bars[buf]));
  };
  int tile = blockIdx.x;
  if (tile < ntiles) {
    fetch(tile);
    stash(0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tile < ntiles && tid == 0) issue(tmem_base, 0);
  uint32_t it = 0;
  for (; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int next = tile + gridDim.x;
    if (next < ntiles) fetch(next);                         // in flight during the epilogue below
    mbar_wait(smem_u32(&bars[buf]), (it >> 1) & 1u);
    tc_fence_after();
    {
      const int b = tile / tiles_per_img, p = (tile - b * tiles_per_img) * TPX + warp * 32 + lane;
      float* op = out + (size_t)b * NJ * HW + p;
      for (int n0 = 0; n0 < NJp; n0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * NJp + n0), v);
        tmem_ld_wait();
        if (p < HW) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int n = n0 + e;
            if (n < NJ) op[(size_t)n * HW] = __uint_as_float(v[e]) + (bias ? __ldg(bias + n) : 0.f);
          }
        }
      }
    }
    if (next < ntiles) {
      stash(buf ^ 1);                                       // that buffer's MMA (tile it-1) completed before its epilogue ran
      fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (next < ntiles && tid == 0) issue(tmem_base, buf ^ 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

// weight / bias gradient: D[o][c] = sum_p dy[o][p] x[c][p], both operands MN-major (K = pixels).  Column Cin of the N
// dimension is a constant-one plane, so D[o][Cin] = sum_p dy[o][p] = the bias gradient.  A CTA walks its share of the
// 128-pixel tiles, accumulating in TMEM the whole time; partial[cta][h][m][Np] for o = h*128 + m.
__global__ void __launch_bounds__(128) pw_wgrad_tc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ partial, int Cin, int Cout, int HW, int ntiles, int MH,
                                                          int Np, int tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int GA = MH * 16, GB = Np / 8;                      // A: dy groups padded to 128 rows per half; B: x groups + ones + zero group
  uint8_t* a_base = smem;                                   // [GA][128][16 B]
  uint8_t* b_base = smem + (size_t)GA * PLANE;              // [GB][128][16 B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_base + (size_t)GB * PLANE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (GA + GB) * PLANE / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)tmem_cols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc_bf16_mn(128, Np);
  const int tiles_per_img = (HW + TPX - 1) / TPX;
  const int gx = Cin / 8, gy = Cout / 8;
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * TPX;
    if (it > 0) {
      mbar_wait(smem_u32(bar), (it - 1) & 1u);              // the previous tile's MMAs have read shared memory
      tc_fence_after();
    }
    stage_planes(dy, a_base, gy, Cout, HW, b, p0);
    stage_planes(x, b_base, gx, Cin, HW, b, p0);
    {                                                       // ones plane: element 0 of group gx = 1 for valid pixels
      const int p = p0 + tid;
      *reinterpret_cast<uint4*>(b_base + (size_t)gx * PLANE + (size_t)tid * 16) = make_uint4(p < HW ? 0x00003F80u : 0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      const uint32_t a0 = smem_u32(a_base), b0 = smem_u32(b_base);
      for (int h = 0; h < MH; ++h) {
#pragma unroll
        for (int ks = 0; ks < TPX / 16; ++ks) {
          // MN-major: LBO = step between the two 8-pixel K groups of one instruction, SBO = step between 8-channel MN groups
          const uint64_t adesc = make_smem_desc(a0 + (uint32_t)(h * 16 * PLANE + ks * 256), 128, PLANE);
          const uint64_t bdesc = make_smem_desc(b0 + (uint32_t)(ks * 256), 128, PLANE);
          umma_bf16(tmem_base + (uint32_t)(h * Np), adesc, bdesc, idesc, (it > 0 || ks > 0) ? 1u : 0u);
        }
      }
      umma_commit(smem_u32(bar));
    }
  }
  if (it > 0) {
    mbar_wait(smem_u32(bar), (it - 1) & 1u);
    tc_fence_after();
    const int m = warp * 32 + lane;
    for (int h = 0; h < MH; ++h) {
      float* pp = partial + (((size_t)blockIdx.x * MH + h) * 128 + m) * Np;
      for (int n0 = 0; n0 < Np; n0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * Np + n0), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) pp[n0 + e] = __uint_as_float(v[e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

// ---- software-pipelined variant for layers with at most 10 channel groups in total (Cin + Cout <= 80, i.e. the 16-channel
// stages that hold most of the pixels): two shared-memory buffers; the next tile's dy / x values are fetched into registers
// while the tensor core contracts the current tile.
constexpr int WG_MAXG = 10;
__global__ void __launch_bounds__(128) pw_wgrad_tc_pipe_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ partial, int Cin, int Cout, int HW, int ntiles,
                                                               int Np, int tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int GA = 16, GB = Np / 8, GT = GA + GB;             // MH == 1 here
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)2 * GT * PLANE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * GT * PLANE / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)tmem_cols);
  const int tiles_per_img = (HW + TPX - 1) / TPX;
  const int gx = Cin / 8, gy = Cout / 8;
  float f[WG_MAXG][8];
  bool pvalid = false;
  auto fetch = [&](int tile) {
    const int b = tile / tiles_per_img, p = (tile - b * tiles_per_img) * TPX + tid;
    pvalid = p < HW;
    const float* yp = dy + (size_t)b * Cout * HW + p;
    const float* xp = x + (size_t)b * Cin * HW + p;
#pragma unroll
    for (int g = 0; g < WG_MAXG; ++g) {
      if (g < gy) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[g][e] = pvalid ? __ldg(yp + (size_t)(g * 8 + e) * HW) : 0.f;
      } else if (g < gy + gx) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[g][e] = pvalid ? __ldg(xp + (size_t)((g - gy) * 8 + e) * HW) : 0.f;
      }
    }
  };
  auto stash = [&](int buf) {
    uint8_t* a_base = smem + (size_t)buf * GT * PLANE;
    uint8_t* b_base = a_base + (size_t)GA * PLANE;
#pragma unroll
    for (int g = 0; g < WG_MAXG; ++g) {
      if (g < gy + gx) {
        uint8_t* dst = (g < gy ? a_base + (size_t)g * PLANE : b_base + (size_t)(g - gy) * PLANE) + (size_t)tid * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(f[g][0], f[g][1]), pack_bf16x2(f[g][2], f[g][3]),
                                                    pack_bf16x2(f[g][4], f[g][5]), pack_bf16x2(f[g][6], f[g][7]));
      }
    }
    *reinterpret_cast<uint4*>(b_base + (size_t)gx * PLANE + (size_t)tid * 16) = make_uint4(pvalid ? 0x00003F80u : 0u, 0u, 0u, 0u);
  };
  const uint32_t idesc = make_idesc_bf16_mn(128, Np);
  auto issue = [&](uint32_t tmem_base, int buf, bool first) {
    const uint32_t a0 = smem_u32(smem + (size_t)buf * GT * PLANE), b0 = a0 + (uint32_t)(GA * PLANE);
#pragma unroll
    for (int ks = 0; ks < TPX / 16; ++ks) {
      const uint64_t adesc = make_smem_desc(a0 + (uint32_t)(ks * 256), 128, PLANE);
      const uint64_t bdesc = make_smem_desc(b0 + (uint32_t)(ks * 256), 128, PLANE);
      umma_bf16(tmem_base, adesc, bdesc, idesc, (!first || ks > 0) ? 1u : 0u);
    }
    umma_commit(smem_u32(&bars[buf]));
  };
  int tile = blockIdx.x;
  __syncthreads();                                          // zero fill done before the first stash
  if (tile < ntiles) {
    fetch(tile);
    stash(0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tile < ntiles && tid == 0) issue(tmem_base, 0, true);
  uint32_t it = 0;
  for (; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int next = tile + gridDim.x;
    if (next < ntiles) {
      fetch(next);                                          // overlaps MMA(it)
      if (it >= 1) {                                        // buffer buf^1 was read by MMA(it-1)
        mbar_wait(smem_u32(&bars[buf ^ 1]), ((it - 1) >> 1) & 1u);
        tc_fence_after();
      }
      stash(buf ^ 1);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tid == 0) issue(tmem_base, buf ^ 1, false);
    }
  }
  if (it > 0) {
    // the last two MMAs (tiles it-1 and, if any, it-2) must be complete: waiting for the last one is enough (in-order pipe)
    mbar_wait(smem_u32(&bars[(it - 1) & 1]), ((it - 1) >> 1) & 1u);
    tc_fence_after();
    const int m = warp * 32 + lane;
    float* pp = partial + ((size_t)blockIdx.x * 128 + m) * Np;
    for (int n0 = 0; n0 < Np; n0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) pp[n0 + e] = __uint_as_float(v[e]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

// dW[o][c] = sum over CTAs; db[o] = column Cin.  CTAs with no tile wrote nothing: only the first `nparts` are summed.
__global__ void __launch_bounds__(256) pw_wreduce_tc_kernel(const float* __restrict__ partial, int nparts, int MH, int Np, int Cin,
                                                            int Cout, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float red[8][33];
  const int o_l = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + o_l;                    // over Cout * (Cin + 1)
  const int total = Cout * (Cin + 1);
  float s = 0.f;
  int o = 0, c = 0;
  if (idx < total) {
    o = idx / (Cin + 1);
    c = idx - o * (Cin + 1);
    const int h = o >> 7, m = o & 127;
    for (int k = sl; k < nparts; k += 8) s += partial[(((size_t)k * MH + h) * 128 + m) * Np + c];
  }
  red[sl][o_l] = s;
  __syncthreads();
  if (sl == 0 && idx < total) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][o_l];
    if (c < Cin) dw[(size_t)o * Cin + c] = t;
    else if (db) db[o] = t;
  }
}

// host-side launchers for pwconv_bwd_tc.cu (kernels cannot be launched across translation units without -rdc)
void launch_pack(const float* w, __nv_bfloat16* wpack, int NI, int NJ, int NJp, int dgrad, cudaStream_t st) {
  pw_tc_pack_kernel<<<cdiv(NI * NJp, 256), 256, 0, st>>>(w, wpack, NI, NJ, NJp, dgrad);
}
void launch_wreduce(const float* partial, int nparts, int MH, int Np, int Cin, int Cout, float* dw, float* db, cudaStream_t st) {
  pw_wreduce_tc_kernel<<<cdiv(Cout * (Cin + 1), 32), 256, 0, st>>>(partial, nparts, MH, Np, Cin, Cout, dw, db);
}

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}
static bool supported(const kmu_pwconv_desc& d) {
  return d.Cin % 16 == 0 && d.Cout % 16 == 0 && d.Cin >= 16 && d.Cin <= 256 && d.Cout >= 16 && d.Cout <= 256;
}
static bool wgrad_supported(const kmu_pwconv_desc& d) {
  // UMMA N = Cin + 16 (x groups + the ones group + a zero group) must stay <= 256, and all accumulators must fit the 512 TMEM columns
  return supported(d) && d.Cin + 16 <= 256 && cdiv(d.Cout, 128) * (d.Cin + 16) <= 512;
}
static int wgrad_ctas(const kmu_pwconv_desc& d) {
  long long tiles = (long long)d.B * cdiv(d.HW, TPX);
  return (int)(tiles < 296 ? tiles : 296);
}
struct Ws { size_t wpack, partial, total; };
static Ws ws_layout(const kmu_pwconv_desc& d) {
  Ws w;
  size_t o = 0;
  const int mx = d.Cin > d.Cout ? d.Cin : d.Cout;
  w.wpack = o; o += align_up((size_t)mx * mx * 2, 256);
  const int MH = cdiv(d.Cout, 128), Np = d.Cin + 16;
  w.partial = o; o += align_up((size_t)wgrad_ctas(d) * MH * 128 * Np * 4, 256);
  w.total = o;
  return w;
}

static int run_gemm(const float* in, const float* w, const float* bias, float* out, int NI, int NJ, int HW, int B, int dgrad,
                    __nv_bfloat16* wpack, cudaStream_t st) {
  const int NJp = (NJ + 15) / 16 * 16;
  pw_tc_pack_kernel<<<cdiv(NI * NJp, 256), 256, 0, st>>>(w, wpack, NI, NJ, NJp, dgrad);
  KMU_LAUNCH_CHECK("pw_tc_pack");
  if (NI == 16 || NI == 32 || NI == 64) {
    // persistent pipelined kernel: a few CTAs per SM, each walking its share of the tiles
    const int G = NI / 8, ntiles = B * cdiv(HW, TPX);
    const size_t smem = (size_t)2 * G * PLANE + (size_t)NI * NJp * 2 + 64;
    int per_sm = (int)(200 * 1024 / (smem + 1024));
    const int cols = pow2_cols(2 * NJp);
    if (per_sm > 512 / cols) per_sm = 512 / cols;
    if (per_sm > 6) per_sm = 6;
    if (per_sm < 1) per_sm = 1;
    int grid = 148 * per_sm;
    if (grid > ntiles) grid = ntiles;
#define KMU_PW_TC_P(GG)                                                                                                     \
  do {                                                                                                                      \
    cudaError_t e = cudaFuncSetAttribute(pw_tc_persistent_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "pw_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e)); \
    pw_tc_persistent_kernel<GG><<<grid, 128, smem, st>>>(in, wpack, bias, out, NJ, NJp, HW, ntiles, cols);                    \
  } while (0)
    if (G == 2) KMU_PW_TC_P(2);
    else if (G == 4) KMU_PW_TC_P(4);
    else KMU_PW_TC_P(8);
#undef KMU_PW_TC_P
  } else {
    const size_t smem = (size_t)(NI / 8) * PLANE + (size_t)NI * NJp * 2 + 64;
    cudaError_t e = cudaFuncSetAttribute(pw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "pw_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
    pw_tc_kernel<<<dim3(cdiv(HW, TPX), B), 128, smem, st>>>(in, wpack, bias, out, NI, NJ, NJp, HW, pow2_cols(NJp));
  }
  KMU_LAUNCH_CHECK(dgrad ? "pw_tc_dgrad" : "pw_tc_fwd");
  return KMU_OK;
}

}  // namespace pwtc
}  // namespace kmu

using namespace kmu;
using namespace kmu::pwtc;

extern "C" {

int kmu_pwconv_tc_supported(const kmu_pwconv_desc* d) { return (d && supported(*d)) ? 1 : 0; }
int kmu_pwconv_tc_wgrad_supported(const kmu_pwconv_desc* d) { return (d && wgrad_supported(*d)) ? 1 : 0; }

size_t kmu_pwconv_tc_workspace_bytes(const kmu_pwconv_desc* d) {
  if (!d || !supported(*d) || d->B <= 0 || d->HW <= 0) return 0;
  return ws_layout(*d).total;
}

int kmu_pwconv_tc_fwd(const kmu_pwconv_desc* d, const float* x, const float* w, const float* bias, float* y, void* workspace,
                      size_t workspace_bytes, kmu_stream stream) {
  KMU_REQUIRE(d && supported(*d) && d->B > 0 && d->HW > 0 && d->B <= 65535, KMU_ERR_UNSUPPORTED, "pwconv_tc_fwd: unsupported shape");
  KMU_REQUIRE(x && w && y, KMU_ERR_BAD_ARG, "pwconv_tc_fwd: null tensor");
  const Ws wl = ws_layout(*d);
  KMU_REQUIRE(workspace && workspace_bytes >= wl.total, KMU_ERR_WORKSPACE, "pwconv_tc_fwd: workspace too small");
  return run_gemm(x, w, bias, y, d->Cin, d->Cout, d->HW, d->B, 0, (__nv_bfloat16*)((char*)workspace + wl.wpack), (cudaStream_t)stream);
}

int kmu_pwconv_tc_bwd(const kmu_pwconv_desc* d, const float* x, const float* dy, const float* w, float* dx, float* dw, float* dbias,
                      void* workspace, size_t workspace_bytes, kmu_stream stream) {
  KMU_REQUIRE(d && supported(*d) && d->B > 0 && d->HW > 0 && d->B <= 65535, KMU_ERR_UNSUPPORTED, "pwconv_tc_bwd: unsupported shape");
  KMU_REQUIRE(dy && w, KMU_ERR_BAD_ARG, "pwconv_tc_bwd: null tensor");
  const Ws wl = ws_layout(*d);
  KMU_REQUIRE(workspace && workspace_bytes >= wl.total, KMU_ERR_WORKSPACE, "pwconv_tc_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) {
    int rc = run_gemm(dy, w, nullptr, dx, d->Cout, d->Cin, d->HW, d->B, 1, (__nv_bfloat16*)((char*)workspace + wl.wpack), st);
    if (rc != KMU_OK) return rc;
  }
  if (dw) {
    KMU_REQUIRE(x != nullptr, KMU_ERR_BAD_ARG, "pwconv_tc_bwd: weight gradient needs x");
    KMU_REQUIRE(wgrad_supported(*d), KMU_ERR_UNSUPPORTED, "pwconv_tc_bwd: weight gradient needs Cin <= 240 (got %d -> %d)", d->Cin, d->Cout);
    const int MH = cdiv(d->Cout, 128), Np = d->Cin + 16;
    const int ctas = wgrad_ctas(*d), ntiles = d->B * cdiv(d->HW, TPX);
    float* partial = (float*)((char*)workspace + wl.partial);
    if (MH == 1 && (d->Cin + d->Cout) / 8 <= WG_MAXG) {
      const size_t smem = (size_t)2 * (16 + Np / 8) * PLANE + 64;
      cudaError_t e = cudaFuncSetAttribute(pw_wgrad_tc_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "pw_wgrad_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
      pw_wgrad_tc_pipe_kernel<<<ctas, 128, smem, st>>>(x, dy, partial, d->Cin, d->Cout, d->HW, ntiles, Np, pow2_cols(Np));
    } else {
      const size_t smem = (size_t)(MH * 16 + Np / 8) * PLANE + 64;
      cudaError_t e = cudaFuncSetAttribute(pw_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "pw_wgrad_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
      pw_wgrad_tc_kernel<<<ctas, 128, smem, st>>>(x, dy, partial, d->Cin, d->Cout, d->HW, ntiles, MH, Np, pow2_cols(MH * Np));
    }
    KMU_LAUNCH_CHECK("pw_wgrad_tc");
    const int total = d->Cout * (d->Cin + 1);
    pw_wreduce_tc_kernel<<<cdiv(total, 32), 256, 0, st>>>(partial, ctas, MH, Np, d->Cin, d->Cout, dw, dbias);
    KMU_LAUNCH_CHECK("pw_wreduce_tc");
  }
  return KMU_OK;
}

}  // extern "C"
