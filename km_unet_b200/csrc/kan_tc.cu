// kan_tc.cu -- tcgen05 / TMEM / bulk-TMA family of KANConv2d forward (KMU_PREC_BF16), 3x3 / stride 1 / padding 1.
//
// Replaces convKAN/KANConv2Dlayers.py:15-37 + convKAN/KANlayers.py:577-610,644-660 for the shapes KM-UNet uses.
// Formulation (checked by oracle.kan.kanconv2d_as_phi_conv): the layer is a 3x3 convolution over the
// per-pixel expansion Phi(x) = [SiLU(x), B_0(x)..B_7(x)] with Phi(0) at the border, i.e. an implicit GEMM
//     Y[pixel, o] = sum_{tap, c, q} Phi_q(x[c, pixel + tap]) * Wfull[o, c, tap, q],      K = 81 * Cin.
// Phi is evaluated ONCE per input pixel by CUDA-core producer warps straight into shared memory in the UMMA canonical
// no-swizzle K-major layout; it never exists in HBM.  Because a K-group of 8 bf16 (16 bytes) per pixel is exactly one
// core-matrix row, the nine taps are nine SHIFTED VIEWS of the same shared-memory planes: the A descriptor of tap (ki,kj)
// just starts (ki*pitch + kj) * 16 bytes later.  No im2col, no 9x recomputation.
//
// CTA = one 8-wide strip of 16*TT output rows of one image (TT accumulators of 128 pixels x Cout in TMEM, double
// buffered -> 2*TT*Cout <= 512 columns).  Warp roles (448 threads, 1 CTA/SM, persistent over strips):
//     warps 0-7   Phi producers: x (fp32, global) -> SiLU / cubic B-spline closed form -> bf16 K-group planes (ring)
//     warp  8     MMA issuer (one thread): tcgen05.mma.cta_group::1.kind::f16, M=128, N=Cout, K=16 per instruction
//     warp  9     weight loader (one thread): cp.async.bulk global->shared of pre-packed bf16 weight blocks (ring)
//     warps 10-13 epilogue: tcgen05.ld accumulators -> fp32 NCHW stores, overlapped with the next strip's MMAs
// mbarrier pipelines: phi full/empty, weight full/empty (tx-count), accumulator full/empty.
//
// K order: per 16-channel block, slots 0..7 = spline groups of channel pairs (ring of Phi slots), slot 8 = the two SiLU
// groups (8 channels each) which the producers fill 2 bytes at a time while they evaluate the splines of the same x
// (double-buffered side region).  Each slot is used by 9 taps x TT tiles = 9*TT MMAs, then released.  x values are
// prefetched into registers one slot ahead so the global-load latency overlaps the previous slot's math.  Weights are
// re-streamed from L2 per strip; TT=4 keeps that at ~2.7 TB/s aggregate for the 64->64 microbench.
#include "common.cuh"
#include "kan_common.cuh"
#include "tc_common.cuh"

namespace kmu {
namespace kan {
namespace tc {

using namespace kmu::tcx;

constexpr int PITCH = 10;  // 8-wide strip + 1-pixel halo each side (positions per halo row)
constexpr int NUM_PRODUCER_WARPS = 8;
constexpr int WARP_MMA = 8, WARP_LOAD = 9, WARP_EPI0 = 10;
constexpr int NUM_THREADS = 14 * 32;
constexpr int W_STAGES = 4;
constexpr int SMEM_BUDGET = 200 * 1024;

struct TcDims {
  int B, Cin, H, W;
  int tiles_x, tiles_y, num_tiles;
  int slots;       // K slots per strip = 9 * Cin / 16
  int phi_stages;  // ring depth of Phi slots
  float t0, inv_h; // uniform knots t_j = t0 + j*h
};

// ------------------------------------------------------------------------------------------------ weight packing
// wpack[slot][tap][gi][n][e] (bf16): the shared-memory image of every (slot, tap) B block, K-major no-swizzle:
// 8 output rows x 16 B core matrices, SBO = 128 B (next 8 rows), LBO = N*16 B (second K-group).
__global__ void kan_tc_pack_kernel(const float* __restrict__ base_w, const float* __restrict__ spline_w,
                                   const float* __restrict__ scaler, __nv_bfloat16* __restrict__ wpack, int Cin, int N) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)81 * Cin * N;
  if (idx >= total) return;
  int e = (int)(idx & 7);
  long long r = idx >> 3;
  int n = (int)(r % N);
  r /= N;
  int gi = (int)(r & 1);
  r >>= 1;
  int tap = (int)(r % 9);
  int slot = (int)(r / 9);
  int cb = slot / 9, j = slot - cb * 9;
  int F = Cin * 9;
  float v;
  if (j == 8) {
    int c = cb * 16 + gi * 8 + e;
    v = base_w[(size_t)n * F + c * 9 + tap];
  } else {
    int c = cb * 16 + 2 * j + gi;
    size_t of = (size_t)n * F + c * 9 + tap;
    v = spline_w[of * NB + e] * (scaler ? scaler[of] : 1.0f);
  }
  wpack[idx] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------ main kernel
template <int N, int TT>
__global__ void __launch_bounds__(NUM_THREADS, 1) kan_fwd_tc_kernel(const float* __restrict__ x,
                                                                    const __nv_bfloat16* __restrict__ wpack,
                                                                    float* __restrict__ y, TcDims d) {
  constexpr int NPOS = PITCH * (16 * TT + 2);
  constexpr int PLANE = NPOS * 16;
  constexpr int SLOT = 2 * PLANE;
  constexpr int WSTAGE = 9 * 2 * N * 16;
  constexpr uint32_t IDESC = make_idesc_bf16(128, N);
  static_assert(2 * TT * N <= 512, "accumulators exceed TMEM");
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int UPT = (2 * NPOS + NUM_PRODUCER_WARPS * 32 - 1) / (NUM_PRODUCER_WARPS * 32);  // units per producer thread
  const int R = d.phi_stages;
  uint8_t* phi_base = smem;
  uint8_t* silu_base = smem + (size_t)R * SLOT;            // 2 x SLOT
  uint8_t* w_base = silu_base + 2 * (size_t)SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + (size_t)W_STAGES * WSTAGE);
  uint64_t* phi_full = bars;
  uint64_t* phi_empty = bars + R;
  uint64_t* w_full = bars + 2 * R;
  uint64_t* w_empty = w_full + W_STAGES;
  uint64_t* acc_full = w_empty + W_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* silu_full = acc_empty + 2;
  uint64_t* silu_empty = silu_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(silu_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < R; ++i) {
      mbar_init(smem_u32(&phi_full[i]), NUM_PRODUCER_WARPS);
      mbar_init(smem_u32(&phi_empty[i]), 1);
    }
    for (int i = 0; i < W_STAGES; ++i) {
      mbar_init(smem_u32(&w_full[i]), 1);
      mbar_init(smem_u32(&w_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), 4);
      mbar_init(smem_u32(&silu_full[i]), NUM_PRODUCER_WARPS);
      mbar_init(smem_u32(&silu_empty[i]), 1);
    }
    fence_barrier_init();
  }
  if (warp == WARP_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const size_t HW = (size_t)d.H * d.W;
  const int tiles_per_img = d.tiles_x * d.tiles_y;

  if (warp < NUM_PRODUCER_WARPS) {
    // ===================================================================== Phi producers
    uint32_t it = 0, bc = 0;  // ring iteration, 16-channel block counter
    const int pairs = d.Cin / 2;
    for (int tile = blockIdx.x; tile < d.num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty0 = (tr / d.tiles_x) * (16 * TT), tx0 = (tr % d.tiles_x) * 8;
      const float* xb = x + (size_t)b * d.Cin * HW;
      // per-thread units: (halo position, which channel of the pair); fixed for the whole strip
      int uoff[UPT];   // pixel offset inside the image plane, -1 = zero padding, -2 = no unit
      int ugi[UPT], upos[UPT];
#pragma unroll
      for (int k = 0; k < UPT; ++k) {
        const int u = tid + k * (NUM_PRODUCER_WARPS * 32);
        const int gi = u >= NPOS ? 1 : 0;
        const int pos = u - gi * NPOS;
        const int py = pos / PITCH, px = pos - py * PITCH;
        const int gy = ty0 - 1 + py, gx = tx0 - 1 + px;
        ugi[k] = gi;
        upos[k] = pos;
        uoff[k] = (u >= 2 * NPOS) ? -2 : ((gy >= 0 && gy < d.H && gx >= 0 && gx < d.W) ? gy * d.W + gx : -1);
      }
      float xr[UPT];
#pragma unroll
      for (int k = 0; k < UPT; ++k) xr[k] = uoff[k] >= 0 ? __ldg(xb + (size_t)ugi[k] * HW + uoff[k]) : 0.f;
      for (int ps = 0; ps < pairs; ++ps, ++it) {
        float xn[UPT];
        if (ps + 1 < pairs) {                      // prefetch the next channel pair: latency overlaps this slot's math
          const float* xc = xb + (size_t)(2 * (ps + 1)) * HW;
#pragma unroll
          for (int k = 0; k < UPT; ++k) xn[k] = uoff[k] >= 0 ? __ldg(xc + (size_t)ugi[k] * HW + uoff[k]) : 0.f;
        } else {
#pragma unroll
          for (int k = 0; k < UPT; ++k) xn[k] = 0.f;
        }
        const int j = ps & 7;
        const uint32_t sb = bc & 1u;
        if (j == 0) mbar_wait(smem_u32(&silu_empty[sb]), ((bc >> 1) & 1u) ^ 1u);
        const int r = it % R;
        mbar_wait(smem_u32(&phi_empty[r]), ((it / R) & 1u) ^ 1u);
        uint8_t* slot = phi_base + (size_t)r * SLOT;
        uint8_t* sslot = silu_base + (size_t)sb * SLOT;
#pragma unroll
        for (int k = 0; k < UPT; ++k) {
          if (uoff[k] != -2) {
            const uint4 v = spline_group(xr[k], d.t0, d.inv_h);
            *reinterpret_cast<uint4*>(slot + (size_t)ugi[k] * PLANE + (size_t)upos[k] * 16) = v;
            const int cl = 2 * j + ugi[k];          // channel inside the 16-channel block
            const __nv_bfloat16 sv = __float2bfloat16_rn(silu_fast(xr[k]));
            *reinterpret_cast<__nv_bfloat16*>(sslot + (size_t)(cl >> 3) * PLANE + (size_t)upos[k] * 16 + (cl & 7) * 2) = sv;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&phi_full[r]));
          if (j == 7) mbar_arrive(smem_u32(&silu_full[sb]));
        }
        if (j == 7) ++bc;
#pragma unroll
        for (int k = 0; k < UPT; ++k) xr[k] = xn[k];
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      uint32_t it = 0, wit = 0, bc = 0, tcount = 0;
      const int blocks = d.Cin / 16;
      for (int tile = blockIdx.x; tile < d.num_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount & 1u;
        mbar_wait_hot(smem_u32(&acc_empty[buf]), ((tcount >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256u;
        for (int cb = 0; cb < blocks; ++cb, ++bc) {
          for (int j = 0; j < 9; ++j, ++wit) {
            const int ws = wit % W_STAGES;
            uint32_t a0;
            int r = 0;
            const uint32_t sb = bc & 1u;
            if (j < 8) {
              r = it % R;
              mbar_wait_hot(smem_u32(&phi_full[r]), (it / R) & 1u);
              a0 = smem_u32(phi_base + (size_t)r * SLOT);
            } else {
              mbar_wait_hot(smem_u32(&silu_full[sb]), (bc >> 1) & 1u);
              a0 = smem_u32(silu_base + (size_t)sb * SLOT);
            }
            mbar_wait_hot(smem_u32(&w_full[ws]), (wit / W_STAGES) & 1u);
            tc_fence_after();
            const uint64_t bdesc0 = make_smem_desc(smem_u32(w_base + (size_t)ws * WSTAGE), N * 16, 128);
            const uint64_t adesc0 = make_smem_desc(a0, PLANE, PITCH * 16);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const int ki = t / 3, kj = t - ki * 3;
              const uint64_t bdesc = desc_advance(bdesc0, (uint32_t)t * (2 * N * 16));
#pragma unroll
              for (int i = 0; i < TT; ++i) {
                const uint64_t adesc = desc_advance(adesc0, (uint32_t)(((i * 16 + ki) * PITCH + kj) * 16));
                umma_bf16(d_tmem + (uint32_t)(i * N), adesc, bdesc, IDESC, (cb > 0 || j > 0 || t > 0) ? 1u : 0u);
              }
            }
            if (j < 8) {
              umma_commit(smem_u32(&phi_empty[r]));
              ++it;
            } else {
              umma_commit(smem_u32(&silu_empty[sb]));
            }
            umma_commit(smem_u32(&w_empty[ws]));
          }
        }
        umma_commit(smem_u32(&acc_full[buf]));
      }
    }
  } else if (warp == WARP_LOAD) {
    // ===================================================================== weight loader
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < d.num_tiles; tile += gridDim.x) {
        for (int s = 0; s < d.slots; ++s, ++it) {
          const int ws = it % W_STAGES;
          mbar_wait(smem_u32(&w_empty[ws]), ((it / W_STAGES) & 1u) ^ 1u);
          const uint32_t bar = smem_u32(&w_full[ws]);
          mbar_expect_tx(bar, WSTAGE);
          bulk_g2s(smem_u32(w_base + (size_t)ws * WSTAGE), reinterpret_cast<const uint8_t*>(wpack) + (size_t)s * WSTAGE, WSTAGE,
                   bar);
        }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 10..13)
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < d.num_tiles; tile += gridDim.x, ++tcount) {
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty0 = (tr / d.tiles_x) * (16 * TT), tx0 = (tr % d.tiles_x) * 8;
      const uint32_t buf = tcount & 1u;
      mbar_wait(smem_u32(&acc_full[buf]), (tcount >> 1) & 1u);
      tc_fence_after();
      const int m = q * 32 + lane;  // accumulator row = pixel of the tile
      const int gx = tx0 + (m & 7);
#pragma unroll
      for (int i = 0; i < TT; ++i) {
        const int gy = ty0 + i * 16 + (m >> 3);
        const bool ok = gy < d.H && gx < d.W;
        float* yp = y + ((size_t)b * N) * HW + (size_t)gy * d.W + gx;
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u + (uint32_t)(i * N + c0), v);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int e = 0; e < 16; ++e) yp[(size_t)(c0 + e) * HW] = __uint_as_float(v[e]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static size_t smem_bytes(int N, int TT, int R) {
  size_t npos = (size_t)PITCH * (16 * TT + 2);
  return (size_t)(R + 2) * 2 * npos * 16 + (size_t)W_STAGES * 9 * 2 * N * 16 + (size_t)(2 * R + 2 * W_STAGES + 8) * 8 + 16;
}
static int pick_phi_stages(int N, int TT) {
  int r = 8;
  while (r > 2 && smem_bytes(N, TT, r) > (size_t)SMEM_BUDGET) --r;
  return r;
}

bool supported(const kmu_kanconv2d_desc& s) {
  return s.ksize == 3 && s.stride == 1 && s.padding == 1 && s.spline_order == 3 && s.grid_size == 5 && s.Cin % 16 == 0 &&
         (s.Cout == 16 || s.Cout == 32 || s.Cout == 64) && s.grid_uniform != 0;
}

size_t fwd_workspace(const Dims& d) { return align_up((size_t)81 * d.Cin * d.Cout * 2, 256); }

template <int N, int TT>
static int launch(const kmu_kanconv2d_fwd_args* a, const Dims& d, const __nv_bfloat16* wpack, cudaStream_t st) {
  TcDims t;
  t.B = d.B; t.Cin = d.Cin; t.H = d.H; t.W = d.W;
  t.tiles_x = cdiv(d.W, 8);
  t.tiles_y = cdiv(d.H, 16 * TT);
  t.num_tiles = d.B * t.tiles_x * t.tiles_y;
  t.slots = 9 * d.Cin / 16;
  t.phi_stages = pick_phi_stages(N, TT);
  t.t0 = a->d.grid_t0;
  t.inv_h = 1.0f / a->d.grid_h;
  size_t smem = smem_bytes(N, TT, t.phi_stages);
  cudaError_t e = cudaFuncSetAttribute(kan_fwd_tc_kernel<N, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "kan_fwd_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = t.num_tiles < sms ? t.num_tiles : sms;
  kan_fwd_tc_kernel<N, TT><<<grid, NUM_THREADS, smem, st>>>(a->x, wpack, a->y, t);
  KMU_LAUNCH_CHECK("kan_fwd_tc");
  return KMU_OK;
}

template <int N>
static int launch_n(const kmu_kanconv2d_fwd_args* a, const Dims& d, const __nv_bfloat16* wpack, cudaStream_t st) {
  // strips of 16*TT rows: take the tallest strip that still gives every SM about two strips of work
  long long m_tiles = (long long)d.B * cdiv(d.W, 8) * cdiv(d.H, 16);
  if (N <= 64 && d.H >= 64 && m_tiles >= 4LL * 2 * 148) return launch<N, 4>(a, d, wpack, st);
  if (d.H >= 32 && m_tiles >= 2LL * 2 * 148) return launch<N, 2>(a, d, wpack, st);
  return launch<N, 1>(a, d, wpack, st);
}

int forward(const kmu_kanconv2d_fwd_args* a, const Dims& d, cudaStream_t st) {
  KMU_REQUIRE(a->workspace && a->workspace_bytes >= fwd_workspace(d), KMU_ERR_WORKSPACE, "kanconv2d_fwd(tc): workspace %zu < %zu",
              a->workspace_bytes, fwd_workspace(d));
  KMU_REQUIRE(a->d.grid_h > 0.f, KMU_ERR_BAD_ARG, "kanconv2d_fwd(tc): grid_h must be positive");
  __nv_bfloat16* wpack = (__nv_bfloat16*)a->workspace;
  long long total = (long long)81 * d.Cin * d.Cout;
  kan_tc_pack_kernel<<<cdiv(total, 256), 256, 0, st>>>(a->base_weight, a->spline_weight, a->spline_scaler, wpack, d.Cin, d.Cout);
  KMU_LAUNCH_CHECK("kan_tc_pack");
  switch (d.Cout) {
    case 16: return launch_n<16>(a, d, wpack, st);
    case 32: return launch_n<32>(a, d, wpack, st);
    case 64: return launch_n<64>(a, d, wpack, st);
  }
  set_error("kanconv2d_fwd(tc): unsupported Cout %d", d.Cout);
  return KMU_ERR_UNSUPPORTED;
}

}  // namespace tc
}  // namespace kan
}  // namespace kmu
