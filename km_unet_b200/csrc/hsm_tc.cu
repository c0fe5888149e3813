// hsm_tc.cu -- tcgen05 / TMEM path of the HSM-SSD projection (KMU_PREC_BF16): P = dw3x3(Wp x) as ONE dense 3x3 convolution.
//
// vim_block_init/efficient_vim_init.py:39 computes BCdt = dw(BCdt_proj(x)): a 1x1 projection C -> 192 followed by a depthwise
// 3x3.  Neither layer has a bias or a non-linearity, so the pair is exactly the dense convolution
//     P[n, p] = sum_{tap, c} (wd[n, tap] * Wp[n, c]) x[c, p + tap]            K = 9 C, zero padding of x == zero padding of Wp x,
// i.e. the same shifted-window implicit GEMM as the KAN convolution (kan_tc.cu) with the identity in place of Phi: the x tile
// (+1-pixel halo) is converted once to bf16 K-group planes in shared memory in the UMMA canonical no-swizzle K-major layout,
// the nine taps are nine descriptor start offsets into those planes, accumulators live in TMEM.  The 192 outputs go as three
// 64-column slices (blockIdx.z), so a CTA needs 64 TMEM columns and 9*C*64*2 bytes of weights: several CTAs are resident per
// SM and hide each other's staging / epilogue latency -- no warp specialisation, no rings.
// CTA = 128 threads = one 8-wide x 16-row pixel tile (M = 128) of one image and one 64-channel slice.
#include "common.cuh"
#include "tc_common.cuh"

namespace kmu {
namespace hsm {
namespace tc {

using namespace kmu::tcx;

constexpr int PITCH = 10, ROWS = 18, NPOS = PITCH * ROWS;  // 8x16 tile + 1-pixel halo
constexpr int PLANE = NPOS * 16;                           // bytes of one K-group plane (8 channels x bf16 per position)
constexpr int NS = 64;                                     // output channels per CTA of the projection
constexpr int WBLK = 2 * NS * 16;                          // bytes of one (tap, k-step) weight block [gi][n][8]

// wpack[slice][tap][ks][gi][n][e] = bf16( wd[slice*64+n][tap] * Wp[slice*64+n][ks*16 + gi*8 + e] )
__global__ void hsm_tc_pack_kernel(const float* __restrict__ wp, const float* __restrict__ wd, __nv_bfloat16* __restrict__ wpack, int C) {
  const int KS = C / 16;
  const int total = 192 * 9 * C;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int e = idx & 7;
  int r = idx >> 3;
  int n = r % NS; r /= NS;
  int gi = r & 1; r >>= 1;
  int ks = r % KS; r /= KS;
  int tap = r % 9;
  int slice = r / 9;
  const int ng = slice * NS + n, c = ks * 16 + gi * 8 + e;
  wpack[idx] = __float2bfloat16_rn(wd[ng * 9 + tap] * wp[(size_t)ng * C + c]);
}

// Generic form: out[b, z * oz_step + n, p] = sum_{tap, c} w[b][z][n, tap, c] x[b, c, p + tap], n < NSL.  The projection uses
// NSL = 64 with weights shared by the batch (wb_stride = 0) and z = the Bm / dt slices; the output contraction y = ho Cm uses
// NSL = C with per-batch weights (see out() below).
template <int C, int NSL>
__global__ void __launch_bounds__(128) hsm_proj_tc_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ wpack,
                                                          float* __restrict__ P, int H, int tiles_x, int out_ch, int oz_step,
                                                          size_t wz_stride, size_t wb_stride) {
  constexpr int G = C / 8, KS = C / 16;
  constexpr int NS = NSL, WBLK = 2 * NSL * 16, TCOLS = NSL < 32 ? 32 : NSL;
  constexpr uint32_t IDESC = make_idesc_bf16(128, NS);
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* a_base = smem;                               // [G][NPOS][16 B]
  uint8_t* w_base = smem + (size_t)G * PLANE;           // [9][KS][WBLK]
  uint64_t* bar = reinterpret_cast<uint64_t*>(w_base + (size_t)9 * KS * WBLK);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = H * H;
  const int b = blockIdx.y, slice = blockIdx.z;
  const int ty0 = (blockIdx.x / tiles_x) * 16, tx0 = (blockIdx.x % tiles_x) * 8;

  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TCOLS);
  // weights of this slice: plain 16-byte copies (L2-resident, shared by every CTA of the slice)
  {
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(wpack) + (size_t)slice * wz_stride + (size_t)b * wb_stride);
    uint4* dst = reinterpret_cast<uint4*>(w_base);
    for (int i = tid; i < 9 * KS * WBLK / 16; i += 128) dst[i] = __ldg(src + i);
  }
  // x tile + halo -> bf16 K-group planes (zero outside the image)
  const float* xb = x + (size_t)b * C * L;
  for (int u = tid; u < G * NPOS; u += 128) {
    const int g = u / NPOS, pos = u - g * NPOS;
    const int py = pos / PITCH, px = pos - py * PITCH;
    const int gy = ty0 - 1 + py, gx = tx0 - 1 + px;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (gy >= 0 && gy < H && gx >= 0 && gx < H) {
      const float* xp = xb + (size_t)(g * 8) * L + (size_t)gy * H + gx;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __ldg(xp + (size_t)e * L);
      v = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
    *reinterpret_cast<uint4*>(a_base + (size_t)g * PLANE + (size_t)pos * 16) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    const uint64_t adesc0 = make_smem_desc(smem_u32(a_base), PLANE, PITCH * 16);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(w_base), NS * 16, 128);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int ki = t / 3, kj = t - ki * 3;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint64_t adesc = desc_advance(adesc0, (uint32_t)(ks * 2 * PLANE + (ki * PITCH + kj) * 16));
        const uint64_t bdesc = desc_advance(bdesc0, (uint32_t)((t * KS + ks) * WBLK));
        umma_bf16(tmem_base, adesc, bdesc, IDESC, (t > 0 || ks > 0) ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after();
  // epilogue: TMEM lane = pixel of the tile (m & 7 = column, m >> 3 = row), 64 fp32 columns = this slice's channels
  {
    const int m = warp * 32 + lane;
    const int gy = ty0 + (m >> 3), gx = tx0 + (m & 7);
    const bool ok = gy < H && gx < H;
    float* pp = P + ((size_t)b * out_ch + slice * oz_step) * L + (size_t)gy * H + gx;
#pragma unroll
    for (int c0 = 0; c0 < NS; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int e = 0; e < 16; ++e) pp[(size_t)(c0 + e) * L] = __uint_as_float(v[e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

// weff[b][tap][ks][gi][n][e] = bf16( sum_m ho[b, n, m] wd[64 + m, tap] Wp[64 + m, ks*16 + gi*8 + e] ): y = ho Cm with
// Cm = conv3x3(x; W rows 64..127) is itself a 3x3 convolution of x with these per-batch C x C weights (everything is linear),
// so the Cm slice of P never has to exist.  grid (9 taps, B), 256 threads; ho[b] and the tap's weight rows staged in shared memory.
__global__ void __launch_bounds__(256) hsm_tc_weff_kernel(const float* __restrict__ ho, const float* __restrict__ wp,
                                                          const float* __restrict__ wd, __nv_bfloat16* __restrict__ weff, int C) {
  extern __shared__ float sm[];
  float* ho_s = sm;             // [C][65]
  float* w_s = sm + C * 65;     // [C][65]: w_s[c][m] = wd[64 + m, tap] Wp[64 + m, c]
  const int tap = blockIdx.x, b = blockIdx.y, KS = C / 16;
  for (int i = threadIdx.x; i < C * 64; i += 256) {
    const int r = i >> 6, m = i & 63;
    ho_s[r * 65 + m] = ho[(size_t)b * C * 64 + i];
  }
  for (int i = threadIdx.x; i < C * 64; i += 256) {
    const int m = i / C, c = i - m * C;              // coalesced along c (rows of Wp)
    w_s[c * 65 + m] = wd[(64 + m) * 9 + tap] * wp[(size_t)(64 + m) * C + c];
  }
  __syncthreads();
  __nv_bfloat16* wb = weff + ((size_t)b * 9 + tap) * C * C;
  for (int o = threadIdx.x; o < C * C; o += 256) {  // o = ((ks*2 + gi)*C + n)*8 + e
    const int e = o & 7, n = (o >> 3) % C, kg = o / (8 * C);
    const int c = kg * 8 + e;                          // = ks*16 + gi*8 + e
    const float* hr = ho_s + n * 65;
    const float* wr = w_s + c * 65;
    float s = 0.f;
#pragma unroll 8
    for (int m = 0; m < 64; ++m) s = fmaf(hr[m], wr[m], s);
    wb[o] = __float2bfloat16_rn(s);
  }
  (void)KS;
}

size_t pack_bytes(int C) { return align_up((size_t)192 * 9 * C * 2, 256); }

// slices: 3 = Bm, Cm, dt (everything materialised); 2 = Bm and dt only (the Cm slice is folded into out() / the backward's
// correlation and is neither written nor read)
template <int C>
static int launch(const float* x, const __nv_bfloat16* wpack, float* P, int B, int H, int slices, cudaStream_t st) {
  constexpr int G = C / 8, KS = C / 16;
  const size_t smem = (size_t)G * PLANE + (size_t)9 * KS * WBLK + 64;
  cudaError_t e = cudaFuncSetAttribute(hsm_proj_tc_kernel<C, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "hsm_proj_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
  const int tiles_x = cdiv(H, 8), tiles_y = cdiv(H, 16);
  const size_t wslice = (size_t)9 * KS * WBLK;
  if (slices == 3)
    hsm_proj_tc_kernel<C, NS><<<dim3(tiles_x * tiles_y, B, 3), 128, smem, st>>>(x, wpack, P, H, tiles_x, 192, NS, wslice, 0);
  else
    hsm_proj_tc_kernel<C, NS><<<dim3(tiles_x * tiles_y, B, 2), 128, smem, st>>>(x, wpack, P, H, tiles_x, 192, 2 * NS, 2 * wslice, 0);
  KMU_LAUNCH_CHECK("hsm_proj_tc");
  return KMU_OK;
}

template <int C>
static int launch_out(const float* x, const __nv_bfloat16* weff, float* y, int B, int H, cudaStream_t st) {
  constexpr int G = C / 8, KS = C / 16;
  const size_t wbytes = (size_t)9 * KS * 2 * C * 16;
  const size_t smem = (size_t)G * PLANE + wbytes + 64;
  cudaError_t e = cudaFuncSetAttribute(hsm_proj_tc_kernel<C, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  KMU_REQUIRE(e == cudaSuccess, KMU_ERR_LAUNCH, "hsm_out_tc: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e));
  const int tiles_x = cdiv(H, 8), tiles_y = cdiv(H, 16);
  hsm_proj_tc_kernel<C, C><<<dim3(tiles_x * tiles_y, B, 1), 128, smem, st>>>(x, weff, y, H, tiles_x, C, 0, 0, wbytes);
  KMU_LAUNCH_CHECK("hsm_out_tc");
  return KMU_OK;
}

size_t weff_bytes(int B, int C) { return align_up((size_t)B * 9 * C * C * 2, 256); }

// y[b] = ho[b] Cm[b] = conv3x3(x[b]; ho[b] (x) W_Cm) on tensor cores.  workspace >= weff_bytes(B, C).
int out(const float* x, const float* ho, const float* wp, const float* wd, float* y, int B, int C, int H, void* workspace,
        cudaStream_t st) {
  __nv_bfloat16* weff = (__nv_bfloat16*)workspace;
  hsm_tc_weff_kernel<<<dim3(9, B), 256, (size_t)2 * C * 65 * sizeof(float), st>>>(ho, wp, wd, weff, C);
  KMU_LAUNCH_CHECK("hsm_tc_weff");
  switch (C) {
    case 16: return launch_out<16>(x, weff, y, B, H, st);
    case 32: return launch_out<32>(x, weff, y, B, H, st);
    case 64: return launch_out<64>(x, weff, y, B, H, st);
  }
  set_error("hsm_out_tc: unsupported C=%d", C);
  return KMU_ERR_UNSUPPORTED;
}

// P = conv3x3(x; wd (x) Wp) on tensor cores.  workspace >= pack_bytes(C).
int project(const float* x, const float* wp, const float* wd, float* P, int B, int C, int H, int slices, void* workspace,
            cudaStream_t st) {
  __nv_bfloat16* wpack = (__nv_bfloat16*)workspace;
  const int total = 192 * 9 * C;
  hsm_tc_pack_kernel<<<cdiv(total, 256), 256, 0, st>>>(wp, wd, wpack, C);
  KMU_LAUNCH_CHECK("hsm_tc_pack");
  switch (C) {
    case 16: return launch<16>(x, wpack, P, B, H, slices, st);
    case 32: return launch<32>(x, wpack, P, B, H, slices, st);
    case 64: return launch<64>(x, wpack, P, B, H, slices, st);
  }
  set_error("hsm_proj_tc: unsupported C=%d", C);
  return KMU_ERR_UNSUPPORTED;
}

}  // namespace tc
}  // namespace hsm
}  // namespace kmu
