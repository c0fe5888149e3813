// kan_common.cuh -- shapes, knot-table layout and the Phi(x) evaluators shared by the KAN kernel families.
#pragma once
#include "common.cuh"

namespace kmu {
namespace kan {

constexpr int ORDER = 3;   // cubic
constexpr int NB = 8;      // basis functions per feature (grid_size 5 + order 3)
constexpr int NK = 12;     // knots per feature
constexpr int NPHI = 9;    // [SiLU, B_0..B_7]
// per-feature knot table: knots, then reciprocal spans 1/(t[j+p]-t[j]) for p = 1,2,3
constexpr int KT = 44, KT_R1 = 12, KT_R2 = 23, KT_R3 = 33;

struct Dims {
  int B, Cin, H, W, Cout, k, stride, pad, Ho, Wo, F;
  long long M;  // B*Ho*Wo
};

inline Dims make_dims(const kmu_kanconv2d_desc& s) {
  Dims d;
  d.B = s.B; d.Cin = s.Cin; d.H = s.H; d.W = s.W; d.Cout = s.Cout;
  d.k = s.ksize; d.stride = s.stride; d.pad = s.padding;
  d.Ho = (s.H + 2 * s.padding - s.ksize) / s.stride + 1;
  d.Wo = (s.W + 2 * s.padding - s.ksize) / s.stride + 1;
  d.F = s.Cin * s.ksize * s.ksize;
  d.M = (long long)s.B * d.Ho * d.Wo;
  return d;
}

// Cox-de Boor on an arbitrary knot row (convKAN/KANlayers.py:593-603): half-open order-0 indicators, three levels.
// `lvl2` receives the 9 quadratic values the derivative needs.
__device__ __forceinline__ void cox_de_boor(float x, const float* __restrict__ k, float* __restrict__ b3, float* __restrict__ lvl2) {
  float b[11];
#pragma unroll
  for (int j = 0; j < 11; ++j) b[j] = (x >= k[j] && x < k[j + 1]) ? 1.0f : 0.0f;
#pragma unroll
  for (int j = 0; j < 10; ++j) b[j] = (x - k[j]) * k[KT_R1 + j] * b[j] + (k[j + 2] - x) * k[KT_R1 + j + 1] * b[j + 1];
#pragma unroll
  for (int j = 0; j < 9; ++j) b[j] = (x - k[j]) * k[KT_R2 + j] * b[j] + (k[j + 3] - x) * k[KT_R2 + j + 1] * b[j + 1];
  if (lvl2) {
#pragma unroll
    for (int j = 0; j < 9; ++j) lvl2[j] = b[j];
  }
  if (b3) {
#pragma unroll
    for (int j = 0; j < 8; ++j) b3[j] = (x - k[j]) * k[KT_R3 + j] * b[j] + (k[j + 4] - x) * k[KT_R3 + j + 1] * b[j + 1];
  }
}

__device__ __forceinline__ void eval_phi(float x, const float* __restrict__ k, float* __restrict__ phi) {
  phi[0] = siluf_(x);
  cox_de_boor(x, k, phi + 1, nullptr);
}

// d/dx of every Phi component: SiLU'(x) and B'_{j,3} = 3 (B_{j,2}/(t_{j+3}-t_j) - B_{j+1,2}/(t_{j+4}-t_{j+1})).
__device__ __forceinline__ void eval_dphi(float x, const float* __restrict__ k, float* __restrict__ dphi) {
  float q[9];
  cox_de_boor(x, k, nullptr, q);
  dphi[0] = silu_gradf_(x);
#pragma unroll
  for (int j = 0; j < 8; ++j) dphi[1 + j] = 3.0f * (q[j] * k[KT_R3 + j] - q[j + 1] * k[KT_R3 + j + 1]);
}

// fp32 CUDA-core family (kan_simt.cu)
size_t simt_fwd_workspace(const Dims& d);
size_t simt_bwd_workspace(const Dims& d);
int simt_forward(const kmu_kanconv2d_fwd_args* a, const Dims& d, cudaStream_t st);
int simt_backward(const kmu_kanconv2d_bwd_args* a, const Dims& d, cudaStream_t st);

// tcgen05 / TMEM family (kan_tc.cu)
namespace tc {
bool supported(const kmu_kanconv2d_desc& s);
size_t fwd_workspace(const Dims& d);
int forward(const kmu_kanconv2d_fwd_args* a, const Dims& d, cudaStream_t st);
size_t bwd_workspace(const Dims& d);
int backward(const kmu_kanconv2d_bwd_args* a, const Dims& d, cudaStream_t st);
void set_debug_flags(int flags);
}  // namespace tc

}  // namespace kan
}  // namespace kmu
