// optim.cu -- the AdamW update of train_shanghai.py:342 (`optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05)`, stepped at
// :180) over ALL parameter tensors of a group in one launch.
//
//   p <- p (1 - lr wd);   m <- m + (1 - b1)(g - m);   v <- b2 v + (1 - b2) g g;
//   p <- p - (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)                    t = step count after this update
//
// torch's fused AdamW walks KM-UNet's 664 live tensors (1.29 M elements, most of them 16 .. 256 elements long) in 19
// multi_tensor_apply launches at the very end of the step, where nothing else is left to overlap with.  Here the caller keeps a
// device table of (p, g, m, v, n) entries and a chunk list (entry, first element); a CTA owns one chunk of 1024 elements, so the
// whole update is one launch of ~2 k CTAs over 36 MB of L2-resident traffic, plus a one-thread kernel that advances the device
// step counter (the counter lives on the device so the update can sit inside a CUDA graph).
#include "common.cuh"

namespace kmu {
namespace optim {

constexpr int CHUNK = 1024;     // elements per CTA: 256 threads x 4

struct Hyper {
  double beta1, beta2;                                         // for the bias corrections (fp64, once per CTA)
  float lr, weight_decay, omb1, b2, omb2, eps, grad_scale;     // omb = 1 - beta, rounded once from the fp64 difference
};

__global__ void __launch_bounds__(256) adamw_kernel(const kmu_adamw_entry* __restrict__ entries, const long long* __restrict__ chunks,
                                                    const float* __restrict__ step, const float* __restrict__ lr_dev, Hyper h) {
  __shared__ float s_c[2];
  if (threadIdx.x == 0) {
    const double t = (double)step[0] + 1.0;
    s_c[0] = (float)(1.0 - pow(h.beta1, t));
    s_c[1] = (float)sqrt(1.0 - pow(h.beta2, t));
  }
  const long long code = chunks[blockIdx.x];
  const kmu_adamw_entry e = entries[(int)(code >> 32)];
  const long long first = (long long)(unsigned)(code & 0xffffffffll) * CHUNK;
  __syncthreads();
  const float lr = lr_dev ? lr_dev[0] : h.lr;
  const float step_size = lr / s_c[0], inv_bc2 = 1.0f / s_c[1], decay = 1.0f - lr * h.weight_decay;
  const long long i0 = first + 4 * threadIdx.x;
  if (i0 >= e.n) return;
  float p[4], g[4], m[4], v[4];
  const bool vec = (i0 + 4 <= e.n) && ((((size_t)e.p | (size_t)e.g | (size_t)e.m | (size_t)e.v) & 15) == 0);
  const int cnt = vec ? 4 : (int)min(4ll, e.n - i0);
  if (vec) {
    *(float4*)p = *(const float4*)(e.p + i0);
    *(float4*)g = *(const float4*)(e.g + i0);
    *(float4*)m = *(const float4*)(e.m + i0);
    *(float4*)v = *(const float4*)(e.v + i0);
  } else {
    for (int j = 0; j < cnt; ++j) { p[j] = e.p[i0 + j]; g[j] = e.g[i0 + j]; m[j] = e.m[i0 + j]; v[j] = e.v[i0 + j]; }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < cnt) {
      const float gj = g[j] * h.grad_scale;
      const float pj = p[j] * decay;
      const float mj = m[j] + h.omb1 * (gj - m[j]);
      const float vj = h.b2 * v[j] + h.omb2 * gj * gj;
      p[j] = pj - step_size * (mj / (sqrtf(vj) * inv_bc2 + h.eps));
      m[j] = mj;
      v[j] = vj;
    }
  }
  if (vec) {
    *(float4*)(e.p + i0) = *(float4*)p;
    *(float4*)(e.m + i0) = *(float4*)m;
    *(float4*)(e.v + i0) = *(float4*)v;
  } else {
    for (int j = 0; j < cnt; ++j) { e.p[i0 + j] = p[j]; e.m[i0 + j] = m[j]; e.v[i0 + j] = v[j]; }
  }
}

__global__ void adamw_tick_kernel(float* step) { step[0] += 1.0f; }

}  // namespace optim
}  // namespace kmu

extern "C" {

int32_t kmu_adamw_chunk_elems(void) { return kmu::optim::CHUNK; }

int kmu_adamw_step(const kmu_adamw_args* a, kmu_stream stream) {
  KMU_REQUIRE(a && a->entries && a->chunks && a->step, KMU_ERR_BAD_ARG, "adamw_step: null argument");
  KMU_REQUIRE(a->n_chunks > 0 && a->n_entries > 0, KMU_ERR_BAD_ARG, "adamw_step: empty table");
  KMU_REQUIRE(a->beta1 >= 0. && a->beta1 < 1. && a->beta2 >= 0. && a->beta2 < 1. && a->eps >= 0., KMU_ERR_BAD_ARG,
              "adamw_step: betas must lie in [0, 1), eps >= 0");
  kmu::optim::Hyper h{a->beta1, a->beta2, (float)a->lr, (float)a->weight_decay, (float)(1.0 - a->beta1), (float)a->beta2,
                      (float)(1.0 - a->beta2), (float)a->eps, (float)a->grad_scale};
  cudaStream_t st = (cudaStream_t)stream;
  kmu::optim::adamw_kernel<<<(unsigned)a->n_chunks, 256, 0, st>>>(a->entries, (const long long*)a->chunks, a->step, a->lr_dev, h);
  KMU_LAUNCH_CHECK("adamw_step");
  kmu::optim::adamw_tick_kernel<<<1, 1, 0, st>>>(a->step);
  KMU_LAUNCH_CHECK("adamw_tick");
  return KMU_OK;
}

}  // extern "C"
