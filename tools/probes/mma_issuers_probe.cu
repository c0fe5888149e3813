// Micro-probe: does a second MMA-issuing thread (another warp, its own TMEM accumulator) lift the per-instruction floor of small-N
// tcgen05.mma?  One CTA per SM; 1, 2 or 4 warps each issue `nmma` MMAs (M = 128, K = 16, bf16, constant descriptors) back to back and
// commit to their own mbarrier; reports cycles per MMA seen by the SM (wall cycles / total MMAs).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../km_unet_b200/csrc -I../../include -o mma_issuers_probe mma_issuers_probe.cu
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

using namespace kmu::tcx;

__global__ void __launch_bounds__(128) probe(int N, int nmma, int issuers, int cta2, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  __shared__ long long t_end[4];
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if ((threadIdx.x & 31) == 0 && warp < issuers) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t a0 = make_smem_desc(smem_u32(smem) + warp * 16 * 1024, 2048, 128);
    const uint64_t b0 = make_smem_desc(smem_u32(smem) + 96 * 1024 + warp * 8 * 1024, (uint32_t)N * 16, 128);
    const uint32_t d = tmem + (uint32_t)warp * 128u;
    for (int i = 0; i < nmma; i += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) umma_bf16(d, a0, b0, idesc, 1u);
    }
    umma_commit(smem_u32(&bar[warp]));
    mbar_wait_hot(smem_u32(&bar[warp]), 0);
    t_end[warp] = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t1 = 0;
    for (int i = 0; i < issuers; ++i) t1 = t_end[i] > t1 ? t_end[i] : t1;
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
  (void)cta2;
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int nmma = 2048;
  for (int N : {16, 64, 128}) {
    for (int issuers : {1, 2, 4}) {
      probe<<<148, 128, 160 * 1024>>>(N, nmma, issuers, 0, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0;
      for (int i = 0; i < 148; ++i) s += (double)h[i];
      printf("N=%3d issuers=%d : %.1f cycles per MMA (SM-level), ideal math %.1f\n", N, issuers, s / 148.0 / (nmma * issuers), N / 2.0);
    }
  }
  return 0;
}
