// Micro-probe: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N and of the alignment of the 8-row x 16-byte core
// matrices of the no-swizzle K-major A operand.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../km_unet_b200/csrc
// -I../../include -o mma_align_probe mma_align_probe.cu ;  run on one B200.
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

using namespace kmu::tcx;

__global__ void __launch_bounds__(128) probe(int N, int nmma, uint32_t a_off, uint32_t sbo, uint32_t a_step, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t a0 = make_smem_desc(smem_u32(smem) + a_off, 2880 * 8, sbo);          // LBO: K halves far apart
    const uint64_t b0 = make_smem_desc(smem_u32(smem) + 96 * 1024, (uint32_t)N * 16, 128);
    long long t0 = clock64();
    if (a_step == 0xFFFFFFFFu) {   // issue-overhead control: constant descriptors, 16 MMAs per loop iteration
      for (int i = 0; i < nmma; i += 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j) umma_bf16(tmem, a0, b0, idesc, 1u);
      }
    } else {
      for (int i = 0; i < nmma; ++i) umma_bf16(tmem, desc_advance(a0, (uint32_t)(i & 7) * a_step), b0, idesc, i > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait_hot(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int nmma = 2048;
  struct Cfg { const char* name; uint32_t off, sbo, step; } cfgs[] = {
      {"aligned   (off 0,  SBO 128, step 2048)", 0, 128, 2048},
      {"pitch 160 (off 16, SBO 160, step 160)", 16, 160, 160},
      {"const desc, unrolled x16 (off 0, SBO 128)", 0, 128, 0xFFFFFFFFu}, {"const desc, unrolled x16 (off 16, SBO 160)", 16, 160, 0xFFFFFFFFu}};
  for (int N : {8, 16, 64, 128, 192, 256}) {
    for (auto& c : cfgs) {
      for (int grid : {148}) {
        probe<<<grid, 128, 160 * 1024>>>(N, nmma, c.off, c.sbo, c.step, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("N=%3d %-42s grid %3d : %.1f cycles / MMA\n", N, c.name, grid, (double)mx / nmma);
      }
    }
  }
  return 0;
}
