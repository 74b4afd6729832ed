// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, M=128, K=16) for N in {32,64,128,256},
// K-major vs MN-major operands (SWIZZLE_128B), A from shared memory.  One CTA per SM, one issuing thread,
// ITERS back-to-back MMAs on fixed shared-memory operands (values irrelevant), timed with clock64 between
// the first issue and the commit's mbarrier completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vision4leg_b200/csrc -I include -o /tmp/mma_rate tools/ubench/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"

__global__ void __launch_bounds__(128, 1) k(int N, int a_mn, int b_mn, int iters, int ndesc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::umma_idesc_f16(128, N, a_mn, b_mn);
    const uint32_t sa = tc::smem_u32(smem), sb = sa + 64 * 1024;
    long long t0 = clock64();
    if (ndesc == 0) {
      // minimal issue loop: 8 MMAs per iteration, descriptors precomputed (no address arithmetic in the loop)
      uint64_t ad[4], bd[4];
      for (int j = 0; j < 4; ++j) {
        ad[j] = a_mn ? tc::umma_smem_desc(sa + j * 2048, 16384, 1024) : tc::umma_smem_desc(sa + j * 32, 0, 1024);
        bd[j] = b_mn ? tc::umma_smem_desc(sb + j * 2048, 16384, 1024) : tc::umma_smem_desc(sb + j * 32, 0, 1024);
      }
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) tc::umma_f16(tm + (u & 1) * 256, ad[u & 3], bd[u & 3], idesc, 1u);
      }
    } else
    for (int i = 0; i < iters; ++i) {
      const int j = i % ndesc;
      // K-major: advance 32 B inside the swizzle row; MN-major: 16 reduction rows = 2048 B
      const uint64_t ad = a_mn ? tc::umma_smem_desc(sa + j * 2048, 16384, 1024) : tc::umma_smem_desc(sa + j * 32, 0, 1024);
      const uint64_t bd = b_mn ? tc::umma_smem_desc(sb + j * 2048, 16384, 1024) : tc::umma_smem_desc(sb + j * 32, 0, 1024);
      tc::umma_f16(tm + (i & 1) * 256, ad, bd, idesc, 1u);
    }
    tc::umma_commit(&bar);
    long long t1 = clock64();
    tc::mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) { tc::tc_fence_after(); tc::tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2048;
  printf("N a_mn b_mn ndesc  issue_cyc/mma  total_cyc/mma\n");
  for (int nd : {0, 4})
  for (int amn = 0; amn < 2; ++amn) for (int bmn = 0; bmn < 2; ++bmn)
    for (int N : {32, 64, 128, 256}) {
      
      k<<<148, 128, 180 * 1024>>>(N, amn, bmn, iters, nd, d);
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      printf("%3d %d %d %d   %8.1f   %8.1f   %s\n", N, amn, bmn, nd, (double)h[0] / iters, (double)h[1] / iters, cudaGetErrorString(e));
    }
  return 0;
}
