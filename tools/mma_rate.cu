// Developer microbenchmark: issue-to-completion rate of tcgen05.mma kind::f16 (SS operands) per instruction shape,
// one CTA per SM, optionally with four warps hammering tcgen05.ld at the same time.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I navillm_b200/csrc -o gpurun_out/mma_rate tools/mma_rate.cu
//   gpurun_out/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "nv_common.cuh"

using namespace nv;

template <uint32_t N, uint32_t B_MN, uint32_t LDERS, uint32_t NACC>
__global__ void __launch_bounds__(192, 1) rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 128 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  volatile uint32_t* stop = tptr + 1;
  const uint32_t warp = warp_id_uniform(), lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < 32 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); *stop = 0; }
  if (warp == 0) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, B_MN);
    const uint64_t ad = umma_smem_desc_sw128(smem_u32(smem), 0, 1024);                       // A: 128 x 64 K-major (16 KB) x 2 atoms
    const uint64_t bd = umma_smem_desc_sw128(smem_u32(smem + 32 * 1024), B_MN ? 8192 : 0, 1024);
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k)
          umma_f16_ss(tmem + (NACC > 1 ? (k & (NACC - 1)) * N : 0), ad + (k >> 2) * 1024 + (k & 3) * 2,
                      bd + (B_MN ? (k & 3) * 128 : ((k >> 2) * (N * 8) + (k & 3) * 2)), idesc, 1u);
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (lane == 0) {
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      *stop = 1;
    }
  } else if (warp >= 2 && warp < 2 + LDERS) {
    uint32_t v[32];
    uint32_t acc = 0;
    long long n = 0;
    while (!*stop) {
      tmem_ld_32x32b_x32(tmem + (((warp & 3) * 32) << 16) + 256, v);
      tmem_ld_wait();
      acc += v[0] + v[31];
      ++n;
    }
    if (blockIdx.x == 0 && lane == 0) out[1 + warp] = n + (acc == 0x12345 ? 1 : 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <uint32_t N, uint32_t B_MN, uint32_t LDERS, uint32_t NACC>
void run(const char* name) {
  long long* out;
  cudaMalloc(&out, 64 * 8);
  cudaMemset(out, 0, 64 * 8);
  auto k = rate_kernel<N, B_MN, LDERS, NACC>;
  const int smem = 128 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512;
  k<<<148, 192, smem>>>(iters, out);
  k<<<148, 192, smem>>>(iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[64];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const double cyc = (double)h[0] / (iters * 8.0);
  printf("%-34s N=%3u b_mn=%u loaders=%u nacc=%u : %7.1f cycles/MMA  (ideal %u, smem operand bytes %u -> %.0f B/clk)  lds/warp=%lld  %s\n",
         name, N, B_MN, LDERS, NACC, cyc, N / 2, 4096 + N * 32, (4096 + N * 32) / cyc, h[3], e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  run<32, 0, 0, 1>("S-like");
  run<64, 0, 0, 1>("S-like");
  run<64, 0, 0, 2>("S-like 2 acc");
  run<128, 0, 0, 1>("S-like");
  run<256, 0, 0, 1>("gemm-like");
  run<64, 1, 0, 1>("PV-like");
  run<128, 1, 0, 1>("PV-like");
  run<256, 1, 0, 1>("PV-like");
  run<64, 0, 4, 1>("S-like + tmem loads");
  run<128, 0, 4, 1>("S-like + tmem loads");
  run<128, 1, 4, 1>("PV-like + tmem loads");
  run<256, 0, 4, 1>("gemm-like + tmem loads");
  return 0;
}
