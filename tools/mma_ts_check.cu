// Developer check for the NEXT step of the attention kernels: tcgen05.mma with the A operand in TENSOR
// MEMORY (".ts" form), the FA4 arrangement that removes the 128 B/clk shared-memory operand bound measured by
// tools/mma_rate.cu.  Two questions, one run on the GPU box:
//   1. layout   is a bf16 A tile [128 x 64] written with tcgen05.st.32x32b (thread t of warp w = TMEM lane 32w+t = row,
//               32-bit column c = elements k = 2c, 2c+1) what the TS-form MMA expects?  D_ts is compared with D_ss
//               (same A staged in 128B-swizzled smem) element by element on small-integer data (exact in fp32).
//   2. rate     cycles per TS MMA for N = 64 / 128 (expected: the N/2 floor, no A read from smem).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I navillm_b200/csrc -o /tmp/mma_ts_check tools/mma_ts_check.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "nv_common.cuh"

using namespace nv;


// deterministic small integers: exact products / sums in bf16 x bf16 -> fp32
__device__ __forceinline__ float aval(uint32_t m, uint32_t k) { return (float)((int)((m * 7 + k * 13) % 9) - 4); }
__device__ __forceinline__ float bval(uint32_t n, uint32_t k) { return (float)((int)((n * 5 + k * 3) % 7) - 3); }

template <uint32_t N>
__global__ void __launch_bounds__(192, 1) ts_kernel(float* out_ss, float* out_ts, long long* cycles, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                      // [128 x 64] bf16, one 16 KB swizzle atom
  uint8_t* sB = smem + 16384;              // [N x 64] bf16 K-major
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 256 * 128);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const uint32_t warp = warp_id_uniform(), lane = threadIdx.x & 31;
  // fill A and B in the TMA SWIZZLE_128B layout (16-byte chunk c of row r at sw128_offset(r, c))
  for (uint32_t i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const uint32_t r = i >> 3, c = i & 7;
    uint32_t w[4];
    for (uint32_t q = 0; q < 4; ++q) w[q] = pack_bf16x2(aval(r, c * 8 + 2 * q), aval(r, c * 8 + 2 * q + 1));
    *reinterpret_cast<uint4*>(sA + sw128_offset(r, c)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (uint32_t i = threadIdx.x; i < N * 8; i += blockDim.x) {
    const uint32_t r = i >> 3, c = i & 7;
    uint32_t w[4];
    for (uint32_t q = 0; q < 4; ++q) w[q] = pack_bf16x2(bval(r, c * 8 + 2 * q), bval(r, c * 8 + 2 * q + 1));
    *reinterpret_cast<uint4*>(sB + sw128_offset(r, c)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  static_assert(N <= 128, "column budget: D_ss [0,128) D_ts [128,256) A [256,288)");
  const uint32_t tD_ss = tmem, tD_ts = tmem + 128, tA = tmem + 256;        // A: 32 columns (64 bf16 per row)
  // warps 2..5: write A into TMEM: lane = row, column c = (k = 2c, 2c + 1)
  if (warp >= 2) {
    const uint32_t r = (warp & 3) * 32 + lane;
    uint32_t v[32];
#pragma unroll
    for (uint32_t c = 0; c < 32; ++c) v[c] = pack_bf16x2(aval(r, 2 * c), aval(r, 2 * c + 1));
    tmem_st_32x32b_x32(tA + (((warp & 3) * 32) << 16), v);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
  const uint64_t ad = umma_smem_desc_sw128(smem_u32(sA), 0, 1024), bd = umma_smem_desc_sw128(smem_u32(sB), 0, 1024);
  if (warp == 1) {
    if (elect_one()) {
#pragma unroll
      for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_ss(tD_ss, ad + ks * 2, bd + ks * 2, idesc, ks ? 1u : 0u);
#pragma unroll
      for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_ts(tD_ts, tA + ks * 8, bd + ks * 2, idesc, ks ? 1u : 0u);
      umma_commit(&bar[0]);
    }
    __syncwarp();
    mbar_wait(&bar[0], 0);
    // rate: TS MMAs back to back
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it)
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_ts(tmem + 384, tA + ks * 8, bd + ks * 2, idesc, 1u);   // scratch D: the
      // compute warps are reading tD_ts at this moment (round 1 accumulated the rate loop INTO tD_ts and reported the layout
      // as wrong: that was this race, not the layout - the probes below and the backward kernels confirm it)
      umma_commit(&bar[1]);
    }
    __syncwarp();
    mbar_wait(&bar[1], 0);
    if (lane == 0 && blockIdx.x == 0) cycles[0] = clock64() - t0;
  }
  if (warp >= 2) {
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t r = (warp & 3) * 32 + lane;
    for (uint32_t c = 0; c < N; c += 32) {
      uint32_t a[32], b[32];
      tmem_ld_32x32b_x32(tD_ss + (((warp & 3) * 32) << 16) + c, a);
      tmem_ld_32x32b_x32(tD_ts + (((warp & 3) * 32) << 16) + c, b);
      tmem_ld_wait();
      if (blockIdx.x == 0)
        for (uint32_t j = 0; j < 32; ++j) {
          out_ss[r * N + c + j] = __uint_as_float(a[j]);
          out_ts[r * N + c + j] = __uint_as_float(b[j]);
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// Probe of the TMEM A-operand layout (run when the hypothesis above fails): TMEM cell (row r, 32-bit column c) holds the
// bf16 pair (1 + 2c, 2 + 2c) [probe 0] or (r, r) for r < 128 [probe 1]; B = identity (B[n][k] = delta(n, k), N = K = 64).
// Then D[m][n] = the value the tensor core read as A[m][k = n]: probe 0 names the (column, half) feeding each k, probe 1
// names the TMEM lane feeding each output row.
template <int PROBE>
__global__ void __launch_bounds__(192, 1) probe_kernel(float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 128);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t warp = warp_id_uniform(), lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const uint32_t r = i >> 3, c = i & 7;
    uint32_t w[4];
    for (uint32_t q = 0; q < 4; ++q)
      w[q] = pack_bf16x2(r == c * 8 + 2 * q ? 1.f : 0.f, r == c * 8 + 2 * q + 1 ? 1.f : 0.f);
    *reinterpret_cast<uint4*>(sB + sw128_offset(r, c)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr, tD = tmem, tA = tmem + 256;
  if (warp >= 2) {
    const uint32_t r = (warp & 3) * 32 + lane;
    uint32_t v[32];
#pragma unroll
    for (uint32_t c = 0; c < 32; ++c)
      v[c] = PROBE == 0 ? pack_bf16x2((float)(1 + 2 * c), (float)(2 + 2 * c)) : pack_bf16x2((float)r, (float)r);
    tmem_st_32x32b_x32(tA + (((warp & 3) * 32) << 16), v);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
  const uint64_t bd = umma_smem_desc_sw128(smem_u32(sB), 0, 1024);
  if (warp == 1) {
    if (elect_one()) {
#pragma unroll
      for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_ts(tD, tA + ks * 8, bd + ks * 2, idesc, ks ? 1u : 0u);
      umma_commit(bar);
    }
    __syncwarp();
  }
  if (warp >= 2) {
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t r = (warp & 3) * 32 + lane;
    for (uint32_t c = 0; c < 64; c += 32) {
      uint32_t a[32];
      tmem_ld_32x32b_x32(tD + (((warp & 3) * 32) << 16) + c, a);
      tmem_ld_wait();
      for (uint32_t j = 0; j < 32; ++j) out[r * 64 + c + j] = __uint_as_float(a[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int PROBE>
void probe() {
  float* out;
  cudaMalloc(&out, 128 * 64 * 4);
  auto k = probe_kernel<PROBE>;
  const int smem = 64 * 128 + 64 + 1024;
  k<<<1, 192, smem>>>(out);
  cudaError_t e = cudaDeviceSynchronize();
  static float h[128 * 64];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("probe %d (%s)  %s\n", PROBE, PROBE == 0 ? "value 1+2c+half read as k, rows 0, 1, 37, 64, 100" : "TMEM lane read as output row m (column 0), all rows",
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  if (PROBE == 0) {
    const int rows[5] = {0, 1, 37, 64, 100};
    for (int ri = 0; ri < 5; ++ri) {
      printf("  row %3d:", rows[ri]);
      for (int n = 0; n < 64; ++n) printf(" %g", h[rows[ri] * 64 + n]);
      printf("\n");
    }
  } else {
    printf("  ");
    for (int m = 0; m < 128; ++m) printf(" %g", h[m * 64]);
    printf("\n");
  }
  cudaFree(out);
}

template <uint32_t N>
void run() {
  float *ss, *ts;
  long long* cyc;
  cudaMalloc(&ss, 128 * N * 4); cudaMalloc(&ts, 128 * N * 4); cudaMalloc(&cyc, 8);
  cudaMemset(ss, 0, 128 * N * 4); cudaMemset(ts, 0xff, 128 * N * 4);
  auto k = ts_kernel<N>;
  const int smem = 16384 + 256 * 128 + 64 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 1024;
  k<<<148, 192, smem>>>(ss, ts, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  static float hs[128 * 256], ht[128 * 256];
  long long hc = 0;
  cudaMemcpy(hs, ss, 128 * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(ht, ts, 128 * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  double maxd = 0, maxref = 0;
  int bad = 0;
  for (uint32_t m = 0; m < 128; ++m)
    for (uint32_t n = 0; n < N; ++n) {
      double ref = 0;
      for (uint32_t kk = 0; kk < 64; ++kk) ref += (double)((int)((m * 7 + kk * 13) % 9) - 4) * (double)((int)((n * 5 + kk * 3) % 7) - 3);
      const double d_ss = fabs(hs[m * N + n] - ref), d_ts = fabs(ht[m * N + n] - ref);
      if (d_ss > maxref) maxref = d_ss;
      if (d_ts > maxd) maxd = d_ts;
      if (d_ts != 0) ++bad;
    }
  printf("N=%3u: SS max|err| %.1f, TS max|err| %.1f (%d of %u elements differ) -> TMEM A layout hypothesis %s;  TS rate %.1f cycles/MMA (floor %u)  %s\n",
         N, maxref, maxd, bad, 128 * N, bad == 0 ? "CONFIRMED" : "WRONG", (double)hc / (iters * 4.0), N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(ss); cudaFree(ts); cudaFree(cyc);
}

int main() {
  run<64>();
  run<128>();
  probe<0>();
  probe<1>();
  return 0;
}
