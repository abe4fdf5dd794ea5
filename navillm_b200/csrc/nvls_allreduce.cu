// navillm_b200 — in-switch all-reduce of the flat gradient buffers over NVLink SHARP (NVLS multicast), sm_100a.
//
// The one exchange step of the path (reference: DDP's bucketed NCCL all-reduce, tools/optims.py:52-54,
// fired by the last backward outside `no_sync`, tasks/agents/mp3d_agent.py:661-667).  Round 1 overlapped NCCL ring
// all-reduces of finished layer slices with the remaining backward; NCCL's channel CTAs (dozens of them, each moving data
// through SM registers, twice per element in a ring) take SMs and HBM bandwidth from the persistent wgrad GEMMs.
//
// Here every rank's gradient buffer is a SYMMETRIC allocation bound to one NVLS multicast object, and the reduction happens
// in the NVSwitch: rank r owns the r-th 1/W of a range; for each 16-byte chunk of its part it issues ONE
// `multimem.ld_reduce` (the switch reads the chunk from all W replicas and returns the sum: fp32 accumulation of bf16x2) and
// ONE `multimem.st` (the switch writes the averaged chunk back into all W replicas).  Per GPU that is 1/W of the range read
// and written once, issued by a handful of CTAs: no ring, no staging buffers, almost no SM footprint next to the GEMMs.
// Cross-rank ordering (all replicas complete before the reduce, all parts stored before anyone reads) is a barrier over the
// symmetric signal pads, issued by the host layer (navillm_b200/parallel.py) on the same side stream before and after.
#include "nv_common.cuh"
#include "nv_host.h"

namespace nv {

__device__ __forceinline__ void mm_ld_reduce_bf16x2(uint64_t mc_addr, uint32_t (&v)[4]) {
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "l"(mc_addr)
               : "memory");
}
__device__ __forceinline__ void mm_ld_reduce_f32(uint64_t mc_addr, uint32_t (&v)[4]) {
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "l"(mc_addr)
               : "memory");
}
__device__ __forceinline__ void mm_st_16(uint64_t mc_addr, const uint32_t (&v)[4]) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3])
               : "memory");
}

// mc_base: multicast address of byte 0 of the symmetric buffer; the range is [byte_off, byte_off + n_chunks * 16).
template <bool BF16>
__global__ void __launch_bounds__(512) multimem_allreduce_kernel(uint64_t mc_base, int64_t byte_off, int64_t n_chunks, int rank,
                                                                 int world, float scale) {
  const int64_t per = (n_chunks + world - 1) / world;
  const int64_t c0 = min((int64_t)rank * per, n_chunks), c1 = min(c0 + per, n_chunks);
  const uint64_t base = mc_base + (uint64_t)byte_off;
  // four independent chunks per thread and iteration: a multimem round trip through the switch takes microseconds, the
  // bandwidth comes from bytes in flight, and the CTA count has to stay small (the CTAs share SMs with the persistent
  // wgrad GEMMs: every SM they claim is taken from a CTA pair)
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = c0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < c1; c += stride * U) {
    uint32_t v[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t cc = c + u * stride;
      if (cc < c1) {
        if (BF16) mm_ld_reduce_bf16x2(base + (uint64_t)cc * 16u, v[u]);
        else mm_ld_reduce_f32(base + (uint64_t)cc * 16u, v[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t cc = c + u * stride;
      if (cc < c1) {
        if (scale != 1.f) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            v[u][i] = BF16 ? pack_bf16x2(bf16_lo(v[u][i]) * scale, bf16_hi(v[u][i]) * scale)
                           : __float_as_uint(__uint_as_float(v[u][i]) * scale);
        }
        mm_st_16(base + (uint64_t)cc * 16u, v[u]);
      }
    }
  }
}

}  // namespace nv

// In-switch all-reduce (sum, then * scale) of elements [elem_off, elem_off + n) of a symmetric buffer whose multicast
// address is mc_ptr (torch.distributed._symmetric_memory handle .multicast_ptr).  is_bf16: 1 = bf16 elements (fp32
// accumulation in the switch), 0 = fp32.  The range must start on a 16-byte boundary and hold a multiple of 16 bytes.
// Every rank of the group calls this with its own rank between two cross-rank barriers; `ctas` (<= 0: default 32) CTAs of
// 512 threads issue the multimem operations.
extern "C" int nv_multimem_allreduce(uint64_t mc_ptr, int64_t elem_off, int64_t n, int is_bf16, int rank, int world, float scale,
                                     int ctas, void* stream_) {
  using namespace nv;
  const int esz = is_bf16 ? 2 : 4;
  NV_REQUIRE(mc_ptr != 0 && n >= 0 && world >= 1 && rank >= 0 && rank < world, "nv_multimem_allreduce: bad arguments");
  NV_REQUIRE(((mc_ptr + (uint64_t)elem_off * esz) & 15) == 0 && ((n * esz) & 15) == 0,
             "nv_multimem_allreduce: range must be 16-byte aligned (off %lld, n %lld)", (long long)elem_off, (long long)n);
  if (n == 0) return NV_OK;
  const int64_t n_chunks = n * esz / 16;
  const int grid = ctas > 0 ? ctas : 32;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (is_bf16)
    multimem_allreduce_kernel<true><<<grid, 512, 0, stream>>>(mc_ptr, elem_off * esz, n_chunks, rank, world, scale);
  else
    multimem_allreduce_kernel<false><<<grid, 512, 0, stream>>>(mc_ptr, elem_off * esz, n_chunks, rank, world, scale);
  NV_LAUNCH_CHECK();
  return NV_OK;
}
