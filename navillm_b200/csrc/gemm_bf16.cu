// navillm_b200 — bf16 GEMM on tcgen05 tensor cores (the >97 %-of-FLOPs kernel of the path).
//
// Replaces the cuBLAS bf16 GEMMs the reference reaches through nn.Linear inside HF LLaMA
// (reference call sites: models/modified_lm.py:112-120 -> LlamaDecoderLayer q/k/v/o/gate/up/down
// projections and lm_head) and their autograd backward (dgrad / wgrad).
//
//   C[M,N] = A · B  (+ addend),  fp32 accumulation in TMEM, bf16 output
//
// Operand forms (all row-major bf16 in HBM, leading dimension in elements):
//   a_mn = 0 : A stored [M,K] (K contiguous)      a_mn = 1 : A stored [K,M] (M contiguous)
//   b_mn = 0 : B stored [N,K] (K contiguous)      b_mn = 1 : B stored [K,N] (N contiguous)
// so   linear fwd  Y = X W^T   -> (a_mn=0, b_mn=0),
//      dgrad       dX = dY W   -> (a_mn=0, b_mn=1),
//      wgrad       dW = dY^T X -> (a_mn=1, b_mn=1).
//
// Structure: persistent, warp-specialised, one CTA per SM.
//   warp 0 (1 lane)  TMA producer: 128-byte-swizzled tiles, STAGES-deep mbarrier ring
//   warp 1 (1 lane)  tcgen05.mma issuer, 128 x BLOCK_N x 16 atoms, accumulators in TMEM (2 stages)
//   warps 2..5       epilogue: tcgen05.ld -> bf16 round -> (+addend) -> 16-byte global stores,
//                    overlapped with the next tile's MMAs through the TMEM double buffer
// HBM layout assumptions: base pointers 16-byte aligned, leading dimensions multiples of 8 elements.
// Ragged M/N/K are handled by TMA zero-fill on loads and predicated stores.
#include "nv_common.cuh"
#include "nv_host.h"

namespace nv {

constexpr uint32_t GEMM_BLOCK_M = 128;
constexpr uint32_t GEMM_BLOCK_K = 64;
constexpr uint32_t GEMM_UMMA_K = 16;
constexpr uint32_t GEMM_THREADS = 192;
constexpr uint32_t GEMM_GROUP_M = 16;  // m-blocks per rasterisation group (L2 reuse of B tiles)

enum GemmFlags : uint32_t {
  GEMM_ADD = 1u,        // C = bf16(bf16(acc) + addend)   (residual add / gradient accumulation)
  GEMM_OUT_F32 = 2u,    // C is fp32 (no bf16 rounding of the accumulator)
};

template <uint32_t BLOCK_N, uint32_t STAGES>
struct GemmSmem {
  static constexpr uint32_t A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr uint32_t NUM_BARS = 2 * STAGES + 4;
  static constexpr uint32_t TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16;
  static constexpr uint32_t DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-byte alignment
};

__device__ __forceinline__ void tile_coords(uint32_t tile, uint32_t num_m, uint32_t num_n, uint32_t& m_blk,
                                            uint32_t& n_blk) {
  const uint32_t group_size = GEMM_GROUP_M * num_n;
  const uint32_t g = tile / group_size;
  const uint32_t first_m = g * GEMM_GROUP_M;
  const uint32_t gm = min(num_m - first_m, GEMM_GROUP_M);
  const uint32_t r = tile - g * group_size;
  m_blk = first_m + r % gm;
  n_blk = r / gm;
}

template <uint32_t BLOCK_N, uint32_t STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  void* __restrict__ Cout, int64_t ldc, const __nv_bfloat16* __restrict__ addend, int64_t ld_add,
                  uint32_t M, uint32_t N, uint32_t K, uint32_t flags) {
  using L = GemmSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * L::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = threadIdx.x & 31;

  constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                 : (2 * BLOCK_N <= 256) ? 256 : 512;
  static_assert(2 * BLOCK_N <= 512, "two accumulator stages must fit TMEM");

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (uint32_t i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (uint32_t i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const uint32_t num_m = ceil_div_u32(M, GEMM_BLOCK_M);
  const uint32_t num_n = ceil_div_u32(N, BLOCK_N);
  const uint32_t num_tiles = num_m * num_n;
  const uint32_t num_kb = ceil_div_u32(K, GEMM_BLOCK_K);

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (the whole warp runs the loop and one elected lane issues: operands stay in uniform registers, see elect_one())
    uint32_t stage = 0, phase = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      uint32_t m_blk, n_blk;
      tile_coords(tile, num_m, num_n, m_blk, n_blk);
      const int32_t m0 = m_blk * GEMM_BLOCK_M, n0 = n_blk * BLOCK_N;
      for (uint32_t kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          const int32_t k0 = kb * GEMM_BLOCK_K;
          uint8_t* sa = smem_a + stage * L::A_BYTES;
          uint8_t* sb = smem_b + stage * L::B_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);  // box {64 k, 128 m}
          } else {
#pragma unroll
            for (uint32_t i = 0; i < GEMM_BLOCK_M / 64; ++i)  // box {64 m, 64 k} per 64-wide MN atom
              tma_load_2d(sa + i * (GEMM_BLOCK_K * 128), &tmap_a, &full_bar[stage], m0 + i * 64, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);  // box {64 k, BLOCK_N n}
          } else {
#pragma unroll
            for (uint32_t i = 0; i < BLOCK_N / 64; ++i)
              tma_load_2d(sb + i * (GEMM_BLOCK_K * 128), &tmap_b, &full_bar[stage], n0 + i * 64, k0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BLOCK_M, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
    // K-major: SBO = 8 rows * 128 B; advance 32 B per UMMA_K inside the 128-byte swizzle span.
    // MN-major: LBO = one 64-wide MN atom (BLOCK_K rows * 128 B), SBO = 8 K rows; advance 16 K rows.
    constexpr uint32_t A_LBO = A_MN ? GEMM_BLOCK_K * 128 : 0, B_LBO = B_MN ? GEMM_BLOCK_K * 128 : 0;
    constexpr uint32_t A_KADV = A_MN ? (GEMM_UMMA_K * 128) >> 4 : (GEMM_UMMA_K * 2) >> 4;
    constexpr uint32_t B_KADV = B_MN ? (GEMM_UMMA_K * 128) >> 4 : (GEMM_UMMA_K * 2) >> 4;
    uint32_t stage = 0, phase = 0, iter = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
      const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (uint32_t kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem_a + stage * L::A_BYTES), A_LBO, 1024);
          const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem_b + stage * L::B_BYTES), B_LBO, 1024);
#pragma unroll
          for (uint32_t k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k)
            umma_f16_ss(tmem_d, adesc + k * A_KADV, bdesc + k * B_KADV, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (kb + 1 == num_kb) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const uint32_t quarter = warp & 3;  // TMEM lane quarter this warp may access
    const bool do_add = (flags & GEMM_ADD) != 0;
    const bool out_f32 = (flags & GEMM_OUT_F32) != 0;
    uint32_t iter = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
      uint32_t m_blk, n_blk;
      tile_coords(tile, num_m, num_n, m_blk, n_blk);
      const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t row = m_blk * GEMM_BLOCK_M + quarter * 32 + lane;
      const uint32_t col0 = n_blk * BLOCK_N;
      const uint32_t taddr = tmem_base + ((quarter * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (uint32_t c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_wait();
        const uint32_t col = col0 + c;
        if (row < M && col < N) {
          if (out_f32) {
            float* dst = reinterpret_cast<float*>(Cout) + static_cast<int64_t>(row) * ldc + col;
            if (col + 32 <= N && (ldc & 3) == 0) {
#pragma unroll
              for (uint32_t j = 0; j < 32; j += 4)
                *reinterpret_cast<uint4*>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (uint32_t j = 0; j < 32; ++j)
                if (col + j < N) dst[j] = __uint_as_float(v[j]);
            }
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(Cout) + static_cast<int64_t>(row) * ldc + col;
            const __nv_bfloat16* add = do_add ? addend + static_cast<int64_t>(row) * ld_add + col : nullptr;
            if (col + 32 <= N && (ldc & 7) == 0 && (!do_add || (ld_add & 7) == 0)) {
#pragma unroll
              for (uint32_t j = 0; j < 32; j += 8) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]));
                o.y = pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                o.z = pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                o.w = pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                if (do_add) {
                  const uint4 a = *reinterpret_cast<const uint4*>(add + j);
                  o.x = pack_bf16x2(bf16_lo(o.x) + bf16_lo(a.x), bf16_hi(o.x) + bf16_hi(a.x));
                  o.y = pack_bf16x2(bf16_lo(o.y) + bf16_lo(a.y), bf16_hi(o.y) + bf16_hi(a.y));
                  o.z = pack_bf16x2(bf16_lo(o.z) + bf16_lo(a.z), bf16_hi(o.z) + bf16_hi(a.z));
                  o.w = pack_bf16x2(bf16_lo(o.w) + bf16_lo(a.w), bf16_hi(o.w) + bf16_hi(a.w));
                }
                *reinterpret_cast<uint4*>(dst + j) = o;
              }
            } else {
#pragma unroll
              for (uint32_t j = 0; j < 32; ++j) {
                if (col + j < N) {
                  float x = bf16_round(__uint_as_float(v[j]));
                  if (do_add) x = x + __bfloat162float(add[j]);
                  dst[j] = __float2bfloat16_rn(x);
                }
              }
            }
          }
        }
      }
      // all of this warp's TMEM reads for this accumulator stage are complete (wait::ld above)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <uint32_t BLOCK_N, uint32_t STAGES, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, void* C, int64_t ldc, const void* addend,
                       int64_t ld_add, uint32_t M, uint32_t N, uint32_t K, uint32_t flags, cudaStream_t stream) {
  using L = GemmSmem<BLOCK_N, STAGES>;
  auto kern = gemm_bf16_tcgen05<BLOCK_N, STAGES, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    NV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
    attr_set = true;
  }
  const uint32_t tiles = ceil_div_u32(M, GEMM_BLOCK_M) * ceil_div_u32(N, BLOCK_N);
  const uint32_t grid = min(tiles, (uint32_t)sm_count());
  kern<<<grid, GEMM_THREADS, L::DYN_BYTES, stream>>>(ta, tb, C, ldc, reinterpret_cast<const __nv_bfloat16*>(addend),
                                                     ld_add, M, N, K, flags);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

}  // namespace nv

extern "C" int nv_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* C,
                            int64_t ldc, const void* addend, int64_t ld_add, int M, int N, int K, unsigned flags,
                            int block_n, void* stream_) {
  using namespace nv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NV_REQUIRE(M > 0 && N > 0 && K > 0, "nv_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  NV_REQUIRE(A && B && C, "nv_gemm_bf16: null operand");
  NV_REQUIRE(!(flags & GEMM_ADD) || addend, "nv_gemm_bf16: GEMM_ADD without addend");
  NV_REQUIRE(!((flags & GEMM_ADD) && (flags & GEMM_OUT_F32)), "nv_gemm_bf16: ADD with fp32 output unsupported");
  NV_REQUIRE((lda & 7) == 0 && (ldb & 7) == 0, "nv_gemm_bf16: lda/ldb must be multiples of 8 (got %lld, %lld)",
             (long long)lda, (long long)ldb);
  // auto: CTA-pair kernel (256x256 per SM pair) when there is at least one full wave of pair tiles; otherwise
  // the single-CTA kernel whose smaller tiles give skinny problems (decode, pruned rows) more parallelism
  if (block_n == 0) {
    const long pair_tiles = ((long)(M + 255) / 256) * ((long)(N + 255) / 256);
    block_n = (pair_tiles >= sm_count() / 2 && M >= 512) ? 512 : (M <= 128) ? 128 : (N >= 2048 ? 256 : 128);
    // 16 < M < 512 (one forward of a small batch, evaluation rollouts, prefix-reuse suffixes): weight streaming, and the
    // tile shape decides how many SMs pull on HBM.  Measured per variant at M = 48..384 on the four Vicuna-7B
    // projections (tools/midm_bench.py, profiles/r02_midm_gemm.txt); every variant accumulates the k-blocks in the same
    // order, so the choice does not change a single output bit.
    if (M < 512 && !a_mn && !b_mn) {
      if (M <= 128) block_n = N <= 4096 ? 32 : (N >= 16384 ? 256 : 128);
      else if (M <= 256) block_n = (N <= 4096 || N >= 16384) ? 128 : 512;
      else block_n = N <= 4096 ? 128 : 256;
    }
  }
  if (block_n == 512)   // CTA-pair kernel (cta_group::2), 256 x 256 tile per SM pair
    return gemm_bf16_2cta_dispatch(A, lda, a_mn, B, ldb, b_mn, C, ldc, addend, ld_add, M, N, K, flags, stream);
  NV_REQUIRE(block_n == 32 || block_n == 128 || block_n == 256, "nv_gemm_bf16: block_n must be 32, 128, 256 or 512 (2-CTA)");

  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = make_tmap_2d(&ta, A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, 64, GEMM_BLOCK_M);
  else       rc = make_tmap_2d(&ta, A, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, GEMM_BLOCK_K);
  if (rc) return rc;
  NV_REQUIRE(!(block_n == 32 && b_mn), "nv_gemm_bf16: block_n = 32 (skinny-M weight streaming) needs a K-major B");
  if (!b_mn) rc = make_tmap_2d(&tb, B, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, 64, (uint32_t)block_n);
  else       rc = make_tmap_2d(&tb, B, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, GEMM_BLOCK_K);
  if (rc) return rc;

#define NV_GEMM_CASE(BN, ST)                                                                                       \
  do {                                                                                                             \
    if (!a_mn && !b_mn) return launch_gemm<BN, ST, false, false>(ta, tb, C, ldc, addend, ld_add, M, N, K, flags, stream); \
    if (!a_mn && b_mn) return launch_gemm<BN, ST, false, true>(ta, tb, C, ldc, addend, ld_add, M, N, K, flags, stream);   \
    if (a_mn && b_mn) return launch_gemm<BN, ST, true, true>(ta, tb, C, ldc, addend, ld_add, M, N, K, flags, stream);     \
    return launch_gemm<BN, ST, true, false>(ta, tb, C, ldc, addend, ld_add, M, N, K, flags, stream);                      \
  } while (0)

  if (block_n == 256) NV_GEMM_CASE(256, 4);
  if (block_n == 32) {
    // decode / pruned-row GEMMs (M <= 128): HBM-bound weight streaming.  32-column tiles give every projection of
    // the model >= 128 CTAs, and 10 stages keep ~40 KB of weights in flight per SM.
    if (!a_mn) return launch_gemm<32, 10, false, false>(ta, tb, C, ldc, addend, ld_add, M, N, K, flags, stream);
    return launch_gemm<32, 10, true, false>(ta, tb, C, ldc, addend, ld_add, M, N, K, flags, stream);
  }
  NV_GEMM_CASE(128, 6);
#undef NV_GEMM_CASE
}
