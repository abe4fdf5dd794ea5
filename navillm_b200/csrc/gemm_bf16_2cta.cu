// navillm_b200 — bf16 GEMM on tcgen05 with CTA pairs (cta_group::2): 256 x 256 output tile per SM pair.
//
// Same contract as gemm_bf16.cu (C[M,N] = A·B (+addend), K-/MN-major operands), different mapping:
// two CTAs of a cluster (adjacent SMs) cooperate on one UMMA_M = 256 instruction stream issued by the
// leader CTA.  Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N rows); the
// tensor cores of the pair read both halves of B across the pair, so per-SM shared-memory traffic and
// capacity for B halve (32 KB per stage instead of 48 KB -> 6 stages, and less energy per FLOP on a
// power-capped part).  Accumulators: each CTA's TMEM holds its own 128 rows x 256 columns, 2 stages.
//
//   warp 0 (1 lane, both CTAs)  TMA producer; completes transactions on the LEADER's full barrier
//   warp 1 (1 lane, leader)     tcgen05.mma.cta_group::2 issuer; commits are multicast to both CTAs
//   warps 2..5 (both CTAs)      epilogue for the CTA's own 128 rows; arrive on the leader's tmem_empty
//
// Tile scheduling is DYNAMIC: the leader's producer warp claims pair tiles from a global atomic counter and
// publishes each claim to every consumer warp of both CTAs through a 4-slot ring in shared memory (TileRing).
// With a static "tile = cluster + i * clusters" schedule a CTA pair that becomes resident late - because NCCL's
// all-reduce CTAs (overlapped with the backward) or another kernel still hold its SMs - runs its whole tile list
// after everybody else has finished, which doubles the GEMM's duration; with claims the late pair simply finds
// the counter exhausted.  The counter resets itself (the claim that returns the last value stores 0), so launches
// on one stream, including CUDA-graph replays, need no memset.  tile_counter == nullptr selects the static order.
#include "nv_common.cuh"
#include "nv_host.h"
#include <cstdlib>
#include <mutex>

namespace nv {

constexpr uint32_t G2_BM = 128;       // rows per CTA (256 per pair)
constexpr uint32_t G2_BN = 256;       // columns per pair
constexpr uint32_t G2_BNH = 128;      // B rows staged per CTA
constexpr uint32_t G2_BK = 64;
constexpr uint32_t G2_STAGES = 6;
constexpr uint32_t G2_THREADS = 192;
constexpr uint32_t G2_GROUP_M = 8;    // pair-tiles (256 rows) per rasterisation group
constexpr uint32_t G2_A_BYTES = G2_BM * G2_BK * 2;
constexpr uint32_t G2_B_BYTES = G2_BNH * G2_BK * 2;
constexpr uint32_t G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr uint32_t G2_BAR_OFF = G2_STAGES * G2_STAGE_BYTES;
constexpr uint32_t G2_SCHED = 4;      // tile-claim ring slots
constexpr uint32_t G2_NUM_BARS = 2 * G2_STAGES + 4 + 2 * G2_SCHED;
constexpr uint32_t G2_DYN_BYTES = G2_BAR_OFF + G2_NUM_BARS * 8 + 16 + G2_SCHED * 4 + 1024;
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> CTA 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose transaction bytes complete on the LEADER CTA's mbarrier (same smem offset, rank bit cleared).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                                int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit: arrive (once the issued MMAs are done) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- cluster-scope release/acquire on mbarriers (the tile ring carries DATA between the two CTAs) -------------
__device__ __forceinline__ void mbar_arrive_release_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t* p, uint32_t cta, uint32_t v) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.u32 [ra], %2;\n\t}"
      ::"r"(smem_u32(p)), "r"(cta), "r"(v)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}

// Ring of claimed tile indices, replicated in both CTAs of the pair.  One publisher (leader CTA, warp 0) and ten
// consumer warps (leader: MMA + 4 epilogue; peer: producer + 4 epilogue); `empty` lives in the leader.
struct TileRing {
  uint64_t* full;    // [G2_SCHED] per CTA, count 1
  uint64_t* empty;   // [G2_SCHED] in the leader, count 10
  uint32_t* tiles;   // [G2_SCHED] per CTA
  uint32_t idx;
  // whole warp; returns the tile index (>= num_tiles: no more work)
  __device__ __forceinline__ uint32_t claim(uint32_t* counter, uint32_t total_claims) {
    const uint32_t slot = idx % G2_SCHED, ph = (idx / G2_SCHED) & 1;
    ++idx;
    mbar_wait(&empty[slot], ph ^ 1);
    if (elect_one()) {
      const uint32_t t = atomicAdd(counter, 1u);
      if (t == total_claims - 1) atomicExch(counter, 0u);   // last claim of this launch: ready for the next one
      tiles[slot] = t;
      st_shared_cluster_u32(&tiles[slot], 1, t);
      mbar_arrive_release_cluster(&full[slot], 0);
      mbar_arrive_release_cluster(&full[slot], 1);
    }
    __syncwarp();
    return *reinterpret_cast<volatile uint32_t*>(&tiles[slot]);
  }
  __device__ __forceinline__ uint32_t take(bool leader, uint32_t lane) {
    const uint32_t slot = idx % G2_SCHED, ph = (idx / G2_SCHED) & 1;
    ++idx;
    mbar_wait_acquire_cluster(&full[slot], ph);
    const uint32_t t = *reinterpret_cast<volatile uint32_t*>(&tiles[slot]);
    __syncwarp();
    if (lane == 0) {
      if (leader) mbar_arrive(&empty[slot]);
      else mbar_arrive_remote(&empty[slot], 0);
    }
    return t;
  }
};

__device__ __forceinline__ void tile_coords2(uint32_t tile, uint32_t num_m, uint32_t num_n, uint32_t group_m,
                                             uint32_t& m_blk, uint32_t& n_blk) {
  const uint32_t group_size = group_m * num_n;
  const uint32_t g = tile / group_size;
  const uint32_t first_m = g * group_m;
  const uint32_t gm = min(num_m - first_m, group_m);
  const uint32_t r = tile - g * group_size;
  m_blk = first_m + r % gm;
  n_blk = r / gm;
}

// Fused epilogues (EPI):
//   EPI_PLAIN   C = bf16(acc) (+addend) | fp32(acc)
//   EPI_SWIGLU  gate|up projection: the pair tile holds 128 gate columns (CTA 0's half of B = gate rows) and the
//               SAME 128 up columns (CTA 1's half = up rows F + n); writes g,u into the fused [T,2F] buffer (kept
//               for the backward) and h = bf16(bf16(silu(g)) * u) into aux [T,F]   (HF LlamaMLP act_fn(gate)*up)
//   EPI_DSWIGLU down-projection dgrad: acc = dh; reads g,u from aux [T,2F] and writes dg = dh*u*silu'(g),
//               du = dh*silu(g) into C [T,2F] (the separate swiglu_bwd pass and the dh round trip disappear)
//   EPI_ATTND   o_proj dgrad: C = dO (gradient of the attention output) and, per row and 128-column head, the
//               D[h, t] = sum_d dO[t,h,d] * O[t,h,d] vector of the attention backward (aux = O; replaces attn_bwd_prep)
//   EPI_ROPE    fused q|k|v projection: columns < rope_cols get the rotate-half rotary embedding
//               (HF apply_rotary_pos_emb, bf16 rounding points preserved) with cos/sin[pos[row]] before the store
enum { EPI_PLAIN = 0, EPI_SWIGLU = 1, EPI_DSWIGLU = 2, EPI_ROPE = 3, EPI_ATTND = 4 };

struct EpiAux {
  void* aux;            // SWIGLU: h out [T,F];  DSWIGLU: gu in [T,2F]
  int64_t ld_aux;
  uint32_t F;           // SWIGLU / DSWIGLU: intermediate size (columns of gate and of up)
  float* dvec;          // ATTND: D out [H, T] fp32 (T = M rows)
  const int* pos;       // ROPE
  const __nv_bfloat16* cos_t;
  const __nv_bfloat16* sin_t;
  uint32_t rope_cols;   // ROPE: q and k columns (2 * H * 128)
};

__device__ __forceinline__ float silu_bf16(float g) { return bf16_round(g / (1.f + __expf(-g))); }

template <bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_bf16_tcgen05_2cta(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       void* __restrict__ Cout, int64_t ldc, const __nv_bfloat16* __restrict__ addend, int64_t ld_add,
                       uint32_t M, uint32_t N, uint32_t K, uint32_t flags, EpiAux ea, uint32_t* __restrict__ tile_counter,
                       uint32_t group_m) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + G2_STAGES * G2_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G2_BAR_OFF);
  uint64_t* full_bar = bars;                       // used in the leader
  uint64_t* empty_bar = bars + G2_STAGES;          // per CTA
  uint64_t* tmem_full = bars + 2 * G2_STAGES;      // per CTA [2]
  uint64_t* tmem_empty = bars + 2 * G2_STAGES + 2; // used in the leader [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + G2_NUM_BARS);
  TileRing ring;
  ring.full = bars + 2 * G2_STAGES + 4;
  ring.empty = ring.full + G2_SCHED;
  ring.tiles = tmem_ptr_smem + 4;
  ring.idx = 0;
  const bool dyn = tile_counter != nullptr;

  const uint32_t warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (uint32_t i = 0; i < G2_STAGES; ++i) {
      mbar_init(&full_bar[i], 2);   // leader's arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty_bar[i], 1);  // multicast commit from the leader
    }
    for (uint32_t i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);   // multicast commit
      mbar_init(&tmem_empty[i], 8);  // 4 epilogue warps of each CTA
    }
    for (uint32_t i = 0; i < G2_SCHED; ++i) {
      mbar_init(&ring.full[i], 1);
      mbar_init(&ring.empty[i], 10);
    }
    fence_mbar_init();
  }
  cluster_sync_all();                 // barrier inits of both CTAs visible before any remote arrive / TMA
  if (warp == 1) {
    tmem_alloc_2sm(tmem_ptr_smem, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // logical output columns per pair tile: 256, except SWIGLU where the 256 accumulator columns are 128 gate + 128 up
  constexpr uint32_t TILE_N = (EPI == EPI_SWIGLU) ? 128u : G2_BN;
  const uint32_t num_m = ceil_div_u32(M, 2 * G2_BM);
  const uint32_t num_n = ceil_div_u32(N, TILE_N);
  const uint32_t num_tiles = num_m * num_n;
  const uint32_t num_kb = ceil_div_u32(K, G2_BK);
  const uint32_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    // (the whole warp runs the loop and one elected lane issues: operands stay in uniform registers, see elect_one())
    uint32_t stage = 0, phase = 0;
    for (uint32_t it = 0;; ++it) {
      const uint32_t tile = !dyn ? cluster_id + it * num_clusters
                                 : (leader ? ring.claim(tile_counter, num_tiles + num_clusters) : ring.take(false, lane));
      if (tile >= num_tiles) break;
      uint32_t m_blk, n_blk;
      tile_coords2(tile, num_m, num_n, group_m, m_blk, n_blk);
      const int32_t m0 = m_blk * 2 * G2_BM + rank * G2_BM;    // this CTA's A rows
      // this CTA's half of the B tile (SWIGLU: CTA 0 stages gate rows n.., CTA 1 the matching up rows F + n..)
      const int32_t n0 = (EPI == EPI_SWIGLU) ? (int32_t)(n_blk * 128 + rank * ea.F) : (int32_t)(n_blk * G2_BN + rank * G2_BNH);
      for (uint32_t kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
          else mbar_arrive_remote(&full_bar[stage], 0);
          const int32_t k0 = kb * G2_BK;
          uint8_t* sa = smem_a + stage * G2_A_BYTES;
          uint8_t* sb = smem_b + stage * G2_B_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (uint32_t i = 0; i < G2_BM / 64; ++i)
              tma_load_2d_2sm(sa + i * (G2_BK * 128), &tmap_a, &full_bar[stage], m0 + i * 64, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d_2sm(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (uint32_t i = 0; i < G2_BNH / 64; ++i)
              tma_load_2d_2sm(sb + i * (G2_BK * 128), &tmap_b, &full_bar[stage], n0 + i * 64, k0);
          }
        }
        __syncwarp();
        if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * G2_BM, G2_BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      constexpr uint32_t A_LBO = A_MN ? G2_BK * 128 : 0, B_LBO = B_MN ? G2_BK * 128 : 0;
      constexpr uint32_t A_KADV = A_MN ? (16 * 128) >> 4 : (16 * 2) >> 4;
      constexpr uint32_t B_KADV = B_MN ? (16 * 128) >> 4 : (16 * 2) >> 4;
      uint32_t stage = 0, phase = 0;
      for (uint32_t iter = 0;; ++iter) {
        const uint32_t tile = dyn ? ring.take(true, lane) : cluster_id + iter * num_clusters;
        if (tile >= num_tiles) break;
        const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * G2_BN;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem_a + stage * G2_A_BYTES), A_LBO, 1024);
            const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem_b + stage * G2_B_BYTES), B_LBO, 1024);
#pragma unroll
            for (uint32_t k = 0; k < G2_BK / 16; ++k)
              umma_f16_ss_2sm(tmem_d, adesc + k * A_KADV, bdesc + k * B_KADV, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage]);
            if (kb + 1 == num_kb) umma_commit_2sm(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5, both CTAs: own 128 rows) =====================
    const uint32_t quarter = warp & 3;
    const bool do_add = (flags & 1u) != 0;
    const bool out_f32 = (flags & 2u) != 0;
    for (uint32_t iter = 0;; ++iter) {
      const uint32_t tile = dyn ? ring.take(leader, lane) : cluster_id + iter * num_clusters;
      if (tile >= num_tiles) break;
      uint32_t m_blk, n_blk;
      tile_coords2(tile, num_m, num_n, group_m, m_blk, n_blk);
      const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t row = m_blk * 2 * G2_BM + rank * G2_BM + quarter * 32 + lane;
      const uint32_t col0 = n_blk * TILE_N;
      const uint32_t taddr = tmem_base + ((quarter * 32) << 16) + acc * G2_BN;
      if constexpr (EPI == EPI_SWIGLU) {
        __nv_bfloat16* gu = reinterpret_cast<__nv_bfloat16*>(Cout) + static_cast<int64_t>(row) * ldc;
        __nv_bfloat16* hrow = reinterpret_cast<__nv_bfloat16*>(ea.aux) + static_cast<int64_t>(row) * ea.ld_aux;
#pragma unroll 1
        for (uint32_t c = 0; c < 128; c += 32) {
          uint32_t g[32], u[32];
          tmem_ld_32x32b_x32(taddr + c, g);
          tmem_ld_32x32b_x32(taddr + 128 + c, u);
          tmem_ld_wait();
          const uint32_t col = col0 + c;
          if (row < M && col < ea.F) {   // F % 32 == 0 is required by the host wrapper
#pragma unroll
            for (uint32_t j = 0; j < 32; j += 8) {
              uint4 og, ou, oh;
              uint32_t* pg = &og.x; uint32_t* pu = &ou.x; uint32_t* ph = &oh.x;
#pragma unroll
              for (uint32_t q = 0; q < 4; ++q) {
                pg[q] = pack_bf16x2(__uint_as_float(g[j + 2 * q]), __uint_as_float(g[j + 2 * q + 1]));
                pu[q] = pack_bf16x2(__uint_as_float(u[j + 2 * q]), __uint_as_float(u[j + 2 * q + 1]));
                ph[q] = pack_bf16x2(silu_bf16(bf16_lo(pg[q])) * bf16_lo(pu[q]), silu_bf16(bf16_hi(pg[q])) * bf16_hi(pu[q]));
              }
              if ((flags & 4u) == 0) {                                   // flag 4: inference, do not keep g|u
                *reinterpret_cast<uint4*>(gu + col + j) = og;
                *reinterpret_cast<uint4*>(gu + ea.F + col + j) = ou;
              }
              *reinterpret_cast<uint4*>(hrow + col + j) = oh;
            }
          }
        }
      } else if constexpr (EPI == EPI_DSWIGLU) {
        __nv_bfloat16* dgu = reinterpret_cast<__nv_bfloat16*>(Cout) + static_cast<int64_t>(row) * ldc;
        const __nv_bfloat16* gu = reinterpret_cast<const __nv_bfloat16*>(ea.aux) + static_cast<int64_t>(row) * ea.ld_aux;
#pragma unroll 1
        for (uint32_t c = 0; c < G2_BN; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c, v);
          tmem_ld_wait();
          const uint32_t col = col0 + c;
          if (row < M && col < ea.F) {
#pragma unroll
            for (uint32_t j = 0; j < 32; j += 8) {
              const uint4 gg = *reinterpret_cast<const uint4*>(gu + col + j);
              const uint4 uu = *reinterpret_cast<const uint4*>(gu + ea.F + col + j);
              const uint32_t* pg = &gg.x; const uint32_t* pu = &uu.x;
              uint4 og, ou;
              uint32_t* qg = &og.x; uint32_t* qu = &ou.x;
#pragma unroll
              for (uint32_t q = 0; q < 4; ++q) {
                float dg[2], du[2];
#pragma unroll
                for (uint32_t e = 0; e < 2; ++e) {
                  const float d = bf16_round(__uint_as_float(v[j + 2 * q + e]));   // dh as the unfused path stores it
                  const float gv = e ? bf16_hi(pg[q]) : bf16_lo(pg[q]);
                  const float uv = e ? bf16_hi(pu[q]) : bf16_lo(pu[q]);
                  const float sg = 1.f / (1.f + __expf(-gv));
                  dg[e] = d * uv * (sg * (1.f + gv * (1.f - sg)));
                  du[e] = d * (gv * sg);
                }
                qg[q] = pack_bf16x2(dg[0], dg[1]);
                qu[q] = pack_bf16x2(du[0], du[1]);
              }
              *reinterpret_cast<uint4*>(dgu + col + j) = og;
              *reinterpret_cast<uint4*>(dgu + ea.F + col + j) = ou;
            }
          }
        }
      } else if constexpr (EPI == EPI_ROPE) {
        __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(Cout) + static_cast<int64_t>(row) * ldc;
        const int p = (row < M) ? ea.pos[row] : 0;
        const __nv_bfloat16* cs = ea.cos_t + static_cast<int64_t>(p) * 128;
        const __nv_bfloat16* sn = ea.sin_t + static_cast<int64_t>(p) * 128;
#pragma unroll 1
        for (uint32_t hh = 0; hh < G2_BN; hh += 128) {        // two 128-wide heads per tile
#pragma unroll 1
          for (uint32_t c = 0; c < 64; c += 32) {              // chunk c pairs with chunk c + 64 (rotate-half)
            uint32_t a[32], b[32];
            tmem_ld_32x32b_x32(taddr + hh + c, a);
            tmem_ld_32x32b_x32(taddr + hh + 64 + c, b);
            tmem_ld_wait();
            const uint32_t col = col0 + hh + c;
            if (row < M && col < N) {
              const bool rope = col < ea.rope_cols;
#pragma unroll
              for (uint32_t j = 0; j < 32; j += 8) {
                uint4 o1, o2;
                uint32_t* p1 = &o1.x; uint32_t* p2 = &o2.x;
                uint4 c1 = make_uint4(0, 0, 0, 0), s1 = c1, c2 = c1, s2 = c1;
                if (rope) {
                  c1 = *reinterpret_cast<const uint4*>(cs + c + j);      s1 = *reinterpret_cast<const uint4*>(sn + c + j);
                  c2 = *reinterpret_cast<const uint4*>(cs + 64 + c + j); s2 = *reinterpret_cast<const uint4*>(sn + 64 + c + j);
                }
                const uint32_t* pc1 = &c1.x; const uint32_t* ps1 = &s1.x; const uint32_t* pc2 = &c2.x; const uint32_t* ps2 = &s2.x;
#pragma unroll
                for (uint32_t q = 0; q < 4; ++q) {
                  // projection output rounded to bf16 first (the reference ropes the bf16 q/k)
                  const uint32_t xa = pack_bf16x2(__uint_as_float(a[j + 2 * q]), __uint_as_float(a[j + 2 * q + 1]));
                  const uint32_t xb = pack_bf16x2(__uint_as_float(b[j + 2 * q]), __uint_as_float(b[j + 2 * q + 1]));
                  if (rope) {
                    const float y1l = bf16_round(bf16_lo(xa) * bf16_lo(pc1[q])) + bf16_round(-bf16_lo(xb) * bf16_lo(ps1[q]));
                    const float y1h = bf16_round(bf16_hi(xa) * bf16_hi(pc1[q])) + bf16_round(-bf16_hi(xb) * bf16_hi(ps1[q]));
                    const float y2l = bf16_round(bf16_lo(xb) * bf16_lo(pc2[q])) + bf16_round(bf16_lo(xa) * bf16_lo(ps2[q]));
                    const float y2h = bf16_round(bf16_hi(xb) * bf16_hi(pc2[q])) + bf16_round(bf16_hi(xa) * bf16_hi(ps2[q]));
                    p1[q] = pack_bf16x2(y1l, y1h);
                    p2[q] = pack_bf16x2(y2l, y2h);
                  } else {
                    p1[q] = xa;
                    p2[q] = xb;
                  }
                }
                *reinterpret_cast<uint4*>(crow + col + j) = o1;
                *reinterpret_cast<uint4*>(crow + col + 64 + j) = o2;
              }
            }
          }
        }
      } else if constexpr (EPI == EPI_ATTND) {
        // N % 128 == 0 and 16-byte aligned rows are required by the host wrapper
        __nv_bfloat16* drow = reinterpret_cast<__nv_bfloat16*>(Cout) + static_cast<int64_t>(row) * ldc;
        const __nv_bfloat16* orow = reinterpret_cast<const __nv_bfloat16*>(ea.aux) + static_cast<int64_t>(row) * ea.ld_aux;
        float dsum = 0.f;
#pragma unroll 1
        for (uint32_t c = 0; c < G2_BN; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c, v);
          tmem_ld_wait();
          const uint32_t col = col0 + c;
          if (row < M && col < N) {
#pragma unroll
            for (uint32_t j = 0; j < 32; j += 8) {
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]));
              o.y = pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              o.z = pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
              o.w = pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
              const uint4 a = *reinterpret_cast<const uint4*>(orow + col + j);
              dsum += bf16_lo(o.x) * bf16_lo(a.x) + bf16_hi(o.x) * bf16_hi(a.x) + bf16_lo(o.y) * bf16_lo(a.y) + bf16_hi(o.y) * bf16_hi(a.y)
                    + bf16_lo(o.z) * bf16_lo(a.z) + bf16_hi(o.z) * bf16_hi(a.z) + bf16_lo(o.w) * bf16_lo(a.w) + bf16_hi(o.w) * bf16_hi(a.w);
              *reinterpret_cast<uint4*>(drow + col + j) = o;
            }
            if ((c & 127u) == 96u) {                       // a 128-column head is complete
              ea.dvec[static_cast<int64_t>(col >> 7) * M + row] = dsum;
              dsum = 0.f;
            }
          }
        }
      } else {
#pragma unroll 1
      for (uint32_t c = 0; c < G2_BN; c += 32) {
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_remote(&tmem_empty[acc], 0);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // both CTAs done with TMEM and with each other's barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// One self-resetting claim counter per (device, stream): kernels on one stream never overlap, so they can share it;
// different streams get different counters (exact table, no hashing: a collision would let two concurrent GEMMs
// corrupt each other's tile claims).  NV_GEMM_STATIC_SCHED=1 selects the static schedule (A/B measurements).
static int tile_counter_for(cudaStream_t stream, uint32_t** out) {
  constexpr int SLOTS = 64, MAX_DEV = 16;
  static uint32_t* pool[MAX_DEV] = {nullptr};
  static cudaStream_t owner[MAX_DEV][SLOTS];
  static int used[MAX_DEV] = {0};
  static std::mutex mu;
  static int static_sched = -1;
  std::lock_guard<std::mutex> lock(mu);
  if (static_sched < 0) {
    const char* e = getenv("NV_GEMM_STATIC_SCHED");
    static_sched = (e && e[0] == '1') ? 1 : 0;
  }
  *out = nullptr;
  if (static_sched) return NV_OK;
  int dev = 0;
  NV_CUDA(cudaGetDevice(&dev));
  NV_REQUIRE(dev >= 0 && dev < MAX_DEV, "tile_counter_for: device index %d", dev);
  if (!pool[dev]) {
    NV_CUDA(cudaMalloc(&pool[dev], SLOTS * 128));      // one counter per 128-byte line
    NV_CUDA(cudaMemset(pool[dev], 0, SLOTS * 128));
  }
  int slot = -1;
  for (int i = 0; i < used[dev]; ++i)
    if (owner[dev][i] == stream) { slot = i; break; }
  if (slot < 0) {
    if (used[dev] == SLOTS) return NV_OK;              // more streams than counters: static schedule for the extra ones
    slot = used[dev]++;
    owner[dev][slot] = stream;
  }
  *out = pool[dev] + slot * 32;
  return NV_OK;
}

template <bool A_MN, bool B_MN, int EPI = EPI_PLAIN>
static int launch_gemm_2cta(const CUtensorMap& ta, const CUtensorMap& tb, void* C, int64_t ldc, const void* addend,
                            int64_t ld_add, uint32_t M, uint32_t N, uint32_t K, uint32_t flags, cudaStream_t stream,
                            EpiAux ea = EpiAux{}) {
  auto kern = gemm_bf16_tcgen05_2cta<A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    NV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_DYN_BYTES));
    attr_set = true;
  }
  const uint32_t tiles = ceil_div_u32(M, 2 * G2_BM) * ceil_div_u32(N, EPI == EPI_SWIGLU ? 128u : G2_BN);
  uint32_t clusters = min(tiles, (uint32_t)sm_count() / 2);
  uint32_t* counter = nullptr;
  int rc = tile_counter_for(stream, &counter);
  if (rc) return rc;
  // Rasterisation group (pair-tile rows swept together across N).  Measured DRAM reads per launch at T = 10.4k tokens
  // (ncu, tools/gemm_raster.sh): K = 4096 projections 0.40 GB with 16 rows vs 0.68 GB with 8 (operand panels are
  // small, a taller group reuses each B panel more); K >= 11008 problems 2.4 GB with 8 vs 2.7 GB with 16 (the
  // A panels of a 16-row group no longer fit the L2 next to the streamed B panels).
  static int group_env = -1;
  if (group_env < 0) {
    const char* e = getenv("NV_GEMM_GROUP_M");   // rasterisation experiments
    group_env = (e && atoi(e) > 0) ? atoi(e) : 0;
  }
  const int group_m = group_env ? group_env : (K <= 8192 ? 2 * (int)G2_GROUP_M : (int)G2_GROUP_M);
  kern<<<clusters * 2, G2_THREADS, G2_DYN_BYTES, stream>>>(ta, tb, C, ldc, reinterpret_cast<const __nv_bfloat16*>(addend),
                                                            ld_add, M, N, K, flags, ea, counter, (uint32_t)group_m);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

// Called from nv_gemm_bf16 when block_n == 512 (2-CTA, 256x256 pair tile).
int gemm_bf16_2cta_dispatch(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* C,
                            int64_t ldc, const void* addend, int64_t ld_add, int M, int N, int K, unsigned flags,
                            cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = make_tmap_2d(&ta, A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, 64, G2_BM);
  else       rc = make_tmap_2d(&ta, A, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, G2_BK);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_2d(&tb, B, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, 64, G2_BNH);
  else       rc = make_tmap_2d(&tb, B, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, G2_BK);
  if (rc) return rc;
  if (!a_mn && !b_mn) return launch_gemm_2cta<false, false>(ta, tb, C, ldc, addend, ld_add, M, N, K, flags, stream);
  if (!a_mn && b_mn) return launch_gemm_2cta<false, true>(ta, tb, C, ldc, addend, ld_add, M, N, K, flags, stream);
  if (a_mn && b_mn) return launch_gemm_2cta<true, true>(ta, tb, C, ldc, addend, ld_add, M, N, K, flags, stream);
  return launch_gemm_2cta<true, false>(ta, tb, C, ldc, addend, ld_add, M, N, K, flags, stream);
}

}  // namespace nv

// ---- fused-epilogue entry points (K-major activations x nn.Linear weights; full-wave problems) -----------------
extern "C" {

// gu[T, 2F] = x[T,K] · Wgu[2F,K]^T  (gate rows then up rows)  and  h[T,F] = silu(g) * u   in one kernel.
int nv_gemm_swiglu_bf16(const void* x, int64_t ldx, const void* Wgu, int64_t ldw, void* gu, int64_t ldgu, void* h,
                        int64_t ldh, int M, int F, int K, int keep_gu, void* stream) {
  using namespace nv;
  NV_REQUIRE(M > 0 && F > 0 && K > 0 && (F % 128) == 0, "nv_gemm_swiglu_bf16: F must be a multiple of 128 (got %d)", F);
  NV_REQUIRE((ldx & 7) == 0 && (ldw & 7) == 0 && (ldgu & 7) == 0 && (ldh & 7) == 0, "nv_gemm_swiglu_bf16: alignment");
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_2d(&ta, x, 2, (uint64_t)K, (uint64_t)M, (uint64_t)ldx * 2, 64, G2_BM))) return rc;
  if ((rc = make_tmap_2d(&tb, Wgu, 2, (uint64_t)K, (uint64_t)2 * F, (uint64_t)ldw * 2, 64, G2_BNH))) return rc;
  EpiAux ea{};
  ea.aux = h; ea.ld_aux = ldh; ea.F = (uint32_t)F;
  return launch_gemm_2cta<false, false, EPI_SWIGLU>(ta, tb, gu, ldgu, nullptr, 0, M, F, K, keep_gu ? 0u : 4u,
                                                    reinterpret_cast<cudaStream_t>(stream), ea);
}

// dgu[T,2F] = swiglu'(gu) applied to dh = dx[T,D] · Wd[D,F]   (Wd is the nn.Linear weight [D, F]: B stored [K=D, N=F]).
int nv_gemm_dswiglu_bf16(const void* dx, int64_t lddx, const void* Wd, int64_t ldw, const void* gu, int64_t ldgu, void* dgu,
                         int64_t lddgu, int M, int F, int D, void* stream) {
  using namespace nv;
  NV_REQUIRE(M > 0 && F > 0 && D > 0 && (F % 32) == 0, "nv_gemm_dswiglu_bf16: F %% 32");
  NV_REQUIRE((lddx & 7) == 0 && (ldw & 7) == 0 && (ldgu & 7) == 0 && (lddgu & 7) == 0, "nv_gemm_dswiglu_bf16: alignment");
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_2d(&ta, dx, 2, (uint64_t)D, (uint64_t)M, (uint64_t)lddx * 2, 64, G2_BM))) return rc;
  if ((rc = make_tmap_2d(&tb, Wd, 2, (uint64_t)F, (uint64_t)D, (uint64_t)ldw * 2, 64, G2_BK))) return rc;
  EpiAux ea{};
  ea.aux = const_cast<void*>(gu); ea.ld_aux = ldgu; ea.F = (uint32_t)F;
  return launch_gemm_2cta<false, true, EPI_DSWIGLU>(ta, tb, dgu, lddgu, nullptr, 0, M, F, D, 0u,
                                                    reinterpret_cast<cudaStream_t>(stream), ea);
}

// dO[T, D] = dY[T, Dout] · Wo[Dout, D]  (o_proj dgrad; Wo is the nn.Linear weight, B stored [K = Dout, N = D]) and, in the
// same epilogue, dvec[h, t] = sum_d bf16(dO[t, h*128+d]) * O[t, h*128+d]: the D vector of the attention backward.
int nv_gemm_attnd_bf16(const void* dy, int64_t lddy, const void* Wo, int64_t ldw, const void* o, int64_t ldo, void* dout,
                       int64_t lddo, float* dvec, int M, int D, int Dout, void* stream) {
  using namespace nv;
  NV_REQUIRE(M > 0 && D > 0 && Dout > 0 && (D % 128) == 0 && o && dvec, "nv_gemm_attnd_bf16: D %% 128 and O / dvec required");
  NV_REQUIRE((lddy & 7) == 0 && (ldw & 7) == 0 && (ldo & 7) == 0 && (lddo & 7) == 0, "nv_gemm_attnd_bf16: alignment");
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_2d(&ta, dy, 2, (uint64_t)Dout, (uint64_t)M, (uint64_t)lddy * 2, 64, G2_BM))) return rc;
  if ((rc = make_tmap_2d(&tb, Wo, 2, (uint64_t)D, (uint64_t)Dout, (uint64_t)ldw * 2, 64, G2_BK))) return rc;
  EpiAux ea{};
  ea.aux = const_cast<void*>(o); ea.ld_aux = ldo; ea.dvec = dvec;
  return launch_gemm_2cta<false, true, EPI_ATTND>(ta, tb, dout, lddo, nullptr, 0, M, D, Dout, 0u,
                                                  reinterpret_cast<cudaStream_t>(stream), ea);
}

// qkv[T, N] = x[T,K] · Wqkv[N,K]^T with rotate-half RoPE applied to the first rope_cols columns (q and k heads).
int nv_gemm_rope_bf16(const void* x, int64_t ldx, const void* W, int64_t ldw, void* out, int64_t ldo, const int* pos,
                      const void* cos_t, const void* sin_t, int M, int N, int K, int rope_cols, void* stream) {
  using namespace nv;
  NV_REQUIRE(M > 0 && N > 0 && K > 0 && (N % 128) == 0 && (rope_cols % 128) == 0, "nv_gemm_rope_bf16: head alignment");
  NV_REQUIRE((ldx & 7) == 0 && (ldw & 7) == 0 && (ldo & 7) == 0, "nv_gemm_rope_bf16: alignment");
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_2d(&ta, x, 2, (uint64_t)K, (uint64_t)M, (uint64_t)ldx * 2, 64, G2_BM))) return rc;
  if ((rc = make_tmap_2d(&tb, W, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2, 64, G2_BNH))) return rc;
  EpiAux ea{};
  ea.pos = pos; ea.cos_t = reinterpret_cast<const __nv_bfloat16*>(cos_t); ea.sin_t = reinterpret_cast<const __nv_bfloat16*>(sin_t);
  ea.rope_cols = (uint32_t)rope_cols;
  return launch_gemm_2cta<false, false, EPI_ROPE>(ta, tb, out, ldo, nullptr, 0, M, N, K, 0u,
                                                  reinterpret_cast<cudaStream_t>(stream), ea);
}

}  // extern "C"
