// navillm_b200 — causal self-attention forward on tcgen05 (flash-style, packed variable-length rows).
//
// Replaces HF LLaMA's eager attention (reference call site models/modified_lm.py:112-116 ->
// LlamaAttention: scores = QK^T/sqrt(hd) + causal/left-pad mask, fp32 softmax, P·V),
// which materialises [B,32,S,S] scores.  Here the sequences of a batch are PACKED (no pad tokens are
// ever computed): row t of the fused qkv buffer belongs to sequence b with cu_seqlens[b] <= t <
// cu_seqlens[b+1]; the reference's left padding is reproduced by the explicit position ids given to
// the rotary kernel, so results at real tokens are identical (pad positions do not exist here).
//
// One CTA = one (256-query block = two 128-row tiles, head).  head_dim = 128.  The two tiles ping-pong so
// that the tensor pipe (S = QK^T, O += PV) and the MUFU-bound softmax of the other tile overlap:
//   warp 0 (1 lane)   TMA producer: Q0,Q1 once; K_j (1 stage) and V_j (2 stages) per 128-key block
//   warp 1 (1 lane)   tcgen05.mma issuer; issue order per key block: S_0(j+1), PV_0(j), S_1(j+1), PV_1(j)
//   warps 2..5 / 6..9 softmax group of tile 0 / tile 1, one thread per query row: tcgen05.ld S -> online
//                     softmax (exp2, fp32) -> P (bf16) into 128B-swizzled smem (the next MMA's A operand);
//                     rescales O in TMEM when the running max moved; final O / l -> bf16, LSE -> HBM
// TMEM: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512).
// smem: Q 2x32K + K 32K + V 2x32K + P 2x32K = 224 KB.
#include "nv_common.cuh"
#include "nv_host.h"

namespace nv {

constexpr uint32_t ATT_TILE_BYTES = 128 * 128 * 2;  // one 128x128 bf16 tile = two 64-wide swizzle atoms
constexpr uint32_t ATT_ATOM_BYTES = 128 * 128;      // 128 rows x 128 B
constexpr uint32_t ATT_THREADS = 320;
constexpr uint32_t ATT_QROWS = 256;

struct AttnFwdSmem {
  static constexpr uint32_t Q_OFF = 0;                          // 2 tiles
  static constexpr uint32_t K_OFF = Q_OFF + 2 * ATT_TILE_BYTES;  // 1 stage
  static constexpr uint32_t V_OFF = K_OFF + ATT_TILE_BYTES;      // 2 stages
  static constexpr uint32_t P_OFF = V_OFF + 2 * ATT_TILE_BYTES;  // 2 tiles
  static constexpr uint32_t BAR_OFF = P_OFF + 2 * ATT_TILE_BYTES;
  // q_full, k_full, k_empty, v_full[2], v_empty[2], s_full[2], p_ready[2], pv_done[2]
  static constexpr uint32_t NUM_BARS = 13;
  static constexpr uint32_t TOTAL = BAR_OFF + NUM_BARS * 8 + 16;
  static constexpr uint32_t DYN_BYTES = TOTAL + 1024;
};

// Map a flat block id to (sequence, 256-row query block); heavy (late) blocks are launched first.
__device__ __forceinline__ bool locate_qblock(const int* __restrict__ cu, int B, uint32_t blk, int& seq_start,
                                              int& seq_len, uint32_t& qblk, int& seq_idx) {
  for (int b = 0; b < B; ++b) {
    const int s = cu[b], len = cu[b + 1] - s;
    const uint32_t nb = (len + ATT_QROWS - 1) / ATT_QROWS;
    if (blk < nb) { seq_start = s; seq_len = len; qblk = blk; seq_idx = b; return true; }
    blk -= nb;
  }
  return false;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ O, int64_t ldo,
                float* __restrict__ lse, const int* __restrict__ cu_seqlens, int B, int T, float scale,
                const int* __restrict__ kv_start, const int* __restrict__ kv_len) {
  using L = AttnFwdSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + L::Q_OFF;
  uint8_t* sK = smem + L::K_OFF;
  uint8_t* sV = smem + L::V_OFF;
  uint8_t* sP = smem + L::P_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* v_empty = bars + 5;   // [2]
  uint64_t* s_full = bars + 7;    // [2] per tile
  uint64_t* p_ready = bars + 9;   // [2]
  uint64_t* pv_done = bars + 11;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  const uint32_t warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const uint32_t head = blockIdx.y;

  int seq_start = 0, seq_len = 0, seq_idx = 0;
  uint32_t qblk = 0;
  if (!locate_qblock(cu_seqlens, B, gridDim.x - 1 - blockIdx.x, seq_start, seq_len, qblk, seq_idx)) return;  // CTA-uniform
  // Keys: by default the sequence's own rows (self-attention over the packed batch).  With kv_start / kv_len the
  // keys of sequence b are rows kv_start[b] .. +kv_len[b] of the K/V tensors (a KV cache holding an already encoded
  // prefix followed by the new rows) and the queries are its LAST seq_len positions: query i sees keys <= dk + i,
  // dk = kv_len - seq_len.
  const int kv0 = kv_start ? kv_start[seq_idx] : seq_start;
  const uint32_t dk = kv_len ? (uint32_t)(kv_len[seq_idx] - seq_len) : 0u;
  const uint32_t q0 = qblk * ATT_QROWS;                         // first query row (inside the sequence)
  const bool tile1_on = (q0 + 128) < (uint32_t)seq_len;         // second tile has at least one real row
  // key blocks seen by a tile: 0 .. ceil((dk + last real row + 1) / 128) - 1   (dk = 0: 2*qblk+1 and 2*qblk+2)
  const uint32_t nb0 = (dk + min(q0 + 128, (uint32_t)seq_len) + 127) / 128;
  const uint32_t nb1 = tile1_on ? (dk + min(q0 + 256, (uint32_t)seq_len) + 127) / 128 : 0u;
  const uint32_t n_blocks = max(nb0, nb1);
  auto tile_on = [&](uint32_t t, uint32_t j) -> bool { return j < (t == 0 ? nb0 : nb1); };
  auto last_tile = [&](uint32_t j) -> uint32_t { return tile_on(1, j) ? 1u : 0u; };
  auto first_tile = [&](uint32_t j) -> uint32_t { return tile_on(0, j) ? 0u : 1u; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1); mbar_init(k_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 128); mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // (whole warp runs the loop, one elected lane issues: see elect_one() in nv_common.cuh)
    const int32_t qcol = head * 128;
    const int32_t qrow0 = seq_start + q0;
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        tma_load_2d(sQ + t * ATT_TILE_BYTES, &tmap_q, q_full, qcol, qrow0 + t * 128);
        tma_load_2d(sQ + t * ATT_TILE_BYTES + ATT_ATOM_BYTES, &tmap_q, q_full, qcol + 64, qrow0 + t * 128);
      }
    }
    __syncwarp();
    for (uint32_t j = 0; j < n_blocks; ++j) {
      const int32_t krow0 = kv0 + j * 128;
      mbar_wait(k_empty, (j & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(k_full, ATT_TILE_BYTES);
        tma_load_2d(sK, &tmap_k, k_full, qcol, krow0);
        tma_load_2d(sK + ATT_ATOM_BYTES, &tmap_k, k_full, qcol + 64, krow0);
      }
      __syncwarp();
      const uint32_t st = j & 1;
      mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[st], ATT_TILE_BYTES);
        tma_load_2d(sV + st * ATT_TILE_BYTES, &tmap_v, &v_full[st], qcol, krow0);
        tma_load_2d(sV + st * ATT_TILE_BYTES + ATT_ATOM_BYTES, &tmap_v, &v_full[st], qcol + 64, krow0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The whole warp walks the (uniform) schedule; each group of tcgen05.mma + commit is issued by the elected lane,
    // with every operand in uniform registers so that the 32/64-cycle MMAs go out back to back.
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, 0, 1);  // P (K-major) x V (MN-major: hd contiguous)
    const uint64_t qd0 = umma_smem_desc_sw128(smem_u32(sQ), 0, 1024);
    const uint64_t pd0 = umma_smem_desc_sw128(smem_u32(sP), 0, 1024);
    const uint64_t kd0 = umma_smem_desc_sw128(smem_u32(sK), 0, 1024);
    // V tile: rows = keys (K dim), 128 B of hd per row per atom; atoms (hd halves) ATT_ATOM_BYTES apart
    const uint64_t vd0 = umma_smem_desc_sw128(smem_u32(sV), ATT_ATOM_BYTES, 1024);
    auto issue_s = [&](uint32_t t, uint64_t* extra_commit) {
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + t * 128;
#pragma unroll
        for (uint32_t ka = 0; ka < 2; ++ka)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            umma_f16_ss(d_tmem, qd0 + ((t * ATT_TILE_BYTES + ka * ATT_ATOM_BYTES) >> 4) + ks * 2,
                        kd0 + ((ka * ATT_ATOM_BYTES) >> 4) + ks * 2, idesc_s, (ka | ks) ? 1u : 0u);
        umma_commit(&s_full[t]);
        if (extra_commit) umma_commit(extra_commit);
      }
      __syncwarp();
    };
    auto issue_pv = [&](uint32_t t, uint32_t st, bool accumulate, uint64_t* extra_commit) {
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + 256 + t * 128;
#pragma unroll
        for (uint32_t ka = 0; ka < 2; ++ka)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)   // key rows ka*64 + ks*16 -> byte offset * 128 >> 4
            umma_f16_ss(d_tmem, pd0 + ((t * ATT_TILE_BYTES + ka * ATT_ATOM_BYTES) >> 4) + ks * 2,
                        vd0 + ((st * ATT_TILE_BYTES) >> 4) + ((ka * 64 + ks * 16) * 128 >> 4), idesc_pv,
                        (accumulate || (ka | ks)) ? 1u : 0u);
        umma_commit(&pv_done[t]);
        if (extra_commit) umma_commit(extra_commit);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(k_full, 0);
    tc_fence_after();
    for (uint32_t t = 0; t < 2; ++t)
      if (tile_on(t, 0)) issue_s(t, t == last_tile(0) ? k_empty : nullptr);
    for (uint32_t j = 0; j < n_blocks; ++j) {
      const uint32_t st = j & 1;
      const bool have_next = (j + 1 < n_blocks);
      bool k_next_waited = false, v_waited = false;
      for (uint32_t t = 0; t < 2; ++t) {
        if (!tile_on(t, j)) continue;
        mbar_wait(&p_ready[t], j & 1);           // tile t processes every block 0..its last, so its phase index is j
        tc_fence_after();
        if (have_next && tile_on(t, j + 1)) {
          if (!k_next_waited) { mbar_wait(k_full, (j + 1) & 1); tc_fence_after(); k_next_waited = true; }
          issue_s(t, t == last_tile(j + 1) ? k_empty : nullptr);   // S_t(j+1) overlaps the other tile's softmax
        }
        if (!v_waited) { mbar_wait(&v_full[st], (j >> 1) & 1); tc_fence_after(); v_waited = true; }
        issue_pv(t, st, j > 0, t == last_tile(j) ? &v_empty[st] : nullptr);
      }
      // a block seen only by tile 1 whose S could not be issued from tile 0's branch: handled above because
      // tile_on(1, j+1) implies tile_on(1, j); tile 0 dropping out at the last block needs no S.
    }
  } else {
    // ================================ softmax groups ================================
    const uint32_t t = (warp - 2) >> 2;                   // tile of this group
    const uint32_t quarter = warp & 3;
    const uint32_t r = quarter * 32 + lane;               // row inside the tile
    const uint32_t lane_off = (quarter * 32) << 16;       // TMEM lane field
    const uint32_t tmem_S = tmem_base + t * 128, tmem_O = tmem_base + 256 + t * 128;
    uint8_t* sPt = sP + t * ATT_TILE_BYTES;
    const float sl2 = scale * 1.4426950408889634f;
    const uint32_t my_blocks = (t == 0) ? nb0 : nb1;
    const uint32_t vis0 = dk + q0 + t * 128;               // last key visible to the tile's first row
    float m_run = -INFINITY, l_run = 0.f;
    for (uint32_t j = 0; j < my_blocks; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      const bool diag = (j * 128 + 127 > vis0);             // block holds keys that some row of the tile must not see
      const int lim = (int)(vis0 + r) - (int)(j * 128);     // this row sees columns 0 .. lim of the block
      // whole S row -> registers (4 x 32 columns), one wait
      uint32_t s[128];
      tmem_ld_32x32b_x32(tmem_S + lane_off + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld_32x32b_x32(tmem_S + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_32x32b_x32(tmem_S + lane_off + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
      tmem_ld_32x32b_x32(tmem_S + lane_off + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
      tmem_ld_wait();
      // four independent running maxima / sums: a single dependent chain of 128 ops would cost ~4 cycles each
      float mx[4] = {m_run, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (uint32_t i = 0; i < 128; ++i)
        if (!diag || (int)i <= lim) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(s[i]));
      const float m_new = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      const float alpha = exp2f((m_run - m_new) * sl2);
      const float mb = m_new * sl2;
      float rsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (uint32_t i = 0; i < 128; i += 2) {
        float p0 = exp2f(__uint_as_float(s[i]) * sl2 - mb);
        float p1 = exp2f(__uint_as_float(s[i + 1]) * sl2 - mb);
        if (diag) {
          if ((int)i > lim) p0 = 0.f;
          if ((int)i + 1 > lim) p1 = 0.f;
        }
        // the row sum uses the bf16-rounded probabilities that the PV product actually consumes
        const uint32_t pk = pack_bf16x2(p0, p1);
        s[i >> 1] = pk;
        rsum[(i >> 1) & 3] += bf16_lo(pk) + bf16_hi(pk);
      }
      const float rs = (rsum[0] + rsum[1]) + (rsum[2] + rsum[3]);
      if (j > 0) mbar_wait(&pv_done[t], (j - 1) & 1);   // P_t buffer and O_t accumulator free again
      tc_fence_after();
#pragma unroll
      for (uint32_t c = 0; c < 16; ++c) {                // 16 chunks of 8 bf16 (16 B)
        uint8_t* atom = sPt + (c >> 3) * ATT_ATOM_BYTES;
        *reinterpret_cast<uint4*>(atom + sw128_offset(r, c & 7)) = make_uint4(s[c * 4], s[c * 4 + 1], s[c * 4 + 2], s[c * 4 + 3]);
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
        for (uint32_t c = 0; c < 128; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_O + lane_off + c, v);
          tmem_ld_wait();
#pragma unroll
          for (uint32_t i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32b_x32(tmem_O + lane_off + c, v);
        }
        tmem_st_wait();
      }
      fence_proxy_async_smem();  // P stores (generic proxy) -> visible to tcgen05.mma (async proxy)
      tc_fence_before();
      mbar_arrive(&p_ready[t]);
    }
    // ---- epilogue ----
    // Two phases so that HBM sees full 256-byte rows: each thread (= query row) rounds its normalised O row to bf16 into
    // the group's P tile (free after the last PV; 16-byte chunks XOR-swizzled by the row: conflict-free for the row-per-
    // thread writes and the row-per-warp reads), then each warp streams whole rows out with 8-byte coalesced stores.
    if (my_blocks > 0) {
      mbar_wait(&pv_done[t], (my_blocks - 1) & 1);
      tc_fence_after();
      const uint32_t qi = q0 + t * 128 + r;
      const bool valid = qi < (uint32_t)seq_len;
      const float inv_l = 1.f / l_run;
      const int64_t tok = (int64_t)seq_start + qi;
      uint8_t* srow = sPt + r * 256;
#pragma unroll 1
      for (uint32_t c = 0; c < 128; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_O + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (uint32_t i = 0; i < 32; i += 8) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[i + 0]) * inv_l, __uint_as_float(v[i + 1]) * inv_l);
          o.y = pack_bf16x2(__uint_as_float(v[i + 2]) * inv_l, __uint_as_float(v[i + 3]) * inv_l);
          o.z = pack_bf16x2(__uint_as_float(v[i + 4]) * inv_l, __uint_as_float(v[i + 5]) * inv_l);
          o.w = pack_bf16x2(__uint_as_float(v[i + 6]) * inv_l, __uint_as_float(v[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(srow + ((((c + i) >> 3) ^ (r & 15)) << 4)) = o;
        }
      }
      if (valid && lse) lse[(int64_t)head * T + tok] = m_run * scale + __logf(l_run);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");        // the tile's four softmax warps
      const uint32_t rows_valid = min(128u, (uint32_t)seq_len - (q0 + t * 128));
      const uint32_t wq = warp & 3;
      __nv_bfloat16* obase = O + ((int64_t)seq_start + q0 + t * 128) * ldo + head * 128;
      for (uint32_t rr = wq; rr < rows_valid; rr += 4) {
        const uint32_t chunk = (lane >> 1) ^ (rr & 15);
        const uint2 x = *reinterpret_cast<const uint2*>(sPt + rr * 256 + (chunk << 4) + ((lane & 1) << 3));
        *reinterpret_cast<uint2*>(obase + (int64_t)rr * ldo + lane * 4) = x;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace nv

// q, k, v: bf16 row-major [T, *] views with leading dimensions ldq/ldk/ldv (elements); head h occupies
// columns [h*128, (h+1)*128).  o: [T, H*128] bf16 (ldo).  lse: [H, T] fp32 or null.
// cu_seqlens: device int32 [B+1]; total_qblocks = sum_b ceil(len_b / 128) (the host knows the lengths);
// the kernel itself works on 256-row blocks: the grid is sized from an upper bound and surplus CTAs exit.
static int attn_fwd_launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                           int64_t ldo, float* lse, const int* cu_seqlens, const int* kv_start, const int* kv_len, int B, int T,
                           int Tkv, int H, int head_dim, int total_qblocks, float scale, void* stream) {
  using namespace nv;
  NV_REQUIRE(head_dim == 128, "nv_attn_fwd: head_dim must be 128 (got %d)", head_dim);
  NV_REQUIRE(B > 0 && T > 0 && H > 0 && total_qblocks > 0, "nv_attn_fwd: empty problem");
  NV_REQUIRE((ldo & 7) == 0, "nv_attn_fwd: ldo %% 8");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_2d(&tq, q, 2, (uint64_t)H * 128, (uint64_t)T, (uint64_t)ldq * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&tk, k, 2, (uint64_t)H * 128, (uint64_t)Tkv, (uint64_t)ldk * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&tv, v, 2, (uint64_t)H * 128, (uint64_t)Tkv, (uint64_t)ldv * 2, 64, 128))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    NV_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnFwdSmem::DYN_BYTES));
    attr_set = true;
  }
  // sum_b ceil(len_b/256) <= ceil((total_qblocks + B) / 2): launch that many, CTAs beyond the real count return
  const int grid_x = (total_qblocks + B + 1) / 2;
  dim3 grid(grid_x, H);
  attn_fwd_kernel<<<grid, ATT_THREADS, AttnFwdSmem::DYN_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(
      tq, tk, tv, reinterpret_cast<__nv_bfloat16*>(o), ldo, lse, cu_seqlens, B, T, scale, kv_start, kv_len);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

extern "C" int nv_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                           int64_t ldo, float* lse, const int* cu_seqlens, int B, int T, int H, int head_dim,
                           int total_qblocks, float scale, void* stream) {
  return attn_fwd_launch(q, ldq, k, ldk, v, ldv, o, ldo, lse, cu_seqlens, nullptr, nullptr, B, T, T, H, head_dim,
                         total_qblocks, scale, stream);
}

// Suffix ("append") attention over a KV cache: the Tq packed query rows of sequence b (cu_seqlens) are the LAST
// positions of a context whose keys/values are rows kv_start[b] .. kv_start[b] + kv_len[b] of the cache tensors
// (Tkv rows in total; kv_len[b] >= query count, already holding the new rows' K/V).  Rows of the cache past kv_len
// must be finite (allocate it zeroed): they are masked, but 0 * NaN would poison the P·V product.
extern "C" int nv_attn_fwd_kv(const void* q, int64_t ldq, const void* kcache, int64_t ldk, const void* vcache, int64_t ldv,
                              void* o, int64_t ldo, float* lse, const int* cu_seqlens, const int* kv_start,
                              const int* kv_len, int B, int Tq, int Tkv, int H, int head_dim, int total_qblocks, float scale,
                              void* stream) {
  using namespace nv;
  NV_REQUIRE(kv_start && kv_len, "nv_attn_fwd_kv: kv_start / kv_len are required");
  return attn_fwd_launch(q, ldq, kcache, ldk, vcache, ldv, o, ldo, lse, cu_seqlens, kv_start, kv_len, B, Tq, Tkv, H,
                         head_dim, total_qblocks, scale, stream);
}
