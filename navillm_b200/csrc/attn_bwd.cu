// navillm_b200 — causal self-attention backward on tcgen05 (packed variable-length rows), pipelined.
//
// Hand-written twin of attn_fwd.cu; replaces the autograd backward of HF LLaMA's eager attention
// (reference: loss.backward() call sites tasks/agents/mp3d_agent.py:750-757, tasks/agents/llava.py:38-40
// through models/modified_lm.py:112-116).
//
// With P = exp(S*scale - LSE), D_i = sum_d dO_id O_id, dS = P o (dP - D) * scale:
//     dV = P^T dO      dK = dS^T Q      dQ = dS K
// Two deterministic passes (no atomics).  Both walk the partner dimension in 64-wide SUB-blocks so that the
// score tiles are only 64 TMEM columns and can be double-buffered: the tensor pipe computes the scores of
// sub-block n+1 while the compute warps turn sub-block n into P / dS and the accumulate MMAs of n run.
//
//   dq pass   CTA = (128-query block, head).  Q,dO resident; K_n,V_n [64 keys x 128] streamed (3 stages).
//             S = Q K_n^T, dP = dO V_n^T  [128 q x 64 keys] -> dS (bf16, K-major A operand) -> dQ += dS K_n.
//             TMEM: S[2] dP[2] (4 x 64 columns) + dQ (128).
//   dkdv pass CTA = (128-key block, head).  K,V resident; Q_n,dO_n [64 queries x 128] streamed (2 stages).
//             TRANSPOSED scores so that TMEM lanes are keys: S^T = K Q_n^T, dP^T = V dO_n^T [128 keys x 64 q]
//             -> P^T, dS^T (bf16, K-major A operands) -> dV += P^T dO_n, dK += dS^T Q_n.
//             TMEM: S^T[2] dP^T[2] (4 x 64) + dV (128) + dK (128) = 512 columns.
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 compute (two warps per TMEM lane quarter, each owning
// 32 of the 64 score columns; no row reductions are needed in the backward: LSE and D are precomputed).
// The inverse rotary embedding of dQ / dK is applied in the epilogue (optional).
#include "nv_common.cuh"
#include "nv_host.h"

namespace nv {

constexpr uint32_t AB_THREADS = 320;
constexpr uint32_t AB_T128 = 128 * 128 * 2;  // [128 rows x 128 hd] bf16 = two 16 KB swizzle atoms (hd halves)
constexpr uint32_t AB_A128 = 128 * 128;      // one atom of a 128-row tile
constexpr uint32_t AB_T64 = 64 * 128 * 2;    // [64 rows x 128 hd] bf16 = two 8 KB atoms
constexpr uint32_t AB_A64 = 64 * 128;        // one atom of a 64-row tile
constexpr uint32_t AB_PS = 128 * 64 * 2;     // [128 rows x 64 cols] bf16 = one 16 KB atom (P / dS operands)
constexpr float LOG2E = 1.4426950408889634f;

// Developer phase trace (compiled in only with -DNV_ATTN_TRACE, see tools/attn_trace.py): SM-clock stamps of the
// pipeline events of the CTAs of one head, read back through nv_debug_attn_trace().
#ifdef NV_ATTN_TRACE
__device__ unsigned long long g_attn_trace[2][256 * 64];
#define AB_TR(k, slot) do { if (blockIdx.y == 5 && blockIdx.x < 256) g_attn_trace[k][blockIdx.x * 64 + (slot)] = clock64(); } while (0)
#define AB_TRV(k, slot, v) do { if (blockIdx.y == 5 && blockIdx.x < 256) g_attn_trace[k][blockIdx.x * 64 + (slot)] = (v); } while (0)
#else
#define AB_TR(k, slot) do {} while (0)
#define AB_TRV(k, slot, v) do {} while (0)
#endif

__device__ __forceinline__ bool locate_block_bwd(const int* __restrict__ cu, int B, uint32_t blk, int& seq_start,
                                                 int& seq_len, uint32_t& idx) {
  for (int b = 0; b < B; ++b) {
    const int s = cu[b], len = cu[b + 1] - s;
    const uint32_t nb = (len + 127) / 128;
    if (blk < nb) { seq_start = s; seq_len = len; idx = blk; return true; }
    blk -= nb;
  }
  return false;
}

// D[h, t] = sum_d dO[t, h*128+d] * O[t, h*128+d]   (one warp per (t, h))
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ O, int64_t ldo,
                                     const __nv_bfloat16* __restrict__ dO, int64_t lddo, float* __restrict__ D, int T,
                                     int H) {
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (int64_t)T * H) return;
  const int h = gw % H;
  const int64_t t = gw / H;
  const uint2 a = *reinterpret_cast<const uint2*>(O + t * ldo + h * 128 + lane * 4);
  const uint2 b = *reinterpret_cast<const uint2*>(dO + t * lddo + h * 128 + lane * 4);
  float s = bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) +
            bf16_hi(a.y) * bf16_hi(b.y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) D[(int64_t)h * T + t] = s;
}

// Load a [rows x 128 hd] tile as 64-row x 64-col boxes into the two-atom layout (atom = hd half).
template <uint32_t ROWS>
__device__ __forceinline__ void load_tile(uint8_t* dst, const CUtensorMap* m, uint64_t* bar, int32_t col, int32_t row) {
  constexpr uint32_t ATOM = ROWS * 128;
#pragma unroll
  for (uint32_t a = 0; a < 2; ++a)
#pragma unroll
    for (uint32_t r = 0; r < ROWS / 64; ++r)
      tma_load_2d(dst + a * ATOM + r * (64 * 128), m, bar, col + a * 64, row + r * 64);
}

// Epilogue shared by both passes, in two phases so that HBM sees full 256-byte rows (a row-per-thread store - what the
// TMEM lane mapping suggests - touches 32 different rows per instruction and, with the per-row cos/sin loads of the fused
// inverse RoPE, cost 7.5k of a dk/dv CTA's ~28k cycles in the in-kernel trace):
//   stage_acc_tile   thread = accumulator row: tcgen05.ld its 2 x 32 columns (chunks c0 and c0 + 64), round to bf16,
//                    write them into a padded row-major smem tile (272-byte rows: conflict-free 16-byte stores);
//   flush_acc_tile   warp = row: lane l owns columns 4l..4l+3, its rotate-half partner sits in lane l^16; optional
//                    rotation by -theta[pos] with the rounding points of rope_kernel(sign = -1); 8-byte coalesced stores.
// The 8 compute warps synchronise on named barrier 1 between the phases; the staging buffers are operand stages that
// are free once the `done` barrier has fired.
constexpr uint32_t AB_STAGE_LD = 272;                        // bytes per staged row (256 + 16 padding)
constexpr uint32_t AB_STAGE_BYTES = 128 * AB_STAGE_LD;       // 34 816 B per accumulator

__device__ __forceinline__ void compute_warps_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void stage_acc_tile(uint32_t tm, uint32_t lane_off, uint32_t c0, uint32_t r, uint8_t* stage) {
  uint32_t a[32], b[32];
  tmem_ld_32x32b_x32(tm + lane_off + c0, a);
  tmem_ld_32x32b_x32(tm + lane_off + c0 + 64, b);
  tmem_ld_wait();
  uint8_t* row = stage + r * AB_STAGE_LD;
#pragma unroll
  for (uint32_t i = 0; i < 32; i += 8) {
    uint4 o1, o2;
    o1.x = pack_bf16x2(__uint_as_float(a[i]), __uint_as_float(a[i + 1]));     o1.y = pack_bf16x2(__uint_as_float(a[i + 2]), __uint_as_float(a[i + 3]));
    o1.z = pack_bf16x2(__uint_as_float(a[i + 4]), __uint_as_float(a[i + 5])); o1.w = pack_bf16x2(__uint_as_float(a[i + 6]), __uint_as_float(a[i + 7]));
    o2.x = pack_bf16x2(__uint_as_float(b[i]), __uint_as_float(b[i + 1]));     o2.y = pack_bf16x2(__uint_as_float(b[i + 2]), __uint_as_float(b[i + 3]));
    o2.z = pack_bf16x2(__uint_as_float(b[i + 4]), __uint_as_float(b[i + 5])); o2.w = pack_bf16x2(__uint_as_float(b[i + 6]), __uint_as_float(b[i + 7]));
    *reinterpret_cast<uint4*>(row + (c0 + i) * 2) = o1;
    *reinterpret_cast<uint4*>(row + (c0 + 64 + i) * 2) = o2;
  }
}

// cw: compute-warp index 0..7 (rows cw, cw + 8, ...); rows_valid: real rows of the tile; dst: row 0 of the tile at the
// head's column offset; pos: rotary positions of the tile's rows (null = no rotation).
__device__ __forceinline__ void flush_acc_tile(const uint8_t* stage, uint32_t cw, uint32_t lane, uint32_t rows_valid,
                                               __nv_bfloat16* dst, int64_t ld, const int* __restrict__ pos,
                                               const __nv_bfloat16* __restrict__ cos_t,
                                               const __nv_bfloat16* __restrict__ sin_t) {
  const float sgn = lane < 16 ? 1.f : -1.f;
  for (uint32_t r = cw; r < rows_valid; r += 8) {
    uint2 x = *reinterpret_cast<const uint2*>(stage + r * AB_STAGE_LD + lane * 8);
    if (pos != nullptr) {
      const uint32_t px = __shfl_xor_sync(0xffffffffu, x.x, 16), py = __shfl_xor_sync(0xffffffffu, x.y, 16);
      const int p = pos[r];
      const uint2 c = *reinterpret_cast<const uint2*>(cos_t + (int64_t)p * 128 + lane * 4);
      const uint2 sn = *reinterpret_cast<const uint2*>(sin_t + (int64_t)p * 128 + lane * 4);
      // y[d] = bf16(x[d] cos) + bf16(x[d+64] sin),  y[d+64] = bf16(x[d+64] cos) + bf16(-x[d] sin)
      const float y0 = bf16_round(bf16_lo(x.x) * bf16_lo(c.x)) + bf16_round(sgn * bf16_lo(px) * bf16_lo(sn.x));
      const float y1 = bf16_round(bf16_hi(x.x) * bf16_hi(c.x)) + bf16_round(sgn * bf16_hi(px) * bf16_hi(sn.x));
      const float y2 = bf16_round(bf16_lo(x.y) * bf16_lo(c.y)) + bf16_round(sgn * bf16_lo(py) * bf16_lo(sn.y));
      const float y3 = bf16_round(bf16_hi(x.y) * bf16_hi(c.y)) + bf16_round(sgn * bf16_hi(py) * bf16_hi(sn.y));
      x.x = pack_bf16x2(y0, y1);
      x.y = pack_bf16x2(y2, y3);
    }
    *reinterpret_cast<uint2*>(dst + (int64_t)r * ld + lane * 4) = x;
  }
}

// ======================================================================================================
// dq pass -- A operands in TENSOR MEMORY (tcgen05.mma ".ts" form)
// ======================================================================================================
// Round 1 ran this pass with every operand in shared memory and measured it smem-bandwidth-bound: 176 KB of shared-memory
// traffic per 64-key sub-block (1375 cycles at 128 B/clk) against 768 tensor cycles, ~1550 measured (in-kernel trace).
// Q and dO are the A operands of EVERY score MMA of the CTA, and dS is produced by the compute warps from TMEM data:
//   * Q, dO  are written once into TMEM as bf16 (2 x 64 columns): TMA stages the two tiles in shared memory, each compute
//            thread copies half of its own row (a row-per-thread load straight from HBM cost 5300 cycles of prologue)
//   * dS(n)  is written with tcgen05.st straight over the S(n) columns it was computed from (no smem store, no
//            fence.proxy.async), laid out so that every warp overwrites only columns it has already read
// so per sub-block only the B operands (K_n, V_n: 48 KB) and the TMA fills (32 KB) touch shared memory, and the N = 64 score
// MMAs run at their 32-cycle tensor floor instead of 48.
//   TMEM: S[2] [0,128)  dP[2] [128,256)  dQ [256,384)  Q [384,448)  dO [448,512);  dS(n) aliases S[n & 1].
//   smem: Q | dO staging (64 KB, once; the epilogue stages dQ there) + 5 stages of (K_n 16 KB | V_n 16 KB).  With the
//   sub-block time down to ~800 cycles a 3-stage ring left only two sub-blocks (< the ~2000-cycle TMA latency) between a
//   stage being freed by dQ(n) and S(n+3) needing its successor (in-kernel trace: the issuer waited for kv_full).
struct DqSmem {
  static constexpr uint32_t NST = 5;
  static constexpr uint32_t QDO_OFF = 0;                          // Q 32K | dO 32K
  static constexpr uint32_t KV_OFF = 2 * AB_T128;                 // NST x (K 16K | V 16K)
  static constexpr uint32_t BAR_OFF = KV_OFF + NST * 2 * AB_T64;
  // kv_full[5], kv_empty[5], sdp_full[2], sdp_empty[2], ds_full[2], res_full, qdo_ready, done
  static constexpr uint32_t NUM_BARS = 19;
  static constexpr uint32_t DYN_BYTES = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
};
static_assert(AB_STAGE_BYTES <= 2 * AB_T128, "dQ staging reuses the Q/dO staging area");
static_assert(DqSmem::DYN_BYTES <= 232448, "dq pass shared memory");

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                   const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                   const float* __restrict__ lse, const float* __restrict__ Dvec, __nv_bfloat16* __restrict__ dq,
                   int64_t lddq, const int* __restrict__ cu_seqlens, int B, int T, float scale,
                   const int* __restrict__ rope_pos, const __nv_bfloat16* __restrict__ cos_t,
                   const __nv_bfloat16* __restrict__ sin_t) {
  using L = DqSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQdO = smem + L::QDO_OFF;
  uint8_t* sKV = smem + L::KV_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* kv_full = bars + 0;     // [5]
  uint64_t* kv_empty = bars + 5;    // [5]
  uint64_t* sdp_full = bars + 10;   // [2]
  uint64_t* sdp_empty = bars + 12;  // [2]
  uint64_t* ds_full = bars + 14;    // [2]
  uint64_t* res_full = bars + 16;
  uint64_t* qdo_ready = bars + 17;
  uint64_t* done = bars + 18;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  const uint32_t warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const uint32_t head = blockIdx.y;
  int seq_start = 0, seq_len = 0;
  uint32_t own = 0;
  if (!locate_block_bwd(cu_seqlens, B, gridDim.x - 1 - blockIdx.x, seq_start, seq_len, own)) return;  // heavy first
  // 64-key sub-blocks 0 .. n_sub-1 cover keys [0, min((own+1)*128, len))
  const uint32_t n_sub = min(2 * own + 2, (uint32_t)(seq_len + 63) / 64);
  if (threadIdx.x == 0) {
    AB_TR(1, 0); AB_TRV(1, 3, n_sub);
#ifdef NV_ATTN_TRACE
    unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); AB_TRV(1, 4, sm);
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); AB_TRV(1, 5, gt);
#endif
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    for (uint32_t i = 0; i < L::NST; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&sdp_empty[i], 8); mbar_init(&ds_full[i], 8); }
    mbar_init(res_full, 1);
    mbar_init(qdo_ready, 8);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_dQ = tmem_base + 256, tmem_Qa = tmem_base + 384, tmem_dOa = tmem_base + 448;
  if (threadIdx.x == 0) AB_TR(1, 1);

  if (warp == 0) {
    const int32_t col = head * 128;
    if (elect_one()) {
      mbar_arrive_expect_tx(res_full, 2 * AB_T128);
      load_tile<128>(sQdO, &tmap_q, res_full, col, seq_start + own * 128);
      load_tile<128>(sQdO + AB_T128, &tmap_do, res_full, col, seq_start + own * 128);
    }
    __syncwarp();
    for (uint32_t n = 0; n < n_sub; ++n) {
      const uint32_t st = n % L::NST, use = n / L::NST;
      mbar_wait(&kv_empty[st], (use & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[st], 2 * AB_T64);
        load_tile<64>(sKV + st * 2 * AB_T64, &tmap_k, &kv_full[st], col, seq_start + n * 64);
        load_tile<64>(sKV + st * 2 * AB_T64 + AB_T64, &tmap_v, &kv_full[st], col, seq_start + n * 64);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // whole warp runs the issue loop (uniform registers, see elect_one()); one elected lane issues
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);    // [128 q] x [64 keys], both K-major over hd
    constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 128, 0, 1);  // dS (K-major over keys) x K_n (MN-major: hd contiguous)
    auto issue_sdp = [&](uint32_t n) {
      const uint32_t st = n % L::NST, b = n & 1;
      mbar_wait(&kv_full[st], (n / L::NST) & 1);
      mbar_wait(&sdp_empty[b], ((n >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t kd = umma_smem_desc_sw128(smem_u32(sKV + st * 2 * AB_T64), 0, 1024);
        const uint64_t vd = kd + (AB_T64 >> 4);
#pragma unroll
        for (uint32_t ka = 0; ka < 2; ++ka)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)      // hd step ka*64 + ks*16 -> A columns (ka*32 + ks*8), B bytes as before
            umma_f16_ts(tmem_base + b * 64, tmem_Qa + ka * 32 + ks * 8, kd + ka * (AB_A64 >> 4) + ks * 2, idesc_s,
                        (ka | ks) ? 1u : 0u);
#pragma unroll
        for (uint32_t ka = 0; ka < 2; ++ka)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            umma_f16_ts(tmem_base + 128 + b * 64, tmem_dOa + ka * 32 + ks * 8, vd + ka * (AB_A64 >> 4) + ks * 2, idesc_s,
                        (ka | ks) ? 1u : 0u);
        umma_commit(&sdp_full[b]);
        if (n < 8) AB_TR(1, 10 + 2 * n);
      }
      __syncwarp();
    };
    mbar_wait(qdo_ready, 0);
    tc_fence_after();
    if (lane == 0) AB_TR(1, 2);
    issue_sdp(0);
    for (uint32_t n = 0; n < n_sub; ++n) {
      if (n + 1 < n_sub) issue_sdp(n + 1);
      const uint32_t st = n % L::NST, b = n & 1;
      mbar_wait(&ds_full[b], (n >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        if (n < 8) AB_TR(1, 11 + 2 * n);
        // dQ += dS K_n : contraction over the 64 keys (4 k-steps).  A = dS(n) in the S[b] columns: keys 0..31 at columns
        // 0..15, keys 32..63 at columns 32..47 (each compute warp overwrote only what it had read).  B rows = keys, 64-wide
        // hd atoms AB_A64 apart.
        const uint64_t km = umma_smem_desc_sw128(smem_u32(sKV + st * 2 * AB_T64), AB_A64, 1024);
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks)
          umma_f16_ts(tmem_dQ, tmem_base + b * 64 + (ks >> 1) * 32 + (ks & 1) * 8, km + ks * 128, idesc_dq, (n | ks) ? 1u : 0u);
        umma_commit(&kv_empty[st]);
        if (n + 1 == n_sub) umma_commit(done);
      }
      __syncwarp();
    }
  } else {
    const uint32_t quarter = warp & 3, half = (warp - 2) >> 2;
    const uint32_t r = quarter * 32 + lane;
    const uint32_t lane_off = (quarter * 32) << 16;
    const float sl2 = scale * LOG2E;
    const uint32_t qi = own * 128 + r;
    const bool row_valid = qi < (uint32_t)seq_len;
    const int64_t tok = (int64_t)seq_start + qi;
    // ---- Q, dO -> TMEM (bf16 A-operand layout): this thread owns hd [half*64, half*64+64) of its row = atom `half` of
    // the TMA-staged tile (two 64-row boxes per atom, 128-byte rows, 16-byte chunks XOR-swizzled by the row) ----
    {
      uint32_t a[32], d[32];
      mbar_wait(res_full, 0);
      const uint32_t rowoff = half * AB_A128 + (r >> 6) * (64 * 128);
      const uint32_t qb = smem_u32(sQdO) + rowoff, ob = qb + AB_T128;
#pragma unroll
      for (uint32_t i = 0; i < 8; ++i) {
        const uint32_t o = sw128_offset(r & 63, i);
        uint4 x, y;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(qb + o));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(y.x), "=r"(y.y), "=r"(y.z), "=r"(y.w) : "r"(ob + o));
        if (!row_valid) { x = make_uint4(0, 0, 0, 0); y = make_uint4(0, 0, 0, 0); }   // rows past the sequence: next sequence's data
        a[4 * i] = x.x; a[4 * i + 1] = x.y; a[4 * i + 2] = x.z; a[4 * i + 3] = x.w;
        d[4 * i] = y.x; d[4 * i + 1] = y.y; d[4 * i + 2] = y.z; d[4 * i + 3] = y.w;
      }
      tmem_st_32x32b_x32(tmem_Qa + lane_off + half * 32, a);
      tmem_st_32x32b_x32(tmem_dOa + lane_off + half * 32, d);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qdo_ready);
    }
    // an invalid row gets LSE = +inf -> p = exp2(-inf) = 0 without a branch
    const float lse_r = row_valid ? lse[(int64_t)head * T + tok] * LOG2E : INFINITY;
    const float d_r = row_valid ? Dvec[(int64_t)head * T + tok] : 0.f;
    const float2 sl2v = make_float2(sl2, sl2), nl = make_float2(-lse_r, -lse_r);
    const float2 scv = make_float2(scale, scale), nds = make_float2(-d_r * scale, -d_r * scale);
    for (uint32_t n = 0; n < n_sub; ++n) {
      const uint32_t b = n & 1;
      mbar_wait(&sdp_full[b], (n >> 1) & 1);
      if (warp == 2 && lane == 0 && n < 8) AB_TR(1, 30 + 2 * n);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      tmem_ld_32x32b_x32(tmem_base + b * 64 + lane_off + half * 32, sv);
      tmem_ld_32x32b_x32(tmem_base + 128 + b * 64 + lane_off + half * 32, dv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sdp_empty[b]);           // (kept for symmetry; program order already protects S[b])
      const bool diag = (n >= 2 * own);                     // sub-blocks that touch the diagonal 128x128 block
      const uint32_t key0 = n * 64 + half * 32;             // first key of this thread's 32 columns
      uint32_t dd[16];
#pragma unroll
      for (uint32_t i = 0; i < 32; i += 2) {
        float2 e = __ffma2_rn(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2v, nl);
        float p0 = exp2f(e.x), p1 = exp2f(e.y);
        if (diag) {
          if (key0 + i > qi) p0 = 0.f;
          if (key0 + i + 1 > qi) p1 = 0.f;
        }
        // dS = P o (dP - D) * scale
        const float2 t = __ffma2_rn(make_float2(__uint_as_float(dv[i]), __uint_as_float(dv[i + 1])), scv, nds);
        const float2 ds = __fmul2_rn(make_float2(p0, p1), t);
        dd[i >> 1] = pack_bf16x2(ds.x, ds.y);
      }
      // dS(n) over the S columns this warp has just consumed: keys half*32.. -> 16 packed columns at half*32
      tmem_st_32x32b_x16(tmem_base + b * 64 + lane_off + half * 32, dd);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[b]);
      if (warp == 2 && lane == 0 && n < 8) AB_TR(1, 31 + 2 * n);
    }
    mbar_wait(done, 0);
    if (warp == 2 && lane == 0) AB_TR(1, 50);
    tc_fence_after();
    uint8_t* stage = sQdO;                                  // Q/dO staging is free since qdo_ready
    stage_acc_tile(tmem_dQ, lane_off, half * 32, r, stage);
    compute_warps_sync();
    const uint32_t rows_valid = min(128u, (uint32_t)seq_len - own * 128);
    const int64_t tok0 = (int64_t)seq_start + own * 128;
    flush_acc_tile(stage, warp - 2, lane, rows_valid, dq + tok0 * lddq + head * 128, lddq,
                   rope_pos != nullptr ? rope_pos + tok0 : nullptr, cos_t, sin_t);
    if (warp == 2 && lane == 0) AB_TR(1, 51);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
  if (threadIdx.x == 32) AB_TR(1, 52);
}

// ======================================================================================================
// dk/dv pass (transposed scores: TMEM lanes = keys)
// ======================================================================================================
struct DkvSmem {
  static constexpr uint32_t NST = 4;                              // 3 stages left 2 sub-blocks of slack < TMA latency (trace); 5 do not fit next to the 2 KB of static smem
  static constexpr uint32_t K_OFF = 0, V_OFF = AB_T128;
  static constexpr uint32_t QO_OFF = 2 * AB_T128;                 // NST x (Q 16K | dO 16K)
  static constexpr uint32_t BAR_OFF = QO_OFF + NST * 2 * AB_T64;
  // res_full, qo_full[NST], qo_empty[NST] (slots for 5), sdp_full[2], sdp_empty[2], pds_full[2], done
  static constexpr uint32_t NUM_BARS = 18;
  static constexpr uint32_t DYN_BYTES = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
};
static_assert(DkvSmem::DYN_BYTES + 2048 <= 232448, "dk/dv pass shared memory (dynamic + s_stats)");

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                    const float* __restrict__ lse, const float* __restrict__ Dvec, __nv_bfloat16* __restrict__ dk,
                    int64_t lddk, __nv_bfloat16* __restrict__ dv, int64_t lddv, const int* __restrict__ cu_seqlens,
                    int B, int T, float scale, const int* __restrict__ rope_pos,
                    const __nv_bfloat16* __restrict__ cos_t, const __nv_bfloat16* __restrict__ sin_t) {
  using L = DkvSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + L::K_OFF;
  uint8_t* sV = smem + L::V_OFF;
  uint8_t* sQO = smem + L::QO_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* res_full = bars + 0;
  uint64_t* qo_full = bars + 1;     // [5]
  uint64_t* qo_empty = bars + 6;    // [5]
  uint64_t* sdp_full = bars + 11;   // [2]
  uint64_t* sdp_empty = bars + 13;  // [2]
  uint64_t* pds_full = bars + 15;   // [2] P^T(n) / dS^T(n) are in TMEM (over S^T[b] / dP^T[b])
  uint64_t* done = bars + 17;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);
  __shared__ float s_stats[8][2][32];   // per compute warp: LSE*log2e and D of its 32 query columns

  const uint32_t warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const uint32_t head = blockIdx.y;
  int seq_start = 0, seq_len = 0;
  uint32_t own = 0;
  if (!locate_block_bwd(cu_seqlens, B, blockIdx.x, seq_start, seq_len, own)) return;   // early key blocks are the heavy ones
  // 64-query sub-blocks first .. n_qsub-1 can see keys of this block (causal): q >= own*128
  const uint32_t first = 2 * own, n_qsub = (uint32_t)(seq_len + 63) / 64;
  const uint32_t n_it = n_qsub - first;
  if (threadIdx.x == 0) {
    AB_TR(0, 0); AB_TRV(0, 3, n_it);
#ifdef NV_ATTN_TRACE
    unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); AB_TRV(0, 4, sm);
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); AB_TRV(0, 5, gt);
#endif
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do);
    mbar_init(res_full, 1);
    for (uint32_t i = 0; i < L::NST; ++i) { mbar_init(&qo_full[i], 1); mbar_init(&qo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&sdp_empty[i], 8); }
    mbar_init(&pds_full[0], 8); mbar_init(&pds_full[1], 8);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_dV = tmem_base + 256, tmem_dK = tmem_base + 384;
  if (threadIdx.x == 0) AB_TR(0, 1);

  if (warp == 0) {
    const int32_t col = head * 128;
    if (elect_one()) {
      mbar_arrive_expect_tx(res_full, 2 * AB_T128);
      load_tile<128>(sK, &tmap_k, res_full, col, seq_start + own * 128);
      load_tile<128>(sV, &tmap_v, res_full, col, seq_start + own * 128);
    }
    __syncwarp();
    for (uint32_t n = 0; n < n_it; ++n) {
      const uint32_t st = n % L::NST, use = n / L::NST;
      mbar_wait(&qo_empty[st], (use & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&qo_full[st], 2 * AB_T64);
        load_tile<64>(sQO + st * 2 * AB_T64, &tmap_q, &qo_full[st], col, seq_start + (first + n) * 64);
        load_tile<64>(sQO + st * 2 * AB_T64 + AB_T64, &tmap_do, &qo_full[st], col, seq_start + (first + n) * 64);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // whole warp runs the issue loop (uniform registers, see elect_one()); one elected lane issues
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);     // [128 keys] x [64 q], both K-major over hd
    constexpr uint32_t idesc_acc = umma_idesc_bf16(128, 128, 0, 1);  // P^T/dS^T (K-major over q) x dO_n/Q_n (MN-major)
    const uint64_t k_k = umma_smem_desc_sw128(smem_u32(sK), 0, 1024);
    const uint64_t v_k = umma_smem_desc_sw128(smem_u32(sV), 0, 1024);
    auto issue_sdp = [&](uint32_t n) {
      const uint32_t st = n % L::NST, b = n & 1;
      mbar_wait(&qo_full[st], (n / L::NST) & 1);
      mbar_wait(&sdp_empty[b], ((n >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qd = umma_smem_desc_sw128(smem_u32(sQO + st * 2 * AB_T64), 0, 1024);
        const uint64_t od = qd + (AB_T64 >> 4);
#pragma unroll
        for (uint32_t ka = 0; ka < 2; ++ka)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            umma_f16_ss(tmem_base + b * 64, k_k + ka * (AB_A128 >> 4) + ks * 2, qd + ka * (AB_A64 >> 4) + ks * 2, idesc_s,
                        (ka | ks) ? 1u : 0u);
#pragma unroll
        for (uint32_t ka = 0; ka < 2; ++ka)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            umma_f16_ss(tmem_base + 128 + b * 64, v_k + ka * (AB_A128 >> 4) + ks * 2, od + ka * (AB_A64 >> 4) + ks * 2,
                        idesc_s, (ka | ks) ? 1u : 0u);
        umma_commit(&sdp_full[b]);
        if (n < 8) AB_TR(0, 10 + 2 * n);
      }
      __syncwarp();
    };
    mbar_wait(res_full, 0);
    if (lane == 0) AB_TR(0, 2);
    issue_sdp(0);
    for (uint32_t n = 0; n < n_it; ++n) {
      if (n + 1 < n_it) issue_sdp(n + 1);
      const uint32_t st = n % L::NST, b = n & 1;
      mbar_wait(&pds_full[b], (n >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        if (n < 8) AB_TR(0, 11 + 2 * n);
        // contraction over the 64 queries (4 k-steps).  A operands in TENSOR MEMORY (.ts form): P^T(n) lies over the
        // S^T[b] columns it was computed from, dS^T(n) over dP^T[b] (queries 0..31 at columns 0..15, 32..63 at 32..47: each
        // compute warp overwrote only what it had read) - no smem store, no proxy fence, no A-operand smem bandwidth.
        // B rows = queries, 64-wide hd atoms AB_A64 apart.
        const uint64_t qm = umma_smem_desc_sw128(smem_u32(sQO + st * 2 * AB_T64), AB_A64, 1024);
        const uint64_t om = qm + (AB_T64 >> 4);
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks)
          umma_f16_ts(tmem_dV, tmem_base + b * 64 + (ks >> 1) * 32 + (ks & 1) * 8, om + ks * 128, idesc_acc, (n | ks) ? 1u : 0u);
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks)
          umma_f16_ts(tmem_dK, tmem_base + 128 + b * 64 + (ks >> 1) * 32 + (ks & 1) * 8, qm + ks * 128, idesc_acc, (n | ks) ? 1u : 0u);
        umma_commit(&qo_empty[st]);
        if (n + 1 == n_it) umma_commit(done);
      }
      __syncwarp();
    }
  } else {
    const uint32_t cw = warp - 2;
    const uint32_t quarter = warp & 3, half = cw >> 2;
    const uint32_t r = quarter * 32 + lane;                 // key row of this thread
    const uint32_t lane_off = (quarter * 32) << 16;
    const float sl2 = scale * LOG2E;
    const uint32_t kj = own * 128 + r;                      // key index inside the sequence
    float* st_l = s_stats[cw][0];
    float* st_d = s_stats[cw][1];
    // LSE / D of the 32 query columns a warp handles in sub-block n: one coalesced load per lane, shared inside the warp
    // through smem.  The loads for sub-block n+1 are issued before sub-block n is processed - on the critical path they
    // cost ~900 cycles of global-load latency per sub-block (in-kernel trace: gap between P^T/dS^T of n and S^T of n+1).
    auto load_stats = [&](uint32_t n, float& l_out, float& d_out) {
      const uint32_t qi = (first + n) * 64 + half * 32 + lane;
      const bool ok = qi < (uint32_t)seq_len;
      const int64_t tq = (int64_t)seq_start + qi;
      // stored NEGATED / pre-scaled so that the loop is two packed FMAs: -inf -> p = exp2(-inf) = 0 for rows past the sequence
      l_out = ok ? -lse[(int64_t)head * T + tq] * LOG2E : -INFINITY;
      d_out = ok ? -Dvec[(int64_t)head * T + tq] * scale : 0.f;
    };
    float cur_l = -INFINITY, cur_d = 0.f;
    const float2 sl2v = make_float2(sl2, sl2), scv = make_float2(scale, scale);
    if (n_it > 0) load_stats(0, cur_l, cur_d);
    for (uint32_t n = 0; n < n_it; ++n) {
      const uint32_t b = n & 1;
      const uint32_t q0 = (first + n) * 64 + half * 32;     // first query of the 32 columns
      __syncwarp();
      st_l[lane] = cur_l;
      st_d[lane] = cur_d;
      __syncwarp();
      if (n + 1 < n_it) load_stats(n + 1, cur_l, cur_d);   // in flight while this sub-block is processed
      mbar_wait(&sdp_full[b], (n >> 1) & 1);
      if (warp == 2 && lane == 0 && n < 8) AB_TR(0, 30 + 2 * n);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      tmem_ld_32x32b_x32(tmem_base + b * 64 + lane_off + half * 32, sv);
      tmem_ld_32x32b_x32(tmem_base + 128 + b * 64 + lane_off + half * 32, dv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sdp_empty[b]);
      const bool diag = (first + n) < 2 * own + 2;          // query sub-blocks inside the diagonal 128x128 block
      uint32_t pp[16], dd[16];
#pragma unroll
      for (uint32_t i = 0; i < 32; i += 2) {
        const float2 l2 = *reinterpret_cast<const float2*>(st_l + i);       // -LSE * log2e of queries i, i+1
        const float2 d2 = *reinterpret_cast<const float2*>(st_d + i);       // -D * scale
        const float2 e = __ffma2_rn(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2v, l2);
        float p0 = exp2f(e.x), p1 = exp2f(e.y);
        if (diag) {                                         // causal: a query sees keys <= itself
          if (q0 + i < kj) p0 = 0.f;
          if (q0 + i + 1 < kj) p1 = 0.f;
        }
        pp[i >> 1] = pack_bf16x2(p0, p1);
        const float2 t = __ffma2_rn(make_float2(__uint_as_float(dv[i]), __uint_as_float(dv[i + 1])), scv, d2);   // (dP - D) * scale
        const float2 ds = __fmul2_rn(make_float2(p0, p1), t);
        dd[i >> 1] = pack_bf16x2(ds.x, ds.y);
      }
      // P^T(n) / dS^T(n) straight into TMEM, over the score columns this warp has just consumed (queries half*32..)
      tmem_st_32x32b_x16(tmem_base + b * 64 + lane_off + half * 32, pp);
      tmem_st_32x32b_x16(tmem_base + 128 + b * 64 + lane_off + half * 32, dd);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[b]);
      if (warp == 2 && lane == 0 && n < 8) AB_TR(0, 31 + 2 * n);
    }
    mbar_wait(done, 0);
    if (warp == 2 && lane == 0) AB_TR(0, 50);
    tc_fence_after();
    uint8_t* stage_v = sK;                                  // resident K/V and the Q/dO stages are free now
    uint8_t* stage_k = sQO;
    stage_acc_tile(tmem_dV, lane_off, half * 32, r, stage_v);
    stage_acc_tile(tmem_dK, lane_off, half * 32, r, stage_k);
    compute_warps_sync();
    const uint32_t rows_valid = min(128u, (uint32_t)seq_len - own * 128);
    const int64_t tok0 = (int64_t)seq_start + own * 128;
    flush_acc_tile(stage_v, cw, lane, rows_valid, dv + tok0 * lddv + head * 128, lddv, nullptr, cos_t, sin_t);
    flush_acc_tile(stage_k, cw, lane, rows_valid, dk + tok0 * lddk + head * 128, lddk,
                   rope_pos != nullptr ? rope_pos + tok0 : nullptr, cos_t, sin_t);
    if (warp == 2 && lane == 0) AB_TR(0, 51);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
  if (threadIdx.x == 32) AB_TR(0, 52);
}

}  // namespace nv

// Developer hook: copy the phase trace of the last nv_attn_bwd (kernel 0 = dk/dv pass, 1 = dq pass) to the host.
// Returns the number of 64-bit words written, 0 when the library was built without -DNV_ATTN_TRACE.
extern "C" int nv_debug_attn_trace(int kernel, unsigned long long* out, int max_words) {
  if (kernel == 2 && out) return nv::attn_fwd_trace_copy(out, max_words);      // forward kernel (attn_fwd.cu)
#ifdef NV_ATTN_TRACE
  if (kernel < 0 || kernel > 1 || !out) return 0;
  const int n = max_words < 256 * 64 ? max_words : 256 * 64;
  if (cudaMemcpyFromSymbol(out, nv::g_attn_trace, (size_t)n * 8, (size_t)kernel * 256 * 64 * 8) != cudaSuccess) return 0;
  return n;
#else
  (void)kernel; (void)out; (void)max_words;
  return 0;
#endif
}

// Inputs: q,k,v (post-RoPE) and o, do as bf16 [T, H*128]-column views; lse [H,T] from the forward.
// dvec: fp32 workspace [H, T] (computed here from o and dout; with o == nullptr it must already hold
// D[h,t] = sum_d dout[t,h,d] o[t,h,d]).  Outputs dq, dk, dv: bf16 views with their own leading dimensions.
// rope_pos/cos_t/sin_t (nullable): when given, dq and dk are rotated back by -theta[pos] in the epilogue, i.e. the
// outputs are the gradients w.r.t. the PRE-RoPE q/k (fuses the backward of the rotary embedding).
extern "C" int nv_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                           const void* o, int64_t ldo, const void* dout, int64_t lddo, const float* lse, float* dvec,
                           void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                           const int* cu_seqlens, int B, int T, int H, int head_dim, int total_blocks, float scale,
                           const int* rope_pos, const void* cos_t, const void* sin_t, void* stream_) {
  using namespace nv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NV_REQUIRE(head_dim == 128, "nv_attn_bwd: head_dim must be 128 (got %d)", head_dim);
  NV_REQUIRE(B > 0 && T > 0 && H > 0 && total_blocks > 0, "nv_attn_bwd: empty problem");
  NV_REQUIRE((lddq & 7) == 0 && (lddk & 7) == 0 && (lddv & 7) == 0 && (ldo & 3) == 0 && (lddo & 3) == 0,
             "nv_attn_bwd: leading-dimension alignment");
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_tmap_2d(&tq, q, 2, (uint64_t)H * 128, (uint64_t)T, (uint64_t)ldq * 2, 64, 64))) return rc;
  if ((rc = make_tmap_2d(&tk, k, 2, (uint64_t)H * 128, (uint64_t)T, (uint64_t)ldk * 2, 64, 64))) return rc;
  if ((rc = make_tmap_2d(&tv, v, 2, (uint64_t)H * 128, (uint64_t)T, (uint64_t)ldv * 2, 64, 64))) return rc;
  if ((rc = make_tmap_2d(&tdo, dout, 2, (uint64_t)H * 128, (uint64_t)T, (uint64_t)lddo * 2, 64, 64))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    NV_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::DYN_BYTES));
    NV_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::DYN_BYTES));
    attr_set = true;
  }
  if (o != nullptr) {     // o == nullptr: dvec already holds D (written by the o_proj dgrad epilogue, nv_gemm_attnd_bf16)
    const int64_t threads = (int64_t)T * H * 32;
    attn_bwd_prep_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(o), ldo, reinterpret_cast<const __nv_bfloat16*>(dout), lddo, dvec, T, H);
    NV_LAUNCH_CHECK();
  }
  const __nv_bfloat16* c = reinterpret_cast<const __nv_bfloat16*>(cos_t);
  const __nv_bfloat16* s = reinterpret_cast<const __nv_bfloat16*>(sin_t);
  dim3 grid(total_blocks, H);
  attn_bwd_dkv_kernel<<<grid, AB_THREADS, DkvSmem::DYN_BYTES, stream>>>(
      tq, tk, tv, tdo, lse, dvec, reinterpret_cast<__nv_bfloat16*>(dk), lddk, reinterpret_cast<__nv_bfloat16*>(dv), lddv,
      cu_seqlens, B, T, scale, rope_pos, c, s);
  NV_LAUNCH_CHECK();
  attn_bwd_dq_kernel<<<grid, AB_THREADS, DqSmem::DYN_BYTES, stream>>>(
      tq, tdo, tk, tv, lse, dvec, reinterpret_cast<__nv_bfloat16*>(dq), lddq, cu_seqlens, B, T, scale, rope_pos, c, s);
  NV_LAUNCH_CHECK();
  return NV_OK;
}
