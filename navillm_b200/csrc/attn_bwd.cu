// navillm_b200 — causal self-attention backward on tcgen05 (packed variable-length rows).
//
// Hand-written twin of attn_fwd.cu; replaces the autograd backward of HF LLaMA's eager attention
// (reference: loss.backward() call sites tasks/agents/mp3d_agent.py:750-757, tasks/agents/llava.py:38-40
// through models/modified_lm.py:112-116).
//
// With P = exp(S*scale - LSE), D_i = sum_d dO_id O_id, dS = P o (dP - D) * scale:
//     dV = P^T dO      dK = dS^T Q      dQ = dS K
// Two deterministic passes (no atomics) over the same code, selected by MODE:
//   MODE_DKDV : CTA = (128-key block jb, head); loops over query blocks i >= jb; K_jb,V_jb resident,
//               Q_i,dO_i streamed; accumulators dV,dK in TMEM.
//   MODE_DQ   : CTA = (128-query block ib, head); loops over key blocks j <= ib; Q_ib,dO_ib resident,
//               K_j,V_j streamed; accumulator dQ in TMEM.
// Per iteration: S = Q K^T and dP = dO V^T (tcgen05, fp32 in TMEM) -> one thread per query row turns
// them into P and dS (bf16, 128B-swizzled smem, [q rows x keys]) -> accumulate MMAs read them as
// K-major (dQ) or MN-major (dV/dK: P^T, dS^T) operands.
// TMEM: S [0,128) dP [128,256) acc0 [256,384) acc1 [384,512).  smem: 6 x 32 KB tiles = 192 KB.
#include "nv_common.cuh"
#include "nv_host.h"

namespace nv {

constexpr uint32_t AB_TILE = 128 * 128 * 2;
constexpr uint32_t AB_ATOM = 128 * 128;
constexpr uint32_t AB_THREADS = 320;  // TMA warp, MMA warp, 8 compute warps (2 per TMEM lane quarter: column halves)
constexpr int MODE_DKDV = 0, MODE_DQ = 1;

struct AttnBwdSmem {
  static constexpr uint32_t R0_OFF = 0;            // resident tile 0 (K | Q)
  static constexpr uint32_t R1_OFF = 1 * AB_TILE;  // resident tile 1 (V | dO)
  static constexpr uint32_t S0_OFF = 2 * AB_TILE;  // streamed tile 0 (Q_i | K_j)
  static constexpr uint32_t S1_OFF = 3 * AB_TILE;  // streamed tile 1 (dO_i | V_j)
  static constexpr uint32_t P_OFF = 4 * AB_TILE;
  static constexpr uint32_t DS_OFF = 5 * AB_TILE;
  static constexpr uint32_t BAR_OFF = 6 * AB_TILE;
  static constexpr uint32_t NUM_BARS = 5;  // res_full, ld_full, sdp_full, pds_ready, acc_done
  static constexpr uint32_t TOTAL = BAR_OFF + NUM_BARS * 8 + 16;
  static constexpr uint32_t DYN_BYTES = TOTAL + 1024;
};

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

template <int MODE>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                const float* __restrict__ lse, const float* __restrict__ Dvec, __nv_bfloat16* __restrict__ out0,
                int64_t ld0, __nv_bfloat16* __restrict__ out1, int64_t ld1, const int* __restrict__ cu_seqlens, int B,
                int T, float scale, const int* __restrict__ rope_pos, const __nv_bfloat16* __restrict__ cos_t,
                const __nv_bfloat16* __restrict__ sin_t) {
  using L = AttnBwdSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sR0 = smem + L::R0_OFF;
  uint8_t* sR1 = smem + L::R1_OFF;
  uint8_t* sS0 = smem + L::S0_OFF;
  uint8_t* sS1 = smem + L::S1_OFF;
  uint8_t* sP = smem + L::P_OFF;
  uint8_t* sDS = smem + L::DS_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* res_full = bars + 0;
  uint64_t* ld_full = bars + 1;
  uint64_t* sdp_full = bars + 2;
  uint64_t* pds_ready = bars + 3;
  uint64_t* acc_done = bars + 4;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t head = blockIdx.y;
  int seq_start = 0, seq_len = 0;
  uint32_t own = 0;
  const uint32_t flat = (MODE == MODE_DQ) ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;  // heavy blocks first
  if (!locate_block_bwd(cu_seqlens, B, flat, seq_start, seq_len, own)) return;
  const uint32_t nblk = (seq_len + 127) / 128;
  // partner-block range: DKDV -> query blocks own..nblk-1 ; DQ -> key blocks 0..own
  const uint32_t it_begin = (MODE == MODE_DKDV) ? own : 0;
  const uint32_t it_end = (MODE == MODE_DKDV) ? nblk : own + 1;
  const uint32_t n_it = it_end - it_begin;

  // Smem views by role.  Q/dO/K/V tiles are all [128 rows x 128 hd] as two 64-wide swizzle atoms.
  uint8_t* sQ = (MODE == MODE_DKDV) ? sS0 : sR0;
  uint8_t* sdO = (MODE == MODE_DKDV) ? sS1 : sR1;
  uint8_t* sK = (MODE == MODE_DKDV) ? sR0 : sS0;
  uint8_t* sV = (MODE == MODE_DKDV) ? sR1 : sS1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do);
    mbar_init(res_full, 1);
    mbar_init(ld_full, 1);
    mbar_init(sdp_full, 1);
    mbar_init(pds_ready, 256);
    mbar_init(acc_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_A0 = tmem_base + 256, tmem_A1 = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      const int32_t col = head * 128;
      const int32_t own_row = seq_start + own * 128;
      mbar_arrive_expect_tx(res_full, 2 * AB_TILE);
      if (MODE == MODE_DKDV) {
        tma_load_2d(sR0, &tmap_k, res_full, col, own_row);
        tma_load_2d(sR0 + AB_ATOM, &tmap_k, res_full, col + 64, own_row);
        tma_load_2d(sR1, &tmap_v, res_full, col, own_row);
        tma_load_2d(sR1 + AB_ATOM, &tmap_v, res_full, col + 64, own_row);
      } else {
        tma_load_2d(sR0, &tmap_q, res_full, col, own_row);
        tma_load_2d(sR0 + AB_ATOM, &tmap_q, res_full, col + 64, own_row);
        tma_load_2d(sR1, &tmap_do, res_full, col, own_row);
        tma_load_2d(sR1 + AB_ATOM, &tmap_do, res_full, col + 64, own_row);
      }
      for (uint32_t n = 0; n < n_it; ++n) {
        if (n > 0) mbar_wait(acc_done, (n - 1) & 1);  // previous iteration's MMAs finished reading the tiles
        const int32_t row = seq_start + (it_begin + n) * 128;
        mbar_arrive_expect_tx(ld_full, 2 * AB_TILE);
        const CUtensorMap* m0 = (MODE == MODE_DKDV) ? &tmap_q : &tmap_k;
        const CUtensorMap* m1 = (MODE == MODE_DKDV) ? &tmap_do : &tmap_v;
        tma_load_2d(sS0, m0, ld_full, col, row);
        tma_load_2d(sS0 + AB_ATOM, m0, ld_full, col + 64, row);
        tma_load_2d(sS1, m1, ld_full, col, row);
        tma_load_2d(sS1 + AB_ATOM, m1, ld_full, col + 64, row);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_kk = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_mm = umma_idesc_bf16(128, 128, 1, 1);
      constexpr uint32_t idesc_km = umma_idesc_bf16(128, 128, 0, 1);
      // loop-invariant operand descriptors (the issuing thread is the serial resource: keep its loop short)
      uint64_t q_k[2], k_k[2], o_k[2], v_k[2], ds_k[2];
#pragma unroll
      for (uint32_t ka = 0; ka < 2; ++ka) {
        q_k[ka] = umma_smem_desc_sw128(smem_u32(sQ + ka * AB_ATOM), 0, 1024);     // K-major over hd
        k_k[ka] = umma_smem_desc_sw128(smem_u32(sK + ka * AB_ATOM), 0, 1024);
        o_k[ka] = umma_smem_desc_sw128(smem_u32(sdO + ka * AB_ATOM), 0, 1024);
        v_k[ka] = umma_smem_desc_sw128(smem_u32(sV + ka * AB_ATOM), 0, 1024);
        ds_k[ka] = umma_smem_desc_sw128(smem_u32(sDS + ka * AB_ATOM), 0, 1024);   // K-major over keys
      }
      // MN-major views (rows = contraction index, 64-wide MN atoms AB_ATOM bytes apart); k-step = 16 rows = 2048 B
      const uint64_t p_m = umma_smem_desc_sw128(smem_u32(sP), AB_ATOM, 1024);
      const uint64_t ds_m = umma_smem_desc_sw128(smem_u32(sDS), AB_ATOM, 1024);
      const uint64_t o_m = umma_smem_desc_sw128(smem_u32(sdO), AB_ATOM, 1024);
      const uint64_t q_m = umma_smem_desc_sw128(smem_u32(sQ), AB_ATOM, 1024);
      const uint64_t k_m = umma_smem_desc_sw128(smem_u32(sK), AB_ATOM, 1024);
      mbar_wait(res_full, 0);
      for (uint32_t n = 0; n < n_it; ++n) {
        mbar_wait(ld_full, n & 1);
        tc_fence_after();
        // S = Q K^T ; dP = dO V^T   (both operands K-major over hd: 2 atoms x 4 k-steps)
#pragma unroll
        for (uint32_t ka = 0; ka < 2; ++ka)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_ss(tmem_S, q_k[ka] + ks * 2, k_k[ka] + ks * 2, idesc_kk, (ka | ks) ? 1u : 0u);
#pragma unroll
        for (uint32_t ka = 0; ka < 2; ++ka)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_ss(tmem_dP, o_k[ka] + ks * 2, v_k[ka] + ks * 2, idesc_kk, (ka | ks) ? 1u : 0u);
        umma_commit(sdp_full);
        mbar_wait(pds_ready, n & 1);
        tc_fence_after();
        if (MODE == MODE_DKDV) {
          // dV += P^T dO ; dK += dS^T Q : contraction over the 128 query rows (8 k-steps of 16 rows)
#pragma unroll
          for (uint32_t ks = 0; ks < 8; ++ks) umma_f16_ss(tmem_A0, p_m + ks * 128, o_m + ks * 128, idesc_mm, (n | ks) ? 1u : 0u);
#pragma unroll
          for (uint32_t ks = 0; ks < 8; ++ks) umma_f16_ss(tmem_A1, ds_m + ks * 128, q_m + ks * 128, idesc_mm, (n | ks) ? 1u : 0u);
        } else {
          // dQ += dS K : A = dS K-major over keys (2 atoms x 4 k-steps), B = K MN-major (hd contiguous)
#pragma unroll
          for (uint32_t ka = 0; ka < 2; ++ka)
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks)
              umma_f16_ss(tmem_A0, ds_k[ka] + ks * 2, k_m + ((ka * 64 + ks * 16) * 128 >> 4), idesc_km, (n | ka | ks) ? 1u : 0u);
        }
        umma_commit(acc_done);
      }
    }
  } else {
    // Two warps share each TMEM lane quarter (a warp may touch lanes 32*(warp%4)..+31 only) and split the 128
    // key columns in halves: no row reductions are needed in the backward (LSE and D are precomputed), so the
    // element-wise phase parallelises over columns and every SMSP holds two compute warps to hide latencies.
    const uint32_t quarter = warp & 3;
    const uint32_t r = quarter * 32 + lane;
    const uint32_t lane_off = (quarter * 32) << 16;
    // the warp owns the 32-column chunks c0 and c0 + 64: a rotate-half RoPE pair (i, i + 64) stays in one thread,
    // so the inverse rotation of dQ / dK can be applied in the epilogue (models/... HF apply_rotary_pos_emb backward)
    const uint32_t c0 = ((warp - 2) >> 2) * 32;
    const float sl2 = scale * 1.4426950408889634f;
    for (uint32_t n = 0; n < n_it; ++n) {
      const uint32_t qb = (MODE == MODE_DKDV) ? (it_begin + n) : own;  // query block of this iteration
      const uint32_t kb = (MODE == MODE_DKDV) ? own : (it_begin + n);  // key block
      const uint32_t qi = qb * 128 + r;
      const bool row_valid = qi < (uint32_t)seq_len;
      const int64_t t = (int64_t)seq_start + qi;
      const float lse_r = row_valid ? lse[(int64_t)head * T + t] * 1.4426950408889634f : 0.f;
      const float d_r = row_valid ? Dvec[(int64_t)head * T + t] : 0.f;
      const bool diag = (qb == kb);
      mbar_wait(sdp_full, n & 1);
      tc_fence_after();
#pragma unroll 1
      for (uint32_t c = c0; c < 128; c += 64) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32b_x32(tmem_S + lane_off + c, sv);
        tmem_ld_32x32b_x32(tmem_dP + lane_off + c, dv);
        tmem_ld_wait();
        uint32_t pp[16], dd[16];
#pragma unroll
        for (uint32_t i = 0; i < 32; i += 2) {
          float p0 = exp2f(__uint_as_float(sv[i]) * sl2 - lse_r);
          float p1 = exp2f(__uint_as_float(sv[i + 1]) * sl2 - lse_r);
          if (!row_valid || (diag && c + i > r)) p0 = 0.f;
          if (!row_valid || (diag && c + i + 1 > r)) p1 = 0.f;
          const float g0 = p0 * (__uint_as_float(dv[i]) - d_r) * scale;
          const float g1 = p1 * (__uint_as_float(dv[i + 1]) - d_r) * scale;
          pp[i >> 1] = pack_bf16x2(p0, p1);
          dd[i >> 1] = pack_bf16x2(g0, g1);
        }
        const uint32_t atom_off = (c >> 6) * AB_ATOM;
        const uint32_t chunk0 = (c & 63) >> 3;
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
          const uint32_t off = atom_off + sw128_offset(r, chunk0 + q);
          if (MODE == MODE_DKDV)
            *reinterpret_cast<uint4*>(sP + off) = make_uint4(pp[q * 4], pp[q * 4 + 1], pp[q * 4 + 2], pp[q * 4 + 3]);
          *reinterpret_cast<uint4*>(sDS + off) = make_uint4(dd[q * 4], dd[q * 4 + 1], dd[q * 4 + 2], dd[q * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_ready);
    }
    // ---- epilogue: accumulators -> bf16 -> HBM (row = TMEM lane = owned block row) ----
    mbar_wait(acc_done, (n_it - 1) & 1);
    tc_fence_after();
    const uint32_t oi = own * 128 + r;
    const bool valid = oi < (uint32_t)seq_len;
    const int64_t t = (int64_t)seq_start + oi;
    constexpr int NOUT = (MODE == MODE_DKDV) ? 2 : 1;
#pragma unroll
    for (int which = 0; which < NOUT; ++which) {
      // DKDV: acc0 = dV -> out1 (v grads), acc1 = dK -> out0 (k grads).  DQ: acc0 = dQ -> out0.
      const uint32_t tm = (which == 0) ? tmem_A0 : tmem_A1;
      __nv_bfloat16* dst;
      if (MODE == MODE_DKDV) dst = (which == 0) ? (out1 + t * ld1 + head * 128) : (out0 + t * ld0 + head * 128);
      else dst = out0 + t * ld0 + head * 128;
      const bool rope = rope_pos != nullptr && !(MODE == MODE_DKDV && which == 0);   // dQ and dK, never dV
      uint32_t a[32], b[32];
      tmem_ld_32x32b_x32(tm + lane_off + c0, a);
      tmem_ld_32x32b_x32(tm + lane_off + c0 + 64, b);
      tmem_ld_wait();
      if (valid) {
        const int p = rope ? rope_pos[t] : 0;
        const __nv_bfloat16* cs = cos_t + (int64_t)p * 128;
        const __nv_bfloat16* sn = sin_t + (int64_t)p * 128;
#pragma unroll
        for (uint32_t i = 0; i < 32; i += 8) {
          uint4 o1, o2;
          uint32_t* p1 = &o1.x; uint32_t* p2 = &o2.x;
          uint4 c1 = make_uint4(0, 0, 0, 0), s1 = c1, c2 = c1, s2 = c1;
          if (rope) {
            c1 = *reinterpret_cast<const uint4*>(cs + c0 + i);      s1 = *reinterpret_cast<const uint4*>(sn + c0 + i);
            c2 = *reinterpret_cast<const uint4*>(cs + 64 + c0 + i); s2 = *reinterpret_cast<const uint4*>(sn + 64 + c0 + i);
          }
          const uint32_t* pc1 = &c1.x; const uint32_t* ps1 = &s1.x; const uint32_t* pc2 = &c2.x; const uint32_t* ps2 = &s2.x;
#pragma unroll
          for (uint32_t q = 0; q < 4; ++q) {
            const uint32_t xa = pack_bf16x2(__uint_as_float(a[i + 2 * q]), __uint_as_float(a[i + 2 * q + 1]));
            const uint32_t xb = pack_bf16x2(__uint_as_float(b[i + 2 * q]), __uint_as_float(b[i + 2 * q + 1]));
            if (rope) {   // rotation by -theta on the bf16 gradient, same rounding points as rope_kernel(sign = -1)
              const float y1l = bf16_round(bf16_lo(xa) * bf16_lo(pc1[q])) + bf16_round(bf16_lo(xb) * bf16_lo(ps1[q]));
              const float y1h = bf16_round(bf16_hi(xa) * bf16_hi(pc1[q])) + bf16_round(bf16_hi(xb) * bf16_hi(ps1[q]));
              const float y2l = bf16_round(bf16_lo(xb) * bf16_lo(pc2[q])) + bf16_round(-bf16_lo(xa) * bf16_lo(ps2[q]));
              const float y2h = bf16_round(bf16_hi(xb) * bf16_hi(pc2[q])) + bf16_round(-bf16_hi(xa) * bf16_hi(ps2[q]));
              p1[q] = pack_bf16x2(y1l, y1h);
              p2[q] = pack_bf16x2(y2l, y2h);
            } else {
              p1[q] = xa;
              p2[q] = xb;
            }
          }
          *reinterpret_cast<uint4*>(dst + c0 + i) = o1;
          *reinterpret_cast<uint4*>(dst + c0 + 64 + i) = o2;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace nv

// Inputs: q,k,v (post-RoPE) and o, do as bf16 [T, H*128]-column views; lse [H,T] from the forward.
// dvec: fp32 workspace [H, T].  Outputs dq, dk, dv: bf16 views with their own leading dimensions.
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
  if ((rc = make_tmap_2d(&tq, q, 2, (uint64_t)H * 128, (uint64_t)T, (uint64_t)ldq * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&tk, k, 2, (uint64_t)H * 128, (uint64_t)T, (uint64_t)ldk * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&tv, v, 2, (uint64_t)H * 128, (uint64_t)T, (uint64_t)ldv * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&tdo, dout, 2, (uint64_t)H * 128, (uint64_t)T, (uint64_t)lddo * 2, 64, 128))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    NV_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<MODE_DKDV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 AttnBwdSmem::DYN_BYTES));
    NV_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<MODE_DQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 AttnBwdSmem::DYN_BYTES));
    attr_set = true;
  }
  {
    const int64_t threads = (int64_t)T * H * 32;
    attn_bwd_prep_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(o), ldo, reinterpret_cast<const __nv_bfloat16*>(dout), lddo, dvec, T, H);
    NV_LAUNCH_CHECK();
  }
  dim3 grid(total_blocks, H);
  attn_bwd_kernel<MODE_DKDV><<<grid, AB_THREADS, AttnBwdSmem::DYN_BYTES, stream>>>(
      tq, tk, tv, tdo, lse, dvec, reinterpret_cast<__nv_bfloat16*>(dk), lddk, reinterpret_cast<__nv_bfloat16*>(dv), lddv,
      cu_seqlens, B, T, scale, rope_pos, reinterpret_cast<const __nv_bfloat16*>(cos_t),
      reinterpret_cast<const __nv_bfloat16*>(sin_t));
  NV_LAUNCH_CHECK();
  attn_bwd_kernel<MODE_DQ><<<grid, AB_THREADS, AttnBwdSmem::DYN_BYTES, stream>>>(
      tq, tk, tv, tdo, lse, dvec, reinterpret_cast<__nv_bfloat16*>(dq), lddq, nullptr, 0, cu_seqlens, B, T, scale, rope_pos,
      reinterpret_cast<const __nv_bfloat16*>(cos_t), reinterpret_cast<const __nv_bfloat16*>(sin_t));
  NV_LAUNCH_CHECK();
  return NV_OK;
}
