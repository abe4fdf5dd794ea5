// navillm_b200 — one decoder layer of the INFERENCE forward as a single C-ABI call.
//
// Host-side composite, no new kernels: nv_llama_layer_infer launches the ten kernels of a LLaMA decoder layer
// (reference: transformers LlamaDecoderLayer reached through models/modified_lm.py:112-116) in order, on
// the caller's stream, with every intermediate in a caller-provided workspace.  Why it exists: small packed batches
// (one navigation step at batch 1, evaluation rollouts with cross-step prefix-KV reuse) are HOST-bound
// when each kernel is its own ctypes call from Python (~20 us per call against a few microseconds of GPU work: 328 calls
// = 10 ms per forward at T = 192); one call per layer leaves the GPU as the limit.  Training-size batches do not
// need it (their step is GPU-bound) and keep the per-kernel path, which also saves the activations for the backward.
#include <stdint.h>

#include "navillm_b200.h"
#include "nv_host.h"

namespace {

inline int64_t al(int64_t x) { return (x + 255) & ~int64_t(255); }

struct Carve {
  uint8_t* p;
  int64_t left;
  void* take(int64_t bytes) {
    bytes = al(bytes);
    if (bytes > left) return nullptr;
    void* r = p;
    p += bytes;
    left -= bytes;
    return r;
  }
};

}  // namespace

extern "C" int nv_layer_args_size(void) { return (int)sizeof(nv_layer_args); }

extern "C" int64_t nv_llama_layer_ws_bytes(int T, int R, int D, int F) {
  const int64_t Tm = R > 0 ? R : T;
  int64_t b = al((int64_t)T * D * 2) + al((int64_t)T * 4) + al((int64_t)T * 3 * D * 2) + al((int64_t)T * D * 2);
  if (R > 0) b += 2 * al((int64_t)R * D * 2);
  b += 2 * al(Tm * D * 2) + al(Tm * 4) + al(Tm * 2 * F * 2) + al(Tm * F * 2);
  return b + 256;
}

extern "C" int nv_llama_layer_infer(const nv_layer_args* a, void* stream) {
  using namespace nv;
  NV_REQUIRE(a && a->x && a->y && a->ws, "nv_llama_layer_infer: null argument");
  const int T = a->T, D = a->D, F = a->F, H = a->H, R = a->out_rows ? a->R : 0;
  NV_REQUIRE(T > 0 && D == H * 128 && F > 0, "nv_llama_layer_infer: needs head_dim 128 (D=%d H=%d)", D, H);
  NV_REQUIRE(a->ws_bytes >= nv_llama_layer_ws_bytes(T, R, D, F), "nv_llama_layer_infer: workspace too small");
  uintptr_t base = (reinterpret_cast<uintptr_t>(a->ws) + 255) & ~uintptr_t(255);
  Carve c{reinterpret_cast<uint8_t*>(base), a->ws_bytes - (int64_t)(base - reinterpret_cast<uintptr_t>(a->ws))};
  const int Tm = R > 0 ? R : T;
  void* xn = c.take((int64_t)T * D * 2);
  float* rstd = static_cast<float*>(c.take((int64_t)T * 4));
  uint8_t* qkv = static_cast<uint8_t*>(c.take((int64_t)T * 3 * D * 2));
  void* ao = c.take((int64_t)T * D * 2);
  void* aor = R > 0 ? c.take((int64_t)R * D * 2) : ao;
  void* xr = R > 0 ? c.take((int64_t)R * D * 2) : const_cast<void*>(a->x);
  void* xm = c.take((int64_t)Tm * D * 2);
  void* xn2 = c.take((int64_t)Tm * D * 2);
  float* rstd2 = static_cast<float*>(c.take((int64_t)Tm * 4));
  void* gu = c.take((int64_t)Tm * 2 * F * 2);
  void* h = c.take((int64_t)Tm * F * 2);
  NV_REQUIRE(h != nullptr, "nv_llama_layer_infer: workspace carve failed");
  int rc;
#define STEP(call) do { rc = (call); if (rc != NV_OK) return rc; } while (0)
  // ---- attention block:  xm = x + o_proj(attn(rope(qkv(rmsnorm1(x))))) ----
  STEP(nv_rmsnorm_fwd(a->x, D, a->ln1, xn, D, rstd, T, D, a->eps, stream));
  STEP(nv_gemm_bf16(xn, D, 0, a->wqkv, D, 0, qkv, 3 * (int64_t)D, nullptr, 0, T, 3 * D, D, 0u, 0, stream));
  STEP(nv_rope_inplace(qkv, 3 * (int64_t)D, a->pos, a->cos_t, a->sin_t, T, 2 * H, 128, 0, stream));
  if (a->kv_mode == 2) {            // new rows appended to a cache that already holds a prefix; attention over the cache
    STEP(nv_kv_store_suffix(qkv, 3 * (int64_t)D, a->cu_seqlens, a->cached, a->kcache, a->vcache, a->B, T, a->Smax, D, stream));
    STEP(nv_attn_fwd_kv(qkv, 3 * (int64_t)D, a->kcache, D, a->vcache, D, ao, D, nullptr, a->cu_seqlens, a->kv_start, a->kv_len, a->B,
                        T, a->Tkv, H, 128, a->total_qblocks, a->scale, stream));
  } else {
    if (a->kv_mode == 1)            // prefill of generate(): post-RoPE K, V also go to the cache
      STEP(nv_kv_store_prefill(qkv, 3 * (int64_t)D, a->cu_seqlens, a->kcache, a->vcache, a->B, T, a->Smax, D, stream));
    STEP(nv_attn_fwd(qkv, 3 * (int64_t)D, qkv + (int64_t)D * 2, 3 * (int64_t)D, qkv + (int64_t)D * 4, 3 * (int64_t)D, ao, D, nullptr,
                     a->cu_seqlens, a->B, T, H, 128, a->total_qblocks, a->scale, stream));
  }
  if (R > 0) {                      // last layer: only the requested rows continue (their K, V came from all rows)
    STEP(nv_gather_rows(ao, D, a->out_rows, aor, D, R, D, stream));
    STEP(nv_gather_rows(a->x, D, a->out_rows, xr, D, R, D, stream));
  }
  STEP(nv_gemm_bf16(aor, D, 0, a->wo, D, 0, xm, D, xr, D, Tm, D, D, 1u /* + addend */, 0, stream));
  // ---- MLP:  y = xm + down(silu(gate(xn2)) * up(xn2)) ----
  STEP(nv_rmsnorm_fwd(xm, D, a->ln2, xn2, D, rstd2, Tm, D, a->eps, stream));
  STEP(nv_gemm_bf16(xn2, D, 0, a->wgu, D, 0, gu, 2 * (int64_t)F, nullptr, 0, Tm, 2 * F, D, 0u, 0, stream));
  STEP(nv_swiglu_fwd(gu, 2 * (int64_t)F, h, F, Tm, F, stream));
  STEP(nv_gemm_bf16(h, F, 0, a->wd, F, 0, a->y, D, xm, D, Tm, D, F, 1u, 0, stream));
#undef STEP
  return NV_OK;
}
