// navillm_b200 — fused gradient-norm clip + AdamW over the flat parameter / gradient buffers ("next" row n3 of
// SURVEY.md §8f).  Replaces train.py:86-89 of the reference:
//     torch.nn.utils.clip_grad_norm_(model.parameters(), 40.); optimizer.step()   (AdamW, tools/optims.py:43)
// which walks ~700 tensors with several elementwise passes each.  Here: one sum-of-squares pass per flat
// buffer, a one-thread kernel that turns the partial sums into the clip coefficient ON THE DEVICE (no host
// sync), and one update pass per buffer that reads p,g,m,v and writes p,m,v once (7 x bytes(p) of HBM traffic).
// Math in fp32; parameters and moments are stored in the parameter dtype (bf16 for the LM, like torch AdamW on
// bf16 parameters; fp32 for the encoder).
#include "nv_common.cuh"
#include "nv_host.h"

namespace nv {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ g, int64_t n, float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = ldf(g + i);
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}

// state[0] = total grad norm, state[1] = clip coefficient = min(1, max_norm / (norm + 1e-6))  (torch semantics)
__global__ void clip_coef_kernel(const float* __restrict__ partial, int n_partial, float max_norm, float* __restrict__ state) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t = 0.0;
  for (int i = 0; i < n_partial; ++i) t += (double)partial[i];
  const float norm = (float)sqrt(t);
  state[0] = norm;
  const float c = max_norm / (norm + 1e-6f);
  state[1] = c < 1.f ? c : 1.f;
}

template <typename T>
__global__ void __launch_bounds__(256) adamw_kernel(T* __restrict__ p, T* __restrict__ g, T* __restrict__ m,
                                                    T* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                    float wd, float bc1, float bc2_sqrt, const float* __restrict__ state,
                                                    int write_clipped_grad) {
  const float coef = state ? state[1] : 1.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = ldf(g + i) * coef;
    float pi = ldf(p + i);
    float mi = ldf(m + i), vi = ldf(v + i);
    pi *= (1.f - lr * wd);                       // decoupled weight decay
    mi = mi + (1.f - b1) * (gi - mi);            // lerp, like torch
    vi = b2 * vi + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    stf(p + i, pi); stf(m + i, mi); stf(v + i, vi);
    if (write_clipped_grad) stf(g + i, gi);      // leave the clipped gradient behind, like clip_grad_norm_
  }
}

}  // namespace nv

using namespace nv;
#define S_(x) reinterpret_cast<cudaStream_t>(x)

extern "C" {

int nv_optim_partials(void) { return sm_count() * 8; }

// partial: fp32 [nv_optim_partials()] slice for this buffer (caller concatenates the slices of all buffers)
int nv_grad_sumsq(const void* g, int64_t n, int is_bf16, float* partial, void* stream) {
  const int blocks = sm_count() * 8;
  if (is_bf16) sumsq_kernel<__nv_bfloat16><<<blocks, 256, 0, S_(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(g), n, partial);
  else sumsq_kernel<float><<<blocks, 256, 0, S_(stream)>>>(reinterpret_cast<const float*>(g), n, partial);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_clip_coef(const float* partial, int n_partial, float max_norm, float* state, void* stream) {
  clip_coef_kernel<<<1, 32, 0, S_(stream)>>>(partial, n_partial, max_norm, state);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_adamw_flat(void* p, void* g, void* m, void* v, int64_t n, int is_bf16, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int step, const float* clip_state, int write_clipped_grad, void* stream) {
  NV_REQUIRE(step >= 1, "nv_adamw_flat: step must be >= 1");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  const int blocks = sm_count() * 8;
  if (is_bf16)
    adamw_kernel<__nv_bfloat16><<<blocks, 256, 0, S_(stream)>>>(
        reinterpret_cast<__nv_bfloat16*>(p), reinterpret_cast<__nv_bfloat16*>(g), reinterpret_cast<__nv_bfloat16*>(m),
        reinterpret_cast<__nv_bfloat16*>(v), n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, clip_state, write_clipped_grad);
  else
    adamw_kernel<float><<<blocks, 256, 0, S_(stream)>>>(reinterpret_cast<float*>(p), reinterpret_cast<float*>(g),
                                                        reinterpret_cast<float*>(m), reinterpret_cast<float*>(v), n, lr, beta1,
                                                        beta2, eps, weight_decay, bc1, bc2s, clip_state, write_clipped_grad);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

}  // extern "C"
