// navillm_b200 — HBM-bound element/row-wise kernels of the modified-LLaMA path (forward + backward).
//
// These replace the ~20 eager elementwise launches per decoder layer the reference runs through HF
// LLaMA: token-embedding gather + visual-token scatter-add
// (models/modified_lm.py:100-110), LlamaRMSNorm, rotary embedding (rotate-half), SwiGLU, and the
// <cls_1> action head gather/GEMV (models/nav_model.py:234-242).  Rounding points follow the
// reference's bf16 eager arithmetic (each op result rounded to bf16) so that outputs agree to bf16
// tolerance; internal math is fp32.
//
// All kernels use 128-bit global accesses; rows are D contiguous bf16 with D % 8 == 0.
#include "nv_common.cuh"
#include "nv_host.h"

namespace nv {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the whole CTA (blockDim.x multiple of 32, <= 1024); result broadcast to all threads.
__device__ __forceinline__ float block_sum(float v, float* red /* >= 33 floats */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();  // protect `red` reuse across consecutive calls
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.f;
  t = warp_sum(t);
  return t;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------------------------------------
// RMSNorm forward:  y = bf16( w * bf16(x * rsqrt(mean(x^2) + eps)) )      (HF LlamaRMSNorm)
// One CTA per row; the row stays in registers between the two passes (D <= 8 * 4 * blockDim).
// ------------------------------------------------------------------------------------------------
constexpr int RMS_THREADS = 128;
constexpr int RMS_MAX_VEC = 4;  // uint4 per thread kept in registers -> D <= 4096

__global__ void __launch_bounds__(RMS_THREADS) rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                                  const __nv_bfloat16* __restrict__ w,
                                                                  __nv_bfloat16* __restrict__ y, int64_t ldy,
                                                                  float* __restrict__ rstd_out, int D, float eps) {
  __shared__ float red[33];
  griddep_launch();
  griddep_wait();
  const int row = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (int64_t)row * ldx);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  uint4* yr = reinterpret_cast<uint4*>(y + (int64_t)row * ldy);
  const int nvec = D >> 3;
  uint4 xv[RMS_MAX_VEC];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < RMS_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * RMS_THREADS;
    if (v < nvec) {
      xv[i] = xr[v];
      float f[8];
      unpack8(xv[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / (float)D + eps);
  if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
#pragma unroll
  for (int i = 0; i < RMS_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * RMS_THREADS;
    if (v < nvec) {
      float f[8], g[8];
      unpack8(xv[i], f);
      unpack8(wr[v], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = g[j] * bf16_round(f[j] * rstd);
      yr[v] = pack8(f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// RMSNorm backward.  With xh = x*rstd, g = dy*w:
//   dx = rstd * (g - xh * mean(g*xh)) [+ dres]        dw[j] = sum_t dy[t,j] * xh[t,j]
// Persistent CTAs stride over rows and keep their dw partial in registers; partials [grid, D] fp32 go
// to a caller workspace and are folded into the bf16 weight gradient by rmsnorm_dw_reduce_kernel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RMS_THREADS) rmsnorm_bwd_kernel(
    const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w,
    const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ dy, int64_t lddy,
    const __nv_bfloat16* __restrict__ dres, int64_t lddres, __nv_bfloat16* __restrict__ dx, int64_t lddx,
    float* __restrict__ dw_partial, int T, int D) {
  __shared__ float red[33];
  const int nvec = D >> 3;
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  float dwacc[RMS_MAX_VEC][8];
#pragma unroll
  for (int i = 0; i < RMS_MAX_VEC; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dwacc[i][j] = 0.f;

  // Software pipeline: the x / dy / dres vectors of the NEXT row are requested before the block-wide reduction of the
  // current row, so loads stay in flight across the two __syncthreads of block_sum (without this the kernel ran at 45 %
  // of the HBM peak, tools/rowops_bench.py).
  uint4 xv[RMS_MAX_VEC], gv[RMS_MAX_VEC], rv[RMS_MAX_VEC];
  auto fetch = [&](int row, uint4 (&xo)[RMS_MAX_VEC], uint4 (&go)[RMS_MAX_VEC], uint4 (&ro)[RMS_MAX_VEC]) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (int64_t)row * ldx);
    const uint4* dyr = reinterpret_cast<const uint4*>(dy + (int64_t)row * lddy);
    const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + (int64_t)row * lddres) : nullptr;
#pragma unroll
    for (int i = 0; i < RMS_MAX_VEC; ++i) {
      const int v = threadIdx.x + i * RMS_THREADS;
      if (v < nvec) {
        xo[i] = xr[v];
        go[i] = dyr[v];
        if (rr) ro[i] = rr[v];
      }
    }
  };
  int row = blockIdx.x;
  if (row < T) fetch(row, xv, gv, rv);
  for (; row < T; row += gridDim.x) {
    const int nrow = row + gridDim.x;
    uint4 nx[RMS_MAX_VEC], ng[RMS_MAX_VEC], nr[RMS_MAX_VEC];
    if (nrow < T) fetch(nrow, nx, ng, nr);
    const float rs = rstd[row];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < RMS_MAX_VEC; ++i) {
      const int v = threadIdx.x + i * RMS_THREADS;
      if (v < nvec) {
        float xf[8], df[8], wf[8];
        unpack8(xv[i], xf); unpack8(gv[i], df); unpack8(wr[v], wf);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = xf[j] * rs;
          dwacc[i][j] += df[j] * xh;
          dot += df[j] * wf[j] * xh;
        }
      }
    }
    dot = block_sum(dot, red) / (float)D;
    uint4* dxr = reinterpret_cast<uint4*>(dx + (int64_t)row * lddx);
#pragma unroll
    for (int i = 0; i < RMS_MAX_VEC; ++i) {
      const int v = threadIdx.x + i * RMS_THREADS;
      if (v < nvec) {
        float xf[8], df[8], wf[8], o[8];
        unpack8(xv[i], xf); unpack8(gv[i], df); unpack8(wr[v], wf);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rs * (df[j] * wf[j] - xf[j] * rs * dot);
        if (dres) {
          float rf[8];
          unpack8(rv[i], rf);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += rf[j];
        }
        dxr[v] = pack8(o);
      }
    }
#pragma unroll
    for (int i = 0; i < RMS_MAX_VEC; ++i) { xv[i] = nx[i]; gv[i] = ng[i]; rv[i] = nr[i]; }
  }
  float* out = dw_partial + (int64_t)blockIdx.x * D;
#pragma unroll
  for (int i = 0; i < RMS_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * RMS_THREADS;
    if (v < nvec) {
      *reinterpret_cast<float4*>(out + v * 8) = make_float4(dwacc[i][0], dwacc[i][1], dwacc[i][2], dwacc[i][3]);
      *reinterpret_cast<float4*>(out + v * 8 + 4) = make_float4(dwacc[i][4], dwacc[i][5], dwacc[i][6], dwacc[i][7]);
    }
  }
}

// dst[j] = bf16(dst[j] + sum_p partial[p, j])   (accumulate = 1)   or   bf16(sum)   (accumulate = 0)
// Block = 32 columns x 8 row groups (a 16-block launch with one thread per column left the ~10 MB of partials
// to 4096 threads: 41 us); fixed summation order, so the result is deterministic.
__global__ void __launch_bounds__(256) colsum_accum_bf16_kernel(const float* __restrict__ partial, int P, int D,
                                                                __nv_bfloat16* __restrict__ dst, int accumulate) {
  __shared__ float red[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (j < D)
    for (int p = threadIdx.y; p < P; p += 8) s += partial[(int64_t)p * D + j];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < D) {
#pragma unroll
    for (int y = 1; y < 8; ++y) s += red[y][threadIdx.x];
    if (accumulate) s += __bfloat162float(dst[j]);
    dst[j] = __float2bfloat16_rn(s);
  }
}

// ------------------------------------------------------------------------------------------------
// Rotary embedding, rotate-half form, in place on the q and k column blocks of the fused qkv buffer.
//   y = bf16( bf16(x*cos) + bf16(rot(x)*sin) ),  cos/sin tables [max_pos, hd] bf16 built by the host
//   exactly like HF (fp32 cos/sin of pos*inv_freq, cast to bf16).  sign = -1 gives the backward.
// One thread handles one 8-element chunk i (<hd/2) and its partner chunk i + hd/2 of one head.
// ------------------------------------------------------------------------------------------------
__global__ void rope_kernel(__nv_bfloat16* __restrict__ qkv, int64_t ld, const int* __restrict__ pos,
                            const __nv_bfloat16* __restrict__ cos_t, const __nv_bfloat16* __restrict__ sin_t, int T,
                            int n_heads, int hd, float sign) {
  const int chunks = hd >> 4;  // 8-element chunks in half a head
  const int64_t total = (int64_t)T * n_heads * chunks;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = idx % chunks;
    const int h = (idx / chunks) % n_heads;
    const int t = idx / ((int64_t)chunks * n_heads);
    __nv_bfloat16* base = qkv + (int64_t)t * ld + h * hd;
    const int p = pos[t];
    const uint4 c1 = *reinterpret_cast<const uint4*>(cos_t + (int64_t)p * hd + c * 8);
    const uint4 s1 = *reinterpret_cast<const uint4*>(sin_t + (int64_t)p * hd + c * 8);
    const uint4 c2 = *reinterpret_cast<const uint4*>(cos_t + (int64_t)p * hd + hd / 2 + c * 8);
    const uint4 s2 = *reinterpret_cast<const uint4*>(sin_t + (int64_t)p * hd + hd / 2 + c * 8);
    uint4* p1 = reinterpret_cast<uint4*>(base + c * 8);
    uint4* p2 = reinterpret_cast<uint4*>(base + hd / 2 + c * 8);
    float x1[8], x2[8], cf1[8], sf1[8], cf2[8], sf2[8], y1[8], y2[8];
    unpack8(*p1, x1); unpack8(*p2, x2);
    unpack8(c1, cf1); unpack8(s1, sf1); unpack8(c2, cf2); unpack8(s2, sf2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      y1[j] = bf16_round(x1[j] * cf1[j]) + bf16_round(-x2[j] * sf1[j] * sign);
      y2[j] = bf16_round(x2[j] * cf2[j]) + bf16_round(x1[j] * sf2[j] * sign);
    }
    *p1 = pack8(y1);
    *p2 = pack8(y2);
  }
}

// ------------------------------------------------------------------------------------------------
// SwiGLU on the fused gate|up buffer gu[T, 2F] (gate = cols [0,F), up = cols [F,2F)).
//   fwd: h = bf16( bf16(silu(g)) * u )        bwd: dg = dh*u*silu'(g), du = dh*silu(g)
// ------------------------------------------------------------------------------------------------
__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, int64_t ldgu, __nv_bfloat16* __restrict__ h,
                                  int64_t ldh, int T, int F) {
  const int vecs = F >> 3;
  const int64_t total = (int64_t)T * vecs;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = idx % vecs;
    const int64_t t = idx / vecs;
    float g[8], u[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(gu + t * ldgu + v * 8), g);
    unpack8(*reinterpret_cast<const uint4*>(gu + t * ldgu + F + v * 8), u);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = bf16_round(g[j] / (1.f + __expf(-g[j]))) * u[j];
    *reinterpret_cast<uint4*>(h + t * ldh + v * 8) = pack8(o);
  }
}

// x[r, c] = bf16(x[r, c] * s[0]) in place, s a device scalar: folds the upstream gradient of a scalar loss into the stored
// dlogits without a host read of it (autograd hands CrossEntropyLoss's backward the same factor, models/modified_lm.py:126-137).
__global__ void scale_bf16_kernel(__nv_bfloat16* __restrict__ x, int64_t ld, int rows, int cols, const float* __restrict__ s) {
  const float f = s[0];
  const int pairs = cols >> 1;
  const int64_t total = (int64_t)rows * pairs;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = idx % pairs;
    const int64_t r = idx / pairs;
    uint32_t* p = reinterpret_cast<uint32_t*>(x + r * ld) + c;            // ld even: 4-byte aligned pairs
    const uint32_t v = *p;
    *p = pack_bf16x2(bf16_lo(v) * f, bf16_hi(v) * f);
  }
  if (cols & 1) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
      __nv_bfloat16* q = x + r * ld + (cols - 1);
      *q = __float2bfloat16_rn(__bfloat162float(*q) * f);
    }
  }
}

__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, int64_t ldgu,
                                  const __nv_bfloat16* __restrict__ dh, int64_t lddh, __nv_bfloat16* __restrict__ dgu,
                                  int64_t lddgu, int T, int F) {
  const int vecs = F >> 3;
  const int64_t total = (int64_t)T * vecs;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = idx % vecs;
    const int64_t t = idx / vecs;
    float g[8], u[8], d[8], og[8], ou[8];
    unpack8(*reinterpret_cast<const uint4*>(gu + t * ldgu + v * 8), g);
    unpack8(*reinterpret_cast<const uint4*>(gu + t * ldgu + F + v * 8), u);
    unpack8(*reinterpret_cast<const uint4*>(dh + t * lddh + v * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = 1.f / (1.f + __expf(-g[j]));
      const float silu = g[j] * sg;
      og[j] = d[j] * u[j] * (sg * (1.f + g[j] * (1.f - sg)));
      ou[j] = d[j] * silu;
    }
    *reinterpret_cast<uint4*>(dgu + t * lddgu + v * 8) = pack8(og);
    *reinterpret_cast<uint4*>(dgu + t * lddgu + F + v * 8) = pack8(ou);
  }
}

// ------------------------------------------------------------------------------------------------
// Token embedding gather + visual-token scatter-add (reference models/modified_lm.py:100-110):
//   out[t] = bf16( float(E[ids[t]]) + vis[vis_src[t]] )  when vis_src[t] >= 0, else E[ids[t]].
// vis is fp32 (the panorama encoder / history vectors are fp32 modules).
// ------------------------------------------------------------------------------------------------
__global__ void embed_fwd_kernel(const int* __restrict__ ids, const __nv_bfloat16* __restrict__ E, int V,
                                 const int* __restrict__ vis_src, const float* __restrict__ vis,
                                 __nv_bfloat16* __restrict__ out, int T, int D) {
  griddep_launch();
  griddep_wait();
  const int vecs = D >> 3;
  const int64_t total = (int64_t)T * vecs;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = idx % vecs;
    const int t = idx / vecs;
    int id = ids[t];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    uint4 e = *reinterpret_cast<const uint4*>(E + (int64_t)id * D + v * 8);
    const int s = vis_src ? vis_src[t] : -1;
    if (s >= 0) {
      float f[8];
      unpack8(e, f);
      const float4 a = *reinterpret_cast<const float4*>(vis + (int64_t)s * D + v * 8);
      const float4 b = *reinterpret_cast<const float4*>(vis + (int64_t)s * D + v * 8 + 4);
      f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w;
      f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
      e = pack8(f);
    }
    *reinterpret_cast<uint4*>(out + (int64_t)t * D + v * 8) = e;
  }
}

// d vis[vis_src[t]] = float(dx[t])  (each visual row feeds exactly one token)
__global__ void embed_bwd_vis_kernel(const __nv_bfloat16* __restrict__ dx, const int* __restrict__ vis_src,
                                     float* __restrict__ dvis, int T, int D) {
  const int vecs = D >> 3;
  const int64_t total = (int64_t)T * vecs;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = idx % vecs;
    const int t = idx / vecs;
    const int s = vis_src[t];
    if (s < 0) continue;
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(dx + (int64_t)t * D + v * 8), f);
    *reinterpret_cast<float4*>(dvis + (int64_t)s * D + v * 8) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(dvis + (int64_t)s * D + v * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// dE[id] += sum over tokens with that id of dx[t].  `order` lists token indices sorted by id; the CTA
// at the first position of each run of equal ids owns the whole run (deterministic, no atomics).
__global__ void embed_bwd_weight_kernel(const __nv_bfloat16* __restrict__ dx, const int* __restrict__ order,
                                        const int* __restrict__ sorted_ids, __nv_bfloat16* __restrict__ dE, int T,
                                        int D) {
  const int p = blockIdx.x;
  const int id = sorted_ids[p];
  if (p > 0 && sorted_ids[p - 1] == id) return;
  const int vecs = D >> 3;
  for (int v = threadIdx.x; v < vecs; v += blockDim.x) {
    float acc[8];
    unpack8(*reinterpret_cast<const uint4*>(dE + (int64_t)id * D + v * 8), acc);
    for (int q = p; q < T && sorted_ids[q] == id; ++q) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(dx + (int64_t)order[q] * D + v * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    *reinterpret_cast<uint4*>(dE + (int64_t)id * D + v * 8) = pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// Row gather / scatter of bf16 rows (e.g. hidden states at the <cls_1> positions).
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, int64_t lds, const int* __restrict__ rows,
                                   __nv_bfloat16* __restrict__ dst, int64_t ldd, int R, int D) {
  const int vecs = D >> 3;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < (int64_t)R * vecs;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = idx % vecs, r = idx / vecs;
    *reinterpret_cast<uint4*>(dst + (int64_t)r * ldd + v * 8) =
        *reinterpret_cast<const uint4*>(src + (int64_t)rows[r] * lds + v * 8);
  }
}
__global__ void scatter_rows_kernel(const __nv_bfloat16* __restrict__ src, int64_t lds, const int* __restrict__ rows,
                                    __nv_bfloat16* __restrict__ dst, int64_t ldd, int R, int D) {
  const int vecs = D >> 3;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < (int64_t)R * vecs;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = idx % vecs, r = idx / vecs;
    *reinterpret_cast<uint4*>(dst + (int64_t)rows[r] * ldd + v * 8) =
        *reinterpret_cast<const uint4*>(src + (int64_t)r * lds + v * 8);
  }
}

// ------------------------------------------------------------------------------------------------
// Small dense head on R gathered rows (R = batch): out[r, o] = bf16( sum_k x[r,k] W[o,k] + b[o] ).
// Reference: out_head = Linear(4096, 100) in bf16 at the <cls_1> hidden state (nav_model.py:83-85,237).
// One warp per output element; weights are tiny (0.8 MB) and stay in L2.
// ------------------------------------------------------------------------------------------------
__global__ void head_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ W,
                                const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, int R, int O,
                                int D) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= R * O) return;
  const int r = gw / O, o = gw % O;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (int64_t)r * ldx);
  const uint4* wr = reinterpret_cast<const uint4*>(W + (int64_t)o * D);
  float acc = 0.f;
  for (int v = lane; v < (D >> 3); v += 32) {
    float a[8], b[8];
    unpack8(xr[v], a); unpack8(wr[v], b);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += a[j] * b[j];
  }
  acc = warp_sum(acc);
  if (lane == 0) out[(int64_t)r * O + o] = __float2bfloat16_rn(acc + (bias ? __bfloat162float(bias[o]) : 0.f));
}

// dx[r,k] = sum_o dy[r,o] W[o,k]   (bf16 out)
__global__ void head_bwd_dx_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ W,
                                   __nv_bfloat16* __restrict__ dx, int64_t lddx, int R, int O, int D) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)R * D) return;
  const int r = idx / D, k = idx % D;
  float acc = 0.f;
  for (int o = 0; o < O; ++o) acc += __bfloat162float(dy[(int64_t)r * O + o]) * __bfloat162float(W[(int64_t)o * D + k]);
  dx[(int64_t)r * lddx + k] = __float2bfloat16_rn(acc);
}
// dW[o,k] += sum_r dy[r,o] x[r,k] ;  db[o] += sum_r dy[r,o]
__global__ void head_bwd_dw_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                   int64_t ldx, __nv_bfloat16* __restrict__ dW, __nv_bfloat16* __restrict__ db, int R,
                                   int O, int D) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)O * D) return;
  const int o = idx / D, k = idx % D;
  float acc = __bfloat162float(dW[idx]);
  float bacc = 0.f;
  for (int r = 0; r < R; ++r) {
    const float g = __bfloat162float(dy[(int64_t)r * O + o]);
    acc += g * __bfloat162float(x[(int64_t)r * ldx + k]);
    bacc += g;
  }
  dW[idx] = __float2bfloat16_rn(acc);
  if (k == 0 && db) db[o] = __float2bfloat16_rn(__bfloat162float(db[o]) + bacc);
}

// ------------------------------------------------------------------------------------------------
// Masked token cross-entropy on bf16 logits [N, V] (reference models/modified_lm.py:122-137):
// the special-token columns are forced to -inf, loss = mean over rows with label != -100 of
// -log_softmax(logits)[label].  One CTA per row; writes per-row loss (0 for ignored rows) and, when
// dlogits != null, dlogits = (softmax - onehot) * grad_scale (bf16; zero for ignored rows / specials).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld,
                                                 const int* __restrict__ labels, const int* __restrict__ special,
                                                 int n_special, float* __restrict__ row_loss,
                                                 __nv_bfloat16* __restrict__ dlogits, int64_t ldd, int V,
                                                 float grad_scale) {
  __shared__ float red[33];
  __shared__ float s_bcast;
  const int row = blockIdx.x;
  const __nv_bfloat16* lr = logits + (int64_t)row * ld;
  const int label = labels[row];
  auto is_special = [&](int c) {
    for (int s = 0; s < n_special; ++s)
      if (special[s] == c) return true;
    return false;
  };
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < V; c += blockDim.x)
    if (!is_special(c)) mx = fmaxf(mx, __bfloat162float(lr[c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
    s_bcast = m;
  }
  __syncthreads();
  mx = s_bcast;
  float se = 0.f;
  for (int c = threadIdx.x; c < V; c += blockDim.x)
    if (!is_special(c)) se += __expf(__bfloat162float(lr[c]) - mx);
  se = block_sum(se, red);
  const float lse = mx + __logf(se);
  const bool active = label >= 0 && label < V;
  if (threadIdx.x == 0) row_loss[row] = active ? (lse - __bfloat162float(lr[label])) : 0.f;
  if (dlogits) {
    __nv_bfloat16* dr = dlogits + (int64_t)row * ldd;
    for (int c = threadIdx.x; c < V; c += blockDim.x) {
      float g = 0.f;
      if (active && !is_special(c)) {
        g = __expf(__bfloat162float(lr[c]) - lse);
        if (c == label) g -= 1.f;
        g *= grad_scale;
      }
      dr[c] = __float2bfloat16_rn(g);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Action-logit scatter (models/nav_model.py:234-242): out[b,g] = slot[b,g] >= 0 ? pred[b, slot[b,g]] : -inf
// and its transpose for the gradient (each prediction column feeds at most one slot).
// ------------------------------------------------------------------------------------------------
__global__ void logit_scatter_fwd_kernel(const __nv_bfloat16* __restrict__ pred, int O, const int* __restrict__ slot,
                                         __nv_bfloat16* __restrict__ out, int B, int G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * G) return;
  const int s = slot[i];
  out[i] = s >= 0 ? pred[(i / G) * O + s] : __float2bfloat16_rn(-INFINITY);
}
__global__ void logit_scatter_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const int* __restrict__ slot,
                                         __nv_bfloat16* __restrict__ dpred, int O, int B, int G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * G) return;
  const int s = slot[i];
  if (s >= 0) dpred[(i / G) * O + s] = dout[i];
}

static inline int grid_for(int64_t work, int block, int cap_mult = 8) {
  int64_t g = (work + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * cap_mult;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace nv

using namespace nv;
#define S_(x) reinterpret_cast<cudaStream_t>(x)
#define BF(x) reinterpret_cast<__nv_bfloat16*>(x)
#define CBF(x) reinterpret_cast<const __nv_bfloat16*>(x)

extern "C" {

int nv_rmsnorm_fwd(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, float* rstd, int T, int D,
                   float eps, void* stream) {
  NV_REQUIRE(T >= 0 && D > 0 && (D & 7) == 0 && D <= 8 * RMS_MAX_VEC * RMS_THREADS, "nv_rmsnorm_fwd: bad D=%d", D);
  NV_REQUIRE((ldx & 7) == 0 && (ldy & 7) == 0, "nv_rmsnorm_fwd: leading dims must be multiples of 8");
  if (T == 0) return NV_OK;
  NV_CUDA(launch_pdl(rmsnorm_fwd_kernel, dim3(T), dim3(RMS_THREADS), 0, S_(stream), CBF(x), ldx, CBF(w), BF(y), ldy, rstd, D, eps));
  return NV_OK;
}

// workspace: fp32 [nv_rmsnorm_bwd_partials(), D]
int nv_rmsnorm_bwd_partials(void) { return sm_count() * 3; }   // 155 registers x 128 threads: three CTAs per SM

int nv_rmsnorm_bwd(const void* x, int64_t ldx, const void* w, const float* rstd, const void* dy, int64_t lddy,
                   const void* dres, int64_t lddres, void* dx, int64_t lddx, void* dw, int accumulate_dw,
                   float* workspace, int T, int D, void* stream) {
  NV_REQUIRE(T > 0 && D > 0 && (D & 7) == 0 && D <= 8 * RMS_MAX_VEC * RMS_THREADS, "nv_rmsnorm_bwd: bad T=%d D=%d", T, D);
  int P = nv_rmsnorm_bwd_partials();
  if (P > T) P = T;
  rmsnorm_bwd_kernel<<<P, RMS_THREADS, 0, S_(stream)>>>(CBF(x), ldx, CBF(w), rstd, CBF(dy), lddy, CBF(dres), lddres,
                                                       BF(dx), lddx, workspace, T, D);
  NV_LAUNCH_CHECK();
  if (dw) {
    colsum_accum_bf16_kernel<<<(D + 31) / 32, dim3(32, 8), 0, S_(stream)>>>(workspace, P, D, BF(dw), accumulate_dw);
    NV_LAUNCH_CHECK();
  }
  return NV_OK;
}

int nv_rope_inplace(void* qkv, int64_t ld, const int* pos, const void* cos_t, const void* sin_t, int T, int n_heads,
                    int head_dim, int backward, void* stream) {
  NV_REQUIRE((head_dim & 15) == 0 && (ld & 7) == 0, "nv_rope_inplace: head_dim %% 16 and ld %% 8 required");
  if (T == 0) return NV_OK;
  const int64_t work = (int64_t)T * n_heads * (head_dim >> 4);
  rope_kernel<<<grid_for(work, 256), 256, 0, S_(stream)>>>(BF(qkv), ld, pos, CBF(cos_t), CBF(sin_t), T, n_heads,
                                                         head_dim, backward ? -1.f : 1.f);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_swiglu_fwd(const void* gu, int64_t ldgu, void* h, int64_t ldh, int T, int F, void* stream) {
  NV_REQUIRE((F & 7) == 0 && (ldgu & 7) == 0 && (ldh & 7) == 0, "nv_swiglu_fwd: alignment");
  if (T == 0) return NV_OK;
  swiglu_fwd_kernel<<<grid_for((int64_t)T * (F >> 3), 256), 256, 0, S_(stream)>>>(CBF(gu), ldgu, BF(h), ldh, T, F);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_scale_bf16(void* x, int64_t ld, int rows, int cols, const float* scale, void* stream) {
  NV_REQUIRE(x && scale && (ld & 1) == 0 && rows >= 0 && cols >= 0, "nv_scale_bf16: null pointer or odd leading dimension");
  if (rows == 0 || cols == 0) return NV_OK;
  scale_bf16_kernel<<<grid_for((int64_t)rows * ((cols + 1) >> 1), 256), 256, 0, S_(stream)>>>(BF(x), ld, rows, cols, scale);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_swiglu_bwd(const void* gu, int64_t ldgu, const void* dh, int64_t lddh, void* dgu, int64_t lddgu, int T, int F,
                  void* stream) {
  NV_REQUIRE((F & 7) == 0 && (ldgu & 7) == 0 && (lddh & 7) == 0 && (lddgu & 7) == 0, "nv_swiglu_bwd: alignment");
  if (T == 0) return NV_OK;
  swiglu_bwd_kernel<<<grid_for((int64_t)T * (F >> 3), 256), 256, 0, S_(stream)>>>(CBF(gu), ldgu, CBF(dh), lddh, BF(dgu),
                                                                                lddgu, T, F);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_embed_fwd(const int* ids, const void* E, int V, const int* vis_src, const float* vis, void* out, int T, int D,
                 void* stream) {
  NV_REQUIRE((D & 7) == 0, "nv_embed_fwd: D %% 8");
  if (T == 0) return NV_OK;
  NV_CUDA(launch_pdl(embed_fwd_kernel, dim3(grid_for((int64_t)T * (D >> 3), 256)), dim3(256), 0, S_(stream), ids, CBF(E), V, vis_src, vis, BF(out), T, D));
  return NV_OK;
}

int nv_embed_bwd_vis(const void* dx, const int* vis_src, float* dvis, int T, int D, void* stream) {
  if (T == 0) return NV_OK;
  embed_bwd_vis_kernel<<<grid_for((int64_t)T * (D >> 3), 256), 256, 0, S_(stream)>>>(CBF(dx), vis_src, dvis, T, D);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_embed_bwd_weight(const void* dx, const int* order, const int* sorted_ids, void* dE, int T, int D, void* stream) {
  if (T == 0) return NV_OK;
  embed_bwd_weight_kernel<<<T, 128, 0, S_(stream)>>>(CBF(dx), order, sorted_ids, BF(dE), T, D);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_gather_rows(const void* src, int64_t lds, const int* rows, void* dst, int64_t ldd, int R, int D, void* stream) {
  if (R == 0) return NV_OK;
  gather_rows_kernel<<<grid_for((int64_t)R * (D >> 3), 256), 256, 0, S_(stream)>>>(CBF(src), lds, rows, BF(dst), ldd, R, D);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_scatter_rows(const void* src, int64_t lds, const int* rows, void* dst, int64_t ldd, int R, int D, void* stream) {
  if (R == 0) return NV_OK;
  scatter_rows_kernel<<<grid_for((int64_t)R * (D >> 3), 256), 256, 0, S_(stream)>>>(CBF(src), lds, rows, BF(dst), ldd, R, D);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_head_fwd(const void* x, int64_t ldx, const void* W, const void* bias, void* out, int R, int O, int D,
                void* stream) {
  NV_REQUIRE((D & 7) == 0 && (ldx & 7) == 0, "nv_head_fwd: alignment");
  if (R == 0) return NV_OK;
  const int64_t threads = (int64_t)R * O * 32;
  head_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, S_(stream)>>>(CBF(x), ldx, CBF(W), CBF(bias), BF(out), R, O, D);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_head_bwd(const void* dy, const void* x, int64_t ldx, const void* W, void* dx, int64_t lddx, void* dW, void* db,
                int R, int O, int D, void* stream) {
  if (R == 0) return NV_OK;
  if (dx) {
    head_bwd_dx_kernel<<<(unsigned)(((int64_t)R * D + 255) / 256), 256, 0, S_(stream)>>>(CBF(dy), CBF(W), BF(dx), lddx, R, O, D);
    NV_LAUNCH_CHECK();
  }
  if (dW) {
    head_bwd_dw_kernel<<<(unsigned)(((int64_t)O * D + 255) / 256), 256, 0, S_(stream)>>>(CBF(dy), CBF(x), ldx, BF(dW), BF(db), R, O, D);
    NV_LAUNCH_CHECK();
  }
  return NV_OK;
}

int nv_ce_fwd_bwd(const void* logits, int64_t ld, const int* labels, const int* special, int n_special,
                  float* row_loss, void* dlogits, int64_t ldd, int N, int V, float grad_scale, void* stream) {
  if (N == 0) return NV_OK;
  NV_REQUIRE(n_special >= 0 && n_special <= 16, "nv_ce_fwd_bwd: n_special out of range");
  ce_kernel<<<N, 256, 0, S_(stream)>>>(CBF(logits), ld, labels, special, n_special, row_loss, BF(dlogits), ldd, V, grad_scale);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

int nv_logit_scatter_fwd(const void* pred, int O, const int* slot, void* out, int B, int G, void* stream) {
  if (B * G == 0) return NV_OK;
  logit_scatter_fwd_kernel<<<(B * G + 127) / 128, 128, 0, S_(stream)>>>(CBF(pred), O, slot, BF(out), B, G);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

// dpred must be zero-initialised by the caller
int nv_logit_scatter_bwd(const void* dout, const int* slot, void* dpred, int O, int B, int G, void* stream) {
  if (B * G == 0) return NV_OK;
  logit_scatter_bwd_kernel<<<(B * G + 127) / 128, 128, 0, S_(stream)>>>(CBF(dout), slot, BF(dpred), O, B, G);
  NV_LAUNCH_CHECK();
  return NV_OK;
}

}  // extern "C"
